"""Generate the golden input/output vectors under tests/golden/ FROM THE REFERENCE ITSELF.

Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

Every output array in the .npz files is produced by oracle/_ref/libspmv_ref.so, i.e. by the
reference's own SpMV_cpu_navie / SpMV_genl_cpu_navie / SearchMergePath / LoadCoo+ToCsr
compiled unmodified by oracle/Makefile.  The inputs come from oracle/generators.py (seeded).
The fixtures let the oracle be pinned on machines where /root/reference does not exist.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import cpu, generators as g  # noqa: E402

TILES = (2048, 896, 320)  # ours, the reference's fp32 tile, the reference's fp64 tile


def lattice():
    # reference/include/spmv/merge_based/device_spmv.cuh:95-128
    Ap = np.array([0, 2, 5, 7, 10, 14, 17, 19, 22, 24], dtype=np.int32)
    Aj = np.array([1, 3, 0, 2, 4, 1, 5, 0, 4, 6, 1, 3, 5, 7, 2, 4, 8, 3, 7, 4, 6, 8, 5, 7],
                  dtype=np.int32)
    return Ap, Aj, np.ones(24, dtype=np.float32)


CASES = {
    "lattice3x3": (lattice, 9, "ones"),
    "lap2d_16": (lambda: g.lap2d(16), 256, 101),
    "uniform_256x16": (lambda: g.uniform_rows(256, 256, 16, 7), 256, 102),
    "rmat_s8": (lambda: g.rmat(8, 16, 5), 256, 103),
    "ragged_300": (lambda: g.ragged(300, 200, 5.0, 3, heavy_len=5000), 200, 104),
    "ragged_empty_tail": (lambda: g.ragged(64, 32, 2.0, 9, empty_frac=0.7), 32, 105),
    "uniform_f64_64x128": (lambda: g.uniform_rows(64, 128, 128, 11, dtype=np.float64), 128, 106),
    "ragged_f64_200": (lambda: g.ragged(200, 150, 9.0, 13, dtype=np.float64, heavy_len=3000),
                       150, 107),
}


def main():
    assert cpu.have_ref(), "build oracle/_ref first: make -C oracle"
    for name, (gen, n_cols, xseed) in CASES.items():
        Ap, Aj, Ax = gen()
        n_rows = Ap.shape[0] - 1
        nnz = int(Ap[-1])
        if xseed == "ones":
            x = np.ones(n_cols, dtype=Ax.dtype)  # the reference driver's x (main.cu:41)
        else:
            x = g.gen_x(xseed, n_cols, Ax.dtype)
        out = dict(Ap=Ap, Aj=Aj, Ax=Ax, x=x, n_cols=np.int64(n_cols))
        out["y_ref"] = cpu.ref_spmv(Ap, Aj, Ax, x)
        out["y_ref64"] = cpu.ref_spmv_fp64(Ap, Aj, Ax, x)
        out["abs_ref"] = cpu.ref_abs_scale(Ap, Aj, Ax, x)
        for sr in cpu.SEMIRINGS:   # SpMV_genl_cpu_navie with the fixed functor menu
            out[f"y_{sr}"] = cpu.ref_spmv_semiring(Ap, Aj, Ax, x, sr)
        for tile in TILES:
            tiles = (n_rows + nnz + tile - 1) // tile
            diags = np.minimum(np.arange(tiles + 1, dtype=np.int64) * tile, n_rows + nnz)
            xy = np.array([cpu.ref_merge_path_search(Ap, int(d)) for d in diags], dtype=np.int64)
            out[f"coords_{tile}"] = xy
        # every diagonal of the small cases: the full path
        if n_rows + nnz <= 4096:
            out["path_all"] = np.array(
                [cpu.ref_merge_path_search(Ap, d) for d in range(n_rows + nnz + 1)], dtype=np.int64)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(f"{name}: rows={n_rows} nnz={nnz}")

    # Matrix Market fixtures through the reference loader (LoadCoo + ToCsr)
    mtx = {
        "general_real.mtx": "%%MatrixMarket matrix coordinate real general\n% comment\n4 5 6\n"
                            "1 1 1.5\n3 2 -2\n1 5 0.25\n4 4 3\n3 1 7\n1 1 0.5\n",
        "symmetric_pattern.mtx": "%%MatrixMarket matrix coordinate pattern symmetric\n5 5 5\n"
                                 "2 1\n3 1\n3 3\n5 2\n5 4\n",
        "symmetric_integer.mtx": "%%MatrixMarket matrix coordinate integer symmetric\n3 3 4\n"
                                 "1 1 2\n2 1 -1\n3 2 -1\n3 3 2\n",
    }
    for fname, text in mtx.items():
        path = os.path.join(HERE, fname)
        with open(path, "w") as f:
            f.write(text)
        n_rows, n_cols, Ap, Aj, Ax = cpu.ref_load_mtx(path)
        np.savez_compressed(path + ".npz", n_rows=np.int64(n_rows), n_cols=np.int64(n_cols),
                            Ap=Ap, Aj=Aj, Ax=Ax)
        print(f"{fname}: {n_rows}x{n_cols} nnz={int(Ap[-1])}")


if __name__ == "__main__":
    main()
