"""Device-side synthetic matrices of BASELINE.json's five configurations.

Everything is generated on the GPU by libspmvb200.so (csrc/gen.cu) into torch-owned device
memory; oracle/generators.py is the host restatement the tests compare against.  The
reference has no generator: its driver only reads Matrix Market files
(reference/main.cu:27-28), and none are available offline.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import _lib

STREAM_VAL, STREAM_X, STREAM_COL, STREAM_RMAT = 1, 2, 3, 4
_BITS = {torch.int32: 32, torch.int64: 64, torch.float32: 32, torch.float64: 64}


@dataclass
class Csr:
    n_rows: int
    n_cols: int
    nnz: int
    Ap: torch.Tensor
    Aj: torch.Tensor
    Ax: torch.Tensor
    name: str = ""

    def algorithmic_bytes(self) -> int:
        """nnz*(idx+val) + (n_rows+1)*off + n_cols*x + n_rows*y  (BASELINE.json north_star)."""
        vb = self.Ax.element_size()
        return (self.nnz * (4 + vb) + (self.n_rows + 1) * self.Ap.element_size()
                + self.n_cols * vb + self.n_rows * vb)

    def flops(self) -> int:
        return 2 * self.nnz


# name -> description of BASELINE.json configs[0..4]
CONFIGS = {
    "c1": dict(kind="lap2d", grid=1024, dtype=torch.float32, offset=torch.int32,
               desc="5-point 2-D Laplacian 1024x1024, fp32"),
    "c2": dict(kind="uniform", n=4 * 2 ** 20, row_len=16, dtype=torch.float32, offset=torch.int32,
               desc="uniform random 4Mi x 4Mi, 16 nnz/row, fp32"),
    "c3": dict(kind="rmat", scale=24, edge_factor=16, dtype=torch.float32, offset=torch.int32,
               desc="R-MAT scale 24, edge factor 16, fp32"),
    "c4": dict(kind="uniform", n=65536, row_len=2048, dtype=torch.float64, offset=torch.int32,
               desc="long-row 65536 x 65536, 2048 nnz/row, fp64"),
    "c5": dict(kind="rmat", scale=27, edge_factor=16, dtype=torch.float32, offset=torch.int64,
               desc="R-MAT scale 27, edge factor 16, fp32, int64 offsets"),
}
DEFAULT_SEED = 0x5EED_B200


def _sp(stream=None):
    return C.c_void_p((stream or torch.cuda.current_stream()).cuda_stream)


def uniform_pm1(n: int, seed: int, stream_id: int, dtype=torch.float32, first: int = 0,
                device="cuda") -> torch.Tensor:
    out = torch.empty(n, dtype=dtype, device=device)
    with torch.cuda.device(out.device):
        st = _lib.lib().spmvb200_gen_uniform_pm1(_BITS[dtype], seed, stream_id, first, n,
                                                 out.data_ptr(), _sp())
    _lib.check(st, "gen_uniform_pm1")
    return out


def gen_x(n: int, seed: int, dtype=torch.float32, device="cuda") -> torch.Tensor:
    return uniform_pm1(n, seed, STREAM_X, dtype, 0, device)


def lap2d(grid: int, dtype=torch.float32, offset=torch.int32, device="cuda") -> Csr:
    n = grid * grid
    nnz = 5 * n - 4 * grid
    Ap = torch.empty(n + 1, dtype=offset, device=device)
    Aj = torch.empty(nnz, dtype=torch.int32, device=device)
    Ax = torch.empty(nnz, dtype=dtype, device=device)
    with torch.cuda.device(Ap.device):
        st = _lib.lib().spmvb200_gen_lap2d(_BITS[offset], _BITS[dtype], grid, Ap.data_ptr(),
                                           Aj.data_ptr(), Ax.data_ptr(), _sp())
    _lib.check(st, "gen_lap2d")
    return Csr(n, n, nnz, Ap, Aj, Ax, f"lap2d_{grid}")


def uniform_rows(n_rows: int, n_cols: int, row_len: int, seed: int, dtype=torch.float32,
                 offset=torch.int32, device="cuda") -> Csr:
    nnz = n_rows * row_len
    Ap = torch.empty(n_rows + 1, dtype=offset, device=device)
    Aj = torch.empty(nnz, dtype=torch.int32, device=device)
    Ax = torch.empty(nnz, dtype=dtype, device=device)
    with torch.cuda.device(Ap.device):
        st = _lib.lib().spmvb200_gen_uniform_rows(_BITS[offset], _BITS[dtype], n_rows, n_cols,
                                                  row_len, seed, Ap.data_ptr(), Aj.data_ptr(),
                                                  Ax.data_ptr(), _sp())
    _lib.check(st, "gen_uniform_rows")
    return Csr(n_rows, n_cols, nnz, Ap, Aj, Ax, f"uniform_{n_rows}x{row_len}")


def rmat_edges(scale: int, seed: int, first: int, count: int, device="cuda"):
    rows = torch.empty(count, dtype=torch.int32, device=device)
    cols = torch.empty(count, dtype=torch.int32, device=device)
    with torch.cuda.device(rows.device):
        st = _lib.lib().spmvb200_gen_rmat_edges(scale, seed, first, count, rows.data_ptr(),
                                                cols.data_ptr(), _sp())
    _lib.check(st, "gen_rmat_edges")
    return rows, cols


def coo_to_csr(n_rows: int, rows, cols, vals=None, offset=torch.int32, value_dtype=torch.float32):
    """Stable device COO -> CSR (clobbers rows/cols).  Returns (Ap, Aj, Ax or None)."""
    nnz = rows.numel()
    dev = rows.device
    Ap = torch.empty(n_rows + 1, dtype=offset, device=dev)
    Aj = torch.empty(nnz, dtype=torch.int32, device=dev)
    Ax = torch.empty(nnz, dtype=vals.dtype, device=dev) if vals is not None else None
    vbits = _BITS[vals.dtype] if vals is not None else _BITS[value_dtype]
    with torch.cuda.device(dev):
        st = _lib.lib().spmvb200_coo_to_csr(
            _BITS[offset], vbits, n_rows, nnz, rows.data_ptr(), cols.data_ptr(),
            vals.data_ptr() if vals is not None else None, Ap.data_ptr(), Aj.data_ptr(),
            Ax.data_ptr() if Ax is not None else None, _sp())
    _lib.check(st, "coo_to_csr")
    return Ap, Aj, Ax


def rmat(scale: int, edge_factor: int, seed: int, dtype=torch.float32, offset=torch.int32,
         device="cuda") -> Csr:
    """R-MAT CSR: edges in generation order, stable sort by row, duplicates kept, columns
    unsorted within a row (what the reference's ToCsr yields, load.hpp:457-473); the value of
    the nonzero at CSR position k is the k-th draw of the value stream."""
    n = 1 << scale
    nnz = n * edge_factor
    rows, cols = rmat_edges(scale, seed, 0, nnz, device)
    Ap, Aj, _ = coo_to_csr(n, rows, cols, None, offset, dtype)
    del rows, cols
    Ax = uniform_pm1(nnz, seed, STREAM_VAL, dtype, 0, device)
    return Csr(n, n, nnz, Ap, Aj, Ax, f"rmat_s{scale}_ef{edge_factor}")


def make_config(name: str, seed: int = DEFAULT_SEED, device="cuda", scale_override=None) -> Csr:
    """One of BASELINE.json's configs ("c1".."c5"); scale_override shrinks it for tests."""
    cfg = dict(CONFIGS[name])
    if cfg["kind"] == "lap2d":
        m = lap2d(scale_override or cfg["grid"], cfg["dtype"], cfg["offset"], device)
    elif cfg["kind"] == "uniform":
        n = scale_override or cfg["n"]
        m = uniform_rows(n, n, cfg["row_len"], seed, cfg["dtype"], cfg["offset"], device)
    else:
        m = rmat(scale_override or cfg["scale"], cfg["edge_factor"], seed, cfg["dtype"],
                 cfg["offset"], device)
    m.name = name if scale_override is None else f"{name}@{scale_override}"
    return m
