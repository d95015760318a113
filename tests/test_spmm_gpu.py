"""GPU suite: SpMM (K right-hand sides at once) against K oracle SpMVs, column by column, to the
same per-row tolerance as the single-vector path."""
import numpy as np
import pytest

from oracle import cpu, generators as g

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
TOL = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-13}

FAMILIES = {
    "uniform_16": lambda: g.uniform_rows(20000, 20000, 16, 7),
    "lap2d": lambda: g.lap2d(120),
    "rmat_s15_hubs": lambda: g.rmat(15, 16, 5),                                   # hub rows: warp + CTA tiers
    "ragged_huge_rows": lambda: g.ragged(3000, 4000, 6.0, 1, heavy_rows=3, heavy_len=40000),
    "ragged_f64": lambda: g.ragged(4000, 1500, 20.0, 3, dtype=np.float64, heavy_len=20000),
    "rmat_o64": lambda: g.rmat(13, 16, 8, offset_dtype=np.int64),
    "longrow_f64": lambda: g.uniform_rows(128, 8192, 2048, 9, dtype=np.float64),
}


@pytest.fixture(scope="module", autouse=True)
def _need_cuda(built_lib):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def make_X(n_cols, k, dtype):
    return np.stack([g.gen_x(100 + j, n_cols, dtype) for j in range(k)], axis=1).copy()


def check(Y, Ap, Aj, Ax, X, alpha=1.0):
    tol = TOL[np.dtype(Ax.dtype)]
    for j in range(X.shape[1]):
        xj = np.ascontiguousarray(X[:, j])
        y64 = alpha * cpu.spmv_fp64(Ap, Aj, Ax, xj)
        scale = abs(alpha) * cpu.abs_scale(Ap, Aj, Ax, xj)
        err = np.abs(Y[:, j].astype(np.float64) - y64)
        bad = np.nonzero(~(err <= tol * scale))[0]
        assert bad.size == 0, (j, bad[:5], err[bad[:5]], tol * scale[bad[:5]])


ROUTES = {"auto": {}, "vector": {"spmm_force_vector": 1}, "merge": {"spmm_force_merge": 1},
          "columns": {"spmm_by_columns": 1}}


def run_route(route, fn):
    from spmv_samples_b200 import spmv
    for name, v in ROUTES[route].items():
        spmv.set_option(name, v)
    try:
        fn()
    finally:
        for name in ROUTES[route]:
            spmv.set_option(name, 0)
    torch.cuda.synchronize()


@pytest.mark.parametrize("route", sorted(ROUTES))
@pytest.mark.parametrize("k", [2, 4, 8])
@pytest.mark.parametrize("family", sorted(FAMILIES))
def test_spmm_matches_k_spmvs(family, k, route):
    """auto: the selector's route (power-law matrices take the merge-path SpMM tile kernel, the
    rest the row kernel); vector / merge: that kernel on every matrix, hub rows included;
    columns: the ablation that runs K merge-path SpMVs on merge-class matrices."""
    from spmv_samples_b200 import spmv
    Ap, Aj, Ax = FAMILIES[family]()
    n_rows, n_cols = Ap.shape[0] - 1, int(Aj.max()) + 1
    X = make_X(n_cols, k, Ax.dtype)
    Y = torch.full((n_rows, k), float("nan"), dtype=dev(Ax).dtype, device="cuda")
    run_route(route, lambda: spmv.spmm(dev(Ap), dev(Aj), dev(Ax), dev(X), Y))
    check(Y.cpu().numpy(), Ap, Aj, Ax, X)


def _csr(lens, n_cols, seed=0, dtype=np.float32, off=np.int32):
    rng = np.random.default_rng(seed)
    Ap = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=Ap[1:])
    nnz = int(Ap[-1])
    return Ap.astype(off), rng.integers(0, n_cols, nnz).astype(np.int32), rng.uniform(-1, 1, nnz).astype(dtype)


MERGE_EDGE = {
    "all_rows_empty": lambda: _csr([0] * 777, 10),
    "single_row_single_nnz": lambda: _csr([1], 5),
    "single_row_100k": lambda: _csr([100003], 4096),
    "hub_between_empties": lambda: _csr([0] * 100 + [50001] + [0] * 100, 999),
    "all_nnz_in_last_row": lambda: _csr([0] * 5000 + [7777], 100),
    "all_nnz_in_first_row": lambda: _csr([7777] + [0] * 5000, 100),
    "nnz_not_multiple_of_4": lambda: _csr([3, 1, 2, 5, 0, 7, 1], 9),
    "exact_tile_multiple": lambda: _csr([3] * 127, 64),            # 127 rows + 381 nnz = 508 path items
    "tile_boundary_on_row_end": lambda: _csr([507, 507], 64),
    "exact_tile_multiple_1020": lambda: _csr([9] * 102, 64),       # 102 rows + 918 nnz = 1020 (8 items/thread)
    "tile_boundary_on_row_end_1020": lambda: _csr([1019, 1019, 3], 64),
    "n_cols_1": lambda: _csr([1, 0, 1, 1, 0] * 50, 1),
    "long_rows_f64": lambda: _csr([4099, 1, 4097, 0, 8191], 512, dtype=np.float64),
    "o64_mixed": lambda: _csr([5, 0, 300, 2, 2, 9000, 1], 700, off=np.int64),
}


@pytest.mark.parametrize("k", [2, 4, 8])
@pytest.mark.parametrize("case", sorted(MERGE_EDGE))
def test_merge_spmm_edge_cases_strides_and_alpha(case, k):
    """The merge-path SpMM tile kernel on the edge cases of the SpMV suite, with leading
    dimensions larger than k, alpha, and sentinels around the k columns of Y."""
    from spmv_samples_b200 import spmv
    Ap, Aj, Ax = MERGE_EDGE[case]()
    n_rows = Ap.shape[0] - 1
    n_cols = {"n_cols_1": 1}.get(case, int(Aj.max()) + 1 if Aj.size else 10)
    X = make_X(n_cols, k, Ax.dtype)
    tdt = dev(Ax).dtype
    Xd = torch.full((n_cols, 16), float("nan"), dtype=tdt, device="cuda")
    Xd[:, :k] = dev(X)
    Yd = torch.full((n_rows, 24), 777.0, dtype=tdt, device="cuda")
    Yd[:, :k] = float("nan")
    alpha = torch.tensor([1.5], dtype=tdt, device="cuda")
    run_route("merge", lambda: spmv.spmm(dev(Ap), dev(Aj), dev(Ax), Xd[:, :k], Yd[:, :k], alpha_dev=alpha))
    check(Yd[:, :k].cpu().numpy(), Ap, Aj, Ax, X, alpha=1.5)
    assert bool((Yd[:, k:] == 777.0).all())              # nothing written outside the k columns
    if case == "all_rows_empty":
        assert bool((Yd[:, :k] == 0).all())


@pytest.mark.parametrize("width", [1, 2, 8, 32])
def test_spmm_every_width_and_alpha_and_strides(width):
    from spmv_samples_b200 import spmv
    Ap, Aj, Ax = g.ragged(5000, 2000, 13.0, 4, heavy_len=30000)
    X = make_X(2000, 4, np.float32)
    Xd = torch.zeros(2000, 12, device="cuda")          # leading dimension 12 > k
    Xd[:, :4] = dev(X)
    Yd = torch.full((5000, 8), float("nan"), device="cuda")
    alpha = torch.tensor([-2.5], device="cuda")
    spmv.set_option("vector_width", width)
    spmv.set_option("spmm_force_vector", 1)
    try:
        spmv.spmm(dev(Ap), dev(Aj), dev(Ax), Xd[:, :4], Yd[:, :4], alpha_dev=alpha)
    finally:
        spmv.set_option("vector_width", 0)
        spmv.set_option("spmm_force_vector", 0)
    torch.cuda.synchronize()
    check(Yd[:, :4].cpu().numpy(), Ap, Aj, Ax, X, alpha=-2.5)
    assert bool(torch.isnan(Yd[:, 4:]).all())           # nothing written outside the k columns


def test_spmm_rejects_bad_k_and_misalignment():
    from spmv_samples_b200 import spmv, _lib
    Ap, Aj, Ax = g.uniform_rows(64, 64, 16, 1)
    d = [dev(a) for a in (Ap, Aj, Ax)]
    with pytest.raises(_lib.SpmvB200Error) as ei:
        spmv.spmm(*d, torch.zeros(64, 3, device="cuda"), torch.zeros(64, 3, device="cuda"))
    assert ei.value.status == 4
    Xbig = torch.zeros(64, 5, device="cuda")
    with pytest.raises(_lib.SpmvB200Error) as ei:      # rows of X start at multiples of 20 bytes
        spmv.spmm(*d, Xbig[:, :4], torch.zeros(64, 4, device="cuda"))
    assert ei.value.status == 2


def test_vector_kernel_on_hub_rows():
    """CSR-vector on a matrix with hub rows: they go through the whole-warp tier."""
    from spmv_samples_b200 import spmv
    Ap, Aj, Ax = g.ragged(6000, 9000, 5.0, 21, heavy_rows=12, heavy_len=70000)
    x = g.gen_x(3, 9000)
    for width in (1, 4, 32):
        spmv.set_option("vector_width", width)
        try:
            y = torch.full((6000,), float("nan"), device="cuda")
            spmv.SpMV("vector", 6000, 9000, int(Ap[-1]), dev(Ap), dev(Aj), dev(Ax), dev(x), y)
        finally:
            spmv.set_option("vector_width", 0)
        torch.cuda.synchronize()
        err = np.abs(y.cpu().numpy().astype(np.float64) - cpu.spmv_fp64(Ap, Aj, Ax, x))
        assert np.all(err <= 1e-5 * cpu.abs_scale(Ap, Aj, Ax, x))
