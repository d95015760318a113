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


@pytest.mark.parametrize("force_vector", [0, 1])
@pytest.mark.parametrize("k", [2, 4, 8])
@pytest.mark.parametrize("family", sorted(FAMILIES))
def test_spmm_matches_k_spmvs(family, k, force_vector):
    """force_vector=0: the selector's route (power-law matrices go column by column through
    merge-path); force_vector=1: the multi-vector row kernel on every matrix, hub rows included."""
    from spmv_samples_b200 import spmv
    Ap, Aj, Ax = FAMILIES[family]()
    n_rows, n_cols = Ap.shape[0] - 1, int(Aj.max()) + 1
    X = make_X(n_cols, k, Ax.dtype)
    Y = torch.full((n_rows, k), float("nan"), dtype=dev(Ax).dtype, device="cuda")
    spmv.set_option("spmm_force_vector", force_vector)
    try:
        spmv.spmm(dev(Ap), dev(Aj), dev(Ax), dev(X), Y)
    finally:
        spmv.set_option("spmm_force_vector", 0)
    torch.cuda.synchronize()
    check(Y.cpu().numpy(), Ap, Aj, Ax, X)


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
