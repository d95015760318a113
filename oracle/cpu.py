"""ctypes bindings of the CPU checkers.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

``liboracle.so``  -- C restatement of reference/include/spmv/cpu_navie.hpp:3-35,
                     merge_based/thread_search.cuh:16-49 and load.hpp:420-474.
``_ref/libspmv_ref.so`` -- those reference files themselves, compiled unmodified.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "liboracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libspmv_ref.so")

_lib = None
_ref = None


def build(force: bool = False) -> None:
    """Run oracle/Makefile (C restatement always; _ref only where /root/reference exists)."""
    if force or not os.path.exists(_ORACLE_SO) or (
        os.path.isdir("/root/reference") and not os.path.exists(_REF_SO)
    ):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_ORACLE_SO)
    return _lib


def have_ref() -> bool:
    return os.path.exists(_REF_SO)


def ref() -> C.CDLL:
    global _ref
    if _ref is None:
        if not have_ref():
            raise RuntimeError("oracle/_ref/libspmv_ref.so is not built (needs /root/reference)")
        _ref = C.CDLL(_REF_SO)
    return _ref


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _off_tag(Ap: np.ndarray) -> str:
    if Ap.dtype == np.int32:
        return "o32"
    if Ap.dtype == np.int64:
        return "o64"
    raise TypeError(f"row offsets must be int32 or int64, got {Ap.dtype}")


def _val_tag(Ax: np.ndarray) -> str:
    if Ax.dtype == np.float32:
        return "f32"
    if Ax.dtype == np.float64:
        return "f64"
    raise TypeError(f"values must be float32 or float64, got {Ax.dtype}")


def _check(Ap, Aj, Ax):
    assert Ap.flags.c_contiguous and Aj.flags.c_contiguous and Ax.flags.c_contiguous
    assert Aj.dtype == np.int32, "column indices are int32 in every configuration"


# --------------------------------------------------------------------------- SpMV
def spmv(Ap, Aj, Ax, x) -> np.ndarray:
    """y = A x with the accumulator in the value type (cpu_navie.hpp:3-17)."""
    _check(Ap, Aj, Ax)
    n_rows = Ap.shape[0] - 1
    x = np.ascontiguousarray(x, dtype=Ax.dtype)
    y = np.empty(n_rows, dtype=Ax.dtype)
    fn = getattr(lib(), f"oracle_spmv_{_off_tag(Ap)}_{_val_tag(Ax)}")
    fn(C.c_int64(n_rows), _p(Ap), _p(Aj), _p(Ax), _p(x), _p(y))
    return y


def spmv_fp64(Ap, Aj, Ax, x) -> np.ndarray:
    """fp64 host reference: products and sums in double
    (SpMV_cpu_navie<int, off, float|double, double, double>)."""
    _check(Ap, Aj, Ax)
    n_rows = Ap.shape[0] - 1
    xd = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty(n_rows, dtype=np.float64)
    if Ax.dtype == np.float64:
        fn = getattr(lib(), f"oracle_spmv_{_off_tag(Ap)}_f64")
    else:
        fn = getattr(lib(), f"oracle_spmv_{_off_tag(Ap)}_f32_acc64")
    fn(C.c_int64(n_rows), _p(Ap), _p(Aj), _p(Ax), _p(xd), _p(y))
    return y


def abs_scale(Ap, Aj, Ax, x) -> np.ndarray:
    """Per-row tolerance scale sum_k |a_k x_k| in fp64 (cpu_navie.hpp:20-35)."""
    _check(Ap, Aj, Ax)
    n_rows = Ap.shape[0] - 1
    xd = np.ascontiguousarray(x, dtype=np.float64)
    s = np.empty(n_rows, dtype=np.float64)
    fn = getattr(lib(), f"oracle_abs_{_off_tag(Ap)}_{_val_tag(Ax)}")
    fn(C.c_int64(n_rows), _p(Ap), _p(Aj), _p(Ax), _p(xd), _p(s))
    return s


SEMIRINGS = {"min_plus": 1, "max_plus": 2, "or_and": 3}


def spmv_semiring(Ap, Aj, Ax, x, semiring: str) -> np.ndarray:
    """Generalised SpMV over a fixed semiring (cpu_navie.hpp:20-35, SpMV_genl_cpu_navie)."""
    _check(Ap, Aj, Ax)
    n_rows = Ap.shape[0] - 1
    x = np.ascontiguousarray(x, dtype=Ax.dtype)
    y = np.empty(n_rows, dtype=Ax.dtype)
    fn = getattr(lib(), f"oracle_genl_{_off_tag(Ap)}_{_val_tag(Ax)}")
    fn.restype = C.c_int
    rc = fn(C.c_int(SEMIRINGS[semiring]), C.c_int64(n_rows), _p(Ap), _p(Aj), _p(Ax), _p(x), _p(y))
    assert rc == 0
    return y


def ref_spmv_semiring(Ap, Aj, Ax, x, semiring: str) -> np.ndarray:
    """The reference's SpMV_genl_cpu_navie with the same functor (int32 offsets)."""
    _check(Ap, Aj, Ax)
    assert Ap.dtype == np.int32
    n_rows = Ap.shape[0] - 1
    x = np.ascontiguousarray(x, dtype=Ax.dtype)
    y = np.empty(n_rows, dtype=Ax.dtype)
    fn = getattr(ref(), f"ref_genl_o32_{_val_tag(Ax)}")
    fn.restype = C.c_int
    rc = fn(C.c_int(SEMIRINGS[semiring]), C.c_int32(n_rows), C.c_int32(x.shape[0]),
            C.c_int32(int(Ap[-1])), _p(Ap), _p(Aj), _p(Ax), _p(x), _p(y))
    assert rc == 0
    return y


def spmv_mt(Ap, Aj, Ax, x, n_threads: int):
    """Row-block-parallel run of the same loop; returns (y, threads_used)."""
    _check(Ap, Aj, Ax)
    n_rows = Ap.shape[0] - 1
    x = np.ascontiguousarray(x, dtype=Ax.dtype)
    y = np.empty(n_rows, dtype=Ax.dtype)
    fn = getattr(lib(), f"oracle_spmv_mt_{_off_tag(Ap)}_{_val_tag(Ax)}")
    fn.restype = C.c_int
    used = fn(C.c_int64(n_rows), _p(Ap), _p(Aj), _p(Ax), _p(x), _p(y), C.c_int(n_threads))
    return y, int(used)


def max_threads() -> int:
    fn = lib().oracle_max_threads
    fn.restype = C.c_int
    return int(fn())


# ---------------------------------------------------------------- merge-path search
def merge_path_search(Ap, diagonal: int):
    n_rows = Ap.shape[0] - 1
    nnz = int(Ap[-1])
    cx, cy = C.c_int64(0), C.c_int64(0)
    fn = getattr(lib(), f"oracle_merge_path_search_{_off_tag(Ap)}")
    fn(C.c_int64(diagonal), C.c_int64(n_rows), C.c_int64(nnz), _p(Ap), C.byref(cx), C.byref(cy))
    return int(cx.value), int(cy.value)


def merge_tile_coords(Ap, tile_items: int):
    """(x, y) for the diagonals t*tile_items, t = 0..tiles (dispatch_spmv_orig.cuh:109-148)."""
    n_rows = Ap.shape[0] - 1
    nnz = int(Ap[-1])
    tiles = (n_rows + nnz + tile_items - 1) // tile_items
    n = tiles + 1
    cx = np.empty(n, dtype=np.int64)
    cy = np.empty(n, dtype=np.int64)
    fn = getattr(lib(), f"oracle_merge_tile_coords_{_off_tag(Ap)}")
    fn(C.c_int64(n_rows), C.c_int64(nnz), _p(Ap), C.c_int64(tile_items), C.c_int64(n), _p(cx),
       _p(cy))
    return cx, cy


def row_split(Ap, parts: int) -> np.ndarray:
    """nnz-balanced row boundaries for `parts` shards."""
    n_rows = Ap.shape[0] - 1
    nnz = int(Ap[-1])
    out = np.empty(parts + 1, dtype=np.int64)
    fn = getattr(lib(), f"oracle_row_split_{_off_tag(Ap)}")
    fn(C.c_int64(n_rows), C.c_int64(nnz), _p(Ap), C.c_int64(parts), _p(out))
    return out


def rows_at_cost(Ap, targets, weight=(1, 1)) -> np.ndarray:
    """Weighted split search, restated with numpy: f(r) = w_den*Ap[r] + w_num*r; for each target
    the largest r in [0, n_rows] with f(r) <= target.  With weight (1, 1) and targets
    floor(g*(n_rows+nnz)/P) this is row_split above (thread_search.cuh:16-49 returns the first p
    with f(p+1) > d)."""
    f = int(weight[1]) * np.asarray(Ap, dtype=np.int64) + int(weight[0]) * np.arange(Ap.shape[0], dtype=np.int64)
    t = np.asarray(targets, dtype=np.int64)
    # f is non-decreasing; side="right" counts the entries <= target, f[0] = 0 is always one
    return np.searchsorted(f, t, side="right").astype(np.int64) - 1


# ------------------------------------------------------------------- COO -> CSR
def coo_to_csr(n_rows: int, rows, cols, vals, offset_dtype=np.int32):
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    cols = np.ascontiguousarray(cols, dtype=np.int32)
    vals = np.ascontiguousarray(vals)
    nnz = rows.shape[0]
    Ap = np.empty(n_rows + 1, dtype=offset_dtype)
    Aj = np.empty(nnz, dtype=np.int32)
    Ax = np.empty(nnz, dtype=vals.dtype)
    fn = getattr(lib(), f"oracle_coo_to_csr_{_off_tag(Ap)}_{_val_tag(vals)}")
    fn.restype = C.c_int
    rc = fn(C.c_int64(n_rows), C.c_int64(nnz), _p(rows), _p(cols), _p(vals), _p(Ap), _p(Aj),
            _p(Ax))
    if rc:
        raise MemoryError("oracle_coo_to_csr")
    return Ap, Aj, Ax


# ------------------------------------------------------- the reference itself (_ref)
def ref_spmv(Ap, Aj, Ax, x) -> np.ndarray:
    """The reference's SpMV_cpu_navie, unmodified (int32 offsets; int64 while < 2^31)."""
    _check(Ap, Aj, Ax)
    n_rows = Ap.shape[0] - 1
    nnz = int(Ap[-1])
    x = np.ascontiguousarray(x, dtype=Ax.dtype)
    y = np.empty(n_rows, dtype=Ax.dtype)
    tag = f"{_off_tag(Ap)}_{_val_tag(Ax)}"
    if tag == "o64_f64":
        raise NotImplementedError("no reference instantiation for int64 offsets + fp64")
    fn = getattr(ref(), f"ref_spmv_{tag}")
    nnz_c = C.c_int64(nnz) if Ap.dtype == np.int64 else C.c_int32(nnz)
    fn(C.c_int32(n_rows), C.c_int32(x.shape[0]), nnz_c, _p(Ap), _p(Aj), _p(Ax), _p(x), _p(y))
    return y


def ref_spmv_fp64(Ap, Aj, Ax, x) -> np.ndarray:
    _check(Ap, Aj, Ax)
    assert Ap.dtype == np.int32
    n_rows = Ap.shape[0] - 1
    xd = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty(n_rows, dtype=np.float64)
    name = "ref_spmv_o32_f64" if Ax.dtype == np.float64 else "ref_spmv_o32_f32_acc64"
    getattr(ref(), name)(C.c_int32(n_rows), C.c_int32(xd.shape[0]), C.c_int32(int(Ap[-1])),
                         _p(Ap), _p(Aj), _p(Ax), _p(xd), _p(y))
    return y


def ref_abs_scale(Ap, Aj, Ax, x) -> np.ndarray:
    _check(Ap, Aj, Ax)
    assert Ap.dtype == np.int32
    n_rows = Ap.shape[0] - 1
    xd = np.ascontiguousarray(x, dtype=np.float64)
    s = np.empty(n_rows, dtype=np.float64)
    getattr(ref(), f"ref_abs_o32_{_val_tag(Ax)}")(
        C.c_int32(n_rows), C.c_int32(xd.shape[0]), C.c_int32(int(Ap[-1])), _p(Ap), _p(Aj),
        _p(Ax), _p(xd), _p(s))
    return s


def ref_spmv_mt(Ap, Aj, Ax, x, n_threads: int):
    """Unmodified SpMV_cpu_navie called on contiguous row blocks from all host threads."""
    _check(Ap, Aj, Ax)
    n_rows = Ap.shape[0] - 1
    nnz = int(Ap[-1])
    if nnz >= 2 ** 31:
        raise ValueError("the reference's inner counter is int32 (cpu_navie.hpp:12)")
    x = np.ascontiguousarray(x, dtype=Ax.dtype)
    y = np.empty(n_rows, dtype=Ax.dtype)
    tag = f"{_off_tag(Ap)}_{_val_tag(Ax)}"
    fn = getattr(ref(), f"ref_spmv_mt_{tag}")
    fn.restype = C.c_int
    nnz_c = C.c_int64(nnz) if Ap.dtype == np.int64 else C.c_int32(nnz)
    used = fn(C.c_int32(n_rows), C.c_int32(x.shape[0]), nnz_c, _p(Ap), _p(Aj), _p(Ax), _p(x),
              _p(y), C.c_int(n_threads))
    return y, int(used)


def ref_merge_path_search(Ap, diagonal: int):
    n_rows = Ap.shape[0] - 1
    nnz = int(Ap[-1])
    if Ap.dtype == np.int32:
        cx, cy = C.c_int32(0), C.c_int32(0)
        ref().ref_merge_path_search_o32(C.c_int32(diagonal), C.c_int32(n_rows), C.c_int32(nnz),
                                        _p(Ap), C.byref(cx), C.byref(cy))
    else:
        cx, cy = C.c_int64(0), C.c_int64(0)
        ref().ref_merge_path_search_o64(C.c_int64(diagonal), C.c_int64(n_rows), C.c_int64(nnz),
                                        _p(Ap), C.byref(cx), C.byref(cy))
    return int(cx.value), int(cy.value)


def ref_load_mtx(filename: str):
    """Matrix Market file -> CSR through the reference's LoadCoo + ToCsr (int32, fp32)."""
    n_rows, n_cols, nnz = C.c_int64(0), C.c_int64(0), C.c_int64(0)
    fn = ref().ref_load_mtx_f32
    fn.restype = C.c_void_p
    h = fn(filename.encode(), C.byref(n_rows), C.byref(n_cols), C.byref(nnz))
    if not h:
        raise RuntimeError(f"reference loader rejected {filename}")
    Ap = np.empty(n_rows.value + 1, dtype=np.int32)
    Aj = np.empty(nnz.value, dtype=np.int32)
    Ax = np.empty(nnz.value, dtype=np.float32)
    ref().ref_load_mtx_f32_copy(C.c_void_p(h), _p(Ap), _p(Aj), _p(Ax))
    return int(n_rows.value), int(n_cols.value), Ap, Aj, Ax


def ref_coo_to_csr(n_rows: int, n_cols: int, rows, cols, vals):
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    cols = np.ascontiguousarray(cols, dtype=np.int32)
    vals = np.ascontiguousarray(vals, dtype=np.float32)
    nnz = rows.shape[0]
    Ap = np.empty(n_rows + 1, dtype=np.int32)
    Aj = np.empty(nnz, dtype=np.int32)
    Ax = np.empty(nnz, dtype=np.float32)
    ref().ref_coo_to_csr_o32_f32(C.c_int32(n_rows), C.c_int32(n_cols), C.c_int32(nnz), _p(rows),
                                 _p(cols), _p(vals), _p(Ap), _p(Aj), _p(Ax))
    return Ap, Aj, Ax
