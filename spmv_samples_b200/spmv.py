"""Host-side mirror of the reference's registry (reference/include/spmv.h:18-48) over the C ABI.

    SpMV(kind_str, n_rows, n_cols, nnz, Ap, Aj, Ax, x, y)

has the reference's argument list and meaning: the five arrays live on the device (here:
CUDA torch tensors, used only as owners of device memory), `y` is fully overwritten with
A @ x.  `SPMV_KINDS` plays the role of the X-macro table: label -> function with the
8-argument per-kind signature (spmv.h:39).  An unknown label is an error, as in spmv.h:46-47
(there: message on stderr + exit; here: the same message in a SpMVKindError).

PyTorch is plumbing only (device memory, streams).  Every call goes through
libspmvb200.so; there is no PyTorch or CPU implementation to fall back to.
"""
from __future__ import annotations

import ctypes as C
import sys

import torch

from . import _lib


class SpMVKindError(ValueError):
    pass


_OFF = {torch.int32: ("o32", 32), torch.int64: ("o64", 64)}
_VAL = {torch.float32: ("f32", 32), torch.float64: ("f64", 64)}


def _ptr(t):
    return C.c_void_p(t.data_ptr() if t is not None else 0)


def _stream_ptr(stream=None):
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


def _check_tensors(Ap, Aj, Ax, x, y):
    for name, t in (("Ap", Ap), ("Aj", Aj), ("Ax", Ax), ("x", x), ("y", y)):
        if not t.is_cuda:
            raise ValueError(f"{name} must be a CUDA tensor (device pointer), got {t.device}")
        if not t.is_contiguous():
            raise ValueError(f"{name} must be contiguous")
    if Ap.dtype not in _OFF:
        raise TypeError(f"offset type {Ap.dtype} unsupported (int32 / int64)")
    if Aj.dtype != torch.int32:
        raise TypeError("index type must be int32")
    if Ax.dtype not in _VAL or x.dtype != Ax.dtype or y.dtype != Ax.dtype:
        raise TypeError("Ax, x, y must share one dtype, float32 or float64")


def _typed_call(kind: str, n_rows, n_cols, nnz, Ap, Aj, Ax, x, y, stream=None):
    _check_tensors(Ap, Aj, Ax, x, y)
    otag, _ = _OFF[Ap.dtype]
    vtag, _ = _VAL[Ax.dtype]
    fn = getattr(_lib.lib(), f"spmvb200_{kind}_i32_{otag}_{vtag}")
    with torch.cuda.device(Ap.device):
        st = fn(int(n_rows), int(n_cols), int(nnz), _ptr(Ap), _ptr(Aj), _ptr(Ax), _ptr(x), _ptr(y),
                _stream_ptr(stream))
    _lib.check(st, f"spmvb200_{kind}_i32_{otag}_{vtag}")


def SpMV_merge(n_rows, n_cols, nnz, Ap, Aj, Ax, x, y, stream=None):
    """Merge-path kernel (replaces SpMV_merge_based, merge_based/merge_based.cuh:22)."""
    _typed_call("merge", n_rows, n_cols, nnz, Ap, Aj, Ax, x, y, stream)


def SpMV_vector(n_rows, n_cols, nnz, Ap, Aj, Ax, x, y, stream=None):
    """CSR-vector kernel (replaces SpMV_cusp_*, cusp/cusp.cuh:227)."""
    _typed_call("vector", n_rows, n_cols, nnz, Ap, Aj, Ax, x, y, stream)


def SpMV_light(n_rows, n_cols, nnz, Ap, Aj, Ax, x, y, stream=None):
    """Dynamic-row kernel (replaces SpMV_light_vector / SpMV_light_warp, LightSpMV.cuh:379)."""
    _typed_call("light", n_rows, n_cols, nnz, Ap, Aj, Ax, x, y, stream)


def SpMV_stream(n_rows, n_cols, nnz, Ap, Aj, Ax, x, y, stream=None):
    """CSR-stream kernel: TMA-staged row tiles, thread per row (short regular rows; the
    THREADS_PER_VECTOR = 2 / 4 case of cusp/cusp.cuh:189-203)."""
    _typed_call("stream", n_rows, n_cols, nnz, Ap, Aj, Ax, x, y, stream)


def SpMV_auto(n_rows, n_cols, nnz, Ap, Aj, Ax, x, y, stream=None):
    """Host selector: row statistics -> one of the kernels above."""
    _typed_call("auto", n_rows, n_cols, nnz, Ap, Aj, Ax, x, y, stream)


def SpMV_cusparse(n_rows, n_cols, nnz, Ap, Aj, Ax, x, y, stream=None):
    """cuSPARSE baseline, setup hoisted (replaces SpMV_cusparse, cusparse.cuh:37)."""
    _typed_call("cusparse", n_rows, n_cols, nnz, Ap, Aj, Ax, x, y, stream)


# the X-macro table (spmv.h:18-27): label -> function
SPMV_KINDS = {
    "merge": SpMV_merge,
    "vector": SpMV_vector,
    "light": SpMV_light,
    "stream": SpMV_stream,
    "auto": SpMV_auto,
    "cusparse": SpMV_cusparse,
    # the reference's own labels (spmv.h:18-27), each served by the kernel that replaces it
    "cusp": SpMV_vector,
    "cusp1": SpMV_vector,
    "cusp2": SpMV_vector,
    "light_vec": SpMV_light,
    "light_warp": SpMV_light,
    "cub_merge": SpMV_merge,
}
REFERENCE_ALIASES = {"cusp": "vector", "cusp1": "vector", "cusp2": "vector", "light_vec": "light",
                     "light_warp": "light", "cub_merge": "merge"}
KIND_IDS = {"merge": _lib.KIND_MERGE, "vector": _lib.KIND_VECTOR, "light": _lib.KIND_LIGHT,
            "stream": _lib.KIND_STREAM, "auto": _lib.KIND_AUTO, "cusparse": _lib.KIND_CUSPARSE}
KIND_IDS.update({alias: KIND_IDS[target] for alias, target in REFERENCE_ALIASES.items()})


def SpMV(kind_str, n_rows, n_cols, nnz, Ap, Aj, Ax, x, y, stream=None):
    """reference/include/spmv.h:29-48."""
    fn = SPMV_KINDS.get(kind_str)
    if fn is None:
        msg = f'SpMV kind "{kind_str}" is NOT SUPPROT'  # the reference's wording, spmv.h:46
        print(msg, file=sys.stderr)
        raise SpMVKindError(msg)
    fn(n_rows, n_cols, nnz, Ap, Aj, Ax, x, y, stream)


def spmv_ex(kind_str, Ap, Aj, Ax, x, y, n_cols=None, alpha_dev=None, y_peers=(), stream=None,
            multicast=False, semiring="plus_times", beta_dev=None, static_pattern=False):
    """Untyped entry (spmvb200_spmv): optional device alpha, optional peer replicas of y.
    y_peers: iterable of raw device addresses (ints), each indexed like y; with
    multicast=True it holds ONE NVLink multicast address that reaches every replica.
    semiring: "plus_times" | "min_plus" | "max_plus" | "or_and" (merge-path kernel; the fixed
    menu standing in for the reference's functor_t, merge_genl/merge_genl.cuh:19-38);
    beta_dev: device scalar, y = alpha*A*x + beta*y (plus-times only).
    static_pattern: the caller vouches that Ap is unchanged since the previous call on this
    stream, so the merge-path tile coordinates of that call are reused (never on a first call)."""
    _check_tensors(Ap, Aj, Ax, x, y)
    if kind_str not in KIND_IDS:
        raise SpMVKindError(f'SpMV kind "{kind_str}" is NOT SUPPROT')
    a = _lib.Args()
    a.kind = KIND_IDS[kind_str]
    a.offset_bits = _OFF[Ap.dtype][1]
    a.value_bits = _VAL[Ax.dtype][1]
    a.n_rows = Ap.numel() - 1
    a.n_cols = int(n_cols) if n_cols is not None else x.numel()
    a.nnz = Aj.numel()
    a.Ap, a.Aj, a.Ax, a.x, a.y = (Ap.data_ptr(), Aj.data_ptr(), Ax.data_ptr(), x.data_ptr(),
                                  y.data_ptr())
    a.alpha_dev = alpha_dev.data_ptr() if alpha_dev is not None else None
    a.beta_dev = beta_dev.data_ptr() if beta_dev is not None else None
    a.semiring = _lib.SEMIRINGS[semiring]
    a.flags = _lib.FLAG_STATIC_PATTERN if static_pattern else 0
    peers = list(y_peers)
    if multicast and len(peers) != 1:
        raise ValueError("multicast=True takes exactly one address")
    a.n_peers = -1 if multicast else len(peers)
    arr = (C.c_void_p * max(1, len(peers)))(*peers)
    a.y_peers = C.cast(arr, C.POINTER(C.c_void_p))
    a.stream = (stream if stream is not None else torch.cuda.current_stream()).cuda_stream
    with torch.cuda.device(Ap.device):
        st = _lib.lib().spmvb200_spmv(C.byref(a))
    _lib.check(st, f"spmvb200_spmv[{kind_str}]")


def spmm(Ap, Aj, Ax, X, Y, alpha_dev=None, stream=None):
    """Y = alpha * A @ X for k = X.shape[1] in {2, 4, 8} right-hand sides at once
    (spmvb200_spmm).  X [n_cols, k] and Y [n_rows, k] are row-major CUDA tensors (a row stride
    larger than k is allowed)."""
    if X.dim() != 2 or Y.dim() != 2 or X.shape[1] != Y.shape[1] or X.stride(1) != 1 or Y.stride(1) != 1:
        raise ValueError("X and Y must be 2-D, row-major, with the same number of columns")
    if X.dtype != Ax.dtype or Y.dtype != Ax.dtype or not (X.is_cuda and Y.is_cuda):
        raise TypeError("X, Y must be CUDA tensors of the matrix dtype")
    a = _lib.SpmmArgs()
    a.offset_bits = _OFF[Ap.dtype][1]
    a.value_bits = _VAL[Ax.dtype][1]
    a.k = X.shape[1]
    a.n_rows, a.n_cols, a.nnz = Ap.numel() - 1, X.shape[0], Aj.numel()
    a.Ap, a.Aj, a.Ax = Ap.data_ptr(), Aj.data_ptr(), Ax.data_ptr()
    a.X, a.ldx, a.Y, a.ldy = X.data_ptr(), X.stride(0), Y.data_ptr(), Y.stride(0)
    a.alpha_dev = alpha_dev.data_ptr() if alpha_dev is not None else None
    a.stream = (stream if stream is not None else torch.cuda.current_stream()).cuda_stream
    if Y.shape[0] != a.n_rows:
        raise ValueError("Y must have n_rows rows")
    with torch.cuda.device(Ap.device):
        st = _lib.lib().spmvb200_spmm(C.byref(a))
    _lib.check(st, "spmvb200_spmm")


def merge_path_partition(Ap, tile_items=None, stream=None):
    """Row coordinates of the merge path on the tile diagonals (int32 CUDA tensor, tiles+1)."""
    L = _lib.lib()
    n_rows = Ap.numel() - 1
    nnz = int(Ap[-1].item())
    if tile_items is None:
        tile_items = L.spmvb200_merge_tile_items(_OFF[Ap.dtype][1], 32)
    tiles = (n_rows + nnz + tile_items - 1) // tile_items
    out = torch.empty(tiles + 1, dtype=torch.int32, device=Ap.device)
    fn = getattr(L, f"spmvb200_merge_path_partition_{_OFF[Ap.dtype][0]}")
    with torch.cuda.device(Ap.device):
        st = fn(n_rows, nnz, _ptr(Ap), int(tile_items), tiles + 1, _ptr(out), _stream_ptr(stream))
    _lib.check(st, "spmvb200_merge_path_partition")
    return out


def rows_at_cost(Ap, targets, weight=(1, 1), stream=None):
    """For each target t: the largest r with f(r) = w_den*Ap[r] + w_num*r <= t, where
    weight = (w_num, w_den) is the cost of a row in nonzeros ((1, 1): the merge path)."""
    L = _lib.lib()
    n_rows = Ap.numel() - 1
    n = len(targets)
    t = (C.c_int64 * max(n, 1))(*[int(v) for v in targets])
    out = (C.c_int64 * max(n, 1))()
    fn = getattr(L, f"spmvb200_rows_at_cost_{_OFF[Ap.dtype][0]}")
    with torch.cuda.device(Ap.device):
        st = fn(n_rows, _ptr(Ap), int(weight[0]), int(weight[1]), n, t, out, _stream_ptr(stream))
    _lib.check(st, "spmvb200_rows_at_cost")
    return [int(out[k]) for k in range(n)]


def split_targets(total_cost: int, parts: int):
    """Equal-cost targets floor(g * total / parts), g = 1 .. parts-1 (exact integers)."""
    return [(g * int(total_cost)) // parts for g in range(1, parts)]


def rebalance_targets(cost_bounds, times):
    """Targets for a re-split from measured shard times.  cost_bounds[g] = f(row_bounds[g]) of the
    current split (parts+1 values), times[g] its measured time.  Time is taken as uniform in cost
    inside a shard, and the new boundaries are put at the equal-time quantiles of that
    piecewise-linear curve.  Integer arithmetic on times quantised to 1/65536 of their sum, so
    every rank computes the same targets from the same inputs."""
    parts = len(times)
    tot = float(sum(times))
    if parts < 2 or tot <= 0:
        return [int(c) for c in cost_bounds[1:-1]]
    q = [max(1, int(round(65536.0 * float(t) / tot))) for t in times]
    qs = sum(q)
    cum = [0]
    for v in q:
        cum.append(cum[-1] + v)
    out = []
    for k in range(1, parts):
        want = k * qs  # compared against cum[g] * parts
        g = 0
        while g + 1 < parts and cum[g + 1] * parts <= want:
            g += 1
        c0, c1 = int(cost_bounds[g]), int(cost_bounds[g + 1])
        out.append(c0 + ((c1 - c0) * (want - cum[g] * parts)) // (q[g] * parts))
    return out


def row_split(Ap, parts: int, nnz=None, stream=None, weight=(1, 1)):
    """nnz-balanced row boundaries (list of parts+1 ints) from the device merge-path search.
    weight = (w_num, w_den): cost of a row relative to a nonzero; (1, 1) is the merge path of
    SURVEY.md 8(e) and goes through spmvb200_row_split_*, any other weight through
    spmvb200_rows_at_cost_* on the equal-cost targets."""
    L = _lib.lib()
    n_rows = Ap.numel() - 1
    if nnz is None:
        nnz = int(Ap[-1].item())
    if tuple(weight) != (1, 1):
        total = int(weight[1]) * int(nnz) + int(weight[0]) * n_rows
        return [0] + rows_at_cost(Ap, split_targets(total, parts), weight, stream) + [n_rows]
    out = (C.c_int64 * (parts + 1))()
    fn = getattr(L, f"spmvb200_row_split_{_OFF[Ap.dtype][0]}")
    with torch.cuda.device(Ap.device):
        st = fn(n_rows, nnz, _ptr(Ap), parts, out, _stream_ptr(stream))
    _lib.check(st, "spmvb200_row_split")
    return [int(v) for v in out]


def row_stats(Ap, nnz=None, stream=None):
    L = _lib.lib()
    n_rows = Ap.numel() - 1
    if nnz is None:
        nnz = int(Ap[-1].item())
    st_out = _lib.RowStats()
    with torch.cuda.device(Ap.device):
        st = L.spmvb200_row_stats(_OFF[Ap.dtype][1], n_rows, nnz, _ptr(Ap), C.byref(st_out),
                                  _stream_ptr(stream))
    _lib.check(st, "spmvb200_row_stats")
    return {f: getattr(st_out, f) for f, _ in _lib.RowStats._fields_}


def set_option(name: str, value: int) -> None:
    _lib.check(_lib.lib().spmvb200_set_option(name.encode(), int(value)), f"set_option({name})")


def get_option(name: str) -> int:
    return int(_lib.lib().spmvb200_get_option(name.encode()))


def launch_count() -> int:
    return int(_lib.lib().spmvb200_launch_count())


def hot_x_info(Aj) -> dict:
    """What the merge-path kernel's hot-x plan holds for this Aj on its device (csrc/hotx.cu):
    hot_columns = 0 when no plan exists."""
    k, share, ms = C.c_int64(0), C.c_double(0.0), C.c_double(0.0)
    with torch.cuda.device(Aj.device):
        st = _lib.lib().spmvb200_hot_x_info(_ptr(Aj), C.byref(k), C.byref(share), C.byref(ms))
        _lib.check(st, "spmvb200_hot_x_info")
        tk, tshare = C.c_int64(0), C.c_double(0.0)
        st = _lib.lib().spmvb200_hot_x_table_info(_ptr(Aj), C.byref(tk), C.byref(tshare))
    _lib.check(st, "spmvb200_hot_x_table_info")
    return {"hot_columns": int(k.value), "hot_share": float(share.value), "build_ms": float(ms.value),
            "table_columns": int(tk.value), "table_share": float(tshare.value)}


def release_cache() -> None:
    _lib.lib().spmvb200_release_cache()


def gather_yardstick(x_elements: int, gathers: int = 1 << 27, reps: int = 3, stream=None) -> float:
    """G gathers/s of nothing but uniformly random 4-byte gathers over x_elements floats
    (csrc/diag.cu): the yardstick beside the gather-bound configurations."""
    ms = C.c_double(0.0)
    st = _lib.lib().spmvb200_gather_yardstick(int(x_elements), int(gathers), int(reps),
                                              _stream_ptr(stream), C.byref(ms))
    _lib.check(st, "spmvb200_gather_yardstick")
    return gathers / (ms.value * 1e-3) / 1e9
