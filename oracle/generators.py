"""numpy restatement of the device matrix generators.  TEST INFRASTRUCTURE ONLY.

The reference ships no synthetic generator (SURVEY.md 2.1 "Absent"); BASELINE.json's
configs are synthetic shapes, so the generators are new.  The product generates on the
device (spmv_samples_b200/csrc/gen.cu); this file restates the same counter-based
arithmetic on the host so tests can demand bit-identical CSR arrays and feed the CPU
oracle the very same inputs.

Random numbers: splitmix64.  key = mix64(seed ^ (stream * PHI)); the ctr-th draw of a
stream is mix64(key + (ctr + 1) * PHI).  All arithmetic is modulo 2^64.
"""
from __future__ import annotations

import numpy as np

PHI = np.uint64(0x9E3779B97F4A7C15)
M1 = np.uint64(0xBF58476D1CE4E5B9)
M2 = np.uint64(0x94D049BB133111EB)

# stream ids (must match csrc/gen.cu)
STREAM_VAL = 1   # matrix values, counter = position k in Ax
STREAM_X = 2     # vector x, counter = column j
STREAM_COL = 3   # uniform-K columns, counter = row * K + k
STREAM_RMAT = 4  # R-MAT quadrant draws, counter = edge * 16 + (level // 2)

# R-MAT (a, b, c, d) = (0.57, 0.19, 0.19, 0.05) as 24-bit integer thresholds
RMAT_A = int(round(0.57 * (1 << 24)))
RMAT_AB = int(round(0.76 * (1 << 24)))
RMAT_ABC = int(round(0.95 * (1 << 24)))


def mix64(z):
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * M1
        z = (z ^ (z >> np.uint64(27))) * M2
        z = z ^ (z >> np.uint64(31))
    return z


def stream_key(seed: int, stream: int) -> np.uint64:
    with np.errstate(over="ignore"):
        return mix64(np.uint64(seed & 0xFFFFFFFFFFFFFFFF) ^ (np.uint64(stream) * PHI))[()]


def draw(key, ctr):
    ctr = np.asarray(ctr, dtype=np.uint64)
    with np.errstate(over="ignore"):
        return mix64(key + (ctr + np.uint64(1)) * PHI)


def uniform_pm1(seed: int, stream: int, first: int, count: int, dtype):
    """U(-1, 1) on a fixed grid: fp32 (h>>40 - 2^23) * 2^-23, fp64 (h>>11 - 2^52) * 2^-52."""
    h = draw(stream_key(seed, stream), np.arange(first, first + count, dtype=np.uint64))
    if np.dtype(dtype) == np.float32:
        i = (h >> np.uint64(40)).astype(np.int64) - (1 << 23)
        return (i.astype(np.float32) * np.float32(2.0 ** -23)).astype(np.float32)
    i = (h >> np.uint64(11)).astype(np.int64) - (1 << 52)
    return i.astype(np.float64) * (2.0 ** -52)


def gen_x(seed: int, n: int, dtype=np.float32):
    return uniform_pm1(seed, STREAM_X, 0, n, dtype)


# ------------------------------------------------------------------ C1: 2-D Laplacian
def lap2d(n: int, dtype=np.float32, offset_dtype=np.int32):
    """5-point stencil on an n x n grid, row-major, columns ascending, values 4 / -1."""
    N = n * n
    r = np.arange(N + 1, dtype=np.int64)
    Ap = 5 * r - np.minimum(r, n) - np.maximum(0, r - n * (n - 1)) - (r + n - 1) // n - r // n
    nnz = int(Ap[-1])
    rows = np.arange(N, dtype=np.int64)
    i, j = rows // n, rows % n
    cand = np.stack([rows - n, rows - 1, rows, rows + 1, rows + n], axis=1)
    ok = np.stack([i > 0, j > 0, np.ones(N, bool), j < n - 1, i < n - 1], axis=1)
    vals = np.broadcast_to(np.array([-1, -1, 4, -1, -1], dtype=dtype), (N, 5))
    Aj = cand[ok].astype(np.int32)
    Ax = np.ascontiguousarray(vals[ok], dtype=dtype)
    assert Aj.shape[0] == nnz
    return Ap.astype(offset_dtype), Aj, Ax


# ------------------------------------------------- C2 / C4: K distinct sorted columns
def uniform_rows(n_rows: int, n_cols: int, K: int, seed: int, dtype=np.float32,
                 offset_dtype=np.int32):
    """K columns per row, one drawn uniformly from each of K equal strata of [0, n_cols):
    distinct and ascending by construction.  Values U(-1,1)."""
    assert n_cols % K == 0
    S = n_cols // K
    key = stream_key(seed, STREAM_COL)
    ctr = np.arange(n_rows * K, dtype=np.uint64)
    h = draw(key, ctr)
    jitter = ((h >> np.uint64(32)) * np.uint64(S)) >> np.uint64(32)
    k = (ctr % np.uint64(K)).astype(np.int64)
    Aj = (k * S + jitter.astype(np.int64)).astype(np.int32)
    Ap = (np.arange(n_rows + 1, dtype=np.int64) * K).astype(offset_dtype)
    Ax = uniform_pm1(seed, STREAM_VAL, 0, n_rows * K, dtype)
    return Ap, Aj, Ax


# ---------------------------------------------------------------- C3 / C5: R-MAT
def rmat_edges(scale: int, seed: int, first: int, count: int):
    """Edge e -> (row, col): `scale` quadrant draws, most significant bit first; two
    24-bit draws per 64-bit hash (bits 63..40, then bits 39..16)."""
    assert scale <= 32
    key = stream_key(seed, STREAM_RMAT)
    e = np.arange(first, first + count, dtype=np.uint64)
    row = np.zeros(count, dtype=np.int64)
    col = np.zeros(count, dtype=np.int64)
    h = None
    for level in range(scale):
        if level % 2 == 0:
            h = draw(key, e * np.uint64(16) + np.uint64(level // 2))
            u = (h >> np.uint64(40)).astype(np.int64)
        else:
            u = ((h >> np.uint64(16)) & np.uint64(0xFFFFFF)).astype(np.int64)
        rb = (u >= RMAT_AB).astype(np.int64)
        cb = np.where(u < RMAT_A, 0, np.where(u < RMAT_AB, 1, np.where(u < RMAT_ABC, 0, 1)))
        row = (row << 1) | rb
        col = (col << 1) | cb
    return row.astype(np.int32), col.astype(np.int32)


def rmat(scale: int, edge_factor: int, seed: int, dtype=np.float32, offset_dtype=np.int32):
    """R-MAT CSR: edges in generation order, stable counting sort by row (duplicates kept,
    columns unsorted within a row -- what the reference's ToCsr produces, load.hpp:457-473).
    Values are U(-1,1) of the CSR position."""
    n = 1 << scale
    E = n * edge_factor
    rows, cols = rmat_edges(scale, seed, 0, E)
    order = np.argsort(rows, kind="stable")
    Aj = cols[order].astype(np.int32)
    counts = np.bincount(rows, minlength=n)
    Ap = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(counts, out=Ap[1:])
    Ax = uniform_pm1(seed, STREAM_VAL, 0, E, dtype)
    return Ap.astype(offset_dtype), Aj, Ax


# ------------------------------------------------------ test-only ragged matrices
def ragged(n_rows: int, n_cols: int, mean_len: float, seed: int, dtype=np.float32,
           offset_dtype=np.int32, empty_frac: float = 0.2, heavy_rows: int = 2,
           heavy_len: int = 0):
    """Random row lengths (geometric-ish, many empty rows, a few very long rows), random
    unsorted columns with duplicates.  numpy Generator -- host-only, for edge-case tests."""
    rng = np.random.default_rng(seed)
    lens = rng.geometric(1.0 / max(mean_len, 1.0), size=n_rows).astype(np.int64) - 1
    lens[rng.random(n_rows) < empty_frac] = 0
    if heavy_len and n_rows:
        idx = rng.choice(n_rows, size=min(heavy_rows, n_rows), replace=False)
        lens[idx] = heavy_len
    Ap = np.zeros(n_rows + 1, dtype=np.int64)
    np.cumsum(lens, out=Ap[1:])
    nnz = int(Ap[-1])
    Aj = rng.integers(0, max(n_cols, 1), size=nnz, dtype=np.int64).astype(np.int32)
    Ax = rng.uniform(-1, 1, size=nnz).astype(dtype)
    return Ap.astype(offset_dtype), Aj, Ax
