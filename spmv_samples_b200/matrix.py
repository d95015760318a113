"""Host-buffer front end: a CSR matrix uploaded once, then y = A @ x with x, y in host memory.

This is the end-to-end call a user of the reference's driver makes (reference/main.cu:55-97:
upload the CSR arrays once, then SpMV + copy y back per call), wrapped over
spmvb200_matrix_{create,spmv_host,destroy}.  The H2D copy of x, the kernel and the D2H copy
of y all happen inside `spmv`; it returns after the stream has been synchronised.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .spmv import KIND_IDS, SpMVKindError


MAX_SLOTS = 4  # SPMVB200_MAX_SLOTS in include/spmv_b200.h


class CsrMatrix:
    def __init__(self, n_rows: int, n_cols: int, Ap: np.ndarray, Aj: np.ndarray, Ax: np.ndarray):
        if Ap.dtype not in (np.int32, np.int64):
            raise TypeError("Ap must be int32 or int64")
        if Aj.dtype != np.int32:
            raise TypeError("Aj must be int32")
        if Ax.dtype not in (np.float32, np.float64):
            raise TypeError("Ax must be float32 or float64")
        if Ap.shape[0] != n_rows + 1:
            raise ValueError("Ap must have n_rows + 1 entries")
        self.n_rows, self.n_cols = int(n_rows), int(n_cols)
        self.nnz = int(Ap[-1])
        self.dtype = Ax.dtype
        Ap, Aj, Ax = map(np.ascontiguousarray, (Ap, Aj, Ax))
        h = C.c_void_p()
        st = _lib.lib().spmvb200_matrix_create(
            Ap.dtype.itemsize * 8, Ax.dtype.itemsize * 8, self.n_rows, self.n_cols, self.nnz,
            Ap.ctypes.data, Aj.ctypes.data if self.nnz else None,
            Ax.ctypes.data if self.nnz else None, C.byref(h))
        _lib.check(st, "spmvb200_matrix_create")
        self._h = h

    @classmethod
    def from_device(cls, csr) -> "CsrMatrix":
        """Wrap CSR arrays that already live on the device (a generate.Csr); they are borrowed,
        not copied, and must outlive the object."""
        self = cls.__new__(cls)
        self.n_rows, self.n_cols, self.nnz = int(csr.n_rows), int(csr.n_cols), int(csr.nnz)
        self.dtype = np.dtype(np.float32 if csr.Ax.element_size() == 4 else np.float64)
        self._keep = csr
        h = C.c_void_p()
        st = _lib.lib().spmvb200_matrix_create_from_device(
            csr.Ap.element_size() * 8, csr.Ax.element_size() * 8, self.n_rows, self.n_cols,
            self.nnz, csr.Ap.data_ptr(), csr.Aj.data_ptr(), csr.Ax.data_ptr(), C.byref(h))
        _lib.check(st, "spmvb200_matrix_create_from_device")
        self._h = h
        return self

    def spmv(self, x: np.ndarray, y: np.ndarray | None = None, kind: str = "auto") -> np.ndarray:
        """y = A @ x, host in / host out (pinned or pageable)."""
        if kind not in KIND_IDS:
            raise SpMVKindError(f'SpMV kind "{kind}" is NOT SUPPROT')
        if x.dtype != self.dtype or x.shape[0] != self.n_cols or not x.flags.c_contiguous:
            raise ValueError("x must be a contiguous vector of n_cols values of the matrix dtype")
        if y is None:
            y = np.empty(self.n_rows, dtype=self.dtype)
        st = _lib.lib().spmvb200_matrix_spmv_host(self._h, KIND_IDS[kind], x.ctypes.data,
                                                  y.ctypes.data)
        _lib.check(st, "spmvb200_matrix_spmv_host")
        return y

    def submit(self, slot: int, x: np.ndarray, y: np.ndarray, kind: str = "auto") -> None:
        """Asynchronous y = A @ x on pipeline slot 0..MAX_SLOTS-1 (own stream, own device buffers): the
        upload, the kernel and the download are enqueued and the call returns.  x and y should be
        pinned host arrays and must not be touched until wait(slot)."""
        if kind not in KIND_IDS:
            raise SpMVKindError(f'SpMV kind "{kind}" is NOT SUPPROT')
        if x.dtype != self.dtype or x.shape[0] != self.n_cols or y.dtype != self.dtype \
                or y.shape[0] != self.n_rows or not x.flags.c_contiguous or not y.flags.c_contiguous:
            raise ValueError("x / y must be contiguous vectors of the matrix dtype and shape")
        st = _lib.lib().spmvb200_matrix_submit_host(self._h, KIND_IDS[kind], int(slot), x.ctypes.data,
                                                    y.ctypes.data)
        _lib.check(st, "spmvb200_matrix_submit_host")

    def wait(self, slot: int) -> None:
        _lib.check(_lib.lib().spmvb200_matrix_wait(self._h, int(slot)), "spmvb200_matrix_wait")

    def spmv_many(self, xs, ys, kind: str = "auto", slots: int = MAX_SLOTS) -> None:
        """ys[i] = A @ xs[i] for a sequence of independent right-hand sides, `slots` in flight
        (3: one uploading, one in the kernel, one downloading)."""
        if not 1 <= slots <= MAX_SLOTS:
            raise ValueError(f"slots must be in 1..{MAX_SLOTS}")
        n = 0
        for i, (x, y) in enumerate(zip(xs, ys)):
            if i >= slots:
                self.wait(i % slots)
            self.submit(i % slots, x, y, kind)
            n = i + 1
        for s in range(min(n, slots)):
            self.wait(s)

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().spmvb200_matrix_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
