"""Matrix Market loader throughput on the host: include/load.hpp of this
repository against the reference's loader (oracle/_ref, compiled from /root/reference), same file.
CPU only.

    python tools/loader_bench.py [--nnz 8000000] [--rows 1000000]
"""
import argparse
import ctypes as C
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
p = argparse.ArgumentParser()
p.add_argument("--nnz", type=int, default=8_000_000)
p.add_argument("--rows", type=int, default=1_000_000)
p.add_argument("--keep", action="store_true")
a = p.parse_args()

scratch = os.path.join(ROOT, "gpurun_out")
os.makedirs(scratch, exist_ok=True)
path = os.path.join(scratch, "loader_bench.mtx")
rng = np.random.default_rng(1)
r = rng.integers(1, a.rows + 1, a.nnz)
c = rng.integers(1, a.rows + 1, a.nnz)
v = rng.uniform(-1, 1, a.nnz)
t0 = time.perf_counter()
with open(path, "w") as f:
    f.write("%%MatrixMarket matrix coordinate real general\n")
    f.write(f"{a.rows} {a.rows} {a.nnz}\n")
    np.savetxt(f, np.column_stack([r, c, v]), fmt="%d %d %.9g")
size = os.path.getsize(path)
print(f"wrote {path}: {size / 1e6:.1f} MB, {a.nnz} entries ({time.perf_counter() - t0:.1f} s)")

shim_src = os.path.join(ROOT, "tests", "cxx", "loader_shim.cpp")
shim_so = os.path.join(ROOT, "tests", "cxx", "libloader_shim.so")
gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
subprocess.run([gxx, "-std=c++17", "-O2", "-fPIC", "-shared", "-fvisibility=hidden",
                "-I" + os.path.join(ROOT, "include"), shim_src, "-o", shim_so], check=True)
shim = C.CDLL(shim_so)


def ours():
    h = C.c_void_p()
    n_rows, n_cols, nnz, scheme = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int()
    err = C.create_string_buffer(256)
    t = time.perf_counter()
    rc = shim.shim_load_o32_f32(path.encode(), C.byref(h), C.byref(n_rows), C.byref(n_cols), C.byref(nnz),
                                C.byref(scheme), err, 256)
    dt = time.perf_counter() - t
    assert rc == 0, (rc, err.value)
    Ap = np.empty(n_rows.value + 1, dtype=np.int32)
    Aj = np.empty(nnz.value, dtype=np.int32)
    Ax = np.empty(nnz.value, dtype=np.float32)
    shim.shim_copy_o32_f32(h, Ap.ctypes.data_as(C.c_void_p), Aj.ctypes.data_as(C.c_void_p),
                           Ax.ctypes.data_as(C.c_void_p))
    return dt, (Ap, Aj, Ax)


best = min(ours()[0] for _ in range(3))
_, mine = ours()
print(f"include/load.hpp (this repository): {best:.2f} s  {size / best / 1e6:.0f} MB/s  {a.nnz / best / 1e6:.1f} M entries/s")

try:
    from oracle import cpu
    t = time.perf_counter()
    ref = cpu.ref_load_mtx(path)
    dt = time.perf_counter() - t
    print(f"reference load.hpp (oracle/_ref):   {dt:.2f} s  {size / dt / 1e6:.0f} MB/s  {a.nnz / dt / 1e6:.1f} M entries/s   -> {dt / best:.1f}x")
    same = all(np.array_equal(x, y) for x, y in zip(mine, ref[-3:]))
    print("CSR arrays identical to the reference loader's:", same)
except Exception as e:  # noqa: BLE001
    print("reference loader unavailable here:", e)
if not a.keep:
    os.remove(path)
