"""Does write-combined pinned memory for the upload buffer change what the host link gives when
upload and download run at once?  (cudaHostAlloc default vs cudaHostAllocWriteCombined for the H2D
source; the D2H target stays ordinary pinned memory.)  One GPU."""
import ctypes as C
import subprocess

import numpy as np
import torch

rt = C.CDLL("libcudart.so.12")
rt.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
n = 128 << 20   # floats: 512 MiB
print(subprocess.run("lscpu | grep -i 'numa\\|socket\\|model name'; nvidia-smi topo -m | head -8", shell=True, capture_output=True, text=True).stdout)


def host(flags):
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), n * 4, flags) == 0
    a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(n,))
    a[:] = 1.0
    return torch.from_numpy(a)


torch.cuda.init()
dx = torch.empty(n, device="cuda")
dy = torch.ones(n, device="cuda")
s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()
for name, fx, fy in (("default / default", 0, 0), ("write-combined x / default y", 4, 0), ("default / default again", 0, 0)):
    x, y = host(fx), host(fy)
    print(name, "pinned:", x.is_pinned(), y.is_pinned())
    for mode in ("h2d", "d2h", "both"):
        def go(r):
            for _ in range(r):
                if mode != "d2h":
                    with torch.cuda.stream(s_up):
                        dx.copy_(x, non_blocking=True)
                if mode != "h2d":
                    with torch.cuda.stream(s_down):
                        y.copy_(dy, non_blocking=True)
            torch.cuda.synchronize()
        go(2)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        torch.cuda.synchronize()
        import time
        t0 = time.perf_counter()
        go(8)
        dt = (time.perf_counter() - t0) / 8
        print(f"   {mode:5s} {dt*1e3:7.2f} ms per 512 MiB each way   {n*4/dt/1e9:6.1f} GB/s per direction", flush=True)
