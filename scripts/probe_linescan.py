"""Device-resident timing of the line-window kernels (rows and columns variants) with CUDA events."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import massiv_b200 as M
from massiv_b200 import _abi, integral as I

dev = M.Device.default()
L = dev._L
rng = np.random.default_rng(0)
for name, shape, dim, n in (("rows n=16", (16384, 16385), 1, 16), ("cols n=16", (16385, 16384), 2, 16), ("rows n=100", (8192, 20001), 1, 100),
                            ("cols n=100", (20001, 8192), 2, 100), ("rows midpoint n=2", (16384, 16384), 1, 2)):
    a = dev.fromHost(rng.random(shape))
    rank = 2
    stride = tuple(n if i == rank - dim else 1 for i in range(rank))
    st = I.simpsonsStencil(np.float64(0.01), dim, n) if "midpoint" not in name else I.midpointStencil(np.float64(0.01), dim, n)
    dw = M.mapStencil(M.Fill(0.0), st, a)
    d = dw.desc(stride)
    out = dev.alloc(dw.size(stride), a.dtype, a.elem)
    dims = _abi.dims_array(shape)
    for _ in range(3):
        dev.check(L.mb_run_dev(dev.handle, C.byref(d), dims, C.c_void_p(a.ptr), C.c_void_p(out.ptr)))
    dev.sync()
    dev.check(L.mb_timer_start(dev.handle))
    reps = 10
    for _ in range(reps):
        dev.check(L.mb_run_dev(dev.handle, C.byref(d), dims, C.c_void_p(a.ptr), C.c_void_p(out.ptr)))
    ms = C.c_float()
    dev.check(L.mb_timer_stop(dev.handle, C.byref(ms)))
    nbytes = (np.prod(shape) + np.prod(dw.size(stride))) * 8
    print(f"{name:20s} {dev.last_kernel:14s} {ms.value / reps:7.3f} ms  {nbytes / (ms.value / reps * 1e-3) / 1e9:7.0f} GB/s", flush=True)
    a.free(); out.free()
