"""Small-window kernel, dense windows: Gcells/s with and without the vertical reuse (MASSIV_B200_SW_NO_DENSE=1)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import massiv_b200 as M

dev = M.Device(0)
rng = np.random.default_rng(0)
n = 16384
cases = []
for dt, tdt in ((np.uint8, torch.uint8), (np.float32, torch.float32), (np.float64, torch.float64), (np.int16, torch.int16)):
    if dt in (np.float32, np.float64):
        t = torch.rand((n, n), dtype=tdt, device="cuda")
    else:
        t = torch.randint(0, 100, (n, n), dtype=tdt, device="cuda")
    o = torch.empty_like(t)
    a, b = dev.wrap(t.data_ptr(), (n, n), dt), dev.wrap(o.data_ptr(), (n, n), dt)
    todo = {"sum9x9": M.sumStencil(M.Sz(9, 9)), "sum5x5": M.sumStencil(M.Sz(5, 5)), "sum3x11": M.sumStencil(M.Sz(3, 11))}
    if dt in (np.float32, np.float64):
        todo["conv9x9"] = M.makeConvolutionStencilFromKernel(rng.standard_normal((9, 9)).astype(dt))
        todo["corr11x11"] = M.makeCorrelationStencilFromKernel(rng.standard_normal((11, 11)).astype(dt))
    else:
        todo["max9x9"] = M.maxStencil(M.Sz(9, 9))
    for name, st in todo.items():
        dw = M.mapStencil(M.Edge, st, a)
        for _ in range(2):
            M.computeInto(b, dw)
        dev.sync()
        dev.timer_start()
        for _ in range(5):
            M.computeInto(b, dw)
        ms = dev.timer_stop() / 5
        print(f"{np.dtype(dt).name:8s} {name:10s} {dev.last_kernel:10s} {n * n / ms / 1e6:8.1f} Gcells/s", flush=True)
