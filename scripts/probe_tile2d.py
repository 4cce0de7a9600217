"""Gcells/s of the 8-byte rank-2 tiled kernels (tile shape experiments: MB_T2_WX64 / MB_T2_R64 variants of the library
are selected with MASSIV_B200_LIB).  Prints one line per case."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import massiv_b200 as M

dev = M.Device(0)
rng = np.random.default_rng(0)
n = 16384
for dt in (np.float64, np.int64):
    if dt == np.float64:
        t = torch.rand((n, n), dtype=torch.float64, device="cuda")
    else:
        t = torch.randint(-1000, 1000, (n, n), dtype=torch.int64, device="cuda")
    o = torch.empty_like(t)
    a, b = dev.wrap(t.data_ptr(), (n, n), dt), dev.wrap(o.data_ptr(), (n, n), dt)
    if dt == np.float64:
        k3, k5, k7 = (rng.standard_normal((k, k)) for k in (3, 5, 7))
        cases = {"avg3x3": (M.Edge, M.avgStencil(M.Sz(3, 3))), "conv3x3": (M.Reflect, M.makeConvolutionStencilFromKernel(k3)),
                 "conv5x5": (M.Edge, M.makeConvolutionStencilFromKernel(k5)), "corr7x7": (M.Edge, M.makeCorrelationStencilFromKernel(k7)),
                 "sum5x5": (M.Wrap, M.sumStencil(M.Sz(5, 5))), "conv1x5": (M.Edge, M.makeConvolutionStencilFromKernel(k5[:1])),
                 "conv5x1": (M.Edge, M.makeConvolutionStencilFromKernel(k5[:, :1].copy())), "fill0 conv3x3": (M.Fill(0.0), M.makeConvolutionStencilFromKernel(k3))}
    else:
        cases = {"max3x3 i64": (M.Edge, M.maxStencil(M.Sz(3, 3))), "sum3x3 i64": (M.Edge, M.sumStencil(M.Sz(3, 3))), "min2x2 i64": (M.Edge, M.minStencil(M.Sz(2, 2)))}
    for name, (border, st) in cases.items():
        dw = M.mapStencil(border, st, a)
        for _ in range(3):
            M.computeInto(b, dw)
        dev.sync()
        dev.timer_start()
        for _ in range(10):
            M.computeInto(b, dw)
        ms = dev.timer_stop() / 10
        print(f"{name:16s} {dev.last_kernel:12s} {n * n / ms / 1e6:8.1f} Gcells/s", flush=True)
