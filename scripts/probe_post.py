"""Where do mapped (fmap'ed) 2-D stencils run fastest: the compile-time-shaped tiled kernel followed by the post-map
pass over the output, or the small-window kernel with the maps fused into its store epilogue?  Prints Gcells/s."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import massiv_b200 as M

dev = M.Device(0)
rng = np.random.default_rng(0)
for dt, n in ((np.float32, 16384), (np.float64, 16384)):
    t = torch.rand((n, n), dtype=torch.float32 if dt == np.float32 else torch.float64, device="cuda")
    o = torch.empty_like(t)
    a, b = dev.wrap(t.data_ptr(), (n, n), dt), dev.wrap(o.data_ptr(), (n, n), dt)
    k3 = rng.standard_normal((3, 3)).astype(dt)
    k5 = rng.standard_normal((5, 5)).astype(dt)
    cases = {"conv3x3": M.makeConvolutionStencilFromKernel(k3), "conv3x3 |.|*c": abs(M.makeConvolutionStencilFromKernel(k3)) * dt(0.5),
             "avg3x3": M.avgStencil(M.Sz(3, 3)), "avg3x3 +c": M.avgStencil(M.Sz(3, 3)) + dt(1),
             "conv5x5": M.makeConvolutionStencilFromKernel(k5), "conv5x5 sqrt|.|": M.sqrt(abs(M.makeConvolutionStencilFromKernel(k5)))}
    for name, st in cases.items():
        for pref in (0, 1):
            dev.set_option("prefer_smallwin", pref)
            dw = M.mapStencil(M.Edge, st, a)
            for _ in range(3):
                M.computeInto(b, dw)
            dev.sync()
            dev.timer_start()
            for _ in range(10):
                M.computeInto(b, dw)
            ms = dev.timer_stop() / 10
            print(f"{np.dtype(dt).name:8s} {name:18s} prefer_smallwin={pref} {dev.last_kernel:12s} {n * n / ms / 1e6:8.1f} Gcells/s", flush=True)
dev.set_option("prefer_smallwin", 0)
