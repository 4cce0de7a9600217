#!/usr/bin/env python
"""bench.py — stencil Gcells/s and achieved HBM GB/s vs roofline (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One JSON line on stdout (rank 0).  A "step" is one pass of the stencil over the workload's grid.
Workloads (BASELINE.json configs):
    heat3d   3-D 7-point heat step, Double, Fill 0, 1536^3 per GPU, iterated   (default; the config
             that runs at 1/2/4/8 GPUs — slab-sharded with per-step halo exchange, weak scaling)
    avg3x3   avgStencil 3x3 Double, Edge, 16384^2, single pass
    sobel    Sobel magnitude Float, Reflect, 32768^2, single pass
    gauss7   7x7 Gaussian conv Float, Edge, 32768^2, direct
    gauss7sep  same, separable two-pass (mb_run_sep2)
    life     Game of Life Word8, Wrap, 2048^2, iterated

`--impl reference` times the CPU restatement of massiv's Par path (oracle/, all host threads) on a
bounded sample of the same workload: no GHC exists in this image, so massiv itself cannot run.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (dtype, full grid, bytes/cell algorithmic, iterated, config text)
    "heat3d": dict(dtype="f64", grid=(1536, 1536, 1536), bpc=16, iterated=True,
                   text="3D 7-point Laplacian heat step Double, Fill 0 border, makeConvolutionStencilFromKernel 3x3x3"),
    "avg3x3": dict(dtype="f64", grid=(16384, 16384), bpc=16, iterated=False, text="avgStencil 3x3 Double, Edge border"),
    "sobel": dict(dtype="f32", grid=(32768, 32768), bpc=8, iterated=False,
                  text="Sobel gradient magnitude (two 3x3 convolution stencils, fused) Float, Reflect border"),
    "gauss7": dict(dtype="f32", grid=(32768, 32768), bpc=8, iterated=False,
                   text="7x7 Gaussian makeConvolutionStencilFromKernel Float, Edge border, direct"),
    "gauss7sep": dict(dtype="f32", grid=(32768, 32768), bpc=8, iterated=False,
                      text="7x7 Gaussian as separable 1x7 then 7x1 two-pass, Float, Edge border"),
    "life": dict(dtype="u8", grid=(2048, 2048), bpc=2, iterated=True, text="Game of Life 3x3 Word8, Wrap border"),
    # strided evaluation.  Cells are OUTPUT cells; bytes are input + output.
    "pool3": dict(dtype="i64", grid=(24576, 24576), stride=(3, 3), iterated=False,
                  text="computeWithStride (Stride 3) . mapStencil Edge (maxStencil (Sz 3)), Int (Stencil.hs:360-364 at scale)"),
    "simpson": dict(dtype="f64", grid=(16384, 16385), stride=(1, 16), iterated=False,
                    text="integrateWith simpsonsStencil (Dim 1) 16: Simpson's rule along the rows, Double, Fill 0 "
                         "(Numeric/Integral.hs:99-135)"),
}
NPDT = {"f64": np.float64, "f32": np.float32, "u8": np.uint8, "i64": np.int64}
SIMPSON_DX = 1.0 / 16.0


def out_shape(name, shape):
    """output extents of the workload on an input of `shape` (strideSize of the same-size mapStencil result)"""
    st = WORKLOADS[name].get("stride")
    return tuple(shape) if st is None else tuple((n - 1) // s + 1 for n, s in zip(shape, st))


def bytes_per_cell(name, shape):
    wl = WORKLOADS[name]
    if "bpc" in wl:
        return wl["bpc"]
    item = np.dtype(NPDT[wl["dtype"]]).itemsize
    return item * (float(np.prod(shape)) + float(np.prod(out_shape(name, shape)))) / float(np.prod(out_shape(name, shape)))


def gauss7():
    x = np.arange(-3, 4, dtype=np.float64)
    g = np.exp(-0.5 * x * x)
    g /= g.sum()
    return g.astype(np.float32)


def heat_kernel(r=0.1):
    k = np.zeros((3, 3, 3), np.float64)
    k[1, 1, 1] = 1 - 6 * r
    for ix in ((0, 1, 1), (2, 1, 1), (1, 0, 1), (1, 2, 1), (1, 1, 0), (1, 1, 2)):
        k[ix] = r
    return k


SOBEL_X = np.array([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]], np.float32)


def oracle_spec(name):
    """(kind, kwargs) for oracle.apply of one pass of the workload (CPU legs only)"""
    if name == "heat3d":
        return "conv", dict(size=(3, 3, 3), coeffs=heat_kernel().reshape(-1).tolist(), border="fill", fill=0.0)
    if name == "avg3x3":
        return "avg", dict(size=(3, 3), border="edge")
    if name == "sobel":
        return "mag2", dict(size=(3, 3), coeffs=SOBEL_X.reshape(-1).tolist(), coeffs2=SOBEL_X.T.reshape(-1).tolist(),
                            border="reflect")
    if name == "gauss7":
        g = gauss7()
        return "conv", dict(size=(7, 7), coeffs=np.outer(g, g).astype(np.float32).reshape(-1).tolist(), border="edge")
    if name == "life":
        return "life", dict(size=(3, 3), border="wrap")
    if name == "pool3":
        return "max", dict(size=(3, 3), border="edge", stride=(3, 3))
    if name == "simpson":
        return "int_simpson", dict(size=(1, 17), coeffs=[SIMPSON_DX], border="fill", fill=0.0, stride=(1, 16))
    raise KeyError(name)


def stencil_spec(name, M):
    """(border, stencil[, second stencil]) in the product's host API"""
    if name == "heat3d":
        st = M.makeConvolutionStencilFromKernel(heat_kernel())
        # No finite-input promise: the library's checked optimistic kernels give the reference's result for every
        # input.  MASSIV_B200_PROMISE=1 adds .assumeFinite() (the unchecked 7-tap kernels) for comparison.
        return M.Fill(0.0), (st.assumeFinite() if os.environ.get("MASSIV_B200_PROMISE") else st)
    if name == "avg3x3":
        return M.Edge, M.avgStencil(M.Sz(3, 3))
    if name == "sobel":
        sx, sy = M.makeConvolutionStencilFromKernel(SOBEL_X), M.makeConvolutionStencilFromKernel(SOBEL_X.T.copy())
        return M.Reflect, M.sqrt(sx ** 2 + sy ** 2)
    if name == "gauss7":
        g = gauss7()
        return M.Edge, M.makeConvolutionStencilFromKernel(np.outer(g, g).astype(np.float32))
    if name == "gauss7sep":
        g = gauss7()
        return M.Edge, M.makeConvolutionStencilFromKernel(g.reshape(1, 7)), M.makeConvolutionStencilFromKernel(g.reshape(7, 1))
    if name == "life":
        return M.Wrap, M.lifeStencil
    if name == "pool3":
        return M.Edge, M.maxStencil(M.Sz(3, 3))
    if name == "simpson":
        from massiv_b200 import integral as I
        return M.Fill(0.0), I.simpsonsStencil(np.float64(SIMPSON_DX), 1, 16)
    raise KeyError(name)


# ---- synthetic inputs -----------------------------------------------------------------------------

def splitmix64_np(idx: np.ndarray, seed: int) -> np.ndarray:
    x = (idx.astype(np.uint64) + np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15))
    z = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def synth_host(name, shape, seed=2020, first=0):
    """synthetic input of SURVEY.md §8d: uniform [0,1) from splitmix64(seed + linear index)"""
    n = int(np.prod(shape))
    h = splitmix64_np(np.arange(first, first + n, dtype=np.uint64), seed)
    dt = WORKLOADS[name]["dtype"]
    if dt == "f64":
        return ((h >> np.uint64(11)).astype(np.float64) * 2.0 ** -53).reshape(shape)
    if dt == "f32":
        return ((h >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)).reshape(shape)
    if dt == "i64":
        return (h >> np.uint64(11)).astype(np.int64).reshape(shape)
    return (h >> np.uint64(63)).astype(np.uint8).reshape(shape)


def synth_device(torch, name, shape, device, seed=2020, first=0):
    """same generator on the GPU (int64 arithmetic wraps like uint64; logical shifts emulated)"""
    n = int(np.prod(shape))
    dt = WORKLOADS[name]["dtype"]
    tdt = {"f64": torch.float64, "f32": torch.float32, "u8": torch.uint8, "i64": torch.int64}[dt]
    out = torch.empty(n, dtype=tdt, device=device)
    chunk = 1 << 26

    def lsr(x, s):
        return (x >> s) & ((1 << (64 - s)) - 1)

    def wrap(v):
        v &= (1 << 64) - 1
        return v - (1 << 64) if v >= (1 << 63) else v

    for a in range(0, n, chunk):
        b = min(n, a + chunk)
        x = torch.arange(first + a, first + b, dtype=torch.int64, device=device) + wrap(seed + 0x9E3779B97F4A7C15)
        z = (x ^ lsr(x, 30)) * wrap(0xBF58476D1CE4E5B9)
        z = (z ^ lsr(z, 27)) * wrap(0x94D049BB133111EB)
        h = z ^ lsr(z, 31)
        if dt == "f64":
            out[a:b] = lsr(h, 11).to(torch.float64) * 2.0 ** -53
        elif dt == "f32":
            out[a:b] = lsr(h, 40).to(torch.float32) * 2.0 ** -24
        elif dt == "i64":
            out[a:b] = lsr(h, 11)
        else:
            out[a:b] = lsr(h, 63).to(torch.uint8)
        del x, z, h
    return out.view(*shape)


# ---- clocks ------------------------------------------------------------------------------------------

class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons during the timed region"""

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def load_traffic(name, kernel, grid):
    """measured DRAM bytes per launch of the dominant kernel (ncu capture summarised under profiles/), or None"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(name)
        if t and t["kernel"] == kernel and list(t["grid"]) == list(grid):
            return t["bytes_per_launch"]
    except Exception:
        pass
    return None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---- CPU reference arm ------------------------------------------------------------------------------------

def cpu_sample_shape(name):
    """bounded sample of the workload: same row/plane extents, fewer outer planes/rows"""
    g = WORKLOADS[name]["grid"]
    if name == "heat3d":
        return (48, g[1], g[2]), 2
    if name == "life":
        return g, 20
    if name == "avg3x3":
        return (4096, g[1]), 1
    return (4096, g[1]), 1


def run_cpu(name, steps, warmup, threads=None):
    from oracle import oracle as O
    threads = threads or os.cpu_count() or 1
    wl = "gauss7" if name == "gauss7sep" else name
    shape, inner = cpu_sample_shape(wl)
    a = synth_host(wl, shape)
    kind, kw = oracle_spec(wl)
    iterated = WORKLOADS[wl]["iterated"]

    def once():
        if name == "gauss7sep":
            g = gauss7().tolist()
            t = O.apply("conv", a, size=(1, 7), coeffs=g, border="edge", nthreads=threads)
            return O.apply("conv", t, size=(7, 1), coeffs=g, border="edge", nthreads=threads)
        if iterated:
            return O.iterate(kind, a, inner, nthreads=threads, **kw)
        return O.apply(kind, a, nthreads=threads, **kw)

    for _ in range(warmup):
        once()
    t0 = time.perf_counter()
    for _ in range(steps):
        once()
    dt = time.perf_counter() - t0
    passes = inner if iterated else 1
    cells = float(np.prod(out_shape(wl, shape))) * passes * steps
    return dict(value=cells / dt / 1e9, seconds=dt, shape=list(shape), passes_per_step=passes, threads=threads)


def reference_main(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 5))
    warm = max(1, min(args.warmup, 1))
    r = run_cpu(args.workload, steps, warm)
    wl = WORKLOADS[args.workload]
    sample = (f"{'x'.join(map(str, r['shape']))} sample of the {'x'.join(map(str, wl['grid']))} grid, "
              f"{r['passes_per_step']} pass(es) per step, {steps} steps; C restatement of massiv's Par schedule "
              f"(oracle/, gcc -O2, no FMA), {r['threads']} threads — not GHC-compiled massiv (no GHC in the image)")
    line = {
        "impl": "reference", "metric": "stencil_gcells_per_s", "value": r["value"], "unit": "Gcells/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": r["seconds"] / steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": wl["dtype"], "data": "synthetic",
        "config": {"workload": args.workload, "description": wl["text"], "grid": list(wl["grid"])},
        "cpu_baseline": {"value": r["value"], "unit": "Gcells/s", "cores": r["threads"], "kind": "port", "sample": sample},
        "e2e": {"value": r["value"], "unit": "Gcells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---- GPU arm ---------------------------------------------------------------------------------------------

def gpu_main(args):
    import torch
    import torch.distributed as dist

    import massiv_b200 as M
    from massiv_b200 import _abi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the stencil path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    name = args.workload
    wl = WORKLOADS[name]
    grid = tuple(args.grid) if args.grid else wl["grid"]
    npdt = NPDT[wl["dtype"]]
    dev = M.Device(local)
    L = dev._L
    # the library launches on the stream it is given; torch's legacy default stream is handle 0, which
    # the ABI reads as "use your own stream", so the bench runs everything on an explicit torch stream
    tstream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(tstream)
    dev.set_stream(tstream.cuda_stream)
    spec = stencil_spec(name, M)
    border = spec[0]
    stride = wl.get("stride")
    ogrid = out_shape(name, grid)
    cells = float(np.prod(ogrid))
    tdev = torch.device("cuda", local)

    # device-resident input (synthetic, generated on the GPU; rank r continues the index stream)
    a_t = synth_device(torch, name, grid, tdev, first=rank * int(np.prod(grid)))
    b_t = torch.empty(ogrid, dtype=a_t.dtype, device=tdev)
    a = dev.wrap(a_t.data_ptr(), grid, npdt)
    b = dev.wrap(b_t.data_ptr(), ogrid, npdt)
    dims = _abi.dims_array(grid)

    sharded = (world > 1 or os.environ.get("MASSIV_B200_BENCH_SHARDED")) and wl["iterated"] and name == "heat3d"
    if sharded:
        # weak scaling: every rank owns a full-size slab of a (world x taller) global grid; the slab sits
        # between its ghost planes and the library exchanges halos with ncclSend/ncclRecv every step
        from massiv_b200 import sharding
        sharding.init_comm(dev, dist if world > 1 else None)
        d = M.mapStencil(border, spec[1], a).desc()
        lo, hi = C.c_int64(), C.c_int64()
        dev.check(L.mb_shard_halo(C.byref(d), C.byref(lo), C.byref(hi)))
        plane = int(np.prod(grid[1:]))
        ga_t = torch.empty((lo.value + grid[0] + hi.value) * plane, dtype=a_t.dtype, device=tdev)
        gb_t = torch.empty_like(ga_t)
        ga_t[lo.value * plane:(lo.value + grid[0]) * plane].copy_(a_t.view(-1))
        del a_t, b_t
        torch.cuda.synchronize()
        gdims = _abi.dims_array((grid[0] * world,) + tuple(grid[1:]))
        res = C.c_void_p()

        def run(k):
            dev.check(L.mb_iterate_sharded_dev(dev.handle, C.byref(d), gdims, dims, C.c_void_p(ga_t.data_ptr()),
                                               C.c_void_p(gb_t.data_ptr()), k, C.byref(res)))
    elif name == "gauss7sep":
        d1 = M.mapStencil(border, spec[1], a).desc()
        d2 = M.mapStencil(border, spec[2], a).desc()
        tmp_t = torch.empty_like(a_t)

        def run(k):
            for _ in range(k):
                dev.check(L.mb_run_sep2_dev(dev.handle, C.byref(d1), C.byref(d2), dims, C.c_void_p(a.ptr), C.c_void_p(b.ptr),
                                            C.c_void_p(tmp_t.data_ptr())))
    elif wl["iterated"]:
        d = M.mapStencil(border, spec[1], a).desc()
        res = C.c_void_p()

        def run(k):
            dev.check(L.mb_iterate_dev(dev.handle, C.byref(d), dims, C.c_void_p(a.ptr), C.c_void_p(b.ptr), k, C.byref(res)))
    else:
        dw = M.mapStencil(border, spec[1], a)
        d = dw.desc(stride)
        assert dw.size(stride) == ogrid, (dw.size(stride), ogrid)

        def run(k):
            for _ in range(k):
                dev.check(L.mb_run_dev(dev.handle, C.byref(d), dims, C.c_void_p(a.ptr), C.c_void_p(b.ptr)))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    run(max(args.warmup, 3))
    barrier()
    l0, x0 = dev.launch_count, dev.aux_launch_count
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    run(args.steps)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    launches = dev.launch_count - l0
    main_launches = launches - (dev.aux_launch_count - x0)  # the stencil kernels proper (helpers return in microseconds)
    kernel = dev.last_kernel
    if world > 1:
        t = torch.tensor([ms], device=tdev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = cells * world * args.steps / (ms * 1e-3) / 1e9

    # ---- e2e: the reference-facing host->host call on pinned host buffers -------------------------------
    e2e = None
    if not args.no_e2e:
        run = None
        a_t = b_t = ga_t = gb_t = tmp_t = None  # release the device-resident buffers before the host->host leg
        torch.cuda.empty_cache()
        e2e = measure_e2e(args, dev, M, name, grid, npdt, spec, world, rank, dist if world > 1 else None, torch, tdev, sharded)

    if rank == 0:
        peak, peak_src = load_peaks()
        per_launch_ms = ms / max(main_launches, 1)
        passes_per_launch = args.steps / max(main_launches, 1)
        bpc = bytes_per_cell(name, grid)
        alg_bytes = cells * bpc * passes_per_launch
        if name == "gauss7sep" and main_launches == 2 * args.steps:
            alg_bytes = cells * bpc  # two-pass fallback: each launch moves one read + one write
        achieved = alg_bytes / (per_launch_ms * 1e-3) / 1e9
        cpu = None
        if not args.no_cpu:
            r = run_cpu(name, 2 if name != "life" else 1, 1)
            cpu = {"value": r["value"], "unit": "Gcells/s", "cores": r["threads"], "kind": "port",
                   "sample": f"{'x'.join(map(str, r['shape']))} sample, {r['passes_per_step']} pass(es)/step; C restatement of "
                             "massiv's Par schedule (oracle/), not GHC-compiled massiv"}
        line = {
            "metric": "stencil_gcells_per_s", "value": value, "unit": "Gcells/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": wl["dtype"], "data": "synthetic",
            "config": {"workload": name, "description": wl["text"], "grid_per_gpu": list(grid), "cells": "output cells" if stride else "cells",
                       "l2": "inputs larger than L2" if cells * bpc > 256e6 else "grid is L2-resident by definition of the workload",
                       "kernel": kernel, "parallelism": f"slab{world}" if world > 1 else "single",
                       "halo_transport": L.mb_shard_transport(dev.handle).decode() if sharded else None},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": load_traffic(name, kernel, grid) if world == 1 else None, "peak_source": peak_src, "kernel": kernel,
                         "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": per_launch_ms},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "gpu_launches_main": int(main_launches), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def measure_e2e(args, dev, M, name, grid, npdt, spec, world, rank, dist, torch, tdev, sharded):
    """The same metric through the reference-facing host->host call — mb_run (compute . mapStencil),
    mb_run_sep2, mb_iterate (iterateUntil) or mb_iterate_sharded — on pinned HOST buffers: the H2D copy
    of the input and the D2H copy of the result are inside the timed region.  An iterated workload is
    ONE call of the configured number of generations / steps, as the reference's iterateUntil is."""
    from massiv_b200 import _abi
    wl = WORKLOADS[name]
    L = dev._L
    border = spec[0]
    item = np.dtype(npdt).itemsize
    try:
        avail = int([ln for ln in open("/proc/meminfo") if ln.startswith("MemAvailable")][0].split()[1]) * 1024
    except Exception:
        avail = 32 << 30
    budget = int(avail * 0.45 / max(world, 1))  # two pinned buffers per rank must fit in host memory
    g = list(grid)
    while int(np.prod(g)) * item * 2 > budget and g[0] > 64:
        g[0] //= 2
    g = tuple(g)
    n = int(np.prod(g))
    nbytes = n * item
    stride = wl.get("stride")
    n_out = int(np.prod(out_shape(name, g)))
    obytes = n_out * item
    hin, hout = C.c_void_p(), C.c_void_p()
    dev.check(L.mb_host_alloc(dev.handle, nbytes, C.byref(hin)))
    dev.check(L.mb_host_alloc(dev.handle, obytes, C.byref(hout)))
    src = np.ctypeslib.as_array(C.cast(hin, C.POINTER(C.c_uint8)), shape=(nbytes,)).view(npdt).reshape(g)
    t = synth_device(torch, name, g, tdev, first=rank * n)  # fill the pinned input from the device generator
    torch.cuda.synchronize()
    dev.check(L.mb_download(dev.handle, hin, C.c_void_p(t.data_ptr()), nbytes))
    del t
    torch.cuda.empty_cache()
    dims = _abi.dims_array(g)
    steps = args.e2e_steps or (500 if name == "heat3d" else 1000 if name == "life" else 1)
    if name == "gauss7sep":
        d1, d2 = M.mapStencil(border, spec[1], src).desc(), M.mapStencil(border, spec[2], src).desc()
        call, passes, cname = (lambda k: dev.check(L.mb_run_sep2(dev.handle, C.byref(d1), C.byref(d2), dims, hin, hout))), 1, "mb_run_sep2"
    elif sharded:
        d = M.mapStencil(border, spec[1], src).desc()
        gdims = _abi.dims_array((g[0] * world,) + tuple(g[1:]))
        call, passes, cname = (lambda k: dev.check(L.mb_iterate_sharded(dev.handle, C.byref(d), gdims, dims, hin, hout, k))), steps, "mb_iterate_sharded"
    elif wl["iterated"]:
        d = M.mapStencil(border, spec[1], src).desc()
        call, passes, cname = (lambda k: dev.check(L.mb_iterate(dev.handle, C.byref(d), dims, hin, hout, k))), steps, "mb_iterate"
    else:
        d = M.mapStencil(border, spec[1], src).desc(stride)
        call, passes, cname = (lambda k: dev.check(L.mb_run(dev.handle, C.byref(d), dims, hin, hout))), 1, "mb_run"
    for _ in range(3):
        call(2)  # warm-up: allocates the context's device scratch, touches the pinned pages
    reps = 1 if wl["iterated"] else 5
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        call(steps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if dist is not None:
        tt = torch.tensor([dt], device=tdev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
    dev.check(L.mb_host_free(dev.handle, hin))
    dev.check(L.mb_host_free(dev.handle, hout))
    val = float(n_out) * world * passes * reps / dt / 1e9
    return {"value": val, "unit": "Gcells/s", "h2d_bytes_per_step": nbytes / passes, "d2h_bytes_per_step": obytes / passes,
            "grid_per_gpu": list(g), "passes_per_call": passes, "seconds_per_call": dt / reps, "call": cname,
            "timing": "host wall clock around the blocking C call, max over ranks"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=os.environ.get("MASSIV_B200_WORKLOAD", "heat3d"), choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", type=int, nargs="+", default=None, help="override the grid (debugging only)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0)
    args = ap.parse_args()
    if args.impl == "reference":
        reference_main(args)
    else:
        gpu_main(args)


if __name__ == "__main__":
    main()
