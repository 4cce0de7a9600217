#!/usr/bin/env python
"""bench.py — stencil Gcells/s and achieved HBM GB/s vs roofline (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One JSON line on stdout (rank 0).  A "step" is one pass of the stencil over the workload's grid.
The headline workload is heat3d (BASELINE.json's config 5, the one that runs at 1/2/4/8 GPUs).  At N=1 the
default run also carries, in the same line,
    "configs": one record per other BASELINE config (Life at 1000 generations, avg3x3, Sobel, Gauss 7x7 direct and
               separable): value, roofline{frac, traffic}, cpu_baseline, e2e, clocks, verify;
    "verify":  full-size correctness checks made OUTSIDE the timed regions: independent kernels of the library
               compared bit for bit on the device at the size the number is quoted on (two-step vs single-step
               vs strict 27-tap heat kernel; bit-packed vs byte-wise Life; tiled vs generic 2-D kernels), and an
               oracle check of a random sub-slab (with its halo) of every config;
and at N>1
    "parity":  the slab-sharded loop against the single-GPU iteration of the same global array, bit for bit,
               per halo transport (peer-memory copies and ncclSend/ncclRecv), including Inf / NaN on slab boundaries.
Workloads (BASELINE.json configs):
    heat3d   3-D 7-point heat step, Double, Fill 0, 1536^3 per GPU, iterated   (default)
    avg3x3   avgStencil 3x3 Double, Edge, 16384^2, single pass
    sobel    Sobel magnitude Float, Reflect, 32768^2, single pass
    gauss7   7x7 Gaussian conv Float, Edge, 32768^2, direct
    gauss7sep  same, separable two-pass (mb_run_sep2)
    life     Game of Life Word8, Wrap, 2048^2, iterated
    pool3 / simpson   SURVEY §8 (f) rows: strided evaluation
    sobel_operator / sobel_closure   the reference's published Sobel benchmark (closure-form stencils) at its own size and large
    sum9x9_u8   a 9x9 window over a Word8 image (small-window kernel)
    avg3d       avgStencil 3x3x3 Float 1024^3 (small-window kernel, rank 3)

`--impl reference` times the CPU restatement of massiv's Par path (oracle/, a pinned number of host threads) on a
bounded sample of the same workload: no GHC exists in this image, so massiv itself cannot run.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (dtype, full grid, bytes/cell algorithmic, iterated, config text)
    "heat3d": dict(dtype="f64", grid=(1536, 1536, 1536), bpc=16, iterated=True, config_id="C5",
                   text="3D 7-point Laplacian heat step Double, Fill 0 border, makeConvolutionStencilFromKernel 3x3x3"),
    "avg3x3": dict(dtype="f64", grid=(16384, 16384), bpc=16, iterated=False, config_id="C2", text="avgStencil 3x3 Double, Edge border"),
    "sobel": dict(dtype="f32", grid=(32768, 32768), bpc=8, iterated=False, config_id="C3",
                  text="Sobel gradient magnitude (two 3x3 convolution stencils, fused) Float, Reflect border"),
    "gauss7": dict(dtype="f32", grid=(32768, 32768), bpc=8, iterated=False, config_id="C4-direct",
                   text="7x7 Gaussian makeConvolutionStencilFromKernel Float, Edge border, direct"),
    "gauss7sep": dict(dtype="f32", grid=(32768, 32768), bpc=8, iterated=False, config_id="C4-separable",
                      text="7x7 Gaussian as separable 1x7 then 7x1 two-pass, Float, Edge border"),
    "life": dict(dtype="u8", grid=(2048, 2048), bpc=2, iterated=True, config_id="C1", text="Game of Life 3x3 Word8, Wrap border"),
    # strided evaluation.  Cells are OUTPUT cells; bytes are input + output.
    "pool3": dict(dtype="i64", grid=(24576, 24576), stride=(3, 3), iterated=False, config_id="f1",
                  text="computeWithStride (Stride 3) . mapStencil Edge (maxStencil (Sz 3)), Int (Stencil.hs:360-364 at scale)"),
    # the reference's OWN published benchmark (massiv/README.md:598-632, "Sobel/Par/Operator - Massiv"): the closure-form
    # Sobel operator of massiv-bench/src/Data/Massiv/Bench/Sobel.hs:15-44, 1600x1200 Double, Edge — 7.421 ms on a
    # 4-core i7-3740QM.  sobel_operator is that exact size (L2-resident: a latency figure), sobel_closure the same
    # operator on an array larger than L2.  Both run the small-window kernel (two tap lists + sqrt (x*x + y*y)).
    "sobel_operator": dict(dtype="f64", grid=(1600, 1200), bpc=16, iterated=False, config_id="published",
                           published_gcells=1600 * 1200 / 7.421e-3 / 1e9,
                           text="Sobel operator, closure form (makeConvolutionStencil), Double 1600x1200, Edge: the reference's published benchmark"),
    "sobel_closure": dict(dtype="f64", grid=(16384, 16384), bpc=16, iterated=False, config_id="a13",
                          text="Sobel operator, closure form (makeConvolutionStencil x2, sqrt (sX*sX + sY*sY)), Double 16384^2, Edge"),
    "sum9x9_u8": dict(dtype="u8", grid=(32768, 32768), bpc=2, iterated=False, config_id="a17", u8_full=True,
                      text="sumStencil (Sz2 9 9) Word8 (mod 256), Edge, 32768^2: a 9x9 window on a byte image"),
    "avg3d": dict(dtype="f32", grid=(1024, 1024, 1024), bpc=8, iterated=False, config_id="a8",
                  text="avgStencil (Sz3 3 3 3) Float, Edge, 1024^3: a 3-D fold (small-window kernel, rank 3)"),
    "simpson": dict(dtype="f64", grid=(16384, 16385), stride=(1, 16), iterated=False, config_id="f2",
                    text="integrateWith simpsonsStencil (Dim 1) 16: Simpson's rule along the rows, Double, Fill 0 "
                         "(Numeric/Integral.hs:99-135)"),
}
NPDT = {"f64": np.float64, "f32": np.float32, "u8": np.uint8, "i64": np.int64}
SIMPSON_DX = 1.0 / 16.0
# the other BASELINE configs, measured after the headline in the default N=1 run: (workload, timed steps)
# (Life: 1000 generations per call as the config says; the single passes repeat until the timed region is long
# enough for the clock sampler to see it)
CONFIG_RUNS = [("life", 1000), ("avg3x3", 200), ("sobel", 100), ("gauss7", 50), ("gauss7sep", 100),
               # beyond BASELINE.json's list: the reference's published benchmark and the small-window kernel's proofs
               ("sobel_operator", 2000), ("sobel_closure", 50), ("sum9x9_u8", 20), ("avg3d", 20)]
# measured unfused FP32 rate (mul and add issued separately, scripts/micro/fp_rate.cu on this pool's B200)
FP32_OPS_PER_S = 35.0e12
REF_THREADS = 16  # the reference arm and cpu_baseline use min(this, host cores) threads so that ratios compare across boxes


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


def oracle_from_desc(O, dw, a, threads, stride=None):
    """the CPU oracle on the very descriptor the product would run (the two structures share their layout)"""
    d = dw.desc(stride)
    od = O._Desc.from_buffer_copy(bytes(d))
    dims = O._dims(a.shape)
    out_dims = (C.c_int64 * O.MAX_RANK)()
    assert O.lib().mo_out_dims(C.byref(od), dims, out_dims) == 0
    out = np.empty(tuple(out_dims[i] for i in range(a.ndim)), a.dtype)
    st = O.lib().mo_apply(C.byref(od), dims, np.ascontiguousarray(a).ctypes.data, out.ctypes.data, threads)
    assert st == 0, st
    return out


def oracle_pass(O, name, a, threads):
    """one pass of the workload through the CPU oracle (gauss7sep: massiv's two computes, intermediate rounded)"""
    if name in ("sobel_operator", "sobel_closure", "sum9x9_u8", "avg3d"):
        import massiv_b200 as M   # host-side lowering only (numpy in, descriptor out): no device involved
        spec = stencil_spec(name, M)
        return oracle_from_desc(O, M.mapStencil(spec[0], spec[1], a), a, threads)
    if name == "gauss7sep":
        g = gauss7().tolist()
        t = O.apply("conv", a, size=(1, 7), coeffs=g, border="edge", nthreads=threads)
        return O.apply("conv", t, size=(7, 1), coeffs=g, border="edge", nthreads=threads)
    kind, kw = oracle_spec(name)
    return O.apply(kind, a, nthreads=threads, **kw)


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
    if name in ("sobel_operator", "sobel_closure"):
        # Bench/Sobel.hs:15-37: makeConvolutionStencil (Sz 3) (1 :. 1) $ \f -> f (-1 :. -1) (-1) . f (0 :. -1) (-2) . ...
        sx = M.makeConvolutionStencil(M.Sz(3, 3), (1, 1), [((-1, -1), -1), ((0, -1), -2), ((1, -1), -1), ((-1, 1), 1), ((0, 1), 2), ((1, 1), 1)])
        sy = M.makeConvolutionStencil(M.Sz(3, 3), (1, 1), [((-1, -1), -1), ((-1, 0), -2), ((-1, 1), -1), ((1, -1), 1), ((1, 0), 2), ((1, 1), 1)])
        return M.Edge, M.sqrt(sx * sx + sy * sy)   # Bench/Sobel.hs:39-44
    if name == "sum9x9_u8":
        return M.Edge, M.sumStencil(M.Sz(9, 9))
    if name == "avg3d":
        return M.Edge, M.avgStencil(M.Sz(3, 3, 3))
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
    return (h >> np.uint64(56 if WORKLOADS[name].get("u8_full") else 63)).astype(np.uint8).reshape(shape)


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
            out[a:b] = lsr(h, 56 if WORKLOADS[name].get("u8_full") else 63).to(torch.uint8)
        del x, z, h
    return out.view(*shape)


# ---- clocks ------------------------------------------------------------------------------------------

class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons during the timed region"""

    def __init__(self, index: int, period_ms: int = 50):
        self.index, self.rows, self.proc, self.period = index, [], None, period_ms

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", str(self.period)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def wait_first(self, timeout=3.0):
        """the first sample takes nvidia-smi a while: do not start the timed region before it exists"""
        t0 = time.time()
        while self.proc and not self.rows and time.time() - t0 < timeout:
            time.sleep(0.01)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
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

def ref_threads():
    return max(1, min(REF_THREADS, os.cpu_count() or 1))


def cpu_sample_shape(name):
    """bounded sample of the workload: same row/plane extents, fewer outer planes/rows"""
    g = WORKLOADS[name]["grid"]
    if name == "heat3d":
        return (48, g[1], g[2]), 2
    if name == "life":
        return g, 20
    return ((min(4096, g[0]), g[1]) if len(g) == 2 else (48,) + tuple(g[1:])), 1


def run_cpu(name, steps, warmup, threads=None):
    from oracle import oracle as O
    threads = threads or ref_threads()
    wl = "gauss7" if name == "gauss7sep" else name
    shape, inner = cpu_sample_shape(wl)
    a = synth_host(wl, shape)
    iterated = WORKLOADS[wl]["iterated"]

    def once():
        if iterated:
            kind, kw = oracle_spec(wl)
            return O.iterate(kind, a, inner, nthreads=threads, **kw)
        return oracle_pass(O, name, a, threads)

    for _ in range(warmup):
        once()
    t0 = time.perf_counter()
    for _ in range(steps):
        once()
    dt = time.perf_counter() - t0
    passes = inner if iterated else 1
    cells = float(np.prod(out_shape(wl, shape))) * passes * steps
    return dict(value=cells / dt / 1e9, seconds=dt, shape=list(shape), passes_per_step=passes, threads=threads)


def cpu_sample_text(name, r, steps):
    wl = WORKLOADS[name]
    return (f"{'x'.join(map(str, r['shape']))} sample of the {'x'.join(map(str, wl['grid']))} grid, "
            f"{r['passes_per_step']} pass(es) per step, {steps} steps; C restatement of massiv's Par schedule "
            f"(oracle/, gcc -O3, no FMA), {r['threads']} threads (pinned to min({REF_THREADS}, host cores)) — not GHC-compiled "
            "massiv (no GHC in the image)")


def reference_main(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # --steps / --warmup are honoured; each step is a bounded sample of the workload (a few hundred ms), so the
    # whole run stays within a couple of minutes for any sensible K and W
    steps, warm = max(1, args.steps), max(0, args.warmup)
    r = run_cpu(args.workload, steps, warm)
    wl = WORKLOADS[args.workload]
    sample = cpu_sample_text(args.workload, r, steps)
    line = {
        "impl": "reference", "metric": "stencil_gcells_per_s", "value": r["value"], "unit": "Gcells/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": r["seconds"] / steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": wl["dtype"], "data": "synthetic",
        "config": {"workload": args.workload, "description": wl["text"], "grid_per_gpu": list(wl["grid"]), "cells": "cells",
                   "sample": sample, "sample_grid": r["shape"], "passes_per_step": r["passes_per_step"],
                   "parallelism": f"{r['threads']} host threads"},
        "cpu_baseline": {"value": r["value"], "unit": "Gcells/s", "cores": r["threads"], "kind": "port", "sample": sample},
        "e2e": {"value": r["value"], "unit": "Gcells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---- GPU arm ---------------------------------------------------------------------------------------------

class Env:
    """what every leg of the GPU arm needs"""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        import massiv_b200 as M
        from massiv_b200 import _abi
        self.torch, self.dist, self.M, self.abi, self.args = torch, dist, M, _abi, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the stencil path has no CPU fallback")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.dev = M.Device(self.local)
        self.L = self.dev._L
        # the library launches on the stream it is given; torch's legacy default stream is handle 0, which
        # the ABI reads as "use your own stream", so the bench runs everything on an explicit torch stream
        self.tstream = torch.cuda.Stream(device=self.local)
        torch.cuda.set_stream(self.tstream)
        self.dev.set_stream(self.tstream.cuda_stream)
        self.tdev = torch.device("cuda", self.local)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def equal_dev(self, a_ptr, b_ptr, nbytes):
        """bitwise comparison on the device through the library's own compare kernel -> mismatching words"""
        n, first = C.c_uint64(), C.c_uint64()
        self.dev.check(self.L.mb_compare_dev(self.dev.handle, C.c_void_p(a_ptr), C.c_void_p(b_ptr), nbytes, C.byref(n), C.byref(first)))
        return int(n.value)


def measure_device(env, name, grid, steps, warmup, sharded=False, keep=None):
    """device-resident timing of one workload: W warm-up steps, K timed steps between CUDA events on the launching
    stream.  Returns the record and (through `keep`) the buffers for the verify leg."""
    torch, M, L, dev, _abi = env.torch, env.M, env.L, env.dev, env.abi
    wl = WORKLOADS[name]
    npdt = NPDT[wl["dtype"]]
    spec = stencil_spec(name, M)
    border = spec[0]
    stride = wl.get("stride")
    ogrid = out_shape(name, grid)
    cells = float(np.prod(ogrid))
    world, rank = env.world, env.rank
    a_t = synth_device(torch, name, grid, env.tdev, first=rank * int(np.prod(grid)))
    dims = _abi.dims_array(grid)
    calls = 1
    if sharded:
        # weak scaling: every rank owns a full-size slab of a (world x taller) global grid; the slab sits between its
        # ghost planes and the library exchanges halos every round (peer-memory copies, NCCL send/recv as fallback)
        from massiv_b200 import sharding
        sharding.init_comm(dev, env.dist if world > 1 else None)
        a = dev.wrap(a_t.data_ptr(), grid, npdt)
        d = M.mapStencil(border, spec[1], a).desc()
        lo, hi = C.c_int64(), C.c_int64()
        dev.check(L.mb_shard_halo(C.byref(d), C.byref(lo), C.byref(hi)))
        plane = int(np.prod(grid[1:]))
        ga_t = torch.empty((lo.value + grid[0] + hi.value) * plane, dtype=a_t.dtype, device=env.tdev)
        gb_t = torch.empty_like(ga_t)
        ga_t[lo.value * plane:(lo.value + grid[0]) * plane].copy_(a_t.view(-1))
        del a_t
        torch.cuda.synchronize()
        gdims = _abi.dims_array((grid[0] * world,) + tuple(grid[1:]))
        res = C.c_void_p()
        bufs = (ga_t, gb_t)

        def run(k):
            dev.check(L.mb_iterate_sharded_dev(dev.handle, C.byref(d), gdims, dims, C.c_void_p(ga_t.data_ptr()),
                                               C.c_void_p(gb_t.data_ptr()), k, C.byref(res)))
    else:
        b_t = torch.empty(ogrid, dtype=a_t.dtype, device=env.tdev)
        a = dev.wrap(a_t.data_ptr(), grid, npdt)
        b = dev.wrap(b_t.data_ptr(), ogrid, npdt)
        bufs = (a_t, b_t)
        if name == "gauss7sep":
            d1 = M.mapStencil(border, spec[1], a).desc()
            d2 = M.mapStencil(border, spec[2], a).desc()

            def run(k):
                for _ in range(k):
                    dev.check(L.mb_run_sep2_dev(dev.handle, C.byref(d1), C.byref(d2), dims, C.c_void_p(a.ptr), C.c_void_p(b.ptr), None))
        elif wl["iterated"]:
            d = M.mapStencil(border, spec[1], a).desc()
            res = C.c_void_p()
            if name == "life":
                calls = 50  # 1000 generations take about a millisecond: repeat the call so the clocks can be sampled

            def run(k):
                dev.check(L.mb_iterate_dev(dev.handle, C.byref(d), dims, C.c_void_p(a.ptr), C.c_void_p(b.ptr), k, C.byref(res)))
        else:
            dw = M.mapStencil(border, spec[1], a)
            d = dw.desc(stride)
            assert dw.size(stride) == ogrid, (dw.size(stride), ogrid)

            def run(k):
                for _ in range(k):
                    dev.check(L.mb_run_dev(dev.handle, C.byref(d), dims, C.c_void_p(a.ptr), C.c_void_p(b.ptr)))

    warm = max(warmup, 3)
    if wl["iterated"] and not sharded:
        # the paths for long runs (Life's on-chip generations, the padded pair of an array TMA cannot describe, with its
        # one-time scratch allocation) only exist from a few steps on: warm up THE CALL that is timed
        run(steps)
    run(warm)
    env.barrier()
    l0, x0 = dev.launch_count, dev.aux_launch_count
    sampler = ClockSampler(env.local)
    if rank == 0:
        sampler.start()
        sampler.wait_first()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    env.barrier()
    ev0.record()
    for _ in range(calls):
        run(steps)
    ev1.record()
    env.barrier()
    ms = ev0.elapsed_time(ev1) / calls
    clocks = sampler.stop() if rank == 0 else None
    launches = (dev.launch_count - l0) // calls
    main_launches = launches - (dev.aux_launch_count - x0) // calls  # the stencil kernels proper (helpers return in microseconds)
    kernel = dev.last_kernel
    if world > 1:
        t = torch.tensor([ms], device=env.tdev, dtype=torch.float64)
        env.dist.all_reduce(t, op=env.dist.ReduceOp.MAX)
        ms = float(t.item())
    value = cells * world * steps / (ms * 1e-3) / 1e9
    peak, peak_src = load_peaks()
    per_launch_ms = ms / max(main_launches, 1)
    passes_per_launch = steps / max(main_launches, 1)
    bpc = bytes_per_cell(name, grid)
    alg_bytes = cells * bpc * passes_per_launch
    if name == "gauss7sep" and main_launches == 2 * steps:
        alg_bytes = cells * bpc  # two-pass fallback: each launch moves one read + one write
    achieved = alg_bytes / (per_launch_ms * 1e-3) / 1e9
    traffic = load_traffic(name, kernel, grid) if world == 1 else None
    rec = {
        "value": value, "unit": "Gcells/s", "steps": steps, "warmup": warm, "ms_per_step": ms / steps, "dtype": wl["dtype"],
        "config": {"workload": name, "id": wl["config_id"], "description": wl["text"], "grid_per_gpu": list(grid),
                   "cells": "output cells" if stride else "cells",
                   "l2": "inputs larger than L2" if cells * bpc > 256e6 else "grid is L2-resident by definition of the workload",
                   "kernel": kernel, "parallelism": f"slab{world}" if world > 1 else "single",
                   "halo_transport": L.mb_shard_transport(dev.handle).decode() if sharded else None},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "kernel": kernel,
                     "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": per_launch_ms},
        "gpu_launches": int(launches * calls), "gpu_launches_main": int(main_launches * calls), "clocks": clocks,
    }
    if traffic:
        # the kernel's real DRAM rate: measured traffic of one launch over the launch time measured here
        rec["roofline"]["dram_gbs"] = traffic / (per_launch_ms * 1e-3) / 1e9
        rec["roofline"]["dram_frac"] = rec["roofline"]["dram_gbs"] / peak
    if "published_gcells" in wl:
        rec["vs_baseline"] = value / wl["published_gcells"]
        rec["baseline"] = {"value": wl["published_gcells"], "unit": "Gcells/s", "source": "massiv/README.md:598-632 (7.421 ms, i7-3740QM, 8 threads, GHC 8.8.4)"}
    if name == "gauss7":
        # 49 multiplies + 49 adds per cell, never fused (bit-identical to the reference): the FP32 issue rate is the
        # nearer roof for the direct form
        ops = cells * 98.0 / (per_launch_ms * 1e-3)
        rec["roofline_fp32"] = {"bound": "fp32", "achieved": ops / 1e12, "peak": FP32_OPS_PER_S / 1e12, "unit": "Tops/s (unfused mul + add)",
                                "frac": ops / FP32_OPS_PER_S, "peak_source": "scripts/micro/fp_rate.cu on this pool's B200"}
    if keep is not None:
        keep["bufs"], keep["run"] = bufs, run
    return rec


# ---- verification (outside every timed region) ----------------------------------------------------------

def _nan_aware_equal(got, want):
    if got.tobytes() == want.tobytes():
        return True
    if got.dtype.kind != "f":
        return False
    gn, wn = np.isnan(got), np.isnan(want)
    return bool(np.array_equal(gn, wn) and got[~gn].tobytes() == want[~wn].tobytes())


def verify_heat3d(env, grid, n_steps=4, slab=48):
    """Full size: the two-step kernel, the iterated single-step star kernel and the strict 27-tap kernel must agree
    bit for bit after n_steps; a random `slab`-plane sub-slab of the result must equal the CPU oracle's."""
    torch, M, L, dev, _abi = env.torch, env.M, env.L, env.dev, env.abi
    from oracle import oracle as O
    name = "heat3d"
    npdt = NPDT["f64"]
    border, st = stencil_spec(name, M)
    x_t = synth_device(torch, name, grid, env.tdev)
    a_t, b_t = torch.empty_like(x_t), torch.empty_like(x_t)
    nbytes = x_t.numel() * x_t.element_size()
    a = dev.wrap(a_t.data_ptr(), grid, npdt)
    d = M.mapStencil(border, st, a).desc()
    dims = _abi.dims_array(grid)
    out = {"grid": list(grid), "steps": n_steps}

    def iterate(opts):
        nonlocal n_steps
        for k, v in opts.items():
            dev.set_option(k, v)
        a_t.copy_(x_t)
        res = C.c_void_p()
        dev.check(L.mb_iterate_dev(dev.handle, C.byref(d), dims, C.c_void_p(a_t.data_ptr()), C.c_void_p(b_t.data_ptr()), n_steps, C.byref(res)))
        torch.cuda.synchronize()
        kern = dev.last_kernel
        for k in opts:
            dev.set_option(k, {"tb2": 1, "no_optimistic": 0}[k])
        return (a_t if res.value == a_t.data_ptr() else b_t), kern

    r, k1 = iterate({})
    r1 = r.clone()
    r, k2 = iterate({"tb2": 0})
    out[f"{k1}_vs_{k2}_mismatches"] = env.equal_dev(r1.data_ptr(), r.data_ptr(), nbytes)
    r, k3 = iterate({"tb2": 0, "no_optimistic": 1})
    out[f"{k1}_vs_{k3}_mismatches"] = env.equal_dev(r1.data_ptr(), r.data_ptr(), nbytes)
    out["kernels"] = [k1, k2, k3]
    # oracle: planes [z0, z0 + slab) of the result need n_steps planes of halo each side (the array's own Fill 0
    # border when the slab touches it)
    rng = np.random.default_rng(2020)
    z0 = int(rng.integers(0, max(1, grid[0] - slab)))
    lo, hi = max(0, z0 - n_steps), min(grid[0], z0 + slab + n_steps)
    sub = x_t[lo:hi].cpu().numpy()
    kind, kw = oracle_spec(name)
    want = O.iterate(kind, sub, n_steps, nthreads=ref_threads(), **kw)[z0 - lo:z0 - lo + min(slab, grid[0] - z0)]
    got = r1[z0:z0 + want.shape[0]].cpu().numpy()
    out["oracle_slab"] = {"planes": [z0, z0 + want.shape[0]], "equal": _nan_aware_equal(got, want)}
    out["ok"] = bool(out[f"{k1}_vs_{k2}_mismatches"] == 0 and out[f"{k1}_vs_{k3}_mismatches"] == 0 and out["oracle_slab"]["equal"]
                     and k1 == "vol3d_tb2" and k2 == "vol3d_star7" and k3 == "vol3d_dense27")
    if (grid[2] * 8) % 16 != 0:
        # rows TMA cannot describe: a longer iteration runs on an internal pair with padded rows (mb_api.cu: iterate_pitch);
        # it must give the bits of the same iteration done in place.  (Checksums of the bit patterns: a third full-size
        # copy next to the padded pair does not fit.)
        del r1
        torch.cuda.empty_cache()

        def checksum(t):
            s1 = s2 = 0
            for z in range(0, t.shape[0], 32):  # (in chunks: the temporaries of a whole-array expression do not fit either)
                v = t[z:z + 32].view(torch.int64)
                s1 = (s1 + int(v.sum().item())) & (2 ** 64 - 1)
                s2 = (s2 + int((v ^ (v >> 17)).sum().item())) & (2 ** 64 - 1)
            return s1, s2

        n_steps = 9
        r, _ = iterate({})
        c_padded = checksum(r)
        dev.set_option("no_pitch", 1)
        try:
            r, _ = iterate({})
        finally:
            dev.set_option("no_pitch", 0)
        out["padded_rows_vs_in_place"] = {"steps": n_steps, "checksums_equal": c_padded == checksum(r)}
        out["ok"] = bool(out["ok"] and out["padded_rows_vs_in_place"]["checksums_equal"])
    return out


def verify_life(env, grid, gens=64):
    torch, M, L, dev, _abi = env.torch, env.M, env.L, env.dev, env.abi
    from oracle import oracle as O
    x_t = synth_device(torch, "life", grid, env.tdev)
    a_t, b_t = torch.empty_like(x_t), torch.empty_like(x_t)
    a = dev.wrap(a_t.data_ptr(), grid, np.uint8)
    d = M.mapStencil(M.Wrap, M.lifeStencil, a).desc()
    dims = _abi.dims_array(grid)

    def iterate(bitpack):
        dev.set_option("life_bitpack", bitpack)
        a_t.copy_(x_t)
        res = C.c_void_p()
        dev.check(L.mb_iterate_dev(dev.handle, C.byref(d), dims, C.c_void_p(a_t.data_ptr()), C.c_void_p(b_t.data_ptr()), gens, C.byref(res)))
        torch.cuda.synchronize()
        dev.set_option("life_bitpack", 1)
        return (a_t if res.value == a_t.data_ptr() else b_t), dev.last_kernel

    r, k1 = iterate(1)
    r1 = r.clone()
    r, k2 = iterate(0)
    mism = env.equal_dev(r1.data_ptr(), r.data_ptr(), r1.numel())
    want = O.iterate("life", x_t.cpu().numpy(), gens, size=(3, 3), border="wrap", nthreads=ref_threads())
    eq = r1.cpu().numpy().tobytes() == want.tobytes()
    return {"grid": list(grid), "generations": gens, "kernels": [k1, k2], f"{k1}_vs_{k2}_mismatches": mism, "oracle_full_grid_equal": bool(eq),
            "ok": bool(mism == 0 and eq and k1 == "life_bits")}


def verify_single_pass(env, name, bufs, run, rows=None):
    """Full size: the dispatched kernel against the library's generic kernel (an independent code path) bit for bit
    on the device; a random `rows`-row sub-slab of the output (with the stencil's reach of input rows either side)
    against the CPU oracle."""
    torch, dev = env.torch, env.dev
    from oracle import oracle as O
    a_t, b_t = bufs
    grid = tuple(a_t.shape)
    rows = rows or (4096 if len(grid) == 2 else 48)
    run(1)
    torch.cuda.synchronize()
    k1 = dev.last_kernel
    fast = b_t.clone()
    dev.set_option("force_generic", 1)
    try:
        run(1)
        torch.cuda.synchronize()
        k2 = dev.last_kernel
    finally:
        dev.set_option("force_generic", 0)
    mism = env.equal_dev(fast.data_ptr(), b_t.data_ptr(), fast.numel() * fast.element_size())
    reach = 16  # rows either side: more than the largest window here (9x9) reaches
    rng = np.random.default_rng(2021)
    y0 = int(rng.integers(0, max(1, grid[0] - rows)))
    lo, hi = max(0, y0 - reach), min(grid[0], y0 + rows + reach)
    sub = a_t[lo:hi].cpu().numpy()
    want = oracle_pass(O, name, sub, ref_threads())
    n = min(rows, grid[0] - y0)
    # rows within `reach` of a cut that is not the array's own edge saw the border rule instead of real data
    first = 0 if lo == 0 else reach
    last = want.shape[0] if hi == grid[0] else want.shape[0] - reach
    got = fast[lo + first:lo + last].cpu().numpy()
    eq = _nan_aware_equal(got, want[first:last])
    return {"grid": list(grid), "kernels": [k1, k2 + " (force_generic)"], f"{k1}_vs_generic_mismatches": mism,
            "oracle_slab": {"rows": [lo + first, lo + last], "equal": bool(eq)}, "ok": bool(mism == 0 and eq and k1 != k2)}


def verify_sharded(env):
    """N > 1: the native slab-sharded loop against the single-GPU iteration of the same global array, bit for bit,
    once per halo transport (massiv_b200/selfcheck.py holds the cases; tests/multi_gpu_check.py runs the same)."""
    from massiv_b200 import selfcheck
    return selfcheck.sharded_parity(env.dev, env.dist, env.world, env.rank)


# ---- end to end ------------------------------------------------------------------------------------------

def measure_e2e(env, name, grid, sharded):
    """The same metric through the reference-facing host->host call — mb_run (compute . mapStencil),
    mb_run_sep2, mb_iterate (iterateUntil) or mb_iterate_sharded — on pinned HOST buffers: the H2D copy
    of the input and the D2H copy of the result are inside the timed region.  An iterated workload is
    ONE call of the configured number of generations / steps, as the reference's iterateUntil is."""
    args, dev, M, torch, _abi = env.args, env.dev, env.M, env.torch, env.abi
    world, rank = env.world, env.rank
    dist = env.dist if world > 1 else None
    wl = WORKLOADS[name]
    npdt = NPDT[wl["dtype"]]
    spec = stencil_spec(name, M)
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
    t = synth_device(torch, name, g, env.tdev, first=rank * n)  # fill the pinned input from the device generator
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
    reps = 1 if name == "heat3d" else 5
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        call(steps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if dist is not None:
        tt = torch.tensor([dt], device=env.tdev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
    dev.check(L.mb_host_free(dev.handle, hin))
    dev.check(L.mb_host_free(dev.handle, hout))
    val = float(n_out) * world * passes * reps / dt / 1e9
    return {"value": val, "unit": "Gcells/s", "h2d_bytes_per_step": nbytes / passes, "d2h_bytes_per_step": obytes / passes,
            "grid_per_gpu": list(g), "passes_per_call": passes, "seconds_per_call": dt / reps, "call": cname,
            "timing": "host wall clock around the blocking C call, max over ranks"}


def cpu_baseline(name):
    steps = 2 if name != "life" else 1
    r = run_cpu(name, steps, 1)
    return {"value": r["value"], "unit": "Gcells/s", "cores": r["threads"], "kind": "port", "sample": cpu_sample_text(name, r, steps)}


def gpu_main(args):
    env = Env(args)
    torch = env.torch
    world, rank = env.world, env.rank
    name = args.workload
    wl = WORKLOADS[name]
    grid = tuple(args.grid) if args.grid else wl["grid"]
    sharded = bool((world > 1 or os.environ.get("MASSIV_B200_BENCH_SHARDED")) and name == "heat3d")

    def release():
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
        env.dev.trim()  # the library's scratch (staging arrays, padded pairs) too

    parity = None
    if world > 1 and sharded and not args.no_verify:
        from massiv_b200 import sharding
        sharding.init_comm(env.dev, env.dist)
        parity = verify_sharded(env)

    keep = {}
    head = measure_device(env, name, grid, args.steps, args.warmup, sharded=sharded, keep=keep)
    verify = {}
    if world == 1 and not args.no_verify and not wl["iterated"] and not wl.get("stride"):
        verify[name] = verify_single_pass(env, name, keep["bufs"], keep["run"])
    keep.clear()
    release()
    if world == 1 and not args.no_verify and name == "heat3d" and (not args.grid or min(grid) >= 256):
        verify["heat3d"] = verify_heat3d(env, grid)
        release()
    if world == 1 and not args.no_verify and name == "life":
        verify["life"] = verify_life(env, grid)
    if not args.no_e2e:
        head["e2e"] = measure_e2e(env, name, grid, sharded)
        release()
    else:
        head["e2e"] = None
    head["cpu_baseline"] = cpu_baseline(name) if (rank == 0 and not args.no_cpu) else None

    configs = []
    if world == 1 and name == "heat3d" and not args.grid and args.configs == "all":
        for cname, csteps in CONFIG_RUNS:
            cwl = WORKLOADS[cname]
            keep = {}
            rec = measure_device(env, cname, cwl["grid"], csteps, 3, keep=keep)
            if not args.no_verify:
                if cname == "life":
                    keep.clear()
                    rec["verify"] = verify_life(env, cwl["grid"])
                else:
                    rec["verify"] = verify_single_pass(env, cname, keep["bufs"], keep["run"])
                verify[cname] = {"ok": rec["verify"]["ok"]}
            keep.clear()
            release()
            rec["e2e"] = None if args.no_e2e else measure_e2e(env, cname, cwl["grid"], False)
            release()
            rec["cpu_baseline"] = None if args.no_cpu else cpu_baseline(cname)
            configs.append(rec)

    if rank == 0:
        line = {
            "metric": "stencil_gcells_per_s", "value": head["value"], "unit": "Gcells/s", "n_gpus": world, "steps": args.steps,
            "warmup": head["warmup"], "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": wl["dtype"], "data": "synthetic", "config": head["config"], "roofline": head["roofline"],
            "cpu_baseline": head["cpu_baseline"], "e2e": head["e2e"], "gpu_launches": head["gpu_launches"],
            "gpu_launches_main": head["gpu_launches_main"], "clocks": head["clocks"],
        }
        if "roofline_fp32" in head:
            line["roofline_fp32"] = head["roofline_fp32"]
        if verify:
            verify["all_ok"] = bool(all(v.get("ok") for v in verify.values()))
            line["verify"] = verify
        if parity is not None:
            line["parity"] = parity
        if configs:
            line["configs"] = configs
        print(json.dumps(line), flush=True)
    if world > 1:
        env.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default: 20; the step counts of CONFIG_RUNS for the other workloads, e.g. 1000 generations of Life)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=os.environ.get("MASSIV_B200_WORKLOAD", "heat3d"), choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", type=int, nargs="+", default=None, help="override the grid (debugging only)")
    ap.add_argument("--configs", default="all", choices=["all", "none"],
                    help="all: the default heat3d run at N=1 also measures the other BASELINE configs")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0)
    args = ap.parse_args()
    if args.steps is None:
        # heat3d (the driver's workload): 20; the others: what their record in the default run uses (Life only reaches its
        # on-chip path on long runs: 1000 generations as BASELINE.json's config says)
        args.steps = dict(CONFIG_RUNS).get(args.workload, 20) if args.impl != "reference" else 20
    if args.impl == "reference":
        reference_main(args)
    else:
        gpu_main(args)


if __name__ == "__main__":
    main()
