"""Self-checks of the multi-GPU path that need nothing but the library itself.

:func:`sharded_parity` runs the native slab-sharded loop (``mb_iterate_sharded_dev``) against the single-GPU
iteration (``mb_iterate_dev``) of the same global array and compares bit for bit, once per halo transport.
``bench.py`` prints its verdict at N > 1 and ``tests/multi_gpu_check.py`` asserts it.
Reference semantics of the loop: ``iterateUntil`` / ``iterateLoop``, Array/Manifest/Internal.hs:331-397.
"""
from __future__ import annotations

import numpy as np

from . import sharding
from . import stencil as M


def _same(got: np.ndarray, want: np.ndarray) -> bool:
    """bit-identical, NaN payloads excepted (different producers of a NaN may choose different payloads)"""
    if got.shape != want.shape:
        return False
    if got.tobytes() == want.tobytes():
        return True
    if got.dtype.kind != "f":
        return False
    gn, wn = np.isnan(got), np.isnan(want)
    return bool(np.array_equal(gn, wn) and got[~gn].tobytes() == want[~wn].tobytes())


def sharded_cases(world: int):
    """(name, border, stencil, global array, steps): fused and single-step rounds, thin slabs, Wrap rings, Reflect
    copies, a 2-D fold with a one-sided halo, a shifted (non-mapStencil) padding, Life, and promise-free runs with
    Inf / NaN on, next to and two planes away from slab boundaries."""
    rng = np.random.default_rng(5)
    k = np.zeros((3, 3, 3))
    k[1, 1, 1] = 0.4
    for ix in ((0, 1, 1), (2, 1, 1), (1, 0, 1), (1, 2, 1), (1, 1, 0), (1, 1, 2)):
        k[ix] = 0.1
    conv = M.makeConvolutionStencilFromKernel
    cases = [("heat fill0", M.Fill(0.0), conv(k).assumeFinite(), rng.random((24 * world + 3, 40, 72)), 7),
             ("heat fused even", M.Fill(0.0), conv(k).assumeFinite(), rng.random((40 * world + 1, 64, 128)), 6),
             ("heat thin slabs", M.Fill(0.0), conv(k).assumeFinite(), rng.random((2 * world + 1, 40, 72)), 5),
             ("heat f32 fused", M.Fill(0.0), conv(k.astype(np.float32)).assumeFinite(), rng.random((16 * world, 48, 128)).astype(np.float32), 4),
             ("heat strict", M.Fill(0.0), conv(k), rng.random((16 * world, 33, 65)), 4),
             ("heat strict inf", M.Fill(0.0), conv(k), "inf", 4),
             ("strict fused", M.Fill(0.0), conv(k), rng.random((24 * world + 2, 64, 128)), 6),
             ("strict fused inf", M.Fill(0.0), conv(k), "inf128", 5),
             ("strict ghost hi", M.Fill(0.0), conv(k), "one@16", 4),
             ("strict ghost lo", M.Fill(0.0), conv(k), "one@15", 4),
             ("strict ghost far", M.Fill(0.0), conv(k), "one@17", 2),
             ("heat odd pitch", M.Fill(0.0), conv(k), rng.random((12 * world, 35, 67)), 4),
             ("heat wrap", M.Wrap, conv(k).assumeFinite(), rng.random((20 * world, 32, 64)), 5),
             ("heat reflect", M.Reflect, conv(rng.standard_normal((3, 3, 3))), rng.random((12 * world, 32, 64)), 3),
             ("avg2d edge", M.Edge, M.avgStencil(M.Sz(3, 3)), rng.random((64 * world + 5, 256)), 6),
             ("sum shifted", M.Padding((2, 0), (0, 2), M.Fill(0)), M.sumStencil(M.Sz(3, 3)),
              rng.integers(-99, 99, (24 * world + 1, 200)).astype(np.int64), 4),
             ("life wrap", M.Wrap, M.lifeStencil, (rng.random((32 * world, 128)) < 0.4).astype(np.uint8), 9)]
    built = []
    for name, border, st, arr, n in cases:
        if isinstance(arr, str):  # non-finite values on and next to the slab boundaries, no finite promise
            spec = arr
            arr = rng.random((16 * world, 33, 65) if spec == "inf" else (16 * world, 64, 128))
            if spec.startswith("one@"):  # a single value one rank owns and its neighbour only sees as a ghost
                arr[int(spec[4:]), 10, 10] = np.inf
            else:
                for r in range(world):
                    arr[16 * r, 5 + r, 7], arr[16 * r + 15, 20, 40 + r], arr[16 * r + 8, 0, 0] = np.inf, np.nan, -np.inf
        built.append((name, border, st, arr, n))
    return built


def sharded_parity(dev: M.Device, dist, world: int, rank: int, verbose: bool = False) -> dict:
    """{"p2p": {...}, "nccl": {...}, "parity_ok": bool}; `sharding.init_comm(dev, dist)` must have been called.
    Collective: every rank must call it."""
    import torch
    L = dev._L
    tdev = torch.device("cuda", dev.ordinal)
    built = sharded_cases(world)
    result = {}
    for transport, no_p2p in (("p2p", 0), ("nccl", 1)):
        dev.set_option("no_p2p", no_p2p)
        bad, used = [], set()
        for name, border, st, arr, n in built:
            padded = isinstance(border, M.Padding)
            dw = M.applyStencil(border, st, arr) if padded else M.mapStencil(border, st, arr)
            sp = sharding.shard_plan(dw.desc(), arr.shape, world, rank)
            mine = np.ascontiguousarray(arr[sp.first_plane:sp.first_plane + sp.local_planes])
            got = sharding.iterate_sharded(n, border, st, dev.fromHost(mine), arr.shape).toHost()
            used.add(L.mb_shard_transport(dev.handle).decode())
            whole = M.iterateStencil(n, border, st, dev.fromHost(arr)).toHost()
            ok = _same(got, whole[sp.first_plane:sp.first_plane + sp.local_planes])
            flag = torch.tensor([1 if ok else 0], device=tdev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if verbose and rank == 0:
                print(f"{name:18s} world={world} halo via {sorted(used)[-1]:4s} {'ok' if flag.item() else 'MISMATCH'}", flush=True)
            if not flag.item():
                bad.append(name)
        result[transport] = {"parity_ok": not bad, "cases": len(built), "mismatching": bad, "transport_reported": sorted(used)}
    dev.set_option("no_p2p", 0)
    result["parity_ok"] = bool(result["p2p"]["parity_ok"] and result["nccl"]["parity_ok"])
    return result
