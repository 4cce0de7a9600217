"""Run under torchrun on >= 2 GPUs: the native slab-sharded iteration (NCCL halo exchange) must equal
the single-GPU iteration of the global array, bit for bit.  Prints MULTI_GPU_OK on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import massiv_b200 as M  # noqa: E402
from massiv_b200 import sharding  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = M.Device(local)
    sharding.init_comm(dev, dist)
    rng = np.random.default_rng(5)
    k = np.zeros((3, 3, 3))
    k[1, 1, 1] = 0.4
    for ix in ((0, 1, 1), (2, 1, 1), (1, 0, 1), (1, 2, 1), (1, 1, 0), (1, 1, 2)):
        k[ix] = 0.1
    cases = [("heat fill0", M.Fill(0.0), M.makeConvolutionStencilFromKernel(k).assumeFinite(), rng.random((24 * world + 3, 40, 72)), 7),
             ("heat fused even", M.Fill(0.0), M.makeConvolutionStencilFromKernel(k).assumeFinite(), rng.random((40 * world + 1, 64, 128)), 6),
             ("heat thin slabs", M.Fill(0.0), M.makeConvolutionStencilFromKernel(k).assumeFinite(), rng.random((2 * world + 1, 40, 72)), 5),
             ("heat f32 fused", M.Fill(0.0), M.makeConvolutionStencilFromKernel(k.astype(np.float32)).assumeFinite(),
              rng.random((16 * world, 48, 128)).astype(np.float32), 4),
             ("heat strict", M.Fill(0.0), M.makeConvolutionStencilFromKernel(k), rng.random((16 * world, 33, 65)), 4),
             ("heat strict inf", M.Fill(0.0), M.makeConvolutionStencilFromKernel(k), "inf", 4),
             ("strict fused", M.Fill(0.0), M.makeConvolutionStencilFromKernel(k), rng.random((24 * world + 2, 64, 128)), 6),
             ("strict fused inf", M.Fill(0.0), M.makeConvolutionStencilFromKernel(k), "inf128", 5),
             ("strict ghost hi", M.Fill(0.0), M.makeConvolutionStencilFromKernel(k), "one@16", 4),
             ("strict ghost lo", M.Fill(0.0), M.makeConvolutionStencilFromKernel(k), "one@15", 4),
             ("strict ghost far", M.Fill(0.0), M.makeConvolutionStencilFromKernel(k), "one@17", 2),
             ("heat wrap", M.Wrap, M.makeConvolutionStencilFromKernel(k).assumeFinite(), rng.random((20 * world, 32, 64)), 5),
             ("heat reflect", M.Reflect, M.makeConvolutionStencilFromKernel(rng.standard_normal((3, 3, 3))), rng.random((12 * world, 32, 64)), 3),
             ("avg2d edge", M.Edge, M.avgStencil(M.Sz(3, 3)), rng.random((64 * world + 5, 256)), 6),
             ("life wrap", M.Wrap, M.lifeStencil, (rng.random((32 * world, 128)) < 0.4).astype(np.uint8), 9)]
    ok = True
    for name, border, st, arr, n in cases:
        if isinstance(arr, str):  # non-finite values on and next to the slab boundaries, no finite promise
            arr_spec = arr
            arr = rng.random((16 * world, 33, 65) if arr == "inf" else (16 * world, 64, 128))
            if arr_spec.startswith("one@"):   # a single value that one rank owns and its neighbour only sees as a ghost
                arr[int(arr_spec[4:]), 10, 10] = np.inf
            else:
                for r in range(world):
                    arr[16 * r, 5 + r, 7], arr[16 * r + 15, 20, 40 + r], arr[16 * r + 8, 0, 0] = np.inf, np.nan, -np.inf
        desc = M.mapStencil(border, st, arr).desc()
        sp = sharding.shard_plan(desc, arr.shape, world, rank)
        mine = np.ascontiguousarray(arr[sp.first_plane:sp.first_plane + sp.local_planes])
        got = sharding.iterate_sharded(n, border, st, dev.fromHost(mine), arr.shape).toHost()
        want = M.iterateStencil(n, border, st, dev.fromHost(arr)).toHost()[sp.first_plane:sp.first_plane + sp.local_planes]
        same = got.tobytes() == want.tobytes() or (np.array_equal(np.isnan(got), np.isnan(want)) and
                                                   np.array_equal(got[~np.isnan(got)], want[~np.isnan(want)]))
        flags = torch.tensor([1 if same else 0], device="cuda")
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        if rank == 0:
            how = dev._L.mb_shard_transport(dev.handle).decode()
            print(f"{name:16s} world={world} halo via {how:4s} {'ok' if flags.item() else 'MISMATCH'}", flush=True)
        ok = ok and bool(flags.item())
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_OK" if ok else "MULTI_GPU_FAIL", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
