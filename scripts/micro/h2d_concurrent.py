"""How fast do pinned host <-> device copies run when every rank of the box copies at once?
torchrun --nproc-per-node N scripts/micro/h2d_concurrent.py   (prints GB/s per rank: all ranks together, then rank 0 alone)"""
import os
import time

import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl")
n = 4 << 30
host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
host.fill_(1)
devt = torch.empty(n, dtype=torch.uint8, device="cuda")


def bar():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def run(active, h2d):
    bar()
    t0 = time.perf_counter()
    if active:
        for _ in range(3):
            if h2d:
                devt.copy_(host, non_blocking=True)
            else:
                host.copy_(devt, non_blocking=True)
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    bar()
    return 3 * n / dt / 1e9 if active else 0.0


for label, act in (("all ranks", True), ("rank 0 alone", rank == 0)):
    for h2d in (True, False):
        run(act, h2d)  # warm
        g = run(act, h2d)
        vals = [torch.zeros(1, device="cuda") for _ in range(world)]
        me = torch.tensor([g], device="cuda", dtype=torch.float32)
        if world > 1:
            dist.all_gather(vals, me)
        else:
            vals = [me]
        if rank == 0:
            print(f"{label:13s} {'H2D' if h2d else 'D2H'}: " + " ".join(f"{float(v):5.1f}" for v in vals) + " GB/s per rank", flush=True)
if world > 1:
    dist.destroy_process_group()
