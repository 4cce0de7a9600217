"""Run under torchrun on >= 2 GPUs: the native slab-sharded iteration must equal the single-GPU iteration of the
global array, bit for bit, for both halo transports (peer-memory copies and ncclSend/ncclRecv).  The cases live in
massiv_b200/selfcheck.py (bench.py runs the same at N > 1).  Prints MULTI_GPU_OK on rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import massiv_b200 as M  # noqa: E402
from massiv_b200 import selfcheck, sharding  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = M.Device(local)
    sharding.init_comm(dev, dist)
    res = selfcheck.sharded_parity(dev, dist, world, rank, verbose=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(res), flush=True)
        print("MULTI_GPU_OK" if res["parity_ok"] else "MULTI_GPU_FAIL", flush=True)
    sys.exit(0 if res["parity_ok"] else 1)


if __name__ == "__main__":
    main()
