"""Per-GPU host-to-device bandwidth, one rank at a time and all ranks together (which GPUs share the slow path?).
usage: torchrun --nproc-per-node N tools/h2d_probe.py"""
import os
import time

import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl")
n = 256 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h.fill_(rank)
d = torch.empty(n, dtype=torch.uint8, device="cuda")


def rate(reps=8):
    d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    return reps * n / (time.perf_counter() - t0) / 1e9


alone = torch.zeros(world, device="cuda")
for r in range(world):
    if world > 1:
        dist.barrier()
    if r == rank:
        alone[r] = rate()
together = torch.zeros(world, device="cuda")
if world > 1:
    dist.barrier()
together[rank] = rate(16)
if world > 1:
    dist.all_reduce(alone)
    dist.all_reduce(together)
if rank == 0:
    print("alone    GB/s:", [round(float(x), 1) for x in alone])
    print("together GB/s:", [round(float(x), 1) for x in together], "sum", round(float(together.sum()), 1))
if world > 1:
    dist.destroy_process_group()
