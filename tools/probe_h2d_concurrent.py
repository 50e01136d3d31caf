"""Concurrent pinned host->device copy ceiling of one box (the bound of the raw-raster e2e path at N ranks).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/probe_h2d_concurrent.py

Every rank owns one GPU and copies a pinned 1.67 GB buffer (one 64-page A4 batch) to it, all ranks at the same time
(barrier before every round); per-rank and aggregate GB/s, alone (rank by rank) and together; optional D2H of the result
size at the same time (what run_host_stream does)."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
NB = 64 * 3508 * 2480 * 3
ND = 166_625_792
src = torch.empty(NB, dtype=torch.uint8, pin_memory=True)
src.fill_(rank + 1)
dst = torch.empty(NB, dtype=torch.uint8, device=dev)
back_d = torch.empty(ND, dtype=torch.uint8, device=dev)
back_h = torch.empty(ND, dtype=torch.uint8, pin_memory=True)
s2 = torch.cuda.Stream(dev)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def run(reps, with_d2h, only_rank=None):
    barrier()
    if only_rank is not None and rank != only_rank:
        barrier()
        return 0.0
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
        if with_d2h:
            with torch.cuda.stream(s2):
                back_h.copy_(back_d, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    if only_rank is not None:
        barrier()
    return NB * reps / (a.elapsed_time(b) * 1e-3) / 1e9


def gather(v):
    if world == 1:
        return [v]
    t = torch.tensor([v], dtype=torch.float64, device=dev)
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [float(x.item()) for x in out]


run(2, False)
alone = []
for r in range(world):
    v = run(4, False, only_rank=r)
    alone.append(max(gather(v)))
together = gather(run(8, False))
together_d2h = gather(run(8, True))
if rank == 0:
    print(json.dumps({"ranks": world, "bytes_per_copy": NB,
                      "h2d_GBps_each_rank_alone": [round(x, 1) for x in alone],
                      "h2d_GBps_per_rank_all_together": [round(x, 1) for x in together],
                      "h2d_GBps_aggregate_all_together": round(sum(together), 1),
                      "h2d_GBps_per_rank_with_concurrent_d2h": [round(x, 1) for x in together_d2h],
                      "h2d_GBps_aggregate_with_concurrent_d2h": round(sum(together_d2h), 1),
                      "ms_per_64_page_batch_per_rank_all_together": round(NB / (min(together) * 1e9) * 1e3, 1),
                      "host_cores": os.cpu_count()}))
if world > 1:
    dist.destroy_process_group()
