"""Diagnostic (run under torchrun on the GPU box): latency of the one collective of this path -- an all-gather of a
fixed-capacity (hash, key) int64 buffer per rank -- for a few capacities.  Prints one line on rank 0."""
import os
import sys
import time

import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
world, rank = dist.get_world_size(), dist.get_rank()
res = {}
for cap in (256, 2048, 16000):
    buf = torch.full((cap, 2), -1, dtype=torch.int64, device="cuda")
    out = torch.empty((world * cap, 2), dtype=torch.int64, device="cuda")
    for _ in range(5):
        dist.all_gather_into_tensor(out, buf)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        dist.all_gather_into_tensor(out, buf)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 20], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res[cap] = float(t.item())
if rank == 0:
    print("all_gather_into_tensor latency ms per call (max over ranks), capacity -> ms:", res, "world", world, file=sys.stderr)
    print(res)
dist.destroy_process_group()
