#!/usr/bin/env python
"""Floor of north_star's OTHER row-sharding variant ("an all-reduce of K x (3 sums + count) per k-means iteration"), measured
instead of estimated.  In that variant every rank keeps its own shard's colours and the split runs level-synchronously:
each of the ~ 8 rounds x 11 passes of a K=256 call ends in one all-reduce of the per-job partial sums (at most
256 jobs x 8 u64 = 16 KB), and the next pass cannot start before it returns.  This script times exactly that dependency
chain and nothing else: `passes` times { a one-block kernel that touches the buffer ; ncclAllReduce(16 KB, u64 sum) },
enqueued back to back on one stream without any host wait -- the time the variant spends in its exchange even with
kernels that take no time at all.  The in-library all-gather form
(dq_rows_*, tools/rows_check.py) does its ONE exchange up front and then runs the single-GPU split kernel on every rank.
usage: torchrun --nproc-per-node N tools/rows_allreduce_floor.py [passes words reps]"""
import json
import os
import sys

import torch
import torch.distributed as dist

passes, words, reps = [int(x) for x in (sys.argv[1:4] + ["88", "2048", "20"][len(sys.argv) - 1:])]
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
buf = torch.zeros(words, dtype=torch.int64, device="cuda")


def chain():
    for _ in range(passes):
        buf.add_(1)              # stands for the pass's kernels: depends on the previous all-reduce, feeds the next
        dist.all_reduce(buf)


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


eager_ms = timed(chain)
if rank == 0:
    print(json.dumps({"variant": "allreduce_per_pass_floor", "n_gpus": world, "passes": passes, "bytes_per_allreduce": words * 8,
                      "ms_per_image": round(eager_ms, 4), "us_per_pass": round(eager_ms * 1e3 / passes, 2)}), flush=True)
dist.destroy_process_group()
