"""CPU tests of the multi-GPU host logic with torch.distributed (gloo, world_size 2).

The N>1 path shards frames (no collective) or pixel rows (one exchange of per-shard (colour, count) lists).
What has to hold for the row-sharded result to be GPU-count invariant is additivity of the exact integer
sums; this test exercises the partitioning helpers and the exchange layout across two real processes."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, height, width, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = importlib.import_module("clusteringsegmentation-1_b200")
    from oracle import Oracle
    o = Oracle()
    img = o.generate(1, width, height).reshape(height, width)
    r0, r1 = pkg.rows_for_rank(height, world, rank)
    colours, _, counts = o.calc_color_table(img[r0:r1].ravel())
    # exchange: padded to the largest shard, true lengths gathered first (what the NCCL path does on device)
    n = torch.tensor([colours.size], dtype=torch.int64)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, n)
    cap = int(max(s.item() for s in sizes))
    buf = torch.zeros(2, cap, dtype=torch.int64)
    buf[0, :colours.size] = torch.from_numpy(colours.astype(np.int64))
    buf[1, :counts.size] = torch.from_numpy(counts.astype(np.int64))
    gathered = [torch.zeros(2, cap, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, buf)
    cl = [g[0, :int(s.item())].numpy().astype(np.uint32) for g, s in zip(gathered, sizes)]
    ct = [g[1, :int(s.item())].numpy().astype(np.uint64) for g, s in zip(gathered, sizes)]
    uniq, merged = pkg.merge_histograms(cl, ct)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), uniq=uniq, merged=merged, rows=np.array([r0, r1]))
    dist.barrier()
    dist.destroy_process_group()


def test_row_sharded_histogram_is_shard_count_invariant(tmp_path, oracle, pkg):
    height, width, world = 90, 160, 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, height, width, str(tmp_path)), nprocs=world, join=True)
    img = oracle.generate(1, width, height)
    colours, _, counts = oracle.calc_color_table(img)
    order = np.argsort(colours)
    results = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    rows = sorted(tuple(int(x) for x in res["rows"]) for res in results)
    assert rows[0][0] == 0 and rows[-1][1] == height and rows[0][1] == rows[1][0]
    for res in results:  # every rank ends with the same, exact, global histogram
        assert np.array_equal(res["uniq"], colours[order])
        assert np.array_equal(res["merged"], counts[order].astype(np.uint64))
    assert int(results[0]["merged"].sum()) == height * width


def test_frame_sharding_covers_every_frame_once(pkg):
    for frames in (0, 1, 7, 8, 1024, 1025):
        for world in (1, 2, 4, 8):
            seen = [f for r in range(world) for f in pkg.frames_for_rank(frames, world, r)]
            assert seen == list(range(frames))
            sizes = [len(pkg.frames_for_rank(frames, world, r)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
