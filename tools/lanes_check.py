#!/usr/bin/env python
"""Experiment (not a test): throughput of the frame pipeline for different lane counts / SM partitions, with
device-resident frames and with pinned host frames.  Usage: python tools/lanes_check.py W H K frames [configs]"""
import ctypes as C
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import Oracle  # noqa: E402

pkg = importlib.import_module("clusteringsegmentation-1_b200")
W, H, K, FR = (int(x) for x in (sys.argv[1:5] + ["3840", "2160", "256", "24"][len(sys.argv) - 1:]))
CONFIGS = [tuple(int(v) for v in c.split("x")) for c in sys.argv[5].split(",")] if len(sys.argv) > 5 else \
    [(1, 148), (2, 74), (4, 37), (6, 24), (8, 18), (8, 37), (12, 24), (12, 12)]
N = W * H
o = Oracle()
lib = pkg.load_library()
lib.dq_set_display_timings(0)
ring = max(6, -(-200_000_000 // (N * 4)))
SEED0 = int(os.environ.get("LANES_SEED0", "12345"))
host = [torch.from_numpy(o.generate(1, W, H, SEED0 + s).view(np.int32)).pin_memory() for s in range(ring)]
dev = [h.cuda() for h in host]
u32p = C.POINTER(C.c_uint32)
ref = {}


def run(lanes, ctas, on_device):
    pipe = lib.dq_pipeline_create_lanes(0, 0 if on_device else N, lanes, ctas)
    lib.dq_pipeline_set_blocking_wait(pipe, int(os.environ.get("LANES_BLOCKING", "0")))
    if os.environ.get("LANES_PROFILE"):
        lib.dq_context_set_profiling(lib.dq_pipeline_context(pipe), 1)
    nbuf = max(lanes + 2, 4)
    outs = [torch.empty(N, dtype=torch.int32, device="cuda") if on_device else torch.empty(N, dtype=torch.int32).pin_memory()
            for _ in range(nbuf)]
    nks = [C.c_uint32(K) for _ in range(FR)]
    cts = [(C.c_uint32 * K)() for _ in range(FR)]
    for phase, count in ((0, 2 * lanes), (1, FR)):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tickets = []
        for f in range(count):
            nks[f].value = K
            if len(tickets) >= nbuf:  # the out buffer we are about to reuse must be complete
                lib.dq_pipeline_wait(pipe, tickets[f - nbuf])
            if on_device:
                tickets.append(lib.dq_pipeline_submit_device(pipe, N, dev[f % ring].data_ptr(), outs[f % nbuf].data_ptr(),
                                                             C.byref(nks[f]), cts[f], 0))
            else:
                tickets.append(lib.dq_pipeline_submit(pipe, N, C.cast(host[f % ring].data_ptr(), u32p),
                                                      C.cast(outs[f % nbuf].data_ptr(), u32p), C.byref(nks[f]), cts[f], 0))
        lib.dq_pipeline_flush(pipe)
        dt = time.perf_counter() - t0
    ms = lib.dq_pipeline_last_elapsed_ms(pipe)
    key = [(nks[f].value, bytes(cts[f])[: 4 * nks[f].value]) for f in range(FR)]
    last = int(outs[(FR - 1) % nbuf].to(torch.int64).sum().item())
    tag = "dev" if on_device else "host"
    if tag not in ref:
        ref[tag] = (key, last)
    print(f"{tag:4s} lanes={lanes:2d} ctas={ctas:3d}: wall {dt / FR * 1e3:.3f} ms/frame, events {ms / FR:.3f} ms/frame = "
          f"{N * FR / ms / 1e6:.2f} Gpix/s  same={ref[tag] == (key, last)}", flush=True)
    st = pkg.CallStats()
    lib.dq_context_last_stats(lib.dq_pipeline_context(pipe), C.byref(st))
    print("      flagged frames:", lib.dq_pipeline_flagged_frames(pipe), flush=True)
    print("      lane 0, last frame, stage ms:", " ".join(f"{n}={v:.3f}" for n, v in zip(pkg.CallStats.STAGES, st.stage_ms)), flush=True)
    lib.dq_pipeline_destroy(pipe)


for lanes, ctas in CONFIGS:
    run(lanes, ctas, True)
for lanes, ctas in CONFIGS:
    run(lanes, ctas, False)
