#!/usr/bin/env python
"""Times the brute-force remap kernel (N*K evaluations) and the table path on one synthetic frame.
usage: python tools/bench_map.py [width height K reps]"""
import ctypes as C
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import Oracle, muted  # generator + checker

w, h, k, reps = [int(x) for x in (sys.argv[1:5] + ["3840", "2160", "256", "10"][len(sys.argv) - 1:])]
pkg = importlib.import_module("clusteringsegmentation-1_b200")
lib = pkg.load_library()
lib.dq_set_display_timings(0)
ctx = lib.dq_context_create(0)
stream = torch.cuda.ExternalStream(lib.dq_context_stream(ctx))
o = Oracle()
host = o.generate(1, w, h)
with muted():
    pal, _ = o.quant_varpart_fast(host[:: 16].copy(), k)
px = torch.from_numpy(host.view(np.int32)).cuda()
out = torch.empty_like(px)
ct = pal.copy()
ctp = ct.ctypes.data_as(C.POINTER(C.c_uint32))
ref = o.map_colors_mps(host, pal)
for prefer in (0, 1):
    lib.dq_map_colors_device(ctx, px.data_ptr(), px.numel(), out.data_ptr(), ctp, ct.size, prefer)
    ok = np.array_equal(out.cpu().numpy().view(np.uint32), ref)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(reps):
        lib.dq_map_colors_device(ctx, px.data_ptr(), px.numel(), out.data_ptr(), ctp, ct.size, prefer)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    evals = px.numel() * ct.size
    t_pipe = evals * 4 / (148 * 128 * 1.965e9) * 1e3
    print(f"{'table' if prefer else 'brute'}: {ms:.3f} ms/frame ({px.numel()/ms/1e3:.0f} Mpix/s) exact={ok} K={ct.size} "
          f"pipe-roofline(N*K*4)={t_pipe:.3f} ms frac={t_pipe/ms:.2f}", flush=True)
