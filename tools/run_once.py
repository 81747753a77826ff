#!/usr/bin/env python
"""Runs quant_recurse_device a few times on one synthetic frame (for ncu / sanitizer captures).
usage: python tools/run_once.py [width height K reps kind seed]"""
import ctypes as C
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import Oracle  # input generator only

w, h, k, reps, kind, seed = [int(x) for x in (sys.argv[1:7] + ["3840", "2160", "256", "3", "1", "12345"][len(sys.argv) - 1:])]
pkg = importlib.import_module("clusteringsegmentation-1_b200")
lib = pkg.load_library()
lib.dq_set_display_timings(0)
ctx = lib.dq_context_create(0)
px = torch.from_numpy(Oracle().generate(kind, w, h, seed).view(np.int32)).cuda()
out = torch.empty_like(px)
ct = np.zeros(k, np.uint32)
nk = C.c_uint32(k)
lib.dq_context_set_profiling(ctx, 1)
st = pkg.CallStats()
for i in range(reps):
    nk.value = k
    lib.dq_quant_recurse_device(ctx, px.numel(), px.data_ptr(), out.data_ptr(), C.byref(nk), ct.ctypes.data_as(C.POINTER(C.c_uint32)), 0)
    lib.dq_context_last_stats(ctx, C.byref(st))
    print(i, nk.value, st.as_dict(), flush=True)
torch.cuda.synchronize()
if os.environ.get("DQ_TRACE"):
    lib.dq_debug_split_timeline(ctx, 1, None, 0)
    nk.value = k
    lib.dq_quant_recurse_device(ctx, px.numel(), px.data_ptr(), out.data_ptr(), C.byref(nk), ct.ctypes.data_as(C.POINTER(C.c_uint32)), 0)
    buf = (C.c_uint64 * (2 * 4096))()
    n = lib.dq_debug_split_timeline(ctx, 0, buf, 4096)
    names = {1: "round", 2: "phaseA", 3: "phaseB", 4: "phaseC", 5: "ctlbar", 6: "pass", 7: "part", 8: "root", 9: "begin", 10: "collected", 11: "reduced", 12: "rootbar"}
    prev = None
    agg = {}
    for i in range(n):
        tag, arg, clk = buf[2 * i] >> 32, buf[2 * i] & 0xFFFFFFFF, buf[2 * i + 1]
        dt = 0 if prev is None else (clk - prev) / 1.965e3
        prev = clk
        agg[names[tag]] = agg.get(names[tag], 0.0) + dt
        if os.environ.get("DQ_TRACE") == "2":
            print(f"{names[tag]:7s} {arg:5d} +{dt:8.2f} us")
    print("timeline totals (us):", {k2: round(v, 1) for k2, v in agg.items()}, "sum", round(sum(agg.values()), 1))
