#!/usr/bin/env python
"""Summarise an ncu raw CSV page (ncu -i X.ncu-rep --page raw --csv) into the handful of numbers
profiles/*.md quote.  usage: ncu -i rep --page raw --csv | python tools/ncu_summary.py [pattern ...]"""
import csv
import re
import sys

rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
default = [r"gpu__time_duration\.sum", r"^dram__bytes_(read|write)\.sum$", r"launch__registers_per_thread", r"launch__grid_size",
           r"launch__block_size", r"sm__warps_active\.avg\.pct_of_peak_sustained_active", r"smsp__inst_executed\.sum$",
           r"smsp__issue_active\.avg\.pct", r"sm__throughput\.avg\.pct", r"gpu__dram_throughput\.avg\.pct",
           r"lts__t_sector_hit_rate\.pct", r"sm__inst_executed_pipe_(fp64|fma|alu|lsu|xu|fmaheavy|fmalite)\.sum$",
           r"smsp__average_warps?_issue_stalled_.*_per_issue_active", r"sm__cycles_elapsed\.max", r"smsp__cycles_active\.avg$",
           r"lts__t_requests_srcunit_tex_op_(atom|red)\.sum$", r"l1tex__data_bank_conflicts_pipe_lsu\.sum$"]
pats = [re.compile(p) for p in (sys.argv[1:] or default)]
for row in rows[2:]:
    print("== kernel:", row[hdr.index("Kernel Name")], "grid", row[hdr.index("Grid Size")], "block", row[hdr.index("Block Size")])
    for h, u, v in zip(hdr, units, row):
        if any(p.search(h) for p in pats):
            try:
                if float(v.replace(",", "")) == 0 and "stalled" in h:
                    continue
            except ValueError:
                pass
            print(f"  {h} [{u}] = {v}")
