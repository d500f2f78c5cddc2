#!/usr/bin/env python
"""Condense ncu's launch list (--metrics gpu__time_duration.sum --csv) to kernel, grid, block, duration in us."""
import csv
import re
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]
ki, gi, bi, vi, ui = (h.index(k) for k in ("Kernel Name", "Grid Size", "Block Size", "Metric Value", "Metric Unit"))
print("kernel,grid,block,gpu__time_duration_us")
for r in rows[1:]:
    name = re.sub(r"\(dmel::FusedParams\)|dmel::|void ", "", r[ki])[:90]
    v = float(r[vi].replace(",", ""))
    us = v / 1e3 if r[ui].startswith("ns") else (v if r[ui].startswith("us") else v * 1e3)
    print(f'"{name}","{r[gi]}","{r[bi]}",{us:.2f}')
