#!/usr/bin/env python
"""Attribute the per-instruction counters of an ncu report to CUDA source lines.

    python profiles/line_profile.py <report.ncu-rep> <lib.so> <mangled-kernel-substring> [frames]

ncu's CSV export carries per-SASS-instruction counts but no line numbers; nvdisasm -g
carries line numbers.  Both list the kernel's instructions in the same order, so zip them.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def main():
    report, lib, kern = sys.argv[1:4]
    frames = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
    out = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    pick = 0  # the SASS table of the first dmel_fused_kernel launch if there is one
    for n, hi in enumerate(heads):
        if hi > 0 and any("dmel_fused_kernel" in c for c in rows[hi - 1]):
            pick = n
            break
    h = rows[heads[pick]]
    body = [r for r in rows[heads[pick] + 1: heads[pick + 1] - 1 if len(heads) > pick + 1 else None] if len(r) == len(h)]
    ci, si, wi = h.index("Instructions Executed"), h.index("# Samples"), h.index("L1 Wavefronts Shared")
    wx = h.index("L1 Wavefronts Shared Excessive")
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=td, capture_output=True)
        # one cubin per translation unit of the library: disassemble them all, the kernel filter below picks
        dis = "".join(subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, f)], capture_output=True, text=True).stdout
                      for f in sorted(os.listdir(td)) if f.endswith(".cubin"))
    lines, cur, active = [], None, False
    for ln in dis.splitlines():
        if ln.startswith("//---") and ".text." in ln:
            active = kern in ln
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            lines.append(cur)
    if len(lines) != len(body):
        print(f"warning: {len(lines)} disassembled instructions vs {len(body)} profiled", file=sys.stderr)
    agg = collections.defaultdict(lambda: [0, 0, 0, 0])
    for loc, r in zip(lines, body):
        a = agg[loc]
        a[0] += int(r[ci]); a[1] += int(r[si]); a[2] += int(r[wi] or 0); a[3] += int(r[wx] or 0)
    total = sum(a[0] for a in agg.values())
    print(f"{'file:line':32s} {'inst/frame':>10s} {'%':>6s} {'samples':>8s} {'smem wf/frame':>14s} {'excess':>8s}")
    for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(os.environ.get("LP_LINES", "45"))]:
        name = f"{loc[0]}:{loc[1]}" if loc else "?"
        print(f"{name:32s} {a[0] / frames:10.1f} {100 * a[0] / total:6.1f} {a[1]:8d} {a[2] / frames:14.1f} {a[3] / frames:8.1f}")


if __name__ == "__main__":
    main()
