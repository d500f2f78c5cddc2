#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed):  python profiles/ncu_summary.py <report> [frames_per_launch]

Prints the headline metrics of each captured launch and the dynamic SASS
instruction mix (per frame when frames_per_launch is given)."""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_issued.avg.per_cycle_active", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__maximum_warps_per_active_cycle_pct"]


def ncu_csv(report, page, extra=()):
    out = subprocess.run(["ncu", "-i", report, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def traffic(report):
    """dram__bytes_read.sum + dram__bytes_write.sum of the fused launch, as the JSON bench.py cites in roofline.traffic"""
    import json
    rows = ncu_csv(report, "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in data:
        if "dmel_fused_kernel" in r[name_i]:
            rd = float(r[hdr.index("dram__bytes_read.sum")].replace(",", "")) * scale[units[hdr.index("dram__bytes_read.sum")]]
            wr = float(r[hdr.index("dram__bytes_write.sum")].replace(",", "")) * scale[units[hdr.index("dram__bytes_write.sum")]]
            print(json.dumps({"kernel": r[name_i][:80], "dram_bytes_read": int(rd), "dram_bytes_written": int(wr),
                              "dram_bytes_per_launch": int(rd + wr),
                              "source": "ncu --set full --clock-control none, one launch of the benchmark step (profiles/capture.sh); "
                                        "not measured in the bench run. Writes still in L2 when the launch ends are not counted"}, indent=1))
            return


def main():
    report = sys.argv[1]
    if len(sys.argv) > 2 and sys.argv[2] == "--traffic":
        return traffic(report)
    frames = float(sys.argv[2]) if len(sys.argv) > 2 else None
    rows = ncu_csv(report, "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    for r in data:
        print("==", r[name_i][:100])
        for k in KEYS:
            if k in hdr:
                print(f"   {k:75s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
                v = float(r[i].replace(",", ""))
                if v >= 0.05:
                    print(f"   stall {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:40s} {v:6.2f}")
    rows = ncu_csv(report, "source", ["--print-source", "sass"])
    heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    if not heads:
        return
    # the SASS table of the first dmel_fused_kernel launch if there is one, else of the first launch
    pick = 0
    for n, hi in enumerate(heads):
        if hi > 0 and any("dmel_fused_kernel" in c for c in rows[hi - 1]):
            pick = n
            break
    h = rows[heads[pick]]
    body = rows[heads[pick] + 1: heads[pick + 1] - 1 if len(heads) > pick + 1 else None]
    ci, si, srci = h.index("Instructions Executed"), h.index("# Samples"), h.index("Source")
    tot, samp, total = collections.Counter(), collections.Counter(), 0
    for r in body:
        try:
            n, s = int(r[ci]), int(r[si])
        except (ValueError, IndexError):
            continue
        op = re.sub(r"^@!?U?P\w+\s+", "", r[srci].strip()).split()[0].split(".")[0]
        tot[op] += n
        samp[op] += s
        total += n
    div = frames or 1.0
    print(f"-- dynamic SASS mix of the fused launch: {total} warp instructions" + (f", {total / div:.1f} per frame" if frames else ""))
    for op, n in tot.most_common(28):
        print(f"   {op:12s} {n / div:12.1f}   stall samples {samp[op]}")


if __name__ == "__main__":
    main()
