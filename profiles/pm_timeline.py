#!/usr/bin/env python
"""Time series of one launch from an `ncu --section PmSampling` report (read here, no GPU needed).

    python profiles/pm_timeline.py <report.ncu-rep>

One value per sample (about 1.1 us apart at --pm-sampling-interval 1000); leading / trailing all-zero samples
(before the launch starts, after it ends) are dropped."""
import csv
import io
import subprocess
import sys

WANT = [("TPC.TriageCompute.sm__inst_executed_realtime.avg.pct_of_peak_sustained_elapsed", "instructions executed, % of peak"),
        ("SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_shared.avg", "shared-memory wavefronts per sample per SM"),
        ("FBSP.TriageCompute.dram__read_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM read, % of peak"),
        ("FBSP.TriageCompute.dram__write_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM write, % of peak"),
        ("TPC.TriageCompute.sm__inst_executed_pipe_alu_realtime.avg.pct_of_peak_sustained_elapsed", "ALU pipe, % of peak")]


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv", "--print-metric-instances", "values"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, data = rows[0], rows[2]
    series = {}
    for key, label in WANT:
        if key not in hdr:
            continue
        cell = data[hdr.index(key)]
        inner = cell[cell.index("(") + 1: cell.rindex(")")]
        series[label] = [float(v) for v in inner.split(";") if v.strip()]
    lead = series.get("instructions executed, % of peak", [])
    nz = [i for i, v in enumerate(lead) if v > 0]
    lo, hi = (max(nz[0] - 1, 0), min(nz[-1] + 2, len(lead))) if nz else (0, len(lead))
    print(f"# {sys.argv[1]}: {hi - lo} samples across the launch")
    for label, vals in series.items():
        print(label)
        print("  " + " ".join(f"{v:.0f}" for v in vals[lo:hi]))


if __name__ == "__main__":
    main()
