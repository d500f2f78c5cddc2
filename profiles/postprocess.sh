#!/bin/bash
# Turns the reports written by profiles/capture.sh (merged back into gpurun_out/) into the committed summaries.
set -e
R=${1:-r1}
python profiles/ncu_summary.py gpurun_out/prof_full.ncu-rep 59968 > profiles/${R}_encode_dequant_ncu_summary.txt
python profiles/line_profile.py gpurun_out/prof_full.ncu-rep dmel_codec_b200/libdmel_b200.so ILi1024ELi8ELi65ELi3E 59968 \
    > profiles/${R}_encode_line_profile.txt
echo "launch list: filter gpurun_out/launches.csv to kernel, grid, block, gpu__time_duration (ns -> us) -> profiles/${R}_launches.csv"
echo "traffic: dram__bytes_read.sum + dram__bytes_write.sum of the fused launch in the summary -> profiles/encode_traffic.json"
