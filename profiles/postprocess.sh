#!/bin/bash
# Turns the reports written by profiles/capture.sh (merged back into gpurun_out/) into the committed summaries.
#   bash profiles/postprocess.sh r2
set -e
R=${1:-r2}
python profiles/ncu_summary.py gpurun_out/prof_full.ncu-rep 59968 > profiles/${R}_encode_dequant_ncu_summary.txt
# the library holds one cubin per translation unit: the line profile reads the variant's object file
python profiles/line_profile.py gpurun_out/prof_full.ncu-rep dmel_codec_b200/_obj/libdmel_b200/fused_1024_8_3.o ILi1024ELi8ELi65ELi3E 59968 \
    > profiles/${R}_encode_line_profile.txt
python profiles/launch_list.py gpurun_out/launches.csv > profiles/${R}_launches.csv
python profiles/pm_timeline.py gpurun_out/prof_pm.ncu-rep > profiles/${R}_encode_pm_timeline.txt || true
python profiles/ncu_summary.py gpurun_out/prof_full.ncu-rep --traffic > profiles/encode_traffic.json
