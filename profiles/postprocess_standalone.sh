#!/bin/bash
# bash profiles/postprocess_standalone.sh r2   (reads gpurun_out/prof_final_*.ncu-rep written by capture_standalone.sh)
set -e
R=${1:-r2}
OUT=profiles/${R}_standalone_kernels_ncu_summary.txt
{
  echo "# ncu --set full --clock-control none, one launch each, sizes of bench.py secondary.kernels (profiles/capture_standalone.sh)."
  echo "# 'before' = the first round-2 versions (one block per tile), kept for the comparison DESIGN.md 4.2-4.6 draws."
  for k in row_absmax fsq_encode antialias_snake; do
    echo; echo "######## $k: final"
    python profiles/kernel_keys.py gpurun_out/prof_final_$k.ncu-rep $k
  done
  if [ -f gpurun_out/prof_kern.ncu-rep ]; then
    echo; echo "######## before: row_absmax (one 256-thread block per 64 KB chunk), fsq_encode (run-time level count, divisions, XU conversions)"
    python profiles/kernel_keys.py gpurun_out/prof_kern.ncu-rep
  fi
  if [ -f gpurun_out/prof_act.ncu-rep ]; then
    echo; echo "######## before: antialias_snake staged in shared memory, packed pairs, one CTA per 1008 outputs"
    python profiles/kernel_keys.py gpurun_out/prof_act.ncu-rep
  fi
} > $OUT
wc -l $OUT
