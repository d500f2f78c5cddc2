#!/bin/bash
# Regenerates the round's ncu evidence.  Run on a GPU box from the repo root (e.g. gpurun --timeout 900 -- 'bash profiles/capture.sh'),
# then post-process here (no GPU needed) with profiles/postprocess.sh.  Follows /opt/skills/guides/B200_PROFILING.md:
# the command runs once without ncu first, clocks are not touched, nothing printed under ncu is a bench value.
set -e
CMD="python bench.py --steps 5 --warmup 3 --quick"
mkdir -p gpurun_out
$CMD > gpurun_out/plain.log 2>&1
# 1. launch list: every launch of our kernels with its duration
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"dmel|quantize|minmax" -c 400 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
# 2. full capture of the step's kernel (fused forward, MODE 65) and of the stand-alone dequantise kernel
ncu --set full --import-source on --clock-control none --kernel-name-base demangled \
    -k regex:"65, \(int\)3|dequantize_kernel" -s 4 -c 2 -o gpurun_out/prof_full -f $CMD > gpurun_out/ncu_full.log 2>&1
# 3. time series of one launch (ramp / steady state / tail)
ncu --section PmSampling --pm-sampling-interval 1000 --clock-control none --kernel-name-base demangled \
    -k regex:"65, \(int\)3" -s 6 -c 1 -o gpurun_out/prof_pm -f $CMD > gpurun_out/ncu_pm.log 2>&1
ls -la gpurun_out/launches.csv gpurun_out/prof_full.ncu-rep gpurun_out/prof_pm.ncu-rep
