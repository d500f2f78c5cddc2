#!/bin/bash
# ncu captures of the stand-alone kernels (row peak, FSQ encode, anti-aliased Snake) at the sizes bench.py's
# secondary.kernels uses.  Run on a GPU box from the repo root (gpurun --timeout 1200 -- 'bash profiles/capture_standalone.sh'),
# then here:  bash profiles/postprocess_standalone.sh r2
set -e
mkdir -p gpurun_out
python -c "
import json, sys, torch
sys.path.insert(0, '.')
import bench
dev = torch.device('cuda:0'); torch.cuda.set_device(dev)
print(json.dumps(bench.standalone_kernels(dev)))" > gpurun_out/kern_plain.log 2>&1    # the plain run first
for k in row_absmax fsq_encode antialias_snake; do
  ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:"$k" -s 3 -c 1 \
      -o gpurun_out/prof_final_$k -f python -c "
import sys, torch
sys.path.insert(0, '.')
import bench
dev = torch.device('cuda:0'); torch.cuda.set_device(dev)
bench.standalone_kernels(dev)" > gpurun_out/ncu_final_$k.log 2>&1
done
ls -la gpurun_out/prof_final_*.ncu-rep
