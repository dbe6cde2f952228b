#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 HBTU_WALK_MASKED=1 HBTU_WALK_MASKED_BLOCKS=7
timeout 900 bash scratch/ncu_top.sh walk_masked_kernel r01_walk_masked_top
ls -la gpurun_out/
ncu -i gpurun_out/r01_walk_masked_top.ncu-rep --page raw --csv > gpurun_out/r01_walk_masked_top_raw.csv 2>/dev/null
ncu -i gpurun_out/r01_walk_masked_top.ncu-rep --page source --csv --print-source sass > gpurun_out/r01_walk_masked_top_sass.csv 2>/dev/null
ncu -i gpurun_out/r01_walk_masked_top.ncu-rep --page source --csv > gpurun_out/r01_walk_masked_top_src.csv 2>/dev/null
ls -la gpurun_out/
