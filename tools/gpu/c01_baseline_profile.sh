#!/bin/bash
# GPU call 1 (round 2): baseline bench of the round-1 kernel on this pool + ncu evidence for the SHIPPED masked walk
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt
timeout 400 python bench.py > gpurun_out/r02_bench_v0.json 2> gpurun_out/r02_bench_v0.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r02_bench_v0.json
timeout 900 bash tools/ncu_top.sh walk_masked_kernel r02_walk_masked_v4_top; echo "ncu rc=$?"
ls -la gpurun_out
