#!/bin/bash
# call 21: one prefix popcount per lane (POPC shares the XU pipe) and the predicated-rsqrt variant, cfg 2 + cfg 4 shard
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
S="lib=default lib=tools/ab/lib_sel0.so lib=tools/ab/lib_rsq1.so lib=default"
timeout 900 python tools/ab_walk.py --workload cfg2 --steps 2 --count $S 2>&1 | grep spec | tee gpurun_out/c21_cfg2.jsonl | cut -c1-330
timeout 900 python tools/ab_walk.py --workload cfg4 --particles 1.7e8 --world 8 --steps 2 $S 2>&1 | grep spec | tee gpurun_out/c21_cfg4.jsonl | cut -c1-330
