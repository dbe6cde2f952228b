#!/bin/bash
# call 18: where does the masked group walk start to pay? (walk_group_min sweep on three workloads)
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
S="lib=default,walk_group_min=8192 lib=default,walk_group_min=4096 lib=default,walk_group_min=2048 lib=default,walk_group_min=1024 lib=default,walk_group_min=512"
timeout 600 python tools/ab_walk.py --workload cfg3 --particles 8.9e6 --steps 3 $S 2>&1 | grep spec | tee gpurun_out/c18_cfg3.jsonl | cut -c1-260
timeout 600 python tools/ab_walk.py --workload cfg5 --particles 2.1e7 --steps 3 $S 2>&1 | grep spec | tee gpurun_out/c18_cfg5.jsonl | cut -c1-260
timeout 900 python tools/ab_walk.py --workload cfg2 --steps 2 lib=default,walk_group_min=8192 lib=default,walk_group_min=2048 lib=default,walk_group_min=1024 2>&1 | grep spec | tee gpurun_out/c18_cfg2.jsonl | cut -c1-260
