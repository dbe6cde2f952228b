#!/bin/bash
# call 20: CTAs/SM of the masked walk on the periodic workloads (cfg 2 chose 7), with the new group_min default
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
S="lib=default,walk_masked_blocks=7 lib=default,walk_masked_blocks=6 lib=default,walk_masked_blocks=8 lib=default,walk_masked_blocks=7,walk_masked_pairs=2,walk_masked_blocks=5"
timeout 900 python tools/ab_walk.py --workload cfg4 --particles 1.7e8 --world 8 --steps 2 --count $S 2>&1 | grep spec | tee gpurun_out/c20_cfg4.jsonl | cut -c1-330
timeout 600 python tools/ab_walk.py --workload cfg3 --particles 8.9e6 --steps 3 --count $S 2>&1 | grep spec | tee gpurun_out/c20_cfg3.jsonl | cut -c1-330
