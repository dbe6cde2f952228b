#!/bin/bash
# call 19: walk_group_min sweep, lower end, with the fallback counters
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
S="lib=default,walk_group_min=512 lib=default,walk_group_min=256 lib=default,walk_group_min=128 lib=default,walk_group_min=64"
timeout 600 python tools/ab_walk.py --workload cfg3 --particles 8.9e6 --steps 3 --count $S 2>&1 | grep spec | tee gpurun_out/c19_cfg3.jsonl | cut -c1-330
timeout 600 python tools/ab_walk.py --workload cfg5 --particles 2.1e7 --steps 3 --count $S 2>&1 | grep spec | tee gpurun_out/c19_cfg5.jsonl | cut -c1-330
timeout 900 python tools/ab_walk.py --workload cfg2 --steps 2 --count lib=default,walk_group_min=512 lib=default,walk_group_min=128 2>&1 | grep spec | tee gpurun_out/c19_cfg2.jsonl | cut -c1-330
timeout 900 python tools/ab_walk.py --workload cfg4 --particles 1.7e8 --world 8 --steps 2 --count lib=default,walk_group_min=8192 lib=default,walk_group_min=512 lib=default,walk_group_min=128 2>&1 | grep spec | tee gpurun_out/c19_cfg4.jsonl | cut -c1-330
