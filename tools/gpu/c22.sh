#!/bin/bash
# call 22: a larger L1 for the node loads of the masked walk (smaller chain stack: 196 / 164 KB shared memory carve-outs)
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
S="lib=default lib=tools/ab/lib_l196.so lib=tools/ab/lib_l164.so"
timeout 900 python tools/ab_walk.py --workload cfg2 --steps 2 --count $S 2>&1 | grep spec | tee gpurun_out/c22_cfg2.jsonl | cut -c1-330
timeout 900 python tools/ab_walk.py --workload cfg4 --particles 1.7e8 --world 8 --steps 2 --count $S 2>&1 | grep spec | tee gpurun_out/c22_cfg4.jsonl | cut -c1-330
