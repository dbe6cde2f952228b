#!/bin/bash
# GPU call 2 (round 2): whole GPU test suite (parity stats), A/B sweep of the walk variants, small-N workloads
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
rm -f gpurun_out/parity_stats.jsonl
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/c02_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/c02_pytest.log
echo "== A/B cfg2"
timeout 900 python tools/ab_walk.py --count --steps 2 \
  lib=default,walk_masked_pairs=2,walk_masked_blocks=5 \
  lib=default,walk_masked_pairs=2,walk_masked_blocks=6 \
  lib=default,walk_masked_pairs=1,walk_masked_blocks=6 \
  lib=default,walk_masked_pairs=1,walk_masked_blocks=7 \
  lib=default,walk_masked_pairs=1,walk_masked_blocks=8 \
  lib=default,walk_masked_pairs=1,walk_masked_blocks=9 \
  lib=tools/ab/lib_u2.so,walk_masked_pairs=1,walk_masked_blocks=7 \
  lib=tools/ab/lib_u2.so,walk_masked_pairs=1,walk_masked_blocks=8 \
  lib=tools/ab/lib_u2.so,walk_masked_pairs=1,walk_masked_blocks=9 \
  lib=tools/ab/lib_u2.so,walk_masked_pairs=2,walk_masked_blocks=5 \
  lib=tools/ab/lib_u2.so,walk_masked_pairs=2,walk_masked_blocks=6 \
  lib=tools/ab/lib_big.so,walk_masked_pairs=2,walk_masked_blocks=5 \
  lib=tools/ab/lib_big.so,walk_masked_pairs=1,walk_masked_blocks=7 \
  lib=tools/ab/lib_big.so,walk_masked_pairs=1,walk_masked_blocks=8 \
  lib=tools/ab/lib_u2big.so,walk_masked_pairs=1,walk_masked_blocks=8 \
  lib=tools/ab/lib_d4.so,walk_masked_pairs=2,walk_masked_blocks=5 \
  lib=tools/ab/lib_d4.so,walk_masked_pairs=1,walk_masked_blocks=8 \
  lib=default,walk_masked_pairs=2,walk_masked_blocks=5,walk_small_max=512 \
  lib=default,walk_masked_pairs=2,walk_masked_blocks=5,walk_small_max=2048 \
  > gpurun_out/c02_ab_cfg2.jsonl 2> gpurun_out/c02_ab_cfg2.err; echo "ab rc=$?"; cat gpurun_out/c02_ab_cfg2.jsonl | cut -c1-260
echo "== A/B cfg5 (small-subhalo path)"
timeout 400 python tools/ab_walk.py --workload cfg5 --particles 2.1e7 --count --steps 2 \
  lib=default,walk_small_max=0 lib=default,walk_small_max=256 lib=default,walk_small_max=1024 lib=default,walk_small_max=4096 \
  > gpurun_out/c02_ab_cfg5.jsonl 2> gpurun_out/c02_ab_cfg5.err; echo "ab5 rc=$?"; cat gpurun_out/c02_ab_cfg5.jsonl | cut -c1-260
echo "== A/B cfg3"
timeout 400 python tools/ab_walk.py --workload cfg3 --particles 8.9e6 --count --steps 2 \
  lib=default,walk_small_max=0 lib=default,walk_small_max=256 lib=default,walk_small_max=1024 lib=default,walk_small_max=4096 \
  > gpurun_out/c02_ab_cfg3.jsonl 2> gpurun_out/c02_ab_cfg3.err; echo "ab3 rc=$?"; cat gpurun_out/c02_ab_cfg3.jsonl | cut -c1-260
tail -3 gpurun_out/*.err
