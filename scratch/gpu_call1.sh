#!/bin/bash
# round-1 late: first GPU contact of the masked group walk (parity with every segment through it, A/B bench, memcheck)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== smoke masked"; HBTU_WALK_MASKED=1 HBTU_WALK_GROUP_MIN=1 timeout 240 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== parity masked (all segments)"; HBTU_WALK_MASKED=1 HBTU_WALK_GROUP_MIN=1 timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -6
for b in 5 4; do
  echo "== bench masked blocks=$b"; HBTU_WALK_MASKED=1 HBTU_WALK_MASKED_BLOCKS=$b timeout 200 python bench.py --profile --steps 2 --warmup 1 2>&1 | tail -1 | python -c "
import sys, json
l = sys.stdin.read().strip()
try:
    d = json.loads(l); print('ms_per_step', d['ms_per_step'], d['config']['phase_ms'], d['clocks'])
except Exception as e: print('ERR', l[-400:])"
done
echo "== bench old group walk"; HBTU_WALK_MASKED=0 timeout 200 python bench.py --profile --steps 2 --warmup 1 2>&1 | tail -1 | python -c "
import sys, json
l = sys.stdin.read().strip()
try:
    d = json.loads(l); print('ms_per_step', d['ms_per_step'], d['config']['phase_ms'], d['clocks'])
except Exception as e: print('ERR', l[-400:])"
echo "== memcheck smoke masked"; HBTU_WALK_MASKED=1 HBTU_WALK_GROUP_MIN=1 timeout 300 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
