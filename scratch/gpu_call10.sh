#!/bin/bash
export PYTHONUNBUFFERED=1 HBTU_WALK_MASKED=1 HBTU_WALK_MASKED_BLOCKS=5
for v in A B C E G; do
  echo "== variant $v"; HBTU_LIB_PATH=$PWD/scratch/ab/lib_$v.so timeout 200 python bench.py --profile --steps 2 --warmup 1 2>&1 | tail -1 | python -c "
import sys, json
l = sys.stdin.read().strip()
try:
    d = json.loads(l); print('ms_per_step', d['ms_per_step'], d['config']['phase_ms'])
except Exception as e: print('ERR', l[-400:])"
done
