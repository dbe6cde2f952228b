#!/bin/bash
export PYTHONUNBUFFERED=1
for v in D4 U8; do
  ms=$(HBTU_LIB_PATH=$PWD/scratch/ab/lib_$v.so timeout 50 python bench.py --profile --steps 2 --warmup 1 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read().strip()); print(d['config']['phase_ms']['walk'])
except Exception as e: print(99999)")
  echo "variant $v walk_ms $ms"
done
