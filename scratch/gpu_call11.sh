#!/bin/bash
export PYTHONUNBUFFERED=1
best=A; bestms=1093.3
for v in S SU U; do
  ms=$(HBTU_LIB_PATH=$PWD/scratch/ab/lib_$v.so timeout 60 python bench.py --profile --steps 2 --warmup 1 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read().strip()); print(d['config']['phase_ms']['walk'])
except Exception as e: print(99999)")
  echo "variant $v walk_ms $ms"
  if python -c "import sys; sys.exit(0 if float('$ms') < float('$bestms') else 1)"; then best=$v; bestms=$ms; fi
done
echo "best $best $bestms"
if [ "$best" != "A" ]; then HBTU_LIB_PATH=$PWD/scratch/ab/lib_$best.so HBTU_WALK_GROUP_MIN=1 timeout 60 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2; fi
