#!/bin/bash
export PYTHONUNBUFFERED=1
for b in 5 6 7; do HBTU_WALK_MASKED=1 HBTU_WALK_MASKED_BLOCKS=$b timeout 200 python scratch/gpu_count.py 2>&1 | tail -1; done
HBTU_WALK_MASKED=0 timeout 200 python scratch/gpu_count.py 2>&1 | tail -1
