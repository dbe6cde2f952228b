#!/bin/bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 280 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
echo "== bench"; timeout 240 python bench.py > gpurun_out/r01_bench_v8.json 2> gpurun_out/bench_v8.err; tail -c 3000 gpurun_out/r01_bench_v8.json; tail -3 gpurun_out/bench_v8.err
