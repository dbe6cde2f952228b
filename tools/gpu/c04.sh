#!/bin/bash
# GPU call 4: remaining GPU tests, ncu evidence (full-size largest launch of the shipped kernel + the 64-target variant), full bench
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
echo "== pytest -m gpu (all)"; timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/c04_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/c04_pytest.log
echo "== ncu full, largest masked launch of a bench step (NP=2, 5 CTAs/SM)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:walk_masked_kernel --launch-skip 37 -c 1 -o gpurun_out/r02_masked_np2b5_full python bench.py --profile --steps 1 --warmup 0 > gpurun_out/c04_ncu1.log 2>&1; echo "ncu1 rc=$?"; tail -2 gpurun_out/c04_ncu1.log
echo "== ncu full, same launch, 64-target groups (NP=1, 7 CTAs/SM)"
HBTU_WALK_MASKED_PAIRS=1 HBTU_WALK_MASKED_BLOCKS=7 timeout 900 ncu --set full --clock-control none --import-source on -k regex:walk_masked_kernel --launch-skip 37 -c 1 -o gpurun_out/r02_masked_np1b7_full python bench.py --profile --steps 1 --warmup 0 > gpurun_out/c04_ncu2.log 2>&1; echo "ncu2 rc=$?"; tail -2 gpurun_out/c04_ncu2.log
echo "== bench (default)"
timeout 900 python bench.py > gpurun_out/r02_bench_v1.json 2> gpurun_out/r02_bench_v1.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r02_bench_v1.json; tail -5 gpurun_out/r02_bench_v1.err
ls -la gpurun_out | head -30
