#!/bin/bash
# GPU call: validation of the reference-style cell moments (tests + bench parity block), e2e breakdown + drop-in row, A/B of the walk variants with the new
# deciding loop, bench lines of cfg3 / cfg5 / sampled mode
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
rm -f gpurun_out/parity_stats.jsonl
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/c06_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/c06_pytest.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
echo "== bench (default)"
timeout 900 python bench.py > gpurun_out/r02_bench_v2.json 2> gpurun_out/r02_bench_v2.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_v2.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['config']['phase_ms'], d['roofline']['frac'])
print('e2e', {k:v for k,v in d['e2e'].items() if k!='drop_in'})
print('dropin', d['e2e'].get('drop_in'))
print('parity', d.get('parity'))
PY
tail -3 gpurun_out/r02_bench_v2.err
echo "== A/B cfg2"
timeout 700 python tools/ab_walk.py --count --steps 2 \
  lib=default,walk_masked_pairs=1,walk_masked_blocks=7 \
  lib=default,walk_masked_pairs=1,walk_masked_blocks=8 \
  lib=default,walk_masked_pairs=2,walk_masked_blocks=5 \
  lib=tools/ab/lib_d4.so,walk_masked_pairs=1,walk_masked_blocks=7 \
  lib=tools/ab/lib_d4.so,walk_masked_pairs=1,walk_masked_blocks=8 \
  lib=tools/ab/lib_d4.so,walk_masked_pairs=2,walk_masked_blocks=5 \
  lib=tools/ab/lib_big.so,walk_masked_pairs=1,walk_masked_blocks=7 \
  lib=tools/ab/lib_u2.so,walk_masked_pairs=1,walk_masked_blocks=8 \
  > gpurun_out/c06_ab_cfg2.jsonl 2> gpurun_out/c06_ab_cfg2.err; echo "ab rc=$?"; cut -c1-230 gpurun_out/c06_ab_cfg2.jsonl
echo "== bench cfg5 / cfg3 / sampled"
timeout 600 python bench.py --workload cfg5 --steps 5 --warmup 3 --dropin-particles 2.1e7 > gpurun_out/r02_bench_cfg5.json 2> gpurun_out/r02_bench_cfg5.err; echo "cfg5 rc=$?"; tail -c 600 gpurun_out/r02_bench_cfg5.json; tail -2 gpurun_out/r02_bench_cfg5.err
timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 --dropin-particles 8.9e6 > gpurun_out/r02_bench_cfg3.json 2> gpurun_out/r02_bench_cfg3.err; echo "cfg3 rc=$?"; tail -c 600 gpurun_out/r02_bench_cfg3.json; tail -2 gpurun_out/r02_bench_cfg3.err
timeout 600 python bench.py --max-sample 1000 --steps 3 --warmup 3 --dropin-particles 1e6 > gpurun_out/r02_bench_sampled.json 2> gpurun_out/r02_bench_sampled.err; echo "sampled rc=$?"; tail -c 600 gpurun_out/r02_bench_sampled.json; tail -2 gpurun_out/r02_bench_sampled.err
ls -la gpurun_out | head -40
