#!/bin/bash
# GPU call 8: validation of the table ring + pairwise exact elements (tests), A/B, bench with e2e breakdown
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/c08_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/c08_pytest.log
echo "== A/B cfg2"
timeout 400 python tools/ab_walk.py --count --steps 2 \
  lib=default,walk_masked_pairs=1,walk_masked_blocks=7 \
  lib=default,walk_masked_pairs=1,walk_masked_blocks=6 \
  lib=default,walk_masked_pairs=2,walk_masked_blocks=5 \
  > gpurun_out/c08_ab_cfg2.jsonl 2> gpurun_out/c08_ab_cfg2.err; echo "ab rc=$?"; cut -c1-230 gpurun_out/c08_ab_cfg2.jsonl
echo "== bench (default)"
timeout 900 python bench.py > gpurun_out/r02_bench_v4.json 2> gpurun_out/r02_bench_v4.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_v4.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['config']['phase_ms'], d['roofline']['frac'])
print('e2e', {k:v for k,v in d['e2e'].items() if k!='drop_in'})
print('dropin', d['e2e'].get('drop_in'))
print('parity ok', d.get('parity',{}).get('ok'), d.get('parity',{}).get('frac_identical_nbound'))
PY
tail -3 gpurun_out/r02_bench_v4.err
