#!/bin/bash
# call 28: two-part pipeline (1/4, 3/4): pipeline tests + the cfg4 line
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pipeline.py -q -m gpu -x 2>&1 | tail -3
timeout 900 python bench.py --workload cfg4 --steps 3 --warmup 3 > gpurun_out/c28_bench_cfg4.json 2> gpurun_out/c28_bench_cfg4.err; echo "cfg4 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c28_bench_cfg4.json").read().strip().splitlines()[-1])
p = d.get("parity") or {}
print("cfg4 value %.4g ms %.1f" % (d["value"], d["ms_per_step"]), "e2e %.4g ms %.1f" % (d["e2e"]["value"], d["e2e"]["ms_per_step"]), d["e2e"]["last_call_breakdown"], "parity", p.get("frac_identical_nbound"), p.get("jaccard_misses"))
PY
