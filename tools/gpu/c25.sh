#!/bin/bash
# call 25: pipelined hbtu_unbind_batch (parts of a many-hierarchy batch upload behind each other's kernels): tests, cfg4 / cfg3 lines,
# drop-in row with the huge-page hint
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_dropin.py -q -m gpu -x 2>&1 | tail -8
for w in cfg4 cfg3; do
  timeout 900 python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/c25_bench_$w.json 2> gpurun_out/c25_bench_$w.err; echo "$w rc=$?"
done
python - <<'PY'
import json
for w in ("cfg4", "cfg3"):
    d = json.loads(open(f"gpurun_out/c25_bench_{w}.json").read().strip().splitlines()[-1])
    p = d.get("parity") or {}
    print(w, "value %.4g ms %.1f" % (d["value"], d["ms_per_step"]), "e2e %.4g ms %.1f" % (d["e2e"]["value"], d["e2e"]["ms_per_step"]), d["e2e"]["last_call_breakdown"],
          "parity", p.get("frac_identical_nbound"), p.get("jaccard_misses"))
PY
cat > /tmp/dropin.py <<'PY'
import os, sys, json
sys.path.insert(0, os.getcwd())
import torch
import bench
wl = bench.WORKLOADS["cfg2"]; dev = torch.device("cuda", 0)
csnap, desc = wl.cpu_sample(1.8e8, 3e5)
row = bench.bench_dropin_row(wl, 4e7, dev, csnap, os.cpu_count())
print("DROPIN", json.dumps({k: v for k, v in row.items() if k != "reference"}))
PY
HBT_B200_TRACE=1 timeout 900 python /tmp/dropin.py 2>&1 | grep -E "DROPIN|hbt_b200|Error|error" | tail -6
