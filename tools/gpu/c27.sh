#!/bin/bash
# call 27: host planning of a round on several threads (run_round prologue / epilogue, per-subhalo init, table uploads):
# full GPU suite, cfg4 / cfg3 lines, cfg2 step as a regression check
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for w in cfg4 cfg3; do
  timeout 900 python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/c27_bench_$w.json 2> gpurun_out/c27_bench_$w.err; echo "$w rc=$?"
done
python - <<'PY'
import json
for w in ("cfg4", "cfg3"):
    d = json.loads(open(f"gpurun_out/c27_bench_{w}.json").read().strip().splitlines()[-1])
    p = d.get("parity") or {}
    print(w, "value %.4g ms %.1f" % (d["value"], d["ms_per_step"]), "e2e %.4g ms %.1f" % (d["e2e"]["value"], d["e2e"]["ms_per_step"]), d["e2e"]["last_call_breakdown"],
          d["config"]["phase_ms"], "parity", p.get("frac_identical_nbound"), p.get("jaccard_misses"))
PY
timeout 600 python tools/ab_walk.py --workload cfg2 --steps 2 lib=default 2>&1 | grep spec | cut -c1-300
