#!/bin/bash
# call 23: validation of the final library: full GPU suite, smoke(), the default bench line, the other configurations
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== bench default"
timeout 1200 python bench.py > gpurun_out/c23_bench_default.json 2> gpurun_out/c23_bench_default.err; echo "rc=$?"
echo "== bench --impl reference"
timeout 600 python bench.py --impl reference > gpurun_out/c23_bench_reference.json 2>> gpurun_out/c23_bench_default.err; echo "rc=$?"
for w in cfg3 cfg4 cfg5; do
  echo "== bench $w"
  timeout 900 python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/c23_bench_$w.json 2> gpurun_out/c23_bench_$w.err; echo "rc=$?"
done
echo "== bench sampled"
timeout 900 python bench.py --max-sample 1000 --steps 5 --warmup 3 > gpurun_out/c23_bench_sampled.json 2> gpurun_out/c23_bench_sampled.err; echo "rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c23_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as ex:
        print(f, "unreadable", ex); continue
    if d.get("impl") == "reference":
        print(f, d.get("value"), d.get("unit"), d.get("cpu_baseline")); continue
    p = d.get("parity") or {}
    print(f, "value %.4g" % d["value"], "ms %.1f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], "frac %.3f" % d["roofline"]["frac"], d["config"].get("phase_ms"),
          "parity nbound", p.get("frac_identical_nbound"), "jaccard misses", p.get("jaccard_misses"), "dropin", (d["e2e"].get("drop_in") or {}).get("value"))
PY
