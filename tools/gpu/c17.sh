#!/bin/bash
# call 17: drop-in row after the unpack fix + ncu launch list of one bench step of the current library
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
cat > /tmp/dropin.py <<'PY'
import os, sys, json
sys.path.insert(0, os.getcwd())
import torch
import bench
wl = bench.WORKLOADS["cfg2"]; dev = torch.device("cuda", 0)
csnap, desc = wl.cpu_sample(1.8e8, 3e6)
class A: pass
row = bench.bench_dropin_row(wl, 4e7, dev, csnap, os.cpu_count())
print("DROPIN", json.dumps(row))
PY
HBT_B200_TRACE=1 timeout 900 python /tmp/dropin.py 2>&1 | grep -E "DROPIN|hbt_b200|Error|error" | tail -8
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c17_launches.csv python bench.py --profile --steps 1 --warmup 0 > gpurun_out/c17_p1.log 2>&1
python - <<'PY'
import csv, re, collections
rows = [r for r in csv.reader(open("gpurun_out/c17_launches.csv")) if len(r) > 5]
hdr = rows[0]; vi = hdr.index("Metric Value"); ki = hdr.index("Kernel Name"); ui = hdr.index("Metric Unit")
agg = collections.OrderedDict()
tot = 0.0
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v if u in ("ms", "msecond") else v * 1e3
    k = r[ki][:110]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += ms; tot += ms
with open("gpurun_out/c17_launches_summary.md", "w") as f:
    f.write(f"{len(rows)-1} launches, {tot:.1f} ms of kernel time\n\n| share | launches | ms | kernel |\n|---|---|---|---|\n")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| {100*ms/tot:.2f} % | {n} | {ms:.2f} | `{k}` |\n")
print(open("gpurun_out/c17_launches_summary.md").read()[:6000])
PY
rm -f gpurun_out/c17_launches.csv
