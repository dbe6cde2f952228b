#!/bin/bash
# usage: ncu_top.sh <kernel-regex> <out-name>   (run on the GPU box from the repo root)
# pass 1: launch list of ALL kernels of one bench step; pass 2: --set full on the longest launch matching the regex
K=$1; OUT=$2
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${OUT}_launches.csv python bench.py --profile --steps 1 --warmup 0 > gpurun_out/${OUT}_p1.log 2>&1
IDX=$(python - <<PY
import csv,re
rows=[r for r in csv.reader(open("gpurun_out/${OUT}_launches.csv")) if len(r)>5]
hdr=rows[0]; vi=hdr.index("Metric Value"); ki=hdr.index("Kernel Name")
m=[(float(r[vi].replace(",","")),i) for i,r in enumerate(r for r in rows[1:] if re.search("$K", r[ki]))]
print(max(m)[1])
PY
)
echo "longest launch index among matches: $IDX"
ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip $IDX -c 1 -o gpurun_out/$OUT python bench.py --profile --steps 1 --warmup 0 > gpurun_out/${OUT}_p2.log 2>&1
tail -2 gpurun_out/${OUT}_p2.log
