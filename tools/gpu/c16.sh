#!/bin/bash
# call 16: hash-table particle query + shim timing breakdown + threaded fetch conversion (cfg4)
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_idtable.py tests/test_gpu_dropin.py -q -m gpu -x 2>&1 | tail -5
HBT_B200_TRACE=1 timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/c16_bench.json 2> gpurun_out/c16_bench.err
echo "bench rc=$?"
grep "hbt_b200" gpurun_out/c16_bench.err | tail -8
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c16_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
print("drop_in", json.dumps(d["e2e"].get("drop_in", {}).get("last_call_ms")), d["e2e"].get("drop_in", {}).get("value"))
print("query", json.dumps(d["config"]["next_rows"]["particle_query"]))
print("parity", json.dumps(d.get("parity"))[:600])
PY
cat > /tmp/probe4.py <<'PY'
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import bench
from hbtplus_b200 import capi
from hbtplus_b200.unbind import UnbindContext
wl = bench.WORKLOADS["cfg4"]; dev = torch.device("cuda", 0)
snap = wl.make(1.7e8, dev, 0, 8); torch.cuda.empty_cache()
ctx = UnbindContext(wl.params(0)); e = capi.make_epoch(1.0)
cap = capi.order_capacity(snap.part_offset, snap.nest_offset, snap.nest_list)
buf = torch.empty(cap, dtype=torch.int32, pin_memory=True).numpy()
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = ctx.unbind_batch(e, snap, flags=1, want_energy=False, order_buf=buf)
    dt = time.perf_counter() - t0; st = ctx.stats()
    print(f"cfg4 e2e call {i}: wall {dt*1e3:.1f} ms  stage {st.stage_wall_ms:.1f} execute_wall {st.execute_wall_ms:.1f} fetch {st.fetch_wall_ms:.1f} upload {st.h2d_ms:.1f}", flush=True)
PY
HBTU_TRACE=1 timeout 600 python /tmp/probe4.py 2>&1 | grep -E "fetch|e2e call" | tail -12
