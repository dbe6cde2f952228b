#!/bin/bash
# GPU call 13: fetch trace at 5e5 subhaloes (cfg4), compute-sanitizer memcheck + racecheck of smoke() (all three walk kernel families)
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
echo "== cfg4 e2e trace"
HBTU_TRACE=1 timeout 600 python bench.py --workload cfg4 --steps 1 --warmup 1 --e2e-steps 2 --profile 2>&1 | grep -E "fetch|done|wave" | tail -12
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
HBTU_TRACE=1 timeout 600 python /tmp/probe4.py 2>&1 | grep -E "fetch|e2e call" | tail -20
echo "== memcheck smoke"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c13_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/c13_memcheck.log
echo "== racecheck smoke"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c13_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/c13_racecheck.log
