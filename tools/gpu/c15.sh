#!/bin/bash
export PYTHONUNBUFFERED=1
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
