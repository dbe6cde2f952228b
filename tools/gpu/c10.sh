#!/bin/bash
# GPU call 10: trace of the end-to-end call (where does the upload serialise?), tests after the pinned readbacks / sampled walk order, sampled + cfg4 bench lines
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
echo "== trace"
HBTU_TRACE=1 timeout 600 python bench.py --steps 1 --warmup 1 --e2e-steps 1 --profile > gpurun_out/c10_trace.json 2> gpurun_out/c10_trace.err; echo "rc=$?"
grep -n "hbtu" gpurun_out/c10_trace.err | tail -45
cat > /tmp/e2e_probe.py <<'PY'
import os, sys, time
sys.path.insert(0, os.getcwd())
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch, numpy as np
import bench
from hbtplus_b200 import capi
from hbtplus_b200.unbind import UnbindContext
wl = bench.WORKLOADS["cfg2"]; dev = torch.device("cuda", 0)
snap = wl.make(1.8e8, dev, 0); torch.cuda.empty_cache()
ctx = UnbindContext(wl.params(0)); e = capi.make_epoch(1.0)
cap = capi.order_capacity(snap.part_offset, snap.nest_offset, snap.nest_list)
buf = torch.empty(cap, dtype=torch.int32, pin_memory=True).numpy()
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = ctx.unbind_batch(e, snap, flags=1, want_energy=False, order_buf=buf)
    dt = time.perf_counter() - t0; st = ctx.stats()
    print(f"e2e call {i}: wall {dt*1e3:.1f} ms  stage {st.stage_wall_ms:.1f} execute_wall {st.execute_wall_ms:.1f} execute_gpu {st.execute_ms:.1f} fetch {st.fetch_wall_ms:.1f} h2d {st.h2d_ms:.1f}", flush=True)
PY
echo "== e2e probe with trace"; HBTU_TRACE=1 timeout 600 python /tmp/e2e_probe.py 2>&1 | grep -v "rounds so far" | tail -30
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/c10_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/c10_pytest.log
echo "== bench sampled"
timeout 600 python bench.py --max-sample 1000 --steps 3 --warmup 2 --dropin-particles 1e6 > gpurun_out/r02_bench_sampled_v2.json 2> gpurun_out/r02_bench_sampled_v2.err; echo "sampled rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_sampled_v2.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['config']['phase_ms'], d['roofline']['frac'], d['config'].get('phase_ms_detail'))
print('parity ok', d.get('parity',{}).get('ok'), d.get('parity',{}).get('frac_identical_nbound'))
PY
tail -3 gpurun_out/r02_bench_sampled_v2.err
echo "== bench cfg4 (one GPU = one eighth-sized snapshot)"
timeout 600 python bench.py --workload cfg4 --steps 3 --warmup 2 --dropin-particles 1e6 > gpurun_out/r02_bench_cfg4_n1.json 2> gpurun_out/r02_bench_cfg4_n1.err; echo "cfg4 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_cfg4_n1.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['config']['phase_ms'], d['roofline']['frac'], d['config'].get('phase_ms_detail'), d['config']['subhaloes_per_gpu'], d['config']['rounds'])
print('e2e', {k:v for k,v in d['e2e'].items() if k not in ('drop_in','overlap','api')})
print('parity ok', d.get('parity',{}).get('ok'), d.get('parity',{}).get('frac_identical_nbound'))
PY
tail -3 gpurun_out/r02_bench_cfg4_n1.err
