#!/bin/bash
# GPU call 11: chunked upload helper (does compute now overlap the upload?), tests, sampled / cfg4 lines after the atomics + search fixes, default bench
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
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
for mb in (32, 32, 8, 128, 512, 100000):
    os.environ["HBTU_UPLOAD_CHUNK_MB"] = str(mb)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = ctx.unbind_batch(e, snap, flags=1, want_energy=False, order_buf=buf)
    dt = time.perf_counter() - t0; st = ctx.stats()
    print(f"chunk {mb} MB: wall {dt*1e3:.1f} ms  stage {st.stage_wall_ms:.1f} execute_wall {st.execute_wall_ms:.1f} execute_gpu {st.execute_ms:.1f} fetch {st.fetch_wall_ms:.1f} upload {st.h2d_ms:.1f}", flush=True)
PY
echo "== e2e probe"; HBTU_TRACE=1 timeout 600 python /tmp/e2e_probe.py 2>&1 | grep -v "rounds so far\|enqueued" | tail -40
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/c11_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/c11_pytest.log
for w in "sampled --max-sample 1000" "cfg4 --workload cfg4" "cfg3 --workload cfg3" "v5 "; do
  set -- $w; tag=$1; shift
  timeout 700 python bench.py "$@" --steps 3 --warmup 2 --dropin-particles 2e7 > gpurun_out/r02_bench_${tag}.json 2> gpurun_out/r02_bench_${tag}.err; echo "== bench $tag rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench_${tag}.json'))
print({k:d[k] for k in ('value','ms_per_step')}, {k:round(v,1) for k,v in d['config']['phase_ms'].items()}, round(d['roofline']['frac'],4), {k:round(v,1) for k,v in d['config'].get('phase_ms_detail',{}).items()})
print('e2e', {k:(round(v,1) if isinstance(v,float) else v) for k,v in d['e2e'].items() if k not in ('drop_in','overlap','api','last_call_breakdown')}, {k:round(v,1) for k,v in d['e2e']['last_call_breakdown'].items()})
print('dropin', {k:(round(v,3) if isinstance(v,float) else v) for k,v in d['e2e']['drop_in'].items() if k in ('particles','seconds','value')}, 'parity ok', d.get('parity',{}).get('ok'), d.get('parity',{}).get('frac_identical_nbound'))
PY
  tail -2 gpurun_out/r02_bench_${tag}.err
done
