"""scratch: counted pass of the bench workload: interactions, node-parallel iterations, masked-walk fallbacks"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from hbtplus_b200 import capi, synth
from hbtplus_b200.unbind import UnbindContext

dev = torch.device("cuda", 0)
sizes, parent = bench.workload_sizes(1.8e8, bench.SEED)
snap = synth.make_snapshot_torch(sizes, device=dev, seed=bench.SEED, box_size=bench.BOX, particle_mass=1e-6, parent=parent,
                                 centre=[bench.BOX / 2] * 3, wrap=False, pin=True)
ctx = UnbindContext(bench.params_for(0))
e = capi.make_epoch(1.0)
ctx.stage(e, snap, capi.HBTU_FLAG_TRUNCATE_SOURCE)
ctx.set_counting(True)
ctx.execute()
st = ctx.stats()
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("HBTU_")}, "walk_ms": st.walk_ms, "pair_interactions": st.pair_interactions,
                  "nodes_visited": st.nodes_visited, "walk_fallbacks": st.walk_fallbacks, "walk_targets": st.walk_targets, "rounds": st.rounds}))
