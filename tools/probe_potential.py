"""scratch: one tree build + one walk launch (hbtu_tree_potential) on a single cuspy halo, for ncu"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hbtplus_b200 import capi, synth
from hbtplus_b200.unbind import UnbindContext

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4000000
p = capi.make_params(box_size=100.0, softening=4.8e-5, periodic=False)
e = capi.make_epoch(1.0)
snap = synth.make_snapshot([n], seed=3, wrap=False)
pm = snap.pos_mass
ctx = UnbindContext(p)
ctx.set_counting(len(sys.argv) > 2)
for _ in range(2):
    g = ctx.tree_potential(e, pm, pm, self_mass=pm[:, 3].copy())
    st = ctx.stats()
    print(f"n={n} walk_ms {st.walk_ms:.3f} inter {st.pair_interactions} visits {st.nodes_visited} fallbacks {st.walk_fallbacks}", flush=True)
