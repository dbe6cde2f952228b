"""Development probe (GPU box): walk/build throughput vs N for the stand-alone tree potential and one big unbind."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hbtplus_b200 import capi, synth
from hbtplus_b200.unbind import UnbindContext

sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [2_000_000, 16_000_000]
do_unbind = len(sys.argv) > 2 and sys.argv[2] == "unbind"
e = capi.make_epoch(1.0)
p = capi.make_params(box_size=62.5, softening=5e-3, periodic=False)
ctx = UnbindContext(p)
for n in sizes:
    snap = synth.make_snapshot([n], seed=n + 3, wrap=False)
    pm = snap.pos_mass
    sm = pm[:, 3].copy()
    ctx.set_counting(True)
    g = ctx.tree_potential(e, pm, pm, self_mass=sm)
    st = ctx.stats()
    inter, vis = st.pair_interactions, st.nodes_visited
    ctx.set_counting(False)
    for rep in range(2):
        g = ctx.tree_potential(e, pm, pm, self_mass=sm)
        st = ctx.stats()
        print(f"n={n} walk_ms {st.walk_ms:.2f} build_ms {st.build_ms:.2f} inter/target {inter/n:.0f} visits/warp {vis/(n/32):.0f} "
              f"=> {inter/st.walk_ms/1e9:.3f} T inter/s ({inter/st.walk_ms/1e9/4.65*100:.1f}% of 4.65e12)  visit-lane-steps/s {vis*32/st.walk_ms/1e9:.3f}T", flush=True)
    if do_unbind:
        t0 = time.time(); r = ctx.unbind_batch(e, snap, want_energy=False); dt = time.time() - t0
        st = ctx.stats()
        print(f"unbind n={n}: wall {dt:.3f}s rounds {st.rounds} walk {st.walk_ms:.1f} build {st.build_ms:.1f} other {st.other_ms:.1f} h2d {st.h2d_ms:.1f} d2h {st.d2h_ms:.1f} "
              f"nbound {r.io['nbound'][0]} iters {r.io['iterations'][0]} -> {n/dt/1e6:.2f} Mpart/s", flush=True)
ctx.close()
