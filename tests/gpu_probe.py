"""Development probe run on the GPU box (not a pytest file): quick parity + timing dump."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hbtplus_b200 import capi, synth
from hbtplus_b200.unbind import UnbindContext
from oracle import pyoracle as po

orc = po.load_oracle()
e = capi.make_epoch(1.0)
for periodic in (False, True):
    p = capi.make_params(box_size=62.5, softening=5e-3, periodic=periodic)
    ctx = UnbindContext(p)
    ctx.set_counting(True)
    for n in (1, 2, 5, 100, 1000, 20000, 200000):
        snap = synth.make_snapshot([n], seed=n + 3, wrap=periodic, centre=[0.1, 31, 62.4] if periodic else None)
        pm = snap.pos_mass
        sm = pm[:, 3].copy()
        t0 = time.time(); g = ctx.tree_potential(e, pm, pm, self_mass=sm); tg = time.time() - t0
        st = ctx.stats()
        t0 = time.time(); o = po.tree_potential(orc, "hbto", p, e, pm, pm, self_mass=sm); to = time.time() - t0
        rel = np.abs(g - o) / np.maximum(np.abs(o), 1e-30)
        print(f"pot periodic={periodic} n={n} relerr max {rel.max():.2e} mean {rel.mean():.2e} gpu {tg:.3f}s cpu {to:.3f}s walk_ms {st.walk_ms:.3f} build_ms {st.build_ms:.3f} "
              f"inter {st.pair_interactions} (oracle {orc.hbto_last_interactions()}) visits {st.nodes_visited}", flush=True)
    sizes = [50, 200, 1000, 5000, 20000, 10, 1, 0, 25, 2, 19, 20, 21, 300, 64]
    snap = synth.make_snapshot(sizes, seed=1, wrap=periodic)
    t0 = time.time(); g = ctx.unbind_batch(e, snap); tg = time.time() - t0
    st = ctx.stats()
    t0 = time.time(); o = po.run_batch(orc, "hbto", p, e, snap); to = time.time() - t0
    print(f"unbind periodic={periodic} gpu {tg:.3f}s cpu {to:.3f}s rounds {st.rounds} launches {st.kernel_launches} walk_ms {st.walk_ms:.2f} build_ms {st.build_ms:.2f} other_ms {st.other_ms:.2f}")
    print(" nbound gpu", g.io["nbound"]); print(" nbound cpu", o.io["nbound"])
    print(" iters gpu", g.io["iterations"]); print(" iters cpu", o.io["iterations"])
    print(" mbound rel", np.abs(g.io["mbound"] - o.io["mbound"]) / np.maximum(o.io["mbound"], 1e-30))
    for s in range(snap.nsub):
        a, b = set(g.bound(s).tolist()), set(o.bound(s).tolist())
        jac = len(a & b) / max(len(a | b), 1)
        same_order = np.array_equal(g.particles(s), o.particles(s))
        print(f"  sub {s} n={sizes[s]} jaccard {jac:.5f} order_equal {same_order} avgpos d {np.abs(g.io['avg_pos'][s]-o.io['avg_pos'][s]).max():.2e} "
              f"avgvel d {np.abs(g.io['avg_vel'][s]-o.io['avg_vel'][s]).max():.2e} pot {g.io['specific_self_potential_energy'][s]:.6g}/{o.io['specific_self_potential_energy'][s]:.6g} "
              f"death {g.io['snapshot_index_of_death'][s]}/{o.io['snapshot_index_of_death'][s]}")
    ctx.close()
print("PROBE DONE")
