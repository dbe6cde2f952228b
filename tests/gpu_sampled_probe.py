"""Development probe (GPU box): sampled-mode parity against the oracle's shuffle mode 1."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hbtplus_b200 import capi, synth
from hbtplus_b200.unbind import UnbindContext
from oracle import pyoracle as po
orc = po.load_oracle(); orc.hbto_set_shuffle_mode(1)
e = capi.make_epoch(1.0)
for M, refine in ((1000, True), (64, True), (200, False)):
    for periodic in (False, True):
        p = capi.make_params(box_size=62.5, softening=5e-3, periodic=periodic, max_sample_size=M, refine_mostbound=refine, shuffle_seed=99)
        sizes = [5000, 1200, 800, 20000, 300, 50, 2500, 70, 30]
        parent = [-1, 0, 0, -1, 3, 4, -1, 6, 6]
        snap = synth.make_snapshot(sizes, seed=3 + M, parent=parent, wrap=periodic, f_contam=0.35)
        ctx = UnbindContext(p)
        g = ctx.unbind_batch(e, snap)
        w = po.run_batch(orc, "hbto", p, e, snap)
        print(f"M={M} refine={refine} periodic={periodic}")
        print("  nbound gpu", g.io["nbound"].tolist()); print("  nbound cpu", w.io["nbound"].tolist())
        print("  iters gpu", g.io["iterations"].tolist(), "cpu", w.io["iterations"].tolist())
        for s in range(snap.nsub):
            a, b = g.particles(s), w.particles(s)
            nb = int(w.io["nbound"][s])
            same = np.array_equal(a, b)
            nd = int((a != b).sum()) if len(a) == len(b) else -1
            first = int(np.nonzero(a != b)[0][0]) if nd > 0 else -1
            print(f"   sub {s} n={len(b)} nb={nb} order_equal={same} ndiff={nd} first={first} mostbound_equal={np.array_equal(g.io['mostbound_pos'][s], w.io['mostbound_pos'][s])} "
                  f"dpos={np.abs(g.io['avg_pos'][s]-w.io['avg_pos'][s]).max():.1e} dm={abs(g.io['mbound'][s]-w.io['mbound'][s]):.1e}")
        ctx.close()
