import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np
import cases
from hbtplus_b200 import capi, synth
from hbtplus_b200.unbind import UnbindContext
from oracle import pyoracle as po
lib = po.load_oracle()

def report(tag, snap, got, want):
    skip = cases.unbound_inputs(snap)
    nb_g, nb_w = got.io["nbound"], want.io["nbound"]
    bad = np.nonzero(nb_g != nb_w)[0]
    print(tag, "nsub", snap.nsub, "nbound differs at", len(bad))
    n = np.diff(snap.part_offset)
    for s in bad[:30]:
        print("  s", s, "n", n[s], "nb_g", nb_g[s], "nb_w", nb_w[s], "it", got.io["iterations"][s] if "iterations" in got.io.dtype.names else "", want.io["iterations"][s] if "iterations" in want.io.dtype.names else "",
              "mb", got.io["mbound"][s], want.io["mbound"][s], "J", cases.jaccard(got.bound(s), want.bound(s)), "death", got.io["snapshot_index_of_death"][s], want.io["snapshot_index_of_death"][s])
    mb_g, mb_w = got.io["mbound"], want.io["mbound"]
    rel = np.abs(mb_g - mb_w) / np.maximum(np.abs(mb_w), 1e-30)
    w = np.nonzero((rel > 1e-3) & ~skip)[0]
    print("  mbound gate failures:", w[:20], rel[w][:20])
    for f in ("avg_pos", "avg_vel", "mostbound_pos", "mostbound_vel"):
        a, b = got.io[f], want.io[f]
        d = ~np.isclose(a, b, rtol=2e-6, atol=1e-6).all(axis=1) & ~skip & (nb_g == nb_w)
        w = np.nonzero(d)[0]
        if len(w):
            print("  ", f, "differs at", w[:10])
            for s in w[:5]:
                print("     s", s, "n", n[s], "nb", nb_g[s], a[s], b[s])

which = sys.argv[1:] or ["5", "1"]
if "5" in which:
    rng = np.random.default_rng(20240005)
    tiny = rng.integers(20, 201, 3000)
    sizes = np.concatenate([[40000, 40000], tiny])
    parent = np.concatenate([[-1, -1], rng.integers(0, 2, len(tiny))])
    p = capi.make_params(box_size=250.0, softening=2.1e-3, periodic=False)
    e = capi.make_epoch(1.0, snapshot_index=30)
    snap = synth.make_snapshot(sizes, seed=20240005, box_size=250.0, particle_mass=0.02, parent=parent, wrap=False, f_contam=0.25)
    ctx = UnbindContext(p)
    got = ctx.unbind_batch(e, snap, flags=capi.HBTU_FLAG_TRUNCATE_SOURCE)
    want = po.run_batch(lib, "hbto", p, e, snap, flags=capi.HBTU_FLAG_TRUNCATE_SOURCE)
    report("cfg5", snap, got, want)
    ctx.close()
if "1" in which:
    from test_gpu_configs import next_snapshot
    rng = np.random.default_rng(20240001)
    subs = synth.subhalo_sizes(rng, 60, 20, 3000)
    sizes = np.concatenate([[70000], subs])
    parent = synth.nest_forest(rng, sizes, max_depth=2, p_nest=0.3, root=0)
    p = capi.make_params(box_size=62.5, softening=5e-3, periodic=True)
    snap = synth.make_snapshot(sizes, seed=20240001, parent=parent, wrap=True, centre=[0.2, 30.0, 62.3], f_contam=0.2)
    ctx = UnbindContext(p)
    snap_g = snap_o = snap
    for k, a in enumerate((0.8, 0.9, 1.0)):
        e = capi.make_epoch(a, snapshot_index=10 + k)
        got = ctx.unbind_batch(e, snap_g, flags=capi.HBTU_FLAG_TRUNCATE_SOURCE)
        want = po.run_batch(lib, "hbto", p, e, snap_o, flags=capi.HBTU_FLAG_TRUNCATE_SOURCE)
        print("same inputs:", np.array_equal(snap_g.part_offset, snap_o.part_offset) and np.array_equal(snap_g.pos_mass, snap_o.pos_mass))
        report(f"cfg1 k={k}", snap_o, got, want)
        snap_g = next_snapshot(snap_g, got, 62.5)
        snap_o = next_snapshot(snap_o, want, 62.5)
        if not np.array_equal(snap_g.part_offset, snap_o.part_offset):
            snap_g = snap_o
