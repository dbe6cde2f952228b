"""Input builders shared by bench.py's secondary rows (kept out of tests/ so that bench.py does not import test code)."""
import numpy as np

from hbtplus_b200 import capi


def profile_inputs(snap, res):
    """Particle lists in their new order (bound first) + the [in] fields of hbtu_profile_io from an unbinding result."""
    n = res.io["nsource"].astype(np.int64)
    part_offset = np.zeros(snap.nsub + 1, np.int64)
    np.cumsum(n, out=part_offset[1:])
    ntot = int(part_offset[-1])
    order = np.asarray(res.order[:ntot]) if int(res.order_offset[-1]) == ntot else np.concatenate([res.particles(s) for s in range(snap.nsub)])
    pm = np.ascontiguousarray(snap.pos_mass[order])
    io = np.zeros(snap.nsub, capi.PROFILEIO_DTYPE)
    io["mostbound_pos"] = res.io["mostbound_pos"]
    io["nbound"] = res.io["nbound"]
    io["mbound"] = res.io["mbound"]
    io["snapshot_index_of_last_max_vmax"] = -1
    return part_offset, pm, io


def mask_inputs(snap, seed=3, share=0.3):
    """Particle-Id lists for MaskSubhalos from a synthetic snapshot: Id = global particle index, then every nested subhalo
    takes `share` of its entries from its parent's list (the overlap exclusive ownership removes)."""
    rng = np.random.default_rng(seed)
    po = np.asarray(snap.part_offset, np.int64)
    ids = np.arange(po[-1], dtype=np.int64)
    if snap.nest_offset is not None:
        for s in range(snap.nsub):
            pb, pe = po[s], po[s + 1]
            if pe == pb:
                continue
            for k in range(snap.nest_offset[s], snap.nest_offset[s + 1]):
                c = int(snap.nest_list[k])
                cb, ce = po[c], po[c + 1]
                m = int(share * (ce - cb))
                if m:
                    ids[cb + rng.choice(ce - cb, m, replace=False)] = ids[pb + rng.integers(0, pe - pb, m)]
    nbound = np.maximum(np.diff(po), 2).astype(np.int64)
    return po, ids, snap.nest_offset, snap.nest_list, nbound
