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
