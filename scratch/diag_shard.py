import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np
from hbtplus_b200 import capi
from oracle import pyoracle as po
from test_gpu_dropin import snapshot_with_hosts
drop = po.load_dropin("v32")
p = capi.make_params(box_size=62.5, softening=5e-3, periodic=True)
e = capi.make_epoch(0.9, snapshot_index=15)
snap, host, n_old, nhalos, mb = snapshot_with_hosts(21, True)
os.environ.pop("HBT_UNBIND_DEVICES", None)
one = po.refine_particles(drop, p, e, snap, host, n_old, nhalos, mb)
os.environ["HBT_UNBIND_DEVICES"] = "0,0,0"
three = po.refine_particles(drop, p, e, snap, host, n_old, nhalos, mb)
for s in range(snap.nsub):
    b = one.order_offset[s]; n = one.io["nsource"][s]; nb = one.io["nbound"][s]
    d = np.nonzero(one.energy[b:b+n] != three.energy[b:b+n])[0]
    if len(d): print("sub", s, "n", n, "nbound", nb, "ndiff", len(d), "first", d[:5], one.energy[b+d[:3]], three.energy[b+d[:3]], "iters", one.io["iterations"][s])
