#!/usr/bin/env python
"""A/B harness for the walk kernels (GPU box): the bench workload is generated ONCE, then every (library build, tuning) variant
runs the same staged batch: `python tools/ab_walk.py [--particles 1.8e8] [--steps 2] [--workload cfg2] spec [spec ...]`.
A spec is  lib=<path or 'default'>,key=value,...  with keys of hbtu_set_tuning (walk_masked_pairs, walk_masked_blocks, ...).
Prints one JSON line per variant (walk / build / other ms per step, fallbacks of a counted pass when --count is given).
Numbers are for ranking variants; bench.py is the measurement of record."""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from hbtplus_b200 import capi  # noqa: E402
from hbtplus_b200.unbind import UnbindContext  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--particles", type=float, default=1.8e8)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--count", action="store_true")
    ap.add_argument("--world", type=int, default=1, help="cfg4: rank 0's shard of a snapshot dealt to this many ranks")
    ap.add_argument("specs", nargs="+")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    wl = bench.WORKLOADS[args.workload]
    snap = wl.make(args.particles, dev, 0, args.world)
    torch.cuda.empty_cache()
    e = capi.make_epoch(1.0)
    for spec in args.specs:
        kv = dict(x.split("=", 1) for x in spec.split(","))
        lib = kv.pop("lib", "default")
        path = None if lib == "default" else (lib if os.path.isabs(lib) else os.path.join(ROOT, lib))
        try:
            ctx = UnbindContext(wl.params(0), lib_path=path)
            ctx._lib.hbtu_set_tuning.argtypes = [C.c_char_p, C.c_int64]
            for k, v in kv.items():
                rc = ctx._lib.hbtu_set_tuning(k.encode(), int(v))
                assert rc == 0, (k, v)
            ctx.stage(e, snap, capi.HBTU_FLAG_TRUNCATE_SOURCE)
            out = {"spec": spec}
            if args.count:
                ctx.set_counting(True)
                ctx.execute()
                st = ctx.stats()
                out.update(pair_interactions=int(st.pair_interactions), walk_fallbacks=int(st.walk_fallbacks))
                ctx.set_counting(False)
            ctx.execute()  # warm-up
            walk, build, other, tot = [], [], [], []
            for _ in range(args.steps):
                ctx.execute()
                st = ctx.stats()
                walk.append(st.walk_ms); build.append(st.build_ms); other.append(st.other_ms); tot.append(st.execute_ms)
            out.update(walk_ms=float(np.mean(walk)), build_ms=float(np.mean(build)), other_ms=float(np.mean(other)), step_ms=float(np.mean(tot)),
                       launches=int(st.kernel_launches), rounds=int(st.rounds))
            ctx.close()
        except Exception as ex:  # a variant that fails must not end the sweep
            out = {"spec": spec, "error": repr(ex)}
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
