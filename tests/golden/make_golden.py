"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libhbtref_v32.so).

Run in the build container (where /root/reference exists and `make -C oracle ref` has been run):
    python tests/golden/make_golden.py
Each fixture stores the exact inputs (so that it does not depend on the generator's RNG stream) and the
reference's outputs for flags=0 and flags=HBTU_FLAG_TRUNCATE_SOURCE, plus a tree-potential vector.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from hbtplus_b200 import capi  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
import cases  # noqa: E402


def main():
    ref = po.load_ref()
    ref.hbtref_set_num_threads(1)
    for name, fn in cases.CASES.items():
        p, e, snap = fn()
        out = {
            "pos_mass": snap.pos_mass, "vel": snap.vel, "part_offset": snap.part_offset, "io_in": snap.io,
            "nest_offset": snap.nest_offset if snap.nest_offset is not None else np.zeros(0, np.int64),
            "nest_list": snap.nest_list if snap.nest_list is not None else np.zeros(0, np.int32),
            "has_nest": np.array(snap.nest_offset is not None),
        }
        for tag, flags in (("full", 0), ("trunc", capi.HBTU_FLAG_TRUNCATE_SOURCE)):
            ref.hbtref_seed(cases.SAMPLED_SRAND)  # only the sampled case draws from rand()
            r = po.run_batch(ref, "hbtref", p, e, snap, flags=flags)
            ntot = int(r.order_offset[-1])
            out[f"{tag}_io"] = r.io
            out[f"{tag}_order_offset"] = r.order_offset
            out[f"{tag}_order"] = r.order[:ntot]
            out[f"{tag}_energy"] = r.energy[:ntot]
        # GravityTree_t::EvaluatePotential / BindingEnergy of the largest subhalo's particles on 256 targets
        s = int(np.argmax(np.diff(snap.part_offset)))
        b, en = snap.part_offset[s], snap.part_offset[s + 1]
        src = snap.pos_mass[b:en]
        tgt = src[:: max(1, len(src) // 256)][:256]
        out["pot_src_range"] = np.array([b, en])
        out["pot_tgt"] = tgt
        out["pot_self"] = po.tree_potential(ref, "hbtref", p, e, src, tgt, self_mass=tgt[:, 3].copy())
        out["pot_foreign"] = po.tree_potential(ref, "hbtref", p, e, src, tgt + np.float32([0.01, 0.0, -0.02, 0.0]))
        tv = snap.vel[b:en][:: max(1, len(src) // 256)][:256]
        out["be"] = po.tree_potential(ref, "hbtref", p, e, src, tgt, self_mass=tgt[:, 3].copy(), tgt_vel=tv,
                                      ref_pos=snap.io["avg_pos"][s], ref_vel=snap.io["avg_vel"][s])
        out["be_vel"] = tv
        # Subhalo_t::CalculateProfileProperties + CalculateShape on the truncated result (src/subhalo.cpp:242-398)
        r = po.Result(out["trunc_io"], out["trunc_order_offset"], out["trunc_order"], out["trunc_energy"])
        ppo, ppm, pio = cases.profile_inputs(snap, r, seed=len(name))
        out["prof_part_offset"], out["prof_pos_mass"], out["prof_io_in"] = ppo, ppm, pio
        out["prof_io"] = po.profile_batch(ref, "hbtref", p, e, ppo, ppm, pio)
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **out)
        print(name, "->", path, os.path.getsize(path) // 1024, "KiB", "nbound", out["full_io"]["nbound"])


def variants():
    """The reference's compile-time physics variants (oracle/_ref/libhbtref_v32ns.so, libhbtref_v32th.so)."""
    for name, (fn, variant, vflag) in cases.VARIANT_CASES.items():
        ref = po.load_ref_variant(variant)
        ref.hbtref_set_num_threads(1)
        p, e, snap = fn()
        out = {"pos_mass": snap.pos_mass, "vel": snap.vel, "part_offset": snap.part_offset, "io_in": snap.io,
               "nest_offset": snap.nest_offset if snap.nest_offset is not None else np.zeros(0, np.int64),
               "nest_list": snap.nest_list if snap.nest_list is not None else np.zeros(0, np.int32), "has_nest": np.array(snap.nest_offset is not None)}
        for tag, flags in (("full", vflag), ("trunc", vflag | capi.HBTU_FLAG_TRUNCATE_SOURCE)):
            r = po.run_batch(ref, "hbtref", p, e, snap, flags=flags)
            ntot = int(r.order_offset[-1])
            out[f"{tag}_io"], out[f"{tag}_order_offset"], out[f"{tag}_order"], out[f"{tag}_energy"] = r.io, r.order_offset, r.order[:ntot], r.energy[:ntot]
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **out)
        print(name, "->", path, os.path.getsize(path) // 1024, "KiB", "nbound", out["full_io"]["nbound"])


def mask():
    """SubhaloSnapshot_t::MaskSubhalos of the unmodified reference on cases.case_mask()."""
    ref = po.load_ref()
    p = capi.make_params(box_size=62.5, softening=5e-3)
    part_offset, ids, nest_offset, nest_list, nbound = cases.case_mask()
    new_count, keep = po.mask_batch(ref, "hbtref", p, part_offset, ids, nest_offset, nest_list, nbound)
    path = os.path.join(HERE, "mask.npz")
    np.savez_compressed(path, part_offset=part_offset, ids=ids, nest_offset=nest_offset, nest_list=nest_list, nbound=nbound,
                        new_count=new_count, keep=keep)
    print("mask ->", path, os.path.getsize(path) // 1024, "KiB", "kept", int(new_count.sum()), "of", int(part_offset[-1]))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "mask":
        mask()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "variants":
        variants()
        sys.exit(0)
    main()
