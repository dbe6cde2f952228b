"""Seeded parity cases shared by the CPU tests, the GPU tests and the golden-fixture generator."""
from __future__ import annotations

import numpy as np

from hbtplus_b200 import capi, synth


def case_flat():
    p = capi.make_params(box_size=62.5, softening=5e-3, periodic=False)
    e = capi.make_epoch(1.0)
    snap = synth.make_snapshot([50, 200, 1000, 3000, 10, 1, 0, 25, 2, 19, 20, 21, 300, 64], seed=11, wrap=False)
    return p, e, snap


def case_periodic_straddle():
    """Subhaloes sitting across the x=0 / z=L faces: exercises the un-wrapped bbox + NEAREST paths."""
    p = capi.make_params(box_size=62.5, softening=5e-3, periodic=True)
    e = capi.make_epoch(0.8, snapshot_index=7)
    snap = synth.make_snapshot([400, 1500, 90, 33], seed=12, wrap=True, centre=[0.05, 31.0, 62.45])
    return p, e, snap


def case_nested():
    """Depth-3 nest with an orphan in the middle (src/subhalo_unbind.cpp:432-447)."""
    p = capi.make_params(box_size=62.5, softening=5e-3, periodic=True)
    e = capi.make_epoch(1.0, snapshot_index=12)
    sizes = [4000, 600, 300, 80, 150, 40, 1, 60, 500, 25]
    parent = [-1, 0, 0, 1, 1, 2, 0, 6, -1, 8]
    snap = synth.make_snapshot(sizes, seed=13, wrap=True, parent=parent, f_contam=0.3)
    snap.io["nbound"][6] = 1  # orphan: one tracer particle, still feeds its children's tails upwards
    snap.io["snapshot_index_of_death"][6] = 3
    snap.io["sink_track_id"][1] = 5
    snap.io["snapshot_index_of_sink"][1] = 9
    return p, e, snap


def case_massvar():
    """Unequal particle masses, a=0.5 (Hubble-flow term matters), hotter contaminants."""
    p = capi.make_params(box_size=100.0, softening=2e-3, periodic=False, min_num_part_of_sub=10)
    e = capi.make_epoch(0.5, snapshot_index=3)
    snap = synth.make_snapshot([2500, 700, 120, 15, 9], seed=14, wrap=False, mass_scatter=0.5, f_contam=0.35, contam_hot=5.0, particle_mass=0.01)
    return p, e, snap


def case_sampled():
    """Reference default MaxSampleSizeOfPotentialEstimate=1000 (+RefineMostboundParticle): sources above and below the
    sample size, nested.  CPU fixtures use srand(7) + one thread (the reference's shuffle draws from libc rand())."""
    p = capi.make_params(box_size=62.5, softening=5e-3, periodic=True, max_sample_size=1000, refine_mostbound=True, shuffle_seed=99)
    e = capi.make_epoch(1.0, snapshot_index=21)
    snap = synth.make_snapshot([5000, 1200, 800, 20000, 300, 50, 2500, 70, 30], seed=1003, parent=[-1, 0, 0, -1, 3, 4, -1, 6, 6], wrap=True, f_contam=0.35)
    return p, e, snap


def case_nostrip():
    """-DNO_STRIPPING build of the reference (batch flag HBTU_FLAG_NO_STRIPPING): Nbound = Nlast, one evaluation, everything
    kept and E-sorted; periodic and across the box face so that the Elist[0] origin matters; one all-unbound source."""
    p = capi.make_params(box_size=62.5, softening=5e-3, periodic=True)
    e = capi.make_epoch(0.9, snapshot_index=8)
    snap = synth.make_snapshot([900, 300, 60, 25, 15, 40], seed=15, wrap=True, parent=[-1, 0, 1, -1, -1, -1], centre=[62.45, 0.03, 31.0], f_contam=0.3)
    b, en = snap.part_offset[5], snap.part_offset[6]
    snap.vel[b:en, :3] *= 40.0  # source 5: nothing is bound
    return p, e, snap


def case_thermal():
    """-DUNBIND_WITH_THERMAL_ENERGY build (batch flag HBTU_FLAG_THERMAL_ENERGY): vel[:,3] carries Particle_t::InternalEnergy,
    here a fraction of each subhalo's velocity dispersion squared with a hot tail."""
    p = capi.make_params(box_size=62.5, softening=5e-3, periodic=False)
    e = capi.make_epoch(1.0, snapshot_index=9)
    snap = synth.make_snapshot([3000, 800, 200, 50, 22], seed=16, wrap=False, parent=[-1, 0, 0, 1, -1], f_contam=0.25)
    rng = np.random.default_rng(16)
    for s in range(snap.nsub):
        b, en = snap.part_offset[s], snap.part_offset[s + 1]
        v = snap.vel[b:en, :3]
        sig2 = float(((v - v.mean(0)) ** 2).sum(1).mean())
        snap.vel[b:en, 3] = (0.05 * sig2 * rng.exponential(1.0, en - b)).astype(np.float32)
    return p, e, snap


SAMPLED_SRAND = 7
VARIANT_CASES = {"nostrip": (case_nostrip, "v32ns", capi.HBTU_FLAG_NO_STRIPPING), "thermal": (case_thermal, "v32th", capi.HBTU_FLAG_THERMAL_ENERGY)}

CASES = {"flat": case_flat, "periodic_straddle": case_periodic_straddle, "nested": case_nested, "massvar": case_massvar, "sampled": case_sampled}

IO_EXACT = ["nbound", "snapshot_index_of_death", "snapshot_index_of_sink", "sink_track_id", "nsource", "nsource_full"]
IO_FLOAT = ["mbound", "avg_pos", "avg_vel", "mostbound_pos", "mostbound_vel", "specific_self_potential_energy",
            "specific_self_kinetic_energy", "specific_angular_momentum"]


def unbound_inputs(snap):
    """Subhaloes whose Unbind leaves the kinetic fields untouched (n<=1 or orphan): those outputs are not compared."""
    n = np.diff(snap.part_offset)
    return (n <= 1) | (snap.io["nbound"] <= 1)


def jaccard(a, b) -> float:
    a, b = set(np.asarray(a).tolist()), set(np.asarray(b).tolist())
    return len(a & b) / len(a | b) if (a or b) else 1.0


def profile_inputs(snap, res, last_vmax=None, seed=0):
    """Inputs of CalculateProfileProperties / CalculateShape from an unbinding result: every subhalo's particle list in
    its new order (bound first), centre = most-bound position, Nbound, Mbound; a previous Vmax record and a previous
    overdensity size are invented so that the [io] semantics are exercised."""
    n = res.io["nsource"].astype(np.int64)
    part_offset = np.zeros(snap.nsub + 1, np.int64)
    np.cumsum(n, out=part_offset[1:])
    idx = np.concatenate([res.particles(s) for s in range(snap.nsub)]) if part_offset[-1] else np.zeros(0, np.int64)
    pm = np.ascontiguousarray(snap.pos_mass[idx])
    io = np.zeros(snap.nsub, capi.PROFILEIO_DTYPE)
    io["mostbound_pos"] = res.io["mostbound_pos"]
    io["nbound"] = res.io["nbound"]
    io["mbound"] = res.io["mbound"]
    rng = np.random.default_rng(seed)
    io["last_max_vmax_physical"] = rng.choice([0.0, 1e9], snap.nsub) if last_vmax is None else last_vmax
    io["snapshot_index_of_last_max_vmax"] = np.where(io["last_max_vmax_physical"] > 0, 3, -1)
    io["bound_r200crit_comoving"] = 7.0
    io["bound_m200crit"] = 11.0
    return part_offset, pm, io


PROFILE_FIELDS = ["rmax_comoving", "vmax_physical", "last_max_vmax_physical", "snapshot_index_of_last_max_vmax", "r2sigma_comoving",
                  "rhalf_comoving", "bound_r200crit_comoving", "bound_m200crit", "inertial_tensor", "inertial_tensor_weighted"]


def case_mask(seed=41, nroots=6, scale=1.0):
    """Particle-Id lists with the overlaps MaskSubhalos exists for: every hierarchy draws its Ids from a shared pool, a
    child shares part of its parent's particles, siblings overlap, some Ids repeat inside one list, some subhaloes are
    orphans (Nbound <= 1) or empty, and different hierarchies reuse the same Ids (they must not exclude each other)."""
    rng = np.random.default_rng(seed)
    sizes, parent = [], []
    for r in range(nroots):
        root = len(sizes)
        sizes.append(int(scale * rng.integers(200, 3000)))
        parent.append(-1)
        for _ in range(int(rng.integers(0, 6))):
            par = int(rng.integers(root, len(sizes)))
            sizes.append(int(scale * rng.integers(0, 400)))
            parent.append(par)
    sizes[-1] = 0
    nsub = len(sizes)
    part_offset = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    ids = np.zeros(part_offset[-1], np.int64)
    pool = (rng.integers(1, 2**62, 5000, dtype=np.int64))  # 64-bit Ids, shared by ALL hierarchies
    for s in range(nsub):
        n = sizes[s]
        own = pool[rng.integers(0, len(pool), n)]
        if parent[s] >= 0 and n:
            pb, pe = part_offset[parent[s]], part_offset[parent[s] + 1]
            if pe > pb:
                share = rng.random(n) < 0.4
                own[share] = ids[pb:pe][rng.integers(0, pe - pb, int(share.sum()))]
        ids[part_offset[s]:part_offset[s + 1]] = own
    nbound = np.array([max(2, n // 2) for n in sizes], np.int64)
    nbound[rng.random(nsub) < 0.15] = 1  # orphans
    nbound[np.array(sizes) == 0] = 0
    children = [[] for _ in range(nsub)]
    for s, p_ in enumerate(parent):
        if p_ >= 0:
            children[p_].append(s)
    nest_offset = np.concatenate([[0], np.cumsum([len(c) for c in children])]).astype(np.int64)
    nest_list = np.array([c for cs in children for c in cs], np.int32)
    return part_offset, ids, nest_offset, nest_list, nbound


def case_traps(seed=51, periodic=True):
    """Nested subhaloes some of which sit inside their host's core in phase space (-> trapped), some do not, orphans as
    satellites and as hosts (skipped as hosts), an already trapped one, depth-3 chains where the first real host is two
    levels up, and a host straddling the box face."""
    rng = np.random.default_rng(seed)
    sizes = [5000, 800, 300, 60, 45, 1, 200, 0, 120, 3000, 500, 90, 30]
    parent = [-1, 0, 0, 1, 3, 1, 5, 0, 2, -1, 9, 10, 11]
    snap = synth.make_snapshot(sizes, seed=seed, wrap=periodic, parent=parent, centre=[0.03, 31.0, 62.47] if periodic else None, f_contam=0.0)
    nsub = snap.nsub
    io = np.zeros(nsub, capi.TRAPIO_DTYPE)
    io["sink_track_id"] = -1
    io["snapshot_index_of_sink"] = -1
    nbound = np.array(sizes, np.int64)
    nbound[5] = 1
    io["nbound"] = nbound
    # cores: mean of the first 20 particles; put half of the satellites right on their host's core, the others far in velocity
    for s in range(nsub):
        b, e = snap.part_offset[s], snap.part_offset[s + 1]
        if e > b:
            io["mostbound_pos"][s] = snap.pos_mass[b, :3]
            io["mostbound_vel"][s] = snap.vel[b, :3]
    for s in range(nsub):
        p_ = parent[s]
        if p_ < 0 or sizes[s] == 0:
            continue
        host = p_
        while host >= 0 and nbound[host] <= 1:
            host = parent[host]
        hb = snap.part_offset[host]
        hp = snap.pos_mass[hb:hb + 20, :3].astype(np.float64)
        hv = snap.vel[hb:hb + 20, :3].astype(np.float64)
        d0 = hp - hp[0]
        if periodic:
            d0 -= 62.5 * np.round(d0 / 62.5)
        cpos, cvel = hp[0] + d0.mean(0), hv.mean(0)
        sr, sv = np.sqrt(d0.var(0).sum()), np.sqrt(hv.var(0).sum())
        if s % 2 == 0:   # inside: 0.4 sigma in position, 0.5 sigma in velocity -> delta ~ 0.9
            io["mostbound_pos"][s] = np.mod(cpos + 0.4 * sr * np.array([1.0, 0, 0]), 62.5) if periodic else cpos + 0.4 * sr * np.array([1.0, 0, 0])
            io["mostbound_vel"][s] = cvel + 0.5 * sv * np.array([0, 1.0, 0])
        else:            # outside: 3 sigma in velocity
            io["mostbound_vel"][s] = cvel + 3.0 * sv * np.array([0, 0, 1.0])
    io["sink_track_id"][8] = 0  # already trapped: left alone
    io["snapshot_index_of_sink"][8] = 4
    children = [[] for _ in range(nsub)]
    for s, p_ in enumerate(parent):
        if p_ >= 0:
            children[p_].append(s)
    nest_offset = np.concatenate([[0], np.cumsum([len(c) for c in children])]).astype(np.int64)
    nest_list = np.array([c for cs in children for c in cs], np.int32)
    return snap, nest_offset, nest_list, io


def case_merge(seed=61, periodic=True, nhosts=12):
    """`nhosts` host haloes, each a central with satellites (and satellites of satellites); about half of the satellites sit on
    their host's 20-particle core in phase space and get trapped, so MergeSubhalos (MergeTrappedSubhalos on) merges them into
    their sinks and re-unbinds those hosts (src/subhalo_merge.cpp:201-214).  Returns a synth.Snapshot whose io carries the
    trap inputs (mostbound position / velocity, Nbound, sink ids)."""
    rng = np.random.default_rng(seed)
    sizes, parent = [], []
    for h in range(nhosts):
        c = len(sizes)
        sizes.append(int(rng.integers(1500, 5000)))
        parent.append(-1)
        for k in range(int(rng.integers(2, 5))):
            s = len(sizes)
            sizes.append(int(rng.integers(60, 600)))
            parent.append(c)
            if rng.random() < 0.5:
                sizes.append(int(rng.integers(25, 50)))
                parent.append(s)
    snap = synth.make_snapshot(sizes, seed=seed, wrap=periodic, parent=parent, f_contam=0.1)
    box = 62.5
    nsub = snap.nsub
    io = snap.io
    for s in range(nsub):
        p_ = parent[s]
        if p_ < 0:
            continue
        hb = snap.part_offset[p_]
        hp = snap.pos_mass[hb:hb + 20, :3].astype(np.float64)
        hv = snap.vel[hb:hb + 20, :3].astype(np.float64)
        d0 = hp - hp[0]
        if periodic:
            d0 -= box * np.round(d0 / box)
        cpos, cvel = hp[0] + d0.mean(0), hv.mean(0)
        sr, sv = np.sqrt(d0.var(0).sum()), np.sqrt(hv.var(0).sum())
        if rng.random() < 0.5:  # inside the host's core: delta ~ 0.9 < 2
            x = cpos + 0.4 * sr * np.array([1.0, 0, 0])
            io["mostbound_pos"][s] = np.mod(x, box) if periodic else x
            io["mostbound_vel"][s] = cvel + 0.5 * sv * np.array([0, 1.0, 0])
        else:
            io["mostbound_vel"][s] = cvel + 3.0 * sv * np.array([0, 0, 1.0])
    return snap


def report(name, **stats):
    """Observed parity statistics of a GPU test, appended to gpurun_out/parity_stats.jsonl (travels back from the GPU box;
    summarised under profiles/) and printed (pytest -s)."""
    import json
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    line = json.dumps({"test": name, **stats})
    print("PARITY", line)
    try:
        os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
        with open(os.path.join(root, "gpurun_out", "parity_stats.jsonl"), "a") as f:
            f.write(line + "\n")
    except OSError:
        pass
