"""BASELINE.json's other configs as parity cases (reduced so that the CPU oracle finishes in seconds):
cfg1 Example.conf-style snapshot series, cfg3 MilliMill-style box of FoF groups, cfg5 DynamicMerger-style swarm of
tiny subhaloes.  (cfg2 is the bench workload; cfg4 is cfg3's shape with the V64 ABI, covered by test_gpu_dropin[v64].)"""
import numpy as np
import pytest

import cases
from hbtplus_b200 import capi, synth
from oracle import pyoracle as po
from test_gpu_parity import check_batch, make_ctx  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu


def test_cfg5_dynamic_merger_swarm(make_ctx, oracle_lib):
    """configs/DynamicMerger.conf shape: two haloes + a swarm of tiny subhaloes (n in [20,200]), periodic off."""
    rng = np.random.default_rng(20240005)
    tiny = rng.integers(20, 201, 3000)
    sizes = np.concatenate([[40000, 40000], tiny])
    parent = np.concatenate([[-1, -1], rng.integers(0, 2, len(tiny))])
    p = capi.make_params(box_size=250.0, softening=2.1e-3, periodic=False)
    e = capi.make_epoch(1.0, snapshot_index=30)
    snap = synth.make_snapshot(sizes, seed=20240005, box_size=250.0, particle_mass=0.02, parent=parent, wrap=False, f_contam=0.25)
    ctx = make_ctx(p)
    got = ctx.unbind_batch(e, snap, flags=capi.HBTU_FLAG_TRUNCATE_SOURCE)
    want = po.run_batch(oracle_lib, "hbto", p, e, snap, flags=capi.HBTU_FLAG_TRUNCATE_SOURCE)
    check_batch(snap, got, want)
    st = ctx.stats()
    assert st.rounds < 40  # thousands of subhaloes share each round's launches
    assert (want.io["nbound"][2:] > 1).sum() > 1000 and (want.io["nbound"][2:] == 1).sum() > 10


def box_of_groups(rng, ngroups):
    """Hosts with 0-5 satellites; a satellite hangs off the host or off an EARLIER, larger satellite (depth <= 3)."""
    sizes, parent, depth = [], [], []
    for g in range(ngroups):
        n_host = int(synth.subhalo_sizes(rng, 1, 200, 30000)[0])
        c = len(sizes)
        sizes.append(n_host)
        parent.append(-1)
        depth.append(0)
        nsat = int(rng.integers(0, 6))
        sats = np.sort(synth.subhalo_sizes(rng, nsat, 20, max(21, n_host // 4)))[::-1] if nsat else []
        first = len(sizes)
        for n in sats:
            me = len(sizes)
            par = c
            if me > first and rng.random() < 0.4:
                cand = int(rng.integers(first, me))  # strictly earlier index: no cycles
                if sizes[cand] > n and depth[cand] < 2:
                    par = cand
            sizes.append(int(n))
            parent.append(par)
            depth.append(depth[par] + 1)
    return np.array(sizes), np.array(parent)


def test_cfg3_box_of_groups(make_ctx, oracle_lib):
    """configs/MilliMill.conf shape: many FoF groups (central + satellites, depth <= 3), periodic box."""
    rng = np.random.default_rng(20240003)
    sizes, parent = box_of_groups(rng, 150)
    p = capi.make_params(box_size=62.5, softening=5e-3, periodic=True)
    e = capi.make_epoch(0.7, snapshot_index=40)
    snap = synth.make_snapshot(sizes, seed=20240003, parent=parent, wrap=True, f_contam=0.2)
    ctx = make_ctx(p)
    got = ctx.unbind_batch(e, snap, flags=capi.HBTU_FLAG_TRUNCATE_SOURCE)
    want = po.run_batch(oracle_lib, "hbto", p, e, snap, flags=capi.HBTU_FLAG_TRUNCATE_SOURCE)
    check_batch(snap, got, want)


def test_cfg1_snapshot_series(make_ctx, oracle_lib):
    """configs/Example.conf shape: one host across the periodic face + nested subhaloes, THREE snapshots in a row: the
    truncated source list of snapshot k (in its new order) is the input of snapshot k+1, so a wrong order, a wrong
    most-bound particle or a wrong truncation would propagate (src/subhalo_unbind.cpp:409-418,449-458)."""
    rng = np.random.default_rng(20240001)
    subs = synth.subhalo_sizes(rng, 60, 20, 3000)
    sizes = np.concatenate([[70000], subs])
    parent = synth.nest_forest(rng, sizes, max_depth=2, p_nest=0.3, root=0)
    p = capi.make_params(box_size=62.5, softening=5e-3, periodic=True)
    snap = synth.make_snapshot(sizes, seed=20240001, parent=parent, wrap=True, centre=[0.2, 30.0, 62.3], f_contam=0.2)
    ctx = make_ctx(p)
    snap_g, snap_o = snap, snap
    for k, a in enumerate((0.8, 0.9, 1.0)):
        e = capi.make_epoch(a, snapshot_index=10 + k)
        got = ctx.unbind_batch(e, snap_g, flags=capi.HBTU_FLAG_TRUNCATE_SOURCE)
        want = po.run_batch(oracle_lib, "hbto", p, e, snap_o, flags=capi.HBTU_FLAG_TRUNCATE_SOURCE)
        check_batch(snap_o, got, want, exact_frames=(k == 0))
        snap_g = next_snapshot(snap_g, got, 62.5)
        snap_o = next_snapshot(snap_o, want, 62.5)
        if not np.array_equal(snap_g.part_offset, snap_o.part_offset):
            snap_g = snap_o  # a round-off flip changed one list length: continue both sides from the oracle's lists


def next_snapshot(snap, res, box, dt=2e-5):
    """Sources of the next snapshot = the truncated particle lists in their new order, drifted by v*dt."""
    sizes = res.io["nsource"].astype(np.int64)
    part_offset = np.zeros(snap.nsub + 1, np.int64)
    np.cumsum(sizes, out=part_offset[1:])
    idx = np.concatenate([res.particles(s) for s in range(snap.nsub)]) if part_offset[-1] else np.zeros(0, np.int64)
    pm = snap.pos_mass[idx].copy()
    vel = snap.vel[idx].copy()
    pm[:, :3] = np.mod(pm[:, :3] + vel[:, :3] * np.float32(dt), np.float32(box))
    io = res.io.copy()
    return synth.Snapshot(part_offset, pm, vel, snap.nest_offset, snap.nest_list, io)
