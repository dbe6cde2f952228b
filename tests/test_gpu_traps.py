"""GPU parity of the merger trap detection through the C-ABI ``hbtu_detect_traps``:
SubHelper_t::BuildPosition/BuildVelocity, SinkDistance, DetectTraps (src/subhalo_merge.cpp:29-172).  Sink ids are index
work: bit-exact (the 20-particle moments are summed serially in the reference's own order)."""
import numpy as np
import pytest

import cases
from hbtplus_b200 import capi
from oracle import pyoracle as po
from test_gpu_parity import make_ctx  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu
FIELDS = ("sink_track_id", "snapshot_index_of_sink", "is_merged")


@pytest.mark.parametrize("periodic", [False, True])
def test_traps_vs_oracle_and_reference(make_ctx, oracle_lib, periodic):
    p = capi.make_params(box_size=62.5, softening=5e-3, periodic=periodic)
    e = capi.make_epoch(0.8, snapshot_index=23)
    snap, no, nl, io = cases.case_traps(periodic=periodic)
    ctx = make_ctx(p)
    got = ctx.detect_traps(e, snap.part_offset, snap.pos_mass, snap.vel, no, nl, io)
    want = po.detect_traps(oracle_lib, "hbto", p, e, snap.part_offset, snap.pos_mass, snap.vel, no, nl, io)
    for f in FIELDS:
        assert np.array_equal(got[f], want[f]), f
    if po.have_ref():
        ref = po.detect_traps(po.load_ref(), "hbtref", p, e, snap.part_offset, snap.pos_mass, snap.vel, no, nl, io)
        for f in FIELDS:
            assert np.array_equal(got[f], ref[f]), f
    assert ((got["sink_track_id"] >= 0) & (io["sink_track_id"] < 0)).sum() >= 3
    # the lists may be truncated to the 20 most bound particles the detection reads
    po20 = np.concatenate([[0], np.cumsum(np.minimum(np.diff(snap.part_offset), 20))]).astype(np.int64)
    keep = np.concatenate([np.arange(snap.part_offset[s], snap.part_offset[s] + po20[s + 1] - po20[s]) for s in range(snap.nsub)])
    short = ctx.detect_traps(e, po20, snap.pos_mass[keep], snap.vel[keep], no, nl, io)
    for f in FIELDS:
        assert np.array_equal(short[f], got[f]), f


def test_traps_many_random_hierarchies(make_ctx, oracle_lib):
    """3000 subhaloes in random forests with random core offsets: a few hundred of each outcome."""
    rng = np.random.default_rng(8)
    nsub = 3000
    sizes = rng.integers(0, 60, nsub)
    parent = np.array([-1 if (s == 0 or rng.random() < 0.1) else int(rng.integers(0, s)) for s in range(nsub)])
    part_offset = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    n = int(part_offset[-1])
    centre = rng.uniform(10, 50, (nsub, 3))
    pm = np.zeros((n, 4), np.float32)
    vel = np.zeros((n, 4), np.float32)
    sub_of = np.repeat(np.arange(nsub), sizes)
    pm[:, :3] = centre[sub_of] + rng.normal(0, 0.01, (n, 3))
    pm[:, 3] = 1e-3 * rng.uniform(0.5, 2, n)
    vel[:, :3] = rng.normal(0, 50, (n, 3))
    io = np.zeros(nsub, capi.TRAPIO_DTYPE)
    io["nbound"] = np.where(rng.random(nsub) < 0.1, np.minimum(sizes, 1), sizes)
    io["sink_track_id"] = np.where(rng.random(nsub) < 0.05, 0, -1)
    io["snapshot_index_of_sink"] = np.where(io["sink_track_id"] >= 0, 2, -1)
    host = np.maximum(parent, 0)
    io["mostbound_pos"] = centre[host] + rng.normal(0, 0.012, (nsub, 3))
    io["mostbound_vel"] = rng.normal(0, 45, (nsub, 3))
    children = [[] for _ in range(nsub)]
    for s, q in enumerate(parent):
        if q >= 0:
            children[q].append(s)
    no = np.concatenate([[0], np.cumsum([len(c) for c in children])]).astype(np.int64)
    nl = np.array([c for cs in children for c in cs], np.int32)
    p = capi.make_params(box_size=62.5, softening=5e-3, periodic=True)
    e = capi.make_epoch(1.0, snapshot_index=9)
    ctx = make_ctx(p)
    got = ctx.detect_traps(e, part_offset, pm, vel, no, nl, io)
    want = po.detect_traps(oracle_lib, "hbto", p, e, part_offset, pm, vel, no, nl, io)
    for f in FIELDS:
        assert np.array_equal(got[f], want[f]), f
    new = (got["sink_track_id"] >= 0) & (io["sink_track_id"] < 0)
    assert new.sum() > 200 and (~new & (io["sink_track_id"] < 0) & (parent >= 0)).sum() > 200
