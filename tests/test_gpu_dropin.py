"""Drop-in test at the reference's own seam: the SAME harness (oracle/ref_build/ref_capi.cpp) drives
``SubhaloSnapshot_t::RefineParticles()`` on an in-memory snapshot once with the reference's subhalo_unbind.o
(libhbtref_v32.so, CPU) and once with integration/subhalo_unbind_b200.o + libhbtunbind.so (libhbtdropin_v32.so, GPU).
Host haloes with central + heads + old nests, field subhaloes and new-born subhaloes are all present."""
import numpy as np
import pytest

import cases
from hbtplus_b200 import capi, synth
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def snapshot_with_hosts(seed, periodic):
    #        halo 0: 0 central, 1 & 2 heads, 3 in 1, 4 in 3 | halo 1: 5 central (old nest 7), 6 head | field 8, 9, 10 | new-born 11 (halo 1), 12 (field), 13 (halo 0)
    sizes = [6000, 900, 400, 120, 45, 2500, 300, 150, 700, 30, 18, 260, 90, 55]
    parent = [-1, -1, -1, 1, 3, -1, -1, 5, -1, -1, -1, -1, -1, -1]
    host = [0, 0, 0, 0, 0, 1, 1, 1, -1, -1, -1, 1, -1, 0]
    n_old = 11
    snap = synth.make_snapshot(sizes, seed=seed, parent=parent, wrap=periodic, f_contam=0.3)
    # satellites of a host sit inside its central (their unbound tails are what the central collects)
    mb = np.asarray(sizes, np.float32)
    return snap, np.asarray(host, np.int32), n_old, 2, mb


@pytest.mark.parametrize("variant", ["v32", "v64", "v32ns", "v32th", "v32im", "v64r8"])
@pytest.mark.parametrize("periodic", [False, True])
def test_refine_particles_drop_in(periodic, variant):
    """v32 = -DDM_ONLY (HBTInt=int); v64 = -DHBT_INT8 (HBTInt=long, Particle_t with Type: the CMake default / EAGLE ABI);
    v32ns = -DNO_STRIPPING; v32th = -DUNBIND_WITH_THERMAL_ENERGY (Particle_t with InternalEnergy + Type);
    v32im = -DINCLUSIVE_MASS (flat RefineParticles loop); v64r8 = -DHBT_INT8 -DHBT_REAL8 (HBTReal = double: the reference then
    does its HBTReal arithmetic in double, the shim narrows the particle arrays to float4 - same gates).  The CUDA library is the same binary for all of them; only the
    shim is compiled with the caller's -D flags and passes the physics variant as batch flags."""
    if not po.have_dropin(variant):
        pytest.skip("oracle/_ref libraries not built (reference sources absent at build time)")
    ref, drop = po.load_ref_variant(variant), po.load_dropin(variant)
    assert ref.hbtref_sizeof_hbtint() == (8 if variant in ("v64", "v64r8") else 4)
    assert ref.hbtref_sizeof_hbtreal() == (8 if variant == "v64r8" else 4)
    ref.hbtref_set_num_threads(4)
    p = capi.make_params(box_size=62.5, softening=5e-3, periodic=periodic)
    e = capi.make_epoch(0.9, snapshot_index=15)
    snap, host, n_old, nhalos, mb = snapshot_with_hosts(31 + periodic, periodic)
    want = po.refine_particles(ref, p, e, snap, host, n_old, nhalos, mb)
    got = po.refine_particles(drop, p, e, snap, host, n_old, nhalos, mb)
    skip = cases.unbound_inputs(snap)
    for f in ("snapshot_index_of_death", "snapshot_index_of_sink", "sink_track_id"):
        assert np.array_equal(got.io[f], want.io[f]), f
    assert np.all(np.abs(got.io["nbound"] - want.io["nbound"]) <= np.maximum(1, 2e-4 * want.io["nbound"]))
    assert np.array_equal(got.io["nbound"] > 1, want.io["nbound"] > 1)
    assert np.allclose(got.io["mbound"][~skip], want.io["mbound"][~skip], rtol=1e-3)
    for f in ("avg_pos", "avg_vel", "mostbound_pos", "mostbound_vel"):
        assert np.allclose(got.io[f][~skip], want.io[f][~skip], rtol=2e-6, atol=1e-6), f
    for s in range(snap.nsub):
        assert cases.jaccard(got.bound(s), want.bound(s)) >= 0.999, s
        assert cases.jaccard(got.particles(s), want.particles(s)) >= 0.999, s
    # the central of halo 0 was fed by its heads (1, 2) and, through them, by 3 and 4
    assert want.io["nsource_full"][0] == 0 or want.io["nsource"][0] >= want.io["nbound"][0]
    assert (want.io["nbound"][[0, 5]] > 1000).all()


@pytest.mark.parametrize("variant", ["v32", "v64", "v64r8"])
def test_profile_properties_drop_in(variant):
    """SURVEY.md 8(f) next-2 through the reference-facing side: the harness fills the reference's own Subhalo_t objects and
    calls either Subhalo_t::CalculateProfileProperties/CalculateShape (libhbtref) or the shim's batched replacement of the
    loop at src/subhalo_tracking.cpp:901-906 (libhbtdropin -> hbtu_profile_batch)."""
    if not po.have_dropin(variant):
        pytest.skip("oracle/_ref libraries not built (reference sources absent at build time)")
    ref, drop = po.load_ref_variant(variant), po.load_dropin(variant)
    ref.hbtref_set_num_threads(1)
    p = capi.make_params(box_size=62.5, softening=5e-3, periodic=True)
    e = capi.make_epoch(0.9, snapshot_index=15)
    snap, host, n_old, nhalos, mb = snapshot_with_hosts(21, True)
    res = po.refine_particles(ref, p, e, snap, host, n_old, nhalos, mb)
    part_offset, pm, io = cases.profile_inputs(snap, res, seed=3)
    want = po.profile_batch(ref, "hbtref", p, e, part_offset, pm, io)
    got = po.profile_batch(drop, "hbtref", p, e, part_offset, pm, io)
    from test_gpu_profile import check_profile
    check_profile(got, want, exact_mass=(variant != "v64r8"))  # HBTReal = double: the reference's radii are rounded once more
    assert (want["nbound"] > 1).sum() >= 5


@pytest.mark.parametrize("variant", ["v32", "v64"])
def test_mask_subhalos_drop_in(variant):
    """SURVEY.md 8(f) next-1 through the reference-facing side: the harness builds a SubhaloSnapshot_t + MemberTable and calls
    either the reference's private SubhaloSnapshot_t::MaskSubhalos (libhbtref) or the shim's HBT_B200_MaskSubhalos
    (libhbtdropin -> hbtu_mask_batch), which shrinks the same vector<Particle_t> lists."""
    if not po.have_dropin(variant):
        pytest.skip("oracle/_ref libraries not built (reference sources absent at build time)")
    ref, drop = po.load_ref_variant(variant), po.load_dropin(variant)
    p = capi.make_params(box_size=62.5, softening=5e-3)
    part_offset, ids, nest_offset, nest_list, nbound = cases.case_mask(seed=11, nroots=12)
    if variant == "v32":
        ids = ids % (2**31 - 1)  # HBTInt = int
    want = po.mask_batch(ref, "hbtref", p, part_offset, ids, nest_offset, nest_list, nbound)
    got = po.mask_batch(drop, "hbtref", p, part_offset, ids, nest_offset, nest_list, nbound)
    assert np.array_equal(got[0], want[0]) and 0 < want[0].sum() < part_offset[-1]
    for s in range(len(nbound)):
        b = part_offset[s]
        assert np.array_equal(got[1][b:b + got[0][s]], want[1][b:b + want[0][s]]), s


def test_refine_particles_sharded_over_devices(monkeypatch):
    """SURVEY.md 8(e) at the shim: with HBT_UNBIND_DEVICES the rank's hierarchies are dealt to several contexts (here three
    contexts on device 0, so that one GPU suffices), each on its own host thread; the result must not depend on the split."""
    if not po.have_dropin("v32"):
        pytest.skip("oracle/_ref libraries not built (reference sources absent at build time)")
    drop = po.load_dropin("v32")
    p = capi.make_params(box_size=62.5, softening=5e-3, periodic=True)
    e = capi.make_epoch(0.9, snapshot_index=15)
    snap, host, n_old, nhalos, mb = snapshot_with_hosts(21, True)
    monkeypatch.delenv("HBT_UNBIND_DEVICES", raising=False)
    one = po.refine_particles(drop, p, e, snap, host, n_old, nhalos, mb)
    monkeypatch.setenv("HBT_UNBIND_DEVICES", "0,0,0")
    three = po.refine_particles(drop, p, e, snap, host, n_old, nhalos, mb)
    # bit for bit: every fp64 sum / scan of the path has a summation tree aligned to the subhalo, not to the batch
    for f in one.io.dtype.names:
        assert np.array_equal(one.io[f], three.io[f]), f
    assert np.array_equal(one.order_offset, three.order_offset)
    ntot = int(one.order_offset[-1])
    assert np.array_equal(one.order[:ntot], three.order[:ntot])
    bad = np.nonzero(one.energy[:ntot] != three.energy[:ntot])[0]
    assert len(bad) == 0, (bad[:10], one.energy[bad[:10]], three.energy[bad[:10]], np.searchsorted(one.order_offset, bad[:10], side='right') - 1, one.io['nbound'], one.io['nsource'])


@pytest.mark.parametrize("variant", ["v32", "v64"])
def test_detect_traps_drop_in(variant):
    """SURVEY.md 8(f) next-3 through the reference-facing side: SubhaloSnapshot_t::MergeSubhalos' detection (libhbtref, with
    MergeTrappedSubhalos off) against the shim's HBT_B200_DetectTraps (libhbtdropin -> hbtu_detect_traps)."""
    if not po.have_dropin(variant):
        pytest.skip("oracle/_ref libraries not built (reference sources absent at build time)")
    ref, drop = po.load_ref_variant(variant), po.load_dropin(variant)
    p = capi.make_params(box_size=62.5, softening=5e-3, periodic=True)
    e = capi.make_epoch(0.8, snapshot_index=23)
    snap, no, nl, io = cases.case_traps(periodic=True)
    want = po.detect_traps(ref, "hbtref", p, e, snap.part_offset, snap.pos_mass, snap.vel, no, nl, io)
    got = po.detect_traps(drop, "hbtref", p, e, snap.part_offset, snap.pos_mass, snap.vel, no, nl, io)
    for f in ("sink_track_id", "snapshot_index_of_sink", "is_merged"):
        assert np.array_equal(got[f], want[f]), f
    assert (want["sink_track_id"] >= 0).sum() >= 4


def _check_merge(got, gm, want, wm, snap):
    assert np.array_equal(gm, wm)
    for f in ("snapshot_index_of_death", "snapshot_index_of_sink", "sink_track_id", "nsource", "nsource_full"):
        assert np.array_equal(got.io[f], want.io[f]), f
    assert np.all(np.abs(got.io["nbound"] - want.io["nbound"]) <= np.maximum(1, 2e-4 * want.io["nbound"]))
    assert np.array_equal(got.io["nbound"] > 1, want.io["nbound"] > 1)
    hosts = np.nonzero(wm)[0]
    assert np.allclose(got.io["mbound"][hosts], want.io["mbound"][hosts], rtol=1e-3)
    for f in ("avg_pos", "avg_vel", "mostbound_pos", "mostbound_vel"):
        assert np.allclose(got.io[f][hosts], want.io[f][hosts], rtol=2e-6, atol=1e-6), f
    for s in range(snap.nsub):
        assert cases.jaccard(got.bound(s), want.bound(s)) >= 0.999, s
        assert cases.jaccard(got.particles(s), want.particles(s)) >= 0.999, s


@pytest.mark.parametrize("variant", ["v32", "v64"])
def test_merge_subhalos_drop_in(variant):
    """SURVEY.md 8(b) threading + 8(f) next-3: the reference's UNMODIFIED SubhaloSnapshot_t::MergeSubhalos with
    MergeTrappedSubhalos on calls Subhalo_t::Unbind from its OpenMP loop (src/subhalo_merge.cpp:207-210).  In libhbtdropin that
    is the shim's Unbind, entered here by 8 concurrent threads (combined into batches under the device mutex); the catalogue must
    equal the reference's.  Then the patched sequence with ONE batch (HBT_B200_UnbindMerged) must give the same again."""
    if not po.have_dropin(variant):
        pytest.skip("oracle/_ref libraries not built (reference sources absent at build time)")
    ref, drop = po.load_ref_variant(variant), po.load_dropin(variant)
    p = capi.make_params(box_size=62.5, softening=5e-3, periodic=True)
    e = capi.make_epoch(0.8, snapshot_index=23)
    snap = cases.case_merge()
    want, wm = po.merge_subhalos(ref, p, e, snap, mode=0, nthreads=4)
    assert wm.sum() >= 8 and (want.io["nbound"] != snap.io["nbound"]).sum() >= 8
    for rep in range(3):  # the race, if any, is timing dependent
        got, gm = po.merge_subhalos(drop, p, e, snap, mode=0, nthreads=8)
        _check_merge(got, gm, want, wm, snap)
    one, om = po.merge_subhalos(drop, p, e, snap, mode=1, nthreads=8)
    _check_merge(one, om, want, wm, snap)
    # one batch or many: bit-identical records
    for f in got.io.dtype.names:
        assert np.array_equal(one.io[f], got.io[f]), f
