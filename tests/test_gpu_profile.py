"""GPU parity of the post-unbinding properties through the C-ABI ``hbtu_profile_batch``:
Subhalo_t::CalculateProfileProperties / CalculateShape (src/subhalo.cpp:242-398) against the fixtures minted from the
unmodified reference and against the CPU oracle on larger batches."""
import numpy as np
import pytest

import cases
from conftest import load_golden
from hbtplus_b200 import capi, synth
from oracle import pyoracle as po
from test_gpu_parity import make_ctx  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu

INT_FIELDS = ["snapshot_index_of_last_max_vmax"]
# radii are sorted inputs picked by index and Vmax/M200 come from an exact (equal masses) or double cumulative sum:
# bit-exact; the tensors are double sums in a different (parallel) order rounded to float: 1 ulp
EXACT = ["rmax_comoving", "vmax_physical", "last_max_vmax_physical", "r2sigma_comoving", "rhalf_comoving", "bound_m200crit"]


def check_profile(got, want, exact_mass=True):
    for f in INT_FIELDS:
        assert np.array_equal(got[f], want[f]), f
    for f in EXACT:
        if exact_mass:
            assert np.array_equal(got[f], want[f]), f
        else:
            assert np.allclose(got[f], want[f], rtol=3e-7, atol=0), f
    assert np.allclose(got["bound_r200crit_comoving"], want["bound_r200crit_comoving"], rtol=2e-7, atol=0)  # pow()
    for f in ("inertial_tensor", "inertial_tensor_weighted"):
        scale = np.abs(want[f]).max(axis=1, keepdims=True)
        assert np.all(np.abs(got[f] - want[f]) <= 4e-7 * scale + 1e-30), f


@pytest.mark.parametrize("name", list(cases.CASES))
def test_profile_matches_reference_golden(make_ctx, name):
    p, e, _ = cases.CASES[name]()
    _, z = load_golden(name)
    ctx = make_ctx(p)
    got = ctx.profile_batch(e, z["prof_part_offset"], z["prof_pos_mass"], z["prof_io_in"])
    check_profile(got, z["prof_io"], exact_mass=(name != "massvar"))


@pytest.mark.parametrize("periodic", [False, True])
def test_profile_after_unbind_vs_oracle(make_ctx, oracle_lib, periodic):
    """The chain a host runs: unbind on the GPU, then the properties of the bound lists (3e5 particles, 400 subhaloes,
    one of 1.5e5: exercises block-level, warp-level and per-lane reductions and the segmented sort/scan)."""
    rng = np.random.default_rng(31)
    sizes = np.concatenate([[150000], synth.subhalo_sizes(rng, 399, 20, 20000)])
    p = capi.make_params(box_size=62.5, softening=5e-3, periodic=periodic)
    e = capi.make_epoch(0.9, snapshot_index=22)
    snap = synth.make_snapshot(sizes, seed=31, wrap=periodic, centre=[0.1, 30.0, 62.4] if periodic else None, f_contam=0.2)
    ctx = make_ctx(p)
    res = ctx.unbind_batch(e, snap, flags=capi.HBTU_FLAG_TRUNCATE_SOURCE)
    part_offset, pm, io = cases.profile_inputs(snap, res, seed=5)
    got = ctx.profile_batch(e, part_offset, pm, io)
    want = po.profile_batch(oracle_lib, "hbto", p, e, part_offset, pm, io)
    assert (want["bound_m200crit"] != 11.0).sum() > 100 and (want["nbound"] > 1).sum() > 300
    check_profile(got, want)
    st = ctx.stats()
    assert 8 <= st.kernel_launches <= 20  # one batched pass, not one launch per subhalo
    # the fused form: the batch is still resident after stage + execute, only the records travel
    ctx.stage(e, snap, capi.HBTU_FLAG_TRUNCATE_SOURCE)
    ctx.execute()
    res2 = ctx.fetch()
    _, _, io2 = cases.profile_inputs(snap, res2, seed=5)
    fused = ctx.profile_executed(io2)
    st = ctx.stats()
    assert st.h2d_bytes < 200 * snap.nsub
    for f in cases.PROFILE_FIELDS:
        assert np.array_equal(fused[f], got[f], equal_nan=True), f
    with pytest.raises(Exception):  # nothing resident any more after another kind of call
        ctx.profile_batch(e, part_offset, pm, io)
        ctx.profile_executed(io2)


def test_profile_edge_cases(make_ctx, oracle_lib):
    """Nbound 0/1/2, lists longer than Nbound, co-located particles (r = 0 -> clamped to the softening; the weighted tensor of
    a particle AT the centre is 0/0 = NaN in the reference as well), unequal masses, no radius above 200 rho_crit."""
    rng = np.random.default_rng(9)
    p = capi.make_params(box_size=100.0, softening=2e-3, periodic=False)
    e = capi.make_epoch(0.5, snapshot_index=4)
    sizes = [0, 1, 2, 2, 30, 64, 300]
    nbound = [0, 1, 1, 2, 20, 64, 290]
    part_offset = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    pm = np.zeros((part_offset[-1], 4), np.float32)
    io = np.zeros(len(sizes), capi.PROFILEIO_DTYPE)
    for s, n in enumerate(sizes):
        c = rng.uniform(20, 80, 3)
        x = c + rng.normal(0, 0.01 if s != 5 else 50.0, (n, 3))  # s=5: so diffuse that nothing reaches 200 rho_crit
        pm[part_offset[s]:part_offset[s + 1], :3] = x
        pm[part_offset[s]:part_offset[s + 1], 3] = (1e-4 if s != 5 else 1e-8) * rng.uniform(0.5, 2.0, n)
        if n >= 30:
            pm[part_offset[s] + 3, :3] = pm[part_offset[s] + 4, :3]  # a tie in radius
        io["mostbound_pos"][s] = pm[part_offset[s], :3] if n else c
        io["nbound"][s] = nbound[s]
        io["mbound"][s] = pm[part_offset[s]:part_offset[s] + nbound[s], 3].sum()
    pm[part_offset[6] + 7, :3] = pm[part_offset[6], :3]  # a second particle exactly at the centre
    io["bound_r200crit_comoving"], io["bound_m200crit"] = 7.0, 11.0
    ctx = make_ctx(p)
    got = ctx.profile_batch(e, part_offset, pm, io)
    want = po.profile_batch(oracle_lib, "hbto", p, e, part_offset, pm, io)
    assert want["bound_m200crit"][5] == 11.0 and want["bound_m200crit"][1] == 0.0
    assert np.isnan(want["inertial_tensor_weighted"][6]).any()
    nan = np.isnan(want["inertial_tensor_weighted"])
    assert np.array_equal(np.isnan(got["inertial_tensor_weighted"]), nan)
    got["inertial_tensor_weighted"][nan] = 0
    want["inertial_tensor_weighted"][nan] = 0
    check_profile(got, want, exact_mass=False)
    with pytest.raises(Exception):
        bad = io.copy()
        bad["nbound"][4] = 31
        ctx.profile_batch(e, part_offset, pm, bad)
