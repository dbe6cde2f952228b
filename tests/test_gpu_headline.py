"""Parity in the HEADLINE configuration's parameter regime (BASELINE.json configs[1], AqA2-shaped): BoxSize 100,
SofteningHalo 4.8e-5 (TreeNodeResolution 4.8e-6 is below the fp32 ulp at x ~ 50), particle mass 1e-6, exact potential,
and at production sizes: default kernel routing (no HBTU_* forcing), so the masked group walk and its >= 2^19 / 2^20 target
classes are the kernels under test.  Everything is compared with the UNMODIFIED reference (oracle/_ref) and the oracle.
Plus the cfg-4-shaped V64 forest (EAGLE ABI: -DHBT_INT8) through the drop-in seam."""
import os

import numpy as np
import pytest

import cases
from hbtplus_b200 import capi, synth
from oracle import pyoracle as po
from test_gpu_parity import POT_TOL, check_batch, make_ctx  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu

BOX, EPS, MP = 100.0, 4.8e-5, 1e-6  # bench.py's constants


def bench_params(**kw):
    return capi.make_params(box_size=BOX, softening=EPS, periodic=False, max_sample_size=0, **kw)


def test_potential_1e6_subhalo_bench_regime(make_ctx, oracle_lib):
    """One 1e6-particle source at the bench's softening: per-particle potentials against the reference's own
    GravityTree_t::EvaluatePotential (src/gravity_tree.cpp:79-164) on 1.5e5 of the targets, and the device-counted accepted
    interactions of ALL targets against the oracle's instrumented walk (same decisions <=> same count)."""
    if not po.have_ref():
        pytest.skip("oracle/_ref not built")
    for k in list(os.environ):
        assert not k.startswith("HBTU_WALK"), "default kernel routing is what this test is about"
    n = 1_000_000
    p, e = bench_params(), capi.make_epoch(1.0)
    snap = synth.make_snapshot([n], seed=20240002, box_size=BOX, particle_mass=MP, wrap=False, centre=[BOX / 2] * 3, f_contam=0.2)
    pm = snap.pos_mass
    ctx = make_ctx(p)
    ctx.set_counting(True)
    got = ctx.tree_potential(e, pm, pm, self_mass=pm[:, 3].copy())
    st = ctx.stats()
    ctx.set_counting(False)
    assert st.walk_fallbacks == 0
    ref = po.load_ref()
    ref.hbtref_set_num_threads(os.cpu_count() or 1)
    pick = np.random.default_rng(1).choice(n, 150_000, replace=False)
    want = po.tree_potential(ref, "hbtref", p, e, pm, pm[pick], self_mass=pm[pick, 3].copy())
    rel = np.abs(got[pick] - want) / np.abs(want)
    cases.report("potential_1e6_bench_regime", max_rel_err=float(rel.max()), mean_rel_err=float(rel.mean()), frac_above_1e_4=float(np.mean(rel > 1e-4)),
                 interactions_per_target=st.pair_interactions / n)
    assert rel.max() <= POT_TOL
    assert np.mean(rel > 1e-4) < 1e-3  # a flipped fp32 criterion decision costs that cell's multipole error; rare
    oracle_lib.hbto_set_num_threads(os.cpu_count() or 1)
    import ctypes as C
    acc = np.zeros(n, np.int64)
    P = capi._ptr
    oracle_lib.hbto_walk_counts(C.byref(p), C.byref(e), n, P(pm, C.c_float), n, P(pm, C.c_float), P(acc, C.c_int64), None)
    want_inter = int(acc.sum())
    cases.report("interactions_1e6_bench_regime", device=int(st.pair_interactions), oracle=want_inter)
    assert abs(st.pair_interactions - want_inter) <= 1e-6 * want_inter


def test_unbind_3e5_subhalo_bench_regime_vs_reference(make_ctx):
    """A full Subhalo_t::Unbind of one 3e5-particle source (+ a nested 2e4 one feeding it) at the bench's parameters against
    the unmodified reference with all host threads (6.7 s on 8 cores for 2e5)."""
    if not po.have_ref():
        pytest.skip("oracle/_ref not built")
    p, e = bench_params(), capi.make_epoch(1.0)
    snap = synth.make_snapshot([300_000, 20_000, 150], seed=20240012, box_size=BOX, particle_mass=MP, wrap=False, centre=[BOX / 2] * 3,
                               parent=[-1, 0, 1], f_contam=0.25)
    ctx = make_ctx(p)
    ctx.set_counting(True)
    got = ctx.unbind_batch(e, snap, flags=capi.HBTU_FLAG_TRUNCATE_SOURCE)
    st = ctx.stats()
    ctx.set_counting(False)
    assert st.walk_fallbacks == 0
    ref = po.load_ref()
    ref.hbtref_set_num_threads(os.cpu_count() or 1)
    want = po.run_batch(ref, "hbtref", p, e, snap, flags=capi.HBTU_FLAG_TRUNCATE_SOURCE)
    assert want.io["nbound"][0] > 150_000
    check_batch(snap, got, want, name="unbind_3e5_bench_regime")
    # per-particle binding energies of the bound part (SAVE_BINDING_ENERGY): same particles, E to the potential gate
    nb = int(min(got.io["nbound"][0], want.io["nbound"][0]))
    eg = dict(zip(got.bound(0).tolist(), got.energy[got.order_offset[0]:got.order_offset[0] + nb].tolist()))
    ew = dict(zip(want.bound(0).tolist(), want.energy[want.order_offset[0]:want.order_offset[0] + nb].tolist()))
    common = sorted(set(eg) & set(ew))
    a, b = np.array([eg[k] for k in common]), np.array([ew[k] for k in common])
    scale = np.abs(b).max()
    assert len(common) > 0.999 * nb and np.max(np.abs(a - b)) <= POT_TOL * scale


def cfg4_forest(ngroups=1200, seed=20240004):
    """EagleL100N1504-shaped (SURVEY.md 8(d) cfg 4) at test size: FoF groups with dN/dn ~ n^-1.9, a central + satellites
    (some nested twice) per group, field subhaloes, BoxSize 67.77, eps 1.80239e-3, periodic, DM particle mass 6.57e-4."""
    rng = np.random.default_rng(seed)
    sizes, parent, host = [], [], []
    for g in range(ngroups):
        c = len(sizes)
        sizes.append(int(synth.subhalo_sizes(rng, 1, 40, 20000)[0]))
        parent.append(-1)
        host.append(g)
        nsat = int(min(rng.poisson(1.2), 6))
        for k in range(nsat):
            s = len(sizes)
            sizes.append(int(min(synth.subhalo_sizes(rng, 1, 20, 3000)[0], max(20, sizes[c] // 2))))
            parent.append(c)
            host.append(g)
            if rng.random() < 0.25:
                sizes.append(int(rng.integers(20, 40)))
                parent.append(s)
                host.append(g)
    for k in range(ngroups // 10):  # field subhaloes
        sizes.append(int(synth.subhalo_sizes(rng, 1, 20, 500)[0]))
        parent.append(-1)
        host.append(-1)
    snap = synth.make_snapshot(sizes, seed=seed, box_size=67.77, particle_mass=6.57e-4, wrap=True, parent=parent, f_contam=0.25)
    return snap, np.asarray(host, np.int32), ngroups, np.asarray(sizes, np.float32)


def test_cfg4_v64_forest_drop_in():
    """>= 1e3 FoF groups through the reference's own RefineParticles seam in the V64 ABI (-DHBT_INT8: HBTInt = long,
    Particle_t with Type - what EagleL100N1504 / HBT.apostle builds use): libhbtref_v64 (CPU) vs libhbtdropin_v64 (GPU)."""
    if not po.have_dropin("v64"):
        pytest.skip("oracle/_ref libraries not built")
    ref, drop = po.load_ref_variant("v64"), po.load_dropin("v64")
    assert ref.hbtref_sizeof_hbtint() == 8 and ref.hbtref_sizeof_particle() == 40
    ref.hbtref_set_num_threads(os.cpu_count() or 1)
    p = capi.make_params(box_size=67.77, softening=1.80239e-3, periodic=True)
    e = capi.make_epoch(1.0, snapshot_index=28)
    snap, host, nhalos, mb = cfg4_forest()
    assert nhalos >= 1000 and snap.nsub > 2000
    want = po.refine_particles(ref, p, e, snap, host, snap.nsub, nhalos, mb)
    got = po.refine_particles(drop, p, e, snap, host, snap.nsub, nhalos, mb)
    check_batch(snap, got, want, name="cfg4_v64_forest")
    assert (want.io["nbound"] > 1).sum() > 0.6 * snap.nsub


def test_plain_unbind_flag_vs_oracle(make_ctx, oracle_lib):
    """HBTU_SUB_PLAIN_UNBIND: roots the reference enters through plain Unbind have no orphan rule (src/subhalo_unbind.cpp:498-510)."""
    from test_oracle import plain_unbind_case

    p = capi.make_params(box_size=62.5, softening=5e-3, periodic=False)
    e = capi.make_epoch(0.9, snapshot_index=7)
    snap = plain_unbind_case()
    ctx = make_ctx(p)
    got = ctx.unbind_batch(e, snap)
    want = po.run_batch(oracle_lib, "hbto", p, e, snap)
    check_batch(snap, got, want, exact_counts=True, name="plain_unbind")
    assert got.io["nbound"][0] > 100 and not np.array_equal(got.particles(0), np.arange(400))
    snap.io["flags"] = 0  # RecursiveUnbind semantics: sub 0 (entry Nbound = 1) is an orphan and keeps its list
    got0 = ctx.unbind_batch(e, snap)
    want0 = po.run_batch(oracle_lib, "hbto", p, e, snap)
    assert np.array_equal(got0.particles(0), np.arange(400)) and np.array_equal(want0.particles(0), np.arange(400))
    assert np.array_equal(got0.io["nbound"], want0.io["nbound"])
