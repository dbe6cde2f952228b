"""Parity tests proper: the CUDA path, called through the C-ABI (ctypes on libhbtunbind.so), against
(a) the golden fixtures generated from the unmodified reference and (b) the oracle on fresh seeded inputs.

Gates (BASELINE.json north_star): per-particle potential rel. err <= 1e-3; per-subhalo bound mass within 0.1 %;
bound-membership Jaccard >= 0.999; subhalo survival (Nbound >= MinNumPartOfSub decisions, death flags) bit-exact.
Observed on B200 and asserted below where deterministic: potentials agree to <= 3e-4 (mean ~1e-7), Nbound/death/frames exactly."""
import os
import subprocess
import sys

import numpy as np
import pytest

import cases
from conftest import load_golden, orders_equal_modulo_ties
from hbtplus_b200 import capi, synth
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

POT_TOL = 1e-3  # north_star gate
POT_OBSERVED = 3e-4  # a criterion decision that flips in fp32 costs that cell's multipole error


@pytest.fixture(scope="module")
def make_ctx():
    from hbtplus_b200.unbind import UnbindContext

    made = []

    def factory(params):
        ctx = UnbindContext(params)
        made.append(ctx)
        return ctx

    yield factory
    for c in made:
        c.close()


JACCARD_MISS_CEILING = 2e-3  # hard ceiling of the fraction of subhaloes allowed below Jaccard 0.999 (each by <= 2 particles)


def check_batch(snap, got, want, exact_frames=True, exact_counts=False, name=None):
    """The north-star gates.  Survival (Nbound > 1, death/sink flags: what subhalo counts and track IDs depend on) must be
    bit-exact; bound mass within 0.1 %; membership Jaccard >= 0.999.  A particle whose |E| is within fp32 round-off of 0
    may flip (observed: ~1 per 1e6 particle evaluations), so Nbound itself is compared to 2e-4 unless exact_counts is
    asked for (the golden fixtures, where no flip occurs).  A subhalo of < 2000 particles cannot lose one particle and
    keep Jaccard >= 0.999; such MISSES are counted and reported (cases.report -> gpurun_out/parity_stats.jsonl), never
    hidden: the miss rate has the hard ceiling JACCARD_MISS_CEILING, every miss must be a flip of <= 2 particles in a
    subhalo of < 2000 particles, and with exact_counts no miss is allowed at all.  The mass gate is applied to subhaloes
    with unchanged Nbound and to all subhaloes above 2000 particles.
    Average positions are compared as they are, NOT modulo the box: in periodic runs the image the reference reports
    depends on which particle its hole-based partition leaves in Elist[0] (src/subhalo_unbind.cpp:21-58,152-154), and the
    library tracks exactly that (unbind_batch.cu, rho_*)."""
    skip = cases.unbound_inputs(snap)
    for f in ("snapshot_index_of_death", "snapshot_index_of_sink", "sink_track_id"):
        assert np.array_equal(got.io[f], want.io[f]), f
    nb_g, nb_w = got.io["nbound"], want.io["nbound"]
    assert np.array_equal(nb_g > 1, nb_w > 1) and np.array_equal(nb_g == 0, nb_w == 0)
    if exact_counts:
        for f in ("nbound", "nsource", "nsource_full"):
            assert np.array_equal(got.io[f], want.io[f]), f
    else:
        assert np.all(np.abs(nb_g - nb_w) <= np.maximum(2, 2e-4 * nb_w))
        assert np.mean(nb_g == nb_w) > 0.99
        for f in ("nsource", "nsource_full"):
            assert np.all(np.abs(got.io[f] - want.io[f]) <= np.maximum(8, 1e-3 * want.io[f])), f
    same = nb_g == nb_w
    big = nb_w >= 2000
    gate = ~skip & (same | big)
    mb_g, mb_w = got.io["mbound"][gate], want.io["mbound"][gate]
    assert np.all(np.abs(mb_g - mb_w) <= 1e-3 * np.abs(mb_w))  # gate: 0.1 %
    misses, min_j = [], 1.0
    for s in range(snap.nsub):
        jb = cases.jaccard(got.bound(s), want.bound(s))
        min_j = min(min_j, jb)
        if jb < 0.999:
            assert not exact_counts and nb_w[s] < 2000, s
            assert len(set(got.bound(s).tolist()) ^ set(want.bound(s).tolist())) <= 2, s
            misses.append(s)
        assert cases.jaccard(got.particles(s), want.particles(s)) >= (0.999 if jb >= 0.999 else 0.97), s
    live = int((nb_w > 1).sum())
    stats = {"subhaloes": int(snap.nsub), "live": live, "frac_identical_nbound": float(np.mean(same)), "jaccard_misses": len(misses),
             "jaccard_miss_rate": len(misses) / max(snap.nsub, 1), "min_jaccard": float(min_j),
             "largest_miss_nbound": int(max((nb_w[s] for s in misses), default=0)),
             "max_rel_dmbound": float(np.max(np.abs(mb_g - mb_w) / np.abs(mb_w))) if mb_w.size else 0.0}
    if name:
        cases.report(name, **stats)
    assert len(misses) <= max(1, JACCARD_MISS_CEILING * snap.nsub), stats
    sel = ~skip & same if not exact_counts else ~skip
    for f in ("avg_pos", "avg_vel", "mostbound_pos", "mostbound_vel"):
        a, b = got.io[f][sel].astype(np.float64), want.io[f][sel].astype(np.float64)
        if exact_frames:
            assert np.allclose(a, b, rtol=2e-6, atol=1e-6), f
    for f in ("specific_self_potential_energy", "specific_self_kinetic_energy", "specific_angular_momentum"):
        a, b = got.io[f][~skip & same], want.io[f][~skip & same]
        assert np.allclose(a, b, rtol=2e-4, atol=1e-3 * np.abs(b).max() if b.size else 0), f
    return stats


@pytest.mark.parametrize("name", list(cases.VARIANT_CASES))
@pytest.mark.parametrize("trunc", [0, capi.HBTU_FLAG_TRUNCATE_SOURCE])
def test_physics_variants_match_reference_golden(make_ctx, name, trunc):
    """HBTU_FLAG_NO_STRIPPING / HBTU_FLAG_THERMAL_ENERGY against fixtures minted from the reference compiled with
    -DNO_STRIPPING / -DUNBIND_WITH_THERMAL_ENERGY (SURVEY.md 8(b) compile-time variants)."""
    fn, _, vflag = cases.VARIANT_CASES[name]
    p, e, _ = fn()
    snap, z = load_golden(name)
    tag = "trunc" if trunc else "full"
    ctx = make_ctx(p)
    got = ctx.unbind_batch(e, snap, flags=vflag | trunc)
    want = po.Result(z[f"{tag}_io"], z[f"{tag}_order_offset"], z[f"{tag}_order"], z[f"{tag}_energy"])
    check_batch(snap, got, want, exact_counts=True)
    assert np.array_equal(got.order_offset, want.order_offset)
    for s in range(snap.nsub):
        nb = int(want.io["nbound"][s])
        gp, wp = got.particles(s), want.particles(s)
        ew_all = want.energy[want.order_offset[s]:][:len(wp)]
        if name == "nostrip" and nb > 1:
            # everything is kept and E-sorted, so entries with E ~ 0 are inside the compared range: two neighbours may
            # swap where their reference energies differ by less than the fp32 round-off of the potential
            assert sorted(gp.tolist()) == sorted(wp.tolist())
            pos_w = {int(q): i for i, q in enumerate(wp)}
            scale = np.abs(ew_all[:nb]).max()
            for i in np.nonzero(gp != wp)[0]:
                j = pos_w[int(gp[i])]
                assert abs(j - i) <= 4 and abs(ew_all[i] - ew_all[j]) <= 1e-5 * scale, (s, i, j)
        else:
            assert orders_equal_modulo_ties(gp, wp, ew_all, nb), s
        if nb > 1:
            eg = got.energy[got.order_offset[s]:][:nb]
            assert np.allclose(eg, ew_all[:nb], rtol=2e-4, atol=2e-5 * np.abs(ew_all[:nb]).max())


EXACT_CASES = [c for c in cases.CASES if c != "sampled"]  # the sampled case depends on the shuffle stream: see below


@pytest.mark.parametrize("name", EXACT_CASES)
@pytest.mark.parametrize("tag,flags", [("full", 0), ("trunc", capi.HBTU_FLAG_TRUNCATE_SOURCE)])
def test_unbind_matches_reference_golden(make_ctx, name, tag, flags):
    p, e, _ = cases.CASES[name]()
    snap, z = load_golden(name)
    ctx = make_ctx(p)
    got = ctx.unbind_batch(e, snap, flags=flags)
    want = po.Result(z[f"{tag}_io"], z[f"{tag}_order_offset"], z[f"{tag}_order"], z[f"{tag}_energy"])
    check_batch(snap, got, want, exact_counts=True)
    assert np.array_equal(got.order_offset, want.order_offset)
    for s in range(snap.nsub):
        nb = int(want.io["nbound"][s])
        b = want.order_offset[s]
        assert orders_equal_modulo_ties(got.particles(s), want.particles(s), want.energy[b:], nb), (name, s)
        if nb > 1:  # SAVE_BINDING_ENERGY energies of the bound part
            eg, ew = np.sort(got.energy[b:b + nb]), np.sort(want.energy[b:b + nb])
            assert np.all(np.abs(eg - ew) <= 1e-4 * np.abs(ew).max())


@pytest.mark.parametrize("name", list(cases.CASES))
def test_tree_potential_matches_reference_golden(make_ctx, name):
    p, e, _ = cases.CASES[name]()
    snap, z = load_golden(name)
    ctx = make_ctx(p)
    b, en = z["pot_src_range"]
    src, tgt = snap.pos_mass[b:en], z["pot_tgt"]
    got = ctx.tree_potential(e, src, tgt, self_mass=tgt[:, 3].copy())
    rel = np.abs(got - z["pot_self"]) / np.abs(z["pot_self"])
    assert rel.max() <= POT_OBSERVED <= POT_TOL
    got = ctx.tree_potential(e, src, tgt + np.float32([0.01, 0.0, -0.02, 0.0]))
    assert (np.abs(got - z["pot_foreign"]) / np.abs(z["pot_foreign"])).max() <= POT_OBSERVED
    s = int(np.argmax(np.diff(snap.part_offset)))
    got = ctx.tree_potential(e, src, tgt, self_mass=tgt[:, 3].copy(), tgt_vel=z["be_vel"], ref_pos=snap.io["avg_pos"][s], ref_vel=snap.io["avg_vel"][s])
    assert np.all(np.abs(got - z["be"]) <= POT_OBSERVED * np.abs(z["pot_self"]))


@pytest.mark.parametrize("periodic", [False, True])
@pytest.mark.parametrize("n", [1, 2, 3, 37, 1000, 60000])
def test_tree_potential_vs_oracle(make_ctx, oracle_lib, n, periodic):
    p = capi.make_params(box_size=62.5, softening=5e-3, periodic=periodic)
    e = capi.make_epoch(1.0)
    snap = synth.make_snapshot([n], seed=n + 3, wrap=periodic, centre=[0.1, 31, 62.4] if periodic else None)
    pm = snap.pos_mass
    ctx = make_ctx(p)
    ctx.set_counting(True)
    got = ctx.tree_potential(e, pm, pm, self_mass=pm[:, 3].copy())
    st = ctx.stats()
    ctx.set_counting(False)
    want = po.tree_potential(oracle_lib, "hbto", p, e, pm, pm, self_mass=pm[:, 3].copy())
    if n > 1:
        assert (np.abs(got - want) / np.abs(want)).max() <= POT_OBSERVED
    # the walk makes the reference's accept/open decisions: same number of accepted interactions
    ref_inter = oracle_lib.hbto_last_interactions()
    assert abs(st.pair_interactions - ref_inter) <= 1e-5 * ref_inter + 2


def test_random_forests_vs_oracle(make_ctx, oracle_lib):
    rng = np.random.default_rng(77)
    for trial in range(4):
        nsub = 40
        sizes = synth.subhalo_sizes(rng, nsub, 10, 4000)
        sizes[rng.integers(0, nsub, 3)] = [0, 1, 19]
        parent = np.full(nsub, -1)
        for s in range(1, nsub):
            if rng.random() < 0.6:
                parent[s] = rng.integers(0, s)
        periodic = bool(trial % 2)
        p = capi.make_params(box_size=62.5, softening=5e-3, periodic=periodic)
        e = capi.make_epoch([1.0, 0.7, 0.5, 0.9][trial], snapshot_index=20 + trial)
        snap = synth.make_snapshot(sizes, seed=500 + trial, parent=parent, wrap=periodic, f_contam=[0.2, 0.4, 0.1, 0.6][trial])
        orphans = rng.integers(0, nsub, 2)
        snap.io["nbound"][orphans] = 1
        flags = capi.HBTU_FLAG_TRUNCATE_SOURCE if trial >= 2 else 0
        ctx = make_ctx(p)
        got = ctx.unbind_batch(e, snap, flags=flags)
        want = po.run_batch(oracle_lib, "hbto", p, e, snap, flags=flags)
        check_batch(snap, got, want)


def test_stage_execute_fetch_is_repeatable(make_ctx, oracle_lib):
    """hbtu_execute does not modify the staged inputs: two executions give identical outputs."""
    p, e, snap = cases.case_nested()
    ctx = make_ctx(p)
    ctx.stage(e, snap, flags=capi.HBTU_FLAG_TRUNCATE_SOURCE)
    ctx.execute()
    a = ctx.fetch()
    ctx.execute()
    b = ctx.fetch()
    ntot = int(a.order_offset[-1])
    assert np.array_equal(a.order_offset, b.order_offset)
    assert np.array_equal(a.order[:ntot], b.order[:ntot]) and np.array_equal(a.energy[:ntot], b.energy[:ntot])
    for f in a.io.dtype.names:
        assert np.array_equal(a.io[f], b.io[f]), f
    want = po.run_batch(oracle_lib, "hbto", p, e, snap, flags=capi.HBTU_FLAG_TRUNCATE_SOURCE)
    check_batch(snap, a, want)


def test_errors_are_reported_not_swallowed(make_ctx):
    from hbtplus_b200.unbind import UnbindError

    p, e, snap = cases.case_nested()
    ctx = make_ctx(p)
    with pytest.raises(UnbindError):
        ctx.execute()  # nothing staged
    bad = synth.Snapshot(snap.part_offset, snap.pos_mass, snap.vel, snap.nest_offset, snap.nest_list.copy(), snap.io)
    bad.nest_list[0] = bad.nest_list[1]
    with pytest.raises(UnbindError) as ei:
        ctx.unbind_batch(e, bad)
    assert ei.value.code == capi.HBTU_ERR_INVALID


def test_full_size_properties(make_ctx):
    """Size-independent invariants at a size the CPU oracle cannot check quickly (3e6 particles):
    the output is a permutation; the bound part has E<0 sorted ascending; every removal batch is E-sorted
    and unbound; Mbound equals the mass sum of the bound particles; most-bound = first particle."""
    p = capi.make_params(box_size=62.5, softening=5e-3, periodic=True)
    e = capi.make_epoch(1.0)
    sizes = [3_000_000, 20000] + [64] * 200
    snap = synth.make_snapshot(sizes, seed=9, wrap=True)
    ctx = make_ctx(p)
    r = ctx.unbind_batch(e, snap)
    for s in (0, 1, 17):
        o = r.particles(s)
        b, en = snap.part_offset[s], snap.part_offset[s + 1]
        assert np.array_equal(np.sort(o), np.arange(b, en))
        nb = int(r.io["nbound"][s])
        E = r.energy[r.order_offset[s]: r.order_offset[s] + nb]
        assert nb >= 20 and (E < 0).all() and (np.diff(E) >= 0).all()
        m = snap.pos_mass[o[:nb], 3].astype(np.float64).sum()
        assert abs(r.io["mbound"][s] / m - 1) < 1e-6
        assert np.array_equal(r.io["mostbound_pos"][s].astype(np.float32), snap.pos_mass[o[0], :3])
        assert 0.5 < nb / (en - b) <= 1.0


@pytest.mark.parametrize("tpl", [2, 4])
def test_multi_target_walk_variants(tpl, tmp_path):
    """The 2- and 4-targets-per-lane walk kernels (used for >5e5-particle subhaloes) on a small case."""
    code = (
        "import sys, numpy as np\n"
        f"sys.path[:0] = [{os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r}, {os.path.dirname(os.path.abspath(__file__))!r}]\n"
        "import cases\nfrom hbtplus_b200.unbind import UnbindContext\nfrom oracle import pyoracle as po\n"
        "orc = po.load_oracle()\n"
        "for name in ('flat', 'nested'):\n"
        "    p, e, snap = cases.CASES[name]()\n"
        "    ctx = UnbindContext(p)\n"
        "    g = ctx.unbind_batch(e, snap)\n"
        "    w = po.run_batch(orc, 'hbto', p, e, snap)\n"
        "    assert np.array_equal(g.io['nbound'], w.io['nbound']), (g.io['nbound'], w.io['nbound'])\n"
        "    pm = snap.pos_mass[:snap.part_offset[1]]\n"
        "    a = ctx.tree_potential(e, pm, pm, self_mass=pm[:, 3].copy())\n"
        "    b = po.tree_potential(orc, 'hbto', p, e, pm, pm, self_mass=pm[:, 3].copy())\n"
        "    assert (np.abs(a - b) / np.abs(b)).max() < 1e-4\n"
        "print('OK')\n"
    )
    env = dict(os.environ, HBTU_WALK_TPL=str(tpl))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr


@pytest.mark.parametrize("route", ["masked128", "masked64", "small"])
def test_forced_walk_kernels_on_every_segment(route):
    """Every segment forced through ONE kernel family: the masked group walk with 128-target groups (HBTU_WALK_GROUP_MIN=1; the
    kernel of segments >= 256 targets), the same with 64-target groups (HBTU_WALK_MASKED_PAIRS=1), and the small-subhalo
    kernel (HBTU_WALK_SMALL_MAX huge: the dense sweep on every tree).  The device-counted accepted interactions must equal
    the oracle's, i.e. every target takes the reference's decisions under every routing."""
    code = (
        "import sys, numpy as np\n"
        f"sys.path[:0] = [{os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r}, {os.path.dirname(os.path.abspath(__file__))!r}]\n"
        "import cases\nfrom hbtplus_b200.unbind import UnbindContext\nfrom oracle import pyoracle as po\n"
        "orc = po.load_oracle()\n"
        "for name in ('flat', 'nested', 'periodic_straddle', 'massvar'):\n"
        "    p, e, snap = cases.CASES[name]()\n"
        "    ctx = UnbindContext(p)\n"
        "    g = ctx.unbind_batch(e, snap)\n"
        "    w = po.run_batch(orc, 'hbto', p, e, snap)\n"
        "    assert np.array_equal(g.io['nbound'], w.io['nbound']), (name, g.io['nbound'], w.io['nbound'])\n"
        "    s = int(np.argmax(np.diff(snap.part_offset)))\n"
        "    pm = np.ascontiguousarray(snap.pos_mass[snap.part_offset[s]:snap.part_offset[s + 1]])\n"
        "    ctx.set_counting(True)\n"
        "    a = ctx.tree_potential(e, pm, pm, self_mass=pm[:, 3].copy())\n"
        "    st = ctx.stats()\n"
        "    b = po.tree_potential(orc, 'hbto', p, e, pm, pm, self_mass=pm[:, 3].copy())\n"
        f"    assert (np.abs(a - b) / np.abs(b)).max() <= {POT_OBSERVED}, name\n"
        "    assert abs(st.pair_interactions - orc.hbto_last_interactions()) <= 1e-4 * st.pair_interactions + 2, name\n"
        "    assert st.walk_fallbacks == 0, name\n"
        "print('OK')\n"
    )
    extra = {"masked128": {"HBTU_WALK_GROUP_MIN": "1", "HBTU_WALK_MASKED_PAIRS": "2", "HBTU_WALK_SMALL_MAX": "0"},
             "masked64": {"HBTU_WALK_GROUP_MIN": "1", "HBTU_WALK_MASKED_PAIRS": "1", "HBTU_WALK_SMALL_MAX": "0"},
             "small": {"HBTU_WALK_SMALL_MAX": "1000000000"}}[route]
    env = dict(os.environ, **extra)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr


@pytest.mark.parametrize("M,refine", [(1000, True), (200, False)])
@pytest.mark.parametrize("periodic", [False, True])
def test_sampled_mode_vs_oracle(make_ctx, oracle_lib, M, refine, periodic):
    """MaxSampleSizeOfPotentialEstimate > 0: shuffle, sampled tree with scaled masses, the hole-based Hoare partition
    order that selects the next sample, and RefineBindingEnergyOrder - against the oracle run with the SAME
    counter-based permutation (oracle shuffle mode 1; its mode 0 is pinned bit-exactly to the reference's rand()
    stream by tests/test_oracle.py::test_oracle_matches_golden[sampled])."""
    p = capi.make_params(box_size=62.5, softening=5e-3, periodic=periodic, max_sample_size=M, refine_mostbound=refine, shuffle_seed=99)
    e = capi.make_epoch(1.0)
    sizes = [5000, 1200, 800, 20000, 300, 50, 2500, 70, 30]
    parent = [-1, 0, 0, -1, 3, 4, -1, 6, 6]
    snap = synth.make_snapshot(sizes, seed=3 + M, parent=parent, wrap=periodic, f_contam=0.35)
    ctx = make_ctx(p)
    got = ctx.unbind_batch(e, snap)
    oracle_lib.hbto_set_shuffle_mode(1)
    try:
        want = po.run_batch(oracle_lib, "hbto", p, e, snap)
    finally:
        oracle_lib.hbto_set_shuffle_mode(0)
    check_batch(snap, got, want, exact_counts=True)
    assert np.array_equal(got.io["iterations"], want.io["iterations"])
    for s in range(snap.nsub):
        nb = int(want.io["nbound"][s])
        a, b = got.particles(s), want.particles(s)
        # same order except where energies of the refined sample / sorted parts are within round-off of each other
        assert (a != b).sum() <= 0.01 * len(b) + 2, s
        assert np.array_equal(got.io["mostbound_pos"][s], want.io["mostbound_pos"][s]) or nb <= 1


def test_sampled_mode_is_statistically_the_references(make_ctx):
    """Against the reference's own sampled run (golden, libc rand() stream): a different sample of the same size, so
    agreement only within the sampling noise the reference documents ('percent level', configs/Example.conf:39)."""
    p, e, _ = cases.case_sampled()
    snap, z = load_golden("sampled")
    got = make_ctx(p).unbind_batch(e, snap)
    want = z["full_io"]
    live = want["nbound"] > 1
    assert np.array_equal(got.io["nbound"] > 1, live)
    assert np.all(np.abs(got.io["mbound"][live] / want["mbound"][live] - 1) < 0.03)
    small = np.diff(snap.part_offset) <= 1000  # sources below the sample size (and no larger descendant feeding them) are exact
    for s in np.nonzero(small & live)[0]:
        if snap.nest_offset[s + 1] == snap.nest_offset[s]:
            assert got.io["nbound"][s] == want["nbound"][s]
