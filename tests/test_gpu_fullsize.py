"""BASELINE.json's full single-GPU size (configs[1]: AqA2-shaped, 1.8e8 particles, 4e4 subhaloes) through the C-ABI, checked
with size-independent properties - the CPU oracle would need hours for the 1.3e8-particle central:

  * every subhalo's new particle list is a permutation of its source (own particles + what its descendants fed upwards)
    truncated to min(size, Nbound*SourceSubRelaxFactor): no index lost, none duplicated, none foreign;
  * binding energies of the bound part are ascending and all negative; Mbound is the mass of the bound part;
  * Nbound >= MinNumPartOfSub or the subhalo is dead with exactly one tracer particle;
  * a second execute of the staged batch reproduces every record, particle order and energy bit for bit (the fp64 reductions
    and scans of the path use fixed summation trees: seg_reduce.cuh, det_scan.cuh);
  * one checksum of checksums: the mass-weighted mean of the per-subhalo average positions of the top-level hierarchies
    stays inside the central halo (a wrong frame anywhere moves it).
The same batch then goes through hbtu_mask_batch (idempotence) and hbtu_profile_batch (monotone radii, M200 <= Mbound)."""
import numpy as np
import pytest

from hbtplus_b200 import capi, synth
from test_gpu_parity import make_ctx  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu


def test_full_size_properties(make_ctx):
    import torch

    import bench
    import cases_bench

    wl = bench.WORKLOADS["cfg2"]
    _, parent = wl.sizes(1.8e8)
    snap = wl.make(1.8e8, torch.device("cuda", 0), 0)
    torch.cuda.empty_cache()
    p = wl.params(0)
    e = capi.make_epoch(1.0)
    ctx = make_ctx(p)
    ctx.stage(e, snap, capi.HBTU_FLAG_TRUNCATE_SOURCE)
    ctx.execute()
    r = ctx.fetch(want_energy=True)
    io = r.io
    nsub = snap.nsub
    n_own = np.diff(snap.part_offset)
    nb, ns = io["nbound"], io["nsource"]
    assert nsub == 40001 and snap.npart > 1.79e8
    # survival rule
    dead = io["snapshot_index_of_death"] >= 0
    assert np.all((nb >= p.min_num_part_of_sub) | (nb <= 1))
    assert np.all(nb[dead] <= 1) and np.all(nb[~dead & (n_own >= 2)] >= p.min_num_part_of_sub)
    # truncation rule (src/subhalo_unbind.cpp:449-458)
    want_ns = np.where(nb <= 1, nb, np.minimum(io["nsource_full"], (nb * np.float32(p.source_sub_relax_factor)).astype(np.int64)))
    assert np.array_equal(ns, want_ns)
    assert np.array_equal(r.order_offset[1:] - r.order_offset[:-1], ns)
    # permutation property on the whole batch: within a subhalo no duplicates; every index belongs to the subhalo or a descendant
    ntot = int(r.order_offset[-1])
    order = r.order[:ntot]
    assert order.min() >= 0 and order.max() < snap.npart
    owner = np.repeat(np.arange(nsub), n_own)[order]  # subhalo that owned each listed particle on input
    sub_of_entry = np.repeat(np.arange(nsub), ns)
    par = np.asarray(parent)
    anc = owner.copy()
    ok = anc == sub_of_entry
    for _ in range(6):  # nest depth <= 4
        anc = np.where(ok | (anc < 0), anc, par[np.maximum(anc, 0)])
        ok |= anc == sub_of_entry
    assert ok.all()
    key = sub_of_entry.astype(np.int64) * (snap.npart + 1) + order
    assert len(np.unique(key)) == len(key)
    # energies: ascending and negative over the bound part; Mbound = mass of the bound part
    bound_entry = (np.arange(ntot) - np.repeat(r.order_offset[:-1], ns)) < np.repeat(nb, ns)
    en = r.energy[:ntot]
    live = np.repeat(nb > 1, ns)
    assert np.all(en[bound_entry & live] < 0)
    d = np.diff(en)
    same_sub = (sub_of_entry[1:] == sub_of_entry[:-1]) & bound_entry[1:] & bound_entry[:-1] & live[1:]
    assert np.all(d[same_sub] >= 0)
    mb = np.bincount(sub_of_entry[bound_entry], weights=snap.pos_mass[order[bound_entry], 3].astype(np.float64), minlength=nsub)
    big = nb > 1
    assert np.allclose(io["mbound"][big], mb[big], rtol=2e-6)
    # checksum of checksums
    top = (par < 0) & big
    com = (io["avg_pos"][top] * io["mbound"][top, None]).sum(0) / io["mbound"][top].sum()
    assert np.all(np.abs(com - wl.box / 2) < 0.05)
    # repeatability
    ctx.execute()
    r2 = ctx.fetch(want_energy=True)
    for f in io.dtype.names:  # bit for bit: every reduction and scan of the path has a fixed summation tree
        assert np.array_equal(r2.io[f], io[f]), f
    assert np.array_equal(r2.order[:ntot], order)
    assert np.array_equal(r2.energy[:ntot], en)
    st = ctx.stats()
    assert st.rounds < 60 and st.kernel_launches < 10000

    # --- next rows on the same batch -------------------------------------------------------------------------------------
    po, ids, no, nl, nbm = cases_bench.mask_inputs(snap)
    cnt, keep = ctx.mask_batch(po, ids, no, nl, nbm)
    assert 0.8 * po[-1] < cnt.sum() < po[-1]
    kept = np.concatenate([keep[po[s]:po[s] + cnt[s]] for s in np.nonzero(cnt)[0]])
    po2 = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    cnt2, _ = ctx.mask_batch(po2, ids[kept], no, nl, nbm)
    assert np.array_equal(cnt2, cnt)  # idempotent
    ppo, ppm, pio = cases_bench.profile_inputs(snap, r)
    prof = ctx.profile_batch(e, ppo, ppm, pio)
    assert np.all(prof["rhalf_comoving"][big] <= prof["r2sigma_comoving"][big])
    assert np.all(prof["bound_m200crit"][big] <= io["mbound"][big] * (1 + 1e-6))
    assert np.all(prof["vmax_physical"][big] > 0) and np.all(prof["rmax_comoving"][big] >= np.float32(p.softening_halo))
    assert np.all(prof["vmax_physical"][~big] == 0)
