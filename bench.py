#!/usr/bin/env python
"""bench.py - bound-particle unbinding throughput (particles/s) of the B200-native path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--particles P]

Workload (BASELINE.json configs[1]): AqA2-shaped single Milky-Way halo, synthetic, one central source of 0.72*P
particles + 4e4 subhaloes (dN/dn ~ n^-1.9 on [20, 5e6], nesting depth <= 4) totalling P = 1.8e8 particles per GPU,
BoxSize 100, softening 4.8e-5, periodic off, exact potential (MaxSampleSizeOfPotentialEstimate 0), theta 0.45.
A "step" is one RefineParticles-equivalent pass (all nesting levels, all iterations, TruncateSource) over the batch.

  value : whole-job particles/s with the batch resident in HBM (hbtu_execute only), CUDA-event timed on the
          library's stream, max over ranks.
  e2e   : the same through the C-ABI call a host shim makes (hbtu_unbind_batch) from pinned HOST buffers:
          H2D of positions/velocities and D2H of the new particle orders + records inside the timed region.
  N > 1 : one process per GPU (torchrun), each rank owns one such halo (weak scaling; hierarchies never span
          ranks); the only collective is the NCCL all-gather of the per-subhalo result records.

--impl reference times the reference's own CPU implementation (oracle/_ref: the unmodified HBT+ sources, else the
oracle port) on a bounded sample of the same workload with all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")  # before CUDA initialises: the library's upload stream gets its own hardware queue

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from hbtplus_b200 import capi, synth  # noqa: E402

METRIC = "bound-particle unbinding throughput"
UNIT = "particles/s"
FLOP_PER_INTERACTION = 12  # 3 FADD + FMUL + 2 FFMA + FSETP + FFMA (+MUFU) = 8 issue slots = 12 flop


class Workload:
    """One synthetic configuration of BASELINE.json / SURVEY.md 8(d): parameters, the size list + nest forest, the generator call."""

    def __init__(self, name, seed, box, eps, periodic, mass, full_particles, desc):
        self.name, self.seed, self.box, self.eps, self.periodic, self.mass, self.full_particles, self.desc = name, seed, box, eps, periodic, mass, full_particles, desc

    max_sample = 0  # MaxSampleSizeOfPotentialEstimate: 0 = exact potential (the parity and headline mode); --max-sample sets it

    def params(self, device: int = 0) -> capi.Params:
        return capi.make_params(box_size=self.box, softening=self.eps, periodic=self.periodic, max_sample_size=Workload.max_sample, device=device,
                                shuffle_seed=20240001)

    def sizes(self, particles: float):
        raise NotImplementedError

    def centres(self, sizes, parent):
        return [self.box / 2] * 3

    def describe(self, particles: float) -> str:
        return self.desc.format(P=particles, box=self.box, eps=self.eps)

    def make(self, particles: float, dev, rank: int, world: int = 1):
        sizes, parent = self.sizes(particles)
        return synth.make_snapshot_torch(sizes, device=dev, seed=self.seed + rank, box_size=self.box, particle_mass=self.mass, parent=parent,
                                         centre=self.centres(sizes, parent), wrap=self.periodic, pin=True)

    def cpu_sample(self, particles: float, target: int):
        """Bounded sample of the workload for the CPU legs: whole hierarchies in a shuffled order until ~`target` particles are
        collected; a hierarchy larger than target/10 is replaced by the hierarchies of its nested subhaloes (so the dominant
        sources - the AqA2 central, the two cfg-5 haloes - are excluded: each alone would take the CPU hours)."""
        sizes, parent = self.sizes(particles)
        nsub = len(sizes)
        children = [[] for _ in range(nsub)]
        for s in range(nsub):
            if parent[s] >= 0:
                children[parent[s]].append(s)
        tot = sizes.astype(np.int64).copy()
        for s in np.argsort(sizes, kind="stable"):  # children are smaller than their parents: they come first
            if parent[s] >= 0:
                tot[parent[s]] += tot[s]
        tops, todo = [], [s for s in range(nsub) if parent[s] < 0]
        while todo:
            s = todo.pop()
            if tot[s] > target // 10:
                todo.extend(children[s])
            else:
                tops.append(s)
        tops = np.array(sorted(tops))
        np.random.default_rng(self.seed + 1).shuffle(tops)
        chosen, n = [], 0
        for t in tops:
            stack = [int(t)]
            while stack:
                q = stack.pop()
                chosen.append(q)
                stack.extend(children[q])
            n += int(tot[t])
            if n >= target:
                break
        chosen = np.array(sorted(chosen))
        local = {int(g): i for i, g in enumerate(chosen)}
        par = np.array([local.get(int(parent[g]), -1) for g in chosen])
        snap = synth.make_snapshot(sizes[chosen], seed=self.seed + 2, box_size=self.box, particle_mass=self.mass, parent=par, wrap=self.periodic,
                                   centre=None if self.periodic else [self.box / 2] * 3, f_contam=0.2)
        desc = (f"{len(chosen)} subhaloes in {int((par < 0).sum())} whole hierarchies (sizes {int(sizes[chosen].min())}..{int(sizes[chosen].max())}, "
                f"{snap.npart} particles), same generator/seed family as the GPU batch; hierarchies above {target // 10} particles are "
                f"replaced by their nested subhaloes (the dominant sources are excluded)")
        return snap, desc


class AqA2(Workload):
    def sizes(self, particles):
        rng = np.random.default_rng(self.seed)
        nsub = max(50, int(40000 * min(1.0, particles / 1.8e8)))
        n_max = 5e6 * min(1.0, particles / 1.8e8)
        sizes = synth.aqa2_sizes(rng, n_total=particles, central_frac=0.72, nsub=nsub, n_max=max(n_max, 2000))
        parent = synth.nest_forest(rng, sizes, max_depth=4, p_nest=0.5, root=0)
        return sizes, parent


class MilliMill(Workload):
    def sizes(self, particles):
        rng = np.random.default_rng(self.seed)
        f = min(1.0, particles / self.full_particles)
        nsub = max(100, int(25000 * f))
        sizes = synth.subhalo_sizes(rng, nsub, 20, int(max(5e5 * f, 2000)))
        for _ in range(4):
            sizes = np.clip((sizes * (particles / sizes.sum())).astype(np.int64), 20, int(max(5e5 * f, 2000)))
        parent = synth.nest_forest(rng, sizes, max_depth=3, p_nest=0.2, root=None)  # ~2e4 FoF groups (roots) + ~5e3 satellites
        return synth.dfs_layout(sizes, parent)  # hierarchy by hierarchy, as RefineParticles visits them (subhalo_unbind.cpp:479-493)

    def centres(self, sizes, parent):
        return None  # groups uniformly in the periodic box


class DynamicMerger(Workload):
    def sizes(self, particles):
        rng = np.random.default_rng(self.seed)
        f = min(1.0, particles / self.full_particles)
        nsub = max(100, int(1e5 * f))
        big = int(5e6 * f)
        small = rng.integers(20, 201, nsub).astype(np.int64)
        sizes = np.concatenate([[big, big], small]).astype(np.int64)
        parent = np.concatenate([[-1, -1], rng.integers(0, 2, nsub)]).astype(np.int64)  # every tiny subhalo sits in one of the two haloes
        return sizes, parent

    def centres(self, sizes, parent):
        c = np.tile(np.array([self.box / 2] * 3), (len(sizes), 1))
        c[0, 0] -= 0.5  # two haloes at 1 Mpc/h separation
        c[1, 0] += 0.5
        return c


class Eagle(Workload):
    """SURVEY 8(d) cfg 4.  ONE snapshot of particles x world grouped particles (full size: 1.36e9 = 40 % of 1504^3 on 8 GPUs,
    1.7e8 per GPU), never materialised on one host: every rank draws the same global size list and nest forest from the fixed
    seed, the hierarchies (FoF groups with their satellites) are dealt to the ranks by the cost-weighted longest-processing-
    time-first queue (sched.lpt_partition, cost = sum n log2 n), and a rank generates only the particles of its own shard."""

    def sizes(self, particles):
        rng = np.random.default_rng(self.seed)
        f = min(1.0, particles / (8 * self.full_particles))  # `particles` = the whole snapshot; full size 8 x 1.7e8 = 1.36e9
        ngroups, nsat = max(50, int(3e6 * f)), max(20, int(1e6 * f))
        n_max = int(max(3e7 * f, 5000))
        groups = synth.subhalo_sizes(rng, ngroups, 20, n_max)
        sats = synth.subhalo_sizes(rng, nsat, 20, max(n_max // 8, 400))
        for _ in range(4):  # rescale to the wanted particle total (satellites hold ~1/8 of it)
            groups = np.clip((groups * (0.875 * particles / groups.sum())).astype(np.int64), 20, n_max)
            sats = np.clip((sats * (0.125 * particles / sats.sum())).astype(np.int64), 20, max(n_max // 8, 400))
        # every satellite hangs off a group (90 %) or off another satellite (10 %, depth 2) at least 4x its own size
        gsort = np.argsort(groups, kind="stable")
        elig = len(groups) - np.searchsorted(groups[gsort], 4 * sats, side="left")
        pick = (rng.random(nsat) * np.maximum(elig, 1)).astype(np.int64)
        par = gsort[len(groups) - 1 - np.minimum(pick, len(groups) - 1)]
        ssort = np.argsort(sats, kind="stable")
        elig2 = nsat - np.searchsorted(sats[ssort], 4 * sats, side="left")
        deep = (rng.random(nsat) < 0.1) & (elig2 > 0)
        pick2 = (rng.random(nsat) * np.maximum(elig2, 1)).astype(np.int64)
        par2 = ssort[nsat - 1 - np.minimum(pick2, nsat - 1)]
        deep &= ~deep[par2]  # a depth-2 satellite hangs off a depth-1 one
        parent = np.concatenate([np.full(ngroups, -1, np.int64), np.where(deep, ngroups + par2, par)])
        return np.concatenate([groups, sats]).astype(np.int64), parent

    def centres(self, sizes, parent):
        return None  # groups uniformly in the periodic box

    def shard(self, particles, rank, world):
        """(sizes, parent, global index) of the subhaloes rank `rank` of `world` owns."""
        from hbtplus_b200 import sched

        sizes, parent = self.sizes(particles * world)
        if world == 1:
            return synth.dfs_layout(sizes, parent, return_order=True)
        root = sched.roots_of(parent)
        cap = sizes.astype(np.float64)
        cost_sub = cap * np.log2(np.maximum(cap, 2.0))
        roots = np.nonzero(parent < 0)[0]
        cost = np.zeros(len(sizes))
        np.add.at(cost, root, cost_sub)
        owner = np.full(len(sizes), -1, np.int64)
        owner[roots] = sched.lpt_partition(cost[roots], world)
        mine = np.nonzero(owner[root] == rank)[0]
        local = np.full(len(sizes), -1, np.int64)
        local[mine] = np.arange(len(mine))
        par = np.where(parent[mine] >= 0, local[np.maximum(parent[mine], 0)], -1)
        # the shard hierarchy by hierarchy, parents first: the order in which RefineParticles visits subhaloes (subhalo_unbind.cpp:479-493)
        # and in which the drop-in shim lays a batch out; it lets hbtu_unbind_batch pipeline the upload of a box of many hierarchies
        s2, p2, order = synth.dfs_layout(sizes[mine], par, return_order=True)
        return s2, p2, mine[order]

    def make(self, particles, dev, rank, world: int = 1):
        sizes, parent, _ = self.shard(particles, rank, world)
        return synth.make_snapshot_torch(sizes, device=dev, seed=self.seed + 17 * rank, box_size=self.box, particle_mass=self.mass, parent=parent,
                                         centre=None, wrap=True, pin=True)



WORKLOADS = {
    "cfg2": AqA2("cfg2", 20240002, 100.0, 4.8e-5, False, 1e-6, 1.8e8,
                 "BASELINE configs[1] / SURVEY 8(d) cfg 2: AqA2-shaped synthetic Milky-Way halo per GPU: central source 0.72*P + subhaloes dN/dn~n^-1.9 on "
                 "[20,5e6], nest depth<=4, P={P:.3g} particles, BoxSize {box}, eps {eps}, periodic off, exact potential (MaxSample 0), theta 0.45"),
    "cfg3": MilliMill("cfg3", 20240003, 62.5, 5e-3, True, 0.086, 8.9e6,
                      "BASELINE configs[2] / SURVEY 8(d) cfg 3: MilliMill-shaped 270^3 box per GPU: P={P:.3g} grouped particles (45 % of 1.97e7) in ~2e4 FoF groups + ~5e3 "
                      "satellites, dN/dn~n^-1.9 on [20,5e5], nest depth<=3, BoxSize {box}, eps {eps}, periodic on, exact potential, theta 0.45"),
    "cfg4": Eagle("cfg4", 20240004, 67.77, 1.80239e-3, True, 6.57e-4, 1.7e8,
                  "BASELINE configs[3] / SURVEY 8(d) cfg 4: EagleL100N1504-shaped DM-only snapshot, P={P:.3g} grouped particles PER GPU (full size 1.7e8 x 8 GPUs = "
                  "1.36e9 = 40 % of 1504^3) in FoF groups dN/dn~n^-1.9 on [20,3e7] + satellites (depth<=2), hierarchies dealt to the ranks by the cost-weighted "
                  "LPT queue and generated shard-locally, BoxSize {box}, eps {eps}, periodic on, exact potential, theta 0.45"),
    "cfg5": DynamicMerger("cfg5", 20240005, 250.0, 2.1e-3, False, 0.086, 1.0e7 + 1.1e7,
                          "BASELINE configs[4] / SURVEY 8(d) cfg 5: DynamicMerger-shaped: two 5e6-particle haloes at 1 Mpc/h separation + 1e5 tiny subhaloes "
                          "n in [20,200] nested in them (small-subhalo batched path), P={P:.3g} particles, BoxSize {box}, eps {eps}, periodic off, exact potential, theta 0.45"),
}
SEED = WORKLOADS["cfg2"].seed


def params_for(device: int = 0, workload: str = "cfg2") -> capi.Params:
    return WORKLOADS[workload].params(device)


def workload_sizes(particles: float, seed: int = SEED):  # kept for tools/ and tests that size the cfg-2 batch
    return WORKLOADS["cfg2"].sizes(particles)


def cpu_sample(particles: float, seed: int, target: int, workload: str = "cfg2"):
    return WORKLOADS[workload].cpu_sample(particles, target)


def run_cpu(snap, threads: int | None = None, workload: str = "cfg2"):
    """One RefineParticles-equivalent pass on the host cores; returns (seconds, kind, threads, sum Nbound, result)."""
    from oracle import pyoracle as po

    p = params_for(0, workload)
    e = capi.make_epoch(1.0)
    ncpu = threads or os.cpu_count() or 1
    if po.have_ref() and p.max_sample_size == 0:
        lib, prefix, kind = po.load_ref(), "hbtref", "reference"
    else:  # sampled mode: the reference's random_shuffle on libc rand() cannot be shared with a GPU; the port takes the library's permutation
        if not os.path.exists(po.ORACLE_PATH):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"])
        lib, prefix, kind = po.load_oracle(), "hbto", "port"
        lib.hbto_set_shuffle_mode(1 if p.max_sample_size > 0 else 0)
    getattr(lib, prefix + "_set_num_threads")(ncpu)
    t0 = time.perf_counter()
    r = po.run_batch(lib, prefix, p, e, snap, flags=capi.HBTU_FLAG_TRUNCATE_SOURCE, want_energy=False)
    dt = time.perf_counter() - t0
    return dt, kind, ncpu, int(r.io["nbound"].sum()), r


def parity_block(ctx, e, csnap, cres, kind, workload: str = "cfg2", pot_targets: int = 100_000):
    """The GPU path on exactly the snapshot the CPU leg just unbound (same generator and parameters - cfg 2: BoxSize 100, eps 4.8e-5,
    m_p 1e-6 - exact potential), compared record by record with the reference's result: what the timed kernels compute is what the reference
    computes.  Gates of BASELINE.json's north_star are reported, not assumed: Nbound / survival / death flags, bound mass,
    bound-membership Jaccard, and the per-particle potential of the sample's largest subhalo against the reference's own
    GravityTree_t::EvaluatePotential (src/gravity_tree.cpp:79-164) on `pot_targets` of its particles."""
    from oracle import pyoracle as po

    flags = capi.HBTU_FLAG_TRUNCATE_SOURCE
    got = ctx.unbind_batch(e, csnap, flags=flags, want_energy=False)
    nb_g, nb_w = got.io["nbound"].astype(np.int64), cres.io["nbound"].astype(np.int64)
    live = nb_w > 1
    mb_g, mb_w = got.io["mbound"].astype(np.float64), cres.io["mbound"].astype(np.float64)
    dm = np.abs(mb_g[live] - mb_w[live]) / np.abs(mb_w[live])
    jac = np.ones(csnap.nsub)
    for s in np.nonzero(live | (nb_g > 1))[0]:
        a, b = got.bound(s), cres.bound(s)
        inter = len(np.intersect1d(a, b, assume_unique=True))
        union = len(a) + len(b) - inter
        jac[s] = inter / union if union else 1.0
    miss = jac < 0.999
    out = {
        "against": kind, "sample": f"the cpu_baseline sample: {csnap.nsub} subhaloes, {csnap.npart} particles", "subhaloes": int(csnap.nsub),
        "frac_identical_nbound": float(np.mean(nb_g == nb_w)), "max_abs_dnbound": int(np.abs(nb_g - nb_w).max()),
        "survival_identical": bool(np.array_equal(nb_g > 1, nb_w > 1)),
        "death_flags_identical": bool(np.array_equal(got.io["snapshot_index_of_death"], cres.io["snapshot_index_of_death"])),
        "sink_flags_identical": bool(np.array_equal(got.io["sink_track_id"], cres.io["sink_track_id"])),
        "nsource_identical_frac": float(np.mean(got.io["nsource"] == cres.io["nsource"])),
        "max_rel_dmbound": float(dm.max()) if dm.size else 0.0, "frac_mbound_within_1e-3": float(np.mean(dm <= 1e-3)) if dm.size else 1.0,
        "min_jaccard": float(jac.min()), "jaccard_miss_rate": float(np.mean(miss)), "jaccard_misses": int(miss.sum()),
        "largest_subhalo_with_jaccard_miss": int(nb_w[miss].max()) if miss.any() else 0,
        "gates": {"potential_rel_err": 1e-3, "mbound_rel": 1e-3, "jaccard": 0.999},
    }
    # per-particle potential, largest subhalo of the sample, default kernel routing
    s = int(np.argmax(np.diff(csnap.part_offset)))
    b, en = int(csnap.part_offset[s]), int(csnap.part_offset[s + 1])
    src = csnap.pos_mass[b:en]
    ctx.set_counting(True)
    pot = ctx.tree_potential(e, src, src, self_mass=src[:, 3].copy())
    st = ctx.stats()
    ctx.set_counting(False)
    pick = np.random.default_rng(7).choice(en - b, size=min(pot_targets, en - b), replace=False)
    lib, prefix = (po.load_ref(), "hbtref") if kind == "reference" else (po.load_oracle(), "hbto")
    want = po.tree_potential(lib, prefix, params_for(0, workload), e, src, src[pick], self_mass=src[pick, 3].copy())
    rel = np.abs(pot[pick] - want) / np.abs(want)
    out["potential"] = {"sources": en - b, "targets_compared": int(len(pick)), "max_rel_err": float(rel.max()), "mean_rel_err": float(rel.mean()),
                        "frac_above_1e-3": float(np.mean(rel > 1e-3)), "interactions_per_target": st.pair_interactions / (en - b),
                        "walk_fallbacks": int(st.walk_fallbacks)}
    out["ok"] = bool(out["survival_identical"] and out["death_flags_identical"] and out["potential"]["max_rel_err"] <= 1e-3
                     and out["frac_mbound_within_1e-3"] >= 0.998 and out["jaccard_miss_rate"] <= 0.002)
    return out


def phase_rooflines(st0, build_ms, other_ms, peaks):
    """HBM rooflines of the two non-walk phases of a step with SURVEY.md 8(d)'s ALGORITHMIC bytes (not this implementation's
    pass list): per source particle and tree build 16 B position read + 12 B key/index write, one sort pass 12 B read + 12 B
    write, 16 B sorted read + 8 B moment traffic + 24 B node write = 100 B; per walk target and round 32 B energy inputs,
    4 B E write, partition 8 B read + 8 B write, one E-sort pass 12 B + 12 B = 76 B.  x particles summed over the step's
    rounds / CUDA-event time of the phase."""
    hbm = peaks.get("hbm_gbs", 6650.0)
    build_bytes, other_bytes = 100, 76
    out = {}
    for name, n, ms, b in (("tree_build", st0.tree_sources, build_ms, build_bytes), ("partition_sort_reduce", st0.walk_targets, other_ms, other_bytes)):
        ach = b * n / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        out[name] = {"bound": "hbm", "particles_over_rounds": int(n), "bytes_per_particle": b, "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm}
    return out


def bench_profile_row(ctx, e, snap, res, csnap, peaks):
    """SURVEY.md 8(f) next-2, measured beside the headline: Subhalo_t::CalculateProfileProperties + CalculateShape
    (src/subhalo.cpp:242-398) of every subhalo of the step through hbtu_profile_batch, HBM roofline, and the reference's
    own functions (oracle/_ref) on the bound lists of the CPU sample."""
    from oracle import pyoracle as po
    import cases_bench

    part_offset, pm, io = cases_bench.profile_inputs(snap, res)
    nb = int(np.where(io["nbound"] > 1, io["nbound"], 0).sum())
    t_e2e, t_k = [], []
    for _ in range(3):
        t0 = time.perf_counter()
        ctx.profile_batch(e, part_offset, pm, io)
        t_e2e.append(time.perf_counter() - t0)
        t_k.append(ctx.stats().execute_ms * 1e-3)
    # algorithmic bytes per bound particle: radius/shape 16 B read + 12 B key/index write; sort 12 + 12;
    # cumulative mass 12 + 4 (mass gather) read + 8 write; select 8 + 8 read
    bytes_per = 16 + 12 + 24 + 24 + 16
    # fused with the resident unbinding batch (hbtu_profile_executed): only the per-subhalo records cross PCIe
    ctx.stage(e, snap, capi.HBTU_FLAG_TRUNCATE_SOURCE)
    ctx.execute()
    t_fused = []
    for _ in range(3):
        t0 = time.perf_counter()
        ctx.profile_executed(io)
        t_fused.append(time.perf_counter() - t0)
    row = {"bound_particles": nb, "subhaloes": int(snap.nsub), "kernel_ms": float(np.min(t_k)) * 1e3, "e2e_ms": float(np.min(t_e2e)) * 1e3,
           "e2e_resident_ms": float(np.min(t_fused)) * 1e3,
           "value": nb / float(np.min(t_k)), "e2e_value": nb / float(np.min(t_e2e)), "unit": "bound particles/s",
           "roofline": {"bound": "hbm", "achieved": bytes_per * nb / float(np.min(t_k)) / 1e9, "peak": peaks.get("hbm_gbs", 6650.0), "unit": "GB/s",
                        "frac": bytes_per * nb / float(np.min(t_k)) / 1e9 / peaks.get("hbm_gbs", 6650.0), "bytes_per_particle": bytes_per}}
    if po.have_ref():
        ref = po.load_ref()
        ref.hbtref_set_num_threads(os.cpu_count() or 1)
        cres = po.run_batch(ref, "hbtref", params_for(), e, csnap, flags=capi.HBTU_FLAG_TRUNCATE_SOURCE, want_energy=False)
        cpo, cpm, cio = cases_bench.profile_inputs(csnap, cres)
        t0 = time.perf_counter()
        po.profile_batch(ref, "hbtref", params_for(), e, cpo, cpm, cio)
        dt = time.perf_counter() - t0
        cnb = int(np.where(cio["nbound"] > 1, cio["nbound"], 0).sum())
        row["cpu_baseline"] = {"value": cnb / dt, "unit": "bound particles/s", "cores": os.cpu_count() or 1, "kind": "reference",
                               "sample": f"bound lists of the unbinding CPU sample ({cnb} bound particles, {csnap.nsub} subhaloes)", "seconds": dt}
    return row


def bench_mask_row(ctx, snap, csnap, peaks):
    """SURVEY.md 8(f) next-1, measured beside the headline: SubhaloSnapshot_t::MaskSubhalos (src/subhalo_tracking.cpp:793-841)
    of the step's whole forest through hbtu_mask_batch, HBM roofline, and the reference's own member on the CPU sample."""
    from oracle import pyoracle as po
    import cases_bench

    args = cases_bench.mask_inputs(snap)
    n = int(args[0][-1])
    t_e2e, t_k, kept = [], [], 0
    for _ in range(3):
        t0 = time.perf_counter()
        new_count, _ = ctx.mask_batch(*args)
        t_e2e.append(time.perf_counter() - t0)
        t_k.append(ctx.stats().execute_ms * 1e-3)
        kept = int(new_count.sum())
    # algorithmic bytes per list entry: keys 8 R + 20 W, sort 12 R + 12 W (one algorithmic pass),
    # runs 12 + 4 R + 4 W, scan 4 R + 4 W, scatter 12 R + 4 W
    bytes_per = 28 + 24 + 20 + 8 + 16
    row = {"list_entries": n, "kept": kept, "subhaloes": int(snap.nsub), "kernel_ms": float(np.min(t_k)) * 1e3, "e2e_ms": float(np.min(t_e2e)) * 1e3,
           "value": n / float(np.min(t_k)), "e2e_value": n / float(np.min(t_e2e)), "unit": "list entries/s",
           "roofline": {"bound": "hbm", "achieved": bytes_per * n / float(np.min(t_k)) / 1e9, "peak": peaks.get("hbm_gbs", 6650.0), "unit": "GB/s",
                        "frac": bytes_per * n / float(np.min(t_k)) / 1e9 / peaks.get("hbm_gbs", 6650.0), "bytes_per_entry": bytes_per}}
    if po.have_ref():
        ref = po.load_ref()
        ref.hbtref_set_num_threads(os.cpu_count() or 1)
        cargs = cases_bench.mask_inputs(csnap)
        t0 = time.perf_counter()
        po.mask_batch(ref, "hbtref", params_for(), *cargs)
        dt = time.perf_counter() - t0
        row["cpu_baseline"] = {"value": int(cargs[0][-1]) / dt, "unit": "list entries/s", "cores": os.cpu_count() or 1, "kind": "reference",
                               "sample": f"the forest of the unbinding CPU sample ({int(cargs[0][-1])} list entries, {csnap.nsub} subhaloes); includes the harness' fill of the reference's Subhalo_t lists",
                               "seconds": dt}
    return row


def bench_idtable_row(ctx, n, n_cpu, peaks):
    """SURVEY.md 8(f) next-4, measured beside the headline: MappedIndexTable_t::Fill + GetIndices (src/hash.tpp:18-32,
    src/hash_remote.tpp:9-88) for a snapshot of n particle Ids and n queries (half present, half absent)."""
    from oracle import pyoracle as po

    def make(m, seed):
        rng = np.random.default_rng(seed)
        ids = rng.permutation(m).astype(np.int64) * 7 + 3
        q = np.concatenate([ids[rng.integers(0, m, m // 2)], rng.integers(0, 7 * m, m - m // 2, dtype=np.int64) * 7 + 4])
        return ids, q

    ids, q = make(n, 11)
    tb, tq, eb, eq = [], [], [], []
    for _ in range(2):
        t0 = time.perf_counter(); ctx.idtable_build(ids); eb.append(time.perf_counter() - t0); tb.append(ctx.stats().execute_ms * 1e-3)
        t0 = time.perf_counter(); out = ctx.idtable_query(q); eq.append(time.perf_counter() - t0); tq.append(ctx.stats().execute_ms * 1e-3)
    assert int((out >= 0).sum()) == n // 2
    # algorithmic bytes (kept from round 1 so that the fractions compare): build 8 R + 12 W (keys) + 12 R + 12 W (one pass over the
    # pairs); query 8 R + 8 W + 8 probe.  The hash table of round 2 moves one random 32-byte sector per entry and ~1.5 per query.
    bbytes, qbytes = 44, 24
    k = float(np.min(tb)) + float(np.min(tq))
    row = {"table_entries": int(n), "queries": int(len(q)), "build_kernel_ms": float(np.min(tb)) * 1e3, "query_kernel_ms": float(np.min(tq)) * 1e3,
           "build_e2e_ms": float(np.min(eb)) * 1e3, "query_e2e_ms": float(np.min(eq)) * 1e3, "value": len(q) / k, "unit": "queries/s (build + query kernels)",
           "roofline": {"bound": "hbm", "achieved": (bbytes * n + qbytes * len(q)) / k / 1e9, "peak": peaks.get("hbm_gbs", 6650.0), "unit": "GB/s",
                        "frac": (bbytes * n + qbytes * len(q)) / k / 1e9 / peaks.get("hbm_gbs", 6650.0)}}
    if po.have_ref():
        ref = po.load_ref()
        ref.hbtref_set_num_threads(os.cpu_count() or 1)
        cids, cq = make(n_cpu, 12)
        t0 = time.perf_counter()
        po.idtable_query(ref, "hbtref", params_for(), cids, cq)
        dt = time.perf_counter() - t0
        row["cpu_baseline"] = {"value": len(cq) / dt, "unit": "queries/s (Fill + sort queries + GetIndices)", "cores": os.cpu_count() or 1, "kind": "reference",
                               "sample": f"{n_cpu} table entries, {len(cq)} queries", "seconds": dt}
    return row


def bench_dropin_row(wl, particles, dev, csnap, cpu_threads):
    """The metric as SURVEY.md 8(d) defines it: wall time of SubhaloSnapshot_t::RefineParticles() itself on an in-memory
    SubhaloSnapshot_t (vector<Particle_t> per subhalo) through the drop-in seam - libhbtdropin_v32.so = the reference's own
    objects with subhalo_unbind.o replaced by integration/subhalo_unbind_b200.o.  Inside the timed call: AoS -> pinned SoA pack
    (OpenMP), H2D, all kernels, D2H, the permutation of every vector<Particle_t>.  Beside it: the reference's own
    RefineParticles (libhbtref_v32.so) on the CPU sample, timed the same way."""
    from oracle import pyoracle as po

    if not po.have_dropin("v32"):
        return {"unavailable": "oracle/_ref/libhbtdropin_v32.so not built"}
    drop = po.load_dropin("v32")
    drop.hbtref_set_num_threads(cpu_threads)
    snap = wl.make(particles, dev, 7)
    nsub = snap.nsub
    roots = np.ones(nsub, bool)
    if snap.nest_list is not None:
        roots[snap.nest_list] = False
    # every root is the central of its own host halo; nested subhaloes are its old nests (src/subhalo_unbind.cpp:479-493)
    root_of = np.arange(nsub)
    if snap.nest_offset is not None:
        parent = np.full(nsub, -1, np.int64)
        for s in range(nsub):
            parent[snap.nest_list[snap.nest_offset[s]:snap.nest_offset[s + 1]]] = s
        order = np.argsort(-np.diff(snap.part_offset), kind="stable")
        for s in order:  # parents are larger: they come first
            if parent[s] >= 0:
                root_of[s] = root_of[parent[s]]
    halo_index = {int(r): i for i, r in enumerate(np.nonzero(roots)[0])}
    host = np.array([halo_index[int(r)] for r in root_of], np.int32)
    mb = np.diff(snap.part_offset).astype(np.float32)
    e = capi.make_epoch(1.0)
    secs = []
    for _ in range(3):  # the first call pays the page-locking of the pinned staging buffers
        r = po.refine_particles(drop, wl.params(0), e, snap, host, nsub, len(halo_index), mb)
        secs.append(r.refine_seconds)
    row = {"api": "SubhaloSnapshot_t::RefineParticles() via libhbtdropin_v32.so", "particles": int(snap.npart), "subhaloes": int(nsub),
           "seconds_first_call": secs[0], "seconds": float(np.min(secs[1:])), "value": snap.npart / float(np.min(secs[1:])), "unit": UNIT,
           "host_threads": cpu_threads, "sum_nbound": int(r.io["nbound"].sum())}
    try:  # where the last call's wall time went inside the shim (integration/subhalo_unbind_b200.cpp: HBT_B200_LastBatchTimes)
        import ctypes as C
        t4 = (C.c_double * 4)()
        drop.HBT_B200_LastBatchTimes(t4)
        row["last_call_ms"] = {"pack_aos_to_pinned_soa": 1e3 * t4[0], "hbtu_unbind_batch": 1e3 * t4[1], "permute_particle_vectors": 1e3 * t4[2],
                               "refine_particles_total": 1e3 * secs[-1]}
    except AttributeError:
        pass
    if po.have_ref():
        ref = po.load_ref()
        ref.hbtref_set_num_threads(cpu_threads)
        croots = np.ones(csnap.nsub, bool)
        if csnap.nest_list is not None:
            croots[csnap.nest_list] = False
        cparent = np.full(csnap.nsub, -1, np.int64)
        if csnap.nest_offset is not None:
            for s in range(csnap.nsub):
                cparent[csnap.nest_list[csnap.nest_offset[s]:csnap.nest_offset[s + 1]]] = s
        croot_of = np.arange(csnap.nsub)
        for s in np.argsort(-np.diff(csnap.part_offset), kind="stable"):
            if cparent[s] >= 0:
                croot_of[s] = croot_of[cparent[s]]
        cidx = {int(r_): i for i, r_ in enumerate(np.nonzero(croots)[0])}
        chost = np.array([cidx[int(r_)] for r_ in croot_of], np.int32)
        cr = po.refine_particles(ref, wl.params(0), e, csnap, chost, csnap.nsub, len(cidx), np.diff(csnap.part_offset).astype(np.float32))
        row["reference"] = {"api": "the reference's own SubhaloSnapshot_t::RefineParticles() (libhbtref_v32.so)", "particles": int(csnap.npart),
                            "seconds": cr.refine_seconds, "value": csnap.npart / cr.refine_seconds, "unit": UNIT, "host_threads": cpu_threads}
    return row


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for k, nm in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)), "MEASURED_PEAKS.json"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS), help="cfg2 = BASELINE configs[1] (the headline); cfg3 / cfg5 = SURVEY 8(d) cfg 3 / 5")
    ap.add_argument("--max-sample", type=int, default=0, help="MaxSampleSizeOfPotentialEstimate (0 = exact potential, the headline; 1000 = the reference's default "
                    "sampled mode; the CPU legs then run the oracle port with the library's counter-based permutation)")
    ap.add_argument("--particles", type=float, default=None, help="particles per GPU (default: the configuration's full size; cfg2: 1.8e8)")
    ap.add_argument("--cpu-sample", type=int, default=3_000_000, help="particles in the bounded CPU sample (10-30 s on 16 cores)")
    ap.add_argument("--e2e-steps", type=int, default=None, help="end-to-end steps (default: --steps)")
    ap.add_argument("--dropin-particles", type=float, default=4e7, help="size of the in-memory SubhaloSnapshot_t of the drop-in e2e row (RefineParticles via libhbtdropin)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = one snapshot per GPU (whole hierarchies never span ranks); strong = ONE snapshot on all GPUs, every round's "
                         "walk targets dealt over the ranks (hbtu_set_walk_split), one NCCL all-reduce of 4 B per target per round")
    ap.add_argument("--profile", action="store_true", help="for ncu: no counting pass, no e2e, no CPU leg (numbers printed under a profiler are not bench values)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = WORKLOADS[args.workload]
    Workload.max_sample = args.max_sample
    if args.particles is None:
        args.particles = wl.full_particles
    if args.e2e_steps is None:
        args.e2e_steps = args.steps
    workload = wl.describe(args.particles)
    if args.max_sample > 0:
        workload = workload.replace("exact potential (MaxSample 0)", f"SAMPLED potential (MaxSampleSizeOfPotentialEstimate {args.max_sample})").replace(
            "exact potential", f"SAMPLED potential (MaxSampleSizeOfPotentialEstimate {args.max_sample})")

    if args.impl == "reference":
        if rank != 0:
            return 0
        snap, desc = wl.cpu_sample(args.particles, args.cpu_sample)
        times, kind, ncpu, nb = [], None, None, 0
        for i in range(args.warmup + args.steps):
            dt, kind, ncpu, nb, _ = run_cpu(snap, workload=wl.name)
            if i >= args.warmup:
                times.append(dt)
        dt = float(np.mean(times))
        val = snap.npart / dt
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "sample": desc, "cpu_threads": ncpu, "sum_nbound": nb},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": ncpu, "kind": kind, "sample": desc},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return 0

    import torch
    import torch.distributed as dist

    from hbtplus_b200.unbind import UnbindContext

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the unbinding path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    strong = args.scaling == "strong" and world > 1
    # weak: same shape on every rank, own particle realisation; strong: the SAME snapshot on every rank
    snap = wl.make(args.particles, dev, 0 if strong else rank, 1 if strong else world)
    torch.cuda.empty_cache()
    n_local = snap.npart
    ctx = UnbindContext(wl.params(local_rank))
    if strong:
        from hbtplus_b200 import sched

        ctx.set_walk_split(rank, world, sched.torch_allreduce(dev))
    e = capi.make_epoch(1.0)
    flags = capi.HBTU_FLAG_TRUNCATE_SOURCE

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # one counted pass: exact number of accepted pair interactions of a step (roofline numerator); also warm-up 0
    ctx.stage(e, snap, flags)
    if not args.profile:
        ctx.set_counting(True)
    ctx.execute()
    st0 = ctx.stats()
    ctx.set_counting(False)
    for _ in range(args.warmup):
        ctx.execute()
    barrier()
    exec_ms, walk_ms, build_ms, other_ms, wall = [], [], [], [], []
    with ClockSampler(local_rank) as clk:
        for _ in range(args.steps):
            t0 = time.perf_counter()
            ctx.execute()
            wall.append(time.perf_counter() - t0)
            st = ctx.stats()
            exec_ms.append(st.execute_ms)
            walk_ms.append(st.walk_ms)
            build_ms.append(st.build_ms)
            other_ms.append(st.other_ms)
            step_launches = st.kernel_launches  # kernels of ONE resident step (the e2e call below may run the batch in two parts)
        barrier()
        # end to end through the ABI call, pinned host buffers in, host arrays out
        e2e_wall, e2e_h2d, e2e_d2h, e2e_parts = [], 0, 0, None
        res = None
        cap = capi.order_capacity(snap.part_offset, snap.nest_offset, snap.nest_list)
        order_pinned = torch.empty(max(cap, 1), dtype=torch.int32, pin_memory=True).numpy()  # caller-owned output buffer, as a host shim would keep
        for i in range(0 if args.profile else args.e2e_steps + 1):
            barrier()
            t0 = time.perf_counter()
            res = ctx.unbind_batch(e, snap, flags=flags, want_energy=False, order_buf=order_pinned)
            dt = time.perf_counter() - t0
            if i > 0:
                e2e_wall.append(dt)
            st = ctx.stats()
            e2e_h2d, e2e_d2h = st.h2d_bytes, st.d2h_bytes
            e2e_parts = {"h2d_ms_async": st.h2d_ms, "execute_ms": st.execute_ms, "d2h_ms": st.d2h_ms, "wall_ms": dt * 1e3,
                         "stage_wall_ms": st.stage_wall_ms, "execute_wall_ms": st.execute_wall_ms, "fetch_wall_ms": st.fetch_wall_ms}
    clocks = clk.summary()
    launches = step_launches

    t_dev = torch.tensor([sum(exec_ms) * 1e-3, sum(e2e_wall), sum(walk_ms) * 1e-3], dtype=torch.float64, device=dev)
    if res is None:
        res = ctx.fetch(want_energy=False)
        e2e_wall = []
    once = 1.0 if (not strong or rank == 0) else 0.0  # strong scaling: one snapshot, every rank holds all of its records
    tot = torch.tensor([once * n_local, once * float(res.io["nbound"].sum()), float(st0.pair_interactions)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    if world > 1 and not strong:
        # the path's only collective: gather the per-subhalo result records (~150 B each) of all ranks
        from hbtplus_b200 import sched

        nsub_all = torch.tensor([snap.nsub], device=dev)
        counts = [torch.zeros_like(nsub_all) for _ in range(world)]
        dist.all_gather(counts, nsub_all)
        offs = np.concatenate([[0], np.cumsum([int(c) for c in counts])])
        table = sched.gather_records(res.io, offs[rank] + np.arange(snap.nsub), int(offs[-1]), device=dev)
        assert table is not None and len(table) == offs[-1]
    per_rank = [sum(exec_ms) / args.steps]
    if world > 1:
        allt = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(allt, torch.tensor([per_rank[0]], dtype=torch.float64, device=dev))
        per_rank = [float(x) for x in allt]
    t_exec, t_e2e, t_walk = (float(x) for x in t_dev.cpu())
    n_all, nb_all, inter_all = (float(x) for x in tot.cpu())
    value = n_all * args.steps / t_exec
    e2e_value = n_all * len(e2e_wall) / t_e2e if t_e2e > 0 else None

    out = None
    if rank == 0:
        peaks, peak_src = measured_peaks()
        nsm = torch.cuda.get_device_properties(dev).multi_processor_count
        peak_inter = nsm * 128 * peaks.get("sm_max_mhz", 1965.0) * 1e6 / 8
        inter_rate = st0.pair_interactions * args.steps / (sum(walk_ms) * 1e-3)
        traffic, traffic_note = None, None
        tpath = os.path.join(ROOT, "profiles", "walk_traffic.json")
        if os.path.exists(tpath) and wl.name == "cfg2":
            tj = json.load(open(tpath))
            traffic, traffic_note = tj.get("dram_bytes_per_launch"), tj.get("kernel")
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_exec * 1e3 / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "workload_id": wl.name, "particles_per_gpu": n_local, "subhaloes_per_gpu": snap.nsub, "sum_nbound": nb_all,
                       "l2_policy": f"inputs ({n_local * 32 / 1e9:.2f} GB per GPU) are larger than the 126 MB L2; no explicit flush",
                       "timing": "CUDA events on the library stream around hbtu_execute, max over ranks", "wall_ms_per_step": float(np.mean(wall)) * 1e3,
                       "phase_ms": {"walk": float(np.mean(walk_ms)), "tree_build": float(np.mean(build_ms)), "partition_sort_reduce": float(np.mean(other_ms))},
                       "phase_ms_detail": dict(zip(("gather_bbox", "tree_build", "targets", "walk", "count_state_partition", "energy_sort_permute",
                                                    "frame_reduce", "kinematics_finalize"), [float(x) for x in st.phase_ms])),
                       "rounds": int(st0.rounds), "per_rank_ms_per_step": per_rank, "pair_interactions_per_step": inter_all,
                       "walk_fallbacks": int(st0.walk_fallbacks),
                       "multi_gpu": ("strong scaling: ONE snapshot replicated on every GPU, walk targets of every round dealt block-cyclically over the ranks, "
                                     "one NCCL sum all-reduce of 4 B per walk target per round, everything else replicated") if strong else
                                    ("weak scaling: one snapshot per GPU, hierarchies never span ranks; the only collective is the NCCL all-gather of the result records"),
                       "phase_rooflines": phase_rooflines(st0, float(np.mean(build_ms)), float(np.mean(other_ms)), peaks)},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(e2e_h2d), "d2h_bytes_per_step": int(e2e_d2h),
                    "ms_per_step": t_e2e * 1e3 / max(len(e2e_wall), 1), "steps": len(e2e_wall), "api": "hbtu_unbind_batch from pinned host buffers",
                    "last_call_breakdown": e2e_parts,
                    "overlap": "uploads on a copy stream in two waves: everything but the dominant root first, the dominant root behind the kernels of the deeper levels"},
            "gpu_launches": int(launches * args.steps),
            "roofline": {"bound": "fp32_issue", "achieved": inter_rate * FLOP_PER_INTERACTION / 1e12, "peak": peak_inter * FLOP_PER_INTERACTION / 1e12,
                         "unit": "TFLOP/s", "frac": inter_rate / peak_inter, "traffic": traffic, "traffic_kernel": traffic_note,
                         "kernel": "walk phase: walk_masked_kernel (segments >= 256 targets) + walk_small_kernel / walk_kernel (smaller ones)",
                         "interactions_per_s": inter_rate, "peak_interactions_per_s": peak_inter,
                         "peak_source": f"{nsm} SM x 128 fp32 lanes x sm_max_mhz ({peak_src}) / 8 issue slots per interaction (SURVEY.md 8(d)); 12 flop per interaction",
                         "note": "tensor cores deliberately unused (not a dense contraction); kernel share of the step = walk/total in config.phase_ms"},
        }
        if world == 1 and not args.profile:
            csnap, desc = wl.cpu_sample(args.particles, args.cpu_sample)
            dt, kind, ncpu, _, cres = run_cpu(csnap, workload=wl.name)
            out["cpu_baseline"] = {"value": csnap.npart / dt, "unit": UNIT, "cores": ncpu, "kind": kind, "sample": desc, "seconds": dt}
            out["parity"] = parity_block(ctx, e, csnap, cres, kind, workload=wl.name)
            ctx.close()  # free the HBM of the timed batch: the drop-in row brings its own context
            out["e2e"]["drop_in"] = bench_dropin_row(wl, min(args.particles, args.dropin_particles), dev, csnap, ncpu)
            ctx = UnbindContext(wl.params(local_rank))
            if wl.name == "cfg2":
                res = ctx.unbind_batch(e, snap, flags=flags, want_energy=False, order_buf=order_pinned)
                out["config"]["next_rows"] = {"profile_properties": bench_profile_row(ctx, e, snap, res, csnap, peaks),
                                              "mask_subhalos": bench_mask_row(ctx, snap, csnap, peaks),
                                              "particle_query": bench_idtable_row(ctx, snap.npart, csnap.npart, peaks)}
        print(json.dumps(out), flush=True)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
