"""Synthetic subhalo snapshots shaped like BASELINE.json's configs.

Units follow the reference's defaults (src/config_parser.cpp:95-96): Mpc/h, km/s,
1e10 Msun/h, hence G = 43.0071.  Every subhalo *source* is a truncated Hernquist blob
in rough Jeans equilibrium plus a fraction of hot "host contaminant" particles, so that
the unbinding loop needs several iterations.  Everything is vectorised over particles, so
the same code generates a 1e3-particle test case and (with ``xp=torch`` on the GPU) the
1.8e8-particle AqA2-shaped bench case.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .capi import SUBIO_DTYPE

G_INTERNAL = 43.0071


@dataclass
class Snapshot:
    """One batch in the layout of ``hbtu_unbind_batch`` (include/hbt_unbind.h)."""

    part_offset: np.ndarray  # int64 [nsub+1]
    pos_mass: np.ndarray  # float32 [N,4]
    vel: np.ndarray  # float32 [N,4]
    nest_offset: np.ndarray | None  # int64 [nsub+1]
    nest_list: np.ndarray | None  # int32
    io: np.ndarray  # SUBIO_DTYPE [nsub]

    @property
    def nsub(self) -> int:
        return len(self.part_offset) - 1

    @property
    def npart(self) -> int:
        return int(self.part_offset[-1])


def subhalo_sizes(rng: np.random.Generator, nsub: int, n_min: int, n_max: int, slope: float = -1.9) -> np.ndarray:
    """dN/dn ~ n^slope on [n_min, n_max] (inverse-CDF sampling)."""
    u = rng.random(nsub)
    a = slope + 1.0
    lo, hi = float(n_min) ** a, float(n_max) ** a
    n = (lo + u * (hi - lo)) ** (1.0 / a)
    return np.clip(n.astype(np.int64), n_min, n_max)


def _hernquist_particles(rng, n_tot, sub_of, a_of, mtot_of, f_contam_of, contam_scale, contam_hot):
    """Positions/velocities relative to the owning subhalo's centre.

    Hernquist: M(<r)/M = r^2/(r+a)^2, truncated at r = 10 a.  Velocities: isotropic Gaussian with the
    approximate 1-D dispersion sigma^2 = G M(<r) / (3 r) * (r+a)/(r+a) (close to Jeans for r ~ a).
    Contaminants: envelope x contam_scale, dispersion x contam_hot."""
    a = a_of[sub_of]
    mtot = mtot_of[sub_of]
    umax = (10.0 / 11.0) ** 2
    u = rng.random(n_tot) * umax
    su = np.sqrt(u)
    r = a * su / (1.0 - su)
    is_c = rng.random(n_tot) < f_contam_of[sub_of]
    r = np.where(is_c, r * contam_scale, r)
    cost = rng.uniform(-1.0, 1.0, n_tot)
    sint = np.sqrt(1.0 - cost * cost)
    phi = rng.uniform(0.0, 2.0 * np.pi, n_tot)
    x = np.stack([r * sint * np.cos(phi), r * sint * np.sin(phi), r * cost], axis=1)
    menc = mtot * (r / (r + a)) ** 2
    sigma = np.sqrt(G_INTERNAL * menc / (3.0 * np.maximum(r, 1e-3 * a)))
    sigma = np.where(is_c, sigma * contam_hot, sigma)
    v = rng.standard_normal((n_tot, 3)) * sigma[:, None]
    return x, v


def make_snapshot(
    sizes,
    *,
    seed: int = 20240001,
    box_size: float = 62.5,
    particle_mass: float = 0.086,
    f_contam: float = 0.2,
    contam_scale: float = 2.0,
    contam_hot: float = 4.0,
    parent=None,
    scale_factor: float = 1.0,
    centre=None,
    wrap: bool = True,
    frame_noise: float = 0.05,
    mass_scatter: float = 0.0,
) -> Snapshot:
    """Build a batch of ``len(sizes)`` subhalo sources.

    parent[s] = batch index of the subhalo that nests s (or -1).  A nested subhalo is placed inside
    its parent (within ~1.5 scale radii) and orbits it at about the local circular speed."""
    rng = np.random.default_rng(seed)
    sizes = np.asarray(sizes, np.int64)
    nsub = len(sizes)
    part_offset = np.zeros(nsub + 1, np.int64)
    np.cumsum(sizes, out=part_offset[1:])
    n_tot = int(part_offset[-1])
    mtot = np.maximum(sizes, 1) * particle_mass
    # virial-ish radius from M = 100 H0^2 r^3 / G with H0 = 100 (internal units)
    rvir = (G_INTERNAL * mtot / 1e6) ** (1.0 / 3.0)
    a_of = rvir / 4.0
    fc = np.full(nsub, f_contam) if np.isscalar(f_contam) else np.asarray(f_contam, float)
    sub_of = np.repeat(np.arange(nsub), sizes)
    x, v = _hernquist_particles(rng, n_tot, sub_of, a_of, mtot, fc, contam_scale, contam_hot)

    centres = rng.random((nsub, 3)) * box_size if centre is None else np.broadcast_to(np.asarray(centre, float), (nsub, 3)).copy()
    bulk = rng.standard_normal((nsub, 3)) * 200.0
    if parent is not None:
        parent = np.asarray(parent, np.int64)
        depth = np.zeros(nsub, np.int64)
        for s in range(nsub):
            q, d = s, 0
            while parent[q] >= 0:
                q, d = parent[q], d + 1
            depth[s] = d
        for s in np.argsort(depth, kind="stable"):
            p = parent[s]
            if p < 0:
                continue
            d = rng.standard_normal(3)
            d /= np.linalg.norm(d)
            rad = a_of[p] * rng.uniform(0.3, 1.5)
            centres[s] = centres[p] + d * rad
            vc = np.sqrt(G_INTERNAL * mtot[p] * (rad / (rad + a_of[p])) ** 2 / rad)
            t = np.cross(d, rng.standard_normal(3))
            t /= np.linalg.norm(t)
            bulk[s] = bulk[p] + t * vc * 0.8
    pos = x + centres[sub_of]
    if wrap:
        pos = np.mod(pos, box_size)
    vel = v + bulk[sub_of]
    mass = np.full(n_tot, particle_mass)
    if mass_scatter > 0:
        mass = mass * np.exp(rng.standard_normal(n_tot) * mass_scatter)
    pos_mass = np.empty((n_tot, 4), np.float32)
    pos_mass[:, :3] = pos
    pos_mass[:, 3] = mass
    vel4 = np.zeros((n_tot, 4), np.float32)
    vel4[:, :3] = vel

    io = np.zeros(nsub, SUBIO_DTYPE)
    noise = rng.standard_normal((nsub, 3)) * frame_noise
    ref_pos = centres + noise * a_of[:, None]
    if wrap:
        ref_pos = np.mod(ref_pos, box_size)
    io["avg_pos"] = ref_pos.astype(np.float32)
    io["avg_vel"] = (bulk + noise[:, ::-1] * 20.0).astype(np.float32)
    for s in range(nsub):
        b = part_offset[s]
        if sizes[s] > 0:
            io["mostbound_pos"][s] = pos_mass[b, :3]
            io["mostbound_vel"][s] = vel4[b, :3]
    io["nbound"] = sizes
    io["sink_track_id"] = -1
    io["snapshot_index_of_death"] = -1
    io["snapshot_index_of_sink"] = -1

    nest_offset = nest_list = None
    if parent is not None:
        lists = [[] for _ in range(nsub)]
        for s in range(nsub):
            if parent[s] >= 0:
                lists[parent[s]].append(s)
        nest_offset = np.zeros(nsub + 1, np.int64)
        nest_offset[1:] = np.cumsum([len(l) for l in lists])
        nest_list = np.array([c for l in lists for c in l], np.int32)
    return Snapshot(part_offset, pos_mass, vel4, nest_offset, nest_list, io)


def forest_depth(parent: np.ndarray) -> np.ndarray:
    """Nesting depth of every subhalo of a parent forest (roots: 0), vectorised."""
    parent = np.asarray(parent, np.int64)
    depth = np.zeros(len(parent), np.int64)
    cur = parent.copy()
    while (cur >= 0).any():
        live = cur >= 0
        depth[live] += 1
        cur = np.where(live, parent[np.maximum(cur, 0)], -1)
    return depth


def nest_forest(rng: np.random.Generator, sizes: np.ndarray, max_depth: int = 4, p_nest: float = 0.5, root: int | None = 0) -> np.ndarray:
    """Random nesting: every subhalo is nested in a LARGER one (or in `root`), depth <= max_depth.

    Mirrors what NestSubhalos produces (src/subhalo_tracking.cpp:669): satellites hang off more massive hosts."""
    nsub = len(sizes)
    order = np.argsort(-np.asarray(sizes), kind="stable")
    parent = np.full(nsub, -1, np.int64)
    depth = np.zeros(nsub, np.int64)
    placed = []
    for s in order:
        if root is not None and s == root:
            placed.append(s)
            continue
        cand = -1
        if placed and rng.random() < p_nest:
            c = placed[int(rng.integers(0, len(placed)))]
            if depth[c] + 1 <= max_depth and sizes[c] > sizes[s]:
                cand = c
        if cand < 0 and root is not None:
            cand = root
        parent[s] = cand
        depth[s] = depth[cand] + 1 if cand >= 0 else 0
        placed.append(s)
    return parent


def dfs_layout(sizes: np.ndarray, parent: np.ndarray, return_order: bool = False):
    """Relabel a nest forest depth first: every hierarchy (a root and everything nested in it) becomes a contiguous index range
    with parents in front of their children - the order in which the reference visits subhaloes (RecursiveUnbind from every
    host's central, src/subhalo_unbind.cpp:479-493) and in which integration/subhalo_unbind_b200.cpp::add_hierarchy lays a batch
    out.  Returns (sizes, parent) in the new labelling (and, with return_order, the old index of every new one)."""
    sizes = np.asarray(sizes, np.int64)
    parent = np.asarray(parent, np.int64)
    nsub = len(sizes)
    order = np.argsort(parent, kind="stable")  # children grouped by parent, roots (-1) first, original order inside a group
    first = np.searchsorted(parent[order], np.arange(-1, nsub), side="left")
    last = np.searchsorted(parent[order], np.arange(-1, nsub), side="right")
    new_of = np.full(nsub, -1, np.int64)
    out = []
    stack = list(order[first[0]:last[0]][::-1])
    while stack:
        s = stack.pop()
        new_of[s] = len(out)
        out.append(s)
        stack.extend(order[first[s + 1]:last[s + 1]][::-1])
    out = np.asarray(out, np.int64)
    assert len(out) == nsub, "parent[] is not a forest"
    new_parent = np.where(parent[out] >= 0, new_of[np.maximum(parent[out], 0)], -1)
    return (sizes[out], new_parent, out) if return_order else (sizes[out], new_parent)


def make_snapshot_torch(sizes, *, device, seed: int = 20240002, box_size: float = 100.0, particle_mass: float = 1e-6,
                        f_contam: float = 0.2, contam_scale: float = 2.0, contam_hot: float = 4.0, parent=None, centre=None,
                        wrap: bool = False, frame_noise: float = 0.05, pin: bool = True) -> Snapshot:
    """Same construction as make_snapshot, with the per-particle work done by torch on `device`
    (a 1.8e8-particle AqA2-shaped batch takes seconds instead of minutes).  Returns HOST arrays
    (pinned when pin=True) - torch is used here only as an array library / allocator."""
    import torch

    rng = np.random.default_rng(seed)
    sizes = np.asarray(sizes, np.int64)
    nsub = len(sizes)
    part_offset = np.zeros(nsub + 1, np.int64)
    np.cumsum(sizes, out=part_offset[1:])
    n_tot = int(part_offset[-1])
    mtot = np.maximum(sizes, 1) * particle_mass
    rvir = (G_INTERNAL * mtot / 1e6) ** (1.0 / 3.0)
    a_of = rvir / 4.0
    centres = rng.random((nsub, 3)) * box_size if centre is None else np.broadcast_to(np.asarray(centre, float), (nsub, 3)).copy()
    bulk = rng.standard_normal((nsub, 3)) * 200.0
    if parent is not None:
        parent = np.asarray(parent, np.int64)
        depth = forest_depth(parent)
        dirs = rng.standard_normal((nsub, 3))
        dirs /= np.linalg.norm(dirs, axis=1)[:, None]
        tang = np.cross(dirs, rng.standard_normal((nsub, 3)))
        tang /= np.linalg.norm(tang, axis=1)[:, None]
        frac = rng.uniform(0.3, 1.5, nsub)
        for d in range(1, int(depth.max()) + 1 if nsub else 0):  # level by level: a subhalo only needs its parent's final centre
            idx = np.nonzero(depth == d)[0]
            p = parent[idx]
            rad = a_of[p] * frac[idx]
            centres[idx] = centres[p] + dirs[idx] * rad[:, None]
            vc = np.sqrt(G_INTERNAL * mtot[p] * (rad / (rad + a_of[p])) ** 2 / rad)
            bulk[idx] = bulk[p] + tang[idx] * (vc * 0.8)[:, None]
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    t = lambda x, dt=torch.float32: torch.as_tensor(x, dtype=dt, device=device)
    sizes_t = torch.as_tensor(sizes, device=device)
    sub_of = torch.repeat_interleave(torch.arange(nsub, device=device), sizes_t)
    a = t(a_of)[sub_of]
    umax = (10.0 / 11.0) ** 2
    su = torch.sqrt(torch.rand(n_tot, generator=g, device=device) * umax)
    r = a * su / (1.0 - su)
    del su
    is_c = torch.rand(n_tot, generator=g, device=device) < f_contam
    r = torch.where(is_c, r * contam_scale, r)
    cost = torch.rand(n_tot, generator=g, device=device) * 2 - 1
    sint = torch.sqrt(1 - cost * cost)
    phi = torch.rand(n_tot, generator=g, device=device) * (2 * np.pi)
    pos_mass = torch.empty((n_tot, 4), dtype=torch.float32, device=device)
    c_t = t(centres)[sub_of]
    pos_mass[:, 0] = r * sint * torch.cos(phi) + c_t[:, 0]
    pos_mass[:, 1] = r * sint * torch.sin(phi) + c_t[:, 1]
    pos_mass[:, 2] = r * cost + c_t[:, 2]
    del c_t, cost, sint, phi
    if wrap:
        pos_mass[:, :3] = torch.remainder(pos_mass[:, :3], box_size)
    pos_mass[:, 3] = particle_mass
    menc = t(mtot)[sub_of] * (r / (r + a)) ** 2
    sigma = torch.sqrt(G_INTERNAL * menc / (3.0 * torch.maximum(r, 1e-3 * a)))
    del menc, r, a
    sigma = torch.where(is_c, sigma * contam_hot, sigma)
    del is_c
    vel = torch.zeros((n_tot, 4), dtype=torch.float32, device=device)
    b_t = t(bulk)[sub_of]
    for j in range(3):
        vel[:, j] = torch.randn(n_tot, generator=g, device=device) * sigma + b_t[:, j]
    del b_t, sigma, sub_of

    def to_host(x):
        h = torch.empty(x.shape, dtype=x.dtype, pin_memory=pin and torch.cuda.is_available())
        h.copy_(x)
        return h.numpy()

    first = torch.as_tensor(np.minimum(part_offset[:-1], max(n_tot - 1, 0)), device=device)
    mb_pos = pos_mass[first, :3].cpu().numpy() if n_tot else np.zeros((nsub, 3), np.float32)
    mb_vel = vel[first, :3].cpu().numpy() if n_tot else np.zeros((nsub, 3), np.float32)
    pm_h, vel_h = to_host(pos_mass), to_host(vel)
    del pos_mass, vel
    io = np.zeros(nsub, SUBIO_DTYPE)
    noise = rng.standard_normal((nsub, 3)) * frame_noise
    ref_pos = centres + noise * a_of[:, None]
    if wrap:
        ref_pos = np.mod(ref_pos, box_size)
    io["avg_pos"] = ref_pos.astype(np.float32)
    io["avg_vel"] = (bulk + noise[:, ::-1] * 20.0).astype(np.float32)
    io["mostbound_pos"] = mb_pos
    io["mostbound_vel"] = mb_vel
    io["nbound"] = sizes
    io["sink_track_id"] = -1
    io["snapshot_index_of_death"] = -1
    io["snapshot_index_of_sink"] = -1
    nest_offset = nest_list = None
    if parent is not None:
        order = np.argsort(parent, kind="stable")
        order = order[parent[order] >= 0]
        counts = np.bincount(parent[order], minlength=nsub)
        nest_offset = np.zeros(nsub + 1, np.int64)
        np.cumsum(counts, out=nest_offset[1:])
        nest_list = order.astype(np.int32)
    return Snapshot(part_offset, pm_h, vel_h, nest_offset, nest_list, io)


def aqa2_sizes(rng: np.random.Generator, n_total: float = 1.8e8, central_frac: float = 0.72, nsub: int = 40000, n_max: float = 5e6):
    """AqA2-shaped size list (SURVEY.md section 8(d) cfg 2): one central source + `nsub` subhaloes with
    dN/dn ~ n^-1.9 on [20, n_max], rescaled so that the batch holds ~n_total particles."""
    n_central = int(n_total * central_frac)
    sub = subhalo_sizes(rng, nsub, 20, int(n_max))
    want = n_total - n_central
    for _ in range(4):  # rescale the draw to the wanted particle total, keeping 20 <= n <= n_max
        sub = np.clip((sub * (want / sub.sum())).astype(np.int64), 20, int(n_max))
    return np.concatenate([[n_central], sub]).astype(np.int64)
