"""Sharding of a snapshot's subhalo hierarchies over the GPUs of one box.

The unit of work is a whole hierarchy (a root subhalo with everything nested below it): hierarchies are
independent, so ranks never exchange particles.  Units are dealt out longest-processing-time-first with the
cost model of the north star, cost = sum over members of n*log2(n)*iterations.  The only communication is the
gather of the small per-subhalo result records (torch.distributed: NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import heapq

import numpy as np

from .synth import Snapshot

ITERATIONS = 3.0  # typical potential evaluations per particle


def forest_parents(nsub: int, nest_offset, nest_list) -> np.ndarray:
    parent = np.full(nsub, -1, np.int64)
    if nest_offset is not None:
        counts = np.diff(nest_offset)
        parent[np.asarray(nest_list, np.int64)] = np.repeat(np.arange(nsub), counts)
    return parent


def roots_of(parent: np.ndarray) -> np.ndarray:
    """Root index of every subhalo (path compression by repeated squaring)."""
    nsub = len(parent)
    up = np.where(parent >= 0, parent, np.arange(nsub))
    while True:
        nxt = up[up]
        if np.array_equal(nxt, up):
            return up
        up = nxt


def hierarchy_costs(part_offset, nest_offset, nest_list):
    """(roots, cost per root) with cost = sum n log2 n * ITERATIONS over the members' source capacities."""
    nsub = len(part_offset) - 1
    parent = forest_parents(nsub, nest_offset, nest_list)
    root = roots_of(parent)
    n = np.diff(part_offset).astype(np.float64)
    # a parent also unbinds what its descendants feed it: charge it the capacity, not only its own particles
    cap = n.copy()
    depth = np.zeros(nsub, np.int64)
    q = parent.copy()
    while (q >= 0).any():
        depth[q >= 0] += 1
        q = np.where(q >= 0, parent[np.maximum(q, 0)], -1)
    for s in np.argsort(-depth, kind="stable"):
        if parent[s] >= 0:
            cap[parent[s]] += cap[s]
    cost = ITERATIONS * cap * np.log2(np.maximum(cap, 2.0))
    roots = np.unique(root)
    total = np.zeros(nsub)
    np.add.at(total, root, cost)
    return roots, total[roots], root


def lpt_partition(costs: np.ndarray, nranks: int) -> np.ndarray:
    """Longest-processing-time-first: returns the rank of every unit; deterministic on ties."""
    assign = np.zeros(len(costs), np.int64)
    heap = [(0.0, r) for r in range(nranks)]
    heapq.heapify(heap)
    for u in np.argsort(-np.asarray(costs), kind="stable"):
        load, r = heapq.heappop(heap)
        assign[u] = r
        heapq.heappush(heap, (load + float(costs[u]), r))
    return assign


def shard_snapshot(snap: Snapshot, rank: int, world: int):
    """The sub-batch of `snap` owned by `rank`: (Snapshot, global subhalo index of every local subhalo)."""
    roots, cost, root_of = hierarchy_costs(snap.part_offset, snap.nest_offset, snap.nest_list)
    owner_of_root = np.full(snap.nsub, -1, np.int64)
    owner_of_root[roots] = lpt_partition(cost, world)
    mine = np.nonzero(owner_of_root[root_of] == rank)[0]
    local_of = np.full(snap.nsub, -1, np.int64)
    local_of[mine] = np.arange(len(mine))
    sizes = np.diff(snap.part_offset)[mine]
    part_offset = np.zeros(len(mine) + 1, np.int64)
    np.cumsum(sizes, out=part_offset[1:])
    idx = np.concatenate([np.arange(snap.part_offset[s], snap.part_offset[s + 1]) for s in mine]) if len(mine) else np.zeros(0, np.int64)
    nest_offset = nest_list = None
    if snap.nest_offset is not None:
        lists = [local_of[snap.nest_list[snap.nest_offset[s]:snap.nest_offset[s + 1]]] for s in mine]
        nest_offset = np.zeros(len(mine) + 1, np.int64)
        nest_offset[1:] = np.cumsum([len(l) for l in lists])
        nest_list = (np.concatenate(lists) if lists else np.zeros(0)).astype(np.int32)
        assert (nest_list >= 0).all()  # hierarchies are never split
    return Snapshot(part_offset, snap.pos_mass[idx], snap.vel[idx], nest_offset, nest_list, snap.io[mine].copy()), mine


def gather_records(io_local: np.ndarray, index_local: np.ndarray, nsub_total: int, device=None) -> np.ndarray | None:
    """All-gather the per-subhalo result records (SUBIO_DTYPE, ~150 B each) and scatter them to global order.
    Every rank returns the full table.  Uses the default process group (nccl or gloo)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size()
    dev = device if device is not None else "cpu"
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    counts[dist.get_rank()] = len(io_local)
    dist.all_reduce(counts)
    nmax = int(counts.max())
    rec = io_local.dtype.itemsize
    buf = torch.zeros(nmax * rec + nmax * 8, dtype=torch.uint8, device=dev)
    payload = np.concatenate([io_local.view(np.uint8).reshape(-1), index_local.astype(np.int64).view(np.uint8)])
    buf[: len(io_local) * rec] = torch.from_numpy(payload[: len(io_local) * rec].copy()).to(dev)
    buf[nmax * rec: nmax * rec + len(io_local) * 8] = torch.from_numpy(payload[len(io_local) * rec:].copy()).to(dev)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    full = np.zeros(nsub_total, io_local.dtype)
    for r in range(world):
        k = int(counts[r])
        raw = out[r].cpu().numpy()
        io_r = raw[: k * rec].view(io_local.dtype)
        idx_r = raw[nmax * rec: nmax * rec + k * 8].view(np.int64)
        full[idx_r] = io_r
    return full


class _DevicePtr:
    """A raw device pointer as a CUDA array (for torch.as_tensor): the library's staging array, owned by the library."""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def torch_allreduce(device):
    """The all-reduce callback of UnbindContext.set_walk_split for one-process-per-GPU jobs (torchrun): a sum all-reduce of the
    library's staging array over the default process group (NCCL over NVLink / NVSwitch), complete when it returns.  This is
    the walk split's only collective: 4 bytes per walk target and round (the non-natural case)."""
    import torch
    import torch.distributed as dist

    def allreduce(ptr: int, count: int, stream: int):
        t = torch.as_tensor(_DevicePtr(ptr, count), device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        torch.cuda.current_stream(device).synchronize()

    return allreduce
