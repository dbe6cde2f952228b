"""Host-side multi-GPU logic on CPU: cost-weighted LPT sharding of whole hierarchies and the world_size-2 (gloo)
gather of per-subhalo result records.  The per-rank compute is injected; here the oracle stands in for the CUDA
call (tests may use oracle/), which also proves that sharding by hierarchy does not change any result."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

import cases
from hbtplus_b200 import capi, sched, synth


def test_lpt_partition_balances_and_is_deterministic():
    rng = np.random.default_rng(0)
    costs = rng.pareto(1.2, 500) + 1
    a = sched.lpt_partition(costs, 8)
    b = sched.lpt_partition(costs, 8)
    assert np.array_equal(a, b)
    loads = np.bincount(a, weights=costs, minlength=8)
    assert loads.max() <= max(costs.max(), 1.05 * costs.sum() / 8)
    assert set(a.tolist()) == set(range(8))


def test_hierarchies_are_never_split():
    p, e, snap = cases.case_nested()
    roots, cost, root_of = sched.hierarchy_costs(snap.part_offset, snap.nest_offset, snap.nest_list)
    assert sorted(roots.tolist()) == [0, 8] and (cost > 0).all()
    seen = []
    for r in range(2):
        sub, mine = sched.shard_snapshot(snap, r, 2)
        seen += mine.tolist()
        assert len(set(root_of[mine].tolist())) <= 1 or r == 0
        assert sub.npart == np.diff(snap.part_offset)[mine].sum()
        for k, g in enumerate(mine):
            assert np.array_equal(sub.pos_mass[sub.part_offset[k]:sub.part_offset[k + 1]], snap.pos_mass[snap.part_offset[g]:snap.part_offset[g + 1]])
    assert sorted(seen) == list(range(snap.nsub))
    # cost model: n log2 n * iterations on the source capacity (north star)
    big = sched.hierarchy_costs(np.array([0, 1000]), None, None)[1][0]
    assert abs(big - 3 * 1000 * np.log2(1000)) < 1e-6


def _worker(rank, world, port, q):
    import torch.distributed as dist

    from oracle import pyoracle as po

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p, e, snap = cases.case_nested()
    sub, mine = sched.shard_snapshot(snap, rank, world)
    orc = po.load_oracle()
    orc.hbto_set_num_threads(1)
    r = po.run_batch(orc, "hbto", p, e, sub, flags=capi.HBTU_FLAG_TRUNCATE_SOURCE)
    table = sched.gather_records(r.io, mine, snap.nsub)
    if rank == 0:
        q.put(table.tobytes())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather_matches_single_process(oracle_lib):
    from oracle import pyoracle as po

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    raw = q.get(timeout=120)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    p, e, snap = cases.case_nested()
    table = np.frombuffer(raw, dtype=capi.SUBIO_DTYPE)
    want = po.run_batch(oracle_lib, "hbto", p, e, snap, flags=capi.HBTU_FLAG_TRUNCATE_SOURCE)
    for f in ("nbound", "mbound", "snapshot_index_of_death", "nsource", "avg_pos", "avg_vel"):
        assert np.array_equal(table[f], want.io[f]), f


def test_eagle_workload_shards_partition_the_snapshot_depth_first():
    """bench.py --workload cfg4: every rank draws the same global forest and keeps its own hierarchies.  The shards must
    partition the global subhalo list, keep every hierarchy whole, balance the LPT cost, and come out hierarchy by hierarchy
    with parents first (the layout the drop-in shim produces and hbtu_unbind_batch's pipeline needs)."""
    import os
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench

    wl = bench.WORKLOADS["cfg4"]
    world, particles = 4, 2.0e5  # per rank
    gsizes, gparent = wl.sizes(particles * world)
    seen = np.zeros(len(gsizes), np.int64)
    costs = []
    for rank in range(world):
        sizes, parent, gidx = wl.shard(particles, rank, world)
        seen[gidx] += 1
        assert np.array_equal(sizes, gsizes[gidx])
        # the parent of a local subhalo is the local copy of its global parent (hierarchies are whole)
        has = parent >= 0
        assert np.array_equal(gidx[parent[has]], gparent[gidx[has]]) and np.all(gparent[gidx[~has]] < 0)
        start = 0
        for s in range(len(sizes)):  # depth first: every hierarchy a contiguous range, parents in front
            if parent[s] < 0:
                start = s
            else:
                assert start <= parent[s] < s
        n = sizes.astype(np.float64)
        costs.append(float((n * np.log2(np.maximum(n, 2.0))).sum()))
    assert np.all(seen == 1)
    assert max(costs) / np.mean(costs) < 1.05
    # cfg 3 is generated depth first as well
    sizes, parent = bench.WORKLOADS["cfg3"].sizes(2.0e5)
    start = 0
    for s in range(len(sizes)):
        if parent[s] < 0:
            start = s
        else:
            assert start <= parent[s] < s
