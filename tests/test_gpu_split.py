"""SURVEY.md 8(e), the non-natural case: ONE batch executed by several cooperating contexts that deal every round's walk
targets among themselves (hbtu_set_walk_split / hbtu_split_group_*).  On one GPU the cooperating contexts are simply several
contexts of the same device, each on its own host thread, and the built-in peer-memory all-reduce runs between them - the same
code path as on NVLink peers.  The catalogue must be identical, bit for bit, to the unsplit execution."""
import numpy as np
import pytest

import cases
from hbtplus_b200 import capi, synth
from hbtplus_b200.unbind import SplitGroup, UnbindContext

pytestmark = pytest.mark.gpu


def _same(a, b):
    for f in a.io.dtype.names:
        assert np.array_equal(a.io[f], b.io[f]), f
    assert np.array_equal(a.order_offset, b.order_offset)
    n = int(a.order_offset[-1])
    assert np.array_equal(a.order[:n], b.order[:n])
    assert np.array_equal(a.energy[:n], b.energy[:n])


@pytest.mark.parametrize("nranks", [2, 3])
@pytest.mark.parametrize("periodic", [False, True])
def test_walk_split_matches_single_context(nranks, periodic):
    p = capi.make_params(box_size=62.5, softening=5e-3, periodic=periodic)
    e = capi.make_epoch(0.9, snapshot_index=12)
    sizes = [150_000, 9000, 2500, 700, 60, 25, 12000, 300]
    parent = [-1, 0, 1, 2, 0, 4, -1, 6]
    snap = synth.make_snapshot(sizes, seed=41 + periodic, parent=parent, wrap=periodic, f_contam=0.3)
    flags = capi.HBTU_FLAG_TRUNCATE_SOURCE
    one = UnbindContext(p)
    want = one.unbind_batch(e, snap, flags=flags)
    one.close()
    ctxs = [UnbindContext(p) for _ in range(nranks)]
    group = SplitGroup(ctxs)
    for c in ctxs:
        c.set_counting(True)
    got = group.run(lambda r, c: c.unbind_batch(e, snap, flags=flags))
    inter = sum(c.stats().pair_interactions for c in ctxs)
    for r in range(nranks):
        _same(got[r], want)
    # the walk really was divided: every member did a share of the interactions, together all of them
    one = UnbindContext(p)
    one.set_counting(True)
    one.unbind_batch(e, snap, flags=flags)
    total = one.stats().pair_interactions
    one.close()
    assert inter == total
    shares = [c.stats().pair_interactions / total for c in ctxs]
    assert min(shares) > 0.5 / nranks and max(shares) < 1.6 / nranks, shares
    group.close()
    # after the group is gone the contexts work alone again
    alone = ctxs[0].unbind_batch(e, snap, flags=flags)
    _same(alone, want)
    for c in ctxs:
        c.close()


def test_walk_split_with_a_python_allreduce():
    """The callback form (what bench.py uses with torch.distributed under torchrun), here with a trivial single-member
    'collective' that also counts its calls: one per round."""
    p, e, snap = cases.case_nested()
    ctx = UnbindContext(p)
    want = ctx.unbind_batch(e, snap)
    calls = []
    ctx.set_walk_split(1, 2, lambda ptr, count, stream: calls.append(count))  # rank 1 of 2 with nobody else: rank 0's targets stay 0
    half = ctx.unbind_batch(e, snap)
    assert len(calls) == ctx.stats().rounds and all(c > 0 for c in calls)
    assert not np.array_equal(half.io["nbound"], want.io["nbound"])  # the missing half matters - the split is real
    ctx.set_walk_split(0, 1, None)
    _same(ctx.unbind_batch(e, snap), want)
    ctx.close()


def test_torch_allreduce_wraps_the_device_pointer():
    """sched.torch_allreduce (the NCCL callback of bench.py --scaling strong) on a one-rank process group: the raw device pointer
    becomes a tensor without a copy and the collective runs on it in place."""
    import os

    import torch
    import torch.distributed as dist

    from hbtplus_b200 import sched

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29531")
    created = False
    if not dist.is_initialized():
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
        created = True
    try:
        dev = torch.device("cuda", 0)
        x = torch.arange(1000, dtype=torch.float32, device=dev)
        fn = sched.torch_allreduce(dev)
        fn(x.data_ptr(), 1000, 0)
        assert torch.equal(x, torch.arange(1000, dtype=torch.float32, device=dev))
        y = torch.as_tensor(sched._DevicePtr(x.data_ptr(), 1000), device=dev)
        y += 1  # same memory
        assert float(x[0]) == 1.0 and float(x[999]) == 1000.0
    finally:
        if created:
            dist.destroy_process_group()
