"""The pipelined form of ``hbtu_unbind_batch`` (capi.cu::unbind_batch_pipelined): a batch of many independent hierarchies is run in
two parts, the second uploading behind the kernels of the first.  Hierarchies never interact
(src/subhalo_unbind.cpp:479-513 visits them one host halo at a time), so the catalogue must not depend on it - bit for bit."""
import ctypes as C

import numpy as np
import pytest

from hbtplus_b200 import capi, synth
from oracle import pyoracle as po
from test_gpu_parity import check_batch, make_ctx  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu


def forest(seed, nsub, depth_first=True):
    rng = np.random.default_rng(seed)
    sizes = synth.subhalo_sizes(rng, nsub, 20, 4000)
    parent = synth.nest_forest(rng, sizes, max_depth=3, p_nest=0.3, root=None)
    if depth_first:
        sizes, parent = synth.dfs_layout(sizes, parent)
    return sizes, parent


def set_pipeline(lib, value):
    lib.hbtu_set_tuning.argtypes = [C.c_char_p, C.c_int64]
    lib.hbtu_get_tuning.argtypes = [C.c_char_p]
    lib.hbtu_get_tuning.restype = C.c_int64
    old = lib.hbtu_get_tuning(b"pipeline_min_particles")
    assert lib.hbtu_set_tuning(b"pipeline_min_particles", value) == 0
    return old


@pytest.mark.parametrize("periodic,max_sample", [(False, 0), (True, 0), (False, 300)], ids=["open", "periodic", "sampled"])
def test_pipelined_batch_is_bit_identical(make_ctx, oracle_lib, periodic, max_sample):
    """max_sample = 300: the sampled mode permutes every source above 300 particles with a key of (seed, subhalo index, position) -
    the index is the subhalo's place in the CALLER's batch, whichever part it runs in"""
    sizes, parent = forest(31 + periodic, 1500)
    p = capi.make_params(box_size=62.5, softening=5e-3, periodic=periodic, max_sample_size=max_sample, shuffle_seed=5)
    e = capi.make_epoch(1.0, snapshot_index=12)
    snap = synth.make_snapshot(sizes, seed=77, box_size=62.5, parent=parent, wrap=periodic)
    ctx = make_ctx(p)
    old = set_pipeline(ctx._lib, 0)
    try:
        whole = ctx.unbind_batch(e, snap, flags=capi.HBTU_FLAG_TRUNCATE_SOURCE)
        st_whole = ctx.stats()
        set_pipeline(ctx._lib, 1000)  # any batch of >= 64 subhaloes / 16 hierarchies qualifies
        parts = ctx.unbind_batch(e, snap, flags=capi.HBTU_FLAG_TRUNCATE_SOURCE)
        st_parts = ctx.stats()
    finally:
        set_pipeline(ctx._lib, old)
    assert st_parts.rounds > st_whole.rounds  # it really ran in parts (every part has its own rounds)
    assert st_parts.walk_targets == st_whole.walk_targets and st_parts.tree_sources == st_whole.tree_sources
    assert st_parts.h2d_bytes >= snap.npart * 32
    assert whole.io.tobytes() == parts.io.tobytes()  # every record field, bit for bit
    assert np.array_equal(whole.order_offset, parts.order_offset)
    n = whole.order_offset[-1]
    assert np.array_equal(whole.order[:n], parts.order[:n])
    assert np.array_equal(whole.energy[:n], parts.energy[:n])
    if max_sample == 0:  # and it is the reference's catalogue (sampled mode against the oracle: test_gpu_parity.py)
        want = po.run_batch(oracle_lib, "hbto", p, e, snap, flags=capi.HBTU_FLAG_TRUNCATE_SOURCE)
        check_batch(snap, parts, want, name=f"pipelined periodic={periodic}")
    # the resident-batch shortcut must refuse: only the last part is resident
    with pytest.raises(Exception, match="pipelined"):
        ctx.profile_executed(np.zeros(snap.nsub, capi.PROFILEIO_DTYPE))


def test_layouts_that_do_not_qualify_run_in_one_piece(make_ctx):
    """hierarchies that are not contiguous index ranges (children listed elsewhere) and batches with a dominant hierarchy"""
    p = capi.make_params(box_size=62.5, softening=5e-3, periodic=False)
    e = capi.make_epoch(1.0, snapshot_index=12)
    ctx = make_ctx(p)
    old = set_pipeline(ctx._lib, 1000)
    try:
        sizes, parent = forest(5, 600, depth_first=False)
        snap = synth.make_snapshot(sizes, seed=78, box_size=62.5, parent=parent, wrap=False)
        a = ctx.unbind_batch(e, snap)
        rounds_a = ctx.stats().rounds
        set_pipeline(ctx._lib, 0)
        b = ctx.unbind_batch(e, snap)
        assert ctx.stats().rounds == rounds_a and a.io.tobytes() == b.io.tobytes()
        # a hierarchy above a quarter of the batch: the two-wave upload path
        set_pipeline(ctx._lib, 1000)
        sizes, parent = forest(6, 300)
        sizes = np.concatenate([[int(sizes.sum())], sizes])
        parent = np.concatenate([[-1], np.where(parent >= 0, parent + 1, -1)])
        snap = synth.make_snapshot(sizes, seed=79, box_size=62.5, parent=parent, wrap=False)
        a = ctx.unbind_batch(e, snap)
        rounds_a = ctx.stats().rounds
        set_pipeline(ctx._lib, 0)
        b = ctx.unbind_batch(e, snap)
        assert ctx.stats().rounds == rounds_a and a.io.tobytes() == b.io.tobytes()
    finally:
        set_pipeline(ctx._lib, old)
