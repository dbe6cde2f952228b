"""GPU parity of the source preparation through the C-ABI ``hbtu_mask_batch``:
SubhaloSnapshot_t::MaskSubhalos / SubhaloMasker_t::Mask (src/subhalo_tracking.cpp:793-841).  Index work: bit-exact."""
import os

import numpy as np
import pytest

import cases
from hbtplus_b200 import capi
from oracle import pyoracle as po
from test_gpu_parity import make_ctx  # noqa: F401  (fixture)
from test_oracle import kept_lists

pytestmark = pytest.mark.gpu
P = capi.make_params(box_size=62.5, softening=5e-3)


def same(part_offset, got, want):
    assert np.array_equal(got[0], want[0])
    for a, b in zip(kept_lists(part_offset, *got), kept_lists(part_offset, *want)):
        assert np.array_equal(a, b)


def test_mask_matches_reference_golden(make_ctx):
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "mask.npz"))
    ctx = make_ctx(P)
    got = ctx.mask_batch(z["part_offset"], z["ids"], z["nest_offset"], z["nest_list"], z["nbound"])
    same(z["part_offset"], got, (z["new_count"], z["keep"]))


@pytest.mark.parametrize("seed,nroots,scale", [(5, 40, 1.0), (6, 300, 0.3), (7, 8, 60.0)])
def test_mask_vs_oracle(make_ctx, oracle_lib, seed, nroots, scale):
    """Many small hierarchies, and a few large ones (1e5-particle lists: long equal-Id runs across hierarchies)."""
    part_offset, ids, nest_offset, nest_list, nbound = cases.case_mask(seed=seed, nroots=nroots, scale=scale)
    ctx = make_ctx(P)
    got = ctx.mask_batch(part_offset, ids, nest_offset, nest_list, nbound)
    want = po.mask_batch(oracle_lib, "hbto", P, part_offset, ids, nest_offset, nest_list, nbound)
    same(part_offset, got, want)
    st = ctx.stats()
    assert st.kernel_launches <= 20  # one batched pass, not one launch per hierarchy
    # idempotence (a size-independent property): masking the masked lists changes nothing
    kept = np.concatenate(kept_lists(part_offset, *got))
    po2 = np.concatenate([[0], np.cumsum(got[0])]).astype(np.int64)
    again = ctx.mask_batch(po2, ids[kept], nest_offset, nest_list, nbound)
    assert np.array_equal(again[0], got[0])


def test_mask_edge_cases(make_ctx, oracle_lib):
    ctx = make_ctx(P)
    # no nesting at all, negative and huge Ids, one empty list, everything duplicated inside one list
    part_offset = np.array([0, 4, 4, 10], np.int64)
    ids = np.array([7, 7, -3, 7, 2**62, 5, 5, 2**62, -3, 5], np.int64)
    nbound = np.array([4, 0, 6], np.int64)
    got = ctx.mask_batch(part_offset, ids, None, None, nbound)
    want = po.mask_batch(oracle_lib, "hbto", P, part_offset, ids, None, None, nbound)
    same(part_offset, got, want)
    assert got[0].tolist() == [2, 0, 3]
    # a malformed forest (a subhalo nested twice) is rejected, as in hbtu_unbind_batch
    with pytest.raises(Exception):
        ctx.mask_batch(part_offset, ids, np.array([0, 1, 2, 2], np.int64), np.array([2, 2], np.int32), nbound)
