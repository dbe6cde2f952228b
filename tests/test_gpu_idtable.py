"""GPU parity of the particle query through the C-ABI ``hbtu_idtable_build`` / ``hbtu_idtable_query``:
MappedIndexTable_t::Fill / GetIndices (src/hash.tpp:18-32, src/hash_remote.tpp:9-88).  Index work: bit-exact."""
import numpy as np
import pytest

from hbtplus_b200 import capi
from oracle import pyoracle as po
from test_gpu_parity import make_ctx  # noqa: F401  (fixture)
from test_oracle import idtable_case

pytestmark = pytest.mark.gpu
P = capi.make_params(box_size=62.5, softening=5e-3)


@pytest.mark.parametrize("n,nq,wide", [(1, 5, False), (1000, 3000, False), (300000, 500000, True), (2000000, 1000000, True)])
def test_idtable_vs_oracle(make_ctx, oracle_lib, n, nq, wide):
    ids, q = idtable_case(n, n, nq, wide)
    ctx = make_ctx(P)
    ctx.idtable_build(ids)
    got = ctx.idtable_query(q)
    want = po.idtable_query(oracle_lib, "hbto", P, ids, q)
    assert np.array_equal(got, want)
    found = got >= 0
    assert np.array_equal(ids[got[found]], q[found])  # the defining property, size-independent
    # a second query against the resident table, then a rebuild
    assert np.array_equal(ctx.idtable_query(q[::-1].copy()), want[::-1])
    ctx.idtable_build(ids[::-1].copy())
    again = ctx.idtable_query(q)
    assert np.array_equal(again[found], len(ids) - 1 - got[found]) and np.all(again[~found] == -1)


def test_idtable_vs_reference(make_ctx, ref_lib):
    ids, q = idtable_case(9, 50000, 80000, wide=False)
    ctx = make_ctx(P)
    ctx.idtable_build(ids)
    assert np.array_equal(ctx.idtable_query(q), po.idtable_query(ref_lib, "hbtref", P, ids, q))


def test_idtable_edge_cases(make_ctx):
    ctx = make_ctx(P)
    ctx.idtable_build(np.zeros(0, np.int64))
    assert np.array_equal(ctx.idtable_query(np.array([1, 2, 3])), [-1, -1, -1])
    ctx.idtable_build(np.array([5, -7, 2**62, 5, 0]))
    assert np.array_equal(ctx.idtable_query(np.array([5, -7, 2**62, 0, 1, -8, 2**62 - 1])), [0, 1, 2, 4, -1, -1, -1])  # duplicates: lowest index
    # Id -1 is the table's empty-slot marker (and the reference's NullParticleId): a table that really holds it still answers
    assert np.array_equal(ctx.idtable_query(np.array([-1, 5])), [-1, 0])
    ctx.idtable_build(np.array([3, -1, 9, -1]))
    assert np.array_equal(ctx.idtable_query(np.array([-1, 9, 3, -2, np.iinfo(np.int64).min])), [1, 2, 0, -1, -1])
    # a table small enough that the probe sequence wraps around the end of the slot array
    ids = np.arange(31, dtype=np.int64) * 1000003 + 17
    ctx.idtable_build(ids)
    assert np.array_equal(ctx.idtable_query(np.concatenate([ids[::-1], ids + 1])), np.concatenate([np.arange(30, -1, -1), np.full(31, -1)]))
