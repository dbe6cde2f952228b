"""CPU check of the merge harness (oracle/ref_build/ref_capi.cpp::hbtref_merge_subhalos): the reference's own
SubhaloSnapshot_t::MergeSubhalos with MergeTrappedSubhalos on (src/subhalo_merge.cpp:187-220) gives the same catalogue with 1
and with 8 OpenMP threads in exact mode, so it can serve as the fixed point the drop-in's concurrent Unbind is compared with
(tests/test_gpu_dropin.py::test_merge_subhalos_drop_in)."""
import numpy as np
import pytest

import cases
from hbtplus_b200 import capi
from oracle import pyoracle as po


@pytest.mark.parametrize("variant", ["v32", "v64"])
def test_reference_merge_is_thread_count_independent(variant):
    if not po.have_dropin(variant):
        pytest.skip("oracle/_ref libraries not built (reference sources absent at build time)")
    ref = po.load_ref_variant(variant)
    p = capi.make_params(box_size=62.5, softening=5e-3, periodic=True)
    e = capi.make_epoch(0.8, snapshot_index=23)
    snap = cases.case_merge()
    r1, m1 = po.merge_subhalos(ref, p, e, snap, mode=0, nthreads=1)
    r8, m8 = po.merge_subhalos(ref, p, e, snap, mode=0, nthreads=8)
    assert np.array_equal(m1, m8) and m1.sum() >= 8
    for f in r1.io.dtype.names:
        assert np.array_equal(r1.io[f], r8.io[f]), f
    n = int(r1.order_offset[-1])
    assert np.array_equal(r1.order_offset, r8.order_offset) and np.array_equal(r1.order[:n], r8.order[:n])
    # a trapped real subhalo died into its sink: one particle left, the sink's list grew and was re-unbound
    trapped = np.nonzero((r1.io["sink_track_id"] >= 0) & (snap.io["nbound"] > 1))[0]
    assert len(trapped) >= 8 and (r1.io["nbound"][trapped] == 1).all() and (r1.io["nsource"][trapped] == 1).all()
    assert (r1.io["snapshot_index_of_death"][trapped] == 23).all()
    # mode 1 (the patched sequence) only exists in the drop-in build
    with pytest.raises(RuntimeError):
        po.merge_subhalos(ref, p, e, snap, mode=1)
