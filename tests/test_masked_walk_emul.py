"""CPU unit test of the masked group walk's logic (hbtplus_b200/csrc/walk_masked.cuh).

tests/host_emul/masked_emul.cpp compiles the SAME header with g++ behind a warp-emulation shim (32 fibers per warp,
collectives as rendez-vous that also verify that every lane reached the same collective) and walks groups of 128 targets
over the pre-order node array; a scalar per-target walk with the same fp32 arithmetic is the check: every target must accept
exactly the same number of nodes (the opener masks reproduce each target's own decisions) and get the same sum.
(Test scaffolding only - the product library has no host path; the GPU parity tests cover the compiled kernel.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from hbtplus_b200 import capi, synth

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module", params=[2, 1], ids=["groups128", "groups64"])
def memul(request, tmp_path_factory):
    """both group sizes of the masked walk: two slice pairs per warp (128 targets) and one (64 targets)"""
    out = tmp_path_factory.mktemp("memul") / f"libmaskedemul{request.param}.so"
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++", f"-DEMUL_NP={request.param}",
                           "-I", os.path.join(ROOT, "hbtplus_b200", "csrc"), "-I", os.path.join(ROOT, "include"),
                           os.path.join(HERE, "host_emul", "masked_emul.cpp"), "-o", str(out)])
    return C.CDLL(str(out))


def run(memul, p, pm, ntgt, stride=1):
    n = len(pm)
    ntgt = min(ntgt, n)
    P = capi._ptr
    sm, ss = np.zeros(ntgt), np.zeros(ntgt)
    am, asc = np.zeros(ntgt, np.int64), np.zeros(ntgt, np.int64)
    stats = np.zeros(4, np.int64)
    rc = memul.emul_masked_walk(C.byref(p), C.c_int64(n), P(pm, C.c_float), C.c_int64(ntgt), C.c_int64(stride), P(sm, C.c_double), P(ss, C.c_double),
                                P(am, C.c_int64), P(asc, C.c_int64), P(stats, C.c_int64))
    assert rc == 0
    return sm, ss, am, asc, stats


def per_lane(a, ntgt, memul):
    """sum over the (2 or 4) targets of every lane of every group"""
    g = memul.emul_group_size()
    pad = (-ntgt) % g
    return np.concatenate([a, np.zeros(pad, a.dtype)]).reshape(-1, g // 32, 32).sum(axis=1)


@pytest.mark.parametrize("periodic", [False, True])
@pytest.mark.parametrize("n", [1, 2, 33, 128, 129, 1000, 20000])
def test_masked_walk_reproduces_every_targets_decisions(memul, n, periodic):
    p = capi.make_params(box_size=62.5, softening=5e-3, periodic=periodic)
    snap = synth.make_snapshot([n], seed=n + 11, wrap=periodic, centre=[0.1, 31, 62.4] if periodic else None)
    pm = np.ascontiguousarray(snap.pos_mass)
    ntgt = min(n, 1536)
    sm, ss, am, asc, stats = run(memul, p, pm, ntgt)
    assert stats[0] == 0  # no stack overflow
    if periodic:  # shifted images round differently from NEAREST(dx): a decision may flip where r^2 straddles the criterion
        assert abs(int(am.sum()) - int(asc.sum())) <= 1e-4 * asc.sum() + 2
        assert np.mean(per_lane(am, ntgt, memul) == per_lane(asc, ntgt, memul)) > 0.98
    else:
        assert np.array_equal(per_lane(am, ntgt, memul), per_lane(asc, ntgt, memul))
    # periodic: the group's common image is more accurate than NEAREST(dx) of a wrapped pair (ulp(62) = 4e-6 against
    # pair distances of 1e-2); the same holds for the dense ring
    assert np.allclose(sm, ss, rtol=5e-5 if periodic else 2e-6, atol=0)


def test_masked_walk_softened_and_colocated(memul):
    """softening comparable to the inter-particle distance (many spline pairs, softened cells) and co-located particles"""
    p = capi.make_params(box_size=62.5, softening=0.05, periodic=False)
    snap = synth.make_snapshot([3000], seed=5, wrap=False)
    pm = np.ascontiguousarray(snap.pos_mass)
    pm[100:140, :3] = pm[100, :3]
    pm[200:203, :3] = pm[200, :3] + np.float32(1e-7)
    sm, ss, am, asc, stats = run(memul, p, pm, 3000)
    assert stats[0] == 0
    assert np.array_equal(per_lane(am, 3000, memul), per_lane(asc, 3000, memul))
    assert np.allclose(sm, ss, rtol=2e-6, atol=0)


def test_masked_walk_dense_core(memul):
    """a cuspy 2e5-particle halo with the bench's tiny softening: deep trees, long sibling chains, the stack under load"""
    p = capi.make_params(box_size=100.0, softening=4.8e-5, periodic=False)
    snap = synth.make_snapshot([200000], seed=3, wrap=False)
    pm = np.ascontiguousarray(snap.pos_mass)
    sm, ss, am, asc, stats = run(memul, p, pm, 200000, stride=97)  # 17 groups across the whole halo
    assert stats[0] == 0
    assert asc.sum() > 0 and np.array_equal(per_lane(am, 200000, memul), per_lane(asc, 200000, memul))
    assert np.allclose(sm, ss, rtol=2e-6, atol=0)


def test_masked_walk_clumpy_tree(memul):
    """a host with 40 dense clumps inside it walked as ONE tree (the source of a central that received its satellites'
    particles): strongly varying depth along the key order, groups that straddle clumps"""
    p = capi.make_params(box_size=100.0, softening=4.8e-5, periodic=False)
    sizes = [80000] + [1500] * 40
    snap = synth.make_snapshot(sizes, seed=17, wrap=False, parent=[-1] + [0] * 40, box_size=100.0, particle_mass=1e-6, centre=[50.0, 50.0, 50.0])
    pm = np.ascontiguousarray(snap.pos_mass)
    n = len(pm)
    sm, ss, am, asc, stats = run(memul, p, pm, n, stride=41)  # 27 groups across the key range
    assert stats[0] == 0
    assert asc.sum() > 0 and np.array_equal(per_lane(am, n, memul), per_lane(asc, n, memul))
    assert np.allclose(sm, ss, rtol=2e-6, atol=0)


def test_masked_walk_does_not_depend_on_lane_order(memul, monkeypatch):
    """The emulation runs the lanes of a warp one after the other between two rendez-vous (collectives, __syncwarp).  With a
    random lane order on every pass, a shared-memory hand-over that is not bracketed by a synchronisation would change the
    result: sums and counts must be bit-identical to the run in lane order."""
    p = capi.make_params(box_size=100.0, softening=4.8e-5, periodic=False)
    snap = synth.make_snapshot([60000], seed=8, wrap=False)
    pm = np.ascontiguousarray(snap.pos_mass)
    monkeypatch.delenv("EMUL_LANE_ORDER_SEED", raising=False)
    sm0, ss0, am0, asc0, st0 = run(memul, p, pm, 60000, stride=59)
    for seed in ("12345", "987654321"):
        monkeypatch.setenv("EMUL_LANE_ORDER_SEED", seed)
        sm1, ss1, am1, asc1, st1 = run(memul, p, pm, 60000, stride=59)
        assert st1[0] == 0 and np.array_equal(am0, am1) and np.array_equal(sm0, sm1)
    monkeypatch.delenv("EMUL_LANE_ORDER_SEED", raising=False)
