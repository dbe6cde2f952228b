"""CPU unit test of the device octree's index arithmetic (hbtplus_b200/csrc/tree_core.cuh).

tests/host_emul/emul.cpp compiles the SAME header with g++ and emulates the kernels' per-element logic plus a
scalar fp32 walk over the pre-order node array; it is checked against the oracle's scalar reference walk:
identical accepted-interaction counts per target and potentials to fp32 round-off.  (Test scaffolding only -
the product library has no host path.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from hbtplus_b200 import capi, synth
from oracle import pyoracle as po

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    out = tmp_path_factory.mktemp("emul") / "libemul.so"
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", "-I", os.path.join(ROOT, "hbtplus_b200", "csrc"),
                           "-I", os.path.join(ROOT, "include"), os.path.join(HERE, "host_emul", "emul.cpp"), "-o", str(out)])
    return C.CDLL(str(out))


@pytest.mark.parametrize("periodic", [False, True])
@pytest.mark.parametrize("n", [2, 3, 20, 100, 1000, 30000])
def test_preorder_tree_reproduces_reference_walk(oracle_lib, emul, n, periodic):
    p = capi.make_params(box_size=62.5, softening=5e-3, periodic=periodic)
    e = capi.make_epoch(1.0)
    snap = synth.make_snapshot([n], seed=n + 3, wrap=periodic, centre=[0.1, 31, 62.4] if periodic else None)
    pm = np.ascontiguousarray(snap.pos_mass)
    ntg = min(n, 3000)
    tg = pm[:: max(1, n // ntg)][:ntg].copy()
    sm = tg[:, 3].copy()
    P = capi._ptr
    want = po.tree_potential(oracle_lib, "hbto", p, e, pm, tg, self_mass=sm)
    acc_o = np.zeros(len(tg), np.int64)
    ncell_ref = oracle_lib.hbto_walk_counts(C.byref(p), C.byref(e), n, P(pm, C.c_float), len(tg), P(tg, C.c_float), P(acc_o, C.c_int64), None)
    out = np.zeros(len(tg))
    acc = np.zeros(len(tg), np.int64)
    ncell = C.c_int64(0)
    rc = emul.emul_tree_potential(C.byref(p), C.byref(e), C.c_int64(n), P(pm, C.c_float), C.c_int64(len(tg)), P(tg, C.c_float), P(sm, C.c_float),
                                  P(out, C.c_double), P(acc, C.c_int64), None, C.byref(ncell))
    assert rc == 0  # every pre-order slot written exactly once
    assert ncell.value <= ncell_ref  # single-child chains are collapsed, nothing else
    assert np.mean(acc == acc_o) > 0.99  # decisions differ only where fp32 r^2 straddles the criterion
    assert abs(acc.sum() - acc_o.sum()) <= 1e-5 * acc_o.sum() + 2
    rel = np.abs(out - want) / np.abs(want)
    assert rel.max() < 1e-4


def test_colocated_particles_form_buckets(oracle_lib, emul):
    """Identical positions (the reference randomises octants below TreeNodeResolution, oct_tree.tpp:112-123)."""
    p = capi.make_params(box_size=62.5, softening=5e-3, periodic=False)
    e = capi.make_epoch(1.0)
    snap = synth.make_snapshot([400], seed=9, wrap=False)
    pm = np.ascontiguousarray(snap.pos_mass)
    pm[100:140, :3] = pm[100, :3]  # 40 co-located particles
    pm[200:203, :3] = pm[200, :3] + np.float32(1e-7)
    tg = pm[::7].copy()
    P = capi._ptr
    want = po.tree_potential(oracle_lib, "hbto", p, e, pm, tg)
    out = np.zeros(len(tg))
    rc = emul.emul_tree_potential(C.byref(p), C.byref(e), C.c_int64(len(pm)), P(pm, C.c_float), C.c_int64(len(tg)), P(tg, C.c_float), None,
                                  P(out, C.c_double), None, None, None)
    assert rc == 0
    assert (np.abs(out - want) / np.abs(want)).max() < 1e-3
