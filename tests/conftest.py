import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle_lib():
    """The plain-C restatement (oracle/libhbtoracle.so); built on demand with gcc."""
    from oracle import pyoracle as po

    if not os.path.exists(po.ORACLE_PATH):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    lib = po.load_oracle()
    lib.hbto_set_num_threads(min(8, os.cpu_count() or 1))
    return lib


@pytest.fixture(scope="session")
def ref_lib():
    """The unmodified reference (oracle/_ref); only present where /root/reference was available at build time."""
    from oracle import pyoracle as po

    if not po.have_ref():
        pytest.skip("oracle/_ref/libhbtref_v32.so not built (reference sources absent)")
    lib = po.load_ref()
    lib.hbtref_set_num_threads(1)
    return lib


def load_golden(name):
    from hbtplus_b200 import synth

    z = np.load(os.path.join(HERE, "golden", f"{name}.npz"))
    has_nest = bool(z["has_nest"])
    snap = synth.Snapshot(z["part_offset"], z["pos_mass"], z["vel"], z["nest_offset"] if has_nest else None,
                          z["nest_list"] if has_nest else None, z["io_in"])
    return snap, z


def orders_equal_modulo_ties(order_a, order_b, energy_b, nbound):
    """Same particle sequence, except that runs of (nearly) equal binding energy may be permuted.

    The reference's std::sort is unstable, so ties have no defined order (src/subhalo_unbind.cpp:382,405)."""
    order_a, order_b = np.asarray(order_a), np.asarray(order_b)
    if len(order_a) != len(order_b):
        return False
    if np.array_equal(order_a, order_b):
        return True
    if sorted(order_a[:nbound].tolist()) != sorted(order_b[:nbound].tolist()):
        return False
    bad = np.nonzero(order_a != order_b)[0]
    pos_b = {int(p): i for i, p in enumerate(order_b)}
    for i in bad:
        j = pos_b.get(int(order_a[i]))
        if j is None or abs(j - i) > 8:
            return False
        if i < nbound and energy_b is not None:
            ea, eb = energy_b[i], energy_b[j]
            if abs(ea - eb) > 2e-5 * max(abs(ea), abs(eb), 1e-30):
                return False
    return True
