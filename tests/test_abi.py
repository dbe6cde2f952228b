"""The C-ABI boundary without a GPU: the library loads, exports every symbol include/hbt_unbind.h declares,
the ctypes mirror matches the C structs, and there is NO fallback when no device is present."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest
import torch

from hbtplus_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "hbt_unbind.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hbtu_[a-z_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = capi.load_library()
    syms = declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/hbt_unbind.h but not exported"
    assert set(capi.EXPORTS) <= set(syms)
    assert lib.hbtu_abi_version() == 1


def test_ctypes_structs_match_the_header(tmp_path):
    prog = tmp_path / "sz.c"
    prog.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "hbt_unbind.h"\n'
        'int main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(hbtu_params), sizeof(hbtu_epoch), sizeof(hbtu_sub_io),'
        " sizeof(hbtu_stats), offsetof(hbtu_sub_io, mbound), offsetof(hbtu_sub_io, nsource_full), offsetof(hbtu_params, G),"
        " sizeof(hbtu_profile_io), offsetof(hbtu_profile_io, inertial_tensor));return 0;}\n"
    )
    exe = tmp_path / "sz"
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(capi.Params), C.sizeof(capi.Epoch), C.sizeof(capi.SubIO), C.sizeof(capi.Stats),
            capi.SubIO.mbound.offset, capi.SubIO.nsource_full.offset, capi.Params.G.offset,
            C.sizeof(capi.ProfileIO), capi.ProfileIO.inertial_tensor.offset]
    assert got == want
    assert capi.SUBIO_DTYPE.itemsize == C.sizeof(capi.SubIO)
    for name in capi.SUBIO_DTYPE.names:
        assert capi.SUBIO_DTYPE.fields[name][1] == getattr(capi.SubIO, name).offset, name


def test_order_capacity_matches_host_mirror():
    lib = capi.load_library()
    part_offset = np.array([0, 10, 15, 40, 41, 60], np.int64)
    nest_offset = np.array([0, 2, 3, 3, 3, 3], np.int64)
    nest_list = np.array([1, 2, 3], np.int32)  # 0 -> {1,2}, 1 -> {3}
    P = capi._ptr
    got = lib.hbtu_order_capacity(5, P(part_offset, C.c_int64), P(nest_offset, C.c_int64), P(nest_list, C.c_int32))
    assert got == capi.order_capacity(part_offset, nest_offset, nest_list) == (10 + 5 + 25 + 1) + (5 + 1) + 25 + 1 + 19
    bad = np.array([1, 1, 3], np.int32)  # 1 nested twice
    assert lib.hbtu_order_capacity(5, P(part_offset, C.c_int64), P(nest_offset, C.c_int64), P(bad, C.c_int32)) == capi.HBTU_ERR_INVALID
    cyc_off = np.array([0, 1, 2, 2, 2, 2], np.int64)
    cyc = np.array([1, 0], np.int32)
    assert lib.hbtu_order_capacity(5, P(part_offset, C.c_int64), P(cyc_off, C.c_int64), P(cyc, C.c_int32)) == capi.HBTU_ERR_INVALID


def test_order_capacity_both_layouts_and_dfs_layout():
    """hbtu_order_capacity has a fast path for batches whose parents precede their nested subhaloes (the shim's depth-first
    layout, synth.dfs_layout) and the full forest construction otherwise: both must give sum(own particles x (depth + 1))."""
    from hbtplus_b200 import synth

    lib = capi.load_library()
    P = capi._ptr
    rng = np.random.default_rng(17)
    for depth_first in (False, True):
        sizes = synth.subhalo_sizes(rng, 4000, 20, 5000)
        parent = synth.nest_forest(rng, sizes, max_depth=4, p_nest=0.5, root=None)
        if depth_first:
            old_sizes, old_parent = sizes, parent
            sizes, parent, order = synth.dfs_layout(sizes, parent, return_order=True)
            # a relabelling: same multiset of (size, parent size) pairs, every hierarchy contiguous with parents first
            assert np.array_equal(sizes, old_sizes[order])
            assert np.array_equal(np.where(parent >= 0, sizes[np.maximum(parent, 0)], -1), np.where(old_parent[order] >= 0, old_sizes[np.maximum(old_parent[order], 0)], -1))
            start = 0
            for s in range(len(sizes)):
                if parent[s] < 0:
                    start = s
                else:
                    assert start <= parent[s] < s
        depth = synth.forest_depth(parent)
        part_offset = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        kids = [[] for _ in sizes]
        for s, p_ in enumerate(parent):
            if p_ >= 0:
                kids[p_].append(s)
        nest_offset = np.concatenate([[0], np.cumsum([len(k) for k in kids])]).astype(np.int64)
        nest_list = np.array([c for k in kids for c in k], np.int32)
        want = int((sizes * (depth + 1)).sum())
        got = lib.hbtu_order_capacity(len(sizes), P(part_offset, C.c_int64), P(nest_offset, C.c_int64), P(nest_list, C.c_int32))
        assert got == want == capi.order_capacity(part_offset, nest_offset, nest_list)


def test_pipeline_planner():
    """capi.cu::plan_parts through the host-only diagnostic: a batch of many hierarchies laid out depth first is cut at a
    hierarchy boundary near a quarter of the particles; other layouts, small batches and dominant hierarchies are not cut."""
    from hbtplus_b200 import synth

    lib = capi.load_library()
    lib.hbtu_set_tuning.argtypes = [C.c_char_p, C.c_int64]
    lib.hbtu_get_tuning.argtypes = [C.c_char_p]
    lib.hbtu_get_tuning.restype = C.c_int64
    P = capi._ptr

    def plan(sizes, parent):
        part_offset = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        kids = [[] for _ in sizes]
        for s, p_ in enumerate(parent):
            if p_ >= 0:
                kids[p_].append(s)
        nest_offset = np.concatenate([[0], np.cumsum([len(k) for k in kids])]).astype(np.int64)
        nest_list = np.array([c for k in kids for c in k], np.int32)
        out = np.zeros(8, np.int64)
        n = lib.hbtu_plan_pipeline(len(sizes), P(part_offset, C.c_int64), P(nest_offset, C.c_int64), P(nest_list, C.c_int32), P(out, C.c_int64), 7)
        assert n >= 1 and out[0] == 0 and out[n] == len(sizes)
        return out[:n + 1], part_offset

    rng = np.random.default_rng(23)
    sizes = synth.subhalo_sizes(rng, 3000, 20, 4000)
    parent = synth.nest_forest(rng, sizes, max_depth=3, p_nest=0.3, root=None)
    dsizes, dparent = synth.dfs_layout(sizes, parent)
    old = lib.hbtu_get_tuning(b"pipeline_min_particles")
    try:
        assert len(plan(dsizes, dparent)[0]) == 2  # below the default threshold of 2^24 particles: one piece
        lib.hbtu_set_tuning(b"pipeline_min_particles", 1000)
        cuts, po = plan(dsizes, dparent)
        assert len(cuts) == 3
        assert dparent[cuts[1]] < 0  # a hierarchy starts at the cut ...
        n_all = po[-1]
        assert n_all / 4 <= po[cuts[1]] <= n_all / 4 + dsizes.max() * 4  # ... the first one at or after a quarter of the particles
        assert len(plan(sizes, parent)[0]) == 2  # children listed away from their parents: one piece
        big = np.concatenate([[int(dsizes.sum())], dsizes])  # one hierarchy holds half of the batch: the two-wave upload instead
        bigp = np.concatenate([[-1], np.where(dparent >= 0, dparent + 1, -1)])
        assert len(plan(big, bigp)[0]) == 2
        assert len(plan(dsizes[:40], np.full(40, -1))[0]) == 2  # fewer than 64 subhaloes
        lib.hbtu_set_tuning(b"pipeline_min_particles", 0)
        assert len(plan(dsizes, dparent)[0]) == 2  # switched off
    finally:
        lib.hbtu_set_tuning(b"pipeline_min_particles", old)


def test_tuning_knobs_round_trip():
    lib = capi.load_library()
    lib.hbtu_set_tuning.argtypes = [C.c_char_p, C.c_int64]
    lib.hbtu_get_tuning.argtypes = [C.c_char_p]
    lib.hbtu_get_tuning.restype = C.c_int64
    assert lib.hbtu_get_tuning(b"walk_group_min") == 256 and lib.hbtu_get_tuning(b"pipeline_min_particles") == 1 << 24
    for key in (b"walk_group_min", b"pipeline_min_particles", b"walk_small_max"):
        old = lib.hbtu_get_tuning(key)
        assert lib.hbtu_set_tuning(key, 12345) == 0 and lib.hbtu_get_tuning(key) == 12345
        assert lib.hbtu_set_tuning(key, old) == 0
    assert lib.hbtu_get_tuning(b"no_such_knob") == -1 and lib.hbtu_set_tuning(b"no_such_knob", 1) != 0


def test_create_rejects_bad_abi_and_has_no_cpu_fallback():
    lib = capi.load_library()
    p = capi.make_params(box_size=62.5, softening=5e-3)
    ctx = C.c_void_p()
    p.struct_size = 8
    assert lib.hbtu_create(C.byref(p), C.byref(ctx)) == capi.HBTU_ERR_INVALID
    p = capi.make_params(box_size=62.5, softening=5e-3)
    p.real_bytes = 8
    assert lib.hbtu_create(C.byref(p), C.byref(ctx)) == capi.HBTU_ERR_UNSUPPORTED
    if not torch.cuda.is_available():
        p = capi.make_params(box_size=62.5, softening=5e-3)
        rc = lib.hbtu_create(C.byref(p), C.byref(ctx))
        assert rc == capi.HBTU_ERR_NODEVICE and not ctx.value
        assert b"no CPU fallback" in lib.hbtu_last_error(None)
        from hbtplus_b200.unbind import UnbindContext, UnbindError

        with pytest.raises(UnbindError):
            UnbindContext(p)


def test_product_package_never_touches_the_oracle():
    """hbtplus_b200/ must not import, load or link anything under oracle/."""
    pkg = os.path.join(ROOT, "hbtplus_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in txt and "libhbtoracle" not in txt and "libhbtref" not in txt and "hbto_" not in txt, (dirpath, f)
    out = subprocess.check_output(["ldd", capi.LIB_PATH]).decode()
    assert "hbtoracle" not in out and "hbtref" not in out


def test_params_follow_the_reference_derivations():
    p = capi.make_params(box_size=62.5, softening=5e-3, open_angle=0.45)
    f = np.float32
    assert p.box_half == float(f(62.5) / f(2))
    assert p.tree_node_open_angle_square == float(f(0.45) * f(0.45))
    assert p.tree_node_resolution == float(f(float(f(5e-3)) * 0.1))
    assert p.G == float(f(43.0071))
    e = capi.make_epoch(1.0)
    assert e.hz == 100.0 and e.scale_factor == 1.0
