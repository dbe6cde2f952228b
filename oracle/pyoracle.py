"""ctypes loaders for the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; nothing under hbtplus_b200/ does (the product path has no CPU fallback).

  load_ref()     oracle/_ref/libhbtref_v32.so  - the unmodified reference sources (hbtref_*)
  load_oracle()  oracle/libhbtoracle.so        - the plain-C restatement (hbto_*)
Both export ``*_unbind_batch`` / ``*_tree_potential`` with the argument list of the product's
hbtu_* entry points (include/hbt_unbind.h), the context pointer replaced by ``const hbtu_params*``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from hbtplus_b200 import capi

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_PATH = os.path.join(_HERE, "_ref", "libhbtref_v32.so")
ORACLE_PATH = os.path.join(_HERE, "libhbtoracle.so")


def _bind(lib, prefix):
    ub = getattr(lib, prefix + "_unbind_batch")
    ub.argtypes = [C.POINTER(capi.Params)] + capi.BATCH_ARGTYPES
    ub.restype = C.c_int
    tp = getattr(lib, prefix + "_tree_potential")
    tp.argtypes = [C.POINTER(capi.Params)] + capi.POTENTIAL_ARGTYPES
    tp.restype = C.c_int
    pb = getattr(lib, prefix + "_profile_batch", None)
    if pb is not None:
        pb.argtypes = [C.POINTER(capi.Params)] + capi.PROFILE_ARGTYPES
        pb.restype = C.c_int
    mb = getattr(lib, prefix + "_mask_batch", None)
    if mb is not None:
        mb.argtypes = [C.POINTER(capi.Params)] + capi.MASK_ARGTYPES
        mb.restype = C.c_int
    getattr(lib, prefix + "_set_num_threads").argtypes = [C.c_int]
    getattr(lib, prefix + "_get_max_threads").restype = C.c_int
    return lib


def have_ref() -> bool:
    return os.path.exists(REF_PATH)


def load_ref():
    lib = _bind(C.CDLL(REF_PATH), "hbtref")
    lib.hbtref_seed.argtypes = [C.c_uint]
    return lib


def load_oracle():
    lib = _bind(C.CDLL(ORACLE_PATH), "hbto")
    lib.hbto_walk_counts.argtypes = [C.POINTER(capi.Params), C.POINTER(capi.Epoch), C.c_int64, C.POINTER(C.c_float),
                                     C.c_int64, C.POINTER(C.c_float), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.hbto_walk_counts.restype = C.c_int
    lib.hbto_last_interactions.restype = C.c_int64
    lib.hbto_set_shuffle_mode.argtypes = [C.c_int]
    lib.hbto_seed.argtypes = [C.c_uint]
    return lib


class Result:
    def __init__(self, io, order_offset, order, energy):
        self.io, self.order_offset, self.order, self.energy = io, order_offset, order, energy

    def particles(self, s):
        b = self.order_offset[s]
        return self.order[b : b + self.io["nsource"][s]]

    def bound(self, s):
        b = self.order_offset[s]
        return self.order[b : b + self.io["nbound"][s]]


def run_batch(lib, prefix, params, epoch, snap, flags=0, want_energy=True):
    """Call ``<prefix>_unbind_batch`` on a hbtplus_b200.synth.Snapshot; returns a Result."""
    cap = capi.order_capacity(snap.part_offset, snap.nest_offset, snap.nest_list)
    io = snap.io.copy()
    order_offset = np.zeros(snap.nsub + 1, np.int64)
    order = np.full(max(cap, 1), -1, np.int32)
    energy = np.zeros(max(cap, 1), np.float32) if want_energy else None
    pm = np.ascontiguousarray(snap.pos_mass, np.float32)
    vv = np.ascontiguousarray(snap.vel, np.float32)
    rc = getattr(lib, prefix + "_unbind_batch")(
        C.byref(params), *capi.batch_args(epoch, snap.part_offset, pm, vv, snap.nest_offset, snap.nest_list, io, flags, cap, order_offset, order, energy)
    )
    if rc != 0:
        raise RuntimeError(f"{prefix}_unbind_batch failed: {rc}")
    return Result(io, order_offset, order, energy)


def tree_potential(lib, prefix, params, epoch, src_pos_mass, tgt_pos, self_mass=None, tgt_vel=None, ref_pos=None, ref_vel=None):
    src = np.ascontiguousarray(src_pos_mass, np.float32)
    tgt = np.ascontiguousarray(tgt_pos, np.float32)
    out = np.zeros(len(tgt), np.float64)
    sm = None if self_mass is None else np.ascontiguousarray(self_mass, np.float32)
    tv = None if tgt_vel is None else np.ascontiguousarray(tgt_vel, np.float32)
    rp = None if ref_pos is None else np.ascontiguousarray(ref_pos, np.float64)
    rv = None if ref_vel is None else np.ascontiguousarray(ref_vel, np.float64)
    P = capi._ptr
    rc = getattr(lib, prefix + "_tree_potential")(
        C.byref(params), C.byref(epoch), len(src), P(src, C.c_float), len(tgt), P(tgt, C.c_float), P(sm, C.c_float), P(tv, C.c_float),
        P(rp, C.c_double), P(rv, C.c_double), P(out, C.c_double))
    if rc != 0:
        raise RuntimeError(f"{prefix}_tree_potential failed: {rc}")
    return out


DROPIN_PATH = os.path.join(_HERE, "_ref", "libhbtdropin_v32.so")


def variant_paths(variant: str = "v32"):
    return os.path.join(_HERE, "_ref", f"libhbtref_{variant}.so"), os.path.join(_HERE, "_ref", f"libhbtdropin_{variant}.so")


def have_dropin(variant: str = "v32") -> bool:
    return all(os.path.exists(p) for p in variant_paths(variant))


def load_dropin(variant: str = "v32"):
    """The reference harness linked against integration/subhalo_unbind_b200.o + libhbtunbind.so (GPU backend)."""
    return _bind(C.CDLL(variant_paths(variant)[1]), "hbtref")


def load_ref_variant(variant: str):
    lib = _bind(C.CDLL(variant_paths(variant)[0]), "hbtref")
    lib.hbtref_seed.argtypes = [C.c_uint]
    return lib


def refine_particles(lib, params, epoch, snap, host_halo_id, n_old, nhalos, mbound_in):
    """``SubhaloSnapshot_t::RefineParticles()`` through the harness in ref_build/ref_capi.cpp."""
    f = lib.hbtref_refine_particles
    f.restype = C.c_int
    f.argtypes = [C.POINTER(capi.Params), C.POINTER(capi.Epoch), C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_float), C.POINTER(C.c_float),
                  C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int64, C.c_int32, C.POINTER(C.c_float),
                  C.POINTER(capi.SubIO), C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_float)]
    cap = capi.order_capacity(snap.part_offset, snap.nest_offset, snap.nest_list) + snap.npart
    io = snap.io.copy()
    order_offset = np.zeros(snap.nsub + 1, np.int64)
    order = np.full(max(cap, 1), -1, np.int32)
    energy = np.zeros(max(cap, 1), np.float32)
    pm = np.ascontiguousarray(snap.pos_mass, np.float32)
    vv = np.ascontiguousarray(snap.vel, np.float32)
    hh = np.ascontiguousarray(host_halo_id, np.int32)
    mb = np.ascontiguousarray(mbound_in, np.float32)
    P = capi._ptr
    rc = f(C.byref(params), C.byref(epoch), snap.nsub, P(snap.part_offset, C.c_int64), P(pm, C.c_float), P(vv, C.c_float),
           P(snap.nest_offset, C.c_int64), P(snap.nest_list, C.c_int32), P(hh, C.c_int32), n_old, nhalos, P(mb, C.c_float),
           io.ctypes.data_as(C.POINTER(capi.SubIO)), cap, P(order_offset, C.c_int64), P(order, C.c_int32), P(energy, C.c_float))
    if rc != 0:
        raise RuntimeError(f"hbtref_refine_particles failed: {rc}")
    r = Result(io, order_offset, order, energy)
    lib.hbtref_last_refine_seconds.restype = C.c_double
    r.refine_seconds = lib.hbtref_last_refine_seconds()  # SubhaloSnapshot_t::RefineParticles() alone
    return r


def profile_batch(lib, prefix, params, epoch, part_offset, pos_mass, io):
    """CalculateProfileProperties + CalculateShape on the CPU checker `lib` (same contract as hbtu_profile_batch)."""
    out = np.ascontiguousarray(io, capi.PROFILEIO_DTYPE).copy()
    po = np.ascontiguousarray(part_offset, np.int64)
    pm = np.ascontiguousarray(pos_mass, np.float32)
    rc = getattr(lib, prefix + "_profile_batch")(C.byref(params), C.byref(epoch), len(po) - 1, capi._ptr(po, C.c_int64), capi._ptr(pm, C.c_float),
                                                 out.ctypes.data_as(C.POINTER(capi.ProfileIO)))
    if rc != 0:
        raise RuntimeError(f"{prefix}_profile_batch failed: {rc}")
    return out


def mask_batch(lib, prefix, params, part_offset, particle_id, nest_offset, nest_list, nbound):
    """SubhaloSnapshot_t::MaskSubhalos on the CPU checker `lib` (same contract as hbtu_mask_batch)."""
    po = np.ascontiguousarray(part_offset, np.int64)
    ids = np.ascontiguousarray(particle_id, np.int64)
    nb = np.ascontiguousarray(nbound, np.int64)
    no = None if nest_offset is None else np.ascontiguousarray(nest_offset, np.int64)
    nl = None if nest_list is None else np.ascontiguousarray(nest_list, np.int32)
    new_count = np.zeros(len(po) - 1, np.int64)
    keep = np.full(max(int(po[-1]), 1), -1, np.int32)
    rc = getattr(lib, prefix + "_mask_batch")(C.byref(params), *capi.mask_args(po, ids, no, nl, nb, new_count, keep))
    if rc != 0:
        raise RuntimeError(f"{prefix}_mask_batch failed: {rc}")
    return new_count, keep


def idtable_query(lib, prefix, params, particle_id, query_id):
    """MappedIndexTable_t::Fill + GetIndices on the CPU checker `lib` (same contract as hbtu_idtable_build + _query)."""
    ids = np.ascontiguousarray(particle_id, np.int64)
    q = np.ascontiguousarray(query_id, np.int64)
    out = np.empty(len(q), np.int64)
    f = getattr(lib, prefix + "_idtable_query")
    f.argtypes = [C.POINTER(capi.Params), C.c_int64, C.POINTER(C.c_int64), C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    f.restype = C.c_int
    rc = f(C.byref(params), len(ids), capi._ptr(ids, C.c_int64), len(q), capi._ptr(q, C.c_int64), capi._ptr(out, C.c_int64))
    if rc != 0:
        raise RuntimeError(f"{prefix}_idtable_query failed: {rc}")
    return out


def detect_traps(lib, prefix, params, epoch, part_offset, pos_mass, vel, nest_offset, nest_list, io):
    """Detection part of SubhaloSnapshot_t::MergeSubhalos on the CPU checker `lib` (same contract as hbtu_detect_traps)."""
    out = np.ascontiguousarray(io, capi.TRAPIO_DTYPE).copy()
    po = np.ascontiguousarray(part_offset, np.int64)
    pm = np.ascontiguousarray(pos_mass, np.float32)
    vv = np.ascontiguousarray(vel, np.float32)
    no = None if nest_offset is None else np.ascontiguousarray(nest_offset, np.int64)
    nl = None if nest_list is None else np.ascontiguousarray(nest_list, np.int32)
    f = getattr(lib, prefix + "_detect_traps")
    f.argtypes = [C.POINTER(capi.Params)] + capi.TRAP_ARGTYPES
    f.restype = C.c_int
    rc = f(C.byref(params), *capi.trap_args(epoch, po, pm, vv, no, nl, out))
    if rc != 0:
        raise RuntimeError(f"{prefix}_detect_traps failed: {rc}")
    return out


def merge_subhalos(lib, params, epoch, snap, mode=0, nthreads=0):
    """``SubhaloSnapshot_t::MergeSubhalos()`` with MergeTrappedSubhalos on, through ref_build/ref_capi.cpp::hbtref_merge_subhalos.
    mode 0 = the unmodified member function (its Unbind calls come from `nthreads` OpenMP threads); mode 1 = the patched sequence
    with HBT_B200_UnbindMerged (drop-in libraries only).  Returns (Result, is_merged)."""
    f = lib.hbtref_merge_subhalos
    f.restype = C.c_int
    f.argtypes = [C.POINTER(capi.Params), C.POINTER(capi.Epoch), C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_float), C.POINTER(C.c_float),
                  C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(capi.SubIO), C.c_int32, C.c_int32, C.c_int64, C.POINTER(C.c_int64),
                  C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    cap = 2 * snap.npart + 16  # MergeTo can append a whole satellite list to its sink
    io = snap.io.copy()
    order_offset = np.zeros(snap.nsub + 1, np.int64)
    order = np.full(cap, -1, np.int32)
    merged = np.zeros(snap.nsub, np.int32)
    pm = np.ascontiguousarray(snap.pos_mass, np.float32)
    vv = np.ascontiguousarray(snap.vel, np.float32)
    P = capi._ptr
    rc = f(C.byref(params), C.byref(epoch), snap.nsub, P(snap.part_offset, C.c_int64), P(pm, C.c_float), P(vv, C.c_float),
           P(snap.nest_offset, C.c_int64), P(snap.nest_list, C.c_int32), io.ctypes.data_as(C.POINTER(capi.SubIO)), mode, nthreads, cap,
           P(order_offset, C.c_int64), P(order, C.c_int32), P(merged, C.c_int32))
    if rc != 0:
        raise RuntimeError(f"hbtref_merge_subhalos failed: {rc}")
    return Result(io, order_offset, order, None), merged
