"""Host-side mirror of the reference's unbinding interface on top of the C-ABI.

``UnbindContext`` plays the role of ``SubhaloSnapshot_t::RefineParticles`` /
``Subhalo_t::Unbind`` (reference: src/subhalo_unbind.cpp:263-516) and of
``GravityTree_t::EvaluatePotential/BindingEnergy`` (src/gravity_tree.cpp:79-175) for callers
written in Python (tests, bench.py).  It only marshals arrays into ``hbtu_*`` calls of
``libhbtunbind.so``; there is no Python or CPU implementation of the path behind it.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


class UnbindError(RuntimeError):
    """The reference path cannot fail; the C-ABI reports errors as negative codes, which become this."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"hbtu error {code}: {msg}")
        self.code = code


class BatchResult:
    """Outputs of one batch: per-subhalo records + the new ``Subhalo_t::Particles`` orders."""

    def __init__(self, io, order_offset, order, energy):
        self.io, self.order_offset, self.order, self.energy = io, order_offset, order, energy

    def particles(self, s: int) -> np.ndarray:
        b = self.order_offset[s]
        return self.order[b : b + self.io["nsource"][s]]

    def bound(self, s: int) -> np.ndarray:
        b = self.order_offset[s]
        return self.order[b : b + self.io["nbound"][s]]


class UnbindContext:
    def __init__(self, params: capi.Params, lib_path: str | None = None):
        self._lib = capi.load_library(lib_path)
        self._lib.hbtu_set_counting.argtypes = [C.c_void_p, C.c_int]
        self._ctx = C.c_void_p()
        self.params = params
        rc = self._lib.hbtu_create(C.byref(params), C.byref(self._ctx))
        if rc != capi.HBTU_OK:
            raise UnbindError(rc, (self._lib.hbtu_last_error(None) or b"").decode())

    def close(self):
        if self._ctx:
            self._lib.hbtu_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != capi.HBTU_OK:
            raise UnbindError(rc, (self._lib.hbtu_last_error(self._ctx) or b"").decode())

    # -- target split of the walk over cooperating contexts (the non-natural case) ------------
    def set_walk_split(self, rank: int, nranks: int, allreduce=None):
        """Deal every round's walk targets over `nranks` contexts that execute the SAME batch in lock step.
        ``allreduce(device_ptr: int, count: int, stream: int)`` must sum the `count` floats at `device_ptr` over all
        contexts in place and be complete when it returns (e.g. torch.distributed.all_reduce on a tensor wrapped around the
        pointer, see hbtplus_b200.sched.torch_allreduce).  nranks <= 1 or allreduce None switches the split off."""
        self._lib.hbtu_set_walk_split.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        if nranks <= 1 or allreduce is None:
            self._split_cb = None
            self._check(self._lib.hbtu_set_walk_split(self._ctx, 0, 1, None, None))
            return

        def thunk(user, buf, count, stream):
            try:
                allreduce(int(buf), int(count), int(stream or 0))
                return 0
            except Exception as ex:  # reported as HBTU_ERR_CUDA by the library
                print(f"walk-split all-reduce failed: {ex!r}")
                return 1

        self._split_cb = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p)(thunk)  # keep alive
        self._check(self._lib.hbtu_set_walk_split(self._ctx, rank, nranks, C.cast(self._split_cb, C.c_void_p), None))

    def set_counting(self, on: bool):
        self._check(self._lib.hbtu_set_counting(self._ctx, int(on)))

    def stats(self) -> capi.Stats:
        st = capi.Stats()
        self._check(self._lib.hbtu_get_stats(self._ctx, C.byref(st)))
        return st

    # -- RefineParticles / RecursiveUnbind / Unbind / TruncateSource ---------------------------------
    def unbind_batch(self, epoch, snap, flags: int = 0, want_energy: bool = True, order_buf=None, energy_buf=None) -> BatchResult:
        """One RefineParticles-equivalent call.  ``order_buf`` / ``energy_buf`` may be caller-owned (e.g. pinned) int32 /
        float32 arrays of at least ``capi.order_capacity(...)`` entries; otherwise they are allocated here."""
        cap = capi.order_capacity(snap.part_offset, snap.nest_offset, snap.nest_list)
        io = snap.io.copy()
        order_offset = np.zeros(snap.nsub + 1, np.int64)
        order = order_buf if order_buf is not None else np.empty(max(cap, 1), np.int32)
        energy = (energy_buf if energy_buf is not None else np.empty(max(cap, 1), np.float32)) if want_energy else None
        assert order.dtype == np.int32 and len(order) >= cap
        pm = np.ascontiguousarray(snap.pos_mass, np.float32)
        vv = np.ascontiguousarray(snap.vel, np.float32)
        rc = self._lib.hbtu_unbind_batch(
            self._ctx, *capi.batch_args(epoch, snap.part_offset, pm, vv, snap.nest_offset, snap.nest_list, io, flags, cap, order_offset, order, energy)
        )
        self._check(rc)
        return BatchResult(io, order_offset, order, energy)

    def stage(self, epoch, snap, flags: int = 0):
        pm = np.ascontiguousarray(snap.pos_mass, np.float32)
        vv = np.ascontiguousarray(snap.vel, np.float32)
        args = capi.batch_args(epoch, snap.part_offset, pm, vv, snap.nest_offset, snap.nest_list, snap.io, flags, 0, None, None, None)
        self._check(self._lib.hbtu_stage(self._ctx, *args[:9]))
        self._staged = snap

    def execute(self):
        self._check(self._lib.hbtu_execute(self._ctx))

    def fetch(self, want_energy: bool = True) -> BatchResult:
        snap = self._staged
        cap = capi.order_capacity(snap.part_offset, snap.nest_offset, snap.nest_list)
        io = snap.io.copy()
        order_offset = np.zeros(snap.nsub + 1, np.int64)
        order = np.empty(max(cap, 1), np.int32)
        energy = np.empty(max(cap, 1), np.float32) if want_energy else None
        P = capi._ptr
        rc = self._lib.hbtu_fetch(self._ctx, io.ctypes.data_as(C.POINTER(capi.SubIO)), cap, P(order_offset, C.c_int64), P(order, C.c_int32), P(energy, C.c_float))
        self._check(rc)
        return BatchResult(io, order_offset, order, energy)

    # -- MappedIndexTable_t::Fill / GetIndices (src/hash.tpp:18-32, src/hash_remote.tpp:9-88) ----------------
    def idtable_build(self, particle_id):
        ids = np.ascontiguousarray(particle_id, np.int64)
        self._check(self._lib.hbtu_idtable_build(self._ctx, len(ids), capi._ptr(ids, C.c_int64)))

    def idtable_query(self, query_id) -> np.ndarray:
        """Index of every queried Id in the array given to idtable_build, -1 (NullParticleId) when absent."""
        q = np.ascontiguousarray(query_id, np.int64)
        out = np.empty(len(q), np.int64)
        self._check(self._lib.hbtu_idtable_query(self._ctx, len(q), capi._ptr(q, C.c_int64), capi._ptr(out, C.c_int64)))
        return out

    # -- detection part of SubhaloSnapshot_t::MergeSubhalos (src/subhalo_merge.cpp:29-199) -----------------------
    def detect_traps(self, epoch, part_offset, pos_mass, vel, nest_offset, nest_list, io) -> np.ndarray:
        """``io``: structured array (capi.TRAPIO_DTYPE); returns the updated copy (sink ids, sink snapshot, is_merged)."""
        out = np.ascontiguousarray(io, capi.TRAPIO_DTYPE).copy()
        po = np.ascontiguousarray(part_offset, np.int64)
        pm = np.ascontiguousarray(pos_mass, np.float32)
        vv = np.ascontiguousarray(vel, np.float32)
        no = None if nest_offset is None else np.ascontiguousarray(nest_offset, np.int64)
        nl = None if nest_list is None else np.ascontiguousarray(nest_list, np.int32)
        self._check(self._lib.hbtu_detect_traps(self._ctx, *capi.trap_args(epoch, po, pm, vv, no, nl, out)))
        return out

    # -- SubhaloSnapshot_t::MaskSubhalos (src/subhalo_tracking.cpp:793-841) -----------------------------
    def mask_batch(self, part_offset, particle_id, nest_offset, nest_list, nbound):
        """Exclusive particle ownership inside every hierarchy of the nest forest.  Returns (new_count[nsub], keep_index[N]):
        subhalo s keeps the entries keep_index[part_offset[s] : part_offset[s] + new_count[s]]."""
        po = np.ascontiguousarray(part_offset, np.int64)
        ids = np.ascontiguousarray(particle_id, np.int64)
        nb = np.ascontiguousarray(nbound, np.int64)
        no = None if nest_offset is None else np.ascontiguousarray(nest_offset, np.int64)
        nl = None if nest_list is None else np.ascontiguousarray(nest_list, np.int32)
        new_count = np.zeros(len(po) - 1, np.int64)
        keep = np.full(max(int(po[-1]), 1), -1, np.int32)
        self._check(self._lib.hbtu_mask_batch(self._ctx, *capi.mask_args(po, ids, no, nl, nb, new_count, keep)))
        return new_count, keep

    # -- Subhalo_t::CalculateProfileProperties + CalculateShape (src/subhalo.cpp:242-398) --------------
    def profile_batch(self, epoch, part_offset, pos_mass, io) -> np.ndarray:
        """``io``: structured array (capi.PROFILEIO_DTYPE) with the [in]/[io] fields set; returns the updated copy."""
        out = np.ascontiguousarray(io, capi.PROFILEIO_DTYPE).copy()
        po = np.ascontiguousarray(part_offset, np.int64)
        pm = np.ascontiguousarray(pos_mass, np.float32)
        P = capi._ptr
        rc = self._lib.hbtu_profile_batch(self._ctx, C.byref(epoch), len(po) - 1, P(po, C.c_int64), P(pm, C.c_float),
                                          out.ctypes.data_as(C.POINTER(capi.ProfileIO)))
        self._check(rc)
        return out

    def profile_executed(self, io) -> np.ndarray:
        """profile_batch on the batch hbtu_execute left resident in HBM (stage + execute first): no particle upload."""
        out = np.ascontiguousarray(io, capi.PROFILEIO_DTYPE).copy()
        self._check(self._lib.hbtu_profile_executed(self._ctx, out.ctypes.data_as(C.POINTER(capi.ProfileIO))))
        return out

    # -- GravityTree_t::Build + EvaluatePotential / BindingEnergy -------------------------------------
    def tree_potential(self, epoch, src_pos_mass, tgt_pos, self_mass=None, tgt_vel=None, ref_pos=None, ref_vel=None) -> np.ndarray:
        src = np.ascontiguousarray(src_pos_mass, np.float32)
        tgt = np.ascontiguousarray(tgt_pos, np.float32)
        out = np.zeros(len(tgt), np.float64)
        sm = None if self_mass is None else np.ascontiguousarray(self_mass, np.float32)
        tv = None if tgt_vel is None else np.ascontiguousarray(tgt_vel, np.float32)
        rp = None if ref_pos is None else np.ascontiguousarray(ref_pos, np.float64)
        rv = None if ref_vel is None else np.ascontiguousarray(ref_vel, np.float64)
        P = capi._ptr
        rc = self._lib.hbtu_tree_potential(
            self._ctx, C.byref(epoch), len(src), P(src, C.c_float), len(tgt), P(tgt, C.c_float), P(sm, C.c_float), P(tv, C.c_float),
            P(rp, C.c_double), P(rv, C.c_double), P(out, C.c_double))
        self._check(rc)
        return out


class SplitGroup:
    """Contexts of ONE process that cooperate on the same batch through the library's built-in peer-memory all-reduce
    (hbtu_split_group_*): one host thread per context must call execute()/unbind_batch() concurrently."""

    def __init__(self, contexts):
        self._lib = contexts[0]._lib
        self._lib.hbtu_split_group_create.restype = C.c_void_p
        self._lib.hbtu_split_group_create.argtypes = [C.c_int]
        self._lib.hbtu_split_group_join.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        self._lib.hbtu_split_group_destroy.argtypes = [C.c_void_p]
        self._g = C.c_void_p(self._lib.hbtu_split_group_create(len(contexts)))
        if not self._g:
            raise UnbindError(capi.HBTU_ERR_INVALID, "hbtu_split_group_create failed")
        self.contexts = list(contexts)
        for r, ctx in enumerate(self.contexts):
            ctx._check(self._lib.hbtu_split_group_join(self._g, ctx._ctx, r))

    def close(self):
        if self._g:
            self._lib.hbtu_split_group_destroy(self._g)
            self._g = C.c_void_p()

    def run(self, fn):
        """fn(rank, ctx) on one thread per member; returns the list of results (exceptions are re-raised)."""
        import threading

        out, err = [None] * len(self.contexts), [None] * len(self.contexts)

        def work(r):
            try:
                out[r] = fn(r, self.contexts[r])
            except BaseException as ex:  # noqa: BLE001
                err[r] = ex

        th = [threading.Thread(target=work, args=(r,)) for r in range(len(self.contexts))]
        for t in th:
            t.start()
        for t in th:
            t.join()
        for e in err:
            if e is not None:
                raise e
        return out
