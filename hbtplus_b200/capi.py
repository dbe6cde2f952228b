"""ctypes mirror of include/hbt_unbind.h (the C-ABI drop-in boundary).

The structs here are the POD copies of what ``Subhalo_t::Unbind`` reads and writes
(reference: src/subhalo.h:24-146, src/config_parser.h:22-123, src/snapshot.h:17-39).
This module only *describes* the ABI and loads the product library
``hbtplus_b200/csrc/libhbtunbind.so``; it contains no algorithm and no CPU fallback:
if the CUDA library is missing, loading raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HBTU_OK = 0
HBTU_ERR_INVALID = -1
HBTU_ERR_CUDA = -2
HBTU_ERR_NOMEM = -3
HBTU_ERR_NODEVICE = -4
HBTU_ERR_UNSUPPORTED = -5
HBTU_ERR_CAPACITY = -6
HBTU_FLAG_TRUNCATE_SOURCE = 1
HBTU_SUB_PLAIN_UNBIND = 1  # hbtu_sub_io.flags: entered through plain Subhalo_t::Unbind (no orphan rule)
HBTU_FLAG_NO_STRIPPING = 2  # the reference's -DNO_STRIPPING build
HBTU_FLAG_THERMAL_ENERGY = 4  # -DUNBIND_WITH_THERMAL_ENERGY: vel[:, 3] is Particle_t::InternalEnergy

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libhbtunbind.so")


class Params(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32),
        ("real_bytes", C.c_int32),
        ("min_num_part_of_sub", C.c_int32),
        ("periodic_boundary_on", C.c_int32),
        ("refine_mostbound_particle", C.c_int32),
        ("device", C.c_int32),
        ("max_sample_size", C.c_int64),
        ("bound_mass_precision", C.c_double),
        ("source_sub_relax_factor", C.c_double),
        ("box_size", C.c_double),
        ("box_half", C.c_double),
        ("softening_halo", C.c_double),
        ("tree_node_open_angle_square", C.c_double),
        ("tree_node_resolution", C.c_double),
        ("tree_node_resolution_half", C.c_double),
        ("tree_alloc_factor", C.c_double),
        ("tree_min_num_of_cells", C.c_int64),
        ("G", C.c_double),
        ("direct_sum_max", C.c_int64),
        ("shuffle_seed", C.c_int64),
    ]


class Epoch(C.Structure):
    _fields_ = [
        ("scale_factor", C.c_double),
        ("hz", C.c_double),
        ("snapshot_index", C.c_int32),
        ("reserved", C.c_int32),
    ]


class SubIO(C.Structure):
    _fields_ = [
        ("avg_pos", C.c_double * 3),
        ("avg_vel", C.c_double * 3),
        ("mostbound_pos", C.c_double * 3),
        ("mostbound_vel", C.c_double * 3),
        ("nbound", C.c_int64),
        ("sink_track_id", C.c_int64),
        ("snapshot_index_of_death", C.c_int32),
        ("snapshot_index_of_sink", C.c_int32),
        ("mbound", C.c_float),
        ("specific_self_potential_energy", C.c_float),
        ("specific_self_kinetic_energy", C.c_float),
        ("specific_angular_momentum", C.c_float * 3),
        ("nsource_full", C.c_int64),
        ("nsource", C.c_int64),
        ("iterations", C.c_int32),
        ("flags", C.c_int32),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("kernel_launches", C.c_int64),
        ("tree_builds", C.c_int64),
        ("walk_targets", C.c_int64),
        ("pair_interactions", C.c_int64),
        ("nodes_visited", C.c_int64),
        ("rounds", C.c_int64),
        ("walk_ms", C.c_double),
        ("build_ms", C.c_double),
        ("other_ms", C.c_double),
        ("h2d_ms", C.c_double),
        ("d2h_ms", C.c_double),
        ("h2d_bytes", C.c_int64),
        ("d2h_bytes", C.c_int64),
        ("execute_ms", C.c_double),
        ("tree_sources", C.c_int64),
        ("walk_fallbacks", C.c_int64),
        ("stage_wall_ms", C.c_double),
        ("execute_wall_ms", C.c_double),
        ("fetch_wall_ms", C.c_double),
        ("phase_ms", C.c_double * 8),
    ]


#: numpy view of SubIO so that batches of subhalo records are plain structured arrays
SUBIO_DTYPE = np.dtype(
    [
        ("avg_pos", "<f8", 3),
        ("avg_vel", "<f8", 3),
        ("mostbound_pos", "<f8", 3),
        ("mostbound_vel", "<f8", 3),
        ("nbound", "<i8"),
        ("sink_track_id", "<i8"),
        ("snapshot_index_of_death", "<i4"),
        ("snapshot_index_of_sink", "<i4"),
        ("mbound", "<f4"),
        ("specific_self_potential_energy", "<f4"),
        ("specific_self_kinetic_energy", "<f4"),
        ("specific_angular_momentum", "<f4", 3),
        ("nsource_full", "<i8"),
        ("nsource", "<i8"),
        ("iterations", "<i4"),
        ("flags", "<i4"),
    ],
    align=True,
)
assert SUBIO_DTYPE.itemsize == C.sizeof(SubIO), (SUBIO_DTYPE.itemsize, C.sizeof(SubIO))


class TrapIO(C.Structure):
    _fields_ = [
        ("mostbound_pos", C.c_double * 3),
        ("mostbound_vel", C.c_double * 3),
        ("nbound", C.c_int64),
        ("sink_track_id", C.c_int64),
        ("snapshot_index_of_sink", C.c_int32),
        ("is_merged", C.c_int32),
    ]


#: numpy view of TrapIO (hbtu_trap_io: merger trap detection, src/subhalo_merge.cpp:29-172)
TRAPIO_DTYPE = np.dtype([("mostbound_pos", "<f8", 3), ("mostbound_vel", "<f8", 3), ("nbound", "<i8"), ("sink_track_id", "<i8"),
                         ("snapshot_index_of_sink", "<i4"), ("is_merged", "<i4")], align=True)
assert TRAPIO_DTYPE.itemsize == C.sizeof(TrapIO)
TRAP_ARGTYPES = [C.POINTER(Epoch), C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int64),
                 C.POINTER(C.c_int32), C.POINTER(TrapIO)]


class ProfileIO(C.Structure):
    _fields_ = [
        ("mostbound_pos", C.c_double * 3),
        ("nbound", C.c_int64),
        ("mbound", C.c_float),
        ("rmax_comoving", C.c_float),
        ("vmax_physical", C.c_float),
        ("last_max_vmax_physical", C.c_float),
        ("snapshot_index_of_last_max_vmax", C.c_int32),
        ("r2sigma_comoving", C.c_float),
        ("rhalf_comoving", C.c_float),
        ("bound_r200crit_comoving", C.c_float),
        ("bound_m200crit", C.c_float),
        ("inertial_tensor", C.c_float * 6),
        ("inertial_tensor_weighted", C.c_float * 6),
        ("reserved", C.c_int32),
    ]


#: numpy view of ProfileIO (hbtu_profile_io: Subhalo_t::CalculateProfileProperties / CalculateShape, src/subhalo.cpp:242-398)
PROFILEIO_DTYPE = np.dtype(
    [
        ("mostbound_pos", "<f8", 3),
        ("nbound", "<i8"),
        ("mbound", "<f4"),
        ("rmax_comoving", "<f4"),
        ("vmax_physical", "<f4"),
        ("last_max_vmax_physical", "<f4"),
        ("snapshot_index_of_last_max_vmax", "<i4"),
        ("r2sigma_comoving", "<f4"),
        ("rhalf_comoving", "<f4"),
        ("bound_r200crit_comoving", "<f4"),
        ("bound_m200crit", "<f4"),
        ("inertial_tensor", "<f4", 6),
        ("inertial_tensor_weighted", "<f4", 6),
        ("reserved", "<i4"),
    ],
    align=True,
)
assert PROFILEIO_DTYPE.itemsize == C.sizeof(ProfileIO), (PROFILEIO_DTYPE.itemsize, C.sizeof(ProfileIO))
MASK_ARGTYPES = [C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int64),
                 C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
PROFILE_ARGTYPES = [C.POINTER(Epoch), C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_float), C.POINTER(ProfileIO)]


def f32(x: float) -> float:
    """Round to HBTReal=float, as the reference stores every Parameter_t real (V32 build)."""
    return float(np.float32(x))


def make_params(
    *,
    box_size: float,
    softening: float,
    periodic: bool = True,
    min_num_part_of_sub: int = 20,
    max_sample_size: int = 0,
    refine_mostbound: bool = True,
    bound_mass_precision: float = 0.995,
    source_sub_relax_factor: float = 3.0,
    open_angle: float = 0.45,
    mass_in_msunh: float = 1e10,
    length_in_mpch: float = 1.0,
    vel_in_kms: float = 1.0,
    device: int = 0,
    direct_sum_max: int = 0,
    shuffle_seed: int = 0,
) -> Params:
    """Derive the path's 14 config fields exactly as ``Parameter_t::ParseConfigFile`` does
    (reference: src/config_parser.cpp:95-103), in HBTReal=float arithmetic."""
    f = np.float32
    p = Params()
    p.struct_size = C.sizeof(Params)
    p.real_bytes = 4
    p.min_num_part_of_sub = min_num_part_of_sub
    p.periodic_boundary_on = int(periodic)
    p.refine_mostbound_particle = int(refine_mostbound)
    p.device = device
    p.max_sample_size = max_sample_size
    p.bound_mass_precision = f32(bound_mass_precision)
    p.source_sub_relax_factor = f32(source_sub_relax_factor)
    p.box_size = f32(box_size)
    p.box_half = float(f(f(box_size) / f(2.0)))
    p.softening_halo = f32(softening)
    oa = f(open_angle)
    p.tree_node_open_angle_square = float(f(oa * oa))
    # TreeNodeResolution=SofteningHalo*0.1 : float*double -> double, stored to float
    res = f(float(f(softening)) * 0.1)
    p.tree_node_resolution = float(res)
    p.tree_node_resolution_half = float(f(float(res) / 2.0))
    p.tree_alloc_factor = f32(0.8)
    p.tree_min_num_of_cells = 10
    # G=43.0071*(MassInMsunh/1e10)/VelInKmS/VelInKmS/LengthInMpch (double expr, stored to float)
    p.G = f32(43.0071 * (float(f(mass_in_msunh)) / 1e10) / float(f(vel_in_kms)) / float(f(vel_in_kms)) / float(f(length_in_mpch)))
    p.direct_sum_max = direct_sum_max
    p.shuffle_seed = shuffle_seed
    return p


def make_epoch(scale_factor: float, omega_m: float = 0.3, omega_l: float = 0.7, snapshot_index: int = 10, h0: float = 100.0) -> Epoch:
    """``Cosmology_t::Set`` (reference: src/snapshot.h:27-38) in HBTReal=float storage."""
    e = Epoch()
    a = scale_factor
    hratio = np.sqrt(omega_m / (a * a * a) + (1 - omega_m - omega_l) / (a * a) + omega_l)
    e.scale_factor = f32(a)
    e.hz = f32(hratio * f32(h0))
    e.snapshot_index = snapshot_index
    return e


def _ptr(a, ctype):
    if a is None:
        return None
    return a.ctypes.data_as(C.POINTER(ctype))


BATCH_ARGTYPES = [
    C.POINTER(Epoch),
    C.c_int64,
    C.POINTER(C.c_int64),
    C.POINTER(C.c_float),
    C.POINTER(C.c_float),
    C.POINTER(C.c_int64),
    C.POINTER(C.c_int32),
    C.POINTER(SubIO),
    C.c_int32,
    C.c_int64,
    C.POINTER(C.c_int64),
    C.POINTER(C.c_int32),
    C.POINTER(C.c_float),
]
POTENTIAL_ARGTYPES = [
    C.POINTER(Epoch),
    C.c_int64,
    C.POINTER(C.c_float),
    C.c_int64,
    C.POINTER(C.c_float),
    C.POINTER(C.c_float),
    C.POINTER(C.c_float),
    C.POINTER(C.c_double),
    C.POINTER(C.c_double),
    C.POINTER(C.c_double),
]

EXPORTS = [
    "hbtu_create",
    "hbtu_destroy",
    "hbtu_last_error",
    "hbtu_abi_version",
    "hbtu_order_capacity",
    "hbtu_unbind_batch",
    "hbtu_stage",
    "hbtu_execute",
    "hbtu_fetch",
    "hbtu_tree_potential",
    "hbtu_profile_batch",
    "hbtu_profile_executed",
    "hbtu_mask_batch",
    "hbtu_detect_traps",
    "hbtu_idtable_build",
    "hbtu_idtable_query",
    "hbtu_idtable_clear",
    "hbtu_get_stats",
    "hbtu_set_counting",
]


def order_capacity(part_offset: np.ndarray, nest_offset, nest_list) -> int:
    """Host mirror of ``hbtu_order_capacity``: sum over subhaloes of own + all descendants' particles (vectorised: it sits
    inside the timed end-to-end call of bench.py)."""
    nsub = len(part_offset) - 1
    own = np.diff(part_offset).astype(np.int64)
    if nest_offset is None or nsub == 0:
        return int(own.sum())
    cap = own.copy()
    parent = np.full(nsub, -1, np.int64)
    parent[np.asarray(nest_list, np.int64)] = np.repeat(np.arange(nsub), np.diff(np.asarray(nest_offset, np.int64)))
    depth = np.zeros(nsub, np.int64)
    cur = parent.copy()
    while (cur >= 0).any():  # nesting depth, level by level
        live = cur >= 0
        depth[live] += 1
        cur = np.where(live, parent[np.maximum(cur, 0)], -1)
    for d in range(int(depth.max()), 0, -1):  # children feed their parents, deepest level first
        idx = np.nonzero(depth == d)[0]
        np.add.at(cap, parent[idx], cap[idx])
    return int(cap.sum())


def load_library(path: str | None = None) -> C.CDLL:
    """Load the CUDA product library.  Fails loudly when it has not been built."""
    path = path or os.environ.get("HBTU_LIB_PATH") or LIB_PATH  # HBTU_LIB_PATH: A/B builds of the same library (profiles/)
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the unbinding path)"
        )
    lib = C.CDLL(path)
    lib.hbtu_create.argtypes = [C.POINTER(Params), C.POINTER(C.c_void_p)]
    lib.hbtu_create.restype = C.c_int
    lib.hbtu_destroy.argtypes = [C.c_void_p]
    lib.hbtu_destroy.restype = None
    lib.hbtu_last_error.argtypes = [C.c_void_p]
    lib.hbtu_last_error.restype = C.c_char_p
    lib.hbtu_abi_version.restype = C.c_int
    lib.hbtu_order_capacity.argtypes = [C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
    lib.hbtu_order_capacity.restype = C.c_int64
    lib.hbtu_unbind_batch.argtypes = [C.c_void_p] + BATCH_ARGTYPES
    lib.hbtu_unbind_batch.restype = C.c_int
    lib.hbtu_stage.argtypes = [C.c_void_p] + BATCH_ARGTYPES[:9]
    lib.hbtu_stage.restype = C.c_int
    lib.hbtu_execute.argtypes = [C.c_void_p]
    lib.hbtu_execute.restype = C.c_int
    lib.hbtu_fetch.argtypes = [C.c_void_p, C.POINTER(SubIO), C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_float)]
    lib.hbtu_fetch.restype = C.c_int
    lib.hbtu_tree_potential.argtypes = [C.c_void_p] + POTENTIAL_ARGTYPES
    lib.hbtu_tree_potential.restype = C.c_int
    lib.hbtu_profile_batch.argtypes = [C.c_void_p] + PROFILE_ARGTYPES
    lib.hbtu_profile_batch.restype = C.c_int
    lib.hbtu_profile_executed.argtypes = [C.c_void_p, C.POINTER(ProfileIO)]
    lib.hbtu_profile_executed.restype = C.c_int
    lib.hbtu_detect_traps.argtypes = [C.c_void_p] + TRAP_ARGTYPES
    lib.hbtu_detect_traps.restype = C.c_int
    lib.hbtu_mask_batch.argtypes = [C.c_void_p] + MASK_ARGTYPES
    lib.hbtu_mask_batch.restype = C.c_int
    lib.hbtu_idtable_build.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
    lib.hbtu_idtable_build.restype = C.c_int
    lib.hbtu_idtable_query.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.hbtu_idtable_query.restype = C.c_int
    lib.hbtu_idtable_clear.argtypes = [C.c_void_p]
    lib.hbtu_idtable_clear.restype = C.c_int
    lib.hbtu_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
    lib.hbtu_get_stats.restype = C.c_int
    return lib


def batch_args(epoch, part_offset, pos_mass, vel, nest_offset, nest_list, io, flags, order_capacity_, order_offset, order_out, energy_out):
    """Marshal numpy arrays into the positional tail of ``*_unbind_batch``."""
    nsub = len(part_offset) - 1
    return (
        C.byref(epoch),
        C.c_int64(nsub),
        _ptr(part_offset, C.c_int64),
        _ptr(pos_mass, C.c_float),
        _ptr(vel, C.c_float),
        _ptr(nest_offset, C.c_int64),
        _ptr(nest_list, C.c_int32),
        io.ctypes.data_as(C.POINTER(SubIO)),
        C.c_int32(flags),
        C.c_int64(order_capacity_),
        _ptr(order_offset, C.c_int64),
        _ptr(order_out, C.c_int32),
        _ptr(energy_out, C.c_float),
    )


def mask_args(part_offset, particle_id, nest_offset, nest_list, nbound, new_count, keep_index):
    """Marshal numpy arrays into the positional tail of ``*_mask_batch``."""
    return (C.c_int64(len(part_offset) - 1), _ptr(part_offset, C.c_int64), _ptr(particle_id, C.c_int64), _ptr(nest_offset, C.c_int64),
            _ptr(nest_list, C.c_int32), _ptr(nbound, C.c_int64), _ptr(new_count, C.c_int64), _ptr(keep_index, C.c_int32))


def trap_args(epoch, part_offset, pos_mass, vel, nest_offset, nest_list, io):
    """Marshal numpy arrays into the positional tail of ``*_detect_traps``."""
    return (C.byref(epoch), C.c_int64(len(part_offset) - 1), _ptr(part_offset, C.c_int64), _ptr(pos_mass, C.c_float), _ptr(vel, C.c_float),
            _ptr(nest_offset, C.c_int64), _ptr(nest_list, C.c_int32), io.ctypes.data_as(C.POINTER(TrapIO)))
