"""hbtplus_b200 - B200-native (sm_100a) implementation of HBT+'s subhalo unbinding hot path.

Only what the path needs lives here: ``csrc/`` (CUDA kernels + the C-ABI of include/hbt_unbind.h),
``capi`` (ctypes mirror of that ABI), ``unbind`` (host-side mirror of the reference interface),
``synth`` (synthetic snapshots for tests/bench) and ``sched`` (cost-weighted sharding over GPUs).
"""
__version__ = "0.1.0"
