"""exanbody_b200: B200-native (sm_100a) implementation of the exaNBody LJ hot path behind a C-ABI.

  include/xnb_hotpath.h          the drop-in boundary (C-ABI of libxnb_hotpath.so)
  exanbody_b200/csrc/            hand-written CUDA kernels + host orchestration
  exanbody_b200/capi.py          ctypes binding of the C-ABI (used by tests, bench.py and the operator mirror)
  exanbody_b200/operators.py     host-side mirror of the reference operators (names, slots, YAML keys)
"""
from .buildlib import build  # noqa: F401
