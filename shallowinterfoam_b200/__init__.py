"""shallowinterfoam_b200 -- B200-native drop-in for shallowInterFoam's 3D-region PISO/VOF hot path.

The compute path is hand-written sm_100a CUDA behind the C-ABI of include/b200foam.h (libb200foam.so).
This Python package is only the thin host-side binding used by tests and bench.py.
"""
from .capi import load, B200Lib, B200Error  # noqa: F401
