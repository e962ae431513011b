"""snap_b200: host-side mirror of the C ABI of the B200-native SNAP seed-and-extend engine.

The product is the CUDA library built from snap_b200/csrc (include/snapgpu.h).  Importing this package does
not load it; the first engine call does, and raises if it is missing (no CPU fallback).
"""
from . import synth  # noqa: F401
