"""khmer_b200 — B200-native k-mer ingestion backend for khmer's sketches.

Layers (bottom up): csrc/ (sm_100a kernels + the C ABI of include/kmgpu.h, built into libkmgpu.so),
cabi.py (literal ctypes view of that ABI), and the liboxli-compatible host layer.
"""
__version__ = "0.1.0"
