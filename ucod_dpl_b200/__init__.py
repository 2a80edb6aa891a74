"""ucod_dpl_b200 — B200-native (sm_100a) implementation of the UCOD-DPL inference / pseudo-label hot path.

Host side mirrors the reference's Python API (see `ucod_dpl_b200.dropin`); all device work runs in the
hand-written CUDA library `csrc/libucod_b200.so` through the C-ABI in `include/ucod_b200.h`.
"""
__version__ = "0.1.0"
