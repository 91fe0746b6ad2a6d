"""attentionshift_b200 -- B200-native (sm_100a) implementation of the AttentionShift hot path.

Host side: Python mirror of the reference's mmdet plugin surface
(``VisionTransformerDet`` backbone, ``AttnShiftRoIHead.seed_pseudo_gt``).
Device side: hand-written CUDA kernels behind a C ABI (``include/attnshift_b200.h``),
loaded with ctypes from ``attentionshift_b200/csrc/libattnshift_b200.so``.
There is NO CPU fallback: every op raises if the CUDA library is missing.
"""
__version__ = "0.1.0"
