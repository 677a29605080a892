"""B200-native MultiMAE hot path (drop-in for the reference's ``multimae`` package).

Compute lives in hand-written sm_100a CUDA (``csrc/``) behind the C ABI of ``include/mmf_b200.h``;
``multimae/`` mirrors the reference's Python API on top of it.  There is no CPU / PyTorch fallback.
"""
__version__ = "0.1.0"
