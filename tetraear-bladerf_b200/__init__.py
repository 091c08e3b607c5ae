"""tetraear_b200 — B200-native IQ->bursts receive path behind TetraEar's Python API.

Mirrors ``tetraear.signal`` (SignalProcessor, compute_spectrum) and ``tetraear.core``
(TetraDecoder, TetraProtocolParser, TetraBurst, BurstType) of the reference
(/root/reference/tetraear/__init__.py:25-36).  All arithmetic runs in hand-written sm_100a
CUDA kernels reached through the C-ABI in ``include/tetra_b200.h``; there is no CPU
fallback — importing the compute entry points without the built library or without a
CUDA device raises.
"""
__version__ = "0.1.0"
