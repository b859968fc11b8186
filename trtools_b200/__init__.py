"""
trtools_b200 — B200-native (sm_100a) implementation of the TRTools per-record hot path
(HarmonizeRecord -> TRRecord accessors -> statSTR / dumpSTR / associaTR) behind TRTools' own
Python API.  Host code is Python calling hand-written CUDA through a ctypes C-ABI
(include/trtools_b200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
