"""epilogos_b200 -- B200 (sm_100a) implementation of the Epilogos scoring hot path.

Host layer (Python, mirrors the reference's stage functions):
    epilogos_b200.expected.main / expectedCombination.main / scores.main
Device layer: hand-written CUDA kernels behind the C ABI of include/epilogos_b200.h, bound with ctypes in
epilogos_b200._lib and wrapped for torch device tensors in epilogos_b200.engine.
"""
__version__ = "0.1.0"
