"""
phyly_b200 -- B200-native phylogenetic likelihood engine, drop-in for the hot
path of argriffing/phyly (arbplf-ll / -deriv / -marginal / -dwell / -trans).

The compute path is hand-written CUDA for sm_100a behind a C ABI
(include/plf.h, include/arbplf.h, built into phyly_b200/lib/libarbplf_b200.so).
This Python package is only a binding; it contains no numerical fallback.
"""
__all__ = ["engine", "_lib"]
