"""dynast_b200 -- B200-native (sm_100a) implementation of dynast's count -> estimate hot path.

The public surface mirrors the reference's operator API
(dynast/preprocessing/__init__.py:2-31, dynast/estimation/__init__.py:2-10); the work happens in
hand-written CUDA kernels reached through the C-ABI in ``include/dynast_b200.h``.
There is no CPU fallback: importing the operators works anywhere, calling them needs a B200.
"""
__version__ = '0.1.0'
