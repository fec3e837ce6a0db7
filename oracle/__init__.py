"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference's count -> estimate hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs
may import this package.  Nothing under ``dynast-release_b200/`` does; the product has no CPU path.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, '_build', 'liboracle.so')
_lib = None


def lib():
    """liboracle.so (em_oracle.c, tally_oracle.c, ...), built by build.py:build_oracle()."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            import sys
            sys.path.insert(0, os.path.dirname(_HERE))
            import build
            build.build_oracle()
        _lib = C.CDLL(LIB_PATH)
        vp, i32, i64, u32, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_uint32, C.c_double
        _lib.dyn_oracle_binomial_pmf.argtypes = [i64, i64, dbl, C.POINTER(dbl)]
        _lib.dyn_oracle_em.argtypes = [vp, i32, i32, dbl, dbl, dbl, i32, C.POINTER(dbl), C.POINTER(i32)]
        _lib.dyn_oracle_em_nasc.argtypes = [vp, i32, i32, dbl, dbl, C.POINTER(dbl), C.POINTER(i32)]
        _lib.dyn_oracle_em_batch.argtypes = [vp, vp, vp, vp, i64, vp, i32, vp, vp]
        _lib.dyn_oracle_tally.argtypes = [vp, vp, i64, i64, vp, i64, i32, u32, vp, vp, vp, vp, vp, i64, vp, vp]
    return _lib
