"""CPU: the C-ABI library loads and exports every symbol include/dynast_b200.h declares (no compute)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'dynast_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(dyn_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_exported():
    import build
    lib_path = build.build_product()
    lib = ctypes.CDLL(lib_path)
    syms = declared_symbols()
    assert len(syms) >= 9
    for s in syms:
        assert hasattr(lib, s), f'{s} declared in include/dynast_b200.h but not exported'


def test_python_prototypes_cover_header():
    from dynast_b200 import _lib
    assert set(declared_symbols()) == set(_lib.PROTOTYPES)
    _lib.load()


def test_no_device_is_a_loud_error():
    """On a box without a GPU the product must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        return
    from dynast_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.dyn_ctx_create(0, ctypes.byref(h)) != 0
    assert lib.dyn_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'dynast-release_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.cpp', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f
                assert 'liboracle' not in src, f
