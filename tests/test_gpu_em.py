"""GPU: the CUDA EM (through the C-ABI) is bit-identical to the reference's numba EM (golden vectors)
and to the C oracle on C4-like batches; north-star tolerance is 1e-6 relative, we assert equality."""
import numpy as np
import pytest

from helpers import em_golden, oracle_em_batch, synth_em_batch

pytestmark = pytest.mark.gpu

RTOL = 1e-6   # BASELINE.json north_star: fp64 estimates within 1e-6 relative


def test_em_golden_numba(gpu_ctx):
    from dynast_b200 import device
    g = em_golden()
    p_c, status, iters = device.em_pc_host(g['k'], g['n'], g['count'], g['off'], g['p_e'])
    assert (status == g['status']).all()
    ok = status == 0
    np.testing.assert_allclose(p_c[ok], g['p_c'][ok], rtol=RTOL)
    assert (p_c[ok] == g['p_c'][ok]).all()


def test_em_nasc_golden_numba(gpu_ctx):
    from dynast_b200 import device
    g = em_golden()
    nb = len(g['nasc_p_c'])
    p_c, status, _ = device.em_pc_host(g['k'], g['n'], g['count'], g['off'][:nb + 1], g['p_e'][:nb], nasc=True)
    assert (status == g['nasc_status']).all()
    ok = status == 0
    np.testing.assert_allclose(p_c[ok], g['nasc_p_c'][ok], rtol=RTOL)
    assert (p_c[ok] == g['nasc_p_c'][ok]).all()


def test_em_c4_like_vs_oracle(gpu_ctx, oracle_lib):
    from dynast_b200 import device
    k, n, c, off, pe = synth_em_batch(2000, seed=2004, adversarial_every=100)
    want, wst = oracle_em_batch(oracle_lib, k, n, c, off, pe)
    p_c, status, iters = device.em_pc_host(k, n, c, off, pe)
    assert (status == wst).all()
    ok = status == 0
    assert ok.sum() > 1500
    assert (p_c[ok] == want[ok]).all()
    assert iters.max() <= 300


def test_em_edge_cases(gpu_ctx, oracle_lib):
    from dynast_b200 import device
    # single cell, empty barcode in the middle, k >= 66 (factorial wraps to zero -> reference raises)
    k = np.array([0, 3, 70, 1], dtype=np.int32)
    n = np.array([10, 20, 149, 30], dtype=np.int32)
    c = np.array([5, 7, 1, 9], dtype=np.int64)
    off = np.array([0, 1, 1, 2, 4], dtype=np.int64)
    pe = np.array([1e-3, 1e-3, 1e-3, 1e-3])
    p_c, status, _ = device.em_pc_host(k, n, c, off, pe)
    want, wst = oracle_em_batch(oracle_lib, k, n, c, off, pe)
    assert list(status) == list(wst)
    assert status[3] == 2 and status[1] == 1
    ok = status == 0
    assert (p_c[ok] == want[ok]).all()
