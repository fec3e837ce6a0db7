"""CPU: the pi oracle (oracle/pi_oracle.py, QUADPACK quadrature of dynast/models/pi.stan's posterior means) against an
independent Monte-Carlo estimate of the same means (prior draws weighted by the likelihood).  PARITY WITH THE
REFERENCE STAYS UNPINNED -- pystan is absent and the reference's own test mocks it (tests/estimation/test_pi.py:34-62) --
this pins the oracle to the MODEL by a second method whose error is ~1e-3, ten times tighter than the fixture CSVs'."""
import numpy as np
import pytest


@pytest.mark.parametrize('seed,pi_true,reads', [(1, 0.3, 60), (2, 0.7, 25)])
def test_quadrature_oracle_agrees_with_importance_sampling(seed, pi_true, reads):
    from oracle import pi_oracle
    rng = np.random.default_rng(seed)
    p_e, p_c = 0.001, 0.05
    n = np.clip(rng.poisson(23, reads), 1, 150)
    k = rng.binomial(n, np.where(rng.random(reads) < pi_true, p_c, p_e))
    key, cnt = np.unique(k * 1000 + n, return_counts=True)
    kk, nn = key // 1000, key % 1000
    q = pi_oracle.posterior_means(kk, nn, cnt, p_e, p_c, n_gl=16)
    s = pi_oracle.posterior_means_sampled(kk, nn, cnt, p_e, p_c, draws=3_000_000, seed=seed)
    assert s[3] > 20_000                      # effective sample size
    assert abs(q[2] - s[2]) < 3e-3            # pi
    assert abs(q[0] - s[0]) / s[0] < 2e-2 and abs(q[1] - s[1]) / s[1] < 2e-2       # alpha, beta (wide posteriors)
