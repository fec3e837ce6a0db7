"""TEST INFRASTRUCTURE ONLY -- posterior means of dynast/models/pi.stan (pi.stan:4-34) by QUADPACK.

PARITY UNPINNED: the reference obtains (alpha, beta, pi) as means of 1000 NUTS draws from pystan 2.19
(pi.py:174-186); pystan is not installable here and the reference's own test mocks it
(tests/estimation/test_pi.py:34-62), so there is no deterministic golden vector.  This oracle integrates the
same posterior with a method independent of the CUDA kernel's (scipy QUADPACK `qawse` algebraic-endpoint rules
for pi, 40 x 40 Gauss-Legendre for (log_alpha, log_beta)); agreement with the reference is checked
statistically against its committed pi_*.csv fixtures.
"""
import numpy as np
from scipy import integrate, special


def posterior_means(k, n, count, p_e, p_c, n_gl=40):
    k = np.asarray(k, dtype=np.float64); n = np.asarray(n, dtype=np.float64); c = np.asarray(count, dtype=np.float64)
    keep = (c > 0) & (n > 0)
    k, n, c = k[keep], n[keep], c[keep]
    la = k * np.log(p_c) + (n - k) * np.log1p(-p_c)
    lb = k * np.log(p_e) + (n - k) * np.log1p(-p_e)
    m = np.maximum(la, lb)
    a, b = np.exp(la - m), np.exp(lb - m)

    def loglik(pi):
        return float(np.sum(c * np.log(pi * a + (1.0 - pi) * b)))

    grid = np.linspace(1e-6, 1 - 1e-6, 20001)
    ll = np.array([loglik(x) for x in grid[::40]])
    lmax = ll.max()
    # refine the maximum
    from scipy.optimize import minimize_scalar
    res = minimize_scalar(lambda x: -loglik(x), bounds=(1e-12, 1 - 1e-12), method='bounded', options={'xatol': 1e-14})
    lmax = max(lmax, -res.fun, loglik(1e-300), loglik(1 - 1e-16))
    mode = min(max(res.x, 0.02), 0.98)

    def L(pi):
        return np.exp(loglik(pi) - lmax)

    x, w = np.polynomial.legendre.leggauss(n_gl)
    x, w = 3.0 * x, 3.0 * w
    Z = Ea = Eb = Ep = 0.0
    terms = []
    for i in range(n_gl):
        al = np.exp(x[i])
        for j in range(n_gl):
            be = np.exp(x[j])
            lo0, _ = integrate.quad(lambda t: L(t) * (1 - t)**(be - 1), 0, mode, weight='alg', wvar=(al - 1, 0), epsabs=0, epsrel=1e-11, limit=400)
            hi0, _ = integrate.quad(lambda t: L(t) * t**(al - 1), mode, 1, weight='alg', wvar=(0, be - 1), epsabs=0, epsrel=1e-11, limit=400)
            lo1, _ = integrate.quad(lambda t: t * L(t) * (1 - t)**(be - 1), 0, mode, weight='alg', wvar=(al - 1, 0), epsabs=0, epsrel=1e-11, limit=400)
            hi1, _ = integrate.quad(lambda t: t * L(t) * t**(al - 1), mode, 1, weight='alg', wvar=(0, be - 1), epsabs=0, epsrel=1e-11, limit=400)
            logB = special.gammaln(al) + special.gammaln(be) - special.gammaln(al + be)
            terms.append((np.log(w[i] * w[j]) - logB, lo0 + hi0, lo1 + hi1, al, be))
    sc = max(t[0] + np.log(t[1]) for t in terms if t[1] > 0)
    for lw, m0, m1, al, be in terms:
        if m0 <= 0:
            continue
        wt = np.exp(lw + np.log(m0) - sc)
        Z += wt; Ea += wt * al; Eb += wt * be; Ep += wt * m1 / m0
    return Ea / Z, Eb / Z, Ep / Z


def posterior_means_sampled(k, n, count, p_e, p_c, draws=2_000_000, seed=0):
    """The same posterior means by plain Monte Carlo, independent of any quadrature: (log_alpha, log_beta) ~ U(-3, 3)^2
    and pi ~ Beta(alpha, beta) are drawn from the PRIOR of pi.stan (pi.stan:20-25) and weighted by the likelihood
    (pi.stan:27-33) -- self-normalised importance sampling.  Returns (alpha, beta, pi, effective sample size); the
    Monte-Carlo error is about sd / sqrt(ESS)."""
    rng = np.random.default_rng(seed)
    k = np.asarray(k, dtype=np.float64); n = np.asarray(n, dtype=np.float64); c = np.asarray(count, dtype=np.float64)
    keep = (c > 0) & (n > 0)
    k, n, c = k[keep], n[keep], c[keep]
    la = k * np.log(p_c) + (n - k) * np.log1p(-p_c)
    lb = k * np.log(p_e) + (n - k) * np.log1p(-p_e)
    m = np.maximum(la, lb)
    a, b = np.exp(la - m), np.exp(lb - m)
    tot = np.zeros(4)
    w2 = 0.0
    lmax = None
    for _ in range(max(1, draws // 200_000)):
        al = np.exp(rng.uniform(-3, 3, 200_000))
        be = np.exp(rng.uniform(-3, 3, 200_000))
        pi = np.clip(rng.beta(al, be), 1e-300, 1 - 1e-16)
        ll = (c[None, :] * np.log(pi[:, None] * a[None, :] + (1.0 - pi[:, None]) * b[None, :])).sum(axis=1)
        if lmax is None:
            lmax = ll.max()
        w = np.exp(ll - lmax)
        tot += np.array([w.sum(), (w * al).sum(), (w * be).sum(), (w * pi).sum()])
        w2 += (w * w).sum()
    return tot[1] / tot[0], tot[2] / tot[0], tot[3] / tot[0], tot[0]**2 / w2
