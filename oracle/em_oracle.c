/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's p_c EM.
 *
 * Restates, in plain scalar C, the arithmetic of
 *   dynast/estimation/p_c.py:32-46   binomial_pmf            (numba @njit)
 *   dynast/estimation/p_c.py:111-172 expectation_maximization (numba @njit)
 *   dynast/estimation/p_c.py:49-108  expectation_maximization_nasc
 * including the properties that decide parity (SURVEY.md section 7 "EM parity traps"):
 *   - the binomial coefficient is the true division of two WRAPPING int64 products,
 *   - p**k with an integer exponent is numba's square-and-multiply (numba/cpython/numbers.py
 *     int_power_impl), not libm pow(),
 *   - the working copy of the histogram is float32; E-step sums are sequential float64 sums; the
 *     M-step sums are sequential FLOAT32 sums in row-major order and a float32 division (that is
 *     how numba types `new_values * np.arange(..)`); no fused multiply-add anywhere (compile with
 *     -ffp-contract=off),
 *   - the stop rule is exact equality of successive p_c, at most 300 iterations,
 *   - p_c <= p_e, a zero factorial product (k >= 66) or a zero E-step denominator (nasc) are
 *     failures: the reference raises and estimate_p_c drops the barcode (p_c.py:222-227).
 *
 * Pinned against the reference's own numba functions run in the build container
 * (oracle/make_golden.py -> tests/golden/em_*.npz) and against the fixture p_c_TC.csv.
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / --impl reference legs may
 * call this.  The product path never does.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EM_OK 0
#define EM_FAIL_PC_LE_PE 1   /* ValueError('p_c <= p_e')            p_c.py:169-170 */
#define EM_FAIL_ZERODIV 2    /* ZeroDivisionError inside numba code                */

/* a ** b for float a, int b >= 0 or < 0: numba int_power_impl. Returns 0 on ok, 1 on 1.0/0.0. */
static int ipow(double a, int64_t b, double *out) {
    double r = 1.0;
    int invert = 0;
    uint64_t e;
    if (b < 0) { invert = 1; e = (uint64_t)(-b); } else { e = (uint64_t)b; }
    while (e != 0) {
        if (e & 1) r *= a;
        e >>= 1;
        a *= a;
    }
    if (invert) {
        if (r == 0.0) return 1;
        r = 1.0 / r;
    }
    *out = r;
    return 0;
}

/* p_c.py:45 -- np.prod of int64 ranges wraps silently; "/" is float true division. */
static int binom_coef(int64_t k, int64_t n, double *out) {
    uint64_t num = 1, den = 1;
    for (int64_t i = 0; i < k; i++) {
        num *= (uint64_t)(i + (n - k + 1));
        den *= (uint64_t)(i + 1);
    }
    if ((int64_t)den == 0) return 1;
    *out = (double)(int64_t)num / (double)(int64_t)den;
    return 0;
}

int dyn_oracle_binomial_pmf(int64_t k, int64_t n, double p, double *out) {
    double coef, pk, qnk;
    if (binom_coef(k, n, &coef)) return EM_FAIL_ZERODIV;
    if (ipow(p, k, &pk)) return EM_FAIL_ZERODIV;
    if (ipow(1 - p, n - k, &qnk)) return EM_FAIL_ZERODIV;
    *out = coef * pk * qnk;
    return EM_OK;
}

/* values: dense int64 [K][N] row-major (K = max k + 1, N = max n + 1). */
int dyn_oracle_em(const int64_t *values, int K, int N, double p_e, double p_c, double threshold, int max_iters,
                  double *out_p_c, int *out_iters) {
    int status = EM_OK;
    uint8_t *mask = (uint8_t *)calloc((size_t)K * N, 1);
    float *v = (float *)malloc(sizeof(float) * (size_t)K * N);
    double *pmf = (double *)malloc(sizeof(double) * (size_t)K);
    int iters = 0;

    /* p_c.py:138-143 */
    for (int k = 0; k < K; k++) {
        for (int n = 0; n < N; n++) {
            int64_t s = 0;
            for (int kk = k; kk < K; kk++) s += values[(size_t)kk * N + n];
            double pm;
            if (dyn_oracle_binomial_pmf(k, n, p_e, &pm)) { status = EM_FAIL_ZERODIV; goto done; }
            double expected = (double)s * pm;
            if (expected > threshold * (double)values[(size_t)k * N + n]) mask[(size_t)k * N + n] = 1;
        }
    }
    for (size_t i = 0; i < (size_t)K * N; i++) v[i] = (float)values[i]; /* p_c.py:148 */

    for (int it = 0; it < max_iters; it++) {
        iters = it + 1;
        /* E step, p_c.py:151-160. keys sorted by (k, n); unmasked cells never change, so the sweep
         * order does not matter for the result. */
        for (int n = 0; n < N; n++) {
            int any = 0;
            for (int k = 0; k < K; k++) any |= mask[(size_t)k * N + n];
            if (!any) continue;
            for (int k = 0; k < K; k++)
                if (dyn_oracle_binomial_pmf(k, n, p_c, &pmf[k])) { status = EM_FAIL_ZERODIV; goto done; }
            for (int k = 0; k < K; k++) {
                if (!mask[(size_t)k * N + n]) continue;
                double numerator = 0, denominator = 0;
                for (int kp = 0; kp < K; kp++) {
                    if (mask[(size_t)kp * N + n]) continue;
                    numerator += pmf[k] * (double)v[(size_t)kp * N + n];
                    denominator += pmf[kp];
                }
                if (denominator > 0) v[(size_t)k * N + n] = (float)(numerator / denominator);
            }
        }
        /* M step, p_c.py:162-165.  As compiled by numba the array expression
         * `new_values * np.arange(..)` resolves to the float32 ufunc loop (float32 * int64 -> float32,
         * unlike NumPy), `.sum()` accumulates sequentially in the array dtype (float32) in row-major
         * order, and numerator / denominator is a float32 division widened to float64 afterwards. */
        float numerator = 0.0f, denominator = 0.0f;
        for (int k = 0; k < K; k++)
            for (int n = 0; n < N; n++) numerator += v[(size_t)k * N + n] * (float)k;
        for (int k = 0; k < K; k++)
            for (int n = 0; n < N; n++) denominator += v[(size_t)k * N + n] * (float)n;
        double prev = p_c;
        p_c = denominator > 0 ? (double)(numerator / denominator) : p_c;
        if (prev == p_c) break;
    }
    if (p_c <= p_e) status = EM_FAIL_PC_LE_PE;
done:
    free(mask); free(v); free(pmf);
    *out_p_c = p_c;
    if (out_iters) *out_iters = iters;
    return status;
}

/* p_c.py:49-108 (NASC-seq variant), restated as written (E step sums over kp IN the mask and the
 * sweep is Gauss-Seidel: cells updated earlier in the sorted (k, n) sweep are read by later ones). */
int dyn_oracle_em_nasc(const int64_t *values, int K, int N, double p_e, double threshold, double *out_p_c,
                       int *out_iters) {
    int status = EM_OK;
    uint8_t *mask = (uint8_t *)calloc((size_t)K * N, 1);
    float *v = (float *)malloc(sizeof(float) * (size_t)K * N);
    int iters = 0;
    double r = 1, l = 0, p_c = (r + l) / 2, prev;

    for (int k = 0; k < K; k++) {
        for (int n = 0; n < N; n++) {
            int64_t s = 0;
            for (int kk = k + 1; kk < K; kk++) s += values[(size_t)kk * N + n];
            double pm;
            if (dyn_oracle_binomial_pmf(k, n, p_e, &pm)) { status = EM_FAIL_ZERODIV; goto done; }
            double expected = (double)s * pm;
            if (expected > threshold * (double)values[(size_t)k * N + n]) mask[(size_t)k * N + n] = 1;
        }
    }
    for (size_t i = 0; i < (size_t)K * N; i++) v[i] = (float)values[i];

    while (r - l >= 1e-8) {
        iters++;
        for (int k = 0; k < K; k++) {
            for (int n = 0; n < N; n++) {
                if (!mask[(size_t)k * N + n]) continue;
                double numerator = 0, denominator = 0, pk;
                if (dyn_oracle_binomial_pmf(k, n, p_c, &pk)) { status = EM_FAIL_ZERODIV; goto done; }
                for (int kp = 0; kp < K; kp++) {
                    if (!mask[(size_t)kp * N + n]) continue;
                    double pkp;
                    if (dyn_oracle_binomial_pmf(kp, n, p_c, &pkp)) { status = EM_FAIL_ZERODIV; goto done; }
                    numerator += pk * (double)v[(size_t)kp * N + n];
                    denominator += pkp;
                }
                if (denominator == 0) { status = EM_FAIL_ZERODIV; goto done; }
                v[(size_t)k * N + n] = (float)(numerator / denominator);
            }
        }
        double numerator = 0, denominator = 0;
        for (int k = 0; k < K; k++)
            for (int n = 0; n < N; n++) {
                numerator += (double)k * (double)v[(size_t)k * N + n];
                denominator += (double)n * (double)v[(size_t)k * N + n];
            }
        prev = p_c;
        p_c = denominator > 0 ? numerator / denominator : 0;
        if (prev == p_c) break;
        if (p_c < prev) r = prev; else l = prev;
    }
done:
    free(mask); free(v);
    *out_p_c = p_c;
    if (out_iters) *out_iters = iters;
    return status;
}

/* Batch driver over CSR-like histograms: barcode b owns triplets [off[b], off[b+1]) of
 * (k, n, count); duplicates are summed (scipy csr_matrix semantics, p_c.py:219). */
int dyn_oracle_em_batch(const int32_t *k, const int32_t *n, const int64_t *count, const int64_t *off, int64_t B,
                        const double *p_e, int nasc, double *out_p_c, int32_t *out_status) {
    for (int64_t b = 0; b < B; b++) {
        int K = 0, N = 0;
        for (int64_t i = off[b]; i < off[b + 1]; i++) {
            if (k[i] + 1 > K) K = k[i] + 1;
            if (n[i] + 1 > N) N = n[i] + 1;
        }
        if (K == 0 || N == 0) { out_status[b] = EM_FAIL_PC_LE_PE; out_p_c[b] = 0; continue; }
        int64_t *dense = (int64_t *)calloc((size_t)K * N, sizeof(int64_t));
        for (int64_t i = off[b]; i < off[b + 1]; i++) dense[(size_t)k[i] * N + n[i]] += count[i];
        if (nasc)
            out_status[b] = dyn_oracle_em_nasc(dense, K, N, p_e[b], 0.01, &out_p_c[b], NULL);
        else
            out_status[b] = dyn_oracle_em(dense, K, N, p_e[b], 0.1, 0.01, 300, &out_p_c[b], NULL);
        free(dense);
    }
    return 0;
}
