// Batched labeled-fraction fit.
//
// Replaces dynast/estimation/pi.py:118-187 (fit_stan_mcmc: one NUTS chain, 1000 warm-up + 1000 draws per
// group through pystan) for the model dynast/models/pi.stan:4-34:
//     log_alpha, log_beta ~ U(-3, 3);  pi_g ~ Beta(e^log_alpha, e^log_beta);
//     target += count_i * log( pi_g * Binom(k_i | n_i, p_c) + (1 - pi_g) * Binom(k_i | n_i, p_e) )
// The reference reports the posterior MEANS of alpha, beta, pi_g (pi.py:186).  Those are three ratios of
// integrals over a 3-dimensional box, computed here deterministically, one CTA per group:
//   * the likelihood L(pi) is log-concave in pi, so its mode and the window [a, b] outside which it has
//     dropped by e^-60 are found by a 9-section search (one warp per probe point); all pi-integrals run over that window;
//   * pi: tanh-sinh (double exponential) quadrature on [a, b] -- exact handling of the integrable endpoint
//     singularities pi^(alpha-1), (1-pi)^(beta-1) (alpha, beta go down to e^-3) when the window touches 0 or 1;
//   * (log_alpha, log_beta): 32 x 32 Gauss-Legendre nodes on (-3, 3)^2;
//   * the Beta prior factorises over the nodes, pi^(alpha_a - 1) * (1 - pi)^(beta_b - 1), so the inner
//     integral for all 1024 hyper-parameter nodes is a bilinear form  A * diag(w L) * B^T  (fp64 FMA);
//   * the binomial coefficients cancel between numerator and denominator and are never formed.
// Accuracy ~1e-9 relative (checked against QUADPACK in tests/test_oracle_pi.py); the reference's own Monte-
// Carlo error is ~1e-2, its tests mock Stan (tests/estimation/test_pi.py:34-62): parity unpinned.
// fp64-pipe bound, not HBM bound: 12 B per histogram cell in, 24 B per group out.
#include <math.h>

#include <vector>

#include "common.cuh"

namespace {

constexpr int kPiThreads = 256;
constexpr int kNA = 32;            // Gauss-Legendre nodes per hyper-parameter
constexpr int kNJ = 257;           // tanh-sinh nodes
static_assert(kNJ >= kPiThreads, "one node per thread");
constexpr int kChunk = 64;         // pi nodes per bilinear-form chunk
constexpr double kDrop = 60.0;     // window: log L within this of its maximum

struct PiParams {
    const int32_t *k, *n;
    const long long *count;
    const long long *off;
    long long n_groups;
    const double *p_e, *p_c;
    double *out;        // [G][3] alpha, beta, pi
    int32_t *status;
    const double *gl_x, *gl_w;       // [kNA] nodes (log-parameter) and weights
    const double *log_beta_fn;       // [kNA][kNA] log B(alpha_a, beta_b)
    const double *ts_s, *ts_c, *ts_w;   // [kNJ] s_j in (0,1), 1 - s_j, weight
    int max_cells;
};

__device__ __forceinline__ double block_sum(double v, double *red) {
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < kPiThreads / 32; ++w) s += red[w];
    return s;
}
__device__ __forceinline__ double block_max(double v, double *red) {
    for (int d = 16; d; d >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, d));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = red[0];
    for (int w = 1; w < kPiThreads / 32; ++w) s = fmax(s, red[w]);
    return s;
}

// log L(pi) and d/dpi log L(pi); `q` = 1 - pi passed separately to keep precision near 1
__device__ __forceinline__ void loglik_partial(const double *ca, const double *cb, const double *cc, int cells, double pi,
                                               double q, double &ll, double &dl) {
    ll = 0.0; dl = 0.0;
    for (int i = threadIdx.x; i < cells; i += kPiThreads) {
        const double m = pi * ca[i] + q * cb[i];
        ll += cc[i] * log(m);
        dl += cc[i] * (ca[i] - cb[i]) / m;
    }
}

// By ONE warp (lanes over the cells, shuffle reduction), every lane gets the total: d/dpi log L (DERIV, no
// logarithm needed) or log L (no division needed).
template <bool DERIV>
__device__ __forceinline__ double loglik_warp(const double *ca, const double *cb, const double *cc, int cells, double pi,
                                              double q) {
    double acc = 0.0;
    for (int i = threadIdx.x & 31; i < cells; i += 32) {
        const double m = pi * ca[i] + q * cb[i];
        acc += DERIV ? cc[i] * (ca[i] - cb[i]) / m : cc[i] * log(m);
    }
    for (int d = 16; d; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    return acc;
}

constexpr int kSections = kPiThreads / 32 + 1;     // interior probe points per round = warps per CTA
constexpr int kRounds = 8;                         // (1 / 9)^8 = 2e-8 of the starting interval: the window edges sit where
                                                   // the integrand is e^-60 of its maximum, the mode only fixes that level

// Narrows [lo, hi] around the point where a monotone predicate flips, kSections-section search: in every round
// warp w evaluates the likelihood at its own interior point and the CTA keeps the sub-interval of the flip --
// 8 rounds of one barrier pair instead of 50-60 bisection steps of two block reductions each.
//   MODE = 0: predicate d/dpi log L > 0 (true left of the mode);  MODE = 1: log L >= thr, true right of the flip
//   (left window edge);  MODE = 2: log L >= thr, true left of the flip (right window edge).
template <int MODE>
__device__ __forceinline__ void section_search(const double *ca, const double *cb, const double *cc, int cells, double thr,
                                               double &lo, double &hi, double *s_val) {
    const int w = threadIdx.x >> 5;
    for (int r = 0; r < kRounds; ++r) {
        const double step = (hi - lo) / kSections;
        const double x = lo + step * (w + 1);
        const double val = loglik_warp<MODE == 0>(ca, cb, cc, cells, x, 1.0 - x);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) s_val[w] = val;
        __syncthreads();
        // first interior point where the predicate differs from its value at `lo`
        int f = kSections - 1;
        for (int j = kSections - 2; j >= 0; --j) {
            const bool pred = MODE == 0 ? s_val[j] > 0.0 : s_val[j] >= thr;
            const bool at_lo = MODE != 1;      // MODE 0 / 2: true at lo;  MODE 1: false at lo
            if (pred != at_lo) f = j;
        }
        const double nlo = lo + step * f, nhi = f == kSections - 1 ? hi : lo + step * (f + 1);
        lo = nlo; hi = nhi;
    }
}

__global__ void __launch_bounds__(kPiThreads) pi_kernel(const PiParams p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    double *ca = reinterpret_cast<double *>(smem_raw);        // [max_cells] scaled Binom(k|n,p_c)
    double *cb = ca + p.max_cells;                             // [max_cells] scaled Binom(k|n,p_e)
    double *cc = cb + p.max_cells;                             // [max_cells] read counts
    double *tA = cc + p.max_cells;                             // [kChunk][kNA]
    double *tB = tA + kNA * kChunk;                            // [kChunk][kNA]
    double *v = tB + kNA * kChunk;                             // [kNJ] w_j * L_j
    double *lp = v + kNJ + 1;                                  // [kNJ] log pi_j
    double *lq = lp + kNJ + 1;                                 // [kNJ] log (1 - pi_j)
    double *pj = lq + kNJ + 1;                                 // [kNJ] pi_j
    __shared__ double red[kPiThreads / 32];
    __shared__ double s_val[kPiThreads / 32];
    __shared__ double s_lo, s_hi;

    const int tid = threadIdx.x;
    for (long long g = blockIdx.x; g < p.n_groups; g += gridDim.x) {
        __syncthreads();
        const long long t0 = p.off[g], t1 = p.off[g + 1];
        const int cells = (int)(t1 - t0);
        const double p_e = p.p_e[g], p_c = p.p_c[g];
        double *out = p.out + g * 3;
        // pi.stan:9-10: p_c, p_e in [0, 1]; binomial_lpmf(k | n, 0) is 0 for k == 0 and -inf otherwise
        if (cells <= 0 || cells > p.max_cells || !(p_e >= 0.0) || !(p_c >= 0.0) || !(p_e <= 1.0) || !(p_c <= 1.0)) {
            if (tid == 0) { out[0] = out[1] = out[2] = nan(""); p.status[g] = cells <= 0 ? 1 : (cells > p.max_cells ? 3 : 2); }
            continue;
        }
        const double lpc = log(p_c), lqc = log1p(-p_c), lpe = log(p_e), lqe = log1p(-p_e);
        double total = 0.0, impossible = 0.0;
        for (int i = tid; i < cells; i += kPiThreads) {
            const double k = (double)p.k[t0 + i], n = (double)p.n[t0 + i];
            // x * log(y) with 0 * log(0) = 0
            const double la = (k > 0.0 ? k * lpc : 0.0) + (n - k > 0.0 ? (n - k) * lqc : 0.0);
            const double lb = (k > 0.0 ? k * lpe : 0.0) + (n - k > 0.0 ? (n - k) * lqe : 0.0);
            const double m = fmax(la, lb);
            const long long c = p.count[t0 + i];
            const bool ok = c > 0 && n > 0.0 && k <= n;      // pi.py:223: base > 0 and count > 0
            if (ok && !(m > -INFINITY)) impossible += 1.0;   // the cell has probability zero under both rates
            ca[i] = (m > -INFINITY) ? exp(la - m) : 1.0; cb[i] = (m > -INFINITY) ? exp(lb - m) : 1.0; cc[i] = ok ? (double)c : 0.0;
            total += cc[i];
        }
        total = block_sum(total, red);
        impossible = block_sum(impossible, red);
        if (impossible > 0.0) {   // Stan cannot initialise: the reference drops the group (pi.py:298-304)
            if (tid == 0) { out[0] = out[1] = out[2] = nan(""); p.status[g] = 2; }
            continue;
        }
        if (!(total > 0.0)) {
            if (tid == 0) { out[0] = out[1] = out[2] = nan(""); p.status[g] = 1; }
            continue;
        }
        // ---- mode of log L on [0, 1] (derivative is monotone decreasing)
        double ll, dl, mode;
        {
            double lo = 0.0, hi = 1.0;
            loglik_partial(ca, cb, cc, cells, 0.0, 1.0, ll, dl);
            const double d0 = block_sum(dl, red);
            loglik_partial(ca, cb, cc, cells, 1.0, 0.0, ll, dl);
            const double d1 = block_sum(dl, red);
            if (!(d0 > 0.0)) mode = 0.0;
            else if (!(d1 < 0.0)) mode = 1.0;
            else {
                section_search<0>(ca, cb, cc, cells, 0.0, lo, hi, s_val);
                mode = 0.5 * (lo + hi);
            }
        }
        loglik_partial(ca, cb, cc, cells, mode, 1.0 - mode, ll, dl);
        const double ll_max = block_sum(ll, red);
        // ---- window [a, b]: log L >= ll_max - kDrop
        double wa = 0.0, wb = 1.0;
        {
            loglik_partial(ca, cb, cc, cells, 0.0, 1.0, ll, dl);
            const double l0 = block_sum(ll, red);          // may be -inf
            if (mode > 0.0 && !(l0 >= ll_max - kDrop)) {
                double lo = 0.0, hi = mode;
                section_search<1>(ca, cb, cc, cells, ll_max - kDrop, lo, hi, s_val);
                wa = lo;
            }
            loglik_partial(ca, cb, cc, cells, 1.0, 0.0, ll, dl);
            const double l1 = block_sum(ll, red);
            if (mode < 1.0 && !(l1 >= ll_max - kDrop)) {
                double lo = mode, hi = 1.0;
                section_search<2>(ca, cb, cc, cells, ll_max - kDrop, lo, hi, s_val);
                wb = hi;
            }
        }
        if (tid == 0) { s_lo = wa; s_hi = wb; }
        __syncthreads();
        const double a = s_lo, b = s_hi, width = b - a;
        // ---- tanh-sinh nodes on [a, b]
        // one node per thread; the nodes beyond the thread count (kNJ = 257: one) are summed by the whole CTA instead
        // of making one thread walk the cells twice
        for (int j = kPiThreads; j < kNJ; ++j) {
            const double pi = a + width * p.ts_s[j], q = (1.0 - b) + width * p.ts_c[j];
            double l = 0.0;
            for (int i = tid; i < cells; i += kPiThreads) l += cc[i] * log(pi * ca[i] + q * cb[i]);
            l = block_sum(l, red);
            if (tid == 0) { pj[j] = pi; lp[j] = log(pi); lq[j] = log(q); v[j] = l; }
        }
        {
            const int j = tid;
            const double s = p.ts_s[j], c = p.ts_c[j];
            const double pi = a + width * s, q = (1.0 - b) + width * c;
            double l = 0.0;
            for (int i = 0; i < cells; ++i) l += cc[i] * log(pi * ca[i] + q * cb[i]);
            pj[j] = pi;
            lp[j] = log(pi);
            lq[j] = log(q);
            v[j] = l;   // log L for now
        }
        __syncthreads();
        double lmax = -INFINITY;
        for (int j = tid; j < kNJ; j += kPiThreads) lmax = fmax(lmax, v[j]);
        lmax = block_max(lmax, red);
        for (int j = tid; j < kNJ; j += kPiThreads) v[j] = p.ts_w[j] * width * exp(v[j] - lmax);
        __syncthreads();
        // ---- bilinear forms: thread owns hyper-parameter nodes (a_i, b0..b0+3)
        const int ai = tid >> 3, b0 = (tid & 7) << 2;
        double m0[4] = {0, 0, 0, 0}, m1[4] = {0, 0, 0, 0};
        for (int j0 = 0; j0 < kNJ; j0 += kChunk) {
            const int nj = min(kChunk, kNJ - j0);
            __syncthreads();
            // tables are [node][hyper-parameter]: the 8 x 4 (resp. 4) values a warp reads per node are contiguous
            for (int e = tid; e < kNA * kChunk; e += kPiThreads) {
                const int r = e % kNA, jj = e / kNA;
                if (jj < nj) {
                    const double ex = exp(p.gl_x[r]) - 1.0;     // alpha_r - 1 == beta_r - 1 (same nodes)
                    tA[e] = exp(ex * lp[j0 + jj]) * v[j0 + jj];
                    tB[e] = exp(ex * lq[j0 + jj]);
                } else { tA[e] = 0.0; tB[e] = 0.0; }
            }
            __syncthreads();
            for (int jj = 0; jj < nj; ++jj) {
                const double x = tA[jj * kNA + ai], xp = x * pj[j0 + jj];
                const double2 y01 = *reinterpret_cast<const double2 *>(tB + jj * kNA + b0);
                const double2 y23 = *reinterpret_cast<const double2 *>(tB + jj * kNA + b0 + 2);
                const double y[4] = {y01.x, y01.y, y23.x, y23.y};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    m0[u] = fma(x, y[u], m0[u]);
                    m1[u] = fma(xp, y[u], m1[u]);
                }
            }
        }
        // ---- combine over the hyper-parameter grid.  log B(alpha, beta) and the largest term are taken
        //      out in log space so that nothing under/overflows.
        double lt[4], ltmax = -INFINITY;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            lt[u] = (m0[u] > 0.0) ? log(m0[u]) - p.log_beta_fn[ai * kNA + b0 + u] + log(p.gl_w[ai] * p.gl_w[b0 + u]) : -INFINITY;
            ltmax = fmax(ltmax, lt[u]);
        }
        ltmax = block_max(ltmax, red);
        double z = 0.0, ea = 0.0, eb = 0.0, ep = 0.0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (m0[u] > 0.0) {
                const double w = exp(lt[u] - ltmax);
                z += w;
                ea += w * exp(p.gl_x[ai]);
                eb += w * exp(p.gl_x[b0 + u]);
                ep += w * (m1[u] / m0[u]);
            }
        }
        z = block_sum(z, red); ea = block_sum(ea, red); eb = block_sum(eb, red); ep = block_sum(ep, red);
        if (tid == 0) {
            if (z > 0.0) { out[0] = ea / z; out[1] = eb / z; out[2] = ep / z; p.status[g] = 0; }
            else { out[0] = out[1] = out[2] = nan(""); p.status[g] = 2; }
        }
    }
}

// Gauss-Legendre nodes on (-1, 1) by Newton iteration on P_n
void gauss_legendre(int n, std::vector<double> &x, std::vector<double> &w) {
    x.resize(n); w.resize(n);
    for (int i = 0; i < (n + 1) / 2; ++i) {
        double z = cos(M_PI * (i + 0.75) / (n + 0.5)), pp = 0.0;
        for (int it = 0; it < 100; ++it) {
            double p1 = 1.0, p2 = 0.0;
            for (int j = 0; j < n; ++j) { const double p3 = p2; p2 = p1; p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1.0); }
            pp = n * (z * p1 - p2) / (z * z - 1.0);
            const double z1 = z;
            z = z1 - p1 / pp;
            if (fabs(z - z1) < 1e-16) break;
        }
        x[i] = -z; x[n - 1 - i] = z;
        w[i] = w[n - 1 - i] = 2.0 / ((1.0 - z * z) * pp * pp);
    }
}

}  // namespace

struct dyn_pi_tables {
    double *dev;   // gl_x[kNA] gl_w[kNA] logB[kNA*kNA] ts_s[kNJ] ts_c[kNJ] ts_w[kNJ]
};

static int pi_tables(dyn_ctx *ctx, cudaStream_t stream, const double **gl_x, const double **gl_w, const double **logb,
                     const double **ts_s, const double **ts_c, const double **ts_w) {
    const size_t total = 2 * kNA + kNA * kNA + 3 * kNJ;
    void *buf = nullptr;
    const bool fresh = ctx->scratch_bytes[10] < total * sizeof(double);
    int rc = dyn_ctx_scratch(ctx, 10, total * sizeof(double), &buf);
    if (rc) return rc;
    double *d = static_cast<double *>(buf);
    if (fresh) {
        std::vector<double> h(total), x, w;
        gauss_legendre(kNA, x, w);
        for (int i = 0; i < kNA; ++i) { h[i] = 3.0 * x[i]; h[kNA + i] = 3.0 * w[i]; }
        for (int a = 0; a < kNA; ++a)
            for (int b = 0; b < kNA; ++b) {
                const double al = exp(h[a]), be = exp(h[b]);
                h[2 * kNA + a * kNA + b] = lgamma(al) + lgamma(be) - lgamma(al + be);
            }
        // tanh-sinh on (0, 1): s = 1 / (1 + exp(-2u)), u = (pi/2) sinh t, t = (j - J) h
        const int J = (kNJ - 1) / 2;
        const double T = 6.0, hstep = T / J;   // smallest node ~1e-275: pi^(alpha-1) stays finite for alpha >= e^-3
        double *s = h.data() + 2 * kNA + kNA * kNA, *c = s + kNJ, *ww = c + kNJ;
        for (int j = 0; j < kNJ; ++j) {
            const double t = (j - J) * hstep, u = 0.5 * M_PI * sinh(t);
            s[j] = 1.0 / (1.0 + exp(-2.0 * u));
            c[j] = 1.0 / (1.0 + exp(2.0 * u));
            ww[j] = hstep * 0.5 * M_PI * cosh(t) * 2.0 * s[j] * c[j];
        }
        DYN_CUDA(cudaMemcpyAsync(d, h.data(), total * sizeof(double), cudaMemcpyHostToDevice, stream));
        DYN_CUDA(cudaStreamSynchronize(stream));   // h goes out of scope
    }
    *gl_x = d; *gl_w = d + kNA; *logb = d + 2 * kNA;
    *ts_s = d + 2 * kNA + kNA * kNA; *ts_c = *ts_s + kNJ; *ts_w = *ts_c + kNJ;
    return DYN_OK;
}

extern "C" int dyn_pi_fit(dyn_ctx_t *ctx, const int32_t *d_k, const int32_t *d_n, const int64_t *d_count,
                          const int64_t *d_off, int64_t n_groups, int64_t max_cells, const double *d_p_e,
                          const double *d_p_c, double *d_alpha_beta_pi, int32_t *d_status, void *stream_) {
    DYN_ARG(ctx != nullptr, "null context");
    CtxEnter enter_(ctx, stream_);
    DYN_ARG(n_groups >= 0, "negative n_groups");
    if (n_groups == 0) return DYN_OK;
    DYN_ARG(d_off && d_p_e && d_p_c && d_alpha_beta_pi && d_status, "null buffer");   // triplets may be NULL when all groups are empty
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PiParams p;
    int rc = pi_tables(ctx, stream, &p.gl_x, &p.gl_w, &p.log_beta_fn, &p.ts_s, &p.ts_c, &p.ts_w);
    if (rc) return rc;
    if (max_cells < 1) max_cells = 1;
    max_cells = (max_cells + 1) & ~(int64_t)1;      // even: the tables behind the three cell arrays stay 16-byte aligned
    const size_t fixed = sizeof(double) * (2 * kNA * kChunk + 4 * (kNJ + 1));
    const size_t cap = ((ctx->smem_optin - 2048 - fixed) / (3 * sizeof(double))) & ~(size_t)1;
    if ((size_t)max_cells > cap) max_cells = (int64_t)cap;   // larger groups report status 3
    p.k = d_k; p.n = d_n;
    p.count = reinterpret_cast<const long long *>(d_count);
    p.off = reinterpret_cast<const long long *>(d_off);
    p.n_groups = n_groups; p.p_e = d_p_e; p.p_c = d_p_c;
    p.out = d_alpha_beta_pi; p.status = d_status; p.max_cells = (int)max_cells;
    const size_t smem = fixed + 3 * sizeof(double) * (size_t)max_cells;
    DYN_CUDA(cudaFuncSetAttribute(pi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    DYN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pi_kernel, kPiThreads, smem));
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)ctx->sm_count * per_sm;
    if (grid > n_groups) grid = n_groups;
    pi_kernel<<<(unsigned)grid, kPiThreads, smem, stream>>>(p);
    DYN_CUDA(cudaGetLastError());
    return DYN_OK;
}
