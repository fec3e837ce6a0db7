// Per-barcode binomial-mixture EM for p_c.
//
// Replaces dynast/estimation/p_c.py:32-172 (numba binomial_pmf / expectation_maximization /
// expectation_maximization_nasc) as driven one barcode per task by estimate_p_c (p_c.py:206-227).
//
// One CTA per barcode (persistent grid, barcodes handed out by an atomic counter).  The reference's
// arithmetic is reproduced operation for operation so that results are bit-identical, not merely
// within tolerance (the M step is a FLOAT32 sequential sum in the reference -- reordering it moves
// p_c by ~1e-6 relative, which is the whole parity budget):
//   * binomial coefficients come from a table built with wrapping 64-bit integer products and one
//     int64/int64 true division (p_c.py:45);
//   * p**k is numba's square-and-multiply (numba/cpython/numbers.py int_power_impl);
//   * every floating-point operation uses the explicit round-to-nearest intrinsics so that nvcc can
//     never contract a multiply-add;
//   * E step: unmasked cells never change, so columns are independent and run one per thread; each
//     column's sums run in the reference's kp order;
//   * M step: the two sequential float32 sums (row-major; cells that are zero for good are skipped,
//     x + 0 == x exactly) run as two dependency chains on two lanes of one warp, over a packed (k, n) list.
// The kernel is latency-bound on those serial chains, not HBM- or FP64-throughput-bound; parallelism
// comes from many barcodes resident per SM.
#include "common.cuh"

#include <algorithm>

namespace {

constexpr int kMaxFactorialK = 66;   // k! == 0 (mod 2^64) for k >= 66 -> ZeroDivisionError in the reference

struct EmParams {
    const int32_t *k;
    const int32_t *n;
    const long long *count;
    const long long *off;
    long long n_barcodes;
    const double *p_e;
    double *p_c;
    int32_t *status;
    int32_t *iters;
    const double *coef;   // [coef_K][coef_N]
    int coef_K, coef_N;
    int max_cells;        // dense capacity (cells) of the shared-memory layout
    int cls_lo, cls_hi;   // this launch handles barcodes with cls_lo < K * N <= cls_hi (size classes, see dyn_em_pc)
    int max_K, max_N;     // pow-table capacities
    unsigned long long *work_counter;
};

__global__ void em_coef_kernel(double *coef, int K, int N) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K * N) return;
    const long long k = i / N, n = i % N;
    unsigned long long num = 1, den = 1;
    for (long long j = 0; j < k; ++j) {
        num *= (unsigned long long)(j + (n - k + 1));
        den *= (unsigned long long)(j + 1);
    }
    coef[i] = __ddiv_rn((double)(long long)num, (double)(long long)den);
}

// per-barcode shape discovery: K = max k + 1, N = max n + 1; also global maxima
constexpr int kEmClasses = 3;
__device__ __host__ inline int em_class_of(long long cells, int c0, int c1) { return cells <= c0 ? 0 : (cells <= c1 ? 1 : 2); }

__global__ void em_shape_kernel(const int32_t *k, const int32_t *n, const long long *off, long long B, int *maxima, int c0,
                                int c1) {
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    int K = 0, N = 0;
    if (b < B) {
        for (long long i = off[b]; i < off[b + 1]; ++i) {
            K = max(K, k[i] + 1);
            N = max(N, n[i] + 1);
        }
    }
    // warp reduce then atomics
    long long cells = (long long)K * N;
    int c = cells > 0x7fffffffLL ? 0x7fffffff : (int)cells;
    if (b < B && K > 0) {       // per size class (few barcodes share a value: contention is not an issue next to the EM)
        const int cls = em_class_of(cells, c0, c1);
        atomicMax(maxima + 4 + 2 * cls, K);
        atomicMax(maxima + 5 + 2 * cls, N);
    }
    for (int d = 16; d; d >>= 1) {
        K = max(K, __shfl_xor_sync(0xffffffffu, K, d));
        N = max(N, __shfl_xor_sync(0xffffffffu, N, d));
        c = max(c, __shfl_xor_sync(0xffffffffu, c, d));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(maxima + 0, K);
        atomicMax(maxima + 1, N);
        atomicMax(maxima + 2, c);
    }
}

// numba int_power_impl: a ** b for float a, integer b. ok=false on 1.0 / 0.0.
__device__ __forceinline__ double ipow(double a, int b, bool &ok) {
    double r = 1.0;
    unsigned e = b < 0 ? (unsigned)(-b) : (unsigned)b;
    while (e) {
        if (e & 1u) r = __dmul_rn(r, a);
        e >>= 1;
        a = __dmul_rn(a, a);
    }
    if (b < 0) {
        if (r == 0.0) { ok = false; return 0.0; }
        r = __ddiv_rn(1.0, r);
    }
    return r;
}

struct EmShared {
    double p_c, prev;
    double bis_r, bis_l;
    int K, N, nact, ncols;
    int fail;
    int done;
};

// barrier of the CTA: a one-warp CTA (the small size class) never touches the hardware barrier
template <int THREADS>
__device__ __forceinline__ void cta_sync() {
    if (THREADS == 32) __syncwarp(); else __syncthreads();
}

template <bool NASC, int THREADS>
__global__ void __launch_bounds__(THREADS) em_kernel(const EmParams p) {
    constexpr int kEmThreads = THREADS;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    // layout (capacities from p.max_cells / p.max_K / p.max_N)
    float *v = reinterpret_cast<float *>(smem_raw);                         // [max_cells]
    uint32_t *cnt = reinterpret_cast<uint32_t *>(v + p.max_cells);          // [max_cells]; later: active list
    double *powp = reinterpret_cast<double *>(cnt + p.max_cells);           // [max_K]
    double *powq = powp + p.max_K;                                          // [max_K + max_N]
    double *den_n = powq + p.max_K + p.max_N;                               // [max_N] E-step denominators
    int *cols = reinterpret_cast<int *>(den_n + 2 * p.max_N);               // [max_N] columns with masked cells
    unsigned long long *nzu = reinterpret_cast<unsigned long long *>(den_n + p.max_N);   // [max_N] rows < 64 of the column that
                                                                                         // are unmasked and non-zero
    uint8_t *mask = reinterpret_cast<uint8_t *>(cols + p.max_N);            // [max_cells]
    __shared__ EmShared sh;
    __shared__ __align__(16) double2 s_prod[32];     // M-step staging: one chunk of 32 (k, n) product pairs
    __shared__ unsigned long long s_work;
    __shared__ int s_scan[kEmThreads / 32];

    const int tid = threadIdx.x;
    for (;;) {
        cta_sync<THREADS>();
        if (tid == 0) s_work = atomicAdd(p.work_counter, 1ull);
        cta_sync<THREADS>();
        const long long b = (long long)s_work;
        if (b >= p.n_barcodes) break;

        const long long t0 = p.off[b], t1 = p.off[b + 1];
        const double p_e = p.p_e[b];
        // ---- shape
        int K = 0, N = 0;
        for (long long i = t0 + tid; i < t1; i += kEmThreads) {
            K = max(K, p.k[i] + 1);
            N = max(N, p.n[i] + 1);
        }
        for (int d = 16; d; d >>= 1) {
            K = max(K, __shfl_xor_sync(0xffffffffu, K, d));
            N = max(N, __shfl_xor_sync(0xffffffffu, N, d));
        }
        if ((tid & 31) == 0) { s_scan[tid >> 5] = K; }
        cta_sync<THREADS>();
        if (tid == 0) {
            int kk = 0;
            for (int w = 0; w < kEmThreads / 32; ++w) kk = max(kk, s_scan[w]);
            sh.K = kk;
        }
        cta_sync<THREADS>();
        if ((tid & 31) == 0) { s_scan[tid >> 5] = N; }
        cta_sync<THREADS>();
        if (tid == 0) {
            int nn = 0;
            for (int w = 0; w < kEmThreads / 32; ++w) nn = max(nn, s_scan[w]);
            sh.N = nn;
            sh.fail = 0;
            sh.done = 0;
        }
        cta_sync<THREADS>();
        K = sh.K; N = sh.N;
        const int KN = K * N;
        if (KN <= p.cls_lo || KN > p.cls_hi) continue;      // another launch's size class
        if (K == 0 || N == 0) {
            if (tid == 0) { p.p_c[b] = 0.0; p.status[b] = DYN_EM_PC_LE_PE; if (p.iters) p.iters[b] = 0; }
            continue;
        }
        if (K > kMaxFactorialK) {   // binomial_pmf(k >= 66, ..) divides by a wrapped-to-zero factorial
            if (tid == 0) { p.p_c[b] = 0.0; p.status[b] = DYN_EM_ZERODIV; if (p.iters) p.iters[b] = 0; }
            continue;
        }
        if (KN > p.max_cells || K > p.max_K || N > p.max_N || K > p.coef_K || N > p.coef_N) {
            if (tid == 0) { p.p_c[b] = 0.0; p.status[b] = DYN_EM_UNSUPPORTED; if (p.iters) p.iters[b] = 0; }
            continue;
        }
        // ---- dense histogram (duplicates summed, p_c.py:219)
        for (int i = tid; i < KN; i += kEmThreads) { cnt[i] = 0; mask[i] = 0; }
        cta_sync<THREADS>();
        for (long long i = t0 + tid; i < t1; i += kEmThreads) {
            const long long c = p.count[i];
            if (c < 0 || c > 0x7fffffffLL) sh.fail = DYN_EM_UNSUPPORTED;
            atomicAdd(&cnt[p.k[i] * N + p.n[i]], (uint32_t)c);
        }
        // pow tables for p_e
        {
            bool ok = true;
            for (int i = tid; i < K; i += kEmThreads) powp[i] = ipow(p_e, i, ok);
            const double q = __dsub_rn(1.0, p_e);
            for (int i = tid; i < K + N - 1; i += kEmThreads) powq[i] = ipow(q, i - (K - 1), ok);
            if (!ok) sh.fail = DYN_EM_ZERODIV;
        }
        cta_sync<THREADS>();
        // ---- mask (p_c.py:138-143 / 64-68): one thread per column, suffix sums from the top row down
        for (int n = tid; n < N; n += kEmThreads) {
            long long s = 0;
            bool any = false;
            unsigned long long nz = 0ull;
            for (int k = K - 1; k >= 0; --k) {
                const uint32_t c = cnt[k * N + n];
                if (!NASC) s += c;
                const double pm = __dmul_rn(__dmul_rn(p.coef[k * p.coef_N + n], powp[k]), powq[n - k + K - 1]);
                const double expected = __dmul_rn((double)s, pm);
                const bool m = expected > __dmul_rn(0.01, (double)c);
                mask[k * N + n] = m;
                any |= m;
                if (!m && c != 0 && k < 64) nz |= 1ull << k;
                if (NASC) s += c;
            }
            nzu[n] = nz;
            cols[n] = any ? 1 : 0;
        }
        cta_sync<THREADS>();
        for (int i = tid; i < KN; i += kEmThreads) v[i] = (float)cnt[i];   // p_c.py:148
        cta_sync<THREADS>();
        // ---- active list: row-major indices of cells that can ever be non-zero (non-zero or masked);
        //      compacted list of columns owning masked cells.  Warp 0 compacts 32 entries at a time by ballot, in
        //      place: an entry lands at or before its own index, and a chunk is read completely before it is written.
        if (tid < 32) {
            int na = 0;
            for (int i0 = 0; i0 < KN; i0 += 32) {
                const int i = i0 + tid;
                const bool keep = i < KN && (cnt[i] != 0 || mask[i]);
                const uint32_t packed = ((uint32_t)(i / N) << 16) | (uint32_t)(i % N);
                const uint32_t bal = __ballot_sync(0xffffffffu, keep);
                __syncwarp();
                if (keep) cnt[na + __popc(bal & ((1u << tid) - 1u))] = packed;
                na += __popc(bal);
                __syncwarp();
            }
            int nc = 0;
            for (int n0 = 0; n0 < N; n0 += 32) {
                const int n = n0 + tid;
                const bool keep = n < N && cols[n] != 0;
                const uint32_t bal = __ballot_sync(0xffffffffu, keep);
                __syncwarp();
                if (keep) cols[nc + __popc(bal & ((1u << tid) - 1u))] = n;
                nc += __popc(bal);
                __syncwarp();
            }
            if (tid == 0) {
                sh.nact = na;
                sh.ncols = nc;
                sh.p_c = NASC ? 0.5 : 0.1;
                sh.prev = sh.p_c;
                sh.bis_r = 1.0;
                sh.bis_l = 0.0;
            }
        }
        cta_sync<THREADS>();
        const int nact = sh.nact, ncols = sh.ncols;
        const uint32_t *act = cnt;   // reused as the active list

        int iters = 0;
        const int max_iters = NASC ? 0x7fffffff : 300;
        while (!sh.fail && !sh.done && iters < max_iters) {
            if (NASC && !(__dsub_rn(sh.bis_r, sh.bis_l) >= 1e-8)) break;
            ++iters;
            const double p_c = sh.p_c;
            {
                bool ok = true;
                for (int i = tid; i < K; i += kEmThreads) powp[i] = ipow(p_c, i, ok);
                const double q = __dsub_rn(1.0, p_c);
                for (int i = tid; i < K + N - 1; i += kEmThreads) powq[i] = ipow(q, i - (K - 1), ok);
                if (!ok) sh.fail = DYN_EM_ZERODIV;
            }
            cta_sync<THREADS>();
            // ---- E step.  Unmasked cells never change, so (non-NASC) every masked cell can be updated
            //      independently: first one thread per column for the denominators, then one thread per masked
            //      cell for its numerator -- a column full of masked cells (k up to 30 at n = 150) no longer
            //      serialises on one thread.  Each sum runs in the reference's kp order.
            if (!NASC) {
                for (int ci = tid; ci < ncols; ci += kEmThreads) {
                    const int n = cols[ci];
                    const double *cf = p.coef + n;
                    double den = 0.0;
                    for (int kp = 0; kp < K; ++kp)
                        if (!mask[kp * N + n])
                            den = __dadd_rn(den, __dmul_rn(__dmul_rn(cf[kp * p.coef_N], powp[kp]), powq[n - kp + K - 1]));
                    den_n[n] = den;
                }
                cta_sync<THREADS>();
                for (int i = tid; i < nact; i += kEmThreads) {
                    const uint32_t a = act[i];
                    const int k = (int)(a >> 16), n = (int)(a & 0xffffu);
                    if (!mask[k * N + n]) continue;
                    const double den = den_n[n];
                    if (!(den > 0.0)) continue;
                    const double pk = __dmul_rn(__dmul_rn(p.coef[k * p.coef_N + n], powp[k]), powq[n - k + K - 1]);
                    double num = 0.0;
                    if (K <= 64 && isfinite(pk)) {
                        // unmasked cells keep their counts: only the non-zero ones contribute (pk * 0 == 0 and
                        // num + 0 == num exactly), lowest row first as in the reference
                        unsigned long long m = nzu[n];
                        while (m) {
                            const int kp = __ffsll((long long)m) - 1;
                            m &= m - 1;
                            num = __dadd_rn(num, __dmul_rn(pk, (double)v[kp * N + n]));
                        }
                    } else {
                        for (int kp = 0; kp < K; ++kp)
                            if (!mask[kp * N + n]) num = __dadd_rn(num, __dmul_rn(pk, (double)v[kp * N + n]));
                    }
                    v[k * N + n] = __double2float_rn(__ddiv_rn(num, den));
                }
            }
            for (int ci = tid; NASC && ci < ncols; ci += kEmThreads) {
                const int n = cols[ci];
                const double *cf = p.coef + n;
                {
                    // Gauss-Seidel within the column, sums over kp IN the mask (p_c.py:80-88, as written):
                    // a masked cell reads cells updated earlier in the same sweep, so the column stays serial
                    double den = 0.0;
                    for (int kp = 0; kp < K; ++kp)
                        if (mask[kp * N + n])
                            den = __dadd_rn(den, __dmul_rn(__dmul_rn(cf[kp * p.coef_N], powp[kp]), powq[n - kp + K - 1]));
                    if (den == 0.0) { sh.fail = DYN_EM_ZERODIV; continue; }
                    for (int k = 0; k < K; ++k) {
                        if (!mask[k * N + n]) continue;
                        const double pk = __dmul_rn(__dmul_rn(cf[k * p.coef_N], powp[k]), powq[n - k + K - 1]);
                        double num = 0.0;
                        for (int kp = 0; kp < K; ++kp)
                            if (mask[kp * N + n]) num = __dadd_rn(num, __dmul_rn(pk, (double)v[kp * N + n]));
                        v[k * N + n] = __double2float_rn(__ddiv_rn(num, den));
                    }
                }
            }
            cta_sync<THREADS>();
            // ---- M step: two sequential sums in the reference's cell order (row-major; cells that are zero for
            //      good are skipped, x + 0 == x exactly).  The sums must stay sequential (float32, p_c.py:162-165:
            //      reordering moves p_c by ~1e-6), but the loads and products need not: warp 0 takes 32 active cells
            //      at a time, every lane loads and multiplies one cell, then the 32 products are folded into the
            //      running sums in order out of a 32-entry shared-memory buffer -- the dependency chain is one add per cell.
            if (tid < 32) {
                if (!NASC) {
                    // numba: (new_values * arange).sum() == float32 products, float32 sequential sum
                    float sk = 0.0f, sn = 0.0f;
                    float2 *prod = reinterpret_cast<float2 *>(s_prod);
                    for (int base = 0; base < nact; base += 32) {
                        const int i = base + tid;
                        float pk = 0.0f, pn = 0.0f;
                        if (i < nact) {
                            const uint32_t a = act[i];
                            const uint32_t kk = a >> 16, nn = a & 0xffffu;
                            const float x = v[kk * N + nn];
                            pk = __fmul_rn(x, (float)kk);
                            pn = __fmul_rn(x, (float)nn);
                        }
                        // the 32 product pairs go through shared memory: one broadcast 128-bit load feeds four
                        // adds of the two chains (a shuffle per add before).  Entries past the end are +0: x + 0 == x.
                        prod[tid] = make_float2(pk, pn);
                        __syncwarp();
                        const int groups = (min(32, nact - base) + 7) >> 3;
                        for (int g = 0; g < groups; ++g) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float4 q = reinterpret_cast<const float4 *>(prod)[g * 4 + j];
                                sk = __fadd_rn(sk, q.x); sn = __fadd_rn(sn, q.y);
                                sk = __fadd_rn(sk, q.z); sn = __fadd_rn(sn, q.w);
                            }
                        }
                        __syncwarp();
                    }
                    if (tid == 0) {
                        const double prev = sh.p_c;
                        const double nw = sn > 0.0f ? (double)__fdiv_rn(sk, sn) : prev;
                        sh.prev = prev;
                        sh.p_c = nw;
                        if (prev == nw) sh.done = 1;
                    }
                } else {
                    double sk = 0.0, sn = 0.0;
                    double2 *prod = s_prod;
                    for (int base = 0; base < nact; base += 32) {
                        const int i = base + tid;
                        double pk = 0.0, pn = 0.0;
                        if (i < nact) {
                            const uint32_t a = act[i];
                            const uint32_t kk = a >> 16, nn = a & 0xffffu;
                            const double x = (double)v[kk * N + nn];
                            pk = __dmul_rn((double)kk, x);
                            pn = __dmul_rn((double)nn, x);
                        }
                        prod[tid] = make_double2(pk, pn);
                        __syncwarp();
                        const int cnt32 = min(32, nact - base);
#pragma unroll 8
                        for (int j = 0; j < cnt32; ++j) {
                            const double2 q = prod[j];
                            sk = __dadd_rn(sk, q.x);
                            sn = __dadd_rn(sn, q.y);
                        }
                        __syncwarp();
                    }
                    if (tid == 0) {
                        const double prev = sh.p_c;
                        const double nw = sn > 0.0 ? __ddiv_rn(sk, sn) : 0.0;
                        sh.prev = prev;
                        sh.p_c = nw;
                        if (prev == nw) sh.done = 1;
                        else if (nw < prev) sh.bis_r = prev;
                        else sh.bis_l = prev;
                    }
                }
            }
            cta_sync<THREADS>();
        }
        if (tid == 0) {
            int st = sh.fail;
            if (!st && !NASC && sh.p_c <= p_e) st = DYN_EM_PC_LE_PE;
            p.p_c[b] = sh.p_c;
            p.status[b] = st;
            if (p.iters) p.iters[b] = iters;
        }
    }
}

}  // namespace

static size_t em_smem_bytes(int max_cells, int max_K, int max_N) {
    size_t s = 0;
    s += sizeof(float) * max_cells + sizeof(uint32_t) * max_cells;   // v, cnt/active list
    s += sizeof(double) * (max_K + max_K + max_N + max_N) + sizeof(unsigned long long) * max_N;
    s += sizeof(int) * max_N;
    s += max_cells;
    return (s + 15) & ~(size_t)15;
}

extern "C" int dyn_em_pc(dyn_ctx_t *ctx, const int32_t *d_k, const int32_t *d_n, const int64_t *d_count,
                         const int64_t *d_off, int64_t n_barcodes, const double *d_p_e, int nasc, double *d_p_c,
                         int32_t *d_em_status, int32_t *d_iters, void *stream_) {
    DYN_ARG(ctx != nullptr, "null context");
    CtxEnter enter_(ctx, stream_);
    DYN_ARG(n_barcodes >= 0, "negative n_barcodes");
    if (n_barcodes == 0) return DYN_OK;
    DYN_ARG(d_off && d_p_e && d_p_c && d_em_status, "null buffer");   // triplet pointers may be NULL when every barcode is empty
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);

    // shape discovery (one small kernel + a 40-byte read back: the only synchronisation of this call).
    // Size classes by dense cells K * N: <= kCells0 (one warp per barcode), <= kCells1 (two warps), rest (eight warps).
    const int kCells0 = 512, kCells1 = 2048;
    void *scr = nullptr;
    int rc = dyn_ctx_scratch(ctx, 0, 128, &scr);
    if (rc) return rc;
    int *d_max = static_cast<int *>(scr);
    unsigned long long *d_work = reinterpret_cast<unsigned long long *>(d_max + 16);
    DYN_CUDA(cudaMemsetAsync(scr, 0, 128, stream));
    em_shape_kernel<<<(unsigned)((n_barcodes + 127) / 128), 128, 0, stream>>>(
        d_k, d_n, reinterpret_cast<const long long *>(d_off), n_barcodes, d_max, kCells0, kCells1);
    DYN_CUDA(cudaGetLastError());
    int h_max[10];
    DYN_CUDA(cudaMemcpyAsync(h_max, d_max, sizeof(h_max), cudaMemcpyDeviceToHost, stream));
    DYN_CUDA(cudaStreamSynchronize(stream));
    int max_K = h_max[0] < 1 ? 1 : h_max[0], max_N = h_max[1] < 1 ? 1 : h_max[1];
    long long max_cells = h_max[2] < 1 ? 1 : h_max[2];
    if (max_K > kMaxFactorialK) max_K = kMaxFactorialK;   // larger K fail in-kernel with DYN_EM_ZERODIV

    // shrink the dense capacity until the layout fits; barcodes beyond it report DYN_EM_UNSUPPORTED
    while (em_smem_bytes((int)max_cells, max_K, max_N) > ctx->smem_optin - 1024 && max_cells > 64) max_cells /= 2;
    if (em_smem_bytes((int)max_cells, max_K, max_N) > ctx->smem_optin - 1024) {
        dyn_set_error("dyn_em_pc: histogram shape %d x %d does not fit shared memory", max_K, max_N);
        return DYN_ERR_ARG;
    }

    // coefficient table
    if (!ctx->em_coef || ctx->em_coef_K < max_K || ctx->em_coef_N < max_N) {
        if (ctx->em_coef) DYN_CUDA(cudaFree(ctx->em_coef));
        ctx->em_coef = nullptr;
        const int K = max_K > 32 ? max_K : 32, N = max_N > 160 ? max_N : 160;
        DYN_CUDA(cudaMalloc(&ctx->em_coef, sizeof(double) * (size_t)K * N));
        ctx->em_coef_K = K;
        ctx->em_coef_N = N;
        em_coef_kernel<<<(K * N + 255) / 256, 256, 0, stream>>>(ctx->em_coef, K, N);
        DYN_CUDA(cudaGetLastError());
    }

    EmParams p;
    p.k = d_k; p.n = d_n;
    p.count = reinterpret_cast<const long long *>(d_count);
    p.off = reinterpret_cast<const long long *>(d_off);
    p.n_barcodes = n_barcodes;
    p.p_e = d_p_e; p.p_c = d_p_c; p.status = d_em_status; p.iters = d_iters;
    p.coef = ctx->em_coef; p.coef_K = ctx->em_coef_K; p.coef_N = ctx->em_coef_N;

    // The kernel is latency-bound per barcode (a few dependent fp64 / fp32 chains per iteration), so throughput
    // comes from resident barcodes per SM, and those are limited by the dense shared-memory layout.  One rare
    // large histogram (k = 30, n = 150 -> 4 681 cells, 45 KB) would size every CTA: barcodes are therefore run in
    // three size classes, each with a layout (cells, pow-table lengths) sized by its own largest member:
    //   <= 512 cells   one warp per barcode -- no CTA barrier at all, up to 32 barcodes per SM;
    //   <= 2048 cells  two warps;
    //   larger         eight warps: the E step is the long part there.
    // Every launch walks all barcodes (an atomic counter hands them out) and skips the other classes.
    // (Running the two larger classes on side streams next to the small one was measured and is slower: the small
    // class is issue-bound, 36.4 ms against 35.3 ms back to back for the C4 sweep.)
    for (int c = kEmClasses - 1; c >= 0; --c) {
        int cls_K = h_max[4 + 2 * c], cls_N = h_max[5 + 2 * c];
        if (cls_K < 1 && c > 0) continue;                          // nobody in this class
        if (cls_K < 1) cls_K = cls_N = 1;                          // class 0 also reports the barcodes without reads
        if (cls_K > kMaxFactorialK) cls_K = kMaxFactorialK;
        long long cap = c == 0 ? kCells0 : (c == 1 ? kCells1 : max_cells);
        cap = std::min<long long>(cap, std::min<long long>(max_cells, (long long)cls_K * cls_N));
        p.max_cells = (int)cap;
        p.max_K = cls_K; p.max_N = cls_N;
        p.cls_lo = c == 0 ? -1 : (c == 1 ? kCells0 : kCells1);
        p.cls_hi = c == 0 ? kCells0 : (c == 1 ? kCells1 : 0x7fffffff);
        p.work_counter = d_work + c;
        const size_t smem = em_smem_bytes((int)cap, cls_K, cls_N);
        const int threads = c == 0 ? 32 : (c == 1 ? 64 : 256);
        void (*kern)(const EmParams) =
            c == 0 ? (nasc ? em_kernel<true, 32> : em_kernel<false, 32>)
                   : (c == 1 ? (nasc ? em_kernel<true, 64> : em_kernel<false, 64>)
                             : (nasc ? em_kernel<true, 256> : em_kernel<false, 256>));
        DYN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        DYN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
        if (per_sm < 1) per_sm = 1;
        long long grid = (long long)ctx->sm_count * per_sm;
        if (grid > n_barcodes) grid = n_barcodes;
        kern<<<(unsigned)grid, threads, smem, stream>>>(p);
        DYN_CUDA(cudaGetLastError());
    }
    return DYN_OK;
}
