// Per-group reductions between counting and estimation.
//
// Replaces the pandas group-bys of
//   dynast/preprocessing/aggregation.py:61-89   calculate_mutation_rates  (16-column sums -> 12 rates)
//   dynast/preprocessing/aggregation.py:92-119  aggregate_counts          ((barcode, GX, k, n) -> reads)
//   dynast/estimation/p_e.py:52-108             estimate_p_e              (mean background rate)
//   dynast/estimation/alpha.py:26-95            estimate_alpha            (fraction of reads with a conversion)
// All of them read the 16-byte counter rows of the surviving reads once (HBM-bound, 16 B counters + 8 B
// ids per read in; per-group tables out).
#include <cub/cub.cuh>

#include <math.h>

#include "common.cuh"

#include <algorithm>

namespace {

constexpr int kBlock = 256;
inline unsigned grid_for(long long n) { return (unsigned)((n + kBlock - 1) / kBlock); }

__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }
__device__ __forceinline__ uint4 complement_row(uint4 v) {
    uint4 o;
    o.x = bswap32(v.z); o.y = bswap32(v.y); o.z = bswap32(v.x); o.w = bswap32(v.w);
    return o;
}
__device__ __forceinline__ uint4 load_row(const uint4 *counts, const uint8_t *strand, long long r) {
    uint4 c = counts[r];
    if (strand && strand[r]) c = complement_row(c);
    return c;
}
__device__ __forceinline__ uint32_t col(const uint4 &c, int j) {
    const uint32_t w = j < 4 ? c.x : (j < 8 ? c.y : (j < 12 ? c.z : c.w));
    return (w >> ((j & 3) << 3)) & 0xffu;
}

// ---- a9: 16-column sums per group.  Surviving reads come in split-file order: a stretch of a few thousand rows
// holds a few dozen barcodes.  Every CTA therefore sums its stretch into a small shared-memory table (open
// addressing on the group id, 32-bit shared atomics) and flushes the occupied slots with one 64-bit L2 atomic per
// non-zero counter; a row whose group finds no slot goes straight to L2.
constexpr int kSumSlots = 128;          // table slots per CTA
constexpr int kSumRowsPerThread = 8;    // rows per thread: kBlock * 8 = 2 048 rows per CTA

__global__ void __launch_bounds__(kBlock) group_sums_kernel(const uint4 *counts, const uint8_t *strand, const uint32_t *group,
                                                            const uint32_t *sel, long long n_sel, long long n_groups,
                                                            unsigned long long *sums, unsigned long long *n_reads,
                                                            unsigned long long *n_with, uint32_t conv_mask) {
    __shared__ uint32_t s_key[kSumSlots];
    __shared__ uint32_t s_val[kSumSlots][18];       // 16 columns, reads, reads with a conversion
    for (int i = threadIdx.x; i < kSumSlots; i += kBlock) s_key[i] = 0xffffffffu;
    for (int i = threadIdx.x; i < kSumSlots * 18; i += kBlock) (&s_val[0][0])[i] = 0u;
    __syncthreads();
    const long long base = (long long)blockIdx.x * kBlock * kSumRowsPerThread;
    for (int it = 0; it < kSumRowsPerThread; ++it) {
        const long long i = base + (long long)it * kBlock + threadIdx.x;
        if (i >= n_sel) break;
        const long long r = sel ? (long long)sel[i] : i;
        const uint32_t g = group ? group[r] : 0u;
        if ((long long)g >= n_groups) continue;
        const uint4 c = load_row(counts, strand, r);
        uint32_t k = 0;
#pragma unroll
        for (int j = 0; j < 12; ++j)
            if ((conv_mask >> j) & 1u) k += col(c, j);
        // slot of the group (claimed on first use)
        int slot = -1;
        uint32_t h = (g * 0x9E3779B1u) >> 25;
        for (int probe = 0; probe < 8; ++probe) {
            const int sidx = (int)((h + probe) & (kSumSlots - 1));
            const uint32_t cur = s_key[sidx];
            if (cur == g) { slot = sidx; break; }
            if (cur == 0xffffffffu) {
                const uint32_t old = atomicCAS(&s_key[sidx], 0xffffffffu, g);
                if (old == 0xffffffffu || old == g) { slot = sidx; break; }
            }
        }
        if (slot >= 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const uint32_t v = col(c, j);
                if (v) atomicAdd(&s_val[slot][j], v);
            }
            atomicAdd(&s_val[slot][16], 1u);
            if (k > 0) atomicAdd(&s_val[slot][17], 1u);
        } else {
            unsigned long long *row = sums + (size_t)g * 16;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const uint32_t v = col(c, j);
                if (v) atomicAdd(row + j, (unsigned long long)v);
            }
            if (n_reads) atomicAdd(n_reads + g, 1ull);
            if (n_with && k > 0) atomicAdd(n_with + g, 1ull);
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < kSumSlots * 18; e += kBlock) {
        const int slot = e / 18, j = e % 18;
        const uint32_t g = s_key[slot], v = s_val[slot][j];
        if (g == 0xffffffffu || v == 0u) continue;
        if (j < 16) atomicAdd(sums + (size_t)g * 16 + j, (unsigned long long)v);
        else if (j == 16) { if (n_reads) atomicAdd(n_reads + g, (unsigned long long)v); }
        else if (n_with) atomicAdd(n_with + g, (unsigned long long)v);
    }
}

// ---- a9 / a11: rates and p_e from the sums (uint32 wrap as in `.astype(np.uint32)`, aggregation.py:76)
__global__ void rates_kernel(const unsigned long long *sums, long long n_groups, uint32_t pe_mask, double *rates,
                             double *p_e) {
    const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const unsigned long long *row = sums + (size_t)g * 16;
    double acc = 0.0;
    int cnt = 0;
    for (int j = 0; j < 12; ++j) {
        const double num = (double)(uint32_t)row[j], den = (double)(uint32_t)row[12 + j / 3];
        const double r = num / den;   // 0/0 -> NaN, x/0 -> inf: pandas semantics
        if (rates) rates[(size_t)g * 12 + j] = r;
        if (((pe_mask >> j) & 1u) && !isnan(r)) { acc += r; ++cnt; }   // DataFrame.mean skips NaN
    }
    if (p_e) p_e[g] = cnt ? acc / (double)cnt : nan("");
}

// ---- a10: (group, gene, k, n) keys
struct KnIn {
    const uint4 *counts;
    const uint8_t *strand;
    const uint32_t *group, *gene, *sel;
    long long n_sel;
    uint32_t conv_mask, base_mask, exclude_mask;
    int gene_bits;
};

__global__ void kn_key_kernel(KnIn in, unsigned long long *key, uint32_t *pos) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= in.n_sel) return;
    const long long r = in.sel ? (long long)in.sel[i] : i;
    const uint4 c = load_row(in.counts, in.strand, r);
    uint32_t k = 0, n = 0, other = 0;
#pragma unroll
    for (int j = 0; j < 12; ++j) {
        const uint32_t v = col(c, j);
        if ((in.conv_mask >> j) & 1u) k += v;
        if ((in.exclude_mask >> j) & 1u) other += v;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if ((in.base_mask >> j) & 1u) n += col(c, 12 + j);
    pos[i] = (uint32_t)i;
    if (other != 0) { key[i] = ~0ull; return; }   // estimate.py:248-255: reads with another labelled conversion
    const unsigned long long g = in.group ? in.group[r] : 0u, gx = in.gene ? in.gene[r] : 0u;
    key[i] = (((g << in.gene_bits) | gx) << 22) | ((unsigned long long)k << 10) | n;
}

__global__ void head_flag_kernel(const unsigned long long *key, uint8_t *flag, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    flag[i] = (key[i] != ~0ull) && (i == 0 || key[i] != key[i - 1]);
}

__global__ void iota_kernel(uint32_t *v, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) v[i] = (uint32_t)i;
}

// run lengths from the positions of the run heads; the tail run ends at the first sentinel key
__global__ void run_count_kernel(const unsigned long long *key, const uint32_t *heads, const long long *n_runs, long long n,
                                 uint32_t *count) {
    const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= *n_runs) return;
    long long end;
    if (j + 1 < *n_runs) end = heads[j + 1];
    else {
        long long lo = heads[j], hi = n;     // first index >= heads[j] holding the sentinel
        while (lo < hi) { const long long mid = (lo + hi) >> 1; if (key[mid] == ~0ull) hi = mid; else lo = mid + 1; }
        end = lo;
    }
    count[j] = (uint32_t)(end - heads[j]);
}

__global__ void gather_rows_kernel(const unsigned long long *key, const uint32_t *pos, const uint32_t *heads,
                                   const uint32_t *count, const uint32_t *perm, const long long *n_runs,
                                   unsigned long long *out_key, uint32_t *out_count, uint32_t *out_first) {
    const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= *n_runs) return;
    const uint32_t s = perm ? perm[j] : (uint32_t)j;
    out_key[j] = key[heads[s]];
    out_count[j] = count[s];
    out_first[j] = pos[heads[s]];
}

__global__ void pad_first_kernel(const uint32_t *pos, const uint32_t *heads, const long long *n_runs, long long n, uint32_t *first) {
    const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= n) return;
    first[j] = j < *n_runs ? pos[heads[j]] : 0xffffffffu;
}

// ---- CSR for the EM: rows (barcode, k, n, count) sorted by key -> triplets + per-barcode offsets / totals
__global__ void kn_csr_kernel(const unsigned long long *key, const uint32_t *count, const long long *n_rows_p, long long n_groups,
                              int32_t *k, int32_t *n, long long *cnt, long long *off, unsigned long long *total) {
    const long long n_rows = *n_rows_p;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n_rows) {
        const unsigned long long kk = key[i];
        k[i] = (int32_t)((kk >> 10) & 0xfff);
        n[i] = (int32_t)(kk & 0x3ff);
        cnt[i] = (long long)count[i];
        const unsigned long long g = kk >> 22;
        if ((long long)g < n_groups) atomicAdd(total + g, (unsigned long long)count[i]);
    }
    if (i <= n_groups) {
        // off[g] = first row whose group >= g
        long long lo = 0, hi = n_rows;
        while (lo < hi) { const long long mid = (lo + hi) >> 1; if ((long long)(key[mid] >> 22) < i) lo = mid + 1; else hi = mid; }
        off[i] = lo;
    }
}

// ---- CSR over the (group, gene) pairs that OCCUR (the per-(cell, gene) fit of pi.py:253-296 with group_by = [barcode,
// GX]): rows sorted by key, a pair = the key without its 22 (k, n) bits.  Pair j gets rows [off[j], off[j + 1]).
__global__ void pair_head_kernel(const unsigned long long *key, const long long *n_rows_p, long long max_rows, uint32_t *head) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= max_rows) return;
    head[i] = (i < *n_rows_p && (i == 0 || (key[i] >> 22) != (key[i - 1] >> 22))) ? 1u : 0u;
}

__global__ void pair_fill_kernel(const unsigned long long *key, const uint32_t *count, const long long *n_rows_p, const uint32_t *head,
                                 const uint32_t *pair_of_row, long long max_rows, int32_t *k, int32_t *n, long long *cnt, long long *off,
                                 unsigned long long *pair_key, unsigned long long *total, long long *n_pairs) {
    const long long n_rows = *n_rows_p;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n_rows) {
        const unsigned long long kk = key[i];
        k[i] = (int32_t)((kk >> 10) & 0xfff);
        n[i] = (int32_t)(kk & 0x3ff);
        cnt[i] = (long long)count[i];
        const uint32_t p = pair_of_row[i] - 1u;      // inclusive head count -> 0-based pair number
        if (head[i]) { off[p] = i; pair_key[p] = kk >> 22; }
        atomicAdd(total + p, (unsigned long long)count[i]);
        if (i == n_rows - 1) { off[p + 1] = n_rows; *n_pairs = (long long)p + 1; }
    }
    if (i == 0 && n_rows == 0) { off[0] = 0; *n_pairs = 0; }
}

// ---- a16: alpha = (reads with a conversion / reads) / pi_c   (alpha.py:72-84)
__global__ void alpha_kernel(const unsigned long long *n_reads, const unsigned long long *n_with, const double *pi_c,
                             long long n_groups, double *alpha) {
    const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const double p = pi_c[g];
    alpha[g] = (p > 0.0 && n_reads[g] > 0) ? ((double)n_with[g] / (double)n_reads[g]) / p : nan("");
}

}  // namespace

extern "C" int dyn_group_sums(dyn_ctx_t *ctx, const uint8_t *d_counts, const uint8_t *d_strand, const uint32_t *d_group,
                              const uint32_t *d_sel, int64_t n_sel, int64_t n_groups, uint32_t conv_mask,
                              uint64_t *d_sums, uint64_t *d_n_reads, uint64_t *d_n_with, void *stream_) {
    DYN_ARG(ctx != nullptr, "null context");
    CtxEnter enter_(ctx, stream_);
    DYN_ARG(n_sel >= 0 && n_groups >= 0, "negative size");
    if (n_groups == 0) return DYN_OK;
    DYN_ARG(d_sums != nullptr, "null output");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DYN_CUDA(cudaMemsetAsync(d_sums, 0, sizeof(uint64_t) * 16 * n_groups, stream));
    if (d_n_reads) DYN_CUDA(cudaMemsetAsync(d_n_reads, 0, sizeof(uint64_t) * n_groups, stream));
    if (d_n_with) DYN_CUDA(cudaMemsetAsync(d_n_with, 0, sizeof(uint64_t) * n_groups, stream));
    if (n_sel == 0) return DYN_OK;
    DYN_ARG(d_counts != nullptr, "null counts");
    group_sums_kernel<<<(unsigned)((n_sel + kBlock * kSumRowsPerThread - 1) / (kBlock * kSumRowsPerThread)), kBlock, 0, stream>>>(
        reinterpret_cast<const uint4 *>(d_counts), d_strand, d_group, d_sel, n_sel, n_groups,
        reinterpret_cast<unsigned long long *>(d_sums), reinterpret_cast<unsigned long long *>(d_n_reads),
        reinterpret_cast<unsigned long long *>(d_n_with), conv_mask & 0xfffu);
    DYN_CUDA(cudaGetLastError());
    return DYN_OK;
}

extern "C" int dyn_rates_pe(dyn_ctx_t *ctx, const uint64_t *d_sums, int64_t n_groups, uint32_t pe_mask, double *d_rates,
                            double *d_p_e, void *stream) {
    DYN_ARG(ctx != nullptr, "null context");
    CtxEnter enter_(ctx, stream);
    if (n_groups <= 0) return DYN_OK;
    DYN_ARG(d_sums != nullptr, "null sums");
    rates_kernel<<<grid_for(n_groups), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const unsigned long long *>(d_sums), n_groups, pe_mask & 0xfffu, d_rates, d_p_e);
    DYN_CUDA(cudaGetLastError());
    return DYN_OK;
}

extern "C" int dyn_aggregate_kn(dyn_ctx_t *ctx, const uint8_t *d_counts, const uint8_t *d_strand, const uint32_t *d_group,
                                const uint32_t *d_gene, const uint32_t *d_sel, int64_t n_sel, uint32_t conv_mask,
                                uint32_t base_mask, uint32_t exclude_mask, int group_bits, int gene_bits,
                                int first_appearance_order, uint64_t *d_keys, uint32_t *d_count, uint32_t *d_first,
                                int64_t *d_n_rows, void *stream_) {
    DYN_ARG(ctx != nullptr, "null context");
    CtxEnter enter_(ctx, stream_);
    DYN_ARG(n_sel >= 0 && n_sel < 0xffffffffLL, "n_sel out of range");
    DYN_ARG(d_n_rows != nullptr, "null n_rows");
    DYN_ARG(group_bits >= 0 && gene_bits >= 0 && group_bits + gene_bits <= 41, "group_bits + gene_bits must be <= 41");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DYN_CUDA(cudaMemsetAsync(d_n_rows, 0, sizeof(int64_t), stream));
    if (n_sel == 0) return DYN_OK;
    DYN_ARG(d_counts && d_keys && d_count && d_first, "null buffer");
    const size_t n = (size_t)n_sel;

    size_t tmp_bytes = 0, t = 0;
    {
        cub::DoubleBuffer<unsigned long long> k64(nullptr, nullptr);
        cub::DoubleBuffer<uint32_t> v32(nullptr, nullptr), k32(nullptr, nullptr);
        cub::DeviceRadixSort::SortPairs(nullptr, t, k64, v32, n_sel, 0, 64, stream);
        tmp_bytes = t;
        cub::DeviceRadixSort::SortPairs(nullptr, t, k32, v32, n_sel, 0, 32, stream);
        if (t > tmp_bytes) tmp_bytes = t;
        cub::DeviceSelect::Flagged(nullptr, t, (uint32_t *)nullptr, (uint8_t *)nullptr, (uint32_t *)nullptr, (long long *)nullptr,
                                   n_sel, stream);
        if (t > tmp_bytes) tmp_bytes = t;
    }
    unsigned long long *ka, *kb;
    uint32_t *pa, *pb, *iota, *heads, *cnt, *fa, *fb, *perm_a, *perm_b;
    uint8_t *flag;
    void *tmp;
    Arena ar;
    for (int pass = 0; pass < 2; ++pass) {
        ka = ar.take<unsigned long long>(n); kb = ar.take<unsigned long long>(n);
        pa = ar.take<uint32_t>(n); pb = ar.take<uint32_t>(n);
        iota = ar.take<uint32_t>(n); heads = ar.take<uint32_t>(n); cnt = ar.take<uint32_t>(n);
        fa = ar.take<uint32_t>(n); fb = ar.take<uint32_t>(n);
        perm_a = ar.take<uint32_t>(n); perm_b = ar.take<uint32_t>(n);
        flag = ar.take<uint8_t>(n);
        tmp = ar.take<uint8_t>(tmp_bytes);
        if (pass == 0) {
            void *base = nullptr;
            int rc = dyn_ctx_scratch(ctx, 9, ar.used + 256, &base);
            if (rc) return rc;
            ar = Arena(base);
        }
    }
    KnIn in;
    in.counts = reinterpret_cast<const uint4 *>(d_counts);
    in.strand = d_strand; in.group = d_group; in.gene = d_gene; in.sel = d_sel; in.n_sel = n_sel;
    in.conv_mask = conv_mask & 0xfffu; in.base_mask = base_mask & 0xfu; in.exclude_mask = exclude_mask & 0xfffu;
    in.gene_bits = gene_bits;
    const unsigned g = grid_for(n_sel);
    kn_key_kernel<<<g, kBlock, 0, stream>>>(in, ka, pa);
    cub::DoubleBuffer<unsigned long long> k(ka, kb);
    cub::DoubleBuffer<uint32_t> v(pa, pb);
    t = tmp_bytes;
    // the bits in use plus one: valid keys have that extra bit clear, the sentinel (~0) of excluded reads has it set
    // and sorts last; group_bits == 0 (unknown) sorts all 64 bits
    const int end_bit = group_bits > 0 ? std::min(64, 22 + gene_bits + group_bits + 1) : 64;
    DYN_CUDA(cub::DeviceRadixSort::SortPairs(tmp, t, k, v, n_sel, 0, end_bit, stream));
    head_flag_kernel<<<g, kBlock, 0, stream>>>(k.Current(), flag, n_sel);
    iota_kernel<<<g, kBlock, 0, stream>>>(iota, n_sel);
    t = tmp_bytes;
    DYN_CUDA(cub::DeviceSelect::Flagged(tmp, t, iota, flag, heads, reinterpret_cast<long long *>(d_n_rows), n_sel, stream));
    const long long *n_rows = reinterpret_cast<const long long *>(d_n_rows);
    run_count_kernel<<<g, kBlock, 0, stream>>>(k.Current(), heads, n_rows, n_sel, cnt);
    const uint32_t *perm = nullptr;
    if (first_appearance_order) {
        // groupby(sort=False): rows in order of each group's first read (stable sort kept it at the run head)
        pad_first_kernel<<<g, kBlock, 0, stream>>>(v.Current(), heads, n_rows, n_sel, fa);
        iota_kernel<<<g, kBlock, 0, stream>>>(perm_a, n_sel);
        cub::DoubleBuffer<uint32_t> fk(fa, fb), pv(perm_a, perm_b);
        t = tmp_bytes;
        DYN_CUDA(cub::DeviceRadixSort::SortPairs(tmp, t, fk, pv, n_sel, 0, 32, stream));
        perm = pv.Current();
    }
    gather_rows_kernel<<<g, kBlock, 0, stream>>>(k.Current(), v.Current(), heads, cnt, perm, n_rows,
                                                 reinterpret_cast<unsigned long long *>(d_keys), d_count, d_first);
    DYN_CUDA(cudaGetLastError());
    return DYN_OK;
}

extern "C" int dyn_kn_csr(dyn_ctx_t *ctx, const uint64_t *d_keys, const uint32_t *d_count, const int64_t *d_n_rows,
                          int64_t max_rows, int64_t n_groups, int32_t *d_k, int32_t *d_n, int64_t *d_cnt, int64_t *d_off,
                          uint64_t *d_total, void *stream_) {
    DYN_ARG(ctx != nullptr, "null context");
    CtxEnter enter_(ctx, stream_);
    DYN_ARG(max_rows >= 0 && n_groups >= 0, "negative size");
    DYN_ARG(d_n_rows && d_off && d_total, "null buffer");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DYN_CUDA(cudaMemsetAsync(d_total, 0, sizeof(uint64_t) * (n_groups > 0 ? n_groups : 1), stream));
    const long long work = max_rows > n_groups + 1 ? max_rows : n_groups + 1;
    DYN_ARG(max_rows == 0 || (d_keys && d_count && d_k && d_n && d_cnt), "null rows");
    kn_csr_kernel<<<grid_for(work), kBlock, 0, stream>>>(
        reinterpret_cast<const unsigned long long *>(d_keys), d_count, reinterpret_cast<const long long *>(d_n_rows), n_groups,
        d_k, d_n, reinterpret_cast<long long *>(d_cnt), reinterpret_cast<long long *>(d_off),
        reinterpret_cast<unsigned long long *>(d_total));
    DYN_CUDA(cudaGetLastError());
    return DYN_OK;
}

extern "C" int dyn_alpha(dyn_ctx_t *ctx, const uint64_t *d_n_reads, const uint64_t *d_n_with, const double *d_pi_c,
                         int64_t n_groups, double *d_alpha, void *stream) {
    DYN_ARG(ctx != nullptr, "null context");
    CtxEnter enter_(ctx, stream);
    if (n_groups <= 0) return DYN_OK;
    DYN_ARG(d_n_reads && d_n_with && d_pi_c && d_alpha, "null buffer");
    alpha_kernel<<<grid_for(n_groups), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const unsigned long long *>(d_n_reads), reinterpret_cast<const unsigned long long *>(d_n_with),
        d_pi_c, n_groups, d_alpha);
    DYN_CUDA(cudaGetLastError());
    return DYN_OK;
}

extern "C" int dyn_kn_csr_pairs(dyn_ctx_t *ctx, const uint64_t *d_keys, const uint32_t *d_count, const int64_t *d_n_rows,
                                int64_t max_rows, int32_t *d_k, int32_t *d_n, int64_t *d_cnt, int64_t *d_off, uint64_t *d_pair_key,
                                uint64_t *d_total, int64_t *d_n_pairs, void *stream_) {
    DYN_ARG(ctx != nullptr, "null context");
    CtxEnter enter_(ctx, stream_);
    DYN_ARG(max_rows >= 0 && max_rows < 0x7fffffffLL, "max_rows out of range");
    DYN_ARG(d_n_rows && d_off && d_n_pairs, "null buffer");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DYN_CUDA(cudaMemsetAsync(d_n_pairs, 0, 8, stream));
    DYN_CUDA(cudaMemsetAsync(d_off, 0, 8, stream));
    if (max_rows == 0) return DYN_OK;
    DYN_ARG(d_keys && d_count && d_k && d_n && d_cnt && d_pair_key && d_total, "null rows");
    DYN_CUDA(cudaMemsetAsync(d_total, 0, sizeof(uint64_t) * (size_t)max_rows, stream));
    size_t tmp_bytes = 0;
    DYN_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)max_rows, stream));
    uint32_t *head, *pair_of_row;
    void *tmp;
    Arena ar;
    for (int pass = 0; pass < 2; ++pass) {
        head = ar.take<uint32_t>((size_t)max_rows);
        pair_of_row = ar.take<uint32_t>((size_t)max_rows);
        tmp = ar.take<uint8_t>(tmp_bytes);
        if (pass == 0) {
            void *base = nullptr;
            int rc = dyn_ctx_scratch(ctx, 9, ar.used + 256, &base);
            if (rc) return rc;
            ar = Arena(base);
        }
    }
    pair_head_kernel<<<grid_for(max_rows), kBlock, 0, stream>>>(reinterpret_cast<const unsigned long long *>(d_keys),
                                                                reinterpret_cast<const long long *>(d_n_rows), max_rows, head);
    // pair index of a row = (number of heads up to and including it) - 1
    DYN_CUDA(cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, head, pair_of_row, (int)max_rows, stream));
    pair_fill_kernel<<<grid_for(max_rows), kBlock, 0, stream>>>(
        reinterpret_cast<const unsigned long long *>(d_keys), d_count, reinterpret_cast<const long long *>(d_n_rows), head, pair_of_row,
        max_rows, d_k, d_n, reinterpret_cast<long long *>(d_cnt), reinterpret_cast<long long *>(d_off),
        reinterpret_cast<unsigned long long *>(d_pair_key), reinterpret_cast<unsigned long long *>(d_total),
        reinterpret_cast<long long *>(d_n_pairs));
    DYN_CUDA(cudaGetLastError());
    return DYN_OK;
}
