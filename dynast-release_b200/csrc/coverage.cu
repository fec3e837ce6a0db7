// Per-position coverage and grouped key counts.
//
// Replaces dynast/preprocessing/coverage.py:101-149 (calculate_coverage_contig: a second full pysam pass that
// tests every aligned position of every read against a Python set) and the counting dictionaries of
// dynast/preprocessing/snp.py:108-133 (extract_conversions_part).
//   coverage: one thread per counting unit walks the CIGAR's aligned (M/=/X) blocks -- no MD needed --, finds
//   the block's range in the sorted positions-of-interest table by binary search and bumps those entries;
//   for pairs the union of both mates' positions is counted once (coverage.py:136-142).  The first toucher
//   of every position is kept (64-bit atomicMin of unit << 32 | position) because coverage.csv lists
//   positions in first-touch order (coverage.py:144-149, insertion-ordered Counter).
//   unique counts: radix sort + run heads of arbitrary 64-bit keys (count, first occurrence) -- the
//   group-by behind SNP detection: key = (conversion, contig, position).
// HBM-bound: the packed records in (header + CIGAR words only are touched), 4 B atomics per covered
// position of interest out.
#include <cub/cub.cuh>

#include "common.cuh"

namespace {

constexpr int kBlock = 128;
inline unsigned grid_for(long long n) { return (unsigned)((n + kBlock - 1) / kBlock); }

__device__ __forceinline__ long long lower_bound(const unsigned long long *a, long long n, unsigned long long key) {
    long long lo = 0, hi = n;
    while (lo < hi) { const long long mid = (lo + hi) >> 1; if (__ldg(a + mid) < key) lo = mid + 1; else hi = mid; }
    return lo;
}

__device__ bool in_aligned_block(const uint32_t *cig, uint32_t n_cigar, uint32_t pos, uint32_t p) {
    uint32_t r = pos;
    for (uint32_t c = 0; c < n_cigar; ++c) {
        const uint32_t op = cig[c] & 0xf, len = cig[c] >> 4;
        if (op == 0 || op == 7 || op == 8) { if (p >= r && p < r + len) return true; r += len; }
        else if (op == 2 || op == 3) r += len;
    }
    return false;
}

__global__ void coverage_kernel(const uint8_t *records, const uint32_t *rec_off, long long n_units, const uint8_t *selected,
                                const unsigned long long *poi, long long n_poi, uint32_t *cov, unsigned long long *first) {
    const long long u = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (u >= n_units || (selected && !selected[u])) return;
    const uint8_t *rec = records + ((size_t)rec_off[u] << 4);
    const uint4 h = *reinterpret_cast<const uint4 *>(rec);
    const uint32_t n_cigar = h.z >> 16, l_seq = h.z & 0xffff, n_tok = h.w & 0xffff;
    const uint32_t *cig = reinterpret_cast<const uint32_t *>(rec + 16);
    const bool has_mate = ((h.w >> 16) & DYN_REC_HAS_MATE) != 0;
    const unsigned long long contig = (unsigned long long)h.y << 32;
    for (int which = 0; which < (has_mate ? 2 : 1); ++which) {
        const uint8_t *r2 = rec;
        if (which == 1) {
            const size_t b = 16 + 4 * (size_t)n_cigar + 4 * (size_t)n_tok + (((((size_t)l_seq + 1) >> 1) + 3) & ~(size_t)3) + l_seq;
            r2 = rec + ((b + 15) & ~(size_t)15);
        }
        const uint4 h2 = *reinterpret_cast<const uint4 *>(r2);
        const uint32_t nc2 = h2.z >> 16;
        const uint32_t *cig2 = reinterpret_cast<const uint32_t *>(r2 + 16);
        uint32_t r = h2.x;
        for (uint32_t c = 0; c < nc2; ++c) {
            const uint32_t op = cig2[c] & 0xf, len = cig2[c] >> 4;
            if (op == 0 || op == 7 || op == 8) {
                for (long long i = lower_bound(poi, n_poi, contig | r); i < n_poi; ++i) {
                    const unsigned long long k = __ldg(poi + i);
                    if (k >= (contig | (unsigned long long)(r + len)) ) break;
                    const uint32_t p = (uint32_t)k;
                    if (which == 1 && in_aligned_block(cig, n_cigar, h.x, p)) continue;   // union, counted once
                    atomicAdd(cov + i, 1u);
                    atomicMin(first + i, ((unsigned long long)u << 32) | p);
                }
                r += len;
            } else if (op == 2 || op == 3) r += len;
        }
    }
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
__global__ void run_out_kernel(const unsigned long long *key, const uint32_t *pos, const uint32_t *heads, const long long *n_runs,
                               long long n, unsigned long long *out_key, uint32_t *out_count, uint32_t *out_first) {
    const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= *n_runs) return;
    long long end;
    if (j + 1 < *n_runs) end = heads[j + 1];
    else {
        long long lo = heads[j], hi = n;
        while (lo < hi) { const long long mid = (lo + hi) >> 1; if (key[mid] == ~0ull) hi = mid; else lo = mid + 1; }
        end = lo;
    }
    out_key[j] = key[heads[j]];
    out_count[j] = (uint32_t)(end - heads[j]);
    out_first[j] = pos[heads[j]];     // stable sort: the run head is the first occurrence
}

}  // namespace

extern "C" int dyn_coverage(dyn_ctx_t *ctx, const void *d_records, const uint32_t *d_rec_off, int64_t n_units,
                            const uint8_t *d_selected, const uint64_t *d_poi, int64_t n_poi, uint32_t *d_cov,
                            uint64_t *d_first, void *stream_) {
    DYN_ARG(ctx != nullptr, "null context");
    CtxEnter enter_(ctx, stream_);
    DYN_ARG(n_units >= 0 && n_poi >= 0, "negative size");
    if (n_poi == 0) return DYN_OK;
    DYN_ARG(d_poi && d_cov && d_first, "null buffer");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DYN_CUDA(cudaMemsetAsync(d_cov, 0, sizeof(uint32_t) * n_poi, stream));
    DYN_CUDA(cudaMemsetAsync(d_first, 0xff, sizeof(uint64_t) * n_poi, stream));
    if (n_units == 0) return DYN_OK;
    DYN_ARG(d_records && d_rec_off, "null records");
    coverage_kernel<<<grid_for(n_units), kBlock, 0, stream>>>(static_cast<const uint8_t *>(d_records), d_rec_off, n_units, d_selected,
                                                              reinterpret_cast<const unsigned long long *>(d_poi), n_poi, d_cov,
                                                              reinterpret_cast<unsigned long long *>(d_first));
    DYN_CUDA(cudaGetLastError());
    return DYN_OK;
}

extern "C" int dyn_unique_counts(dyn_ctx_t *ctx, const uint64_t *d_keys, int64_t n_keys, uint64_t *d_out_keys,
                                 uint32_t *d_out_count, uint32_t *d_out_first, int64_t *d_n_out, void *stream_) {
    DYN_ARG(ctx != nullptr, "null context");
    CtxEnter enter_(ctx, stream_);
    DYN_ARG(n_keys >= 0 && n_keys < 0xffffffffLL, "n_keys out of range");
    DYN_ARG(d_n_out != nullptr, "null n_out");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DYN_CUDA(cudaMemsetAsync(d_n_out, 0, sizeof(int64_t), stream));
    if (n_keys == 0) return DYN_OK;
    DYN_ARG(d_keys && d_out_keys && d_out_count && d_out_first, "null buffer");
    const size_t n = (size_t)n_keys;
    size_t tmp_bytes = 0, t = 0;
    {
        cub::DoubleBuffer<unsigned long long> k(nullptr, nullptr);
        cub::DoubleBuffer<uint32_t> v(nullptr, nullptr);
        cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k, v, n_keys, 0, 64, stream);
        cub::DeviceSelect::Flagged(nullptr, t, (uint32_t *)nullptr, (uint8_t *)nullptr, (uint32_t *)nullptr, (long long *)nullptr, n_keys, stream);
        if (t > tmp_bytes) tmp_bytes = t;
    }
    unsigned long long *ka, *kb;
    uint32_t *pa, *pb, *iota, *heads;
    uint8_t *flag;
    void *tmp;
    Arena ar;
    for (int pass = 0; pass < 2; ++pass) {
        ka = ar.take<unsigned long long>(n); kb = ar.take<unsigned long long>(n);
        pa = ar.take<uint32_t>(n); pb = ar.take<uint32_t>(n); iota = ar.take<uint32_t>(n); heads = ar.take<uint32_t>(n);
        flag = ar.take<uint8_t>(n);
        tmp = ar.take<uint8_t>(tmp_bytes);
        if (pass == 0) {
            void *base = nullptr;
            int rc = dyn_ctx_scratch(ctx, 11, ar.used + 256, &base);
            if (rc) return rc;
            ar = Arena(base);
        }
    }
    const unsigned g = grid_for(n_keys);
    DYN_CUDA(cudaMemcpyAsync(ka, d_keys, 8 * n, cudaMemcpyDeviceToDevice, stream));
    iota_kernel<<<g, kBlock, 0, stream>>>(pa, n_keys);
    cub::DoubleBuffer<unsigned long long> k(ka, kb);
    cub::DoubleBuffer<uint32_t> v(pa, pb);
    t = tmp_bytes;
    DYN_CUDA(cub::DeviceRadixSort::SortPairs(tmp, t, k, v, n_keys, 0, 64, stream));
    head_flag_kernel<<<g, kBlock, 0, stream>>>(k.Current(), flag, n_keys);
    iota_kernel<<<g, kBlock, 0, stream>>>(iota, n_keys);
    t = tmp_bytes;
    DYN_CUDA(cub::DeviceSelect::Flagged(tmp, t, iota, flag, heads, reinterpret_cast<long long *>(d_n_out), n_keys, stream));
    run_out_kernel<<<g, kBlock, 0, stream>>>(k.Current(), v.Current(), heads, reinterpret_cast<const long long *>(d_n_out), n_keys,
                                             reinterpret_cast<unsigned long long *>(d_out_keys), d_out_count, d_out_first);
    DYN_CUDA(cudaGetLastError());
    return DYN_OK;
}
