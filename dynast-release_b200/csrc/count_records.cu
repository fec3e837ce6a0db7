// Counters from mismatch records (file-based entry).
//
// Replaces the per-line loop of conversion.count_conversions_part (dynast/preprocessing/conversion.py:417-426)
// for callers that hold conversions.csv / alignments.csv (the reference's on-disk interchange) instead of
// packed BAM records: every mismatch record with Phred > quality that is not a SNP adds one to its read's
// conversion column.  One thread per record, 32-bit atomics on the read's packed 8-bit counters.
// HBM-bound: 16 B per record in (8 B record + 4 B read + 4 B contig), scattered 4-byte atomics out.
#include "common.cuh"

namespace {

__device__ __forceinline__ bool snp_hit(const unsigned long long *snps, long long n, unsigned long long key) {
    long long lo = 0, hi = n;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (__ldg(snps + mid) < key) lo = mid + 1; else hi = mid;
    }
    return lo < n && __ldg(snps + lo) == key;
}

__global__ void count_records_kernel(const unsigned long long *rec, const uint32_t *read, const uint32_t *contig, long long m,
                                     long long n_reads, const unsigned long long *snps, long long n_snps, int quality,
                                     uint32_t *counts, uint32_t *status) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= m) return;
    const unsigned long long r = rec[i];
    const uint32_t q = (uint32_t)(r >> 40) & 0xff, code = (uint32_t)(r >> 32) & 0xff;
    if ((int)q <= quality) return;                       // conversion.py:424 (strict)
    const uint32_t gn = code >> 4, rn = code & 15;
    if (__popc(gn) != 1 || __popc(rn) != 1 || gn == rn) { atomicOr(status, DYN_ST_NON_ACGT_CONV); return; }
    const int g = __ffs((int)gn) - 1, d = __ffs((int)rn) - 1;
    const int idx = g * 3 + (d > g ? d - 1 : d);
    if (n_snps > 0 && snp_hit(snps, n_snps, ((unsigned long long)idx << 56) | ((unsigned long long)contig[i] << 32) |
                                                (r & 0xffffffffull)))
        return;
    const uint32_t rd = read[i];
    if (rd >= (unsigned long long)n_reads) { atomicOr(status, DYN_ST_BAD_RECORD); return; }
    const uint32_t sh = (idx & 3) << 3;
    uint32_t *w = counts + (size_t)rd * 4 + (idx >> 2);
    const uint32_t old = atomicAdd(w, 1u << sh);
    if (((old >> sh) & 0xff) == 0xff) { atomicSub(w, 1u << sh); atomicOr(status, DYN_ST_COUNTER_OVERFLOW); }
}

}  // namespace

extern "C" int dyn_count_records(dyn_ctx_t *ctx, const uint64_t *d_conv_rec, const uint32_t *d_conv_read,
                                 const uint32_t *d_conv_contig, int64_t n_records, int64_t n_reads, const uint64_t *d_snps,
                                 int64_t n_snps, int quality, uint8_t *d_counts, uint32_t *d_status, void *stream) {
    DYN_ARG(ctx != nullptr, "null context");
    CtxEnter enter_(ctx, stream);
    DYN_ARG(n_records >= 0 && n_reads >= 0, "negative size");
    if (n_records == 0) return DYN_OK;
    DYN_ARG(d_conv_rec && d_conv_read && d_counts && d_status, "null buffer");
    DYN_ARG(n_snps == 0 || (d_snps && d_conv_contig), "snp filter needs the table and the contig ids");
    count_records_kernel<<<(unsigned)((n_records + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const unsigned long long *>(d_conv_rec), d_conv_read, d_conv_contig, n_records, n_reads,
        reinterpret_cast<const unsigned long long *>(d_snps), n_snps, quality, reinterpret_cast<uint32_t *>(d_counts), d_status);
    DYN_CUDA(cudaGetLastError());
    return DYN_OK;
}
