// Strand complement and UMI deduplication.
//
// Replaces dynast/preprocessing/conversion.py:99-125 (complement_counts), :203-243 (deduplicate_counts)
// and the ordering logic of the count_conversions driver (:529-606), which in the reference are pandas
// sorts over per-split CSV files in a process pool.
//
// What the reference computes, restated as keys:
//   * frame order F = (gene strand '-' last, reads without any mismatch last, BAM order)
//       conversion.py:533-541 -- parts with conversions, then count_no_conversions, then complement_counts
//   * split-file order = (barcode in first-appearance order of F, F)          conversion.py:543-579
//   * per split: STABLE ascending sort by ([has_conversions,] transcriptome, score, conversion_sum) on the
//     complemented counts, drop_duplicates((barcode, umi, GX), keep='last')   conversion.py:225-243
//     -> winner of a UMI group = max priority, ties -> last in split-file order; survivors keep sorted order
//   * per split: forward-strand rows, then reverse-strand rows                conversion.py:602-606
// On the device that is three stable LSD radix sorts (cub::DeviceRadixSort, keys restricted to the bits in
// use: [priority | barcode order | strand | no-mismatch], the group key, [split | strand | extra]) and one
// stream compaction; the result is the list of surviving read indices in output order.
// HBM-bound: ~20 B/read algorithmic (16 B key + priority in, 4 B index out); the sorts move more.
#include <cub/cub.cuh>

#include "common.cuh"

#include <algorithm>

namespace {

constexpr int kBlock = 256;

__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

// complement_counts on one 16-byte counter row: the 12 conversion columns reversed, the 4 base columns
// reversed (conversion.py:122-123: CONVERSION_COLUMNS[::-1] + BASE_COLUMNS[::-1]).
__device__ __forceinline__ uint4 complement_row(uint4 v) {
    uint4 o;
    o.x = bswap32(v.z); o.y = bswap32(v.y); o.z = bswap32(v.x); o.w = bswap32(v.w);
    return o;
}

__global__ void complement_kernel(uint4 *counts, const uint8_t *strand, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n || !strand[i]) return;
    counts[i] = complement_row(counts[i]);
}

// first read of every barcode (in the order of the reference's split plan) and reads per barcode.  First
// occurrences sit at the very beginning of the array, so the 64-bit atomicMin is almost always skipped by a plain
// load; the counters are privatised per CTA in shared memory when the barcode table fits (HIST).
template <bool HIST>
__global__ void first_pos_kernel(const uint32_t *barcode, const uint8_t *strand, const uint16_t *conv_n, long long n,
                                 long long n_barcodes, unsigned long long *first_pos, uint32_t *n_reads) {
    extern __shared__ uint32_t s_hist[];
    if (HIST) {
        for (long long b = threadIdx.x; b < n_barcodes; b += blockDim.x) s_hist[b] = 0u;
        __syncthreads();
    }
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const uint32_t b = barcode[i];
        if (b >= (unsigned long long)n_barcodes) continue;
        const unsigned long long pos = ((unsigned long long)(strand[i] != 0) << 33) |
                                       ((unsigned long long)(conv_n[i] == 0) << 32) | (unsigned long long)i;
        if (pos < *reinterpret_cast<volatile unsigned long long *>(first_pos + b)) atomicMin(first_pos + b, pos);
        if (HIST) atomicAdd(&s_hist[b], 1u); else atomicAdd(n_reads + b, 1u);
    }
    if (HIST) {
        __syncthreads();
        for (long long b = threadIdx.x; b < n_barcodes; b += blockDim.x)
            if (s_hist[b]) atomicAdd(n_reads + b, s_hist[b]);
    }
}

// (barcode, umi, gene) -> one 64-bit group key, barcode in the high bits (conversion.py:225 groups on the triple)
__global__ void group_key_kernel(const uint32_t *barcode, const unsigned long long *umi, const uint32_t *gene, long long n,
                                 int umi_bits, int gene_bits, unsigned long long *key) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long um = umi_bits >= 64 ? ~0ull : ((1ull << umi_bits) - 1ull);
    key[i] = ((unsigned long long)barcode[i] << (umi_bits + gene_bits)) | ((umi[i] & um) << gene_bits) | (unsigned long long)gene[i];
}

struct DedupIn {
    const uint4 *counts;
    const uint8_t *strand, *transcriptome;
    const uint16_t *score, *conv_n;
    const uint32_t *barcode, *ord_of_barcode, *split_of_barcode;
    long long n;
    uint32_t conv_mask;
    int use_conversions;
};

// dedup sort key of read i in ONE word: [priority:30][barcode order][strand][no-mismatch]; a stable sort of the
// reads (in BAM order) by it == split-file order refined by the stable priority sort of conversion.py:238
__global__ void sort_key_kernel(DedupIn in, int low_bits, unsigned long long *key, uint32_t *val) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= in.n) return;
    uint4 c = in.counts[i];
    if (in.strand[i]) c = complement_row(c);
    const uint32_t w[3] = {c.x, c.y, c.z};
    uint32_t sum = 0;
#pragma unroll
    for (int j = 0; j < 12; ++j)
        if ((in.conv_mask >> j) & 1u) sum += (w[j >> 2] >> ((j & 3) << 3)) & 0xffu;
    const unsigned long long has = (in.use_conversions && sum > 0) ? 1u : 0u;
    const unsigned long long prio = (has << 29) | ((unsigned long long)(in.transcriptome[i] != 0) << 28) |
                                    ((unsigned long long)in.score[i] << 12) | (sum & 0xfffu);
    const unsigned long long ord = in.ord_of_barcode ? in.ord_of_barcode[in.barcode[i]] : 0u;
    key[i] = (prio << low_bits) | (ord << 2) | ((unsigned long long)(in.strand[i] != 0) << 1) | (unsigned long long)(in.conv_n[i] == 0);
    val[i] = (uint32_t)i;
}

__global__ void gather_u64_kernel(const unsigned long long *src, const uint32_t *idx, unsigned long long *dst, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src ? src[idx[i]] : 0ull;
}

// the last element of every run of equal (hi, lo) keys wins its UMI group
__global__ void winner_kernel(const unsigned long long *hi, const unsigned long long *lo, const uint32_t *idx,
                              uint8_t *win, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool last = (i + 1 == n) || lo[i] != lo[i + 1] || (hi && hi[i] != hi[i + 1]);
    if (last) win[idx[i]] = 1;
}

__global__ void flag_kernel(const uint8_t *win, const uint32_t *order, uint8_t *flag, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) flag[i] = win[order[i]];
}

__global__ void final_key_kernel(DedupIn in, const uint32_t *winners, const long long *n_out, unsigned long long pad,
                                 const unsigned long long *extra, int extra_bits, unsigned long long *key) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= in.n) return;
    if (i >= *n_out) { key[i] = pad; return; }
    const uint32_t r = winners[i];
    const unsigned long long split = in.split_of_barcode ? in.split_of_barcode[in.barcode[r]] : 0u;
    key[i] = ((((split << 1) | (unsigned long long)(in.strand[r] != 0))) << extra_bits) | (extra ? extra[r] : 0ull);
}

__global__ void copy_u32_kernel(const uint32_t *src, uint32_t *dst, const long long *n_valid, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n && i < *n_valid) dst[i] = src[i];
}

// ---- hash path: the winner of every (barcode, umi, gene) group by one atomicMax per read ---------------------------
// Within a group the barcode (hence its place in the split-file order) is the same for every member, so the member
// that ends up last after the reference's stable sorts is the one with the largest (priority, gene strand, no-mismatch
// flag, BAM index) -- 30 + 1 + 1 + 32 bits: one 64-bit atomicMax into an open-addressing table keyed by the group key
// (the strand differs inside a group only on the no-UMI path, whose groups are read ids: conversion.py:148-200).
// Key and value share a 16-byte slot, so a read touches one 32-byte sector of the table.  No sort over all reads: only the survivors (less than half of them in UMI data) are sorted afterwards.
constexpr unsigned long long kEmptyKey = ~0ull;

__device__ __forceinline__ unsigned long long mix_key(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

__device__ __forceinline__ unsigned long long priority_of(const DedupIn &in, long long i) {
    uint4 c = in.counts[i];
    if (in.strand[i]) c = complement_row(c);
    const uint32_t w[3] = {c.x, c.y, c.z};
    uint32_t sum = 0;
#pragma unroll
    for (int j = 0; j < 12; ++j)
        if ((in.conv_mask >> j) & 1u) sum += (w[j >> 2] >> ((j & 3) << 3)) & 0xffu;
    const unsigned long long has = (in.use_conversions && sum > 0) ? 1u : 0u;
    return (has << 29) | ((unsigned long long)(in.transcriptome[i] != 0) << 28) | ((unsigned long long)in.score[i] << 12) | (sum & 0xfffu);
}

__global__ void hash_init_kernel(ulonglong2 *table, unsigned long long cap) {
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i < cap) table[i] = make_ulonglong2(kEmptyKey, 0ull);
}

__global__ void hash_insert_kernel(DedupIn in, const unsigned long long *group_key, ulonglong2 *table, unsigned long long mask,
                                   uint32_t *slot_of) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= in.n) return;
    const unsigned long long k = group_key[i];
    const unsigned long long w = (priority_of(in, i) << 34) | ((unsigned long long)(in.strand[i] != 0) << 33) |
                                 ((unsigned long long)(in.conv_n[i] == 0) << 32) | (unsigned long long)i;
    unsigned long long h = mix_key(k) & mask;
    for (;;) {
        unsigned long long *slot_key = &table[h].x;
        unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(slot_key);
        if (cur == kEmptyKey) cur = atomicCAS(slot_key, kEmptyKey, k);
        if (cur == kEmptyKey || cur == k) break;
        h = (h + 1) & mask;
    }
    atomicMax(&table[h].y, w);
    slot_of[i] = (uint32_t)h;
}

__global__ void hash_winner_kernel(DedupIn in, const ulonglong2 *table, const uint32_t *slot_of, uint8_t *flag, uint32_t *iota) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= in.n) return;
    flag[i] = (uint32_t)table[slot_of[i]].y == (uint32_t)i;      // the winner's BAM index sits in the low word
    iota[i] = (uint32_t)i;
}

// dedup sort key (see sort_key_kernel) of the surviving reads only
__global__ void survivor_key_kernel(DedupIn in, const uint32_t *winners, long long n_out, int low_bits, unsigned long long *key) {
    const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= n_out) return;
    const long long i = winners[j];
    const unsigned long long ord = in.ord_of_barcode ? in.ord_of_barcode[in.barcode[i]] : 0u;
    key[j] = (priority_of(in, i) << low_bits) | (ord << 2) | ((unsigned long long)(in.strand[i] != 0) << 1) | (unsigned long long)(in.conv_n[i] == 0);
}

__global__ void survivor_final_key_kernel(DedupIn in, const uint32_t *winners, long long n_out, const unsigned long long *extra, int extra_bits,
                                          unsigned long long *key) {
    const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= n_out) return;
    const uint32_t r = winners[j];
    const unsigned long long split = in.split_of_barcode ? in.split_of_barcode[in.barcode[r]] : 0u;
    key[j] = ((((split << 1) | (unsigned long long)(in.strand[r] != 0))) << extra_bits) | (extra ? extra[r] : 0ull);
}

inline unsigned grid_for(long long n) { return (unsigned)((n + kBlock - 1) / kBlock); }

inline int bits_for(uint64_t max_value) {
    int b = 1;
    while (b < 64 && (max_value >> b)) ++b;
    return b;
}

}  // namespace

extern "C" int dyn_complement_counts(dyn_ctx_t *ctx, uint8_t *d_counts, const uint8_t *d_strand, int64_t n_reads,
                                     void *stream) {
    DYN_ARG(ctx != nullptr, "null context");
    CtxEnter enter_(ctx, stream);
    DYN_ARG(n_reads >= 0, "negative n_reads");
    if (n_reads == 0) return DYN_OK;
    DYN_ARG(d_counts && d_strand, "null buffer");
    DYN_ARG((reinterpret_cast<uintptr_t>(d_counts) & 15) == 0, "counts must be 16-byte aligned");
    complement_kernel<<<grid_for(n_reads), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<uint4 *>(d_counts), d_strand, n_reads);
    DYN_CUDA(cudaGetLastError());
    return DYN_OK;
}

extern "C" int dyn_barcode_first_pos(dyn_ctx_t *ctx, const uint32_t *d_barcode, const uint8_t *d_strand,
                                     const uint16_t *d_conv_n, int64_t n_reads, int64_t n_barcodes,
                                     uint64_t *d_first_pos, uint32_t *d_n_reads, void *stream_) {
    DYN_ARG(ctx != nullptr, "null context");
    CtxEnter enter_(ctx, stream_);
    DYN_ARG(n_reads >= 0 && n_barcodes >= 0, "negative size");
    if (n_barcodes == 0) return DYN_OK;
    DYN_ARG(d_first_pos && d_n_reads, "null output");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DYN_CUDA(cudaMemsetAsync(d_first_pos, 0xff, sizeof(uint64_t) * n_barcodes, stream));
    DYN_CUDA(cudaMemsetAsync(d_n_reads, 0, sizeof(uint32_t) * n_barcodes, stream));
    if (n_reads == 0) return DYN_OK;
    DYN_ARG(d_barcode && d_strand && d_conv_n, "null buffer");
    const size_t hist_bytes = sizeof(uint32_t) * (size_t)n_barcodes;
    if (hist_bytes <= 64 * 1024) {
        DYN_CUDA(cudaFuncSetAttribute(first_pos_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_bytes));
        const unsigned grid = (unsigned)std::min<long long>((long long)ctx->sm_count * 3, (n_reads + kBlock - 1) / kBlock);
        first_pos_kernel<true><<<grid, kBlock, hist_bytes, stream>>>(d_barcode, d_strand, d_conv_n, n_reads, n_barcodes,
                                                                    reinterpret_cast<unsigned long long *>(d_first_pos),
                                                                    d_n_reads);
    } else {
        first_pos_kernel<false><<<grid_for(n_reads), kBlock, 0, stream>>>(d_barcode, d_strand, d_conv_n, n_reads, n_barcodes,
                                                                         reinterpret_cast<unsigned long long *>(d_first_pos),
                                                                         d_n_reads);
    }
    DYN_CUDA(cudaGetLastError());
    return DYN_OK;
}

extern "C" int dyn_dedup_umi(dyn_ctx_t *ctx, const uint64_t *d_key_hi, const uint64_t *d_key_lo, int key_hi_bits,
                             int key_lo_bits, const uint8_t *d_counts, const uint8_t *d_strand,
                             const uint8_t *d_transcriptome, const uint16_t *d_score, const uint16_t *d_conv_n,
                             const uint32_t *d_barcode, const uint32_t *d_ord_of_barcode,
                             const uint32_t *d_split_of_barcode, int64_t n_barcodes, int64_t n_splits,
                             int64_t n_reads, uint32_t conv_mask, int use_conversions,
                             const uint64_t *d_final_extra, int final_extra_bits, uint32_t *d_out_idx,
                             int64_t *d_n_out, void *stream_) {
    DYN_ARG(ctx != nullptr, "null context");
    CtxEnter enter_(ctx, stream_);
    DYN_ARG(n_reads >= 0 && n_reads < 0xffffffffLL, "n_reads out of range");
    DYN_ARG(d_n_out != nullptr, "null n_out");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DYN_CUDA(cudaMemsetAsync(d_n_out, 0, sizeof(int64_t), stream));
    if (n_reads == 0) return DYN_OK;
    DYN_ARG(d_key_lo && d_counts && d_strand && d_transcriptome && d_score && d_conv_n && d_out_idx, "null buffer");
    DYN_ARG((d_ord_of_barcode == nullptr && d_split_of_barcode == nullptr) || d_barcode, "barcode ids needed");
    DYN_ARG(n_barcodes <= (1LL << 29), "too many barcodes");
    if (key_lo_bits <= 0 || key_lo_bits > 64) key_lo_bits = 64;
    if (key_hi_bits < 0 || key_hi_bits > 64) key_hi_bits = 64;
    const bool has_hi = d_key_hi != nullptr && key_hi_bits > 0;
    const size_t n = (size_t)n_reads;

    DedupIn in;
    in.counts = reinterpret_cast<const uint4 *>(d_counts);
    in.strand = d_strand; in.transcriptome = d_transcriptome; in.score = d_score; in.conv_n = d_conv_n;
    in.barcode = d_barcode; in.ord_of_barcode = d_ord_of_barcode; in.split_of_barcode = d_split_of_barcode;
    in.n = n_reads; in.conv_mask = conv_mask & 0xfffu; in.use_conversions = use_conversions;
    const unsigned g = grid_for(n_reads);
    if (!d_final_extra) final_extra_bits = 0;
    if (n_splits < 1 || !d_split_of_barcode) n_splits = 1;
    const int top_bits = bits_for((uint64_t)(2 * n_splits));
    DYN_ARG(final_extra_bits >= 0 && top_bits + final_extra_bits <= 64, "final key does not fit 64 bits");

    // ---- hash path (64-bit group keys; the all-ones key is the table's empty mark): one atomicMax per read finds the
    //      winners, only the survivors are sorted.  DYN_DEDUP_SORT=1 forces the three-sort path below.
    static const bool force_sort = getenv("DYN_DEDUP_SORT") != nullptr;
    if (!has_hi && key_lo_bits < 64 && !force_sort) {
        size_t cap = 1;
        while (cap < 2 * n) cap <<= 1;
        size_t tmp_bytes = 0, t = 0;
        {
            cub::DoubleBuffer<uint32_t> v32(nullptr, nullptr);
            cub::DoubleBuffer<unsigned long long> k64(nullptr, nullptr);
            cub::DeviceRadixSort::SortPairs(nullptr, t, k64, v32, n_reads, 0, 64, stream);
            tmp_bytes = t;
            cub::DeviceSelect::Flagged(nullptr, t, (uint32_t *)nullptr, (uint8_t *)nullptr, (uint32_t *)nullptr, (long long *)nullptr, n_reads, stream);
            if (t > tmp_bytes) tmp_bytes = t;
        }
        ulonglong2 *table;
        unsigned long long *ka, *kb;
        uint32_t *slot_of, *iota, *wa, *wb;
        uint8_t *flag;
        void *tmp;
        Arena ar;
        for (int pass = 0; pass < 2; ++pass) {
            table = ar.take<ulonglong2>(cap);
            slot_of = ar.take<uint32_t>(n); iota = ar.take<uint32_t>(n);
            wa = ar.take<uint32_t>(n); wb = ar.take<uint32_t>(n);
            ka = ar.take<unsigned long long>(n); kb = ar.take<unsigned long long>(n);
            flag = ar.take<uint8_t>(n);
            tmp = ar.take<uint8_t>(tmp_bytes);
            if (pass == 0) {
                void *base = nullptr;
                int rc = dyn_ctx_scratch(ctx, 8, ar.used + 256, &base);
                if (rc) return rc;
                ar = Arena(base);
            }
        }
        hash_init_kernel<<<(unsigned)((cap + kBlock - 1) / kBlock), kBlock, 0, stream>>>(table, cap);
        hash_insert_kernel<<<g, kBlock, 0, stream>>>(in, reinterpret_cast<const unsigned long long *>(d_key_lo), table, cap - 1, slot_of);
        hash_winner_kernel<<<g, kBlock, 0, stream>>>(in, table, slot_of, flag, iota);
        t = tmp_bytes;
        DYN_CUDA(cub::DeviceSelect::Flagged(tmp, t, iota, flag, wa, reinterpret_cast<long long *>(d_n_out), n_reads, stream));
        long long *h_n = reinterpret_cast<long long *>(ctx->h_small + 16);
        DYN_CUDA(cudaMemcpyAsync(h_n, d_n_out, 8, cudaMemcpyDeviceToHost, stream));
        DYN_CUDA(cudaStreamSynchronize(stream));          // the survivor sorts are sized by their number
        const long long n_out = *h_n;
        if (n_out == 0) return DYN_OK;
        const unsigned go = grid_for(n_out);
        const int low_bits = (d_ord_of_barcode ? bits_for((uint64_t)(n_barcodes > 0 ? n_barcodes - 1 : 0)) : 0) + 2;
        survivor_key_kernel<<<go, kBlock, 0, stream>>>(in, wa, n_out, low_bits, ka);
        cub::DoubleBuffer<unsigned long long> k(ka, kb);
        cub::DoubleBuffer<uint32_t> v(wa, wb);
        t = tmp_bytes;
        DYN_CUDA(cub::DeviceRadixSort::SortPairs(tmp, t, k, v, n_out, 0, low_bits + 30, stream));
        survivor_final_key_kernel<<<go, kBlock, 0, stream>>>(in, v.Current(), n_out, reinterpret_cast<const unsigned long long *>(d_final_extra),
                                                             final_extra_bits, k.Alternate());
        cub::DoubleBuffer<unsigned long long> k2(k.Alternate(), k.Current());
        t = tmp_bytes;
        DYN_CUDA(cub::DeviceRadixSort::SortPairs(tmp, t, k2, v, n_out, 0, top_bits + final_extra_bits, stream));
        DYN_CUDA(cudaMemcpyAsync(d_out_idx, v.Current(), 4 * (size_t)n_out, cudaMemcpyDeviceToDevice, stream));
        DYN_CUDA(cudaGetLastError());
        return DYN_OK;
    }

    size_t tmp_bytes = 0, t = 0;
    {
        cub::DoubleBuffer<uint32_t> k32(nullptr, nullptr), v32(nullptr, nullptr);
        cub::DoubleBuffer<unsigned long long> k64(nullptr, nullptr);
        cub::DeviceRadixSort::SortPairs(nullptr, t, k32, v32, n_reads, 0, 32, stream);
        tmp_bytes = t;
        cub::DeviceRadixSort::SortPairs(nullptr, t, k64, v32, n_reads, 0, 64, stream);
        if (t > tmp_bytes) tmp_bytes = t;
        cub::DeviceSelect::Flagged(nullptr, t, (uint32_t *)nullptr, (uint8_t *)nullptr, (uint32_t *)nullptr,
                                   (long long *)nullptr, n_reads, stream);
        if (t > tmp_bytes) tmp_bytes = t;
    }
    uint32_t *v32a, *v32b, *order;
    unsigned long long *k64a, *k64b, *hi_sorted;
    uint8_t *win, *flag;
    void *tmp;
    Arena ar;
    for (int pass = 0; pass < 2; ++pass) {
        v32a = ar.take<uint32_t>(n); v32b = ar.take<uint32_t>(n);
        order = ar.take<uint32_t>(n);
        k64a = ar.take<unsigned long long>(n); k64b = ar.take<unsigned long long>(n);
        hi_sorted = ar.take<unsigned long long>(has_hi ? n : 1);
        win = ar.take<uint8_t>(n); flag = ar.take<uint8_t>(n);
        tmp = ar.take<uint8_t>(tmp_bytes);
        if (pass == 0) {
            void *base = nullptr;
            int rc = dyn_ctx_scratch(ctx, 8, ar.used + 256, &base);
            if (rc) return rc;
            ar = Arena(base);
        }
    }

    // 1 + 2. dedup sort: one stable sort of the reads (BAM order) by [priority | barcode order | strand | no-mismatch]
    //        -> `order` = read indices, ascending (priority, split-file order)
    {
        const int low_bits = (d_ord_of_barcode ? bits_for((uint64_t)(n_barcodes > 0 ? n_barcodes - 1 : 0)) : 0) + 2;
        sort_key_kernel<<<g, kBlock, 0, stream>>>(in, low_bits, k64a, v32a);
        cub::DoubleBuffer<unsigned long long> k(k64a, k64b);
        cub::DoubleBuffer<uint32_t> v(v32a, v32b);
        t = tmp_bytes;
        DYN_CUDA(cub::DeviceRadixSort::SortPairs(tmp, t, k, v, n_reads, 0, low_bits + 30, stream));
        DYN_CUDA(cudaMemcpyAsync(order, v.Current(), 4 * n, cudaMemcpyDeviceToDevice, stream));
    }
    // 3. group by (barcode, umi, gene) key, stable: the last element of a run is the group's winner
    DYN_CUDA(cudaMemcpyAsync(v32a, order, 4 * n, cudaMemcpyDeviceToDevice, stream));
    gather_u64_kernel<<<g, kBlock, 0, stream>>>(reinterpret_cast<const unsigned long long *>(d_key_lo), v32a, k64a, n_reads);
    uint32_t *grouped_idx;
    unsigned long long *lo_sorted;
    {
        cub::DoubleBuffer<unsigned long long> k(k64a, k64b);
        cub::DoubleBuffer<uint32_t> v(v32a, v32b);
        t = tmp_bytes;
        DYN_CUDA(cub::DeviceRadixSort::SortPairs(tmp, t, k, v, n_reads, 0, key_lo_bits, stream));
        if (has_hi) {
            uint32_t *cur = v.Current();
            unsigned long long *other = k.Alternate();
            gather_u64_kernel<<<g, kBlock, 0, stream>>>(reinterpret_cast<const unsigned long long *>(d_key_hi), cur, other, n_reads);
            cub::DoubleBuffer<unsigned long long> kh(other, k.Current());
            t = tmp_bytes;
            DYN_CUDA(cub::DeviceRadixSort::SortPairs(tmp, t, kh, v, n_reads, 0, key_hi_bits, stream));
            DYN_CUDA(cudaMemcpyAsync(hi_sorted, kh.Current(), 8 * n, cudaMemcpyDeviceToDevice, stream));
            grouped_idx = v.Current();
            lo_sorted = (kh.Current() == k64a) ? k64b : k64a;
            gather_u64_kernel<<<g, kBlock, 0, stream>>>(reinterpret_cast<const unsigned long long *>(d_key_lo), grouped_idx, lo_sorted, n_reads);
        } else {
            grouped_idx = v.Current();
            lo_sorted = k.Current();
        }
    }
    DYN_CUDA(cudaMemsetAsync(win, 0, n, stream));
    winner_kernel<<<g, kBlock, 0, stream>>>(has_hi ? hi_sorted : nullptr, lo_sorted, grouped_idx, win, n_reads);
    // 4. survivors in dedup-sorted order
    flag_kernel<<<g, kBlock, 0, stream>>>(win, order, flag, n_reads);
    uint32_t *winners = (grouped_idx == v32a) ? v32b : v32a;
    t = tmp_bytes;
    DYN_CUDA(cub::DeviceSelect::Flagged(tmp, t, order, flag, winners, reinterpret_cast<long long *>(d_n_out), n_reads, stream));
    // 5. per split: forward-strand genes first, then reverse (stable); padding keys sort to the end
    // (an optional per-read extra key orders rows inside a (split, strand) block: the no-UMI path sorts by
    // (barcode, GX), conversion.py:193-196)
    const unsigned long long pad = (top_bits + final_extra_bits == 64) ? ~0ull : ((unsigned long long)(2 * n_splits) << final_extra_bits);
    final_key_kernel<<<g, kBlock, 0, stream>>>(in, winners, reinterpret_cast<const long long *>(d_n_out), pad,
                                               reinterpret_cast<const unsigned long long *>(d_final_extra), final_extra_bits, k64a);
    {
        uint32_t *valt = (winners == v32a) ? v32b : v32a;
        cub::DoubleBuffer<unsigned long long> k(k64a, k64b);
        cub::DoubleBuffer<uint32_t> v(winners, valt);
        t = tmp_bytes;
        DYN_CUDA(cub::DeviceRadixSort::SortPairs(tmp, t, k, v, n_reads, 0, top_bits + final_extra_bits, stream));
        copy_u32_kernel<<<g, kBlock, 0, stream>>>(v.Current(), d_out_idx, reinterpret_cast<const long long *>(d_n_out), n_reads);
    }
    DYN_CUDA(cudaGetLastError());
    return DYN_OK;
}

extern "C" int dyn_group_key(dyn_ctx_t *ctx, const uint32_t *d_barcode, const uint64_t *d_umi, const uint32_t *d_gene, int64_t n,
                             int umi_bits, int gene_bits, uint64_t *d_key, void *stream) {
    DYN_ARG(ctx != nullptr, "null context");
    CtxEnter enter_(ctx, stream);
    DYN_ARG(n >= 0, "negative n");
    DYN_ARG(umi_bits >= 0 && gene_bits >= 0 && umi_bits + gene_bits < 64, "key fields do not fit 64 bits");
    if (n == 0) return DYN_OK;
    DYN_ARG(d_barcode && d_umi && d_gene && d_key, "null buffer");
    group_key_kernel<<<grid_for(n), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
        d_barcode, reinterpret_cast<const unsigned long long *>(d_umi), d_gene, n, umi_bits, gene_bits,
        reinterpret_cast<unsigned long long *>(d_key));
    DYN_CUDA(cudaGetLastError());
    return DYN_OK;
}
