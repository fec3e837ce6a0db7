// Context management, error reporting and the host-buffer (end-to-end) forms of the C-ABI.
#include <stdarg.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"

static thread_local char g_err[512] = "";

void dyn_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *dyn_last_error(void) { return g_err; }
extern "C" int dyn_version(void) { return 100; }

extern "C" int dyn_ctx_create(int device, dyn_ctx_t **out) {
    DYN_ARG(out != nullptr, "null out pointer");
    int n = 0;
    DYN_CUDA(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) {
        dyn_set_error("dyn_ctx_create: device %d not present (%d CUDA devices); there is no CPU fallback", device, n);
        return DYN_ERR_ARG;
    }
    DYN_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    DYN_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        dyn_set_error("dyn_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                      prop.minor);
        return DYN_ERR_ARG;
    }
    dyn_ctx *ctx = new dyn_ctx();     // value-initialised: every plain member is zero
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    for (int i = 0; i < 2; ++i) DYN_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream[i], cudaStreamNonBlocking));
    for (int i = 0; i < 4; ++i) DYN_CUDA(cudaEventCreateWithFlags(&ctx->ev[i], cudaEventDisableTiming));
    DYN_CUDA(cudaEventCreateWithFlags(&ctx->slot_ev, cudaEventDisableTiming));
    DYN_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&ctx->h_small), 256, cudaHostAllocDefault));
    *out = ctx;
    return DYN_OK;
}

extern "C" int dyn_ctx_destroy(dyn_ctx_t *ctx) {
    if (!ctx) return DYN_OK;
    cudaSetDevice(ctx->device);
    for (int i = 0; i < kScratchSlots; ++i)
        if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
    if (ctx->em_coef) cudaFree(ctx->em_coef);
    for (int i = 0; i < 2; ++i) cudaStreamDestroy(ctx->copy_stream[i]);
    if (ctx->h_small) cudaFreeHost(ctx->h_small);
    for (int i = 0; i < 4; ++i) cudaEventDestroy(ctx->ev[i]);
    if (ctx->slot_ev) cudaEventDestroy(ctx->slot_ev);
    if (ctx->ingest_cache && ctx->ingest_cache_free) ctx->ingest_cache_free(ctx->ingest_cache);
    delete ctx;
    return DYN_OK;
}

extern "C" int dyn_ctx_sm_count(dyn_ctx_t *ctx) { return ctx ? ctx->sm_count : 0; }

int dyn_ctx_scratch(dyn_ctx *ctx, int slot, size_t bytes, void **out) {
    if (slot < 0 || slot >= kScratchSlots) { dyn_set_error("bad scratch slot"); return DYN_ERR_ARG; }
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    if (ctx->cur_managed) {
        // the slot's previous user ran on another stream: order this call behind everything queued there
        if (ctx->slot_used[slot] && ctx->slot_stream[slot] != ctx->cur_stream) {
            if (cudaEventRecord(ctx->slot_ev, ctx->slot_stream[slot]) == cudaSuccess)
                DYN_CUDA(cudaStreamWaitEvent(ctx->cur_stream, ctx->slot_ev, 0));
            else {      // that stream no longer exists
                cudaGetLastError();
                DYN_CUDA(cudaDeviceSynchronize());
            }
        }
        ctx->slot_stream[slot] = ctx->cur_stream;
        ctx->slot_used[slot] = true;
    }
    if (ctx->scratch_bytes[slot] < bytes) {
        if (ctx->scratch[slot]) DYN_CUDA(cudaFree(ctx->scratch[slot]));
        ctx->scratch[slot] = nullptr;
        ctx->scratch_bytes[slot] = 0;
        size_t want = std::max(bytes, (size_t)256);
        cudaError_t e = cudaMalloc(&ctx->scratch[slot], want);
        if (e != cudaSuccess) {
            dyn_set_error("scratch allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
            return DYN_ERR_NOMEM;
        }
        ctx->scratch_bytes[slot] = want;
    }
    *out = ctx->scratch[slot];
    return DYN_OK;
}

// --------------------------------------------------------------------------------------------------
// Host-buffer tally: chunked, double-buffered H2D -> kernel -> D2H on two streams.
// Scratch slots: 1,2 = record chunks; 3,4 = offsets; 5,6 = per-chunk outputs; 7 = misc scalars + snps.
// --------------------------------------------------------------------------------------------------
extern "C" int dyn_tally_reads_host(dyn_ctx_t *ctx, const void *h_records, const uint32_t *h_rec_off,
                                    int64_t n_reads, const uint64_t *h_snps, int64_t n_snps, int quality,
                                    uint32_t flags, uint8_t *h_counts, uint32_t *h_conv_off, uint16_t *h_conv_n,
                                    uint64_t *h_conv_rec, uint32_t *h_conv_read, int64_t conv_capacity,
                                    uint64_t *h_conv_total, uint32_t *h_status) {
    DYN_ARG(ctx != nullptr, "null context");
    CtxEnter enter_(ctx, nullptr, false);   // manages its own two streams and slots
    DYN_ARG(n_reads >= 0, "negative n_reads");
    DYN_ARG(h_status != nullptr, "null status");
    *h_status = 0;
    if (h_conv_total) *h_conv_total = 0;
    if (n_reads == 0) return DYN_OK;
    DYN_ARG(h_records && h_rec_off && h_counts, "null buffer");
    const bool emit = (flags & DYN_TALLY_EMIT_CONV) != 0;
    if ((flags >> 16) == 0) {   // mean record size hint for the staging buffer (DYN_TALLY_REC16_HINT)
        const uint64_t avg16 = ((uint64_t)(h_rec_off[n_reads] - h_rec_off[0]) + n_reads - 1) / (uint64_t)n_reads;
        flags |= (uint32_t)std::min<uint64_t>(avg16, 0xffff) << 16;
    }
    if (emit) DYN_ARG(h_conv_off && h_conv_n && h_conv_rec && h_conv_read && h_conv_total, "null conversion buffer");
    DYN_CUDA(cudaSetDevice(ctx->device));

    const int64_t kChunkReads = 1 << 20;   // ~190 MB of records per chunk at L = 91
    const int64_t n_chunks = (n_reads + kChunkReads - 1) / kChunkReads;

    // misc: [0] status u32, [8] conv_total u64, [64..] snps
    void *misc = nullptr;
    int rc = dyn_ctx_scratch(ctx, 7, 64 + sizeof(uint64_t) * (size_t)std::max<int64_t>(n_snps, 1), &misc);
    if (rc) return rc;
    uint32_t *d_status = static_cast<uint32_t *>(misc);
    uint64_t *d_snps = reinterpret_cast<uint64_t *>(static_cast<uint8_t *>(misc) + 64);
    DYN_CUDA(cudaMemsetAsync(misc, 0, 64, ctx->copy_stream[0]));
    if (n_snps > 0)
        DYN_CUDA(cudaMemcpyAsync(d_snps, h_snps, sizeof(uint64_t) * n_snps, cudaMemcpyHostToDevice, ctx->copy_stream[0]));
    DYN_CUDA(cudaStreamSynchronize(ctx->copy_stream[0]));

    // size the per-chunk buffers from the largest chunk
    size_t max_bytes = 0, max_conv = 0;
    for (int64_t c = 0; c < n_chunks; ++c) {
        const int64_t r0 = c * kChunkReads, r1 = std::min(n_reads, r0 + kChunkReads);
        max_bytes = std::max(max_bytes, (size_t)(h_rec_off[r1] - h_rec_off[r0]) << 4);
    }
    // conversion records: per-chunk device capacity; a chunk cannot emit more mismatches than bases,
    // so bound by bytes (every mismatch needs >= 1 quality byte) but cap to what the caller can take.
    max_conv = emit ? std::min<size_t>((size_t)conv_capacity, max_bytes) : 0;

    struct Slot {
        int64_t r0, nr;
        uint32_t *d_conv_off; uint16_t *d_conv_n; uint64_t *d_conv_rec; uint32_t *d_conv_read; uint64_t *d_conv_total;
        uint64_t *chunk_total;  // page-locked landing zone: a copy into pageable memory would block the host until
                                // the chunk is done and serialise the two streams
    } slot[2];
    slot[0].chunk_total = ctx->h_small;
    slot[1].chunk_total = ctx->h_small + 8;
    uint64_t total_conv = 0;

    // Software pipeline: step c enqueues chunk c (H2D, kernel, fixed-size D2H) on stream c&1, then
    // finalises chunk c-1 (whose variable-size conversion records need its total first).
    for (int64_t c = 0; c <= n_chunks; ++c) {
        if (c < n_chunks) {
            const int s = (int)(c & 1);
            cudaStream_t st = ctx->copy_stream[s];
            const int64_t r0 = c * kChunkReads, r1 = std::min(n_reads, r0 + kChunkReads);
            const int64_t nr = r1 - r0;
            const size_t bytes = (size_t)(h_rec_off[r1] - h_rec_off[r0]) << 4;
            void *d_rec = nullptr, *d_off = nullptr, *d_out = nullptr;
            if ((rc = dyn_ctx_scratch(ctx, 1 + s, max_bytes, &d_rec))) return rc;
            if ((rc = dyn_ctx_scratch(ctx, 3 + s, sizeof(uint32_t) * (kChunkReads + 1), &d_off))) return rc;
            // outputs: counts (16 B) + conv_off (4 B) + conv_n (2 B) per read, then conv_rec (8 B) + conv_read (4 B)
            const size_t out_fixed = (size_t)kChunkReads * 24;
            if ((rc = dyn_ctx_scratch(ctx, 5 + s, out_fixed + max_conv * 12 + 64, &d_out))) return rc;
            uint8_t *d_counts = static_cast<uint8_t *>(d_out);
            Slot &sl = slot[s];
            sl.r0 = r0; sl.nr = nr;
            sl.d_conv_off = reinterpret_cast<uint32_t *>(d_counts + (size_t)kChunkReads * 16);
            sl.d_conv_n = reinterpret_cast<uint16_t *>(d_counts + (size_t)kChunkReads * 20);
            sl.d_conv_rec = reinterpret_cast<uint64_t *>(d_counts + out_fixed);
            sl.d_conv_read = reinterpret_cast<uint32_t *>(d_counts + out_fixed + max_conv * 8);
            sl.d_conv_total = reinterpret_cast<uint64_t *>(
                (reinterpret_cast<uintptr_t>(d_counts + out_fixed + max_conv * 12) + 7) & ~(uintptr_t)7);

            // stream order protects the slot's buffers: chunk c-2 on this stream was finalised at step c-1
            DYN_CUDA(cudaMemcpyAsync(d_rec, static_cast<const uint8_t *>(h_records) + ((size_t)h_rec_off[r0] << 4),
                                     bytes, cudaMemcpyHostToDevice, st));
            DYN_CUDA(cudaMemcpyAsync(d_off, h_rec_off + r0, sizeof(uint32_t) * (nr + 1), cudaMemcpyHostToDevice, st));
            DYN_CUDA(cudaMemsetAsync(sl.d_conv_total, 0, 8, st));
            // offsets are absolute: rebase the record pointer instead of rewriting them
            const uint8_t *d_rec_base = static_cast<const uint8_t *>(d_rec) - ((size_t)h_rec_off[r0] << 4);
            rc = dyn_tally_reads(ctx, d_rec_base, static_cast<const uint32_t *>(d_off), nr, n_snps ? d_snps : nullptr,
                                 n_snps, quality, flags, d_counts, sl.d_conv_off, sl.d_conv_n, sl.d_conv_rec,
                                 sl.d_conv_read, (int64_t)max_conv, sl.d_conv_total, d_status, st);
            if (rc) return rc;
            DYN_CUDA(cudaMemcpyAsync(h_counts + (size_t)r0 * 16, d_counts, (size_t)nr * 16, cudaMemcpyDeviceToHost, st));
            if (emit) {
                DYN_CUDA(cudaMemcpyAsync(sl.chunk_total, sl.d_conv_total, 8, cudaMemcpyDeviceToHost, st));
                DYN_CUDA(cudaMemcpyAsync(h_conv_off + r0, sl.d_conv_off, sizeof(uint32_t) * nr, cudaMemcpyDeviceToHost, st));
                DYN_CUDA(cudaMemcpyAsync(h_conv_n + r0, sl.d_conv_n, sizeof(uint16_t) * nr, cudaMemcpyDeviceToHost, st));
            }
        }
        if (c >= 1) {
            const int s = (int)((c - 1) & 1);
            cudaStream_t st = ctx->copy_stream[s];
            Slot &sl = slot[s];
            DYN_CUDA(cudaStreamSynchronize(st));
            if (emit) {
                uint64_t chunk_total = *sl.chunk_total;
                if (chunk_total > max_conv) chunk_total = max_conv;   // status carries DYN_ST_CONV_CAPACITY
                if (total_conv + chunk_total > (uint64_t)conv_capacity) {
                    dyn_set_error("dyn_tally_reads_host: conversion capacity %lld exceeded", (long long)conv_capacity);
                    return DYN_ERR_NOMEM;
                }
                DYN_CUDA(cudaMemcpyAsync(h_conv_rec + total_conv, sl.d_conv_rec, 8 * chunk_total, cudaMemcpyDeviceToHost, st));
                DYN_CUDA(cudaMemcpyAsync(h_conv_read + total_conv, sl.d_conv_read, 4 * chunk_total, cudaMemcpyDeviceToHost, st));
                DYN_CUDA(cudaStreamSynchronize(st));
                // chunk-local -> global numbering
                for (int64_t i = sl.r0; i < sl.r0 + sl.nr; ++i) h_conv_off[i] += (uint32_t)total_conv;
                for (uint64_t i = 0; i < chunk_total; ++i) h_conv_read[total_conv + i] += (uint32_t)sl.r0;
                total_conv += chunk_total;
            }
        }
    }
    DYN_CUDA(cudaStreamSynchronize(ctx->copy_stream[0]));
    DYN_CUDA(cudaStreamSynchronize(ctx->copy_stream[1]));
    DYN_CUDA(cudaMemcpy(h_status, d_status, 4, cudaMemcpyDeviceToHost));
    if (h_conv_total) *h_conv_total = total_conv;
    return DYN_OK;
}

extern "C" int dyn_em_pc_host(dyn_ctx_t *ctx, const int32_t *h_k, const int32_t *h_n, const int64_t *h_count,
                              const int64_t *h_off, int64_t B, const double *h_p_e, int nasc, double *h_p_c,
                              int32_t *h_em_status, int32_t *h_iters) {
    DYN_ARG(ctx != nullptr, "null context");
    CtxEnter enter_(ctx, nullptr, false);   // manages its own two streams and slots
    if (B <= 0) return DYN_OK;
    DYN_ARG(h_off && h_p_e && h_p_c && h_em_status, "null buffer");
    DYN_CUDA(cudaSetDevice(ctx->device));
    const int64_t T = h_off[B];
    DYN_ARG(T == 0 || (h_k && h_n && h_count), "null triplets");
    cudaStream_t st = ctx->copy_stream[0];
    const size_t bytes = (size_t)T * 16 + (size_t)(B + 1) * 8 + (size_t)B * (8 + 8 + 4 + 4) + 256;
    void *buf = nullptr;
    int rc = dyn_ctx_scratch(ctx, 1, bytes, &buf);
    if (rc) return rc;
    uint8_t *p = static_cast<uint8_t *>(buf);
    int64_t *d_count = reinterpret_cast<int64_t *>(p); p += (size_t)T * 8;
    int64_t *d_off = reinterpret_cast<int64_t *>(p); p += (size_t)(B + 1) * 8;
    double *d_pe = reinterpret_cast<double *>(p); p += (size_t)B * 8;
    double *d_pc = reinterpret_cast<double *>(p); p += (size_t)B * 8;
    int32_t *d_k = reinterpret_cast<int32_t *>(p); p += (size_t)T * 4;
    int32_t *d_n = reinterpret_cast<int32_t *>(p); p += (size_t)T * 4;
    int32_t *d_st = reinterpret_cast<int32_t *>(p); p += (size_t)B * 4;
    int32_t *d_it = reinterpret_cast<int32_t *>(p);
    if (T) {
        DYN_CUDA(cudaMemcpyAsync(d_k, h_k, (size_t)T * 4, cudaMemcpyHostToDevice, st));
        DYN_CUDA(cudaMemcpyAsync(d_n, h_n, (size_t)T * 4, cudaMemcpyHostToDevice, st));
        DYN_CUDA(cudaMemcpyAsync(d_count, h_count, (size_t)T * 8, cudaMemcpyHostToDevice, st));
    }
    DYN_CUDA(cudaMemcpyAsync(d_off, h_off, (size_t)(B + 1) * 8, cudaMemcpyHostToDevice, st));
    DYN_CUDA(cudaMemcpyAsync(d_pe, h_p_e, (size_t)B * 8, cudaMemcpyHostToDevice, st));
    rc = dyn_em_pc(ctx, d_k, d_n, d_count, d_off, B, d_pe, nasc, d_pc, d_st, d_it, st);
    if (rc) return rc;
    DYN_CUDA(cudaMemcpyAsync(h_p_c, d_pc, (size_t)B * 8, cudaMemcpyDeviceToHost, st));
    DYN_CUDA(cudaMemcpyAsync(h_em_status, d_st, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
    if (h_iters) DYN_CUDA(cudaMemcpyAsync(h_iters, d_it, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
    DYN_CUDA(cudaStreamSynchronize(st));
    return DYN_OK;
}

// --------------------------------------------------------------------------------------------------
// Host side of a6: barcode order and split assignment (conversion.py:543-579) from the per-barcode first
// frame position / read count that dyn_barcode_first_pos produced.  Sequential by nature (a split closes when
// its running size passes the threshold), O(B log B) on the host for B barcodes.
// --------------------------------------------------------------------------------------------------
#include <numeric>
#include <vector>

extern "C" int dyn_split_plan(const uint64_t *h_first_pos, const uint32_t *h_n_reads, int64_t n_barcodes, int64_t threshold,
                              uint32_t *h_ord_of_barcode, uint32_t *h_split_of_barcode, int64_t *h_n_splits) {
    DYN_ARG(n_barcodes >= 0, "negative n_barcodes");
    DYN_ARG(h_n_splits != nullptr, "null n_splits");
    *h_n_splits = 1;
    if (n_barcodes == 0) return DYN_OK;
    DYN_ARG(h_first_pos && h_n_reads && h_ord_of_barcode && h_split_of_barcode, "null buffer");
    // barcodes in first-appearance order: a stable LSD radix sort of (first position, barcode) on 16-bit digits (the
    // comparison sorts this replaces cost 24 ms at 125 k barcodes -- two thirds of the dedup stage of config C5)
    const size_t B = (size_t)n_barcodes;
    std::vector<uint64_t> key(B), key2(B);
    std::vector<uint32_t> order(B), order2(B);
    uint64_t all_or = 0, all_and = ~0ull;
    for (size_t b = 0; b < B; ++b) {
        key[b] = h_n_reads[b] ? h_first_pos[b] : ~0ull;
        order[b] = (uint32_t)b;
        all_or |= key[b];
        all_and &= key[b];
    }
    std::vector<uint32_t> hist(65536);
    for (int shift = 0; shift < 64; shift += 16) {
        if ((((all_or ^ all_and) >> shift) & 0xffffu) == 0) continue;          // every key has the same digit here
        std::fill(hist.begin(), hist.end(), 0u);
        for (size_t i = 0; i < B; ++i) ++hist[(key[i] >> shift) & 0xffffu];
        uint32_t acc = 0;
        for (auto &h : hist) { const uint32_t c = h; h = acc; acc += c; }
        for (size_t i = 0; i < B; ++i) {
            const uint32_t at = hist[(key[i] >> shift) & 0xffffu]++;
            key2[at] = key[i];
            order2[at] = order[i];
        }
        key.swap(key2);
        order.swap(order2);
    }
    int64_t n_splits = 0, cur = -1, cur_size = 0;
    for (size_t i = 0; i < B; ++i) {
        const uint32_t b = order[i];
        h_split_of_barcode[b] = 0;
        if (!h_n_reads[b]) continue;
        if ((int64_t)h_n_reads[b] > threshold) h_split_of_barcode[b] = (uint32_t)n_splits++;        // its own split
        else if (cur < 0) { cur = n_splits++; h_split_of_barcode[b] = (uint32_t)cur; }             // opens a residual split
        else {
            h_split_of_barcode[b] = (uint32_t)cur;
            cur_size += h_n_reads[b];
            if (cur_size > threshold) { cur = -1; cur_size = 0; }
        }
    }
    // split-file order: by split, then first appearance -- a counting sort of the first-appearance order by split id
    std::vector<uint32_t> start((size_t)std::max<int64_t>(n_splits, 1) + 1, 0u);
    for (size_t b = 0; b < B; ++b) ++start[h_split_of_barcode[b] + 1];
    for (size_t q = 1; q < start.size(); ++q) start[q] += start[q - 1];
    for (size_t i = 0; i < B; ++i) {
        const uint32_t b = order[i];
        h_ord_of_barcode[b] = start[h_split_of_barcode[b]]++;
    }
    *h_n_splits = n_splits > 0 ? n_splits : 1;
    return DYN_OK;
}
