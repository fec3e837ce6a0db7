// Shared helpers for the dynast_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <mutex>

#include "../../include/dynast_b200.h"

void dyn_set_error(const char *fmt, ...);

#define DYN_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (call);                                                               \
        if (_e != cudaSuccess) {                                                               \
            dyn_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
            return DYN_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

#define DYN_ARG(cond, msg)                 \
    do {                                   \
        if (!(cond)) {                     \
            dyn_set_error("%s: %s", __func__, msg); \
            return DYN_ERR_ARG;            \
        }                                  \
    } while (0)

constexpr int kScratchSlots = 24;

struct dyn_ctx {
    int device;
    int sm_count;
    size_t smem_optin;      // max dynamic shared memory per block
    cudaStream_t copy_stream[2];
    cudaEvent_t ev[4];
    uint64_t *h_small;      // 256 page-locked bytes: landing zone for small asynchronous device -> host copies
    // lazily grown device scratch (never shrinks); used by host-buffer entry points and sort temp
    void *scratch[kScratchSlots];
    size_t scratch_bytes[kScratchSlots];
    // lazily built device tables
    double *em_coef;        // binomial coefficient table (wrapped-int64 semantics), [em_coef_K][em_coef_N]
    int em_coef_K, em_coef_N;
    // One call at a time per context (the mutex is held for the whole of every entry point), and a scratch slot
    // that changes streams is ordered by an event: the new stream waits for what the previous one queued.
    std::recursive_mutex mu;
    cudaStream_t cur_stream;
    bool cur_managed;
    cudaStream_t slot_stream[kScratchSlots];
    bool slot_used[kScratchSlots];
    cudaEvent_t slot_ev;
    void *ingest_cache;                      // parked buffers of the device BAM ingest (bamdevice.cu)
    void (*ingest_cache_free)(void *);
};

// Entered at the top of every C-ABI call that takes a context: serialises host threads on the context and tells
// dyn_ctx_scratch which stream the call runs on (managed == false: the call orders its own streams).
struct CtxEnter {
    dyn_ctx *c;
    cudaStream_t prev;
    bool prev_managed;
    CtxEnter(dyn_ctx *ctx, void *stream, bool managed = true) : c(ctx) {
        c->mu.lock();
        prev = c->cur_stream; prev_managed = c->cur_managed;
        c->cur_stream = static_cast<cudaStream_t>(stream); c->cur_managed = managed;
    }
    ~CtxEnter() { c->cur_stream = prev; c->cur_managed = prev_managed; c->mu.unlock(); }
    CtxEnter(const CtxEnter &) = delete;
    CtxEnter &operator=(const CtxEnter &) = delete;
};

int dyn_ctx_scratch(dyn_ctx *ctx, int slot, size_t bytes, void **out);

// Scratch slots (device-pointer entry points reuse them in stream order; one stream per context at a time):
//   0 EM scalars, 1-7 host-buffer tally / EM staging, 8 dedup, 9 aggregation, 10 pi, 11 coverage, 12-13 consensus,
//   14-15 tally parked units (one per copy stream), 16 per-read exchange, 17 consensus sorted path
// Bump allocator over one scratch slot: run the same carve() twice, first with base == nullptr to size it.
struct Arena {
    uint8_t *base;
    size_t used;
    explicit Arena(void *b = nullptr) : base(static_cast<uint8_t *>(b)), used(0) {}
    template <typename T>
    T *take(size_t n) {
        used = (used + 255) & ~(size_t)255;
        T *p = base ? reinterpret_cast<T *>(base + used) : nullptr;
        used += sizeof(T) * n;
        return p;
    }
};

// ---- small device helpers -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA 1-D bulk copy global -> shared (SASS: UBLKCP), completion signalled on an mbarrier.
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
