// Per-read exchange to the barcode owner ("the one exchange"): UMI groups
// (barcode, umi, gene) may straddle the read ranges the GPUs tally, so before deduplication every rank sends
// the fixed-size per-read row -- 16 counters, group key, priority fields -- to rank barcode mod world.  The
// reference has no counterpart (one process reads the whole conversions.csv, conversion.py:529-606); these
// kernels are the device side of dynast_b200.distributed.exchange_reads:
//   dyn_owner_partition  stable partition of the local reads by owner (one radix pass over log2(world) bits)
//                        + per-owner counts = the all-to-all split sizes
//   dyn_pack_reads       gathers the struct-of-arrays columns of the partitioned reads into 40-byte rows, so the
//                        exchange is ONE all-to-all of one buffer
//   dyn_unpack_reads     rows -> columns on the receiving side
//   dyn_merge_kn         the histogram exchange before the EM (no-UMI variant of SURVEY 8e): (group, k, n) rows with
//                        read counts received from several ranks -> one sorted row per key with the counts summed,
//                        group ids rewritten to the owner's local numbering (global id / world)
// HBM-bound streaming passes: 38 B of columns in, 40 B out per read (and back).
#include <cub/cub.cuh>

#include "common.cuh"

#include <algorithm>
#include <cstring>

namespace {

constexpr int kBlock = 256;
inline unsigned grid_for(long long n) { return (unsigned)((n + kBlock - 1) / kBlock); }

__global__ void owner_kernel(const uint32_t *barcode, long long n, uint32_t world, uint8_t *owner, uint32_t *iota,
                             unsigned long long *counts) {
    __shared__ unsigned int local[DYN_MAX_WORLD];
    for (int i = threadIdx.x; i < DYN_MAX_WORLD; i += blockDim.x) local[i] = 0;
    __syncthreads();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const uint32_t o = barcode[i] % world;
        owner[i] = (uint8_t)o;
        iota[i] = (uint32_t)i;
        atomicAdd(&local[o], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (int)world; i += blockDim.x)
        if (local[i]) atomicAdd(&counts[i], (unsigned long long)local[i]);
}

struct alignas(8) ReadRow {     // 40 bytes
    uint32_t counts[4];
    unsigned long long umi;
    uint32_t barcode, gene;
    int16_t score;
    uint16_t conv_n;
    uint8_t strand, transcriptome, pad[2];
};
static_assert(sizeof(ReadRow) == DYN_READ_ROW_BYTES, "row layout");

__global__ void pack_kernel(const uint32_t *perm, long long n, const uint4 *counts, const uint32_t *barcode,
                            const unsigned long long *umi, const uint32_t *gene, const int16_t *score, const uint16_t *conv_n,
                            const uint8_t *strand, const uint8_t *transcriptome, ReadRow *rows) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long s = perm ? (long long)perm[i] : i;
    ReadRow r;
    const uint4 cv = counts[s];
    r.counts[0] = cv.x; r.counts[1] = cv.y; r.counts[2] = cv.z; r.counts[3] = cv.w;
    r.umi = umi[s];
    r.barcode = barcode[s];
    r.gene = gene[s];
    r.score = score[s];
    r.conv_n = conv_n[s];
    r.strand = strand[s];
    r.transcriptome = transcriptome ? transcriptome[s] : (uint8_t)1;
    r.pad[0] = r.pad[1] = 0;
    const unsigned long long *w = reinterpret_cast<const unsigned long long *>(&r);
    unsigned long long *o = reinterpret_cast<unsigned long long *>(rows + i);
#pragma unroll
    for (int k = 0; k < 5; ++k) o[k] = w[k];
}

__global__ void unpack_kernel(const ReadRow *rows, long long n, uint4 *counts, uint32_t *barcode, unsigned long long *umi,
                              uint32_t *gene, int16_t *score, uint16_t *conv_n, uint8_t *strand, uint8_t *transcriptome) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    ReadRow r;
    const unsigned long long *s = reinterpret_cast<const unsigned long long *>(rows + i);
    unsigned long long *w = reinterpret_cast<unsigned long long *>(&r);
#pragma unroll
    for (int k = 0; k < 5; ++k) w[k] = s[k];
    counts[i] = make_uint4(r.counts[0], r.counts[1], r.counts[2], r.counts[3]);
    umi[i] = r.umi;
    barcode[i] = r.barcode;
    gene[i] = r.gene;
    score[i] = r.score;
    conv_n[i] = r.conv_n;
    strand[i] = r.strand;
    transcriptome[i] = r.transcriptome;
}

__global__ void localise_keys_kernel(const unsigned long long *in, long long n, uint32_t world, unsigned long long *out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = in[i];
    out[i] = (((k >> 22) / world) << 22) | (k & 0x3fffffull);
}


// ---- fused exchange: stable partition + row gather + store into the OWNER'S memory over NVLink -----------------------
// dyn_owner_partition + dyn_pack_reads + all_to_all cost three passes over the rows (permutation, packed send buffer, the
// collective's copy).  With every rank's receive buffer mapped into every other rank (cudaIpc) the rows can be written
// where they are needed: tile_count_kernel counts owners per tile of kPushTile reads, one exclusive scan in
// owner-major order turns the counts into the stable position of every (owner, tile) run, and push_kernel gathers the
// columns of its tile into shared memory grouped by owner and streams each owner's run -- contiguous at the
// destination -- with coalesced 8-byte stores through the peer mapping.  No permutation, no send buffer, no collective
// on the data path (the counts still go through one tiny all-gather, which doubles as the buffer-reuse barrier).
constexpr int kPushTile = DYN_PUSH_TILE;
static_assert(kPushTile == kBlock, "one read per thread");

struct PushPeers {
    unsigned long long *rows[DYN_PUSH_MAX_WORLD];     // receive buffers (8-byte words), mine and the mapped ones
    long long base[DYN_PUSH_MAX_WORLD];               // destination row of my first row for that owner - owner_start
};

// stable rank of this thread's read among the tile's reads with the same owner; warp_count[w][o] is left filled
__device__ __forceinline__ uint32_t tile_rank(uint32_t o, uint32_t world, uint32_t (*warp_count)[DYN_PUSH_MAX_WORLD + 1]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (kBlock / 32) * (DYN_PUSH_MAX_WORLD + 1); i += kBlock) (&warp_count[0][0])[i] = 0;
    __syncthreads();
    const unsigned m = __match_any_sync(0xffffffffu, o);
    const uint32_t in_warp = __popc(m & ((1u << lane) - 1u));
    if (in_warp == 0) warp_count[warp][o] = __popc(m);
    __syncthreads();
    uint32_t before = in_warp;
    for (int w = 0; w < warp; ++w) before += warp_count[w][o];
    (void)world;
    return before;
}

__global__ void __launch_bounds__(kBlock) tile_count_kernel(const uint32_t *barcode, long long n, uint32_t world, long long n_tiles,
                                                            uint32_t *tile_count) {
    __shared__ uint32_t warp_count[kBlock / 32][DYN_PUSH_MAX_WORLD + 1];
    const long long i = blockIdx.x * (long long)kBlock + threadIdx.x;
    const uint32_t o = i < n ? barcode[i] % world : world;
    tile_rank(o, world, warp_count);
    if (threadIdx.x < world) {
        uint32_t c = 0;
        for (int w = 0; w < kBlock / 32; ++w) c += warp_count[w][threadIdx.x];
        tile_count[(long long)threadIdx.x * n_tiles + blockIdx.x] = c;
    }
}

__global__ void owner_totals_kernel(const uint32_t *tile_base, long long n, long long n_tiles, uint32_t world, unsigned long long *counts) {
    const uint32_t o = threadIdx.x;
    if (o >= world) return;
    const unsigned long long lo = tile_base[(long long)o * n_tiles];
    const unsigned long long hi = o + 1 < world ? tile_base[(long long)(o + 1) * n_tiles] : (unsigned long long)n;
    counts[o] = hi - lo;
}

__global__ void __launch_bounds__(kBlock) push_kernel(const uint32_t *tile_base, long long n, long long n_tiles, uint32_t world,
                                                      const uint4 *counts, const uint32_t *barcode, const unsigned long long *umi,
                                                      const uint32_t *gene, const int16_t *score, const uint16_t *conv_n,
                                                      const uint8_t *strand, const uint8_t *transcriptome, const PushPeers peers) {
    __shared__ uint32_t warp_count[kBlock / 32][DYN_PUSH_MAX_WORLD + 1];
    __shared__ uint32_t run_start[DYN_PUSH_MAX_WORLD + 2];       // first row of owner o's run inside the tile's staged rows
    __shared__ unsigned long long staged[kBlock * 5];
    const long long i = blockIdx.x * (long long)kBlock + threadIdx.x;
    const bool live = i < n;
    const uint32_t bc = live ? barcode[i] : 0u;
    const uint32_t o = live ? bc % world : world;
    const uint32_t rank = tile_rank(o, world, warp_count);
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (uint32_t q = 0; q <= world; ++q) {
            run_start[q] = acc;
            uint32_t c = 0;
            for (int w = 0; w < kBlock / 32; ++w) c += warp_count[w][q];
            acc += c;
        }
        run_start[world + 1] = acc;
    }
    __syncthreads();
    if (live) {
        ReadRow r;
        const uint4 cv = counts[i];
        r.counts[0] = cv.x; r.counts[1] = cv.y; r.counts[2] = cv.z; r.counts[3] = cv.w;
        r.umi = umi[i];
        r.barcode = bc;
        r.gene = gene[i];
        r.score = score[i];
        r.conv_n = conv_n[i];
        r.strand = strand[i];
        r.transcriptome = transcriptome ? transcriptome[i] : (uint8_t)1;
        r.pad[0] = r.pad[1] = 0;
        const unsigned long long *w = reinterpret_cast<const unsigned long long *>(&r);
        unsigned long long *d = staged + (size_t)(run_start[o] + rank) * 5;
#pragma unroll
        for (int k = 0; k < 5; ++k) d[k] = w[k];
    }
    __syncthreads();
    const uint32_t live_rows = run_start[world];                 // rows of real reads (the dummy owner `world` comes last)
    for (uint32_t word = threadIdx.x; word < live_rows * 5; word += kBlock) {
        const uint32_t row = word / 5;
        uint32_t q = 0;
        while (run_start[q + 1] <= row) ++q;                     // at most `world` steps, the same for most of a warp
        const long long dst_row = peers.base[q] + (long long)tile_base[(long long)q * n_tiles + blockIdx.x] + (row - run_start[q]);
        peers.rows[q][dst_row * 5 + (word - row * 5)] = staged[word];
    }
    __threadfence_system();
}

}  // namespace

extern "C" int dyn_owner_partition(dyn_ctx_t *ctx, const uint32_t *d_barcode, int64_t n, int world, uint32_t *d_perm,
                                   uint64_t *d_owner_counts, void *stream) {
    DYN_ARG(ctx != nullptr, "null context");
    CtxEnter enter_(ctx, stream);
    DYN_ARG(n >= 0 && n <= 0x7fffffffll, "n outside [0, 2^31)");
    DYN_ARG(world >= 1 && world <= DYN_MAX_WORLD, "world out of range");
    DYN_ARG(d_owner_counts != nullptr, "null counts");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DYN_CUDA(cudaMemsetAsync(d_owner_counts, 0, sizeof(uint64_t) * world, st));
    if (n == 0) return DYN_OK;
    DYN_ARG(d_barcode && d_perm, "null buffer");
    int bits = 0;
    while ((1 << bits) < world) ++bits;
    size_t temp = 0;
    DYN_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, temp, (const uint8_t *)nullptr, (uint8_t *)nullptr,
                                             (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)n, 0, bits > 0 ? bits : 1, st));
    Arena sizing;
    for (int pass = 0; pass < 2; ++pass) {
        void *base = nullptr;
        if (pass == 1) {
            int rc = dyn_ctx_scratch(ctx, 16, sizing.used + 256, &base);
            if (rc) return rc;
        }
        Arena a(base);
        uint8_t *owner = a.take<uint8_t>((size_t)n), *owner_sorted = a.take<uint8_t>((size_t)n);
        uint32_t *iota = a.take<uint32_t>((size_t)n);
        uint8_t *tmp = a.take<uint8_t>(temp);
        if (pass == 0) { sizing = a; continue; }
        const unsigned grid = std::min<unsigned>(grid_for(n), (unsigned)ctx->sm_count * 8);
        owner_kernel<<<grid, kBlock, 0, st>>>(d_barcode, n, (uint32_t)world, owner, iota,
                                              reinterpret_cast<unsigned long long *>(d_owner_counts));
        DYN_CUDA(cudaGetLastError());
        if (world == 1) {
            DYN_CUDA(cudaMemcpyAsync(d_perm, iota, sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, st));
        } else {
            DYN_CUDA(cub::DeviceRadixSort::SortPairs(tmp, temp, owner, owner_sorted, iota, d_perm, (int)n, 0, bits, st));
        }
    }
    return DYN_OK;
}


extern "C" int dyn_peer_alloc(dyn_ctx_t *ctx, int64_t bytes, void **d_ptr, void *handle_out) {
    DYN_ARG(ctx != nullptr, "null context");
    CtxEnter enter_(ctx, nullptr);
    DYN_ARG(bytes > 0 && d_ptr && handle_out, "bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == DYN_PEER_HANDLE_BYTES, "handle size");
    void *p = nullptr;
    DYN_CUDA(cudaMalloc(&p, (size_t)bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        DYN_CUDA(e);
    }
    memcpy(handle_out, &h, sizeof(h));
    *d_ptr = p;
    return DYN_OK;
}

extern "C" int dyn_peer_open(dyn_ctx_t *ctx, const void *handle, void **d_ptr) {
    DYN_ARG(ctx != nullptr, "null context");
    CtxEnter enter_(ctx, nullptr);
    DYN_ARG(handle && d_ptr, "bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    DYN_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return DYN_OK;
}

extern "C" int dyn_peer_close(dyn_ctx_t *ctx, void *d_ptr) {
    DYN_ARG(ctx != nullptr, "null context");
    CtxEnter enter_(ctx, nullptr);
    if (d_ptr) DYN_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return DYN_OK;
}

extern "C" int dyn_peer_free(dyn_ctx_t *ctx, void *d_ptr) {
    DYN_ARG(ctx != nullptr, "null context");
    CtxEnter enter_(ctx, nullptr);
    if (d_ptr) DYN_CUDA(cudaFree(d_ptr));
    return DYN_OK;
}

extern "C" int dyn_push_plan(dyn_ctx_t *ctx, const uint32_t *d_barcode, int64_t n, int world, uint32_t *d_tile_base,
                             uint64_t *d_owner_counts, void *stream) {
    DYN_ARG(ctx != nullptr, "null context");
    CtxEnter enter_(ctx, stream);
    DYN_ARG(n >= 0 && n <= 0x7fffffffll, "n outside [0, 2^31)");
    DYN_ARG(world >= 1 && world <= DYN_PUSH_MAX_WORLD, "world out of range");
    DYN_ARG(d_owner_counts != nullptr, "null counts");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n == 0) {
        DYN_CUDA(cudaMemsetAsync(d_owner_counts, 0, sizeof(uint64_t) * world, st));
        return DYN_OK;
    }
    DYN_ARG(d_barcode && d_tile_base, "null buffer");
    const long long n_tiles = (n + kPushTile - 1) / kPushTile;
    const long long cells = n_tiles * world;
    size_t temp = 0;
    DYN_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, temp, (uint32_t *)nullptr, (uint32_t *)nullptr, cells, st));
    void *tmp = nullptr;
    int rc = dyn_ctx_scratch(ctx, 16, temp + 256, &tmp);
    if (rc) return rc;
    tile_count_kernel<<<(unsigned)n_tiles, kBlock, 0, st>>>(d_barcode, n, (uint32_t)world, n_tiles, d_tile_base);
    DYN_CUDA(cudaGetLastError());
    DYN_CUDA(cub::DeviceScan::ExclusiveSum(tmp, temp, d_tile_base, d_tile_base, cells, st));
    owner_totals_kernel<<<1, DYN_PUSH_MAX_WORLD, 0, st>>>(d_tile_base, n, n_tiles, (uint32_t)world,
                                                         reinterpret_cast<unsigned long long *>(d_owner_counts));
    DYN_CUDA(cudaGetLastError());
    return DYN_OK;
}

extern "C" int dyn_push_reads(dyn_ctx_t *ctx, const uint32_t *d_tile_base, int64_t n, int world, const uint8_t *d_counts,
                              const uint32_t *d_barcode, const uint64_t *d_umi, const uint32_t *d_gene, const int16_t *d_score,
                              const uint16_t *d_conv_n, const uint8_t *d_strand, const uint8_t *d_transcriptome,
                              void *const *peer_rows, const int64_t *dst_base, void *stream) {
    DYN_ARG(ctx != nullptr, "null context");
    CtxEnter enter_(ctx, stream);
    DYN_ARG(n >= 0 && n <= 0x7fffffffll, "n outside [0, 2^31)");
    DYN_ARG(world >= 1 && world <= DYN_PUSH_MAX_WORLD, "world out of range");
    if (n == 0) return DYN_OK;
    DYN_ARG(d_tile_base && d_counts && d_barcode && d_umi && d_gene && d_score && d_conv_n && d_strand && peer_rows && dst_base, "null buffer");
    DYN_ARG((reinterpret_cast<uintptr_t>(d_counts) & 15) == 0, "counts must be 16-byte aligned");
    PushPeers peers;
    for (int o = 0; o < DYN_PUSH_MAX_WORLD; ++o) {
        peers.rows[o] = o < world ? static_cast<unsigned long long *>(peer_rows[o]) : nullptr;
        peers.base[o] = o < world ? (long long)dst_base[o] : 0;
        DYN_ARG(o >= world || (peer_rows[o] && (reinterpret_cast<uintptr_t>(peer_rows[o]) & 7) == 0), "peer buffer missing or misaligned");
    }
    const long long n_tiles = (n + kPushTile - 1) / kPushTile;
    push_kernel<<<(unsigned)n_tiles, kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
        d_tile_base, n, n_tiles, (uint32_t)world, reinterpret_cast<const uint4 *>(d_counts), d_barcode,
        reinterpret_cast<const unsigned long long *>(d_umi), d_gene, d_score, d_conv_n, d_strand, d_transcriptome, peers);
    DYN_CUDA(cudaGetLastError());
    return DYN_OK;
}

extern "C" int dyn_pack_reads(dyn_ctx_t *ctx, const uint32_t *d_perm, int64_t n, const uint8_t *d_counts,
                              const uint32_t *d_barcode, const uint64_t *d_umi, const uint32_t *d_gene, const int16_t *d_score,
                              const uint16_t *d_conv_n, const uint8_t *d_strand, const uint8_t *d_transcriptome, void *d_rows,
                              void *stream) {
    DYN_ARG(ctx != nullptr, "null context");
    CtxEnter enter_(ctx, stream);
    DYN_ARG(n >= 0, "negative n");
    if (n == 0) return DYN_OK;
    DYN_ARG(d_counts && d_barcode && d_umi && d_gene && d_score && d_conv_n && d_strand && d_rows, "null buffer");
    DYN_ARG((reinterpret_cast<uintptr_t>(d_counts) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_rows) & 7) == 0, "alignment");
    pack_kernel<<<grid_for(n), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
        d_perm, n, reinterpret_cast<const uint4 *>(d_counts), d_barcode, reinterpret_cast<const unsigned long long *>(d_umi),
        d_gene, d_score, d_conv_n, d_strand, d_transcriptome, static_cast<ReadRow *>(d_rows));
    DYN_CUDA(cudaGetLastError());
    return DYN_OK;
}

extern "C" int dyn_unpack_reads(dyn_ctx_t *ctx, const void *d_rows, int64_t n, uint8_t *d_counts, uint32_t *d_barcode,
                                uint64_t *d_umi, uint32_t *d_gene, int16_t *d_score, uint16_t *d_conv_n, uint8_t *d_strand,
                                uint8_t *d_transcriptome, void *stream) {
    DYN_ARG(ctx != nullptr, "null context");
    CtxEnter enter_(ctx, stream);
    DYN_ARG(n >= 0, "negative n");
    if (n == 0) return DYN_OK;
    DYN_ARG(d_rows && d_counts && d_barcode && d_umi && d_gene && d_score && d_conv_n && d_strand && d_transcriptome, "null buffer");
    DYN_ARG((reinterpret_cast<uintptr_t>(d_counts) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_rows) & 7) == 0, "alignment");
    unpack_kernel<<<grid_for(n), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const ReadRow *>(d_rows), n, reinterpret_cast<uint4 *>(d_counts), d_barcode,
        reinterpret_cast<unsigned long long *>(d_umi), d_gene, d_score, d_conv_n, d_strand, d_transcriptome);
    DYN_CUDA(cudaGetLastError());
    return DYN_OK;
}

extern "C" int dyn_merge_kn(dyn_ctx_t *ctx, const uint64_t *d_keys, const uint32_t *d_count, int64_t n, int world, int group_bits,
                            uint64_t *d_out_keys, uint32_t *d_out_count, int64_t *d_n_rows, void *stream) {
    DYN_ARG(ctx != nullptr, "null context");
    CtxEnter enter_(ctx, stream);
    DYN_ARG(n >= 0 && n < 0x7fffffffll, "bad n");
    DYN_ARG(world >= 1 && world <= DYN_MAX_WORLD, "world out of range");
    DYN_ARG(d_n_rows != nullptr, "null n_rows");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DYN_CUDA(cudaMemsetAsync(d_n_rows, 0, sizeof(int64_t), st));
    if (n == 0) return DYN_OK;
    DYN_ARG(d_keys && d_count && d_out_keys && d_out_count, "null buffer");
    int end_bit = 22 + (group_bits > 0 && group_bits <= 42 ? group_bits : 42);
    size_t t0 = 0, t1 = 0;
    {
        cub::DoubleBuffer<unsigned long long> k(nullptr, nullptr);
        cub::DoubleBuffer<uint32_t> v(nullptr, nullptr);
        DYN_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, t0, k, v, (int)n, 0, end_bit, st));
        DYN_CUDA(cub::DeviceReduce::ReduceByKey(nullptr, t1, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                                (const uint32_t *)nullptr, (uint32_t *)nullptr, (long long *)nullptr, cub::Sum(),
                                                (int)n, st));
    }
    const size_t temp = std::max(t0, t1);
    Arena sizing;
    for (int pass = 0; pass < 2; ++pass) {
        void *base = nullptr;
        if (pass == 1) {
            int rc = dyn_ctx_scratch(ctx, 16, sizing.used + 256, &base);
            if (rc) return rc;
        }
        Arena a(base);
        unsigned long long *ka = a.take<unsigned long long>((size_t)n), *kb = a.take<unsigned long long>((size_t)n);
        uint32_t *va = a.take<uint32_t>((size_t)n), *vb = a.take<uint32_t>((size_t)n);
        uint8_t *tmp = a.take<uint8_t>(temp);
        if (pass == 0) { sizing = a; continue; }
        localise_keys_kernel<<<grid_for(n), kBlock, 0, st>>>(reinterpret_cast<const unsigned long long *>(d_keys), n, (uint32_t)world, ka);
        DYN_CUDA(cudaMemcpyAsync(va, d_count, sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, st));
        cub::DoubleBuffer<unsigned long long> k(ka, kb);
        cub::DoubleBuffer<uint32_t> v(va, vb);
        size_t t = temp;
        DYN_CUDA(cub::DeviceRadixSort::SortPairs(tmp, t, k, v, (int)n, 0, end_bit, st));
        t = temp;
        DYN_CUDA(cub::DeviceReduce::ReduceByKey(tmp, t, k.Current(), reinterpret_cast<unsigned long long *>(d_out_keys), v.Current(),
                                                d_out_count, reinterpret_cast<long long *>(d_n_rows), cub::Sum(), (int)n, st));
    }
    return DYN_OK;
}
