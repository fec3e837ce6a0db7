// Device BAM ingest: BGZF inflate and record packing ON THE GPU (the follow-up the
// host ingest's ceiling calls for: zlib on the host cores tops out near 1.5e7 records/s per box, two orders of
// magnitude below what the tally takes from HBM).
//
// Replaces, for coordinate-sorted single-end BAMs (BGZF blocks that end on record boundaries, as htslib writes them, or
// anywhere, as STAR's sorted BAMs do), everything the reference gets from pysam / htslib between the file and the per-read loop of
// dynast/preprocessing/bam.py:253-312:
//   host    a reader thread copies the COMPRESSED bytes of the next run of BGZF blocks into page-locked staging
//           memory and lists the blocks (BSIZE / ISIZE hops, no inflate); cudaMemcpyAsync moves them -- PCIe carries
//           ~90 B per record instead of the ~280 B of the inflated stream;
//   k1      inflate_kernel: ONE WARP PER BGZF BLOCK (inflate_core.cuh: lane 0 decodes Huffman symbols through
//           shared-memory lookup tables, the warp writes literals lane-parallel and copies matches cooperatively);
//   k2      chain kernels: one thread per block finds the first record that starts in it, walks its records (count, then
//           positions after a scan); where one block's walk ends the next one's must begin (checked);
//   k3      parse_kernel: one thread per record -- read filter, aux-tag scan (MD / HI / AS / barcode / UMI / gene),
//           2-bit barcode / UMI keys, FNV hash of the gene id, packed size;
//   scan    unit index and record offset of every kept record (one cub::DeviceScan over count << 40 | size);
//   k4      pack_kernel: one thread per kept record writes the packed record (header, CIGAR, MD tokens with the
//           Phred folded in for the lean layout, 4-bit sequence) and the unit's keys; the gene id's hash is looked
//           up in the sorted table of the caller's gene list.
// The chunk that comes out is what dyn_bam_stream_next returns with keys = 1, but resident in HBM: the tally, the
// velocity kernel and everything downstream read it where it lies.  Files this path does not take (paired records)
// are reported with DYN_ERR_FORMAT; callers fall back to dyn_bam_stream_*.
#include <cub/cub.cuh>
#include <fcntl.h>
#include <zlib.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <chrono>
#include <condition_variable>
#include <memory>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "bamrec_core.cuh"
#include "common.cuh"
#include "inflate_core.cuh"
#include "inflate_lane.cuh"

struct dyn_bam_dev;
int issue_copy_if_free(dyn_bam_dev *d);
static int inflate_warps_per_sm();

namespace {

constexpr int kInfWarps = 4;          // warps per CTA of the inflate kernel (one Tables set each)

struct BlockDesc { uint32_t in_off, in_len, out_off, out_len; };

#ifndef DYN_INF_QUORUM
#define DYN_INF_QUORUM 8        // lanes that must be waiting at a block header before the header path is taken ...
#endif
#ifndef DYN_INF_ROUNDS
#define DYN_INF_ROUNDS 8         // Huffman codes per lane between two looks at the lanes' states
#endif
#ifndef DYN_INF_WAIT
#define DYN_INF_WAIT 32         // ... unless one of them has waited this many trips, or no lane is decoding (measured: 38-41 ->
                                // 33-36 ms per chunk against taking every header at once; 16 / 24 lanes or longer waits: no better)
#endif

// Pass 1 -- one LANE per BGZF block (inflate_lane.cuh): one-warp CTAs, each lane pulls blocks off the work counter and
// advances its own Huffman decoder one round per iteration, writing tokens; the lanes meet again after every round.
__global__ void __launch_bounds__(32) inflate_lanes_kernel(const uint8_t *in, uint32_t *tokens, uint32_t *n_tok, const BlockDesc *blocks,
                                                           int n_blocks, unsigned int *next_block, uint32_t *status, uint16_t *global_tables) {
    namespace dl = dinflate::lane;
    extern __shared__ uint16_t lane_tables[];                  // [dl::kUnitsAll][32], or none:
    // with the tables in global memory (L1 / L2 hold the entries in use) the warps per SM are not bound by shared memory:
    // the pass is bound by the latency of one lane's dependent instructions, so lanes in flight are what it needs
    uint16_t *tb = global_tables ? global_tables + (size_t)blockIdx.x * dl::kUnitsAll * 32 : lane_tables;
    const dl::Mem<32> m{tb + threadIdx.x};
    dl::Lane s;
    s.state = dl::ST_END; s.err = dinflate::OK; s.op = s.out_len = 0; s.npend = 0; s.nt = 0;
    int cur = -1;                                              // -1 nothing yet, -2 the counter ran out
    int waited = 0;
    for (;;) {
        if (s.state == dl::ST_END && cur != -2) {
            if (cur >= 0) {
                const int err = dl::finish(s);
                n_tok[cur] = err ? 0u : s.nt;                  // pass 2 skips a block that failed
                if (err) atomicOr(status, 1u << err);
            }
            const int b = (int)atomicAdd(next_block, 1u);
            if (b < n_blocks) {
                cur = b;
                const BlockDesc bd = blocks[b];
                dl::lane_init(s, in + bd.in_off, bd.in_len, tokens + bd.out_off, bd.out_len);
            } else cur = -2;
        }
        const bool live = cur >= 0 && s.state != dl::ST_END;
        const bool at_header = live && s.state == dl::ST_HEADER;
        bool go = at_header;
        if (DYN_INF_QUORUM > 1) {
            // the header path is thousands of instructions long: lanes wait so that several take it in one go
            const unsigned hdr = __ballot_sync(0xffffffffu, at_header);
            const unsigned busy = __ballot_sync(0xffffffffu, live && !at_header);
            const unsigned urgent = __ballot_sync(0xffffffffu, at_header && waited >= DYN_INF_WAIT);
            go = at_header && (busy == 0u || __popc(hdr) >= DYN_INF_QUORUM || urgent != 0u);
        }
        if (go) { dl::header(s, m); waited = 0; }
        else if (at_header) ++waited;
        else if (live) {
            // a few rounds per trip through the bookkeeping above (a lane that reaches a header or its block's end
            // idles for the rest of them: round() does nothing in those states)
#pragma unroll 1
            for (int k = 0; k < DYN_INF_ROUNDS; ++k) dl::round(s, m);
        }
        if (__all_sync(0xffffffffu, cur == -2)) break;
    }
}

// Pass 2 -- one WARP per block: 32 tokens at a time into bytes.  Output positions by a prefix sum over the token
// lengths; literals and matches whose source ends before the first match output of the batch (they depend on nothing
// the batch's other matches write) lane-parallel, the remaining matches in stream order with the whole warp on each.
constexpr int kLzWarps = 8;
__global__ void __launch_bounds__(kLzWarps * 32) lz_kernel(const uint32_t *tokens, const uint32_t *n_tok, uint8_t *out, const BlockDesc *blocks,
                                                          int n_blocks, unsigned int *next_block) {
    const int lane = threadIdx.x & 31;
    for (;;) {
        int b = 0;
        if (lane == 0) b = (int)atomicAdd(next_block, 1u);
        b = __shfl_sync(0xffffffffu, b, 0);
        if (b >= n_blocks) return;
        const BlockDesc bd = blocks[b];
        const uint32_t *tok = tokens + bd.out_off;
        const uint32_t nt = n_tok[b];
        uint8_t *dst = out + bd.out_off;
        uint32_t out_pos = 0;
        uint32_t e_next = lane < nt ? tok[lane] : 0u;
        for (uint32_t base = 0; base < nt; base += 32) {
            const uint32_t n = nt - base < 32u ? nt - base : 32u;
            const uint32_t e = e_next;
            if (base + 32 < nt) e_next = base + 32 + lane < nt ? tok[base + 32 + lane] : 0u;       // the next batch is on its way
            const bool live = (uint32_t)lane < n;
            const bool is_match = live && (e >> 31);
            const uint32_t my_len = live ? (is_match ? (e & 0x1ffu) : ((e >> 29) & 3u) + 1u) : 0u;
            uint32_t incl = my_len;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += v;
            }
            const uint32_t my_pos = out_pos + incl - my_len;
            if (live && !is_match) {
                dst[my_pos] = (uint8_t)e;
                if (my_len > 1) dst[my_pos + 1] = (uint8_t)(e >> 8);
                if (my_len > 2) dst[my_pos + 2] = (uint8_t)(e >> 16);
            }
            __syncwarp();
            const uint32_t all_matches = __ballot_sync(0xffffffffu, is_match);
            if (all_matches) {
                const uint32_t first_pos = __shfl_sync(0xffffffffu, my_pos, __ffs((int)all_matches) - 1);
                const uint32_t len = e & 0x1ffu, dist = (e >> 9) & 0xffffu;
                const bool ready = is_match && len <= 32u && my_pos - dist + len <= first_pos;
                if (ready) {
                    const uint8_t *src = dst + my_pos - dist;
                    for (uint32_t k = 0; k < len; ++k) dst[my_pos + k] = src[k];
                }
                __syncwarp();
                uint32_t mm = __ballot_sync(0xffffffffu, is_match && !ready);
                while (mm) {
                    const int i = __ffs((int)mm) - 1;
                    mm &= mm - 1;
                    const uint32_t ei = __shfl_sync(0xffffffffu, e, i), pi = __shfl_sync(0xffffffffu, my_pos, i);
                    const uint32_t li = ei & 0x1ffu, di = (ei >> 9) & 0xffffu;
                    const uint8_t *src = dst + pi - di;
                    if (di >= li) { for (uint32_t k = lane; k < li; k += 32) dst[pi + k] = src[k]; }
                    else { for (uint32_t k = lane; k < li; k += 32) dst[pi + k] = src[k % di]; }      // overlapping copy = a repeating pattern
                    __syncwarp();
                }
            }
            out_pos += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
}

__global__ void __launch_bounds__(kInfWarps * 32) inflate_kernel(const uint8_t *in, uint8_t *out, const BlockDesc *blocks, int n_blocks,
                                                                unsigned int *next_block, uint32_t *status) {
    __shared__ dinflate::Tables tables[kInfWarps];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    dinflate::Tables &t = tables[wid];
    for (;;) {
        int b = 0;
        if (lane == 0) b = (int)atomicAdd(next_block, 1u);
        b = __shfl_sync(0xffffffffu, b, 0);
        if (b >= n_blocks) return;
        const BlockDesc bd = blocks[b];
        uint8_t *dst = out + bd.out_off;
        dinflate::State s;
        dinflate::state_init(s, in + bd.in_off, bd.in_len, bd.out_len);
        if (bd.out_len == 0) s.phase = 3;
        for (;;) {
            int n = 0;
            uint32_t bytes = 0;
            if (lane == 0) {
                if (s.err == dinflate::OK && s.phase == 0) dinflate::begin_block(s, t);
                if (s.err == dinflate::OK && s.phase == 1) n = dinflate::decode_batch(s, t, &bytes);
            }
            __syncwarp();
            n = __shfl_sync(0xffffffffu, n, 0);
            const int err = __shfl_sync(0xffffffffu, s.err, 0);
            const int phase = __shfl_sync(0xffffffffu, s.phase, 0);
            const uint32_t out_pos = __shfl_sync(0xffffffffu, s.out_pos, 0);
            if (n > 0) {
                // ---- flush the queue: positions by a prefix sum over the symbol lengths
                const uint32_t e = lane < n ? t.batch[lane] : 0u;
                const bool is_match = lane < n && (e >> 31);
                const uint32_t my_len = lane < n ? (is_match ? (e & 0x1ffu) : 1u) : 0u;
                uint32_t incl = my_len;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += v;
                }
                const uint32_t my_pos = out_pos + incl - my_len;
                if (lane < n && !is_match) dst[my_pos] = (uint8_t)e;
                __syncwarp();
                const uint32_t all_matches = __ballot_sync(0xffffffffu, is_match);
                if (all_matches) {
                    // Matches are short (7 bytes on average in BAM data): a match whose source ends before the first match
                    // output of this batch depends on nothing the other matches write, so all such matches are copied at
                    // once, every lane walking its own match; the others (near or self-overlapping sources, long matches)
                    // follow in stream order with the whole warp on each.
                    const uint32_t first_pos = __shfl_sync(0xffffffffu, my_pos, __ffs((int)all_matches) - 1);
                    const uint32_t len = e & 0x1ffu, dist = (e >> 9) & 0xffffu;
                    const bool ready = is_match && len <= 32u && my_pos - dist + len <= first_pos;
                    const uint32_t max_len = __reduce_max_sync(0xffffffffu, ready ? len : 0u);
                    if (ready) {
                        const uint8_t *src = dst + my_pos - dist;
                        for (uint32_t k = 0; k < len; ++k) dst[my_pos + k] = src[k];
                    }
                    (void)max_len;
                    __syncwarp();
                    uint32_t mm = __ballot_sync(0xffffffffu, is_match && !ready);
                    while (mm) {
                        const int i = __ffs((int)mm) - 1;
                        mm &= mm - 1;
                        const uint32_t ei = __shfl_sync(0xffffffffu, e, i), pi = __shfl_sync(0xffffffffu, my_pos, i);
                        const uint32_t li = ei & 0x1ffu, di = (ei >> 9) & 0xffffu;
                        const uint8_t *src = dst + pi - di;
                        if (di >= li) { for (uint32_t k = lane; k < li; k += 32) dst[pi + k] = src[k]; }
                        else { for (uint32_t k = lane; k < li; k += 32) dst[pi + k] = src[k % di]; }      // overlapping copy = a repeating pattern
                        __syncwarp();
                    }
                }
            }
            if (phase == 2 && err == dinflate::OK) {
                // stored block: a plain copy by the warp
                const unsigned long long sp = __shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)s.br.p, 0);
                const uint32_t len = __shfl_sync(0xffffffffu, s.stored_left, 0);
                const uint8_t *src = reinterpret_cast<const uint8_t *>((uintptr_t)sp);
                for (uint32_t k = lane; k < len; k += 32) dst[out_pos + k] = src[k];
                __syncwarp();
                if (lane == 0) { s.br.p += len; s.out_pos += len; s.stored_left = 0; s.phase = s.last ? 3 : 0; }
            }
            if (lane == 0) s.out_pos += bytes;
            const int phase2 = __shfl_sync(0xffffffffu, s.phase, 0);
            if (err != dinflate::OK || phase2 == 3) break;
        }
        if (lane == 0) {
            uint32_t st = (uint32_t)s.err;
            if (st == 0 && s.out_pos != bd.out_len) st = dinflate::ERR_OUTPUT;
            if (st) atomicOr(status, 1u << st);
        }
        __syncwarp();
    }
}

// ---- record chain.  A chunk's inflated bytes form ONE stream S = [bytes carried over from the chunk before | block 0 |
// block 1 | ...]; htslib ends BGZF blocks on record boundaries, STAR's sorted BAMs do not (bgzf_write per record, no flush),
// so a record may start in one block and end blocks later -- or in the next chunk.
//   chain_start_kernel  one thread per block: the offset (in S) of the first record that STARTS in the block.  Block 0
//                       knows it (the stream begins at a record; first_skip at the start of a range); the others try their
//                       own first byte, then scan: a start is believed when four plausible records follow one another.
//   chain_walk_kernel   pass 0: walks the records that start in the block, counts them, leaves the offset where the next
//                       block's records begin; pass 1: writes the offsets.
//   chain_check_kernel  induction: where block b's walk ended is where the next block that has a start said it starts --
//                       so every guess is either the true chain or reported (status bit 8: the host reader takes the file).
//                       The last block's end is the start of the bytes to carry into the next chunk.
// Records that start at or after `limit` (blocks read beyond the end of a range to complete its last record) are not counted.
constexpr uint32_t kNoStart = 0xffffffffu;
constexpr uint32_t kCarryRoom = 1u << 20;      // head room in front of a chunk's blocks for the bytes carried over

__device__ __forceinline__ bool plausible_record(const uint8_t *p, unsigned long long avail, int n_ref, uint32_t *size) {
    if (avail < 36) return false;
    const uint32_t bs = bamrec::ld32(p);
    if (bs < 32 || bs > (1u << 26)) return false;
    const int32_t ref = (int32_t)bamrec::ld32(p + 4), pos = (int32_t)bamrec::ld32(p + 8), next_ref = (int32_t)bamrec::ld32(p + 24),
                  next_pos = (int32_t)bamrec::ld32(p + 28);
    if (ref < -1 || ref >= n_ref || pos < -1 || next_ref < -1 || next_ref >= n_ref || next_pos < -1) return false;
    const uint32_t l_name = p[12], n_cigar = bamrec::ld16(p + 16), l_seq = bamrec::ld32(p + 20);
    if (l_name < 1 || l_seq > (1u << 26)) return false;
    const unsigned long long fixed = 32ull + l_name + 4ull * n_cigar + ((unsigned long long)l_seq + 1) / 2 + l_seq;
    if (fixed > bs) return false;
    if (36ull + l_name <= avail && p[36 + l_name - 1] != 0) return false;
    *size = 4 + bs;
    return true;
}

__global__ void chain_start_kernel(const uint8_t *S, uint32_t carry_len, uint32_t stream_len, const BlockDesc *blocks, int n_blocks, uint32_t s0,
                                   int n_ref, uint32_t *start) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    const uint32_t lo = carry_len + blocks[b].out_off, hi = lo + blocks[b].out_len;
    if (b == 0) { start[0] = s0 < hi ? s0 : kNoStart; return; }
    uint32_t found = kNoStart;
    for (uint32_t o = lo; o < hi && found == kNoStart; ++o) {
        uint32_t q = o;
        int ok = 0;
        bool good = true;
        while (ok < 4) {
            uint32_t size = 0;
            if (q >= stream_len || stream_len - q < 36) { good = ok >= 1; break; }       // ran into the end of the stream
            if (!plausible_record(S + q, stream_len - q, n_ref, &size)) { good = false; break; }
            ++ok;
            if ((unsigned long long)q + size >= stream_len) break;
            q += size;
        }
        if (good && ok >= 1) found = o;
    }
    start[b] = found;
}

__global__ void chain_walk_kernel(const uint8_t *S, uint32_t carry_len, uint32_t stream_len, uint32_t limit, const BlockDesc *blocks, int n_blocks,
                                  const uint32_t *start, uint32_t *n_rec, uint32_t *end, const uint32_t *rec_base, uint32_t *rec_at, uint32_t *status) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    const uint32_t lo = carry_len + blocks[b].out_off;
    uint32_t hi = lo + blocks[b].out_len;
    if (hi > limit) hi = limit;
    uint32_t q = start[b], n = 0;
    uint32_t *dst = rec_at ? rec_at + rec_base[b] : nullptr;
    bool paired = false, bad = false;
    if (q != kNoStart) {
        while (q < hi) {
            if (stream_len - q < 4) break;                                  // the size field itself is cut: carried over
            const uint32_t bs = bamrec::ld32(S + q);
            if (bs < 32) { bad = true; break; }
            if ((unsigned long long)q + 4 + bs > stream_len) break;         // incomplete: carried over
            paired |= (bamrec::ld16(S + q + 4 + 14) & 0x1) != 0;
            if (dst) dst[n] = q;
            ++n;
            q += 4 + bs;
        }
    }
    if (!dst) {
        n_rec[b] = n;
        end[b] = q;
        if (bad) atomicOr(status, 1u << 16);
        if (paired) atomicOr(status, 1u << 9);                // paired records: mate pairing is the host path's
    }
}

__global__ void chain_check_kernel(const uint32_t *start, const uint32_t *end, int n_blocks, uint32_t limit, uint32_t *carry_start, uint32_t *status) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks || start[b] == kNoStart) return;
    int nb = b + 1;
    while (nb < n_blocks && start[nb] == kNoStart) ++nb;
    if (nb < n_blocks) {
        // a walk that stopped at the limit (or on an incomplete record) has nothing to hand on
        if (end[b] < limit && end[b] != start[nb]) atomicOr(status, 1u << 8);
    } else {
        *carry_start = end[b];              // the last block in which a record starts
    }
}

__global__ void parse_kernel(const uint8_t *data, const uint32_t *rec_at, uint32_t n_rec, bamrec::Tags tg, bamrec::Parsed *parsed,
                             unsigned long long *scan_in, uint32_t *status) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rec) return;
    bamrec::Parsed pr;
    bamrec::parse_record(data, rec_at[i], tg, pr);
    parsed[i] = pr;
    scan_in[i] = pr.size16 ? ((1ull << 40) | pr.size16) : 0ull;
    if (pr.error) atomicOr(status, 1u << (15 + pr.error));
}

struct UnitOut {
    uint8_t *blob;
    uint32_t *rec_off;
    unsigned long long *bc_key, *umi_key;
    int32_t *gene_idx, *score, *ref_id, *pos, *hi;
    uint16_t *flag;
    uint8_t *gx_assigned;
};

__global__ void pack_kernel(const uint8_t *data, const uint32_t *rec_at, uint32_t n_rec, const bamrec::Parsed *parsed,
                            const unsigned long long *scan, int lean, const unsigned long long *gene_hash, const int32_t *gene_of_hash,
                            int n_gene_hash, UnitOut o) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rec) return;
    const bamrec::Parsed pr = parsed[i];
    if (!pr.size16) return;
    const unsigned long long sc = scan[i];
    const uint32_t unit = (uint32_t)(sc >> 40);
    const unsigned long long off16 = sc & ((1ull << 40) - 1);
    bamrec::pack_record(data, rec_at[i], pr, lean, o.blob + (off16 << 4));
    o.rec_off[unit] = (uint32_t)off16;
    o.bc_key[unit] = pr.bc_key;
    o.umi_key[unit] = pr.umi_key;
    int32_t gi = -1;
    if (pr.gx_assigned) {
        gi = -2;
        int lo = 0, hi = n_gene_hash;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (gene_hash[mid] < pr.gx_hash) lo = mid + 1; else hi = mid; }
        if (lo < n_gene_hash && gene_hash[lo] == pr.gx_hash) gi = gene_of_hash[lo];
    }
    o.gene_idx[unit] = gi;
    o.score[unit] = pr.score;
    o.hi[unit] = pr.hi;
    o.gx_assigned[unit] = pr.gx_assigned;
    const uint8_t *rec = data + rec_at[i] + 4;
    o.ref_id[unit] = (int32_t)bamrec::ld32(rec);
    o.pos[unit] = (int32_t)bamrec::ld32(rec + 4);
    o.flag[unit] = (uint16_t)bamrec::ld16(rec + 14);
}

struct DevBuf {         // grow-only device buffer
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return DYN_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        const size_t want = bytes + bytes / 4 + 256;
        if (cudaMalloc(&p, want) != cudaSuccess) { cudaGetLastError(); dyn_set_error("dyn_bam_dev: device allocation of %zu bytes failed", want); return DYN_ERR_NOMEM; }
        cap = want;
        return DYN_OK;
    }
    ~DevBuf() { if (p) cudaFree(p); }
    template <typename T> T *as() { return static_cast<T *>(p); }
};

struct Staged {          // one run of BGZF blocks, compressed, in page-locked memory
    uint8_t *bytes = nullptr;
    size_t cap = 0, len = 0;
    std::vector<BlockDesc> blocks;
    size_t inflated = 0;
    uint32_t first_skip = 0;
    int n_extra = 0;         // trailing blocks that lie BEHIND the range: read so that the range's last record is whole
    bool last = false;
    int dev_slot = -1;       // >= 0: the compressed bytes are (being) copied to device buffer pair `dev_slot`
};

// Everything that is expensive to create (device buffers of several hundred MB, page-locked staging memory): a closed
// reader parks its set in the context, the next one on that context picks it up (cudaMalloc / cudaHostAlloc of these
// sizes take a few hundred milliseconds).
struct IngestBuffers {
    Staged staged[3];
    DevBuf d_comp[2], d_blocks[2];        // double buffered: chunk i + 1 is copied while chunk i is inflated
    cudaEvent_t ev_copy[2] = {nullptr, nullptr};
    DevBuf d_inflated, d_carry, d_start, d_end, d_tokens, d_ntok, d_nrec, d_recbase, d_recat, d_parsed, d_scan_in, d_scan, d_tmp, d_small;
    struct OutSet { DevBuf blob, rec_off, bc, umi, gene, score, ref_id, pos, hi, flag, gxa; } out[2];
    uint32_t *h_small = nullptr;          // page-locked landing zone
    ~IngestBuffers() {
        for (Staged &s : staged) if (s.bytes) cudaFreeHost(s.bytes);
        if (h_small) cudaFreeHost(h_small);
        for (cudaEvent_t e : ev_copy) if (e) cudaEventDestroy(e);
    }
};

void free_ingest_buffers(void *p) { delete static_cast<IngestBuffers *>(p); }

}  // namespace

struct dyn_bam_dev {
    dyn_ctx *ctx = nullptr;
    int fd = -1;
    size_t file_len = 0;
    // options
    bamrec::Tags tags{};
    int lean = 1;
    size_t chunk_bytes = 0xf0000000u;      // inflated bytes per chunk (below 4 GiB: 32-bit offsets), and
    int read_threads = 16;                 // pread threads of the reader (opts->n_threads; DYN_BAM_READ_THREADS overrides)
    size_t chunk_blocks = 0;               // BGZF blocks per chunk: ONE wave of the Huffman pass (a lane per block, 96 lanes per SM)
    // header / file layout (host)
    std::string header_text;
    std::vector<char> ref_names;
    std::vector<uint32_t> ref_name_off{0};
    std::vector<int32_t> ref_len;
    size_t pos = 0, end = 0;               // this part: the BGZF blocks that start in [pos, end) of the file
    bool first_chunk = true;
    double ratio = 0.5;                    // compressed / inflated bytes seen so far (sizes the next read)
    uint32_t first_skip = 0;
    // gene table
    DevBuf gene_hash, gene_of_hash;
    int n_gene_hash = 0;
    // reader thread
    std::thread reader;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Staged *> ready, free_list;
    bool finished = false, stop = false;
    int err = 0;
    std::string err_msg;
    IngestBuffers *buf = nullptr;         // shared intermediates + two sets of outputs + staging
    int cur = 0, copy_slot = 0;
    uint32_t carry_len = 0;                // bytes of an unfinished record carried from the previous chunk (in d_carry)
    bool slot_busy[2] = {false, false};    // (under mu) the device buffer pair holds a slab that is not inflated yet
    int64_t records_seen = 0;

    ~dyn_bam_dev() {
        {
            std::lock_guard<std::mutex> l(mu);
            stop = true;
        }
        cv.notify_all();
        if (reader.joinable()) reader.join();
        if (buf) {
            cudaDeviceSynchronize();      // nothing may still read (or fill) the buffers when the next reader takes them
            for (Staged &st : buf->staged) st.dev_slot = -1;
            std::lock_guard<std::recursive_mutex> l(ctx->mu);
            if (!ctx->ingest_cache) { ctx->ingest_cache = buf; ctx->ingest_cache_free = free_ingest_buffers; }
            else delete buf;
        }
        if (fd >= 0) close(fd);
    }
};

namespace {

bool pread_all(int fd, uint8_t *dst, size_t len, size_t off) {
    size_t got = 0;
    while (got < len) {
        const ssize_t r = pread(fd, dst + got, len - got, (off_t)(off + got));
        if (r <= 0) return false;
        got += (size_t)r;
    }
    return true;
}

// BGZF block header at p?  *blen = its length in the file
inline bool bgzf_header(const uint8_t *p, size_t avail, size_t *blen) {
    if (avail < 18 || p[0] != 31 || p[1] != 139 || p[2] != 8 || !(p[3] & 4)) return false;
    const uint32_t xlen = bamrec::ld16(p + 10);
    if (avail < 12 + (size_t)xlen) return false;
    for (size_t x = 12; x + 4 <= 12 + (size_t)xlen;) {
        const uint32_t slen = bamrec::ld16(p + x + 2);
        if (p[x] == 'B' && p[x + 1] == 'C' && slen == 2) { *blen = (size_t)bamrec::ld16(p + x + 4) + 1; return *blen >= 12 + (size_t)xlen + 8; }
        x += 4 + slen;
    }
    return false;
}

// The reader thread: reads the next slab of the file straight into page-locked staging memory and lists the BGZF
// blocks it holds by hopping over their BSIZE fields there -- the file is never scanned up front.
void read_ahead(dyn_bam_dev *d) {
    cudaSetDevice(d->ctx->device);
    for (;;) {
        Staged *s = nullptr;
        {
            std::unique_lock<std::mutex> l(d->mu);
            d->cv.wait(l, [&] { return d->stop || !d->free_list.empty(); });
            if (d->stop) return;
            if (d->pos >= d->end) { d->finished = true; d->cv.notify_all(); return; }
            s = d->free_list.front();
            d->free_list.pop_front();
        }
        // blocks of this chunk: at most one wave of the Huffman pass, and the rest of the range in equal parts (a last chunk
        // of a few blocks would take as long as a full one: the pass is bound by the latency of one block)
        size_t block_cap = d->chunk_blocks;
        if (block_cap) {
            const double est_blocks = (double)(d->end - d->pos) / (d->ratio * 65280.0) + 1.0;
            const double n_chunks = std::ceil(est_blocks / (double)block_cap);
            block_cap = std::min(block_cap, (size_t)std::ceil(est_blocks / n_chunks * 1.02) + 8);
        }
        const size_t inflated_cap = block_cap ? std::min(d->chunk_bytes, block_cap * (size_t)65536) : d->chunk_bytes;
        size_t want = std::min(d->file_len - d->pos, (size_t)((double)inflated_cap * d->ratio * 1.05) + (1u << 20));
        if (d->pos + want >= d->end) want = std::min(d->file_len - d->pos, std::max(want, d->end - d->pos + 3 * (size_t)65536));    // + the blocks behind the range
        bool ok = true;
        if (s->cap < want) {
            if (s->bytes) cudaFreeHost(s->bytes);
            s->cap = want + want / 4 + (1 << 20);
            if (cudaHostAlloc((void **)&s->bytes, s->cap, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); s->bytes = nullptr; s->cap = 0; ok = false; }
        }
        // If a device buffer pair is free the slab goes to the GPU stripe by stripe WHILE it is being read (the copy of a
        // 1 GB slab takes 18 ms over PCIe: behind the read it would be 18 ms in which the GPU has nothing to inflate)
        int stream_slot = -1;
        if (ok) {
            std::lock_guard<std::mutex> l(d->mu);
            if (!d->slot_busy[d->copy_slot]) { stream_slot = d->copy_slot; d->slot_busy[stream_slot] = true; d->copy_slot ^= 1; }
        }
        if (stream_slot >= 0 && d->buf->d_comp[stream_slot].ensure(want + 16) != DYN_OK) ok = false;
        if (ok) {
            // the file range is read by several threads: one pread stream does not saturate a page cache / NVMe
            static const int env_nt = [] { const char *e = getenv("DYN_BAM_READ_THREADS"); const int v = e ? atoi(e) : 0; return v < 0 ? 0 : (v > 64 ? 64 : v); }();
            const int nt = env_nt ? env_nt : d->read_threads;
            const size_t stripe = (size_t)128 << 20;
            for (size_t a0 = 0; ok && a0 < want; a0 += stripe) {
                const size_t len = std::min(stripe, want - a0);
                std::vector<std::thread> th;
                std::vector<char> good(nt, 1);
                for (int t = 0; t < nt; ++t)
                    th.emplace_back([&, t] {
                        const size_t a = a0 + len * t / nt, b = a0 + len * (t + 1) / nt;
                        good[t] = pread_all(d->fd, s->bytes + a, b - a, d->pos + a) ? 1 : 0;
                    });
                for (auto &t : th) t.join();
                for (char g : good) ok &= g != 0;
                if (ok && stream_slot >= 0 &&
                    cudaMemcpyAsync(d->buf->d_comp[stream_slot].as<uint8_t>() + a0, s->bytes + a0, len, cudaMemcpyHostToDevice, d->ctx->copy_stream[0]) != cudaSuccess)
                    ok = false;
            }
        }
        s->blocks.clear();
        s->inflated = 0;
        size_t q = 0;
        bool corrupt = false;
        while (ok && d->pos + q < d->end) {
            size_t blen = 0;
            if (want - q < 18) break;
            if (!bgzf_header(s->bytes + q, want - q, &blen)) { corrupt = true; break; }
            if (q + blen > want) break;                                  // the block continues behind this slab
            const uint32_t xlen = bamrec::ld16(s->bytes + q + 10);
            const uint32_t isize = bamrec::ld32(s->bytes + q + blen - 4);
            if (!s->blocks.empty() && (s->inflated + isize > d->chunk_bytes || (block_cap && s->blocks.size() >= block_cap))) break;
            s->blocks.push_back(BlockDesc{(uint32_t)(q + 12 + xlen), (uint32_t)(blen - 12 - xlen - 8), (uint32_t)s->inflated, isize});
            s->inflated += isize;
            q += blen;
        }
        if (ok && !corrupt && s->blocks.empty()) corrupt = true;        // not even one whole block in a slab of > 1 MiB
        const size_t q_range = q;
        s->n_extra = 0;
        if (ok && !corrupt && d->pos + q >= d->end) {
            // the range ends here: its last record may end in the blocks behind it (BGZF blocks need not end on record
            // boundaries) -- two more blocks come along; no record that STARTS in them is taken
            while (s->n_extra < 2 && d->pos + q < d->file_len) {
                size_t blen = 0;
                if (want - q < 18 || !bgzf_header(s->bytes + q, want - q, &blen) || q + blen > want) break;
                const uint32_t xlen = bamrec::ld16(s->bytes + q + 10);
                const uint32_t isize = bamrec::ld32(s->bytes + q + blen - 4);
                if (isize == 0) break;                                      // the end-of-file marker
                s->blocks.push_back(BlockDesc{(uint32_t)(q + 12 + xlen), (uint32_t)(blen - 12 - xlen - 8), (uint32_t)s->inflated, isize});
                s->inflated += isize;
                q += blen;
                ++s->n_extra;
            }
        }
        s->len = q;
        if (s->inflated) d->ratio = std::max(0.05, std::min(1.2, (double)q / (double)s->inflated));
        s->first_skip = d->first_chunk ? d->first_skip : 0;
        d->first_chunk = false;
        d->pos += q_range;
        s->last = d->pos >= d->end;
        std::lock_guard<std::mutex> l(d->mu);
        if (!ok || corrupt) {
            if (stream_slot >= 0) d->slot_busy[stream_slot] = false;
            d->err = corrupt ? DYN_ERR_FORMAT : DYN_ERR_IO;
            d->err_msg = corrupt ? "dyn_bam_dev: corrupt BGZF block" : "dyn_bam_dev: reading the BAM failed";
            d->finished = true; d->cv.notify_all();
            return;
        }
        if (stream_slot >= 0) {      // the compressed bytes are on their way already: block descriptors, then the event the inflate waits on
            IngestBuffers &B = *d->buf;
            if (B.d_blocks[stream_slot].ensure(sizeof(BlockDesc) * s->blocks.size()) != DYN_OK ||
                cudaMemcpyAsync(B.d_blocks[stream_slot].p, s->blocks.data(), sizeof(BlockDesc) * s->blocks.size(), cudaMemcpyHostToDevice, d->ctx->copy_stream[0]) != cudaSuccess ||
                cudaEventRecord(B.ev_copy[stream_slot], d->ctx->copy_stream[0]) != cudaSuccess) {
                d->err = DYN_ERR_CUDA; d->err_msg = "dyn_bam_dev: host to device copy failed"; d->finished = true;
            }
            s->dev_slot = stream_slot;
        }
        d->ready.push_back(s);
        // straight on to the GPU if a device buffer pair is free: the copy then runs under the inflate of the slab before
        if (issue_copy_if_free(d) != DYN_OK) { d->err = DYN_ERR_CUDA; d->err_msg = "dyn_bam_dev: host to device copy failed"; d->finished = true; }
        d->cv.notify_all();
    }
}

// Inflates BGZF blocks on the host from file offset `at` on until `need` inflated bytes are there (header parsing,
// record-boundary guess at a part's start).  *next = file offset behind the last block taken; lens = their inflated sizes.
bool host_inflate_from(dyn_bam_dev *d, size_t at, size_t need, std::vector<uint8_t> &out, std::vector<size_t> *lens, size_t *next) {
    std::vector<uint8_t> raw;
    while (out.size() < need && at < d->file_len) {
        uint8_t head[32];
        const size_t hn = std::min<size_t>(32, d->file_len - at);
        if (!pread_all(d->fd, head, hn, at)) return false;
        size_t blen = 0;
        if (!bgzf_header(head, hn, &blen) || at + blen > d->file_len) return false;
        raw.resize(blen);
        if (!pread_all(d->fd, raw.data(), blen, at)) return false;
        const uint32_t xlen = bamrec::ld16(raw.data() + 10), isize = bamrec::ld32(raw.data() + blen - 4);
        const size_t o = out.size();
        out.resize(o + isize);
        if (isize) {
            z_stream zs;
            memset(&zs, 0, sizeof(zs));
            if (inflateInit2(&zs, -15) != Z_OK) return false;
            zs.next_in = raw.data() + 12 + xlen; zs.avail_in = (uInt)(blen - 12 - xlen - 8);
            zs.next_out = out.data() + o; zs.avail_out = isize;
            const int rc = inflate(&zs, Z_FINISH);
            inflateEnd(&zs);
            if (rc != Z_STREAM_END) return false;
        }
        if (lens) lens->push_back(isize);
        at += blen;
    }
    if (next) *next = at;
    return true;
}

// First BGZF block that starts at or after file offset `from` (magic bytes, confirmed by the headers of the blocks behind it)
size_t find_block_start(dyn_bam_dev *d, size_t from) {
    if (from >= d->file_len) return d->file_len;
    const size_t n = std::min<size_t>(d->file_len - from, 4u << 16);
    std::vector<uint8_t> w(n);
    if (!pread_all(d->fd, w.data(), n, from)) return d->file_len;
    for (size_t o = 0; o + 18 <= n; ++o) {
        size_t q = o, hops = 0, blen = 0;
        while (q + 18 <= n && bgzf_header(w.data() + q, n - q, &blen)) { q += blen; ++hops; if (hops >= 3) break; }
        if (hops >= 3 || (hops >= 1 && from + q == d->file_len)) return from + o;
    }
    return d->file_len;
}

}  // namespace

// bamstream.cpp: the record-boundary guesser at a part's start
size_t dyn_bam_guess_offset(const uint8_t *data, size_t n, size_t len0, int32_t n_ref, bool at_eof);

extern "C" int dyn_bam_dev_open(dyn_ctx_t *ctx, const char *path, const dyn_bam_stream_opts_t *o, dyn_bam_dev_t **out) {
    DYN_ARG(ctx != nullptr && path && o && out, "null argument");
    CtxEnter enter_(ctx, nullptr, false);
    DYN_CUDA(cudaSetDevice(ctx->device));
    std::unique_ptr<dyn_bam_dev> d(new dyn_bam_dev());
    d->ctx = ctx;
    auto tag2 = [](const char *t) -> uint32_t { return (t && t[0] && t[1]) ? ((uint32_t)(uint8_t)t[0] | ((uint32_t)(uint8_t)t[1] << 8)) : 0u; };
    d->tags.bc = tag2(o->barcode_tag); d->tags.umi = tag2(o->umi_tag); d->tags.gx = tag2(o->gene_tag);
    d->tags.require_gene = o->require_gene != 0;
    d->lean = o->lean != 0;
    d->tags.lean = d->lean;
    if (o->chunk_bytes > 0) d->chunk_bytes = (size_t)o->chunk_bytes;
    d->chunk_bytes = std::min(d->chunk_bytes, (size_t)0xf0000000u);      // block offsets inside a chunk are 32 bits wide
    d->chunk_blocks = (size_t)ctx->sm_count * inflate_warps_per_sm() * 32;
    if (o->n_threads > 0) d->read_threads = std::min(o->n_threads, 64);
    d->fd = open(path, O_RDONLY);
    if (d->fd < 0) { dyn_set_error("dyn_bam_dev_open: cannot open %s", path); return DYN_ERR_IO; }
    {
        struct stat sb;
        if (fstat(d->fd, &sb) != 0) { dyn_set_error("dyn_bam_dev_open: cannot stat %s", path); return DYN_ERR_IO; }
        d->file_len = (size_t)sb.st_size;
    }
    // ---- header (inflated on the host: a few blocks), then this part's byte range of the file
    size_t c0 = 0;            // file offset of the block in which the first alignment record starts
    uint32_t skip0 = 0;       // ... and that record's offset inside it
    {
        std::vector<uint8_t> h;
        std::vector<size_t> lens;
        std::vector<size_t> starts;
        size_t at = 0;
        auto need = [&](size_t nbytes) -> bool {
            while (h.size() < nbytes && at < d->file_len) {
                starts.push_back(at);
                std::vector<size_t> one;
                size_t nx = at;
                if (!host_inflate_from(d.get(), at, h.size() + 1, h, &one, &nx) || nx == at) return false;
                // host_inflate_from may take several blocks when some are empty: record each of them
                size_t o = starts.back();
                starts.pop_back();
                for (size_t k = 0; k < one.size(); ++k) { starts.push_back(o); lens.push_back(one[k]); o = 0; }
                // only the first start is exact; empty blocks in a header are not expected -- recompute starts below
                at = nx;
            }
            return h.size() >= nbytes;
        };
        if (!need(12) || memcmp(h.data(), "BAM\1", 4) != 0) { dyn_set_error("dyn_bam_dev_open: %s is not a BAM file", path); return DYN_ERR_FORMAT; }
        size_t p = 4;
        const uint32_t l_text = bamrec::ld32(h.data() + p); p += 4;
        if (!need(p + l_text + 4)) { dyn_set_error("dyn_bam_dev_open: truncated header"); return DYN_ERR_FORMAT; }
        d->header_text.assign((const char *)h.data() + p, l_text); p += l_text;
        const uint32_t n_ref = bamrec::ld32(h.data() + p); p += 4;
        for (uint32_t r = 0; r < n_ref; ++r) {
            if (!need(p + 4)) { dyn_set_error("dyn_bam_dev_open: truncated reference list"); return DYN_ERR_FORMAT; }
            const uint32_t l_name = bamrec::ld32(h.data() + p); p += 4;
            if (!need(p + l_name + 4)) { dyn_set_error("dyn_bam_dev_open: truncated reference list"); return DYN_ERR_FORMAT; }
            d->ref_names.insert(d->ref_names.end(), (const char *)h.data() + p, (const char *)h.data() + p + (l_name ? l_name - 1 : 0));
            d->ref_name_off.push_back((uint32_t)d->ref_names.size());
            p += l_name;
            d->ref_len.push_back((int32_t)bamrec::ld32(h.data() + p)); p += 4;
        }
        // walk the blocks again to find the one that holds inflated offset p
        size_t fo = 0, io = 0;
        for (;;) {
            if (fo >= d->file_len) { c0 = d->file_len; skip0 = 0; break; }
            uint8_t head[32];
            const size_t hn = std::min<size_t>(32, d->file_len - fo);
            size_t blen = 0;
            if (!pread_all(d->fd, head, hn, fo) || !bgzf_header(head, hn, &blen)) { dyn_set_error("dyn_bam_dev_open: corrupt BGZF block in %s", path); return DYN_ERR_FORMAT; }
            uint8_t tail[4];
            if (!pread_all(d->fd, tail, 4, fo + blen - 4)) { dyn_set_error("dyn_bam_dev_open: truncated file"); return DYN_ERR_FORMAT; }
            const size_t isize = bamrec::ld32(tail);
            if (p < io + isize) { c0 = fo; skip0 = (uint32_t)(p - io); break; }
            io += isize; fo += blen;
            if (p == io) { c0 = fo; skip0 = 0; break; }
        }
    }
    DYN_ARG(o->n_parts >= 1 && o->part >= 0 && o->part < o->n_parts, "part out of range");
    const size_t span = d->file_len - c0;
    auto bound = [&](int part) -> size_t {
        if (part <= 0) return c0;
        if (part >= o->n_parts) return d->file_len;
        return find_block_start(d.get(), c0 + (size_t)((double)span * part / o->n_parts));
    };
    d->pos = bound(o->part);
    d->end = bound(o->part + 1);
    d->first_skip = o->part == 0 ? skip0 : 0;
    if (o->part > 0) {
        // the first record that STARTS in this part (blocks in which none starts belong to the previous part's last record)
        while (d->pos < d->end) {
            std::vector<uint8_t> buf;
            std::vector<size_t> lens;
            size_t nx = d->pos;
            if (!host_inflate_from(d.get(), d->pos, 1, buf, &lens, &nx)) { dyn_set_error("dyn_bam_dev_open: inflate failed in %s", path); return DYN_ERR_FORMAT; }
            const size_t len0 = lens.empty() ? 0 : lens[0];
            size_t first_next = d->pos;
            {   // file offset behind the first block
                uint8_t head[32];
                const size_t hn = std::min<size_t>(32, d->file_len - d->pos);
                size_t blen = 0;
                if (!pread_all(d->fd, head, hn, d->pos) || !bgzf_header(head, hn, &blen)) { dyn_set_error("dyn_bam_dev_open: corrupt BGZF block in %s", path); return DYN_ERR_FORMAT; }
                first_next = d->pos + blen;
            }
            buf.resize(len0);
            size_t nx2 = first_next;
            if (!host_inflate_from(d.get(), first_next, len0 + (2u << 16), buf, nullptr, &nx2)) { dyn_set_error("dyn_bam_dev_open: inflate failed in %s", path); return DYN_ERR_FORMAT; }
            const size_t off = dyn_bam_guess_offset(buf.data(), buf.size(), len0, (int32_t)d->ref_len.size(), nx2 >= d->file_len);
            if (off < len0) { d->first_skip = (uint32_t)off; break; }
            d->pos = first_next;
        }
    }
    int rc = DYN_OK;
    // ---- gene table: FNV hashes of the ids, sorted; two ids with one hash cannot be told apart on the device
    {
        std::vector<std::pair<uint64_t, int32_t>> hs;
        for (int64_t i = 0; i < o->n_gene_ids; ++i)
            hs.emplace_back(bamrec::fnv1a(reinterpret_cast<const uint8_t *>(o->gene_ids[i]), (uint32_t)strlen(o->gene_ids[i])), (int32_t)i);
        std::sort(hs.begin(), hs.end());
        for (size_t i = 1; i < hs.size(); ++i)
            if (hs[i].first == hs[i - 1].first) { dyn_set_error("dyn_bam_dev_open: two gene ids share a 64-bit hash"); return DYN_ERR_ARG; }
        std::vector<uint64_t> h(hs.size());
        std::vector<int32_t> g(hs.size());
        for (size_t i = 0; i < hs.size(); ++i) { h[i] = hs[i].first; g[i] = hs[i].second; }
        d->n_gene_hash = (int)hs.size();
        if ((rc = d->gene_hash.ensure(8 * std::max<size_t>(hs.size(), 1)))) return rc;
        if ((rc = d->gene_of_hash.ensure(4 * std::max<size_t>(hs.size(), 1)))) return rc;
        if (!hs.empty()) {
            DYN_CUDA(cudaMemcpy(d->gene_hash.p, h.data(), 8 * h.size(), cudaMemcpyHostToDevice));
            DYN_CUDA(cudaMemcpy(d->gene_of_hash.p, g.data(), 4 * g.size(), cudaMemcpyHostToDevice));
        }
    }
    if (ctx->ingest_cache) { d->buf = static_cast<IngestBuffers *>(ctx->ingest_cache); ctx->ingest_cache = nullptr; }
    else d->buf = new IngestBuffers();
    if (!d->buf->h_small) DYN_CUDA(cudaHostAlloc((void **)&d->buf->h_small, 256, cudaHostAllocDefault));
    for (cudaEvent_t &e : d->buf->ev_copy) if (!e) DYN_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    if ((rc = d->buf->d_small.ensure(256))) return rc;
    for (Staged &s : d->buf->staged) { s.dev_slot = -1; d->free_list.push_back(&s); }
    dyn_bam_dev *dp = d.release();
    dp->reader = std::thread(read_ahead, dp);
    *out = dp;
    return DYN_OK;
}

// Starts the host -> device copy of a staged run of blocks on the context's copy stream (device buffer pair `slot`).
// Warps (of 32 decoding lanes) per SM of the Huffman pass: up to 3 keep their tables in shared memory, more keep them in
// global memory (DYN_INF_WARPS).
static int inflate_warps_per_sm() {
    static const int w = [] { const char *e = getenv("DYN_INF_WARPS"); const int v = e ? atoi(e) : 12; return v < 1 ? 1 : (v > 24 ? 24 : v); }();
    return w;
}

// The inflate launch: two passes, Huffman with one lane per block then LZ77 with one warp per block (default), or
// DYN_INFLATE=warp for the one-pass one-warp-per-block kernel they replaced.  d_tokens: as many u32 as inflated bytes,
// d_ntok: one u32 per block, work: two zeroed counters.
static int launch_inflate(dyn_ctx *ctx, const uint8_t *in, uint8_t *out, const BlockDesc *blocks, int nb, uint32_t *d_tokens, uint32_t *d_ntok,
                          uint32_t *work, uint32_t *status, cudaStream_t stream) {
    static const bool by_warp = [] { const char *e = getenv("DYN_INFLATE"); return e && !strcmp(e, "warp"); }();
    if (by_warp) {
        const int ctas = std::min((nb + kInfWarps - 1) / kInfWarps, ctx->sm_count * 8);
        inflate_kernel<<<ctas, kInfWarps * 32, 0, stream>>>(in, out, blocks, nb, work, status);
    } else {
        const size_t table_bytes = sizeof(uint16_t) * dinflate::lane::kUnitsAll * 32;
        // Few blocks (a small file, one rank's share of a file split many ways): every block gets a lane even with three warps per
        // SM, and then the tables fit shared memory, where a lookup costs a third of an L1 / L2 one (22 vs 33 ms per wave)
        static const bool adaptive = getenv("DYN_INF_WARPS") == nullptr;
        int warps = inflate_warps_per_sm();
        if (adaptive && nb <= ctx->sm_count * 3 * 32) warps = 3;
        if (warps <= 3)       // (a per-device attribute: set on every launch, it is cheap)
            DYN_CUDA(cudaFuncSetAttribute(inflate_lanes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)table_bytes));
        const int ctas = std::min((nb + 31) / 32, ctx->sm_count * warps);
        uint16_t *tables = nullptr;
        if (warps > 3) {
            void *p = nullptr;
            int rc = dyn_ctx_scratch(ctx, 18, table_bytes * (size_t)ctas, &p);
            if (rc) return rc;
            tables = static_cast<uint16_t *>(p);
        }
        inflate_lanes_kernel<<<ctas, 32, tables ? 0 : table_bytes, stream>>>(in, d_tokens, d_ntok, blocks, nb, work, status, tables);
        DYN_CUDA(cudaGetLastError());
        const int lz_ctas = std::min((nb + kLzWarps - 1) / kLzWarps, ctx->sm_count * 8);
        lz_kernel<<<lz_ctas, kLzWarps * 32, 0, stream>>>(d_tokens, d_ntok, out, blocks, nb, work + 4);
    }
    DYN_CUDA(cudaGetLastError());
    return DYN_OK;
}

static int issue_copy(dyn_bam_dev *d, Staged *s, int slot) {
    IngestBuffers &B = *d->buf;
    int rc;
    if ((rc = B.d_comp[slot].ensure(s->len + 16))) return rc;
    if ((rc = B.d_blocks[slot].ensure(sizeof(BlockDesc) * s->blocks.size()))) return rc;
    cudaStream_t cs = d->ctx->copy_stream[0];
    DYN_CUDA(cudaMemcpyAsync(B.d_comp[slot].p, s->bytes, s->len, cudaMemcpyHostToDevice, cs));
    DYN_CUDA(cudaMemcpyAsync(B.d_blocks[slot].p, s->blocks.data(), sizeof(BlockDesc) * s->blocks.size(), cudaMemcpyHostToDevice, cs));
    DYN_CUDA(cudaEventRecord(B.ev_copy[slot], cs));
    s->dev_slot = slot;
    return DYN_OK;
}

// (d->mu held) starts the host to device copy of the first staged slab that has none yet, if the next buffer pair is free
int issue_copy_if_free(dyn_bam_dev *d) {
    for (Staged *s : d->ready) {
        if (s->dev_slot >= 0) continue;
        if (d->slot_busy[d->copy_slot]) return DYN_OK;
        int rc = issue_copy(d, s, d->copy_slot);
        if (rc) return rc;
        d->slot_busy[d->copy_slot] = true;
        d->copy_slot ^= 1;
    }
    return DYN_OK;
}

extern "C" int dyn_bam_dev_next(dyn_bam_dev_t *d, dyn_bam_dev_chunk_t *chunk, void *stream_) {
    if (!d || !chunk) { dyn_set_error("dyn_bam_dev_next: null argument"); return DYN_ERR_ARG; }
    memset(chunk, 0, sizeof(*chunk));
    CtxEnter enter_(d->ctx, stream_, false);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DYN_CUDA(cudaSetDevice(d->ctx->device));
    IngestBuffers &B = *d->buf;
    Staged *s = nullptr;
    static const bool timing = getenv("DYN_BAM_DEV_TIMING") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto t_enter = now();
    {
        std::unique_lock<std::mutex> l(d->mu);
        d->cv.wait(l, [&] { return !d->ready.empty() || d->finished; });
        if (d->ready.empty()) {
            if (d->err) { dyn_set_error("%s", d->err_msg.c_str()); return d->err; }
            chunk->end = 2;         // past the end of the range
            return DYN_OK;
        }
        s = d->ready.front();
        d->ready.pop_front();
    }
    struct Release {        // the staging buffer goes back to the reader when this call is done with it, and the device
        dyn_bam_dev *d; Staged *s;      // buffer pair (compressed bytes + block descriptors, read up to the index kernels) is free for the next copy
        ~Release() {
            std::lock_guard<std::mutex> l(d->mu);
            if (s->dev_slot >= 0) d->slot_busy[s->dev_slot] = false;
            s->dev_slot = -1;
            d->free_list.push_back(s);
            issue_copy_if_free(d);
            d->cv.notify_all();
        }
    } release{d, s};
    const int nb = (int)s->blocks.size();
    auto t_staged = now();
    int rc;
    if (s->dev_slot < 0) {      // not copied ahead: both buffer pairs were busy when the reader staged it
        std::lock_guard<std::mutex> l(d->mu);
        if ((rc = issue_copy(d, s, d->copy_slot))) return rc;
        d->slot_busy[d->copy_slot] = true;
        d->copy_slot ^= 1;
    }
    const int slot = s->dev_slot;
    // the chunk's stream S: [carried bytes | block 0 | block 1 | ...], the blocks inflated behind kCarryRoom bytes of head room
    const uint32_t carry_len = d->carry_len;
    if ((rc = B.d_inflated.ensure(kCarryRoom + s->inflated + 16))) return rc;
    if ((rc = B.d_carry.ensure(kCarryRoom))) return rc;
    if ((rc = B.d_start.ensure(4 * (size_t)(nb + 1))) || (rc = B.d_end.ensure(4 * (size_t)(nb + 1)))) return rc;
    uint8_t *const inflated_at = B.d_inflated.as<uint8_t>() + kCarryRoom;
    const uint8_t *const S = inflated_at - carry_len;
    if ((size_t)carry_len + s->inflated > 0xfffffff0ull) { dyn_set_error("dyn_bam_dev: chunk over 4 GiB"); return DYN_ERR_NOMEM; }
    const uint32_t stream_len = carry_len + (uint32_t)s->inflated;
    const uint32_t limit = s->n_extra > 0 ? carry_len + s->blocks[(size_t)(nb - s->n_extra)].out_off : stream_len;
    const uint32_t s0 = carry_len ? 0u : s->first_skip;
    if ((rc = B.d_nrec.ensure(4 * (size_t)(nb + 1)))) return rc;
    if ((rc = B.d_recbase.ensure(4 * (size_t)(nb + 1)))) return rc;
    uint32_t *d_small = B.d_small.as<uint32_t>();      // [0] inflate work counter, [1] status, [2..3] scan totals
    DYN_CUDA(cudaMemsetAsync(d_small, 0, 64, stream));
    static cudaEvent_t tev[4] = {nullptr, nullptr, nullptr, nullptr};
    if (timing && !tev[0]) for (auto &e : tev) cudaEventCreate(&e);
    if (timing) cudaEventRecord(tev[0], stream);
    DYN_CUDA(cudaStreamWaitEvent(stream, B.ev_copy[slot], 0));
    if (timing) cudaEventRecord(tev[1], stream);
    // ---- k1: inflate, one warp per block
    {
        if ((rc = B.d_tokens.ensure(4 * (s->inflated + 16)))) return rc;
        if ((rc = B.d_ntok.ensure(4 * (size_t)(nb + 1)))) return rc;
        if ((rc = launch_inflate(d->ctx, B.d_comp[slot].as<uint8_t>(), inflated_at, B.d_blocks[slot].as<BlockDesc>(), nb,
                                 B.d_tokens.as<uint32_t>(), B.d_ntok.as<uint32_t>(), d_small, d_small + 1, stream)))
            return rc;
    }
    if (timing) cudaEventRecord(tev[2], stream);
    // ---- k2: where the records start (per block, checked by induction), how many start in each block, their prefix sum
    if (carry_len) DYN_CUDA(cudaMemcpyAsync(inflated_at - carry_len, B.d_carry.p, carry_len, cudaMemcpyDeviceToDevice, stream));
    const int n_ref = (int)d->ref_len.size();
    B.h_small[8] = kNoStart;
    DYN_CUDA(cudaMemcpyAsync(d_small + 5, B.h_small + 8, 4, cudaMemcpyHostToDevice, stream));
    chain_start_kernel<<<(nb + 63) / 64, 64, 0, stream>>>(S, carry_len, stream_len, B.d_blocks[slot].as<BlockDesc>(), nb, s0, n_ref, B.d_start.as<uint32_t>());
    chain_walk_kernel<<<(nb + 127) / 128, 128, 0, stream>>>(S, carry_len, stream_len, limit, B.d_blocks[slot].as<BlockDesc>(), nb, B.d_start.as<uint32_t>(),
                                                            B.d_nrec.as<uint32_t>(), B.d_end.as<uint32_t>(), nullptr, nullptr, d_small + 1);
    chain_check_kernel<<<(nb + 127) / 128, 128, 0, stream>>>(B.d_start.as<uint32_t>(), B.d_end.as<uint32_t>(), nb, limit, d_small + 5, d_small + 1);
    DYN_CUDA(cudaGetLastError());
    DYN_CUDA(cudaMemsetAsync(B.d_nrec.as<uint32_t>() + nb, 0, 4, stream));
    size_t tmp_bytes = 0;
    DYN_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, B.d_nrec.as<uint32_t>(), B.d_recbase.as<uint32_t>(), nb + 1, stream));
    if ((rc = B.d_tmp.ensure(tmp_bytes))) return rc;
    DYN_CUDA(cub::DeviceScan::ExclusiveSum(B.d_tmp.p, tmp_bytes, B.d_nrec.as<uint32_t>(), B.d_recbase.as<uint32_t>(), nb + 1, stream));
    DYN_CUDA(cudaMemcpyAsync(B.h_small, B.d_recbase.as<uint32_t>() + nb, 4, cudaMemcpyDeviceToHost, stream));
    DYN_CUDA(cudaMemcpyAsync(B.h_small + 1, d_small + 1, 4, cudaMemcpyDeviceToHost, stream));
    DYN_CUDA(cudaMemcpyAsync(B.h_small + 5, d_small + 5, 4, cudaMemcpyDeviceToHost, stream));
    auto t_launched = now();
    if (timing) cudaEventRecord(tev[3], stream);
    DYN_CUDA(cudaStreamSynchronize(stream));
    auto t_inflated = now();
    if (timing) {
        float a = 0, b = 0, c = 0;
        cudaEventElapsedTime(&a, tev[0], tev[1]); cudaEventElapsedTime(&b, tev[1], tev[2]); cudaEventElapsedTime(&c, tev[2], tev[3]);
        fprintf(stderr, "   on the stream: wait for the H2D copy %.2f ms, inflate (both passes) %.2f, index + scan %.2f; host: launches done after %.2f ms\n", a, b, c,
                std::chrono::duration<double, std::milli>(t_launched - t_staged).count());
    }
    const uint32_t n_rec = B.h_small[0];
    uint32_t status = B.h_small[1];
    auto status_error = [&](uint32_t st) -> int {
        if (st & 0xeu) { dyn_set_error("dyn_bam_dev: inflate failed (corrupt BGZF block, status 0x%x)", st); return DYN_ERR_FORMAT; }
        if (st & (1u << 8)) { dyn_set_error("dyn_bam_dev: the record chain could not be established block by block (straddle): use the host reader (dyn_bam_stream_*)"); return DYN_ERR_FORMAT; }
        if (st & (1u << 9)) { dyn_set_error("dyn_bam_dev: paired records: use the host reader (dyn_bam_stream_*)"); return DYN_ERR_FORMAT; }
        if (st & (1u << 16)) { dyn_set_error("dyn_bam_dev: corrupt record or aux data"); return DYN_ERR_FORMAT; }
        if (st & (1u << 17)) { dyn_set_error("dyn_bam_dev: alignment without HI / AS tag (KeyError in the reference, bam.py:295-296)"); return DYN_ERR_FORMAT; }
        if (st & (1u << 18)) { dyn_set_error("dyn_bam_dev: alignment without MD tag (count.py:119-124)"); return DYN_ERR_FORMAT; }
        if (st & (1u << 19)) { dyn_set_error("dyn_bam_dev: record too long for the packed layout"); return DYN_ERR_FORMAT; }
        return DYN_OK;
    };
    if ((rc = status_error(status))) return rc;
    // what is left of the stream behind the last whole record goes to the next chunk (a record that starts in this chunk
    // and ends in the next); the last chunk of a range has read the blocks behind the range for that
    uint32_t carry_start = B.h_small[5];
    if (carry_start == kNoStart) {
        if (stream_len > s0) { dyn_set_error("dyn_bam_dev: no alignment record starts in a chunk of %u bytes", stream_len); return DYN_ERR_FORMAT; }
        carry_start = stream_len;
    }
    uint32_t carry_next = 0;
    if (!s->last) {
        carry_next = stream_len - std::min(carry_start, stream_len);
        if (carry_next > kCarryRoom) { dyn_set_error("dyn_bam_dev: an alignment record of more than %u bytes straddles chunks", kCarryRoom); return DYN_ERR_FORMAT; }
    } else if (carry_start < limit) {
        dyn_set_error("dyn_bam_dev: the last alignment record of the range is cut off (truncated file, or a record longer than the %d blocks read behind the range)", s->n_extra);
        return DYN_ERR_FORMAT;
    }
    auto keep_carry = [&]() -> int {      // on the stream, behind every kernel that reads this chunk's stream
        if (carry_next) DYN_CUDA(cudaMemcpyAsync(B.d_carry.p, S + carry_start, carry_next, cudaMemcpyDeviceToDevice, stream));
        d->carry_len = carry_next;
        return DYN_OK;
    };
    d->records_seen += n_rec;
    chunk->n_records_seen = n_rec;
    chunk->end = s->last ? 1 : 0;
    IngestBuffers::OutSet &os = B.out[d->cur];
    d->cur ^= 1;
    if (n_rec == 0) return keep_carry();
    // ---- k2 (positions), k3 (parse), scan
    if ((rc = B.d_recat.ensure(4 * (size_t)n_rec))) return rc;
    if ((rc = B.d_parsed.ensure(sizeof(bamrec::Parsed) * (size_t)n_rec))) return rc;
    if ((rc = B.d_scan_in.ensure(8 * (size_t)(n_rec + 1)))) return rc;
    if ((rc = B.d_scan.ensure(8 * (size_t)(n_rec + 1)))) return rc;
    chain_walk_kernel<<<(nb + 127) / 128, 128, 0, stream>>>(S, carry_len, stream_len, limit, B.d_blocks[slot].as<BlockDesc>(), nb, B.d_start.as<uint32_t>(),
                                                            nullptr, nullptr, B.d_recbase.as<uint32_t>(), B.d_recat.as<uint32_t>(), d_small + 1);
    DYN_CUDA(cudaGetLastError());
    parse_kernel<<<(n_rec + 127) / 128, 128, 0, stream>>>(S, B.d_recat.as<uint32_t>(), n_rec, d->tags,
                                                          B.d_parsed.as<bamrec::Parsed>(), B.d_scan_in.as<unsigned long long>(), d_small + 1);
    DYN_CUDA(cudaGetLastError());
    DYN_CUDA(cudaMemsetAsync(B.d_scan_in.as<unsigned long long>() + n_rec, 0, 8, stream));
    DYN_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, B.d_scan_in.as<unsigned long long>(), B.d_scan.as<unsigned long long>(), (int)(n_rec + 1), stream));
    if ((rc = B.d_tmp.ensure(tmp_bytes))) return rc;
    DYN_CUDA(cub::DeviceScan::ExclusiveSum(B.d_tmp.p, tmp_bytes, B.d_scan_in.as<unsigned long long>(), B.d_scan.as<unsigned long long>(), (int)(n_rec + 1), stream));
    DYN_CUDA(cudaMemcpyAsync(B.h_small + 2, B.d_scan.as<unsigned long long>() + n_rec, 8, cudaMemcpyDeviceToHost, stream));
    DYN_CUDA(cudaMemcpyAsync(B.h_small + 1, d_small + 1, 4, cudaMemcpyDeviceToHost, stream));
    DYN_CUDA(cudaStreamSynchronize(stream));
    auto t_parsed = now();
    status = B.h_small[1];
    if ((rc = status_error(status))) return rc;
    const unsigned long long tot = *reinterpret_cast<unsigned long long *>(B.h_small + 2);
    const uint32_t n_units = (uint32_t)(tot >> 40);
    const unsigned long long total16 = tot & ((1ull << 40) - 1);
    if (total16 > 0xfffffff0ull) { dyn_set_error("dyn_bam_dev: chunk over 64 GiB of packed records"); return DYN_ERR_NOMEM; }
    // ---- k4: pack
    const size_t nu = std::max<uint32_t>(n_units, 1);
    if ((rc = os.blob.ensure((size_t)total16 * 16 + 16)) || (rc = os.rec_off.ensure(4 * (nu + 1))) || (rc = os.bc.ensure(8 * nu)) ||
        (rc = os.umi.ensure(8 * nu)) || (rc = os.gene.ensure(4 * nu)) || (rc = os.score.ensure(4 * nu)) || (rc = os.ref_id.ensure(4 * nu)) ||
        (rc = os.pos.ensure(4 * nu)) || (rc = os.hi.ensure(4 * nu)) || (rc = os.flag.ensure(2 * nu)) || (rc = os.gxa.ensure(nu)))
        return rc;
    UnitOut uo;
    uo.blob = os.blob.as<uint8_t>(); uo.rec_off = os.rec_off.as<uint32_t>();
    uo.bc_key = os.bc.as<unsigned long long>(); uo.umi_key = os.umi.as<unsigned long long>();
    uo.gene_idx = os.gene.as<int32_t>(); uo.score = os.score.as<int32_t>(); uo.ref_id = os.ref_id.as<int32_t>();
    uo.pos = os.pos.as<int32_t>(); uo.hi = os.hi.as<int32_t>(); uo.flag = os.flag.as<uint16_t>(); uo.gx_assigned = os.gxa.as<uint8_t>();
    if (n_units) {
        pack_kernel<<<(n_rec + 127) / 128, 128, 0, stream>>>(S, B.d_recat.as<uint32_t>(), n_rec, B.d_parsed.as<bamrec::Parsed>(),
                                                             B.d_scan.as<unsigned long long>(), d->lean, d->gene_hash.as<unsigned long long>(),
                                                             d->gene_of_hash.as<int32_t>(), d->n_gene_hash, uo);
        DYN_CUDA(cudaGetLastError());
    }
    const uint32_t end16 = (uint32_t)total16;
    if ((rc = keep_carry())) return rc;
    DYN_CUDA(cudaMemcpyAsync(uo.rec_off + n_units, &end16, 4, cudaMemcpyHostToDevice, stream));
    DYN_CUDA(cudaStreamSynchronize(stream));       // (&end16 is a stack variable; the staging buffer is released on return)
    chunk->n_units = n_units;
    chunk->records_bytes = (int64_t)total16 * 16;
    chunk->d_records = uo.blob; chunk->d_rec_off = uo.rec_off;
    chunk->d_barcode_key = reinterpret_cast<uint64_t *>(uo.bc_key); chunk->d_umi_key = reinterpret_cast<uint64_t *>(uo.umi_key);
    chunk->d_gene_idx = uo.gene_idx; chunk->d_score = uo.score; chunk->d_ref_id = uo.ref_id; chunk->d_pos = uo.pos; chunk->d_hi = uo.hi;
    chunk->d_flag = uo.flag; chunk->d_gx_assigned = uo.gx_assigned;
    chunk->compressed_bytes = (int64_t)s->len;
    chunk->inflated_bytes = (int64_t)s->inflated;
    if (timing) {
        auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        fprintf(stderr, "dyn_bam_dev_next: wait for the reader %.2f ms, H2D + inflate + index %.2f, parse + scan %.2f, pack %.2f  (%d blocks, %u records)\n",
                ms(t_enter, t_staged), ms(t_staged, t_inflated), ms(t_inflated, t_parsed), ms(t_parsed, now()), nb, n_rec);
    }
    return DYN_OK;
}

extern "C" void dyn_bam_dev_close(dyn_bam_dev_t *d) {
    if (!d) return;
    cudaSetDevice(d->ctx->device);
    delete d;
}

extern "C" int64_t dyn_bam_dev_n_refs(const dyn_bam_dev_t *d) { return d ? (int64_t)d->ref_len.size() : 0; }

extern "C" const void *dyn_bam_dev_field(const dyn_bam_dev_t *d, int field, int64_t *n_bytes) {
    if (!d) return nullptr;
    const void *p = nullptr;
    size_t nb = 0;
    switch (field) {
        case 9: p = d->ref_len.data(); nb = d->ref_len.size() * 4; break;
        case 10: p = d->header_text.data(); nb = d->header_text.size(); break;
        case 28: p = d->ref_names.data(); nb = d->ref_names.size(); break;
        case 29: p = d->ref_name_off.data(); nb = d->ref_name_off.size() * 4; break;
        default: return nullptr;
    }
    if (n_bytes) *n_bytes = (int64_t)nb;
    return p;
}

// Test / tool entry: inflate BGZF-style raw deflate blocks that already sit in device memory (one warp per block).
extern "C" int dyn_inflate_blocks(dyn_ctx_t *ctx, const void *d_in, void *d_out, const uint32_t *d_block_desc, int32_t n_blocks,
                                  uint32_t *d_work, uint32_t *d_status, void *stream) {
    DYN_ARG(ctx != nullptr, "null context");
    CtxEnter enter_(ctx, stream);
    DYN_ARG(n_blocks >= 0, "negative n_blocks");
    if (n_blocks == 0) return DYN_OK;
    DYN_ARG(d_in && d_out && d_block_desc && d_work && d_status, "null buffer");
    // scratch of the two-pass decoder: work counters, tokens per block, one token word per output byte (the descriptors are
    // read back for the output size: a test / tool entry, not the ingest path)
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    std::vector<BlockDesc> h((size_t)n_blocks);
    DYN_CUDA(cudaMemcpyAsync(h.data(), d_block_desc, sizeof(BlockDesc) * (size_t)n_blocks, cudaMemcpyDeviceToHost, st));
    DYN_CUDA(cudaStreamSynchronize(st));
    size_t out_bytes = 0;
    for (const BlockDesc &bd : h) out_bytes = std::max(out_bytes, (size_t)bd.out_off + bd.out_len);
    void *scratch = nullptr;
    const size_t ntok_at = 256, tok_at = (ntok_at + 4 * (size_t)n_blocks + 255) & ~(size_t)255;
    int rc = dyn_ctx_scratch(ctx, 17, tok_at + 4 * (out_bytes + 16), &scratch);
    if (rc) return rc;
    uint8_t *base = static_cast<uint8_t *>(scratch);
    DYN_CUDA(cudaMemsetAsync(base, 0, 64, st));
    DYN_CUDA(cudaMemsetAsync(d_work, 0, 4, st));
    return launch_inflate(ctx, static_cast<const uint8_t *>(d_in), static_cast<uint8_t *>(d_out), reinterpret_cast<const BlockDesc *>(d_block_desc),
                          n_blocks, reinterpret_cast<uint32_t *>(base + tok_at), reinterpret_cast<uint32_t *>(base + ntok_at),
                          reinterpret_cast<uint32_t *>(base), d_status, st);
}
