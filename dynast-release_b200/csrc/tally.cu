// Per-read conversion tally (SURVEY.md section 8 rows a1 + a3).
//
// Replaces the inner loops of
//   dynast/preprocessing/bam.py:357-429        CIGAR/MD walk, reference base content, mismatch records
//   dynast/preprocessing/conversion.py:403-432 quality / SNP filter and the 12 + 4 counters
// which in the reference run as CPython per-base loops plus a CSV round trip.
//
// Design (HBM-bound integer/byte work, no tensor cores):
//   * one thread per read ("warp per read batch": a warp owns 32 consecutive reads), persistent CTAs,
//     a tile = `tile_reads` consecutive packed records;
//   * the tile's bytes are one contiguous range of the record blob, fetched by ONE TMA 1-D bulk copy
//     (cp.async.bulk -> UBLKCP) into shared memory and awaited on an mbarrier; 4 CTAs share an SM so
//     while one CTA waits for its tile the others walk theirs (multi-buffering across CTAs);
//   * base content = one fixed-trip pass over the whole query, 8 bases per 32-bit word with nibble one-hot
//     bit planes (no popc, no masks), corrected per MD mismatch / deletion / clipped base, so lanes only
//     diverge on the number of MD tokens;
//   * counters leave as one 128-bit store per read (consecutive threads -> consecutive 16 B: coalesced);
//   * mismatch records are placed with one global atomic per CTA tile (block scan of per-read counts),
//     then written by a second, cheaper walk over the same shared-memory bytes.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

struct TallyParams {
    const uint8_t *records;
    const uint32_t *rec_off;
    long long n_reads;
    const unsigned long long *snps;
    long long n_snps;
    int quality;
    uint32_t flags;
    uint8_t *counts;
    uint32_t *conv_off;
    uint16_t *conv_n;
    unsigned long long *conv_rec;
    uint32_t *conv_read;
    long long conv_cap;
    unsigned long long *conv_total;
    uint32_t *status;
    int tile_reads;
    uint32_t stage_bytes;
    long long n_tiles;
};

// BAM 4-bit code of an upper-case IUPAC letter, indexed by (letter & 31); 0 = not a base letter.
__constant__ uint8_t kLetterCode[32] = {0, 1,  14, 2, 13, 0, 0, 4,  11, 0, 0, 12, 0, 3, 15, 0,
                                        0, 0,  5,  6, 8,  0, 7, 9,  0,  10, 0, 0,  0, 0, 0,  0};
//                                      @  A   B   C  D   E  F  G   H   I  J  K   L  M  N   O
//                                      P  Q   R   S  T   U  V  W   X   Y  Z

struct Counters {
    uint32_t a, c, g, t;      // reference base content
    uint32_t w0, w1, w2;      // 12 conversion counters, 8 bits each
    uint32_t n_conv;          // every mismatch (conversions.csv rows), unfiltered
    uint32_t status;
};

__device__ __forceinline__ bool is_digit(uint32_t c) { return (c - '0') <= 9u; }

__device__ __forceinline__ bool snp_lookup(const unsigned long long *snps, long long n, unsigned long long key) {
    long long lo = 0, hi = n;
    while (lo < hi) {
        long long mid = (lo + hi) >> 1;
        unsigned long long v = __ldg(snps + mid);
        if (v < key) lo = mid + 1; else hi = mid;
    }
    return lo < n && __ldg(snps + lo) == key;
}

// Count A/C/G/T over the WHOLE query of a record: a fixed-trip loop (same trip count for every lane
// of a warp when read lengths agree), no masks (the packer zero-pads the last word and code 0 is not a
// base), no popc in the loop: one-hot bit planes are summed per nibble and folded every 15 words.
__device__ __forceinline__ uint32_t fold_nibbles(uint32_t acc) {
    acc = (acc & 0x0f0f0f0fu) + ((acc >> 4) & 0x0f0f0f0fu);
    return (acc * 0x01010101u) >> 24;
}

__device__ __forceinline__ void count_query(const uint32_t *seqw, int l_seq, Counters &cn) {
    const int n_words = (l_seq + 7) >> 3;
    for (int w0 = 0; w0 < n_words; w0 += 15) {
        uint32_t aa = 0, ac = 0, ag = 0, at = 0;
        const int w1 = min(n_words, w0 + 15);
        for (int w = w0; w < w1; ++w) {
            const uint32_t x = seqw[w];
            const uint32_t b0 = x & 0x11111111u, b1 = (x >> 1) & 0x11111111u, b2 = (x >> 2) & 0x11111111u,
                           b3 = (x >> 3) & 0x11111111u;
            const uint32_t s = b0 + b1 + b2 + b3;
            const uint32_t onehot = s & ~(s >> 1) & ~(s >> 2);
            aa += b0 & onehot; ac += b1 & onehot; ag += b2 & onehot; at += b3 & onehot;
        }
        cn.a += fold_nibbles(aa); cn.c += fold_nibbles(ac); cn.g += fold_nibbles(ag); cn.t += fold_nibbles(at);
    }
}

// Walk one record. EMIT == false: fill `cn`. EMIT == true: write mismatch records to out_rec/out_read.
//
// Content = bases of the whole query (count_query) - bases under I/S ops + per-mismatch correction
// (+reference base, -read base) + deleted reference bases: the common path has no per-run work, so
// lanes only diverge on the number of MD tokens, not on run lengths.
template <bool EMIT>
__device__ __forceinline__ void walk_record(const uint8_t *rec, const TallyParams &p, Counters &cn,
                                            unsigned long long *out_rec, uint32_t *out_read, uint32_t read_idx) {
    const uint4 h = *reinterpret_cast<const uint4 *>(rec);
    int rpos = (int)h.x;
    const uint32_t ref_id = h.y;
    const int l_seq = h.z & 0xffff, n_cigar = h.z >> 16;
    const int md_len = h.w & 0xffff;
    const uint32_t *cig = reinterpret_cast<const uint32_t *>(rec + 16);
    const uint32_t *seqw = cig + n_cigar;
    const uint8_t *seqb = reinterpret_cast<const uint8_t *>(seqw);
    const uint8_t *qual = seqb + ((((l_seq + 1) >> 1) + 3) & ~3);
    const uint8_t *md = qual + l_seq;

    if (!EMIT) count_query(seqw, l_seq, cn);
    // signed corrections on top of count_query
    int da = 0, dc = 0, dg = 0, dt = 0;

    int qpos = 0, mdp = 0, run = 0;
    uint32_t emitted = 0;
    bool bad = false;

    for (int ci = 0; ci < n_cigar && !bad; ++ci) {
        const uint32_t c = cig[ci];
        const int op = c & 0xf;
        int len = (int)(c >> 4);
        if (op == 0 || op == 7 || op == 8) {  // M, =, X
            if (qpos + len > l_seq) { bad = true; break; }
            for (;;) {
                if (run >= len) { run -= len; qpos += len; rpos += len; break; }
                qpos += run; rpos += run; len -= run; run = 0;
                if (mdp >= md_len) { bad = true; break; }
                uint32_t ch = md[mdp];
                if (is_digit(ch)) {
                    int v = 0;
                    do { v = v * 10 + (int)(ch - '0'); ++mdp; } while (mdp < md_len && is_digit(ch = md[mdp]));
                    run = v;
                    if (v == 0 && mdp >= md_len) { bad = true; break; }
                    continue;
                }
                if (ch == '^') { bad = true; break; }
                // mismatch: the MD letter is the reference base
                const uint32_t rn = kLetterCode[ch & 31];
                const uint32_t qn = (seqb[qpos >> 1] >> ((~qpos & 1) << 2)) & 0xf;
                if (!EMIT) {
                    da += (int)(rn == 1) - (int)(qn == 1); dc += (int)(rn == 2) - (int)(qn == 2);
                    dg += (int)(rn == 4) - (int)(qn == 4); dt += (int)(rn == 8) - (int)(qn == 8);
                }
                if (rn != 15 && qn != 15 && rn != qn) {   // bam.py:389-390, 405
                    const uint32_t q = qual[qpos];
                    if (EMIT) {
                        out_rec[emitted] = (unsigned long long)(uint32_t)rpos |
                                           ((unsigned long long)((rn << 4) | qn) << 32) |
                                           ((unsigned long long)q << 40);
                        out_read[emitted] = read_idx;
                        ++emitted;
                    } else {
                        ++cn.n_conv;
                        if ((int)q > p.quality) {         // conversion.py:424 (strict)
                            const bool r1 = __popc(rn) == 1, q1 = __popc(qn) == 1;
                            if (r1 && q1) {
                                const int r = __ffs(rn) - 1, d = __ffs(qn) - 1;
                                const int idx = r * 3 + (d > r ? d - 1 : d);
                                bool snp = false;
                                if (p.n_snps > 0)
                                    snp = snp_lookup(p.snps, p.n_snps,
                                                     ((unsigned long long)idx << 56) |
                                                         ((unsigned long long)ref_id << 32) | (uint32_t)rpos);
                                if (!snp) {
                                    const uint32_t sh = (idx & 3) << 3, wi = idx >> 2;
                                    uint32_t cur = wi == 0 ? cn.w0 : (wi == 1 ? cn.w1 : cn.w2);
                                    if (((cur >> sh) & 0xff) == 0xff) cn.status |= DYN_ST_COUNTER_OVERFLOW;
                                    else {
                                        const uint32_t inc = 1u << sh;
                                        cn.w0 += wi == 0 ? inc : 0u;
                                        cn.w1 += wi == 1 ? inc : 0u;
                                        cn.w2 += wi == 2 ? inc : 0u;
                                    }
                                }
                            } else {
                                cn.status |= DYN_ST_NON_ACGT_CONV;
                            }
                        }
                    }
                }
                ++mdp; ++qpos; ++rpos; --len;
            }
        } else if (op == 2) {  // D: "^" + deleted reference bases; they count toward content (bam.py:359-360)
            if (run == 0 && mdp < md_len && is_digit(md[mdp])) {
                uint32_t ch;
                int v = 0;
                while (mdp < md_len && is_digit(ch = md[mdp])) { v = v * 10 + (int)(ch - '0'); ++mdp; }
                run = v;
            }
            if (run != 0 || mdp >= md_len || md[mdp] != '^' || mdp + 1 + len > md_len) { bad = true; break; }
            ++mdp;
            if (!EMIT) {
                for (int j = 0; j < len; ++j) {
                    const uint32_t rn = kLetterCode[md[mdp + j] & 31];
                    da += (rn == 1); dc += (rn == 2); dg += (rn == 4); dt += (rn == 8);
                }
            }
            mdp += len;
            rpos += len;
        } else if (op == 3) {  // N
            rpos += len;
        } else if (op == 1 || op == 4) {  // I, S: query bases that are not aligned -> take them back out
            if (qpos + len > l_seq) { bad = true; break; }
            if (!EMIT) {
                for (int j = qpos; j < qpos + len; ++j) {
                    const uint32_t qn = (seqb[j >> 1] >> ((~j & 1) << 2)) & 0xf;
                    da -= (qn == 1); dc -= (qn == 2); dg -= (qn == 4); dt -= (qn == 8);
                }
            }
            qpos += len;
        }  // H, P: nothing
    }
    if (!EMIT) {
        if (bad) cn.status |= DYN_ST_BAD_RECORD;
        cn.a = (uint32_t)max((int)cn.a + da, 0); cn.c = (uint32_t)max((int)cn.c + dc, 0);
        cn.g = (uint32_t)max((int)cn.g + dg, 0); cn.t = (uint32_t)max((int)cn.t + dt, 0);
    }
}

__global__ void __launch_bounds__(kThreads, 4) tally_kernel(const TallyParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *stage = smem;                                            // [stage_bytes]
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + p.stage_bytes);   // 8 B
    uint32_t *warp_sums = reinterpret_cast<uint32_t *>(smem + p.stage_bytes + 16);  // [kThreads/32 + 2]

    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t parity = 0;
    uint32_t status_acc = 0;
    const bool emit = (p.flags & DYN_TALLY_EMIT_CONV) != 0;

    for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const long long r0 = tile * p.tile_reads;
        const long long r1 = min(p.n_reads, r0 + (long long)p.tile_reads);
        const uint32_t off0 = __ldg(p.rec_off + r0), off1 = __ldg(p.rec_off + r1);
        const unsigned long long bytes = (unsigned long long)(off1 - off0) << 4;
        const bool staged = bytes <= p.stage_bytes;
        if (staged && tid == 0 && bytes > 0) {
            fence_proxy_async();
            mbar_expect_tx(bar, (uint32_t)bytes);
            tma_bulk_g2s(stage, p.records + ((unsigned long long)off0 << 4), (uint32_t)bytes, bar);
        }
        const long long r = r0 + tid;
        const bool active = tid < p.tile_reads && r < r1;
        uint32_t my_off = 0;
        if (active) my_off = __ldg(p.rec_off + r);
        if (staged && bytes > 0) {
            mbar_wait(bar, parity);
            parity ^= 1;
        }

        Counters cn = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        const uint8_t *rec = nullptr;
        if (active) {
            rec = staged ? stage + ((unsigned long long)(my_off - off0) << 4)
                         : p.records + ((unsigned long long)my_off << 4);
            walk_record<false>(rec, p, cn, nullptr, nullptr, 0);
            if ((cn.a | cn.c | cn.g | cn.t) > 255u) {
                cn.status |= DYN_ST_COUNTER_OVERFLOW;
                cn.a = min(cn.a, 255u); cn.c = min(cn.c, 255u); cn.g = min(cn.g, 255u); cn.t = min(cn.t, 255u);
            }
            uint4 o;
            o.x = cn.w0; o.y = cn.w1; o.z = cn.w2;
            o.w = cn.a | (cn.c << 8) | (cn.g << 16) | (cn.t << 24);
            reinterpret_cast<uint4 *>(p.counts)[r] = o;
            status_acc |= cn.status;
        }

        if (emit) {
            // block exclusive scan of n_conv -> one global atomic per tile
            const int lane = tid & 31, wid = tid >> 5;
            uint32_t v = active ? cn.n_conv : 0u, incl = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            if (lane == 31) warp_sums[wid] = incl;
            __syncthreads();
            if (tid == 0) {
                uint32_t run = 0;
                for (int w = 0; w < kThreads / 32; ++w) { uint32_t t = warp_sums[w]; warp_sums[w] = run; run += t; }
                unsigned long long base = run ? atomicAdd(p.conv_total, (unsigned long long)run) : 0ull;
                reinterpret_cast<unsigned long long *>(warp_sums + kThreads / 32)[0] = base;
            }
            __syncthreads();
            const unsigned long long base = reinterpret_cast<unsigned long long *>(warp_sums + kThreads / 32)[0];
            const unsigned long long mine = base + warp_sums[wid] + (incl - v);
            if (active) {
                uint32_t n = cn.n_conv;
                if (n > 0xffffu) { n = 0xffffu; status_acc |= DYN_ST_COUNTER_OVERFLOW; }
                const bool fits = mine + cn.n_conv <= (unsigned long long)p.conv_cap;
                p.conv_off[r] = (uint32_t)mine;
                p.conv_n[r] = fits ? (uint16_t)n : (uint16_t)0;
                if (cn.n_conv > 0) {
                    if (fits) walk_record<true>(rec, p, cn, p.conv_rec + mine, p.conv_read + mine, (uint32_t)r);
                    else status_acc |= DYN_ST_CONV_CAPACITY;
                }
            }
        }
        __syncthreads();   // everyone is done with `stage` before the next bulk copy lands
    }
    if (status_acc) atomicOr(p.status, status_acc);
}

}  // namespace

extern "C" int dyn_tally_reads(dyn_ctx_t *ctx, const void *d_records, const uint32_t *d_rec_off, int64_t n_reads,
                               const uint64_t *d_snps, int64_t n_snps, int quality, uint32_t flags,
                               uint8_t *d_counts, uint32_t *d_conv_off, uint16_t *d_conv_n, uint64_t *d_conv_rec,
                               uint32_t *d_conv_read, int64_t conv_capacity, uint64_t *d_conv_total,
                               uint32_t *d_status, void *stream) {
    DYN_ARG(ctx != nullptr, "null context");
    DYN_ARG(n_reads >= 0, "negative n_reads");
    if (n_reads == 0) return DYN_OK;
    DYN_ARG(d_records && d_rec_off && d_counts && d_status, "null buffer");
    DYN_ARG((reinterpret_cast<uintptr_t>(d_records) & 15) == 0, "records must be 16-byte aligned");
    DYN_ARG((reinterpret_cast<uintptr_t>(d_counts) & 15) == 0, "counts must be 16-byte aligned");
    if (flags & DYN_TALLY_EMIT_CONV)
        DYN_ARG(d_conv_off && d_conv_n && d_conv_rec && d_conv_read && d_conv_total, "null conversion buffer");
    DYN_ARG(n_snps == 0 || d_snps, "null snp table");

    TallyParams p;
    p.records = static_cast<const uint8_t *>(d_records);
    p.rec_off = d_rec_off;
    p.n_reads = n_reads;
    p.snps = reinterpret_cast<const unsigned long long *>(d_snps);
    p.n_snps = n_snps;
    p.quality = quality;
    p.flags = flags;
    p.counts = d_counts;
    p.conv_off = d_conv_off;
    p.conv_n = d_conv_n;
    p.conv_rec = reinterpret_cast<unsigned long long *>(d_conv_rec);
    p.conv_read = d_conv_read;
    p.conv_cap = conv_capacity;
    p.conv_total = reinterpret_cast<unsigned long long *>(d_conv_total);
    p.status = d_status;
    // 4 CTAs per SM: (227 KB - 4 * 1 KB reserved) / 4, rounded down to 128 B
    p.stage_bytes = 55 * 1024;
    p.tile_reads = kThreads;
    p.n_tiles = (n_reads + p.tile_reads - 1) / p.tile_reads;
    const size_t smem = p.stage_bytes + 16 + sizeof(uint32_t) * (kThreads / 32 + 2) + 16;

    static bool attr_set = false;
    if (!attr_set) {
        DYN_CUDA(cudaFuncSetAttribute(tally_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    long long grid = (long long)ctx->sm_count * 4;
    if (grid > p.n_tiles) grid = p.n_tiles;
    tally_kernel<<<(unsigned)grid, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(p);
    DYN_CUDA(cudaGetLastError());
    return DYN_OK;
}
