// Per-read conversion tally (SURVEY.md section 8 rows a1 + a3).
//
// Replaces the inner loops of
//   dynast/preprocessing/bam.py:357-429        CIGAR/MD walk, reference base content, mismatch records
//   dynast/preprocessing/conversion.py:403-432 quality / SNP filter and the 12 + 4 counters
// which in the reference run as CPython per-base loops plus a CSV round trip.
//
// Design (HBM-bound integer/byte work, no tensor cores):
//   * one thread per read ("warp per read batch": a warp owns 32 consecutive reads), persistent CTAs,
//     a tile = up to 256 consecutive packed records;
//   * the tile's bytes are one contiguous range of the record blob, fetched by ONE TMA 1-D bulk copy
//     (cp.async.bulk -> UBLKCP) into shared memory and awaited on an mbarrier; 4 CTAs share an SM so
//     while one CTA waits for its tile the others walk theirs (multi-buffering across CTAs);
//   * the walk is organised so that the lanes of a warp stay converged (the first version, a per-read
//     CIGAR/MD state machine, ran with 5.7 of 32 lanes active):
//       1. base content of the WHOLE query in a fixed-trip loop, 8 bases per 32-bit word with nibble
//          one-hot bit planes (no popc, no masks);
//       2. a short CIGAR pre-scan (clipped / inserted bases are taken back out);
//       3. the MD string is scanned one byte per iteration with a warp-uniform trip count; it only
//          LOCATES events: (aligned offset, reference base) of each mismatch goes to a per-thread slot
//          list in shared memory, deleted bases are added to the content on the spot;
//       4. events are resolved (offset -> query/reference position through the CIGAR, read base,
//          Phred, quality/SNP filter) in a second warp-uniform loop;
//   * counters leave as one 128-bit store per read (consecutive threads -> consecutive 16 B: coalesced);
//   * mismatch records are placed with one global atomic per CTA tile (block scan of per-read counts)
//     and written by replaying the event slots -- the MD string is parsed once.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kEvSlots = 8;   // mismatch events kept per read in shared memory; more -> slow path

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
    uint32_t stage_bytes;
    long long n_tiles;
};

// BAM 4-bit code of an upper/lower-case IUPAC letter, indexed by (letter & 31), 4 bits per entry:
//   @ A B C D E F G H I J K L M N O | P Q R S T U V W X Y Z
//   0 1 14 2 13 0 0 4 11 0 0 12 0 3 15 0 | 0 0 5 6 8 0 7 9 0 10 0 ...
__device__ __forceinline__ uint32_t letter_code(uint32_t ch) {
    // entries 0..15 and 16..31 packed little-nibble-first
    const unsigned long long lo = (0ull) | (1ull << 4) | (14ull << 8) | (2ull << 12) | (13ull << 16) | (0ull << 20) |
                                  (0ull << 24) | (4ull << 28) | (11ull << 32) | (0ull << 36) | (0ull << 40) |
                                  (12ull << 44) | (0ull << 48) | (3ull << 52) | (15ull << 56) | (0ull << 60);
    const unsigned long long hi = (0ull) | (0ull << 4) | (5ull << 8) | (6ull << 12) | (8ull << 16) | (0ull << 20) |
                                  (7ull << 24) | (9ull << 28) | (0ull << 32) | (10ull << 36);
    const uint32_t i = ch & 31;
    const unsigned long long t = (i & 16) ? hi : lo;
    return (uint32_t)(t >> ((i & 15) << 2)) & 0xf;
}

struct Acc {
    int a, c, g, t;           // reference base content
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

__device__ __forceinline__ uint32_t fold_nibbles(uint32_t acc) {
    acc = (acc & 0x0f0f0f0fu) + ((acc >> 4) & 0x0f0f0f0fu);
    return (acc * 0x01010101u) >> 24;
}

// A/C/G/T over the whole query: fixed-trip, no masks (the packer zero-pads, code 0 is not a base).
__device__ __forceinline__ void count_query(const uint32_t *seqw, int l_seq, Acc &cn) {
    const int n_words = (l_seq + 7) >> 3;
    for (int w0 = 0; w0 < n_words; w0 += 15) {
        uint32_t aa = 0, ac = 0, ag = 0, at = 0;
        const int w1 = min(n_words, w0 + 15);
#pragma unroll 4
        for (int w = w0; w < w1; ++w) {
            const uint32_t x = seqw[w];
            const uint32_t b0 = x & 0x11111111u, b1 = (x >> 1) & 0x11111111u, b2 = (x >> 2) & 0x11111111u,
                           b3 = (x >> 3) & 0x11111111u;
            const uint32_t s = b0 + b1 + b2 + b3;
            const uint32_t onehot = s & ~(s >> 1) & ~(s >> 2);
            aa += b0 & onehot; ac += b1 & onehot; ag += b2 & onehot; at += b3 & onehot;
        }
        cn.a += (int)fold_nibbles(aa); cn.c += (int)fold_nibbles(ac);
        cn.g += (int)fold_nibbles(ag); cn.t += (int)fold_nibbles(at);
    }
}

// Subtract the bases of query range [a, b) (soft clips, insertions) from the content.
__device__ __forceinline__ void uncount_range(const uint32_t *seqw, int a, int b, Acc &cn) {
    for (int w = a >> 3; w <= (b - 1) >> 3; ++w) {
        uint32_t x = seqw[w];
        x = ((x & 0x0f0f0f0fu) << 4) | ((x >> 4) & 0x0f0f0f0fu);   // base i at bits 4*(i&7)
        const int lo = max(a - (w << 3), 0), hi = min(b - (w << 3), 8);
        const uint32_t mask = (hi == 8 ? 0xffffffffu : ((1u << (4 * hi)) - 1u)) & ~((1u << (4 * lo)) - 1u);
        const uint32_t b0 = x & 0x11111111u, b1 = (x >> 1) & 0x11111111u, b2 = (x >> 2) & 0x11111111u,
                       b3 = (x >> 3) & 0x11111111u;
        const uint32_t s = b0 + b1 + b2 + b3;
        const uint32_t onehot = s & ~(s >> 1) & ~(s >> 2) & mask;
        cn.a -= __popc(b0 & onehot); cn.c -= __popc(b1 & onehot);
        cn.g -= __popc(b2 & onehot); cn.t -= __popc(b3 & onehot);
    }
}

// Aligned (M/=/X) offset -> query / reference position through the CIGAR.
__device__ __forceinline__ bool map_offset(const uint32_t *cig, int n_cigar, int pos, int aoff, int &qpos, int &rpos) {
    int q = 0, r = pos, acc = 0;
    for (int ci = 0; ci < n_cigar; ++ci) {
        const uint32_t c = cig[ci];
        const int op = c & 0xf, len = (int)(c >> 4);
        if (op == 0 || op == 7 || op == 8) {
            if (aoff < acc + len) { qpos = q + (aoff - acc); rpos = r + (aoff - acc); return true; }
            acc += len; q += len; r += len;
        } else if (op == 1 || op == 4) q += len;
        else if (op == 2 || op == 3) r += len;
    }
    return false;
}

struct RecView {
    const uint32_t *cig;
    const uint32_t *seqw;
    const uint8_t *seqb, *qual, *md;
    int pos, l_seq, n_cigar, md_len;
    uint32_t ref_id;
};

// One mismatch event: reference code `rn` at aligned offset `aoff`.
template <bool EMIT>
__device__ __forceinline__ void do_event(const RecView &v, const TallyParams &p, int aoff, uint32_t rn, Acc &cn,
                                         unsigned long long *out_rec, uint32_t *out_read, uint32_t read_idx,
                                         uint32_t &emitted) {
    int qpos, rpos;
    if (!map_offset(v.cig, v.n_cigar, v.pos, aoff, qpos, rpos)) { if (!EMIT) cn.status |= DYN_ST_BAD_RECORD; return; }
    const uint32_t qn = (v.seqb[qpos >> 1] >> ((~qpos & 1) << 2)) & 0xf;
    if (!EMIT) {
        // count_query took the read base for the reference base: put the reference base in instead
        cn.a += (int)(rn == 1) - (int)(qn == 1); cn.c += (int)(rn == 2) - (int)(qn == 2);
        cn.g += (int)(rn == 4) - (int)(qn == 4); cn.t += (int)(rn == 8) - (int)(qn == 8);
    }
    if (rn == 15 || qn == 15 || rn == qn) return;   // bam.py:389-390, 405
    const uint32_t q = v.qual[qpos];
    if (EMIT) {
        out_rec[emitted] = (unsigned long long)(uint32_t)rpos | ((unsigned long long)((rn << 4) | qn) << 32) |
                           ((unsigned long long)q << 40);
        out_read[emitted] = read_idx;
        ++emitted;
        return;
    }
    ++cn.n_conv;
    if ((int)q <= p.quality) return;                // conversion.py:424 (strict)
    if (__popc(rn) != 1 || __popc(qn) != 1) { cn.status |= DYN_ST_NON_ACGT_CONV; return; }
    const int r = __ffs(rn) - 1, d = __ffs(qn) - 1;
    const int idx = r * 3 + (d > r ? d - 1 : d);
    if (p.n_snps > 0 &&
        snp_lookup(p.snps, p.n_snps,
                   ((unsigned long long)idx << 56) | ((unsigned long long)v.ref_id << 32) | (uint32_t)rpos))
        return;
    const uint32_t sh = (idx & 3) << 3, wi = idx >> 2;
    const uint32_t cur = wi == 0 ? cn.w0 : (wi == 1 ? cn.w1 : cn.w2);
    if (((cur >> sh) & 0xff) == 0xff) { cn.status |= DYN_ST_COUNTER_OVERFLOW; return; }
    const uint32_t inc = 1u << sh;
    cn.w0 += wi == 0 ? inc : 0u;
    cn.w1 += wi == 1 ? inc : 0u;
    cn.w2 += wi == 2 ? inc : 0u;
}

// Slow path for reads with more than kEvSlots mismatches: sequential MD scan, events handled inline
// from slot `first_ev` on (the first kEvSlots were handled from the slot list).
template <bool EMIT>
__device__ __noinline__ void overflow_events(const RecView &v, const TallyParams &p, Acc &cn,
                                             unsigned long long *out_rec, uint32_t *out_read, uint32_t read_idx,
                                             uint32_t &emitted) {
    int val = 0, aoff = 0, n_ev = 0;
    bool in_del = false;
    for (int j = 0; j < v.md_len; ++j) {
        const uint32_t c = v.md[j];
        if (is_digit(c)) { val = val * 10 + (int)(c - '0'); in_del = false; }
        else if (c == '^') { aoff += val; val = 0; in_del = true; }
        else if (!in_del) {
            aoff += val; val = 0;
            if (n_ev >= kEvSlots) do_event<EMIT>(v, p, aoff, letter_code(c), cn, out_rec, out_read, read_idx, emitted);
            ++n_ev; ++aoff;
        }
    }
}

__global__ void __launch_bounds__(kThreads, 4) tally_kernel(const TallyParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *stage = smem;                                                        // [stage_bytes]
    uint32_t *ev = reinterpret_cast<uint32_t *>(smem + p.stage_bytes);            // [kEvSlots][kThreads]
    uint64_t *bar = reinterpret_cast<uint64_t *>(ev + kEvSlots * kThreads);       // 8 B
    uint32_t *warp_sums = reinterpret_cast<uint32_t *>(bar + 2);                  // [kThreads/32 + 2]
    uint32_t *s_sub = warp_sums + kThreads / 32 + 2;                              // sub-tile end broadcast

    const int tid = threadIdx.x;
    const int lane = tid & 31, wid = tid >> 5;
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t parity = 0;
    uint32_t status_acc = 0;
    const bool emit = (p.flags & DYN_TALLY_EMIT_CONV) != 0;
    const uint32_t cap16 = p.stage_bytes >> 4;

    for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const long long t0 = tile * kThreads;
        const long long t1 = min(p.n_reads, t0 + (long long)kThreads);
        const uint32_t off_end = __ldg(p.rec_off + t1);
        long long r0 = t0;
        while (r0 < t1) {
            // ---- sub-tile [r0, r1): as many reads as fit the staging buffer (normally the whole tile)
            const uint32_t off0 = __ldg(p.rec_off + r0);
            long long r1 = t1;
            bool oversized = false;
            if (off_end - off0 > cap16) {                // uniform across the CTA
                if (tid == 0) {
                    long long lo = r0, hi = t1;          // largest e with rec_off[e] - off0 <= cap16
                    while (lo < hi) {
                        const long long mid = (lo + hi + 1) >> 1;
                        if (__ldg(p.rec_off + mid) - off0 <= cap16) lo = mid; else hi = mid - 1;
                    }
                    s_sub[0] = (uint32_t)(lo - r0);
                }
                __syncthreads();
                const uint32_t cnt = s_sub[0];
                __syncthreads();
                if (cnt == 0) { oversized = true; r1 = r0 + 1; } else r1 = r0 + cnt;
            }
            const uint32_t off1 = __ldg(p.rec_off + r1);
            const uint32_t bytes = oversized ? 0u : (off1 - off0) << 4;
            if (tid == 0 && bytes > 0) {
                fence_proxy_async();
                mbar_expect_tx(bar, bytes);
                tma_bulk_g2s(stage, p.records + ((unsigned long long)off0 << 4), bytes, bar);
            }
            const long long r = r0 + tid;
            const bool active = r < r1 && !oversized;
            uint32_t my_off = 0, my_end = 0;
            if (r < r1) { my_off = __ldg(p.rec_off + r); my_end = __ldg(p.rec_off + r + 1); }
            if (bytes > 0) {
                mbar_wait(bar, parity);
                parity ^= 1;
            }

            // ---- header
            Acc cn = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            RecView v;
            v.pos = 0; v.l_seq = 0; v.n_cigar = 0; v.md_len = 0; v.ref_id = 0;
            const uint8_t *rec = stage + ((size_t)(my_off - off0) << 4);
            bool good = active;
            if (active) {
                const uint4 h = *reinterpret_cast<const uint4 *>(rec);
                v.pos = (int)h.x; v.ref_id = h.y;
                v.l_seq = h.z & 0xffff; v.n_cigar = h.z >> 16; v.md_len = h.w & 0xffff;
                const size_t need = 16 + 4 * (size_t)v.n_cigar + (((((size_t)v.l_seq + 1) >> 1) + 3) & ~(size_t)3) +
                                    v.l_seq + v.md_len;
                if (need > ((size_t)(my_end - my_off) << 4)) {   // truncated / corrupt record
                    good = false; v.l_seq = v.n_cigar = v.md_len = 0;
                }
            }
            v.cig = reinterpret_cast<const uint32_t *>(rec + 16);
            v.seqw = v.cig + v.n_cigar;
            v.seqb = reinterpret_cast<const uint8_t *>(v.seqw);
            v.qual = v.seqb + ((((v.l_seq + 1) >> 1) + 3) & ~3);
            v.md = v.qual + v.l_seq;
            if (r < r1 && !good) cn.status |= DYN_ST_BAD_RECORD;

            // ---- 1. whole-query base content
            count_query(v.seqw, v.l_seq, cn);

            // ---- 2. CIGAR pre-scan
            int m_total = 0, d_total = 0;
            {
                int q = 0;
                for (int ci = 0; ci < v.n_cigar; ++ci) {
                    const uint32_t c = v.cig[ci];
                    const int op = c & 0xf, len = (int)(c >> 4);
                    if (op == 0 || op == 7 || op == 8) { m_total += len; q += len; }
                    else if (op == 1 || op == 4) {
                        if (q + len <= v.l_seq && len > 0) uncount_range(v.seqw, q, q + len, cn);
                        q += len;
                    } else if (op == 2) d_total += len;
                }
                if (q != v.l_seq) { cn.status |= DYN_ST_BAD_RECORD; good = false; }
            }

            // ---- 3. MD scan: locate events (warp-uniform trip count)
            int n_ev = 0;
            {
                const int md_len = good ? v.md_len : 0;
                const int max_len = __reduce_max_sync(0xffffffffu, md_len);
                int val = 0, aoff = 0, n_delb = 0;
                bool in_del = false;
                for (int j = 0; j < max_len; ++j) {
                    if (j < md_len) {
                        const uint32_t c = v.md[j];
                        const uint32_t d = c - '0';
                        if (d <= 9u) { val = val * 10 + (int)d; in_del = false; }
                        else if (c == '^') { aoff += val; val = 0; in_del = true; }
                        else {
                            const uint32_t code = letter_code(c);
                            if (in_del) {   // deleted reference base: counts toward content (bam.py:359-360)
                                cn.a += (code == 1); cn.c += (code == 2); cn.g += (code == 4); cn.t += (code == 8);
                                ++n_delb;
                            } else {
                                aoff += val; val = 0;
                                if (n_ev < kEvSlots) ev[n_ev * kThreads + tid] = (uint32_t)aoff | (code << 16);
                                ++n_ev; ++aoff;
                            }
                        }
                    }
                }
                if (good && (aoff + val != m_total || n_delb != d_total)) { cn.status |= DYN_ST_BAD_RECORD; n_ev = 0; }
            }

            // ---- 4. resolve events
            {
                const int n_slot = min(n_ev, kEvSlots);
                const int max_ev = __reduce_max_sync(0xffffffffu, n_slot);
                uint32_t dummy = 0;
                for (int e = 0; e < max_ev; ++e) {
                    if (e < n_slot) {
                        const uint32_t w = ev[e * kThreads + tid];
                        do_event<false>(v, p, (int)(w & 0xffff), w >> 16, cn, nullptr, nullptr, 0, dummy);
                    }
                }
                if (n_ev > kEvSlots) overflow_events<false>(v, p, cn, nullptr, nullptr, 0, dummy);
            }

            if (r < r1) {
                uint32_t a = (uint32_t)max(cn.a, 0), c = (uint32_t)max(cn.c, 0), g = (uint32_t)max(cn.g, 0),
                         t = (uint32_t)max(cn.t, 0);
                if ((a | c | g | t) > 255u) {
                    cn.status |= DYN_ST_COUNTER_OVERFLOW;
                    a = min(a, 255u); c = min(c, 255u); g = min(g, 255u); t = min(t, 255u);
                }
                if (!good) { a = c = g = t = 0; cn.w0 = cn.w1 = cn.w2 = 0; cn.n_conv = 0; }
                uint4 o;
                o.x = cn.w0; o.y = cn.w1; o.z = cn.w2;
                o.w = a | (c << 8) | (g << 16) | (t << 24);
                reinterpret_cast<uint4 *>(p.counts)[r] = o;
                status_acc |= cn.status;
            }

            if (emit) {
                // block exclusive scan of n_conv -> one global atomic per (sub-)tile
                const uint32_t mine_n = (r < r1) ? cn.n_conv : 0u;
                uint32_t incl = mine_n;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += t;
                }
                if (lane == 31) warp_sums[wid] = incl;
                __syncthreads();
                if (tid == 0) {
                    uint32_t run = 0;
                    for (int w = 0; w < kThreads / 32; ++w) { const uint32_t t = warp_sums[w]; warp_sums[w] = run; run += t; }
                    const unsigned long long base = run ? atomicAdd(p.conv_total, (unsigned long long)run) : 0ull;
                    reinterpret_cast<unsigned long long *>(warp_sums + kThreads / 32)[0] = base;
                }
                __syncthreads();
                const unsigned long long base = reinterpret_cast<unsigned long long *>(warp_sums + kThreads / 32)[0];
                const unsigned long long mine = base + warp_sums[wid] + (incl - mine_n);
                if (r < r1) {
                    uint32_t n = mine_n;
                    if (n > 0xffffu) { n = 0xffffu; status_acc |= DYN_ST_COUNTER_OVERFLOW; }
                    const bool fits = mine + mine_n <= (unsigned long long)p.conv_cap;
                    p.conv_off[r] = (uint32_t)mine;
                    p.conv_n[r] = fits ? (uint16_t)n : (uint16_t)0;
                    if (mine_n > 0 && !fits) status_acc |= DYN_ST_CONV_CAPACITY;
                    if (mine_n > 0 && fits) {
                        uint32_t emitted = 0;
                        const int n_slot = min(n_ev, kEvSlots);
                        for (int e = 0; e < n_slot; ++e) {
                            const uint32_t w = ev[e * kThreads + tid];
                            do_event<true>(v, p, (int)(w & 0xffff), w >> 16, cn, p.conv_rec + mine, p.conv_read + mine,
                                           (uint32_t)r, emitted);
                        }
                        if (n_ev > kEvSlots)
                            overflow_events<true>(v, p, cn, p.conv_rec + mine, p.conv_read + mine, (uint32_t)r, emitted);
                    }
                }
            }
            __syncthreads();   // everyone is done with `stage` / `ev` before the next bulk copy lands
            r0 = r1;
        }
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
    DYN_ARG(!(flags & DYN_TALLY_NASC), "paired records / --nasc are handled by dyn_tally_pairs");

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
    // 4 CTAs per SM share 227 KB: 46 KB staging + 8 KB event slots + bookkeeping each
    p.stage_bytes = 46 * 1024;
    p.n_tiles = (n_reads + kThreads - 1) / kThreads;
    const size_t smem = p.stage_bytes + sizeof(uint32_t) * kEvSlots * kThreads + 16 +
                        sizeof(uint32_t) * (kThreads / 32 + 2 + 4) + 16;

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
