// Per-read conversion tally (SURVEY.md section 8 rows a1 + a3).
//
// Replaces the inner loops of
//   dynast/preprocessing/bam.py:297-429        CIGAR/MD walk, reference base content, paired-overlap
//                                              rules, mismatch records
//   dynast/preprocessing/conversion.py:403-432 quality / SNP filter and the 12 + 4 counters
// which in the reference run as CPython per-base loops plus a CSV round trip.
//
// Design (HBM-bound integer/byte work, no tensor cores):
//   * one thread per counting unit ("warp per read batch": a warp owns 32 consecutive reads); a tile is
//     kThreads consecutive packed records = ONE contiguous byte range of the record blob, fetched by one
//     TMA 1-D bulk copy (cp.async.bulk -> UBLKCP) into shared memory and awaited on an mbarrier.
//     CTAs are small (64 threads, ~20 KB) so eight to ten share an SM: while one waits for its tile
//     the others walk theirs (multi-buffering across CTAs, no intra-CTA pipeline state);
//   * the walk keeps the lanes of a warp converged (v1, a per-read CIGAR/MD state machine, ran with
//     5.7 of 32 lanes active; v3, per-lane event loops, spent 45 % of its instructions at 3 lanes):
//       1. CIGAR pre-scan: aligned / deleted totals, the soft clips at either end, number of interior
//          query ranges (insertions);
//       2. base content of the query between the end clips, 4 bases per byte-permute instruction
//          (the 4-bit base codes are the PRMT selectors, see count_query);
//       3. the number of mismatch events is known before the MD tokens are read (tokens - deleted
//          bases), so ONE packed warp scan places every read's events in the warp's shared list --
//          ranges first, then mismatches; the token loop (the packer tokenises the MD text; v4 spent
//          10 of its 57 warp-instructions per read scanning it byte by byte) has a warp-uniform trip
//          count and only LOCATES work: (aligned offset, reference base) goes straight into the list,
//          deleted reference bases are added to the content on the spot;
//       4. the list is resolved LANE-BALANCED -- lane i takes event i, whoever's read it belongs to:
//          offset -> query/reference position through the CIGAR, read base, Phred, quality / SNP
//          filter; results flow back to the owner through shared-memory atomics on that read's private
//          accumulators, the finished 64-bit mismatch record replaces the event in the list; a ballot
//          per chunk of 32 events records which entries are records, and per-read record counts and
//          output positions are popcounts over those ballots;
//   * counters leave as one 128-bit store per read (consecutive threads -> consecutive 16 B: coalesced);
//   * mismatch records are placed with one global atomic per warp and copied out of the list
//     lane-balanced (consecutive lanes -> consecutive records);
//   * paired units (bam.py:307-352, 393-419), warps whose events overflow the list and records larger
//     than the staging buffer are parked on a list and taken by a second kernel with a sequential
//     per-thread walker (tally_slow_kernel / slow_unit) -- same results, no attempt at convergence.  Keeping
//     the walker out of the tile kernel takes it from 128 to ~70 registers (11-12 CTAs per SM instead of 8);
//     the single-end fast path is what the bench times.
#include "common.cuh"
#include "walker.cuh"

#include <algorithm>

namespace {

#ifndef DYN_TALLY_THREADS
#define DYN_TALLY_THREADS 64
#endif
#ifndef DYN_TALLY_SLOTS
#define DYN_TALLY_SLOTS 8
#endif
#ifndef DYN_TALLY_MIN_STAGE
#define DYN_TALLY_MIN_STAGE 8192   // bytes: smaller tiles than this are not worth a bulk copy of their own
#endif
#ifndef DYN_TALLY_MIN_CTAS
#define DYN_TALLY_MIN_CTAS 12
#endif
#ifndef DYN_TALLY_SLACK
#define DYN_TALLY_SLACK 256          // staging bytes beyond kThreads mean-sized records
#endif
constexpr int kThreads = DYN_TALLY_THREADS;
constexpr int kWarps = kThreads / 32;
constexpr int kEvSlots = DYN_TALLY_SLOTS;   // event list entries per read (a warp of 32 reads shares 32 x this many)
constexpr unsigned long long kEvRange = 1ull << 63;

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
    uint32_t *slow_list;     // read indices left to tally_slow_kernel (paired units, event overflow, oversized records)
    uint32_t *n_slow;
};

__device__ __forceinline__ bool snp_lookup(const unsigned long long *snps, long long n, unsigned long long key) {
    long long lo = 0, hi = n;
    while (lo < hi) {
        long long mid = (lo + hi) >> 1;
        unsigned long long v = __ldg(snps + mid);
        if (v < key) lo = mid + 1; else hi = mid;
    }
    return lo < n && __ldg(snps + lo) == key;
}

// column of conversion ref -> read in conversion.py:44-57 (alphabetical AC AG AT CA ...)
__device__ __forceinline__ int conv_idx(int r, int d) { return r * 3 + (d > r ? d - 1 : d); }

// Base counting by byte permute: PRMT's selector nibbles ARE the 4-bit base codes.  Selector value v picks table
// byte v & 7 and, when bit 3 is set, replicates that byte's sign bit instead (PTX prmt.b32, generic mode), so with
//   table AC: byte 1 = 0x01 (A = 1), byte 2 = 0x10 (C = 2), rest 0      -> every other code, 9..15 included, gives 0
//   table GT: byte 4 = 0x01 (G = 4), byte 0 = 0x80 (T = 8 -> 0xff, '=' = 0 -> 0x80; & 0x11 leaves T = 0x11, '=' = 0)
// four bases are classified per instruction, exactly, for all 16 codes; per byte lane the low nibble counts A
// (resp. G + T) and the high nibble C (resp. T).  Nibble fields hold 15: fold after at most 7 words (14 adds).
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
constexpr uint32_t kLutG = 0x00000001u;
// The two tables that go in PRMT's register operand are read from shared memory once per thread: given the
// constants, ptxas keeps them in uniform registers and copies them to a fresh register in front of EVERY PRMT
// (one extra instruction per four bases).
struct BaseLut {
    uint32_t ac, t;
};

__device__ __forceinline__ void fold_fields(uint32_t acc_ac, uint32_t acc_gt, int &ca, int &cc, int &cg, int &ct) {
    const uint32_t a = ((acc_ac & 0x0f0f0f0fu) * 0x01010101u) >> 24, c = (((acc_ac >> 4) & 0x0f0f0f0fu) * 0x01010101u) >> 24;
    const uint32_t gt = ((acc_gt & 0x0f0f0f0fu) * 0x01010101u) >> 24, t = (((acc_gt >> 4) & 0x0f0f0f0fu) * 0x01010101u) >> 24;
    ca += (int)a; cc += (int)c; cg += (int)(gt - t); ct += (int)t;
}

// nibble mask of bases [lo, hi) of a sequence word (0 <= lo < hi <= 8); base i sits at bits 4 * (i ^ 1)
__device__ __forceinline__ uint32_t word_mask(int lo, int hi) {
    const uint32_t m = (0xffffffffu >> (32 - 4 * hi)) & ~((1u << (4 * lo)) - 1u);
    return ((m & 0x0f0f0f0fu) << 4) | ((m >> 4) & 0x0f0f0f0fu);
}

// A/C/G/T over query bases [qa, qb) -- the whole query minus leading / trailing soft clips.  Words are taken in
// fully unrolled rounds of 6 with predicated loads (a word past the end reads as 0 = no base); the first word
// is masked at a fixed slot, the last word is peeled.
__device__ __forceinline__ void count_query(const BaseLut &lut, const uint32_t *seqw, int qa, int qb, int &ca, int &cc, int &cg,
                                            int &ct) {
    if (qb <= qa) return;
    const int wa = qa >> 3, we = (qb - 1) >> 3;
    const uint32_t mask_lo = word_mask(qa & 7, 8), mask_hi = word_mask(0, ((qb - 1) & 7) + 1);
    uint32_t acc_ac, acc_gt;
    {
        const uint32_t x = seqw[we] & mask_hi & (wa == we ? mask_lo : 0xffffffffu), xh = x >> 16;
        acc_ac = prmt(lut.ac, 0u, x) + prmt(lut.ac, 0u, xh);
        acc_gt = (prmt(lut.t, kLutG, x) & 0x11111111u) + (prmt(lut.t, kLutG, xh) & 0x11111111u);
    }
    for (int w0 = wa; w0 < we; w0 += 6) {
        if (w0 != wa) { fold_fields(acc_ac, acc_gt, ca, cc, cg, ct); acc_ac = acc_gt = 0; }
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            uint32_t x = (w0 + k < we) ? seqw[w0 + k] : 0u;
            if (k == 0 && w0 == wa) x &= mask_lo;
            const uint32_t xh = x >> 16;
            acc_ac += prmt(lut.ac, 0u, x) + prmt(lut.ac, 0u, xh);
            acc_gt += (prmt(lut.t, kLutG, x) & 0x11111111u) + (prmt(lut.t, kLutG, xh) & 0x11111111u);
        }
    }
    fold_fields(acc_ac, acc_gt, ca, cc, cg, ct);
}

// A/C/G/T of query range [a, b) (soft clips, insertions).
__device__ __forceinline__ void count_range(const BaseLut &lut, const uint32_t *seqw, int a, int b, int &ca, int &cc, int &cg,
                                            int &ct) {
    for (int w0 = a >> 3; w0 <= (b - 1) >> 3; w0 += 7) {
        uint32_t acc_ac = 0, acc_gt = 0;
        const int w1 = min(((b - 1) >> 3) + 1, w0 + 7);
        for (int w = w0; w < w1; ++w) {
            // base i of the word sits at bits 4 * (i ^ 1): build the mask for bases [lo, hi) and swap its nibbles
            const int lo = max(a - (w << 3), 0), hi = min(b - (w << 3), 8);
            uint32_t mask = (hi == 8 ? 0xffffffffu : ((1u << (4 * hi)) - 1u)) & ~((1u << (4 * lo)) - 1u);
            mask = ((mask & 0x0f0f0f0fu) << 4) | ((mask >> 4) & 0x0f0f0f0fu);
            const uint32_t x = seqw[w] & mask, xh = x >> 16;
            acc_ac += prmt(lut.ac, 0u, x) + prmt(lut.ac, 0u, xh);
            acc_gt += (prmt(lut.t, kLutG, x) & 0x11111111u) + (prmt(lut.t, kLutG, xh) & 0x11111111u);
        }
        fold_fields(acc_ac, acc_gt, ca, cc, cg, ct);
    }
}

// is reference position p an aligned (M/=/X) position of the record? (pysam get_reference_positions)
__device__ bool has_aligned_pos(const uint32_t *cig, int n_cigar, int pos, int p) {
    int r = pos;
    for (int ci = 0; ci < n_cigar; ++ci) {
        const uint32_t c = cig[ci];
        const int op = (int)(c & 0xf), len = (int)(c >> 4);
        if (op == 0 || op == 7 || op == 8) { if (p >= r && p < r + len) return true; r += len; }
        else if (op == 2 || op == 3) r += len;
    }
    return false;
}

struct SlowOut {
    int content[4];
    uint32_t w0, w1, w2, n_conv, status;
};

__device__ __forceinline__ void bump_counter(uint32_t &w0, uint32_t &w1, uint32_t &w2, int idx, uint32_t &status) {
    const uint32_t sh = (idx & 3) << 3, wi = idx >> 2;
    const uint32_t cur = wi == 0 ? w0 : (wi == 1 ? w1 : w2);
    if (((cur >> sh) & 0xff) == 0xff) { status |= DYN_ST_COUNTER_OVERFLOW; return; }
    const uint32_t inc = 1u << sh;
    w0 += wi == 0 ? inc : 0u; w1 += wi == 1 ? inc : 0u; w2 += wi == 2 ? inc : 0u;
}

template <bool EMIT>
__device__ __forceinline__ void slow_call(const TallyParams &p, SlowOut &o, int ref_id, int rpos, uint32_t gn, uint32_t rn,
                                          uint32_t q, unsigned long long *out_rec, uint32_t *out_read, uint32_t read_idx) {
    if (EMIT) {
        out_rec[o.n_conv] = (unsigned long long)(uint32_t)rpos | ((unsigned long long)((gn << 4) | rn) << 32) |
                            ((unsigned long long)q << 40);
        out_read[o.n_conv] = read_idx;
        ++o.n_conv;
        return;
    }
    ++o.n_conv;
    if ((int)q <= p.quality) return;                // conversion.py:424 (strict)
    const int r = base_idx(gn), d = base_idx(rn);
    if (r < 0 || d < 0) { o.status |= DYN_ST_NON_ACGT_CONV; return; }
    const int idx = conv_idx(r, d);
    if (p.n_snps > 0 &&
        snp_lookup(p.snps, p.n_snps, ((unsigned long long)idx << 56) | ((unsigned long long)(uint32_t)ref_id << 32) | (uint32_t)rpos))
        return;
    bump_counter(o.w0, o.w1, o.w2, idx, o.status);
}

// One counting unit, sequentially. EMIT=false: content + counters + n_conv; EMIT=true: write the records.
template <bool EMIT>
__device__ __noinline__ void slow_unit(const uint8_t *rec, size_t unit_bytes, const TallyParams &p, uint32_t read_idx,
                                       SlowOut &o, unsigned long long *out_rec, uint32_t *out_read) {
    o.content[0] = o.content[1] = o.content[2] = o.content[3] = 0;
    o.w0 = o.w1 = o.w2 = 0; o.n_conv = 0; o.status = 0;
    const bool nasc = (p.flags & DYN_TALLY_NASC) != 0;
    Walker rd;
    const size_t rd_bytes = rd.init(rec, unit_bytes);
    if (rd.bad) { o.status |= DYN_ST_BAD_RECORD; return; }
    const bool has_mate = (reinterpret_cast<const uint16_t *>(rec)[7] & DYN_REC_HAS_MATE) != 0;
    if (has_mate && rd.lean) { o.status |= DYN_ST_BAD_RECORD; return; }      // the mate rule needs every quality
    const uint8_t *mrec = rec + rd_bytes;
    const size_t mate_avail = unit_bytes - rd_bytes;
    const int rd_pos = rd.rpos;
    bool bad = false;

    if (has_mate) {
        // content(mate) - overlap, floored at zero (Counter subtraction, bam.py:332-346)
        Walker m;
        m.init(mrec, mate_avail);
        if (m.bad) { o.status |= DYN_ST_BAD_RECORD; return; }
        int overlap[4] = {0, 0, 0, 0};
        while (m.next()) {
            const int b = base_idx(m.gn);
            if (!EMIT && b >= 0) ++o.content[b];
            if (!EMIT && m.gn != 15 && m.rn != 15 && b >= 0 && has_aligned_pos(rd.cig, rd.n_cigar, rd_pos, m.cr)) ++overlap[b];
        }
        bad |= m.bad;
        if (!EMIT)
            for (int b = 0; b < 4; ++b) o.content[b] = max(0, o.content[b] + m.del[b] - overlap[b]);
    }

    // the read's own walk; a second mate walker answers "does the mate hold a call at this position?"
    {
        Walker mm;
        bool mm_live = false;
        if (has_mate) { mm.init(mrec, mate_avail); mm_live = mm.next(); }
        while (rd.next()) {
            if (!EMIT) { const int b = base_idx(rd.gn); if (b >= 0) ++o.content[b]; }
            if (rd.gn == 15 || rd.rn == 15) continue;          // bam.py:389-390
            uint32_t rn = rd.rn, q = rd.q;
            if (has_mate && !nasc) {
                while (mm_live && mm.cr < rd.cr) mm_live = mm.next();
                if (mm_live && mm.cr == rd.cr && mm.gn != 15 && mm.rn != 15 && mm.gn != mm.rn && mm.q > q) {
                    rn = mm.rn; q = mm.q;                        // bam.py:395-403: mate wins only if strictly better
                }
            }
            if (rd.gn != rn) slow_call<EMIT>(p, o, rd.ref_id, rd.cr, rd.gn, rn, q, out_rec, out_read, read_idx);
        }
        bad |= rd.bad | (has_mate && mm.bad);
        if (!EMIT) for (int b = 0; b < 4; ++b) o.content[b] += rd.del[b];
    }

    // leftover mate calls, in the mate's walk order (bam.py:414-419)
    if (has_mate) {
        Walker m, r2;
        m.init(mrec, mate_avail);
        r2.init(rec, unit_bytes);
        bool r2_live = r2.next();
        while (m.next()) {
            if (m.gn == 15 || m.rn == 15 || m.gn == m.rn) continue;
            if (nasc && has_aligned_pos(rd.cig, rd.n_cigar, rd_pos, m.cr)) continue;   // bam.py:339
            while (r2_live && r2.cr < m.cr) r2_live = r2.next();
            if (r2_live && r2.cr == m.cr && r2.gn != 15 && r2.rn != 15) continue;      // consumed by the read's walk
            slow_call<EMIT>(p, o, rd.ref_id, m.cr, m.gn, m.rn, m.q, out_rec, out_read, read_idx);
        }
        bad |= m.bad;
    }
    if (bad) {
        o.status |= DYN_ST_BAD_RECORD;
        if (!EMIT) { o.content[0] = o.content[1] = o.content[2] = o.content[3] = 0; o.w0 = o.w1 = o.w2 = 0; o.n_conv = 0; }
    }
}

// ---------------------------------------------------------------------------------------------------
// Lane-balanced event resolution (fast path, shared-memory records)
// ---------------------------------------------------------------------------------------------------
// A warp's events live in ONE list of 64-bit entries (kWarpEvents per warp): first every query range (insertions,
// interior clips), then every mismatch, each group in read order.  An entry starts as the event
//   mismatch: aligned offset [0,16) | reference base code [16,20) | owner lane [24,29)
//   range:    kEvRange | query start [0,16) | owner lane [24,29) | length [32,48)
//   kEvNull:  nothing to do (the read turned out to be malformed)
// and is overwritten by its result: the finished mismatch record, or 0.
constexpr int kWarpEvents = 32 * kEvSlots;
constexpr unsigned long long kEvNull = 1ull << 62;

struct TileSmem {
    uint8_t *stage;
    unsigned long long *list;   // [kWarps][kWarpEvents]
    int *acc;                   // [7][kThreads]: a c g t w0 w1 w2
    uint32_t *recoff;           // [kThreads] byte offset of the record inside `stage`
    uint32_t *wmask;            // [kWarps][kEvSlots + 1] ballot of "event e left a mismatch record", per chunk of 32 events
    uint32_t *wpre;             // [kWarps][kEvSlots + 1] records in the chunks before
};

// one read's 16 output bytes (conversion.py:46-57 column order), saturating at 255
__device__ __forceinline__ void store_counts(const TallyParams &p, long long r, int ca, int cc, int cg, int ct, uint32_t w0,
                                             uint32_t w1, uint32_t w2, uint32_t &status) {
    uint32_t a = (uint32_t)max(ca, 0), c = (uint32_t)max(cc, 0), g = (uint32_t)max(cg, 0), t = (uint32_t)max(ct, 0);
    if ((a | c | g | t) > 255u) {
        status |= DYN_ST_COUNTER_OVERFLOW;
        a = min(a, 255u); c = min(c, 255u); g = min(g, 255u); t = min(t, 255u);
    }
    uint4 o;
    o.x = w0; o.y = w1; o.z = w2;
    o.w = a | (c << 8) | (g << 16) | (t << 24);
    reinterpret_cast<uint4 *>(p.counts)[r] = o;
}

// Resolves one event of the read owned by thread o_tid; returns the mismatch record it leaves (0: none).
__device__ __forceinline__ unsigned long long resolve_event(const TileSmem &sm, const TallyParams &p, const BaseLut &lut,
                                                            int o_tid, unsigned long long pl, uint32_t &status) {
    if (pl & kEvNull) return 0ull;
    const uint8_t *rec = sm.stage + sm.recoff[o_tid];
    const uint4 h = *reinterpret_cast<const uint4 *>(rec);
    const int pos = (int)h.x;
    const uint32_t l_seq = h.z & 0xffff, n_cigar = h.z >> 16;
    const uint32_t *cig = reinterpret_cast<const uint32_t *>(rec + 16);
    const uint32_t *seqw = cig + n_cigar + (h.w & 0xffff);      // MD tokens sit between the CIGAR and the sequence
    const uint8_t *seqb = reinterpret_cast<const uint8_t *>(seqw);
    int *acc = sm.acc + o_tid;
    if (pl & kEvRange) {
        // inserted / interior clipped query range: count_query took those bases for reference bases
        const int qs = (int)(pl & 0xffff), len = (int)((pl >> 32) & 0xffff);
        int da = 0, dc = 0, dg = 0, dt = 0;
        count_range(lut, seqw, qs, qs + len, da, dc, dg, dt);
        if (da) atomicSub(acc + 0 * kThreads, da);
        if (dc) atomicSub(acc + 1 * kThreads, dc);
        if (dg) atomicSub(acc + 2 * kThreads, dg);
        if (dt) atomicSub(acc + 3 * kThreads, dt);
        return 0ull;
    }
    const int aoff = (int)(pl & 0xffff);
    const uint32_t gn = (uint32_t)(pl >> 16) & 0xf;
    // aligned (M/=/X) offset -> query / reference position through the CIGAR
    int qpos = -1, rpos = 0;
    {
        int q = 0, r = pos, done = 0;
        for (uint32_t ci = 0; ci < n_cigar; ++ci) {
            const uint32_t c = cig[ci];
            const int op = (int)(c & 0xf), len = (int)(c >> 4);
            if (op == 0 || op == 7 || op == 8) {
                if (aoff < done + len) { qpos = q + (aoff - done); rpos = r + (aoff - done); break; }
                done += len; q += len; r += len;
            } else if (op == 1 || op == 4) q += len;
            else if (op == 2 || op == 3) r += len;
        }
    }
    if (qpos < 0 || qpos >= (int)l_seq) { status |= DYN_ST_BAD_RECORD; return 0ull; }
    const uint32_t rn = (seqb[qpos >> 1] >> ((~qpos & 1) << 2)) & 0xf;
    // count_query took the read base for the reference base: put the reference base in instead
    const int gi = base_idx(gn), ri = base_idx(rn);
    if (gi != ri) {
        if (gi >= 0) atomicAdd(acc + gi * kThreads, 1);
        if (ri >= 0) atomicSub(acc + ri * kThreads, 1);
    }
    if (gn == 15 || rn == 15 || gn == rn) return 0ull;   // bam.py:389-390, 405
    // lean records carry the Phred of a mismatching base in its MD token (copied into bits 32..39 of the event)
    const uint32_t q = ((h.w >> 16) & DYN_REC_LEAN) ? (uint32_t)(pl >> 32) & 0xffu : seqb[pad4(((size_t)l_seq + 1) >> 1) + qpos];
    const unsigned long long out =
        (unsigned long long)(uint32_t)rpos | ((unsigned long long)((gn << 4) | rn) << 32) | ((unsigned long long)q << 40);
    if ((int)q <= p.quality) return out;                                // conversion.py:424 (strict)
    if (gi < 0 || ri < 0) { status |= DYN_ST_NON_ACGT_CONV; return out; }
    const int idx = conv_idx(gi, ri);
    if (p.n_snps > 0 &&
        snp_lookup(p.snps, p.n_snps, ((unsigned long long)idx << 56) | ((unsigned long long)h.y << 32) | (uint32_t)rpos))
        return out;
    const uint32_t sh = (idx & 3) << 3;
    uint32_t *w = reinterpret_cast<uint32_t *>(acc + (4 + (idx >> 2)) * kThreads);
    const uint32_t old = atomicAdd(w, 1u << sh);
    if (((old >> sh) & 0xff) == 0xff) { atomicSub(w, 1u << sh); status |= DYN_ST_COUNTER_OVERFLOW; }
    return out;
}

__global__ void __launch_bounds__(kThreads, DYN_TALLY_MIN_CTAS) tally_kernel(const TallyParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    TileSmem sm;
    sm.stage = smem;                                                                   // [stage_bytes]
    sm.list = reinterpret_cast<unsigned long long *>(smem + p.stage_bytes);            // 8 * kWarps * kWarpEvents
    sm.acc = reinterpret_cast<int *>(sm.list + kWarps * kWarpEvents);                  // 4 * 7 * kThreads
    sm.recoff = reinterpret_cast<uint32_t *>(sm.acc + 7 * kThreads);                   // 4 * kThreads
    uint64_t *bar = reinterpret_cast<uint64_t *>(sm.recoff + kThreads);                // 16 B
    uint32_t *s_sub = reinterpret_cast<uint32_t *>(bar + 2);                           // sub-tile broadcast
    sm.wmask = s_sub + 4;                                                              // [kWarps][kEvSlots + 1]
    sm.wpre = sm.wmask + kWarps * (kEvSlots + 1);                                      // [kWarps][kEvSlots + 1]

    const int tid = threadIdx.x;
    const int lane = tid & 31, wid = tid >> 5;
    unsigned long long *wlist = sm.list + wid * kWarpEvents;
    uint32_t *wmask = sm.wmask + wid * (kEvSlots + 1), *wpre = sm.wpre + wid * (kEvSlots + 1);
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_sub[2] = 0x00100100u;     // BaseLut: A -> 0x01, C -> 0x10
        s_sub[3] = 0x00000080u;     //          T -> 0xff (sign of 0x80 replicated), '=' -> 0x80
    }
    __syncthreads();
    BaseLut lut;
    lut.ac = reinterpret_cast<volatile uint32_t *>(s_sub)[2];
    lut.t = reinterpret_cast<volatile uint32_t *>(s_sub)[3];
    uint32_t parity = 0;
    uint32_t status_acc = 0;
    const bool emit = (p.flags & DYN_TALLY_EMIT_CONV) != 0;
    const uint32_t cap16 = p.stage_bytes >> 4;

    // Offsets of the NEXT tile are fetched while the current one is being walked: the dependent chain
    // rec_off -> TMA issue -> mbarrier wait was 14 % of the stall samples (profiles/r1_tally_v4_ncu_summary.txt).
    uint32_t pf_off0 = 0, pf_off_end = 0, pf_my_off = 0, pf_my_end = 0;
    auto prefetch_offsets = [&](long long tile) {
        if (tile >= p.n_tiles) return;
        const long long n0 = tile * kThreads, n1 = min(p.n_reads, n0 + (long long)kThreads);
        // volatile asm: keep the loads HERE (the compiler otherwise sinks them to their first use, one tile later)
        const long long r = min(n0 + tid, n1 - 1);
        asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(pf_off0) : "l"(p.rec_off + n0));
        asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(pf_off_end) : "l"(p.rec_off + n1));
        asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(pf_my_off) : "l"(p.rec_off + r));
        asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(pf_my_end) : "l"(p.rec_off + r + 1));
    };
    prefetch_offsets(blockIdx.x);
    bool pre_issued = false;     // the bulk copy of this tile was started during the previous tile's copy-out

    for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const long long t0 = tile * kThreads;
        const long long t1 = min(p.n_reads, t0 + (long long)kThreads);
        const uint32_t off_end = pf_off_end, tile_off0 = pf_off0, tile_my_off = pf_my_off, tile_my_end = pf_my_end;
        long long r0 = t0;
        while (r0 < t1) {
            // ---- sub-tile [r0, r1): as many reads as fit the staging buffer (normally the whole tile)
            const bool first_sub = r0 == t0;
            const uint32_t off0 = first_sub ? tile_off0 : __ldg(p.rec_off + r0);
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
            const uint32_t off1 = (r1 == t1) ? off_end : __ldg(p.rec_off + r1);
            const uint32_t bytes = oversized ? 0u : (off1 - off0) << 4;
            if (tid == 0 && bytes > 0 && !pre_issued) {
                fence_proxy_async();
                mbar_expect_tx(bar, bytes);
                tma_bulk_g2s(sm.stage, p.records + ((unsigned long long)off0 << 4), bytes, bar);
            }
            pre_issued = false;
            const long long r = r0 + tid;
            const bool inrange = r < r1;
            uint32_t my_off = off0, my_end = off0;
            if (inrange) {
                if (first_sub) { my_off = tile_my_off; my_end = tile_my_end; }
                else { my_off = __ldg(p.rec_off + r); my_end = __ldg(p.rec_off + r + 1); }
            }
            if (first_sub) prefetch_offsets(tile + gridDim.x);
            const size_t unit_bytes = (size_t)(my_end - my_off) << 4;
            if (bytes > 0) {
                mbar_wait(bar, parity);
                parity ^= 1;
            }

            // ---- header
            const uint8_t *rec = sm.stage + ((size_t)(my_off - off0) << 4);
            int ca = 0, cc = 0, cg = 0, ct = 0;
            uint32_t status = 0;
            int l_seq = 0, n_cigar = 0, n_tok = 0;
            bool good = inrange, slow = false;
            if (inrange) {
                // a record larger than the staging buffer is never staged
                if (oversized || unit_bytes < 16) slow = true;
                else {
                    const uint4 h = *reinterpret_cast<const uint4 *>(rec);
                    if ((h.w >> 16) & DYN_REC_HAS_MATE) slow = true;
                    else {
                        l_seq = h.z & 0xffff; n_cigar = h.z >> 16; n_tok = h.w & 0xffff;
                        // truncated / corrupt record, or an MD tag that disagrees with the CIGAR
                        if (record_need(l_seq, n_cigar, n_tok, ((h.w >> 16) & DYN_REC_LEAN) != 0) > unit_bytes ||
                            ((h.w >> 16) & DYN_REC_BAD_MD)) {
                            good = false; l_seq = n_cigar = n_tok = 0;
                        }
                    }
                }
            }
            if (slow) good = false;
            const uint32_t *cig = reinterpret_cast<const uint32_t *>(rec + 16);
            const uint32_t *tok = cig + n_cigar;
            const uint32_t *seqw = tok + n_tok;
            sm.recoff[tid] = (uint32_t)(my_off - off0) << 4;

#ifdef DYN_TALLY_DEBUG_SKIP
            if (DYN_TALLY_DEBUG_SKIP >= 2) { l_seq = 0; n_cigar = 0; }
            n_tok = 0;
#endif
            // ---- 1. CIGAR pre-scan: totals, the soft clips at either end, number of interior query ranges
            int m_total = 0, d_total = 0, n_rng = 0, lead = 0, trail = 0;
            {
                int q = 0;
                for (int ci = 0; ci < n_cigar; ++ci) {
                    const uint32_t c = cig[ci];
                    const int op = c & 0xf, len = (int)(c >> 4);
                    if (op == 0 || op == 7 || op == 8) { m_total += len; q += len; }
                    else if (op == 1 || op == 4) {
                        if (op == 4 && ci == 0) lead = len;
                        else if (op == 4 && ci == n_cigar - 1) trail = len;
                        else if (len > 0 && q + len <= l_seq) ++n_rng;
                        q += len;
                    } else if (op == 2) d_total += len;
                }
                if (q != l_seq) good = false;
            }

            // ---- 2. base content of the query between the end clips
            {
                const int qa = min(lead, l_seq), qb = max(qa, l_seq - trail);
                count_query(lut, seqw, qa, qb, ca, cc, cg, ct);
            }

            // ---- 3. event counts are known before the MD tokens are read (every deleted reference base has a
            //         token, every other token is a mismatch): one packed warp scan places every read's ranges
            //         and mismatches in the warp's list
            if (n_tok < d_total) good = false;
            if (good && (n_tok - d_total > kWarpEvents || n_rng > kWarpEvents)) { slow = true; good = false; }
            const int n_mm = good ? n_tok - d_total : 0;
            const int n_rs = good ? n_rng : 0;
            uint32_t incl = (uint32_t)n_rs | ((uint32_t)n_mm << 16);
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
            const int n_ranges = (int)(tot & 0xffffu);
            int n_events = n_ranges + (int)(tot >> 16);
            const int first_rng = (int)(incl & 0xffffu) - n_rs;
            const int first = n_ranges + (int)(incl >> 16) - n_mm;          // first mismatch event of this read
            if (n_events > kWarpEvents) {
                // the warp's list is full: reads whose events do not fit go to the sequential kernel
                if (good && (first_rng + n_rs > kWarpEvents || first + n_mm > kWarpEvents)) { slow = true; good = false; }
                n_events = kWarpEvents;
                for (int j = lane; j < kWarpEvents; j += 32) wlist[j] = kEvNull;
                __syncwarp();
            }
            const bool listed = good;       // this read's events are in the list

            // ---- 4. fill the list: interior ranges (rare), then the MD tokens with a warp-uniform trip count;
            //         deleted reference bases count toward the content right away (bam.py:359-360)
            if (listed && n_rs > 0) {
                int q = 0, k = 0;
                for (int ci = 0; ci < n_cigar; ++ci) {
                    const uint32_t c = cig[ci];
                    const int op = c & 0xf, len = (int)(c >> 4);
                    if (op == 0 || op == 7 || op == 8) q += len;
                    else if (op == 1 || op == 4) {
                        if (!(op == 4 && (ci == 0 || ci == n_cigar - 1)) && len > 0 && q + len <= l_seq)
                            wlist[first_rng + k++] = kEvRange | (unsigned long long)q | ((unsigned long long)lane << 24) |
                                                     ((unsigned long long)len << 32);
                        q += len;
                    }
                }
            }
            {
                const int my_tok = listed ? n_tok : 0;
                const int max_tok = __reduce_max_sync(0xffffffffu, my_tok);
                int n_delb = 0, last = -1, k = 0;
                for (int j = 0; j < max_tok; ++j) {
                    if (j < my_tok) {
                        const uint32_t tk = tok[j];
                        const uint32_t code = DYN_TOK_BASE(tk);
                        if (DYN_TOK_IS_DEL(tk)) {
                            ca += (code == 1); cc += (code == 2); cg += (code == 4); ct += (code == 8);
                            ++n_delb;
                        } else {
                            const int aoff = (int)DYN_TOK_AOFF(tk);
                            if (aoff <= last || aoff >= m_total) good = false;   // tokens must ascend inside the aligned span
                            last = aoff;
                            if (k < n_mm)
                                wlist[first + k] = (unsigned long long)((uint32_t)aoff | (code << 16) | ((uint32_t)lane << 24)) |
                                                   ((unsigned long long)DYN_TOK_QUAL(tk) << 32);
                            ++k;
                        }
                    }
                }
                if (listed && (n_delb != d_total || k != n_mm)) good = false;
                if (listed && !good)        // malformed after all: take its events back
                    for (int j = 0; j < n_rs + n_mm; ++j) wlist[j < n_rs ? first_rng + j : first + (j - n_rs)] = kEvNull;
            }
            if (inrange && !slow && !good) status |= DYN_ST_BAD_RECORD;

            // ---- 5. resolve the events lane-balanced (lane i takes event i, whoever's read it belongs to).  Every
            //         chunk of 32 events leaves a ballot of "this event is a mismatch record": per-read record
            //         counts and output positions are popcounts over those ballots.
            sm.acc[0 * kThreads + tid] = ca; sm.acc[1 * kThreads + tid] = cc;
            sm.acc[2 * kThreads + tid] = cg; sm.acc[3 * kThreads + tid] = ct;
            sm.acc[4 * kThreads + tid] = 0; sm.acc[5 * kThreads + tid] = 0; sm.acc[6 * kThreads + tid] = 0;
            __syncwarp();
            uint32_t warp_recs = 0;
            {
                int c = 0;
                for (int e0 = 0; e0 < n_events; e0 += 32, ++c) {
                    const int e = e0 + lane;
                    unsigned long long out = 0ull;
                    if (e < n_events) {
                        const unsigned long long pl = wlist[e];
                        const uint32_t owner = (uint32_t)(pl >> 24) & 31u;
                        out = resolve_event(sm, p, lut, (wid << 5) + (int)owner, pl, status);
                        if (out) out |= (unsigned long long)owner << 48;      // kept for the copy-out below
                        wlist[e] = out;
                    }
                    const uint32_t mask = __ballot_sync(0xffffffffu, out != 0ull);
                    if (lane == 0) { wmask[c] = mask; wpre[c] = warp_recs; }
                    warp_recs += __popc(mask);
                }
                if (lane == 0) { wmask[c] = 0u; wpre[c] = warp_recs; }
                __syncwarp();
            }
            // one global atomic per warp reserves the output range of its 32 reads; issued as soon as the count is known, so
            // that its round trip (11 % of the stall samples when it sat right before the copy-out) hides behind the
            // gather, the CTA barrier and the next tile's copy issue below
            unsigned long long wbase = 0ull;
            if (emit && lane == 0 && warp_recs) wbase = atomicAdd(p.conv_total, (unsigned long long)warp_recs);
            // records among the warp's events [0, x)
            auto recs_before = [&](int x) { return wpre[x >> 5] + (uint32_t)__popc(wmask[x >> 5] & ((1u << (x & 31)) - 1u)); };

            // ---- 6. gather the read's results
            uint32_t w0 = 0, w1 = 0, w2 = 0, pre0 = 0, n_conv = 0;
            if (good) {
                ca = sm.acc[0 * kThreads + tid]; cc = sm.acc[1 * kThreads + tid];
                cg = sm.acc[2 * kThreads + tid]; ct = sm.acc[3 * kThreads + tid];
                w0 = (uint32_t)sm.acc[4 * kThreads + tid]; w1 = (uint32_t)sm.acc[5 * kThreads + tid];
                w2 = (uint32_t)sm.acc[6 * kThreads + tid];
                pre0 = recs_before(first);
                n_conv = recs_before(first + n_mm) - pre0;
            } else { ca = cc = cg = ct = 0; }
            {
                // units for tally_slow_kernel: it writes every output of theirs
                const bool park = inrange && slow;
                const uint32_t pm = __ballot_sync(0xffffffffu, park);
                if (pm) {
                    uint32_t base = 0;
                    if (lane == 0) base = atomicAdd(p.n_slow, (uint32_t)__popc(pm));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (park) p.slow_list[base + __popc(pm & ((1u << lane) - 1u))] = (uint32_t)r;
                }
            }
            // ---- 7. the records are dead from here on (the copy-out below reads the list only): start the next
            //         tile's bulk copy now, so that it lands while this tile's output range is reserved and written
            __syncthreads();
            if (r1 == t1 && tile + gridDim.x < p.n_tiles && pf_off_end > pf_off0 && pf_off_end - pf_off0 <= cap16) {
                if (tid == 0) {
                    const uint32_t nbytes = (pf_off_end - pf_off0) << 4;
                    fence_proxy_async();
                    mbar_expect_tx(bar, nbytes);
                    tma_bulk_g2s(sm.stage, p.records + ((unsigned long long)pf_off0 << 4), nbytes, bar);
                }
                pre_issued = true;
            }
            if (inrange && !slow) store_counts(p, r, ca, cc, cg, ct, w0, w1, w2, status);
            status_acc |= status;

            if (emit) {
                wbase = __shfl_sync(0xffffffffu, wbase, 0);
                bool fits = true;
                if (inrange && !slow) {
                    const unsigned long long mine = wbase + pre0;
                    uint32_t n = n_conv;
                    if (n > 0xffffu) { n = 0xffffu; status_acc |= DYN_ST_COUNTER_OVERFLOW; }
                    fits = mine + n_conv <= (unsigned long long)p.conv_cap;
                    p.conv_off[r] = (uint32_t)mine;
                    p.conv_n[r] = fits ? (uint16_t)n : (uint16_t)0;
                    if (n_conv > 0 && !fits) status_acc |= DYN_ST_CONV_CAPACITY;
                }
                // a warp either fits the conversion buffer or (rarely) straddles its end
                const bool all_fit = __all_sync(0xffffffffu, fits);
                if (!all_fit) sm.acc[tid] = fits ? 1 : 0;
                __syncwarp();
                for (int e0 = n_ranges & ~31; e0 < n_events; e0 += 32) {
                    const int e = e0 + lane;
                    if (e < n_events && ((wmask[e >> 5] >> (e & 31)) & 1u)) {
                        const unsigned long long cr = wlist[e];
                        const int o_tid = (wid << 5) + (int)((cr >> 48) & 31);
                        if (all_fit || sm.acc[o_tid]) {
                            const unsigned long long at = wbase + recs_before(e);
                            p.conv_rec[at] = cr & 0xffffffffffffull;
                            p.conv_read[at] = (uint32_t)(r0 + o_tid);
                        }
                    }
                }
            }
            r0 = r1;
        }
    }
    if (status_acc) atomicOr(p.status, status_acc);
}

// One counting unit in ONE pass: the mate is walked once (content, overlap, its mismatch calls into a short list), the
// read once (content, calls, the mate's calls looked up by reference position); every call is tallied and buffered, so
// that the records can be written after one atomic has reserved their place.  Same rules as slow_unit (which stays as
// the fallback when a unit has more mate calls or records than the buffers hold): bam.py:332-346, 389-419.
constexpr int kMateCalls = 16, kRecBuf = 24;

// "is reference position p an aligned position of the record" for a sequence of ascending p: a cursor over the
// record's aligned blocks instead of a CIGAR scan per query
struct AlignedCursor {
    const uint32_t *cig;
    int n_cigar, ci, lo, hi, r;        // current aligned block [lo, hi); r = reference position after the ops consumed
    __device__ __forceinline__ void init(const uint32_t *c, int n, int pos) { cig = c; n_cigar = n; ci = 0; lo = hi = r = pos; }
    __device__ __forceinline__ bool covers(int p) {
        while (p >= hi && ci < n_cigar) {
            const uint32_t c = cig[ci++];
            const int op = (int)(c & 0xf), len = (int)(c >> 4);
            if (op == 0 || op == 7 || op == 8) { lo = r; hi = r + len; r += len; }
            else if (op == 2 || op == 3) r += len;
        }
        return p >= lo && p < hi;
    }
};

__device__ __forceinline__ bool slow_unit_onepass(const uint8_t *rec, size_t unit_bytes, const TallyParams &p, SlowOut &o,
                                                  unsigned long long *buf) {
    o.content[0] = o.content[1] = o.content[2] = o.content[3] = 0;
    o.w0 = o.w1 = o.w2 = 0; o.n_conv = 0; o.status = 0;
    const bool nasc = (p.flags & DYN_TALLY_NASC) != 0;
    Walker rd;
    const size_t rd_bytes = rd.init(rec, unit_bytes);
    if (rd.bad) { o.status |= DYN_ST_BAD_RECORD; return true; }
    const bool has_mate = (reinterpret_cast<const uint16_t *>(rec)[7] & DYN_REC_HAS_MATE) != 0;
    if (has_mate && rd.lean) { o.status |= DYN_ST_BAD_RECORD; return true; }      // the mate rule needs every quality
    const int rd_pos = rd.rpos;
    bool bad = false, overflow = false;
    // A/C/G/T counters as four 16-bit fields of one register (l_seq <= 65535 per record; two records per unit are
    // checked against overflow when unpacked)
    unsigned long long content = 0ull, overlap = 0ull;
    // the mate's mismatch calls in walk order = ascending reference position: position, (ref << 4 | read) codes, Phred
    int m_pos[kMateCalls];
    uint32_t m_val[kMateCalls];
    int n_m = 0;
    auto call = [&](int rpos, uint32_t gn, uint32_t rn, uint32_t q) {
        if (o.n_conv < (uint32_t)kRecBuf)
            buf[o.n_conv] = (unsigned long long)(uint32_t)rpos | ((unsigned long long)((gn << 4) | rn) << 32) | ((unsigned long long)q << 40);
        else overflow = true;
        ++o.n_conv;
        if ((int)q <= p.quality) return;                // conversion.py:424 (strict)
        const int r = base_idx(gn), d = base_idx(rn);
        if (r < 0 || d < 0) { o.status |= DYN_ST_NON_ACGT_CONV; return; }
        const int idx = conv_idx(r, d);
        if (p.n_snps > 0 &&
            snp_lookup(p.snps, p.n_snps, ((unsigned long long)idx << 56) | ((unsigned long long)(uint32_t)rd.ref_id << 32) | (uint32_t)rpos))
            return;
        bump_counter(o.w0, o.w1, o.w2, idx, o.status);
    };
    int mate_del[4] = {0, 0, 0, 0};
    if (has_mate) {
        Walker m;
        m.init(rec + rd_bytes, unit_bytes - rd_bytes);
        if (m.bad) { o.status |= DYN_ST_BAD_RECORD; return true; }
        AlignedCursor cur;
        cur.init(rd.cig, rd.n_cigar, rd_pos);
        int prev_cr = -0x7fffffff;
        while (m.next()) {
            if (m.cr < prev_cr) { overflow = true; break; }          // never expected: positions ascend within a record
            prev_cr = m.cr;
            const int b = base_idx(m.gn);
            if (b >= 0) content += 1ull << (16 * b);
            if (m.gn == 15 || m.rn == 15) continue;
            const bool in_read = cur.covers(m.cr);
            if (b >= 0 && in_read) overlap += 1ull << (16 * b);
            if (m.gn == m.rn) continue;
            if (nasc && in_read) continue;                               // bam.py:339
            if (n_m < kMateCalls) { m_pos[n_m] = m.cr; m_val[n_m] = (m.gn << 4) | m.rn | (m.q << 8); ++n_m; }
            else overflow = true;
        }
        bad |= m.bad;
        for (int b = 0; b < 4; ++b) mate_del[b] = m.del[b];
        if (overflow) return false;
    }
    // content(mate) - overlap, floored at zero, before the read's own content is added (bam.py:332-346)
    int cont[4];
    for (int b = 0; b < 4; ++b)
        cont[b] = max(0, (int)((content >> (16 * b)) & 0xffffu) + mate_del[b] - (int)((overlap >> (16 * b)) & 0xffffu));
    content = 0ull;
    // the read's walk; the mate's calls are met in ascending position, so one cursor over the list is enough
    int e0 = 0;
    int next_pos = n_m > 0 ? m_pos[0] : 0x7fffffff;
    unsigned consumed = 0u;
    while (rd.next()) {
        const int b = base_idx(rd.gn);
        if (b >= 0) content += 1ull << (16 * b);
        if (rd.gn == 15 || rd.rn == 15) continue;              // bam.py:389-390
        uint32_t rn = rd.rn, q = rd.q;
        while (next_pos < rd.cr) { ++e0; next_pos = e0 < n_m ? m_pos[e0] : 0x7fffffff; }
        if (next_pos == rd.cr) {
            // the read holds a call at the mate's position: the entry is consumed; the mate wins only if strictly better
            const uint32_t mv = m_val[e0], mq = (mv >> 8) & 0xffu;
            if (!nasc && mq > q) { rn = mv & 0xfu; q = mq; }
            consumed |= 1u << e0;
        }
        if (rd.gn != rn) call(rd.cr, rd.gn, rn, q);
    }
    bad |= rd.bad;
    for (int b = 0; b < 4; ++b) o.content[b] = cont[b] + (int)((content >> (16 * b)) & 0xffffu) + rd.del[b];
    // leftover mate calls, in the mate's walk order (bam.py:414-419)
    for (int e = 0; e < n_m; ++e)
        if (!((consumed >> e) & 1u)) call(m_pos[e], (m_val[e] >> 4) & 0xfu, m_val[e] & 0xfu, (m_val[e] >> 8) & 0xffu);
    if (bad) {
        o.status |= DYN_ST_BAD_RECORD;
        o.content[0] = o.content[1] = o.content[2] = o.content[3] = 0; o.w0 = o.w1 = o.w2 = 0; o.n_conv = 0;
        return true;
    }
    return !overflow;
}

// Units the tile kernel parked: one thread per unit, straight from global memory.  Normally one pass
// (slow_unit_onepass); units that overflow its buffers take slow_unit twice: count, reserve the output range, write.
#ifndef DYN_SLOW_MIN_CTAS
#define DYN_SLOW_MIN_CTAS 6
#endif
__global__ void __launch_bounds__(128, DYN_SLOW_MIN_CTAS) tally_slow_kernel(const TallyParams p) {
    const uint32_t n = *p.n_slow;
    const bool emit = (p.flags & DYN_TALLY_EMIT_CONV) != 0;
    uint32_t status = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t r = p.slow_list[i];
        const uint32_t off = __ldg(p.rec_off + r), end = __ldg(p.rec_off + r + 1);
        const uint8_t *rec = p.records + ((size_t)off << 4);
        const size_t unit_bytes = (size_t)(end - off) << 4;
        SlowOut so;
        unsigned long long buf[kRecBuf];
        const bool onepass = slow_unit_onepass(rec, unit_bytes, p, so, buf);
        if (!onepass) slow_unit<false>(rec, unit_bytes, p, r, so, nullptr, nullptr);
        status |= so.status;
        store_counts(p, r, so.content[0], so.content[1], so.content[2], so.content[3], so.w0, so.w1, so.w2, status);
        if (emit) {
            const uint32_t n_conv = so.n_conv;
            const unsigned long long mine = n_conv ? atomicAdd(p.conv_total, (unsigned long long)n_conv) : 0ull;
            uint32_t nn = n_conv;
            if (nn > 0xffffu) { nn = 0xffffu; status |= DYN_ST_COUNTER_OVERFLOW; }
            const bool fits = mine + n_conv <= (unsigned long long)p.conv_cap;
            p.conv_off[r] = (uint32_t)mine;
            p.conv_n[r] = fits ? (uint16_t)nn : (uint16_t)0;
            if (n_conv > 0 && !fits) status |= DYN_ST_CONV_CAPACITY;
            if (n_conv > 0 && fits) {
                if (onepass) {
                    for (uint32_t j = 0; j < n_conv; ++j) { p.conv_rec[mine + j] = buf[j]; p.conv_read[mine + j] = r; }
                } else slow_unit<true>(rec, unit_bytes, p, r, so, p.conv_rec + mine, p.conv_read + mine);
            }
        }
    }
    if (status) atomicOr(p.status, status);
}

constexpr size_t kFixedSmem = sizeof(unsigned long long) * kWarps * kWarpEvents + sizeof(int) * 7 * kThreads +
                              sizeof(uint32_t) * kThreads + 16 + sizeof(uint32_t) * 4 +
                              sizeof(uint32_t) * 2 * kWarps * (kEvSlots + 1);

}  // namespace

extern "C" int dyn_tally_reads(dyn_ctx_t *ctx, const void *d_records, const uint32_t *d_rec_off, int64_t n_reads,
                               const uint64_t *d_snps, int64_t n_snps, int quality, uint32_t flags,
                               uint8_t *d_counts, uint32_t *d_conv_off, uint16_t *d_conv_n, uint64_t *d_conv_rec,
                               uint32_t *d_conv_read, int64_t conv_capacity, uint64_t *d_conv_total,
                               uint32_t *d_status, void *stream) {
    DYN_ARG(ctx != nullptr, "null context");
    CtxEnter enter_(ctx, stream);
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
    // staging buffer: one tile of kThreads records.  The caller may pass the mean record size (in 16-byte
    // units) in flags bits 16..31 (DYN_TALLY_REC16_HINT); default fits 91-nt reads (~12 units).
    uint32_t hint16 = flags >> 16;
    if (hint16 == 0) hint16 = 12;
    size_t stage = ((size_t)kThreads * hint16 * 16 + DYN_TALLY_SLACK + 255) & ~(size_t)255;
    const size_t max_stage = ctx->smem_optin - kFixedSmem - 1024;
    if (stage > max_stage) stage = max_stage & ~(size_t)1023;
    if (stage < DYN_TALLY_MIN_STAGE) stage = DYN_TALLY_MIN_STAGE;
    p.stage_bytes = (uint32_t)stage;
    p.n_tiles = (n_reads + kThreads - 1) / kThreads;
    const size_t smem = stage + kFixedSmem;
    // parked units: one list per stream in use (the host-buffer entry point runs two)
    void *park = nullptr;
    const int park_slot = (stream == static_cast<void *>(ctx->copy_stream[1])) ? 15 : 14;
    int rc = dyn_ctx_scratch(ctx, park_slot, 256 + sizeof(uint32_t) * (size_t)n_reads, &park);
    if (rc) return rc;
    p.n_slow = static_cast<uint32_t *>(park);
    p.slow_list = p.n_slow + 64;
    DYN_CUDA(cudaMemsetAsync(p.n_slow, 0, sizeof(uint32_t), static_cast<cudaStream_t>(stream)));

    DYN_CUDA(cudaFuncSetAttribute(tally_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    DYN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tally_kernel, kThreads, smem));
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)ctx->sm_count * per_sm;
    if (grid > p.n_tiles) grid = p.n_tiles;
    tally_kernel<<<(unsigned)grid, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(p);
    DYN_CUDA(cudaGetLastError());
    long long slow_grid = std::min<long long>((long long)ctx->sm_count * 2 * DYN_SLOW_MIN_CTAS, (n_reads + 127) / 128);
    tally_slow_kernel<<<(unsigned)slow_grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(p);
    DYN_CUDA(cudaGetLastError());
    return DYN_OK;
}
