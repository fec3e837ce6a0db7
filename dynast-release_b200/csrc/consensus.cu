// UMI consensus calling.
//
// Replaces dynast/preprocessing/consensus.py:57-200 (call_consensus_from_reads: a Python loop over
// get_aligned_pairs of every read plus a Python loop over every reference position of the group's span,
// run in worker processes fed with SAM strings).
//
// Groups of at most 256 reads that touch at most 32 blocks of 32 reference positions take the DIRECT path: one
// warp per group accumulates those blocks in shared memory and emits the cells (direct_kernel) -- no tuples, no
// sort.  The rest (very deep or very scattered groups) is the sort / segmented-reduce the north star names:
//   1. expand   one warp per member read (lanes over the aligned offsets and the deleted-base tokens, coalesced
//               writes) emits one tuple per aligned or deleted reference base: key = (group << 32 | reference position), value = (Phred, read
//               base, reference base, deletion flag).  Bases the reference skips (reference N, read N,
//               consensus.py:79-91) emit a sentinel key.  Per group min start / max end by atomics.
//   2. sort     cub::DeviceRadixSort on the 64-bit keys (stable: within a position tuples stay in the
//               group's read order, so "first seen reference base" is the first tuple of a run).
//   3. reduce   one thread per run of equal keys: the four quality sums, the consensus call with the tie
//               rules of consensus.py:137-146, quality clipped to 42 -> one 32-bit cell per covered position.
//   4. assemble one thread per group walks its cells: run-length CIGAR (gaps become N ops), MD and NM exactly
//               as consensus.py:116-176 builds them, 4-bit sequence and qualities -> one packed record
//               (the same layout the tally reads -- MD as tokens -- so `dynast count` can run on the result
//               without a BAM round trip) plus the MD text for the BAM writer.  Sizes first, exclusive scans,
//               then the write pass.
// HBM-bound on the sort: 12 B per base tuple, moved ~2 x 8 radix passes; algorithmic bytes = the group's
// packed records in + one record out.
#include <cub/cub.cuh>

#include "common.cuh"

#include <algorithm>
#include "walker.cuh"

namespace {

constexpr int kBlock = 128;
inline unsigned grid_for(long long n) { return (unsigned)((n + kBlock - 1) / kBlock); }

// sentinel of skipped bases: group field = n_groups (one past the last group), set per launch

// tuple value: [0..7] Phred, [8..11] read base (BAM nibble), [12..15] reference base, [16] deletion
__device__ __forceinline__ uint32_t pack_val(uint32_t q, uint32_t rn, uint32_t gn, uint32_t del) {
    return (q & 0xff) | (rn << 8) | (gn << 12) | (del << 16);
}

__device__ __forceinline__ long long group_of(const long long *group_off, long long n_groups, long long item) {
    long long lo = 0, hi = n_groups;   // last g with group_off[g] <= item
    while (lo + 1 < hi) { const long long mid = (lo + hi) >> 1; if (group_off[mid] <= item) lo = mid; else hi = mid; }
    return lo;
}

__global__ void size_kernel(const uint8_t *records, const uint32_t *item_rec16, const long long *group_off, long long n_groups,
                            long long n_items, unsigned long long *tuple_cnt, uint32_t *g_left, uint32_t *g_right, int32_t *g_ref,
                            uint32_t *g_nref, int32_t *status) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    const uint8_t *rec = records + ((size_t)item_rec16[i] << 4);
    const uint4 h = *reinterpret_cast<const uint4 *>(rec);
    const uint32_t n_cigar = h.z >> 16;
    const uint32_t *cig = reinterpret_cast<const uint32_t *>(rec + 16);
    uint32_t n_ref = 0, span = 0;
    for (uint32_t c = 0; c < n_cigar; ++c) {
        const uint32_t op = cig[c] & 0xf, len = cig[c] >> 4;
        if (op == 0 || op == 2 || op == 7 || op == 8) { n_ref += len; span += len; }
        else if (op == 3) span += len;
    }
    tuple_cnt[i] = n_ref;
    const long long g = group_of(group_off, n_groups, i);
    atomicMin(g_left + g, h.x);
    atomicMax(g_right + g, h.x + span);
    atomicAdd(g_nref + g, min(n_ref, 0xffffu));
    const int old = atomicCAS(g_ref + g, -1, (int)h.y);
    if (old != -1 && old != (int)h.y) status[g] = 4;   // reads on several contigs: the reference raises (consensus.py:57-58)
}

// One WARP per member read: lane l takes aligned offsets l, l + 32, ... (query / reference position through the
// CIGAR, reference base = read base unless an MD mismatch token sits at that offset) and the read's deleted-base
// tokens, so the 12-byte tuples of a read are written by consecutive lanes into consecutive slots
// (slot = aligned offset for M/=/X bases, m_total + d for the d-th deleted base).
__device__ __forceinline__ void expand_item(const uint8_t *records, const uint32_t *item_rec16, const long long *group_off,
                                            long long n_groups, long long i, int lane, const unsigned long long *tuple_off,
                                            const unsigned long long *win, unsigned long long *key, uint32_t *val,
                                            int32_t *status) {
    const unsigned long long kInvalid = ((unsigned long long)n_groups << 32) | 0xffffffffull;
    const unsigned long long o0 = tuple_off[i], end = tuple_off[i + 1];
    if (end == o0) return;             // a read of direct_kernel's groups (or one without reference bases)
    const uint8_t *rec = records + ((size_t)item_rec16[i] << 4);
    const unsigned long long g = (unsigned long long)group_of(group_off, n_groups, i);
    if (win[g] != 0ull) return;
    const uint4 h = *reinterpret_cast<const uint4 *>(rec);
    const int pos = (int)h.x;
    const uint32_t l_seq = h.z & 0xffff, n_cigar = h.z >> 16, n_tok = h.w & 0xffff;
    const uint32_t *cig = reinterpret_cast<const uint32_t *>(rec + 16);
    const uint32_t *tok = cig + n_cigar;
    const uint8_t *seq = reinterpret_cast<const uint8_t *>(tok + n_tok);
    const uint8_t *qual = seq + pad4(((size_t)l_seq + 1) >> 1);
    uint32_t m_total = 0, d_total = 0;
    for (uint32_t c = 0; c < n_cigar; ++c) {
        const uint32_t op = cig[c] & 0xf, len = cig[c] >> 4;
        if (op == 0 || op == 7 || op == 8) m_total += len;
        else if (op == 2) d_total += len;
    }
    bool bad = ((h.w >> 16) & (DYN_REC_BAD_MD | DYN_REC_LEAN)) != 0 || (unsigned long long)(m_total + d_total) != end - o0;
    if (bad) {
        for (unsigned long long o = o0 + lane; o < end; o += 32) { key[o] = kInvalid; val[o] = 0; }
        if (lane == 0) status[g] = 2;
        return;
    }
    bool iupac = false;      // a base other than A/C/G/T/N: BASE_IDX[...] raises KeyError in the reference (consensus.py:84-97)
    // aligned bases
    for (uint32_t a = lane; a < m_total; a += 32) {
        int q = 0, r = pos, done = 0, qpos = -1, rpos = 0;
        for (uint32_t c = 0; c < n_cigar; ++c) {
            const uint32_t w = cig[c];
            const int op = (int)(w & 0xf), len = (int)(w >> 4);
            if (op == 0 || op == 7 || op == 8) {
                if ((int)a < done + len) { qpos = q + ((int)a - done); rpos = r + ((int)a - done); break; }
                done += len; q += len; r += len;
            } else if (op == 1 || op == 4) q += len;
            else if (op == 2 || op == 3) r += len;
        }
        unsigned long long k = kInvalid;
        uint32_t v = 0;
        if (qpos >= 0 && qpos < (int)l_seq) {
            const uint32_t rn = (seq[qpos >> 1] >> ((~qpos & 1) << 2)) & 0xf;
            uint32_t gn = rn;
            for (uint32_t t = 0; t < n_tok; ++t) {
                const uint32_t tk = tok[t];
                if (!DYN_TOK_IS_DEL(tk) && DYN_TOK_AOFF(tk) == a) gn = DYN_TOK_BASE(tk);
            }
            if (gn != 15 && rn != 15 && (base_idx(gn) < 0 || base_idx(rn) < 0)) iupac = true;
            if (base_idx(gn) >= 0 && base_idx(rn) >= 0) {      // reference N and read N are skipped (consensus.py:79-91)
                k = (g << 32) | (uint32_t)rpos;
                v = pack_val(qual[qpos], rn, gn, 0);
            }
        } else bad = true;
        key[o0 + a] = k; val[o0 + a] = v;
    }
    // deleted reference bases: the d-th deleted-base token belongs to the d-th base of the CIGAR's D ops
    uint32_t carry = 0;
    for (uint32_t t0 = 0; t0 < n_tok; t0 += 32) {
        const uint32_t t = t0 + lane;
        const uint32_t tk = t < n_tok ? tok[t] : 0u;
        const bool is_del = t < n_tok && DYN_TOK_IS_DEL(tk);
        const uint32_t ballot = __ballot_sync(0xffffffffu, is_del);
        if (is_del) {
            uint32_t d = carry + (uint32_t)__popc(ballot & ((1u << lane) - 1u));
            unsigned long long k = kInvalid;
            uint32_t v = 0;
            if (d < d_total) {
                int r = pos;
                uint32_t left = d;
                int rpos = -1;
                for (uint32_t c = 0; c < n_cigar; ++c) {
                    const uint32_t w = cig[c];
                    const int op = (int)(w & 0xf), len = (int)(w >> 4);
                    if (op == 2) { if (left < (uint32_t)len) { rpos = r + (int)left; break; } left -= (uint32_t)len; r += len; }
                    else if (op == 0 || op == 3 || op == 7 || op == 8) r += len;
                }
                const uint32_t gn = DYN_TOK_BASE(tk);
                if (gn != 15 && base_idx(gn) < 0) iupac = true;
                if (rpos >= 0 && base_idx(gn) >= 0) { k = (g << 32) | (uint32_t)rpos; v = pack_val(0, 0, gn, 1); }
                key[o0 + m_total + d] = k; val[o0 + m_total + d] = v;
            } else bad = true;
        }
        carry += (uint32_t)__popc(ballot);
    }
    if (carry != d_total) {     // fewer deleted-base tokens than the CIGAR deletes: the unwritten slots must not be read
        for (uint32_t d = carry + lane; d < d_total; d += 32) { key[o0 + m_total + d] = kInvalid; val[o0 + m_total + d] = 0; }
        bad = true;
    }
    if (bad) status[g] = 2;
    if (iupac) status[g] = 6;
}

// Most reads belong to direct_kernel's groups and have no tuples: every lane checks one read, and the warp then
// expands the few that do, one after the other (launching a warp per read just to find that out cost 2.5 ms).
__global__ void expand_kernel(const uint8_t *records, const uint32_t *item_rec16, const long long *group_off, long long n_groups,
                              long long n_items, const unsigned long long *tuple_off, const unsigned long long *win,
                              unsigned long long *key, uint32_t *val, int32_t *status) {
    const int lane = threadIdx.x & 31;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long base = warp * 32; base < n_items; base += n_warps * 32) {
        const long long mine = base + lane;
        uint32_t todo = __ballot_sync(0xffffffffu, mine < n_items && tuple_off[mine + 1] != tuple_off[mine]);
        while (todo) {
            const int j = __ffs((int)todo) - 1;
            todo &= todo - 1;
            expand_item(records, item_rec16, group_off, n_groups, base + j, lane, tuple_off, win, key, val, status);
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Direct path: groups of at most 256 reads (16-bit quality sums cannot overflow) whose reads touch at most
// kSlots blocks of kBlk reference positions never become tuples.  One WARP per group keeps those blocks in shared
// memory -- four quality sums and (covered, first-seen reference base) per position; a block is found through a
// 32-entry hash of its block number, so spliced reads cost a few more blocks, not the whole intron --, walks the
// member reads in group order exactly as expand_kernel does (lanes over aligned offsets and deleted-base tokens)
// and then emits the covered positions' cells block by block in ascending order.  A group that touches more
// blocks gives up (win[g] = 0) and takes the tuple sort below with the very deep groups.
// ---------------------------------------------------------------------------------------------------
constexpr int kBlk = 32, kSlots = 32, kWin = kBlk * kSlots;
constexpr int kDirectWarps = 4;
constexpr int kNoBlock = -1;

// win[g] = cell slots reserved for group g on the direct path (0: not eligible)
__global__ void classify_groups_kernel(const uint32_t *g_left, const uint32_t *g_nref, const long long *group_off,
                                       long long n_groups, unsigned long long *win) {
    const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (g > n_groups) return;
    unsigned long long w = 0;
    if (g < n_groups && g_left[g] != 0xffffffffu && group_off[g + 1] - group_off[g] <= 256)
        w = (min(g_nref[g], (uint32_t)kWin) + 3u) & ~3u;      // multiple of 4: a group's cells start 16-byte aligned
    win[g] = w;        // win[n_groups] = 0: the exclusive scan's total lands there
}

__global__ void drop_direct_items_kernel(const long long *group_off, long long n_groups, long long n_items,
                                         const unsigned long long *win, unsigned long long *tuple_cnt) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    if (win[group_of(group_off, n_groups, i)] != 0ull) tuple_cnt[i] = 0ull;
}

__device__ __forceinline__ uint32_t call_cell(const uint32_t s[4], int ref, int quality) {
    const uint32_t base_q = max(max(s[0], s[1]), max(s[2], s[3]));
    if (base_q == 0) return ((uint32_t)ref << 2) | (1u << 4);     // no support: deletion (consensus.py:113-123)
    int base;
    if ((int)base_q < quality) base = ref;
    else if (s[ref] == base_q) base = ref;                         // ties: reference if present
    else base = s[0] == base_q ? 0 : (s[1] == base_q ? 1 : (s[2] == base_q ? 2 : 3));
    return (uint32_t)base | ((uint32_t)ref << 2) | (min(base_q, 42u) << 8);
}

// slot of reference block `blk` in the warp's table (claimed on first use); -1 when the table is full
__device__ __forceinline__ int block_slot(int *blk_id, int blk) {
    const uint32_t h = ((uint32_t)blk * 0x9E3779B1u) >> 27;
    for (int probe = 0; probe < kSlots; ++probe) {
        const int s = (int)((h + probe) & (kSlots - 1));
        const int cur = blk_id[s];
        if (cur == blk) return s;
        if (cur == kNoBlock) {
            const int old = atomicCAS(blk_id + s, kNoBlock, blk);
            if (old == kNoBlock || old == blk) return s;
        }
    }
    return -1;
}

__global__ void __launch_bounds__(32 * kDirectWarps) direct_kernel(const uint8_t *records, const uint32_t *item_rec16,
                                                                  const long long *group_off, long long n_groups,
                                                                  const uint32_t *g_left, unsigned long long *win,
                                                                  const unsigned long long *cell_start, int quality, uint32_t *cell,
                                                                  uint32_t *cell_cnt, int32_t *status) {
    __shared__ __align__(8) uint16_t s_qs[kDirectWarps][kWin][4];
    __shared__ uint8_t s_meta[kDirectWarps][kWin];      // bit 2 covered, bits 0..1 first-seen reference base
    __shared__ int s_blk[kDirectWarps][kSlots];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint16_t(*qs)[4] = s_qs[wid];
    uint8_t *meta = s_meta[wid];
    int *blk_id = s_blk[wid];
    for (int w = lane; w < kWin; w += 32) { *reinterpret_cast<uint2 *>(qs[w]) = make_uint2(0u, 0u); meta[w] = 0; }
    blk_id[lane] = kNoBlock;
    __syncwarp();
    const long long n_warps = (long long)gridDim.x * kDirectWarps;
    for (long long g = blockIdx.x * (long long)kDirectWarps + wid; g < n_groups; g += n_warps) {
        if (win[g] == 0ull) continue;
        bool bad = false, full = false, iupac = false;
        for (long long i = group_off[g]; i < group_off[g + 1] && !full; ++i) {
            const uint8_t *rec = records + ((size_t)item_rec16[i] << 4);
            const uint4 h = *reinterpret_cast<const uint4 *>(rec);
            const int pos = (int)h.x;
            const uint32_t l_seq = h.z & 0xffff, n_cigar = h.z >> 16, n_tok = h.w & 0xffff;
            const uint32_t *cig = reinterpret_cast<const uint32_t *>(rec + 16);
            const uint32_t *tok = cig + n_cigar;
            const uint8_t *seq = reinterpret_cast<const uint8_t *>(tok + n_tok);
            const uint8_t *qual = seq + pad4(((size_t)l_seq + 1) >> 1);
            if ((h.w >> 16) & (DYN_REC_BAD_MD | DYN_REC_LEAN)) { bad = true; continue; }
            uint32_t m_total = 0, d_total = 0;
            for (uint32_t c = 0; c < n_cigar; ++c) {
                const uint32_t op = cig[c] & 0xf, len = cig[c] >> 4;
                if (op == 0 || op == 7 || op == 8) m_total += len;
                else if (op == 2) d_total += len;
            }
            // aligned bases
            for (uint32_t a = lane; a < m_total; a += 32) {
                int q = 0, r = pos, done = 0, qpos = -1, rpos = 0;
                for (uint32_t c = 0; c < n_cigar; ++c) {
                    const uint32_t w = cig[c];
                    const int op = (int)(w & 0xf), len = (int)(w >> 4);
                    if (op == 0 || op == 7 || op == 8) {
                        if ((int)a < done + len) { qpos = q + ((int)a - done); rpos = r + ((int)a - done); break; }
                        done += len; q += len; r += len;
                    } else if (op == 1 || op == 4) q += len;
                    else if (op == 2 || op == 3) r += len;
                }
                if (qpos >= 0 && qpos < (int)l_seq) {
                    const uint32_t rn = (seq[qpos >> 1] >> ((~qpos & 1) << 2)) & 0xf;
                    uint32_t gn = rn;
                    for (uint32_t t = 0; t < n_tok; ++t) {
                        const uint32_t tk = tok[t];
                        if (!DYN_TOK_IS_DEL(tk) && DYN_TOK_AOFF(tk) == a) gn = DYN_TOK_BASE(tk);
                    }
                    const int gi = base_idx(gn), ri = base_idx(rn);
                    if (gn != 15 && rn != 15 && (gi < 0 || ri < 0)) iupac = true;   // BASE_IDX[...] KeyError (consensus.py:84-97)
                    if (gi >= 0 && ri >= 0) {      // reference N and read N are skipped (consensus.py:79-91)
                        const int sl = block_slot(blk_id, rpos >> 5);
                        if (sl < 0) full = true;
                        else {
                            const int w = sl * kBlk + (rpos & (kBlk - 1));
                            if (!(meta[w] & 4)) meta[w] = (uint8_t)(4 | gi);
                            qs[w][ri] = (uint16_t)(qs[w][ri] + qual[qpos]);
                        }
                    }
                } else bad = true;
            }
            // deleted reference bases: the d-th deleted-base token belongs to the d-th base of the CIGAR's D ops
            uint32_t carry = 0;
            for (uint32_t t0 = 0; t0 < n_tok; t0 += 32) {
                const uint32_t t = t0 + lane;
                const uint32_t tk = t < n_tok ? tok[t] : 0u;
                const bool is_del = t < n_tok && DYN_TOK_IS_DEL(tk);
                const uint32_t ballot = __ballot_sync(0xffffffffu, is_del);
                if (is_del) {
                    const uint32_t d = carry + (uint32_t)__popc(ballot & ((1u << lane) - 1u));
                    if (d < d_total) {
                        int r = pos, rpos = -1;
                        uint32_t rest = d;
                        for (uint32_t c = 0; c < n_cigar; ++c) {
                            const uint32_t w = cig[c];
                            const int op = (int)(w & 0xf), len = (int)(w >> 4);
                            if (op == 2) { if (rest < (uint32_t)len) { rpos = r + (int)rest; break; } rest -= (uint32_t)len; r += len; }
                            else if (op == 0 || op == 3 || op == 7 || op == 8) r += len;
                        }
                        const int gi = base_idx(DYN_TOK_BASE(tk));
                        if (gi < 0 && DYN_TOK_BASE(tk) != 15) iupac = true;
                        if (rpos >= 0 && gi >= 0) {
                            const int sl = block_slot(blk_id, rpos >> 5);
                            if (sl < 0) full = true;
                            else {
                                const int w = sl * kBlk + (rpos & (kBlk - 1));
                                if (!(meta[w] & 4)) meta[w] = (uint8_t)(4 | gi);
                            }
                        }
                    } else bad = true;
                }
                carry += (uint32_t)__popc(ballot);
            }
            if (carry != d_total) bad = true;
            full = __any_sync(0xffffffffu, full);
            __syncwarp();       // the next read of the group touches the same positions
        }
        // blocks are emitted in ascending order: repeatedly the smallest block number not yet taken
        const int my_blk = blk_id[lane];
        const uint32_t used = __ballot_sync(0xffffffffu, my_blk != kNoBlock);
        const int n_blk = __popc(used);
        uint32_t my_key = my_blk == kNoBlock ? 0xffffffffu : (uint32_t)my_blk;
        // A direct cell carries, in its upper half, the number of reference positions skipped since the previous
        // cell (the first one: since the group's leftmost start) -- no 8-byte key per cell; a gap that does not fit
        // 16 bits sends the group to the sorted path.
        const unsigned long long base = cell_start[g];
        const unsigned long long cap = win[g];
        uint32_t run = 0;
        uint32_t next_pos = g_left[g];          // position right after the previous cell
        bool wide = false;
        for (int r = 0; r < n_blk; ++r) {
            const uint32_t lowest = __reduce_min_sync(0xffffffffu, my_key);
            const int sl = __ffs((int)__ballot_sync(0xffffffffu, my_key == lowest)) - 1;   // its slot
            const int blk = (int)lowest;
            if (lane == sl) my_key = 0xffffffffu;
            const int w = sl * kBlk + lane;
            const uint32_t m = meta[w];
            const bool covered = !full && (m & 4);
            const uint32_t ballot = __ballot_sync(0xffffffffu, covered);
            if (covered) {
                const uint32_t s[4] = {qs[w][0], qs[w][1], qs[w][2], qs[w][3]};
                const uint32_t below = ballot & ((1u << lane) - 1u);
                const unsigned long long k = run + (uint32_t)__popc(below);
                const uint32_t pos = (uint32_t)(blk * kBlk + lane);
                const uint32_t gap = below ? (uint32_t)(lane - (31 - __clz((int)below)) - 1) : pos - next_pos;
                if (gap > 0xffffu) wide = true;
                if (k < cap) cell[base + k] = call_cell(s, (int)(m & 3), quality) | (gap << 16);
            }
            if (ballot) next_pos = (uint32_t)(blk * kBlk) + (uint32_t)(31 - __clz((int)ballot)) + 1u;
            run += (uint32_t)__popc(ballot);
            // leave the block clean for the next group
            *reinterpret_cast<uint2 *>(qs[w]) = make_uint2(0u, 0u);
            meta[w] = 0;
        }
        full = full || __any_sync(0xffffffffu, wide) || (unsigned long long)run > cap;   // (more cells than reserved: never
                                                                                         // expected, the sorted path is safe)
        __syncwarp();
        blk_id[lane] = kNoBlock;
        __syncwarp();
        if (lane == 0) {
            if (full) win[g] = 0ull;          // too many blocks: the sorted path takes the group
            else cell_cnt[g] = run;
        }
        if (!full && __any_sync(0xffffffffu, bad) && lane == 0) status[g] = 2;
        if (__any_sync(0xffffffffu, iupac) && lane == 0) status[g] = 6;
    }
}

__global__ void head_kernel(const unsigned long long *key, long long n, long long n_groups, uint32_t *flag) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    flag[i] = ((long long)(key[i] >> 32) < n_groups && (i == 0 || key[i] != key[i - 1])) ? 1u : 0u;
}

// cell: [0..1] consensus base, [2..3] reference base, [4] deletion, [8..15] quality
__global__ void reduce_kernel(const unsigned long long *key, const uint32_t *val, const uint32_t *flag, const uint32_t *idx,
                              long long n, int quality, unsigned long long *cell_key, uint32_t *cell) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n || !flag[i]) return;
    const unsigned long long k = key[i];
    uint32_t s[4] = {0, 0, 0, 0};
    const int ref = base_idx((val[i] >> 12) & 0xf);    // first seen (consensus.py:84-90)
    for (long long j = i; j < n && key[j] == k; ++j) {
        const uint32_t v = val[j];
        if (!((v >> 16) & 1u)) s[base_idx((v >> 8) & 0xf)] += v & 0xffu;
    }
    uint32_t out;
    const uint32_t base_q = max(max(s[0], s[1]), max(s[2], s[3]));
    if (base_q == 0) out = ((uint32_t)ref << 2) | (1u << 4);     // no support: deletion (consensus.py:113-123)
    else {
        int base;
        if ((int)base_q < quality) base = ref;
        else if (s[ref] == base_q) base = ref;                   // ties: reference if present
        else base = s[0] == base_q ? 0 : (s[1] == base_q ? 1 : (s[2] == base_q ? 2 : 3));
        out = (uint32_t)base | ((uint32_t)ref << 2) | (min(base_q, 42u) << 8);
    }
    cell_key[idx[i]] = k;
    cell[idx[i]] = out;
}

struct GroupOut { uint32_t n_cigar, l_seq, md_len, nm, n_tok; };

__device__ __forceinline__ int num_digits(uint32_t v) { int d = 1; while (v >= 10) { v /= 10; ++d; } return d; }
__device__ __forceinline__ uint8_t *put_num(uint8_t *p, uint32_t v) {
    const int d = num_digits(v);
    for (int i = d - 1; i >= 0; --i) { p[i] = (uint8_t)('0' + v % 10); v /= 10; }
    return p + d;
}

// One thread per group. WRITE=false: sizes only (stored in `sizes`); WRITE=true: the record at out + 16 * out_off[g].
template <bool WRITE>
__global__ void assemble_kernel(const unsigned long long *cell_key_s, const uint32_t *cell_s, const long long *n_cells_p,
                                const uint32_t *cell_d, const unsigned long long *win,
                                const unsigned long long *cell_start, const uint32_t *cell_cnt, long long n_groups,
                                const uint32_t *g_left, const uint32_t *g_right, const int32_t *g_ref,
                                GroupOut *sizes, uint32_t *out_cap16, uint32_t *md_cap, const uint32_t *out_off, uint8_t *out,
                                long long out_capacity16, const uint32_t *md_off, uint8_t *out_md, long long md_capacity,
                                uint32_t *out_nm, int32_t *status) {
    const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    // cells of the group: direct_kernel's groups know their range; the sorted path's cells are ordered by (group, position)
    // cells of the group: direct_kernel's groups know their range (gap-encoded positions, see there); the sorted path's
    // cells are ordered by (group, position)
    const bool direct = win[g] != 0ull;
    long long c0, c1;
    if (direct) { c0 = (long long)cell_start[g]; c1 = c0 + cell_cnt[g]; }
    else {
        long long lo = 0, hi = *n_cells_p;
        c1 = hi;
        while (lo < hi) { const long long mid = (lo + hi) >> 1; if ((long long)(cell_key_s[mid] >> 32) < g) lo = mid + 1; else hi = mid; }
        c0 = lo;
    }
    const uint32_t left = g_left[g], right = g_right[g];
    const bool empty = left == 0xffffffffu;      // group without reads
    uint32_t *cig = nullptr, *tokp = nullptr;
    uint8_t *seq = nullptr, *qual = nullptr, *md = nullptr;
    GroupOut sz = {0, 0, 0, 0, 0};
    bool fits = true;
    if (WRITE) {
        sz = sizes[g];
        if ((long long)out_off[g + 1] > out_capacity16 || (long long)md_off[g + 1] > md_capacity || status[g] == 3) fits = false;
        uint8_t *rec = out + ((size_t)out_off[g] << 4);
        if (fits) {
            uint32_t *h = reinterpret_cast<uint32_t *>(rec);
            h[0] = empty ? 0u : left;
            h[1] = (uint32_t)g_ref[g];
            h[2] = sz.l_seq | (sz.n_cigar << 16);
            h[3] = sz.n_tok;       // flag 0: the caller sets strand / tags
            cig = reinterpret_cast<uint32_t *>(rec + 16);
            tokp = cig + sz.n_cigar;
            seq = reinterpret_cast<uint8_t *>(tokp + sz.n_tok);
            const size_t seq_bytes = pad4(((size_t)sz.l_seq + 1) >> 1);
            qual = seq + seq_bytes;
            md = out_md + md_off[g];
            // sequence and qualities are written a 32-bit word at a time below; zero what follows the last quality word
            const size_t used4 = 16 + 4 * (size_t)sz.n_cigar + 4 * (size_t)sz.n_tok + seq_bytes + pad4(sz.l_seq);
            for (size_t b = used4; b < (((size_t)out_off[g + 1] - out_off[g]) << 4); b += 4) *reinterpret_cast<uint32_t *>(rec + b) = 0u;
            out_nm[g] = sz.nm;
        }
    }
    uint32_t n_cigar = 0, l_seq = 0, md_len = 0, nm = 0, n_tok = 0;
    uint32_t run_op = 0xffu, run_n = 0;     // ops: 0 M, 2 D, 3 N (BAM codes)
    uint32_t md_n = 0;
    bool md_zero = true, md_del = false;
    uint8_t *mdp = md;
    auto cigar_push = [&](uint32_t op, uint32_t n) {
        if (n == 0) return;
        if (op == run_op) { run_n += n; return; }
        if (run_op != 0xffu) { if (WRITE && fits) cig[n_cigar] = (run_n << 4) | run_op; ++n_cigar; }
        run_op = op; run_n = n;
    };
    auto md_number = [&]() {
        if (md_n > 0 || md_zero) {
            if (WRITE && fits) mdp = put_num(mdp, md_n);
            md_len += num_digits(md_n);
            md_n = 0;
        }
    };
    uint32_t seq_acc = 0, qual_acc = 0;
    uint32_t prev = left;    // next expected reference position
    if (!empty) {
        auto cell_step = [&](uint32_t pos, uint32_t v) {
            cigar_push(3, pos - prev);                 // positions nobody observed (consensus.py:109-111)
            prev = pos + 1;
            const uint32_t ref = (v >> 2) & 3u;
            if ((v >> 4) & 1u) {
                cigar_push(2, 1);
                md_number();
                if (!md_del) { if (WRITE && fits) *mdp++ = '^'; ++md_len; }
                if (WRITE && fits) { *mdp++ = (uint8_t)"ACGT"[ref]; tokp[n_tok] = min(l_seq, 0xffffu) | ((1u << ref) << 16) | (1u << 20); }
                ++md_len; ++n_tok;
                md_del = true;
            } else {
                cigar_push(0, 1);
                md_del = false;
                const uint32_t base = v & 3u;
                if (base == ref) { ++md_n; md_zero = false; }
                else {
                    md_number();
                    if (WRITE && fits) { *mdp++ = (uint8_t)"ACGT"[ref]; tokp[n_tok] = (l_seq & 0xffffu) | ((1u << ref) << 16); }
                    ++md_len; ++n_tok;
                    md_zero = true;
                    ++nm;
                }
                if (WRITE && fits) {
                    // base i of a word sits at byte (i >> 1) & 3, high nibble when i is even
                    seq_acc |= (1u << base) << (8 * ((l_seq >> 1) & 3) + ((l_seq & 1) ? 0 : 4));
                    qual_acc |= ((v >> 8) & 0xffu) << (8 * (l_seq & 3));
                    if ((l_seq & 7) == 7) { reinterpret_cast<uint32_t *>(seq)[l_seq >> 3] = seq_acc; seq_acc = 0; }
                    if ((l_seq & 3) == 3) { reinterpret_cast<uint32_t *>(qual)[l_seq >> 2] = qual_acc; qual_acc = 0; }
                }
                ++l_seq;
            }
        };
        if (direct) {
            // four 4-byte cells per 16-byte load (the group's range starts 16-byte aligned)
            for (long long c = c0; c < c1; c += 4) {
                const uint4 q = *reinterpret_cast<const uint4 *>(cell_d + c);
                const uint32_t vv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (c + u < c1) cell_step(prev + (vv[u] >> 16), vv[u]);
            }
        } else {
            for (long long c = c0; c < c1 && (long long)(cell_key_s[c] >> 32) == g; ++c) cell_step((uint32_t)cell_key_s[c], cell_s[c]);
        }
        cigar_push(3, right - prev);
    }
    if (WRITE && fits) {     // the partial last words (zero padded)
        if (l_seq & 7) reinterpret_cast<uint32_t *>(seq)[l_seq >> 3] = seq_acc;
        if (l_seq & 3) reinterpret_cast<uint32_t *>(qual)[l_seq >> 2] = qual_acc;
    }
    // MD always ends with a number (consensus.py:175); the last CIGAR run is flushed (:176)
    if (WRITE && fits) mdp = put_num(mdp, md_n);
    md_len += num_digits(md_n);
    if (run_op != 0xffu) { if (WRITE && fits) cig[n_cigar] = (run_n << 4) | run_op; ++n_cigar; }
    if (!WRITE) {
        GroupOut o = {n_cigar, l_seq, md_len, nm, n_tok};
        sizes[g] = o;
        const size_t bytes = 16 + 4 * (size_t)n_cigar + 4 * (size_t)n_tok + pad4(((size_t)l_seq + 1) >> 1) + l_seq;
        out_cap16[g] = (uint32_t)((bytes + 15) >> 4);
        md_cap[g] = md_len;
        if (n_cigar > 0xffffu || l_seq > 0xffffu || n_tok > 0xffffu) status[g] = 3;   // does not fit a BAM record
    } else if (!fits && status[g] == 0) status[g] = 5;   // output capacity exceeded
}

__global__ void init_groups_kernel(uint32_t *g_left, uint32_t *g_right, int32_t *g_ref, int32_t *status, long long n) {
    const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (g >= n) return;
    g_left[g] = 0xffffffffu; g_right[g] = 0u; g_ref[g] = -1; status[g] = 0;
}

}  // namespace

extern "C" int dyn_consensus(dyn_ctx_t *ctx, const void *d_records, const uint32_t *d_item_rec16, const int64_t *d_group_off,
                             int64_t n_groups, int64_t n_items, int quality, void *d_out_records,
                             int64_t out_capacity_bytes, uint32_t *d_out_off, void *d_out_md, int64_t md_capacity_bytes,
                             uint32_t *d_out_md_off, uint32_t *d_out_nm, int32_t *d_out_status, void *stream_) {
    DYN_ARG(ctx != nullptr, "null context");
    CtxEnter enter_(ctx, stream_);
    DYN_ARG(n_groups >= 0 && n_items >= 0 && n_groups < 0xffffffffLL, "size out of range");
    if (n_groups == 0) return DYN_OK;
    DYN_ARG(d_group_off && d_out_off && d_out_md_off && d_out_nm && d_out_status, "null buffer");
    DYN_ARG(md_capacity_bytes == 0 || d_out_md, "null MD output");
    DYN_ARG(n_items == 0 || (d_records && d_item_rec16), "null records");
    DYN_ARG(out_capacity_bytes == 0 || d_out_records, "null output");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const long long *group_off = reinterpret_cast<const long long *>(d_group_off);
    const uint8_t *records = static_cast<const uint8_t *>(d_records);

    // ---- per-item tuple counts, per-group span and path (scratch slot 12)
    size_t scan_tmp = 0, t = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, (unsigned long long *)nullptr, (unsigned long long *)nullptr, n_items + 1, stream);
    cub::DeviceScan::ExclusiveSum(nullptr, t, (unsigned long long *)nullptr, (unsigned long long *)nullptr, n_groups + 1, stream);
    if (t > scan_tmp) scan_tmp = t;
    cub::DeviceScan::ExclusiveSum(nullptr, t, (uint32_t *)nullptr, (uint32_t *)nullptr, n_groups + 1, stream);
    if (t > scan_tmp) scan_tmp = t;
    uint32_t *g_left, *g_right, *g_nref, *out_cap16, *md_cap, *cell_cnt;
    int32_t *g_ref;
    unsigned long long *tuple_cnt, *tuple_off, *win, *cell_start;
    GroupOut *sizes;
    long long *n_cells;
    void *tmp0;
    Arena a0;
    for (int pass = 0; pass < 2; ++pass) {
        tuple_cnt = a0.take<unsigned long long>(n_items + 1);
        tuple_off = a0.take<unsigned long long>(n_items + 1);
        win = a0.take<unsigned long long>(n_groups + 1);
        cell_start = a0.take<unsigned long long>(n_groups + 1);
        cell_cnt = a0.take<uint32_t>(n_groups);
        g_left = a0.take<uint32_t>(n_groups); g_right = a0.take<uint32_t>(n_groups); g_ref = a0.take<int32_t>(n_groups);
        g_nref = a0.take<uint32_t>(n_groups);
        out_cap16 = a0.take<uint32_t>(n_groups + 1);
        md_cap = a0.take<uint32_t>(n_groups + 1);
        sizes = a0.take<GroupOut>(n_groups);
        n_cells = a0.take<long long>(2);
        tmp0 = a0.take<uint8_t>(scan_tmp);
        if (pass == 0) {
            void *base = nullptr;
            int rc = dyn_ctx_scratch(ctx, 12, a0.used + 256, &base);
            if (rc) return rc;
            a0 = Arena(base);
        }
    }
    init_groups_kernel<<<grid_for(n_groups), kBlock, 0, stream>>>(g_left, g_right, g_ref, d_out_status, n_groups);
    DYN_CUDA(cudaMemsetAsync(tuple_cnt, 0, sizeof(unsigned long long) * (n_items + 1), stream));
    DYN_CUDA(cudaMemsetAsync(n_cells, 0, 16, stream));
    DYN_CUDA(cudaMemsetAsync(cell_cnt, 0, sizeof(uint32_t) * n_groups, stream));
    DYN_CUDA(cudaMemsetAsync(g_nref, 0, sizeof(uint32_t) * n_groups, stream));
    if (n_items > 0)
        size_kernel<<<grid_for(n_items), kBlock, 0, stream>>>(records, d_item_rec16, group_off, n_groups, n_items, tuple_cnt,
                                                              g_left, g_right, g_ref, g_nref, d_out_status);

    // ---- direct path (scratch slot 13): cell slots per eligible group, then one warp per group
    classify_groups_kernel<<<grid_for(n_groups + 1), kBlock, 0, stream>>>(g_left, g_nref, group_off, n_groups, win);
    t = scan_tmp;
    DYN_CUDA(cub::DeviceScan::ExclusiveSum(tmp0, t, win, cell_start, n_groups + 1, stream));
    unsigned long long n_direct = 0, n_tuples = 0;
    DYN_CUDA(cudaMemcpyAsync(&n_direct, cell_start + n_groups, 8, cudaMemcpyDeviceToHost, stream));
    DYN_CUDA(cudaStreamSynchronize(stream));      // the direct path's cell arrays are sized from this total
    unsigned long long *cell_key = nullptr;
    uint32_t *cell_d = nullptr, *cell = nullptr;
    if (n_direct > 0) {
        Arena a1;
        for (int pass = 0; pass < 2; ++pass) {
            cell_d = a1.take<uint32_t>((size_t)n_direct + 4);
            if (pass == 0) {
                void *base = nullptr;
                int rc = dyn_ctx_scratch(ctx, 13, a1.used + 256, &base);
                if (rc) return rc;
                a1 = Arena(base);
            }
        }
        int per_sm = 0;
        DYN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, direct_kernel, 32 * kDirectWarps, 0));
        if (per_sm < 1) per_sm = 1;
        long long grid = std::min<long long>((long long)ctx->sm_count * per_sm, (n_groups + kDirectWarps - 1) / kDirectWarps);
        direct_kernel<<<(unsigned)grid, 32 * kDirectWarps, 0, stream>>>(records, d_item_rec16, group_off, n_groups, g_left, win,
                                                                        cell_start, quality, cell_d, cell_cnt, d_out_status);
        DYN_CUDA(cudaGetLastError());
    }

    // ---- sorted path for what is left: expand, sort, reduce (scratch slot 17)
    if (n_items > 0) {
        drop_direct_items_kernel<<<grid_for(n_items), kBlock, 0, stream>>>(group_off, n_groups, n_items, win, tuple_cnt);
        t = scan_tmp;
        DYN_CUDA(cub::DeviceScan::ExclusiveSum(tmp0, t, tuple_cnt, tuple_off, n_items + 1, stream));
        DYN_CUDA(cudaMemcpyAsync(&n_tuples, tuple_off + n_items, 8, cudaMemcpyDeviceToHost, stream));
        DYN_CUDA(cudaStreamSynchronize(stream));      // the tuple arrays are sized from this total
    }
    DYN_ARG(n_tuples < 0x7fffffffull, "more than 2^31 base tuples outside the direct path in one call: split the groups into batches");
    if (n_tuples > 0) {
        const size_t n = (size_t)n_tuples;
        size_t sort_tmp = 0;
        {
            cub::DoubleBuffer<unsigned long long> k(nullptr, nullptr);
            cub::DoubleBuffer<uint32_t> v(nullptr, nullptr);
            cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, k, v, (long long)n, 0, 64, stream);
            cub::DeviceScan::ExclusiveSum(nullptr, t, (uint32_t *)nullptr, (uint32_t *)nullptr, (long long)n, stream);
            if (t > sort_tmp) sort_tmp = t;
        }
        unsigned long long *ka, *kb;
        uint32_t *va, *vb, *flag, *idx;
        void *tmp1;
        Arena a1;
        for (int pass = 0; pass < 2; ++pass) {
            ka = a1.take<unsigned long long>(n); kb = a1.take<unsigned long long>(n);
            va = a1.take<uint32_t>(n); vb = a1.take<uint32_t>(n);
            flag = a1.take<uint32_t>(n + 1); idx = a1.take<uint32_t>(n + 1);
            tmp1 = a1.take<uint8_t>(sort_tmp);
            if (pass == 0) {
                void *base = nullptr;
                int rc = dyn_ctx_scratch(ctx, 17, a1.used + 256, &base);
                if (rc) return rc;
                a1 = Arena(base);
            }
        }
        {
            const long long warps_needed = (n_items + 31) / 32;
            const long long grid = std::min<long long>((long long)ctx->sm_count * 16, (warps_needed * 32 + kBlock - 1) / kBlock);
            expand_kernel<<<(unsigned)std::max<long long>(grid, 1), kBlock, 0, stream>>>(records, d_item_rec16, group_off, n_groups,
                                                                                        n_items, tuple_off, win, ka, va, d_out_status);
        }
        cub::DoubleBuffer<unsigned long long> k(ka, kb);
        cub::DoubleBuffer<uint32_t> v(va, vb);
        t = sort_tmp;
        // bits in use: 32 position bits + the group id; the sentinel key is (n_groups << 32 | ~0), one past the last group
        int key_bits = 33;
        while (key_bits < 64 && ((unsigned long long)n_groups >> (key_bits - 32))) ++key_bits;
        DYN_CUDA(cub::DeviceRadixSort::SortPairs(tmp1, t, k, v, (long long)n, 0, key_bits, stream));
        DYN_CUDA(cudaMemsetAsync(flag + n, 0, 4, stream));
        head_kernel<<<grid_for((long long)n), kBlock, 0, stream>>>(k.Current(), (long long)n, n_groups, flag);
        t = sort_tmp;
        DYN_CUDA(cub::DeviceScan::ExclusiveSum(tmp1, t, flag, idx, (long long)n + 1, stream));
        // cells overwrite the alternate buffers (n_cells <= n)
        cell_key = k.Alternate();
        cell = v.Alternate();
        reduce_kernel<<<grid_for((long long)n), kBlock, 0, stream>>>(k.Current(), v.Current(), flag, idx, (long long)n, quality,
                                                                     cell_key, cell);
        // n_cells = idx[n] (uint32) -> int64 slot
        DYN_CUDA(cudaMemcpyAsync(n_cells, idx + n, 4, cudaMemcpyDeviceToDevice, stream));
    }

    // ---- assemble: sizes, offsets, records
    assemble_kernel<false><<<grid_for(n_groups), kBlock, 0, stream>>>(cell_key, cell, n_cells, cell_d, win, cell_start,
                                                                     cell_cnt, n_groups, g_left, g_right, g_ref, sizes,
                                                                     out_cap16, md_cap, nullptr, nullptr, 0, nullptr, nullptr, 0,
                                                                     d_out_nm, d_out_status);
    DYN_CUDA(cudaMemsetAsync(out_cap16 + n_groups, 0, 4, stream));
    DYN_CUDA(cudaMemsetAsync(md_cap + n_groups, 0, 4, stream));
    t = scan_tmp;
    DYN_CUDA(cub::DeviceScan::ExclusiveSum(tmp0, t, out_cap16, d_out_off, n_groups + 1, stream));
    t = scan_tmp;
    DYN_CUDA(cub::DeviceScan::ExclusiveSum(tmp0, t, md_cap, d_out_md_off, n_groups + 1, stream));
    assemble_kernel<true><<<grid_for(n_groups), kBlock, 0, stream>>>(cell_key, cell, n_cells, cell_d, win, cell_start,
                                                                    cell_cnt, n_groups, g_left, g_right, g_ref, sizes,
                                                                    out_cap16, md_cap, d_out_off, static_cast<uint8_t *>(d_out_records),
                                                                    out_capacity_bytes >> 4, d_out_md_off,
                                                                    static_cast<uint8_t *>(d_out_md), md_capacity_bytes, d_out_nm,
                                                                    d_out_status);
    DYN_CUDA(cudaGetLastError());
    return DYN_OK;
}
