// Sequential CIGAR/MD walker over one packed read record (include/dynast_b200.h), shared by the tally's
// slow path (paired units, event overflow, oversized records) and the consensus expansion.
// Restates pysam's get_aligned_pairs(with_seq=True) as used by dynast/preprocessing/bam.py:386-411 and
// consensus.py:73-97: M/=/X pairs with the reference base taken from the record's MD tokens, deleted reference
// bases from the deleted-base tokens, N skips advancing the reference only, I/S advancing the query only.
#pragma once
#include "common.cuh"

namespace {

// BAM 4-bit code of an upper/lower-case IUPAC letter, indexed by (letter & 31), 4 bits per entry:
//   @ A B C D E F G H I J K L M N O | P Q R S T U V W X Y Z
//   0 1 14 2 13 0 0 4 11 0 0 12 0 3 15 0 | 0 0 5 6 8 0 7 9 0 10 0 ...
__device__ __forceinline__ uint32_t letter_code(uint32_t ch) {
    const unsigned long long lo = (0ull) | (1ull << 4) | (14ull << 8) | (2ull << 12) | (13ull << 16) | (0ull << 20) |
                                  (0ull << 24) | (4ull << 28) | (11ull << 32) | (0ull << 36) | (0ull << 40) |
                                  (12ull << 44) | (0ull << 48) | (3ull << 52) | (15ull << 56) | (0ull << 60);
    const unsigned long long hi = (0ull) | (0ull << 4) | (5ull << 8) | (6ull << 12) | (8ull << 16) | (0ull << 20) |
                                  (7ull << 24) | (9ull << 28) | (0ull << 32) | (10ull << 36);
    const uint32_t i = ch & 31;
    const unsigned long long t = (i & 16) ? hi : lo;
    return (uint32_t)(t >> ((i & 15) << 2)) & 0xf;
}

// one-hot BAM base code -> 0..3 (A C G T), anything else -> -1
__device__ __forceinline__ int base_idx(uint32_t code) { return __popc(code) == 1 ? __ffs((int)code) - 1 : -1; }

__device__ __forceinline__ size_t pad4(size_t x) { return (x + 3) & ~(size_t)3; }

// bytes of a record before padding; a lean record (DYN_REC_LEAN) has no qual[] stream
__device__ __forceinline__ size_t record_need(uint32_t l_seq, uint32_t n_cigar, uint32_t n_tok, bool lean = false) {
    return 16 + 4 * (size_t)n_cigar + 4 * (size_t)n_tok + pad4(((size_t)l_seq + 1) >> 1) + (lean ? 0 : l_seq);
}

struct Walker {
    const uint32_t *cig, *tok;
    const uint8_t *seq, *qual;
    int n_cigar, l_seq, n_tok, ref_id;
    int ci, op_left, qpos, rpos, ti, aligned;
    int del[4];
    int del_left;      // deleted reference bases still to be reported (next_event only)
    bool bad, lean;
    // the sequence / quality word and the MD token under the cursor are kept in registers: one load per 8 bases, per 4
    // qualities and per token instead of three loads per base
    uint32_t seq_w, qual_w, tok_w;
    int seq_wi, qual_wi, tok_i;
    // current aligned pair
    int cq, cr;
    uint32_t gn, rn, q;

    // returns padded record size, 0 if the record does not fit `avail`
    __device__ __forceinline__ size_t init(const uint8_t *rec, size_t avail) {
        bad = lean = false;
        ci = op_left = qpos = ti = aligned = 0;
        del_left = 0;
        del[0] = del[1] = del[2] = del[3] = 0;
        seq_w = qual_w = tok_w = 0; seq_wi = qual_wi = tok_i = -1;
        n_cigar = l_seq = n_tok = 0;
        rpos = 0; ref_id = 0;
        cig = tok = nullptr; seq = qual = nullptr;
        if (avail < 16) { bad = true; return 0; }
        const uint4 h = *reinterpret_cast<const uint4 *>(rec);
        const uint32_t ls = h.z & 0xffff, nc = h.z >> 16, nt = h.w & 0xffff;
        lean = ((h.w >> 16) & DYN_REC_LEAN) != 0;
        const size_t need = record_need(ls, nc, nt, lean);
        if (need > avail) { bad = true; return 0; }
        if ((h.w >> 16) & DYN_REC_BAD_MD) bad = true;     // MD disagrees with the CIGAR (pysam raises)
        rpos = (int)h.x; ref_id = (int)h.y;
        l_seq = (int)ls; n_cigar = (int)nc; n_tok = (int)nt;
        cig = reinterpret_cast<const uint32_t *>(rec + 16);
        tok = cig + nc;
        seq = reinterpret_cast<const uint8_t *>(tok + nt);
        qual = seq + pad4(((size_t)ls + 1) >> 1);
        return (need + 15) & ~(size_t)15;
    }
    // one deleted-base token; false if the next token is not one
    __device__ __forceinline__ bool take_del(uint32_t &code) {
        if (ti >= n_tok || !DYN_TOK_IS_DEL(tok[ti])) { bad = true; return false; }
        code = DYN_TOK_BASE(tok[ti++]);
        return true;
    }
    // advance to the next aligned (M/=/X) pair
    __device__ __forceinline__ bool next() {
        for (;;) {
            if (bad) return false;
            if (op_left == 0) {
                if (ci >= n_cigar) { if (ti != n_tok) bad = true; return false; }
                const uint32_t c = cig[ci++];
                const int op = (int)(c & 0xf), len = (int)(c >> 4);
                if (op == 0 || op == 7 || op == 8) { op_left = len; continue; }
                if (op == 2) {
                    for (int j = 0; j < len; ++j) {
                        uint32_t code;
                        if (!take_del(code)) return false;
                        const int b = base_idx(code);
                        if (b >= 0) ++del[b];
                    }
                    rpos += len;
                } else if (op == 3) rpos += len;
                else if (op == 1 || op == 4) { qpos += len; if (qpos > l_seq) { bad = true; return false; } }
                continue;
            }
            if (qpos >= l_seq) { bad = true; return false; }
            if ((qpos >> 3) != seq_wi) { seq_wi = qpos >> 3; seq_w = reinterpret_cast<const uint32_t *>(seq)[seq_wi]; }
            rn = (seq_w >> ((((qpos >> 1) & 3) << 3) + ((~qpos & 1) << 2))) & 0xf;
            gn = rn;
            uint32_t tq = 0;         // lean records: the Phred rides in the MD token (mismatching bases only)
            if (ti < n_tok) {
                if (tok_i != ti) { tok_i = ti; tok_w = tok[ti]; }
                if (!DYN_TOK_IS_DEL(tok_w) && (int)DYN_TOK_AOFF(tok_w) == aligned) { gn = DYN_TOK_BASE(tok_w); tq = DYN_TOK_QUAL(tok_w); ++ti; }
            }
            if (lean) q = tq;
            else {
                if ((qpos >> 2) != qual_wi) { qual_wi = qpos >> 2; qual_w = reinterpret_cast<const uint32_t *>(qual)[qual_wi]; }
                q = (qual_w >> ((qpos & 3) << 3)) & 0xffu;
            }
            cq = qpos; cr = rpos;
            ++qpos; ++rpos; ++aligned; --op_left;
            return true;
        }
    }

    // Like next(), but deleted reference bases are reported too (consensus.py:73-97 walks
    // get_aligned_pairs(matches_only=False)): returns 0 at the end, 1 for an aligned pair (cq, cr, gn, rn, q),
    // 2 for a deleted reference base (cr, gn).
    __device__ int next_event() {
        for (;;) {
            if (bad) return 0;
            if (del_left > 0) {
                uint32_t code;
                if (!take_del(code)) return 0;
                gn = code;
                cr = rpos++;
                --del_left;
                return 2;
            }
            if (op_left == 0) {
                if (ci >= n_cigar) { if (ti != n_tok) bad = true; return 0; }
                const uint32_t c = cig[ci++];
                const int op = (int)(c & 0xf), len = (int)(c >> 4);
                if (op == 0 || op == 7 || op == 8) op_left = len;
                else if (op == 2) del_left = len;
                else if (op == 3) rpos += len;
                else if (op == 1 || op == 4) { qpos += len; if (qpos > l_seq) { bad = true; return 0; } }
                continue;
            }
            return next() ? 1 : 0;
        }
    }
};

}  // namespace
