// BAM alignment record -> packed record, one record per thread (device ingest, bamdevice.cu).  Same rules as the host
// packer (bamhost.h: read filter of bam.py:173-174, 242-248, 286-292; MD tokenisation; lean / full layout; 2-bit
// barcode / UMI keys of bamstream.cpp), written as __host__ __device__ code so that tests/test_bamdevice_host.py can
// run the very same functions on the CPU against the host packer's output.
#pragma once
#include <stdint.h>

#include "../../include/dynast_b200.h"

#ifdef __CUDACC__
#define REC_HD __host__ __device__ __forceinline__
#else
#define REC_HD inline
#endif

namespace bamrec {

REC_HD uint32_t ld16(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
REC_HD uint32_t ld32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

struct Tags {          // two-character tag codes (t0 | t1 << 8), 0 = not used
    uint32_t bc, umi, gx;
    int require_gene, lean, full_keys;
};

struct Parsed {        // what the parse pass leaves per record
    uint32_t size16;   // packed size in 16-byte units, 0 = record filtered out
    uint32_t md_off;   // offset of the MD text inside the inflated chunk
    uint32_t n_tok;
    int32_t hi, score;
    uint64_t bc_key, umi_key, gx_hash;
    uint8_t gx_assigned, error;     // error: 1 corrupt record / aux, 2 HI or AS missing, 3 MD missing, 4 too long for the layout
};

// 2-bit key of an A/C/G/T string under a sentinel bit; "-<digits>" suffix in bits 58..63 (bamstream.cpp acgt_key)
REC_HD uint64_t acgt_key(const uint8_t *s, uint32_t n, bool allow_suffix) {
    uint64_t suffix = 0;
    if (allow_suffix) {
        uint32_t d = n;
        while (d > 0 && s[d - 1] >= '0' && s[d - 1] <= '9') --d;
        if (d < n && d > 0 && s[d - 1] == '-') {
            uint64_t v = 0;
            for (uint32_t i = d; i < n; ++i) v = v * 10 + (uint64_t)(s[i] - '0');
            if (v >= 63) return ~0ull;
            suffix = v + 1;
            n = d - 1;
        }
    }
    if (n > (allow_suffix ? 28u : 31u)) return ~0ull;
    uint64_t k = 1;
    for (uint32_t i = 0; i < n; ++i) {
        uint64_t c;
        const uint8_t ch = s[i];
        if (ch == 'A') c = 0; else if (ch == 'C') c = 1; else if (ch == 'G') c = 2; else if (ch == 'T') c = 3; else return ~0ull;
        k = (k << 2) | c;
    }
    return k | (suffix << 58);
}

REC_HD uint64_t fnv1a(const uint8_t *s, uint32_t n) {
    uint64_t h = 1469598103934665603ull;
    for (uint32_t i = 0; i < n; ++i) { h ^= s[i]; h *= 1099511628211ull; }
    return h;
}

// rec points at block_size; base = start of the inflated chunk (offsets are relative to it)
REC_HD void parse_record(const uint8_t *base, uint32_t at, const Tags &tg, Parsed &o) {
    o.size16 = 0; o.md_off = 0; o.n_tok = 0; o.hi = 0; o.score = 0; o.bc_key = 0; o.umi_key = 0; o.gx_hash = 0;
    o.gx_assigned = 0; o.error = 0;
    const uint8_t *p = base + at;
    const uint32_t block_size = ld32(p);
    const uint8_t *rec = p + 4;
    const int32_t ref_id = (int32_t)ld32(rec);
    const uint32_t l_read_name = rec[8], n_cigar = ld16(rec + 12), flag = ld16(rec + 14), l_seq = ld32(rec + 16);
    if (ref_id < 0) return;
    if (flag & (0x100 | 0x4 | 0x400)) return;                 // bam.py:173: secondary, unmapped, duplicate
    const uint64_t fixed = 32ull + l_read_name + 4ull * n_cigar + ((uint64_t)l_seq + 1) / 2 + l_seq;
    if (fixed > block_size) { o.error = 1; return; }
    const uint8_t *a = rec + fixed, *end = rec + block_size;
    bool has_hi = false, has_as = false, has_md = false, has_bc = false, has_umi = false, has_gx = false;
    uint32_t md_len = 0;
    while (a + 3 <= end) {
        const uint32_t tag = (uint32_t)a[0] | ((uint32_t)a[1] << 8);
        const uint8_t ty = a[2];
        a += 3;
        int64_t ival = 0;
        bool is_int = false;
        const uint8_t *sval = nullptr;
        uint32_t slen = 0;
        if (ty == 'A') { if (a + 1 > end) { o.error = 1; return; } sval = a; slen = 1; a += 1; }
        else if (ty == 'c') { if (a + 1 > end) { o.error = 1; return; } ival = (int8_t)a[0]; is_int = true; a += 1; }
        else if (ty == 'C') { if (a + 1 > end) { o.error = 1; return; } ival = a[0]; is_int = true; a += 1; }
        else if (ty == 's') { if (a + 2 > end) { o.error = 1; return; } ival = (int16_t)ld16(a); is_int = true; a += 2; }
        else if (ty == 'S') { if (a + 2 > end) { o.error = 1; return; } ival = ld16(a); is_int = true; a += 2; }
        else if (ty == 'i') { if (a + 4 > end) { o.error = 1; return; } ival = (int32_t)ld32(a); is_int = true; a += 4; }
        else if (ty == 'I') { if (a + 4 > end) { o.error = 1; return; } ival = ld32(a); is_int = true; a += 4; }
        else if (ty == 'f') { if (a + 4 > end) { o.error = 1; return; } a += 4; }
        else if (ty == 'Z' || ty == 'H') {
            const uint8_t *q = a;
            while (q < end && *q) ++q;
            if (q >= end) { o.error = 1; return; }
            sval = a; slen = (uint32_t)(q - a); a = q + 1;
        } else if (ty == 'B') {
            if (a + 5 > end) { o.error = 1; return; }
            const uint8_t sub = a[0];
            const uint32_t cnt = ld32(a + 1);
            const uint32_t w = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
            if ((uint64_t)cnt * w > (uint64_t)(end - a)) { o.error = 1; return; }
            a += 5 + (size_t)cnt * w;
            if (a > end) { o.error = 1; return; }
        } else { o.error = 1; return; }
        if (tag == ('M' | ('D' << 8)) && sval) { o.md_off = (uint32_t)(sval - base); md_len = slen; has_md = true; }
        if (tag == ('H' | ('I' << 8)) && is_int) { o.hi = (int32_t)ival; has_hi = true; }
        if (tag == ('A' | ('S' << 8)) && is_int) { o.score = (int32_t)ival; has_as = true; }
        if (tg.bc && tag == tg.bc && sval) { o.bc_key = acgt_key(sval, slen, true); has_bc = true; if (slen == 1 && sval[0] == '-') o.bc_key = 1; }
        if (tg.umi && tag == tg.umi && sval) { o.umi_key = acgt_key(sval, slen, false); has_umi = true; }
        if (tg.gx && tag == tg.gx && sval) { o.gx_hash = fnv1a(sval, slen); has_gx = true; }
    }
    if ((tg.bc && !has_bc) || (tg.umi && !has_umi) || (tg.require_gene && tg.gx && !has_gx)) return;       // bam.py:174, 242-248
    if (tg.bc && o.bc_key == 1) return;                                                                        // barcode '-' (bam.py:290)
    if (!has_hi || !has_as) { o.error = 2; return; }
    if (!has_md) { o.error = 3; return; }
    if (l_seq > 0xffff || n_cigar > 0xffff) { o.error = 4; return; }
    o.gx_assigned = has_gx ? 1 : 0;
    // MD letters = tokens
    uint32_t n_tok = 0;
    const uint8_t *md = base + o.md_off;
    for (uint32_t i = 0; i < md_len; ++i) {
        const uint8_t ch = md[i];
        if (!(ch >= '0' && ch <= '9') && ch != '^') ++n_tok;
    }
    if (n_tok > 0xffff) n_tok = 0;        // the host packer drops the tokens and flags the record (DYN_REC_BAD_MD)
    o.n_tok = n_tok;
    const uint64_t raw = 16ull + 4ull * n_cigar + 4ull * n_tok + ((((uint64_t)l_seq + 1) / 2 + 3) & ~3ull) + (tg.lean ? 0 : l_seq);
    o.size16 = (uint32_t)((raw + 15) >> 4);
}

REC_HD uint32_t letter_code(uint8_t ch) {
    if (ch >= 'a' && ch <= 'z') ch = (uint8_t)(ch - 32);
    switch (ch) {
        case '=': return 0; case 'A': return 1; case 'C': return 2; case 'M': return 3; case 'G': return 4; case 'R': return 5;
        case 'S': return 6; case 'V': return 7; case 'T': return 8; case 'W': return 9; case 'Y': return 10; case 'H': return 11;
        case 'K': return 12; case 'D': return 13; case 'B': return 14; case 'N': return 15; default: return 0;
    }
}

// Writes the packed record of the BAM record at base + at to `out` (size16 * 16 bytes, as parse_record computed).
REC_HD void pack_record(const uint8_t *base, uint32_t at, const Parsed &pr, int lean, uint8_t *out) {
    const uint8_t *rec = base + at + 4;
    const uint32_t l_read_name = rec[8], n_cigar = ld16(rec + 12), flag = ld16(rec + 14), l_seq = ld32(rec + 16);
    const uint8_t *cigar = rec + 32 + l_read_name;
    const uint8_t *seq = cigar + 4 * (size_t)n_cigar;
    const uint8_t *qual = seq + (l_seq + 1) / 2;
    const uint32_t bytes = pr.size16 << 4;
    uint32_t *o32 = reinterpret_cast<uint32_t *>(out);
    for (uint32_t i = bytes / 4 - 4; i < bytes / 4; ++i) o32[i] = 0;          // padding (and the tail of whatever ends there)
    // CIGAR words + totals
    uint32_t m_total = 0, d_total = 0;
    uint32_t *oc = o32 + 4;
    for (uint32_t c = 0; c < n_cigar; ++c) {
        const uint32_t w = ld32(cigar + 4 * (size_t)c), op = w & 0xf, len = w >> 4;
        oc[c] = w;
        if (op == 0 || op == 7 || op == 8) m_total += len;
        else if (op == 2) d_total += len;
    }
    // MD text -> tokens (bamhost.h tokenize_md), with the Phred folded in for the lean layout (fold_token_quals)
    uint32_t *ot = oc + n_cigar;
    const uint8_t *md = base + pr.md_off;
    uint64_t aoff = 0, val = 0, n_del = 0;
    bool in_del = false, ok = true;
    uint32_t nt = 0;
    // cursor over the CIGAR for aligned offset -> query position
    uint32_t ci = 0, done = 0, q = 0, cur_len = 0;
    bool in_m = false;
    auto qual_at = [&](uint32_t a) -> uint32_t {
        for (;;) {
            if (in_m && a < done + cur_len) { const uint32_t qpos = q + (a - done); return qpos < l_seq ? qual[qpos] : 0u; }
            if (in_m) { done += cur_len; q += cur_len; in_m = false; }
            if (ci >= n_cigar) return 0u;
            const uint32_t w = ld32(cigar + 4 * (size_t)ci), op = w & 0xf, len = w >> 4;
            ++ci;
            if (op == 0 || op == 7 || op == 8) { in_m = true; cur_len = len; }
            else if (op == 1 || op == 4) q += len;
        }
    };
    const bool keep_tokens = pr.n_tok > 0;
    // behind a '^', as many letters as the CIGAR's next D operation has are deleted bases, zero counts between them skipped;
    // letters behind them are mismatches (bamhost.h tokenize_md)
    uint32_t d_cursor = 0, del_left = 0;
    for (uint32_t i = 0; md[i]; ++i) {
        const uint8_t ch = md[i];
        if (ch >= '0' && ch <= '9') { val = val * 10 + (uint64_t)(ch - '0'); in_del = false; }
        else if (ch == '^') {
            aoff += val; val = 0; in_del = true; del_left = 0;
            while (d_cursor < n_cigar) {
                const uint32_t w = ld32(cigar + 4 * (size_t)d_cursor++);
                if ((w & 0xf) == 2) { del_left = w >> 4; break; }
            }
        }
        else if (del_left > 0) {
            if (val) ok = false;
            val = 0;
            if (keep_tokens && nt < pr.n_tok) ot[nt] = (uint32_t)(aoff < 0xffff ? aoff : 0xffff) | (letter_code(ch) << 16) | (1u << 20);
            ++nt; ++n_del; --del_left;
        } else {
            aoff += val; val = 0;
            if (aoff > 0xffff) ok = false;
            if (keep_tokens && nt < pr.n_tok) {
                uint32_t tk = (uint32_t)(aoff & 0xffff) | (letter_code(ch) << 16);
                if (lean) tk |= qual_at((uint32_t)(aoff & 0xffff)) << 21;
                ot[nt] = tk;
            }
            ++nt; ++aoff;
        }
    }
    aoff += val;
    if (aoff != m_total || n_del != d_total) ok = false;
    if (!keep_tokens && nt > 0) ok = false;            // more than 65 535 tokens: dropped
    // header
    o32[0] = ld32(rec + 4);          // pos
    o32[1] = ld32(rec);              // ref_id
    o32[2] = l_seq | (n_cigar << 16);
    o32[3] = pr.n_tok | (((flag & 0x0fff) | (ok ? 0u : DYN_REC_BAD_MD) | (lean ? DYN_REC_LEAN : 0u)) << 16);
    // sequence (4-bit codes as they are), zero padded to 4 bytes; the spec leaves the low nibble of the last byte undefined
    uint8_t *os = reinterpret_cast<uint8_t *>(ot + pr.n_tok);
    const uint32_t sb = (l_seq + 1) / 2, sp = (sb + 3) & ~3u;
    for (uint32_t i = 0; i < sb; ++i) os[i] = seq[i];
    if (l_seq & 1) os[sb - 1] &= 0xf0;
    for (uint32_t i = sb; i < sp; ++i) os[i] = 0;
    if (!lean) {
        uint8_t *oq = os + sp;
        for (uint32_t i = 0; i < l_seq; ++i) oq[i] = qual[i];
    }
}

}  // namespace bamrec
