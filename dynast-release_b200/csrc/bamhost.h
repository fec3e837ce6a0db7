// Host-side BAM helpers shared by the whole-file packer (bampack.cpp), the streaming reader (bamstream.cpp) and the
// BAM writer (bamwrite.cpp): little-endian reads, aux-tag scan, MD tokenisation, the packed-record writer.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <string>
#include <vector>

#include "../../include/dynast_b200.h"

void dyn_set_error(const char *fmt, ...);

namespace bamhost {

struct PhaseTimer {     // DYN_BAM_PACK_TIMING=1: phase times on stderr
    bool on;
    std::chrono::steady_clock::time_point t0;
    PhaseTimer() : on(getenv("DYN_BAM_PACK_TIMING") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void mark(const char *what) {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "dyn_bam_pack: %-10s %8.3f s\n", what, std::chrono::duration<double>(t1 - t0).count());
        t0 = t1;
    }
};

struct Block { size_t in_off, in_len, out_off, out_len; };

struct Pool {              // string pool with offsets
    std::vector<char> data;
    std::vector<uint32_t> off;   // n + 1
    Pool() { off.push_back(0); }
    void add(const char *s, size_t n) { data.insert(data.end(), s, s + n); off.push_back((uint32_t)data.size()); }
};

struct Aux {               // the tags of one record that the path needs
    const char *md = nullptr; size_t md_len = 0;
    const char *bc = nullptr; size_t bc_len = 0;
    const char *umi = nullptr; size_t umi_len = 0;
    const char *gx = nullptr; size_t gx_len = 0;
    bool has_hi = false, has_as = false, has_bc = false, has_umi = false, has_gx = false, has_md = false;
    int64_t hi = 0, as = 0;
};

inline uint16_t rd16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

inline bool parse_aux(const uint8_t *p, const uint8_t *end, const char *bc_tag, const char *umi_tag, const char *gx_tag, Aux &a) {
    while (p + 3 <= end) {
        const char t0 = (char)p[0], t1 = (char)p[1], ty = (char)p[2];
        p += 3;
        int64_t ival = 0;
        bool is_int = false;
        const uint8_t *sval = nullptr;
        size_t slen = 0;
        switch (ty) {
            case 'A': if (p + 1 > end) return false; sval = p; slen = 1; p += 1; break;
            case 'c': if (p + 1 > end) return false; ival = (int8_t)p[0]; is_int = true; p += 1; break;
            case 'C': if (p + 1 > end) return false; ival = p[0]; is_int = true; p += 1; break;
            case 's': if (p + 2 > end) return false; ival = (int16_t)rd16(p); is_int = true; p += 2; break;
            case 'S': if (p + 2 > end) return false; ival = rd16(p); is_int = true; p += 2; break;
            case 'i': if (p + 4 > end) return false; ival = (int32_t)rd32(p); is_int = true; p += 4; break;
            case 'I': if (p + 4 > end) return false; ival = rd32(p); is_int = true; p += 4; break;
            case 'f': if (p + 4 > end) return false; p += 4; break;
            case 'Z': case 'H': {
                const uint8_t *q = (const uint8_t *)memchr(p, 0, (size_t)(end - p));
                if (!q) return false;
                sval = p; slen = (size_t)(q - p); p = q + 1;
                break;
            }
            case 'B': {
                if (p + 5 > end) return false;
                const char sub = (char)p[0];
                const uint32_t cnt = rd32(p + 1);
                const size_t w = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
                p += 5 + (size_t)cnt * w;
                if (p > end) return false;
                break;
            }
            default: return false;
        }
        auto is = [&](const char *tag) { return tag && tag[0] == t0 && tag[1] == t1; };
        if (t0 == 'M' && t1 == 'D' && sval) { a.md = (const char *)sval; a.md_len = slen; a.has_md = true; }
        if (t0 == 'H' && t1 == 'I' && is_int) { a.hi = ival; a.has_hi = true; }
        if (t0 == 'A' && t1 == 'S' && is_int) { a.as = ival; a.has_as = true; }
        if (is(bc_tag) && sval) { a.bc = (const char *)sval; a.bc_len = slen; a.has_bc = true; }
        if (is(umi_tag) && sval) { a.umi = (const char *)sval; a.umi_len = slen; a.has_umi = true; }
        if (is(gx_tag) && sval) { a.gx = (const char *)sval; a.gx_len = slen; a.has_gx = true; }
    }
    return true;
}

struct Pending { size_t rec_pos; };   // offset of the parked mate's BAM record in the inflated stream

}  // namespace bamhost


struct dyn_bam_result {
    uint8_t *blob = nullptr;        // packed records (page-locked if possible)
    size_t blob_bytes = 0;
    bool blob_pinned = false;
    std::vector<uint32_t> rec_off;  // n_units + 1 (16-byte units)
    std::vector<int32_t> ref_id, pos, hi, score;
    std::vector<uint16_t> flag;
    std::vector<uint8_t> gx_assigned, paired;
    bamhost::Pool name, barcode, umi, gene;
    std::string header_text;
    bamhost::Pool ref_name;
    std::vector<int32_t> ref_len;
    int64_t n_records_seen = 0;
    std::vector<int64_t> rec_index;     // ordinal of the unit's (second-seen) record among all records of the file
    bamhost::Pool raw, raw_mate;        // DYN_BAM_KEEP_RAW: the units' BAM alignment records as they are in the file
    void *extra = nullptr;          // stream chunks: dyn_bam_chunk_extra (bamstream.cpp)
};

void dyn_bam_chunk_extra_free(dyn_bam_result *r);                                   // bamstream.cpp
const void *dyn_bam_chunk_field(const dyn_bam_result *r, int field, size_t *nb);    // bamstream.cpp


namespace bamhost {

// BAM 4-bit code of an MD letter
inline uint32_t letter_code(char ch) {
    static const char codes[] = "=ACMGRSVTWYHKDBN";
    if (ch >= 'a' && ch <= 'z') ch = (char)(ch - 32);
    for (uint32_t i = 0; i < 16; ++i)
        if (codes[i] == ch) return i;
    return 0;
}

// MD text -> tokens (include/dynast_b200.h); returns false when the tag disagrees with the CIGAR
inline bool tokenize_md(const char *md, size_t md_len, const uint8_t *cigar, uint32_t n_cigar, std::vector<uint32_t> &toks) {
    toks.clear();
    uint64_t aoff = 0, val = 0, n_del = 0;
    bool in_del = false, ok = true;
    // The letters behind a '^' that are deleted bases are as many as the CIGAR's next D operation has -- zero counts between
    // them are skipped, letters behind them are mismatches.  dynast's consensus (consensus.py:116-176) appends a count only
    // when it is non-zero or a mismatch came last, and never forgets that mismatch inside a deletion: two deleted bases and
    // a mismatch come out as "148^TAA" (no "0"), a mismatch and three deleted bases as "56C0^G0T0C" -- and a count on its
    // consensus BAM has to read both.
    uint32_t d_cursor = 0, del_left = 0;
    auto next_deletion = [&]() -> uint32_t {
        while (d_cursor < n_cigar) {
            const uint32_t w = rd32(cigar + 4 * (size_t)d_cursor++);
            if ((w & 0xf) == 2) return w >> 4;
        }
        return 0;
    };
    for (size_t i = 0; i < md_len; ++i) {
        const char ch = md[i];
        if (ch >= '0' && ch <= '9') { val = val * 10 + (uint64_t)(ch - '0'); in_del = false; }
        else if (ch == '^') { aoff += val; val = 0; in_del = true; del_left = next_deletion(); }
        else if (del_left > 0) {
            if (val) ok = false;              // matched bases inside a deletion
            val = 0;
            toks.push_back((uint32_t)std::min<uint64_t>(aoff, 0xffff) | (letter_code(ch) << 16) | (1u << 20));
            ++n_del; --del_left;
        }
        else {
            aoff += val; val = 0;
            if (aoff > 0xffff) ok = false;
            toks.push_back((uint32_t)(aoff & 0xffff) | (letter_code(ch) << 16));
            ++aoff;
        }
    }
    aoff += val;
    uint64_t m_total = 0, d_total = 0;
    for (uint32_t c = 0; c < n_cigar; ++c) {
        const uint32_t w = rd32(cigar + 4 * (size_t)c), op = w & 0xf, len = w >> 4;
        if (op == 0 || op == 7 || op == 8) m_total += len;
        else if (op == 2) d_total += len;
    }
    if (aoff != m_total || n_del != d_total) ok = false;
    if (toks.size() > 0xffff) { toks.clear(); ok = false; }
    return ok;
}

// Phred of the query base at every non-deleted token's aligned offset goes into bits 21..28 (lean layout).  Tokens
// and CIGAR both ascend, so one pass over the CIGAR serves all tokens.
inline void fold_token_quals(std::vector<uint32_t> &toks, const uint8_t *cigar, uint32_t n_cigar, const uint8_t *qual, uint32_t l_seq) {
    size_t ti = 0;
    uint32_t done = 0, q = 0;
    for (uint32_t c = 0; c < n_cigar && ti < toks.size(); ++c) {
        const uint32_t w = rd32(cigar + 4 * (size_t)c), op = w & 0xf, len = w >> 4;
        if (op == 0 || op == 7 || op == 8) {
            while (ti < toks.size()) {
                const uint32_t tk = toks[ti];
                if (tk & (1u << 20)) { ++ti; continue; }          // deleted base: no query position
                const uint32_t aoff = tk & 0xffff;
                if (aoff >= done + len) break;
                const uint32_t qpos = q + (aoff - done);
                toks[ti] = tk | ((uint32_t)(qpos < l_seq ? qual[qpos] : 0) << 21);
                ++ti;
            }
            done += len; q += len;
        } else if (op == 1 || op == 4) q += len;
    }
}

// rec points at the BAM alignment record body (after block_size); lean: DYN_REC_LEAN layout (no qual stream)
template <typename Out>
inline size_t pack_record(const uint8_t *rec, const Aux &a, uint16_t extra_flag, std::vector<uint32_t> &toks, Out &out,
                          bool lean = false) {
    const int32_t ref_id = (int32_t)rd32(rec), pos = (int32_t)rd32(rec + 4);
    const uint32_t l_read_name = rec[8];
    const uint32_t n_cigar = rd16(rec + 12);
    const uint16_t flag = rd16(rec + 14);
    const uint32_t l_seq = rd32(rec + 16);
    const uint8_t *cigar = rec + 32 + l_read_name;
    const uint8_t *seq = cigar + 4 * (size_t)n_cigar;
    const uint8_t *qual = seq + (l_seq + 1) / 2;
    const bool ok = tokenize_md(a.md, a.md_len, cigar, n_cigar, toks);
    if (lean) fold_token_quals(toks, cigar, n_cigar, qual, l_seq);
    const size_t raw = 16 + 4 * (size_t)n_cigar + 4 * toks.size() + ((((size_t)l_seq + 1) / 2 + 3) & ~(size_t)3) + (lean ? 0 : l_seq);
    const size_t bytes = (raw + 15) & ~(size_t)15;
    const size_t o = out.size();
    out.resize(o + bytes);
    uint8_t *p = out.data() + o;
    memset(p + bytes - 16, 0, 16);            // the padding (and whatever else of the last 16 bytes is not written)
    memcpy(p, &pos, 4); memcpy(p + 4, &ref_id, 4);
    const uint16_t ls = (uint16_t)l_seq, nc = (uint16_t)n_cigar, nt = (uint16_t)toks.size();
    const uint16_t fl = (uint16_t)((flag & 0x0fff) | extra_flag | (ok ? 0 : DYN_REC_BAD_MD) | (lean ? DYN_REC_LEAN : 0));
    memcpy(p + 8, &ls, 2); memcpy(p + 10, &nc, 2); memcpy(p + 12, &nt, 2); memcpy(p + 14, &fl, 2);
    memcpy(p + 16, cigar, 4 * (size_t)n_cigar);
    uint8_t *pt = p + 16 + 4 * (size_t)n_cigar;
    if (!toks.empty()) memcpy(pt, toks.data(), 4 * toks.size());
    uint8_t *ps = pt + 4 * toks.size();
    const size_t seq_pad = (((size_t)l_seq + 1) / 2 + 3) & ~(size_t)3;
    if (seq_pad) memset(ps + seq_pad - 4, 0, 4);
    memcpy(ps, seq, (l_seq + 1) / 2);
    if (l_seq & 1) ps[l_seq / 2] &= 0xf0;     // the spec leaves the low nibble of the last byte undefined
    if (!lean) memcpy(ps + seq_pad, qual, l_seq);
    return bytes;
}

}  // namespace bamhost

