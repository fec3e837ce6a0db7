// DEFLATE (RFC 1951) decoder in two passes, the Huffman pass with ONE LANE PER BGZF BLOCK:
//   pass 1 (this file)  every thread of a warp walks the Huffman stream of its own block and writes TOKENS -- up to three
//                       literal bytes, or (length, distance) of a match -- to a scratch stream; an issued instruction
//                       advances 32 streams instead of one, and nothing in the pass reads the output it stands for, so
//                       a round costs a shared-memory lookup, not an L2 round trip;
//   pass 2 (lz_kernel)  one warp per block turns 32 tokens at a time into bytes: positions by a prefix sum of the token
//                       lengths, literals and independent matches lane-parallel, dependent matches in stream order.
// (inflate_core.cuh, one warp per block with lane 0 decoding, spent a whole warp's issue slot on every step of a serial
// Huffman walk: 25 warp-instructions per output byte at 2.4 active lanes, profiles/r2_inflate_ncu_summary.txt.  A
// one-pass lane-per-block decoder that copied matches itself was tried first and was SLOWER than that -- 70 vs 30 ms per
// GiB slab: every match waits an L2 round trip for bytes the lane has just written, and 32 lanes wait for the slowest.)
//
// Pass 1 is a state machine advanced one ROUND at a time -- one Huffman symbol, up to kCopy bytes of a stored run, or
// one block header -- so that the lanes of a warp reconverge after every round whatever state each of them is in.
// Tables are per lane, 16 bits wide and interleaved across the lanes in shared memory (element i of lane l at
// [i * 32 + l]: a warp-wide lookup touches 16 banks twice, never one bank 32 times):
//   lit_lut  [512]  9-bit lookup: symbol (9 bits) | code length << 9; 0 = a longer code, resolved by the canonical walk
//   dist_lut [128]  7-bit lookup, same entry format; also holds the code-length code while a dynamic
//                   header is read
//   sorted symbols + per-length counts of both codes for the canonical walk (codes longer than the lookup width)
// 1 984 bytes per lane, 62 KB per warp: three warps per SM.  Code lengths of a dynamic header are staged inside lit_lut
// (320 entries) until both sorted lists exist.  Length / distance bases are arithmetic (no table lookups that diverge).
// Token: bit 31 set: match, distance << 9 | length;  bit 31 clear: (n - 1) << 29 | up to three literal bytes, first lowest.
// Every token stands for at least one output byte, so a block's tokens fit out_len words.
// The same source compiles for the host with STRIDE 1 (tests/native/ingest_host.cpp, checked against zlib).
#pragma once
#include <stdint.h>

#include "inflate_core.cuh"

namespace dinflate {
namespace lane {

#ifndef DYN_INF_LIT_BITS
#define DYN_INF_LIT_BITS 9
#endif
#ifndef DYN_INF_DIST_BITS
#define DYN_INF_DIST_BITS 7
#endif
constexpr int kLitBits = DYN_INF_LIT_BITS, kDistBits = DYN_INF_DIST_BITS, kCopy = 8;
static_assert(kLitBits >= 6 && kLitBits <= 12 && kDistBits >= 5 && kDistBits <= 10, "lookup widths");
constexpr int kLitLut = 0, kDistLut = 1 << kLitBits, kLitSorted = kDistLut + (1 << kDistBits), kDistSorted = kLitSorted + 288,
              kLitCount = kDistSorted + 32, kDistCount = kLitCount + 16, kUnits = kDistCount + 16;
// where the canonical walk stands after the lookup width: `index` in count[0] (symbols without a code: nothing else reads
// it), `first` here -- [0] literal / length code, [1] distance code, [2] code-length code
constexpr int kWalkFirst = kUnits;
// While a dynamic header is read, its 320 code lengths and the 7-bit lookup of the code-length code need a home: inside the
// two lookup tables when those are wide enough (they are built afterwards), else in a scratch region of their own
constexpr bool kOverlay = kLitBits >= 9 && kDistBits >= 7;
constexpr int kLens = kOverlay ? kLitLut : kWalkFirst + 4, kClLut = kOverlay ? kDistLut : kLens + 320;
constexpr int kUnitsAll = kOverlay ? kWalkFirst + 4 : kClLut + 128;

template <int STRIDE>
struct Mem {
    uint16_t *base;
    INF_HD uint32_t get(int i) const { return base[i * STRIDE]; }
    INF_HD void set(int i, uint32_t v) const { base[i * STRIDE] = (uint16_t)v; }
};

enum : int { ST_HEADER = 0, ST_DECODE = 1, ST_DIST = 2, ST_STORED = 3, ST_END = 4 };

struct Lane {
    const uint8_t *ip, *in_end;       // the aligned 4-byte word held in `nw`; end of the block's payload
    uint32_t nw;                      // the next input word, loaded one refill ahead of its use (the load's latency -- 19 % of
                                      // the stall samples when refill() loaded the word it needed -- hides behind a round)
    uint64_t bb;                      // bit buffer, bc valid bits
    int bc;
    uint32_t *tok;                    // the block's token stream
    uint32_t nt;
    uint32_t pend;                    // literal bytes not yet in a token
    int npend;
    uint32_t op, out_len;             // output position the tokens so far stand for
    uint32_t left;                    // bytes left of the current stored run
    int state, err;
    bool last;
};

INF_HD void flush_literals(Lane &s) {
    if (s.npend) { s.tok[s.nt++] = ((uint32_t)(s.npend - 1) << 29) | s.pend; s.pend = 0; s.npend = 0; }
}
INF_HD void literal(Lane &s, uint32_t byte) {
    s.pend |= byte << (8 * s.npend);
    ++s.op;
    if (++s.npend == 3) flush_literals(s);
}

INF_HD uint32_t rev_bits32(uint32_t v) {
#ifdef __CUDA_ARCH__
    return __brev(v);
#else
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0f0f0f0fu) | ((v & 0x0f0f0f0fu) << 4);
    v = ((v >> 8) & 0x00ff00ffu) | ((v & 0x00ff00ffu) << 8);
    return (v >> 16) | (v << 16);
#endif
}
INF_HD uint32_t ld32(const uint8_t *p) { return *reinterpret_cast<const uint32_t *>(p); }

// The input is read in aligned words; up to 11 bytes behind the payload may be touched (the buffers are padded by 16).
INF_HD void lane_init(Lane &s, const uint8_t *in, uint32_t in_len, uint32_t *tok, uint32_t out_len) {
    const int mis = (int)((uintptr_t)in & 3);
    const uint8_t *w = in - mis;
    s.bb = (uint64_t)(ld32(w) >> (8 * mis));
    s.bc = 32 - 8 * mis;
    s.ip = w + 4;
    s.in_end = in + in_len;
    s.nw = (s.ip < s.in_end + 8) ? ld32(s.ip) : 0u;
    s.tok = tok; s.nt = 0; s.pend = 0; s.npend = 0;
    s.op = 0; s.out_len = out_len;
    s.left = 0;
    s.err = OK; s.last = false;
    s.state = out_len ? ST_HEADER : ST_END;
}

// at least 33 valid bits afterwards: a literal / length code with its extra bits needs 20, a distance 28
INF_HD void refill(Lane &s) {
    if (s.bc <= 32) {
        s.bb |= (uint64_t)s.nw << s.bc;
        s.bc += 32;
        s.ip += 4;
        s.nw = (s.ip < s.in_end + 8) ? ld32(s.ip) : 0u;
    }
}
INF_HD uint32_t peek(const Lane &s, int n) { return (uint32_t)s.bb & ((1u << n) - 1u); }
INF_HD void drop(Lane &s, int n) { s.bb >>= n; s.bc -= n; }
INF_HD uint32_t take(Lane &s, int n) { const uint32_t v = peek(s, n); drop(s, n); return v; }
INF_HD bool overrun(const Lane &s) { return (long long)(s.ip - s.in_end) * 8 > (long long)s.bc; }      // s.ip: end of what bb holds
INF_HD void fail(Lane &s, int code) { s.err = code; s.state = ST_END; }

// Canonical Huffman tables from the code lengths at m[lens_at ..): count / sorted for the walk, then the lookup table
// (which may overlay the lengths: they are dead by then).
template <int STRIDE>
INF_HD bool build(const Mem<STRIDE> &m, int lens_at, int n_sym, int lut, int bits, int count, int sorted, int walk_first) {
    for (int i = 0; i < 16; ++i) m.set(count + i, 0);
    for (int s = 0; s < n_sym; ++s) { const int l = (int)m.get(lens_at + s); m.set(count + l, m.get(count + l) + 1); }
    int left = 1;
    uint32_t offs = 0;
    // offsets of each length's symbols in `sorted`, kept in the table itself while the symbols are placed: count[l]
    // is rebuilt afterwards from consecutive offsets
    bool any = false;
    for (int l = 1; l < 16; ++l) {
        const int c = (int)m.get(count + l);
        left = (left << 1) - c;
        if (left < 0) return false;                        // over-subscribed
        any |= c != 0;
        m.set(count + l, offs);
        offs += (uint32_t)c;
    }
    const uint32_t total = offs;
    if (any)
        for (int s = 0; s < n_sym; ++s) {
            const int l = (int)m.get(lens_at + s);
            if (l) { const uint32_t at = m.get(count + l); m.set(sorted + (int)at, (uint32_t)s); m.set(count + l, at + 1); }
        }
    // count[l] now holds the END offset of length l: back to counts
    uint32_t prev = 0;
    for (int l = 1; l < 16; ++l) { const uint32_t e = any ? m.get(count + l) : 0u; m.set(count + l, e - prev); prev = e; }
    (void)total;
    // where the canonical walk stands after `bits` lengths: codes longer than the lookup width start from there
    {
        uint32_t first = 0, index = 0;
        for (int l = 1; l <= bits; ++l) { const uint32_t c = m.get(count + l); index += c; first = (first + c) << 1; }
        m.set(count, index);
        m.set(walk_first, first);
    }
    for (int i = 0; i < (1 << bits); ++i) m.set(lut + i, 0);
    uint32_t code = 0;
    int idx = 0;
    for (int l = 1; l <= bits; ++l) {
        const int c = (int)m.get(count + l);
        for (int k = 0; k < c; ++k, ++idx, ++code) {
            const uint32_t e = m.get(sorted + idx) | ((uint32_t)l << 9);
            for (uint32_t r = rev_bits(code, l); r < (1u << bits); r += 1u << l) m.set(lut + (int)r, e);
        }
        code <<= 1;
    }
    return true;
}

// A code longer than the lookup width: the canonical walk, resumed where `bits` lengths leave it (one or two steps for
// the 10- and 11-bit codes that make up nearly all of these).  -1 on an invalid code.
template <int STRIDE>
INF_HD int walk(Lane &s, const Mem<STRIDE> &m, int bits, int count, int sorted, int walk_first) {
    int code = (int)(rev_bits32(peek(s, bits)) >> (32 - bits)) << 1;
    int first = (int)m.get(walk_first), index = (int)m.get(count);
    uint64_t b = s.bb >> bits;
    for (int l = bits + 1; l <= 15; ++l) {
        code |= (int)(b & 1);
        b >>= 1;
        const int c = (int)m.get(count + l);
        if (code - c < first) { drop(s, l); return (int)m.get(sorted + index + (code - first)); }
        index += c; first += c; first <<= 1; code <<= 1;
    }
    return -1;
}

// one symbol, -1 on an invalid code; the caller has refilled.  Lookup entries: symbol (9 bits) | code length << 9.
template <int STRIDE>
INF_HD int decode_sym(Lane &s, const Mem<STRIDE> &m, int lut, int bits, int count, int sorted, int walk_first) {
    const uint32_t e = m.get(lut + (int)peek(s, bits));
    const int len = (int)(e >> 9);
    if (len) { drop(s, len); return (int)(e & 511u); }
    return walk(s, m, bits, count, sorted, walk_first);
}

// Block header; for a Huffman block also its tables.  The long path of the machine (a few thousand instructions): the
// kernel lets lanes wait for one another in front of it so that several take it together.
template <int STRIDE>
INF_HD void header(Lane &s, const Mem<STRIDE> &m) {
    refill(s);
    s.last = take(s, 1) != 0;
    const uint32_t type = take(s, 2);
    if (type == 0) {
        drop(s, s.bc & 7);
        refill(s);
        const uint32_t len = take(s, 16);
        refill(s);
        const uint32_t nlen = take(s, 16);
        if ((len ^ 0xffffu) != nlen) return fail(s, ERR_DATA);
        if (s.op + len > s.out_len) return fail(s, ERR_OUTPUT);
        s.left = len;
        s.state = len ? ST_STORED : (s.last ? ST_END : ST_HEADER);
        return;
    }
    if (type == 3) return fail(s, ERR_DATA);
    int n_lit = 288, n_dist = 30;
    const int lens = kLens;                            // 288 literal / length + 32 distance code lengths
    if (type == 1) {
        for (int i = 0; i < 288; ++i) m.set(lens + i, i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8);
        for (int i = 0; i < 32; ++i) m.set(lens + 288 + i, i < 30 ? 5 : 0);
    } else {
        n_lit = (int)take(s, 5) + 257;
        n_dist = (int)take(s, 5) + 1;
        const int n_code = (int)take(s, 4) + 4;
        if (n_lit > 286 || n_dist > 30) return fail(s, ERR_DATA);
        const int cl = kLitSorted;                     // 19 lengths of the code-length code, staged where nothing lives yet
        for (int i = 0; i < 19; ++i) m.set(cl + i, 0);
        for (int i = 0; i < n_code; ++i) { refill(s); m.set(cl + INF_K(cl_order)[i], take(s, 3)); }
        if (!build(m, cl, 19, kClLut, 7, kDistCount, kDistSorted, kWalkFirst + 2)) return fail(s, ERR_DATA);
        int i = 0;
        uint32_t prev = 0;
        const int n_all = n_lit + n_dist;
        while (i < n_all) {
            refill(s);
            const int sym = decode_sym(s, m, kClLut, 7, kDistCount, kDistSorted, kWalkFirst + 2);
            if (sym < 0 || sym > 18 || overrun(s)) return fail(s, ERR_DATA);
            if (sym < 16) { prev = (uint32_t)sym; m.set(lens + (i < n_lit ? i : 288 + (i - n_lit)), prev); ++i; continue; }
            int rep;
            uint32_t val = 0;
            if (sym == 16) { if (i == 0) return fail(s, ERR_DATA); val = prev; rep = 3 + (int)take(s, 2); }
            else if (sym == 17) rep = 3 + (int)take(s, 3);
            else rep = 11 + (int)take(s, 7);
            if (i + rep > n_all) return fail(s, ERR_DATA);
            for (; rep > 0; --rep, ++i) m.set(lens + (i < n_lit ? i : 288 + (i - n_lit)), val);
            prev = val;
        }
        if (m.get(lens + 256) == 0) return fail(s, ERR_DATA);
        for (int k = n_lit; k < 288; ++k) m.set(lens + k, 0);
        for (int k = n_dist; k < 32; ++k) m.set(lens + 288 + k, 0);
    }
    // the distance code first: its lengths sit inside the literal lookup table, which the literal build clears
    if (!build(m, lens + 288, 32, kDistLut, kDistBits, kDistCount, kDistSorted, kWalkFirst + 1)) return fail(s, ERR_DATA);
    if (!build(m, lens, 288, kLitLut, kLitBits, kLitCount, kLitSorted, kWalkFirst)) return fail(s, ERR_DATA);
    s.state = ST_DECODE;
}

// One round of the machine in any state but ST_HEADER / ST_END: ONE Huffman code -- a literal / length code, or the
// distance code of the length decoded in the round before (ST_DIST) -- so that every lane of a warp runs the same
// lookup whatever its stream holds, and only the few instructions that act on the symbol differ.
template <int STRIDE>
INF_HD void round(Lane &s, const Mem<STRIDE> &m) {
    if (s.state == ST_DECODE || s.state == ST_DIST) {
        const bool want_dist = s.state == ST_DIST;
        refill(s);
        const int bits = want_dist ? kDistBits : kLitBits;
        const uint32_t e = m.get((want_dist ? kDistLut : kLitLut) + (int)peek(s, bits));
        const int code_len = (int)(e >> 9);
        int sym = (int)(e & 511u);
        if (code_len) drop(s, code_len);
        else {
            sym = want_dist ? walk(s, m, kDistBits, kDistCount, kDistSorted, kWalkFirst + 1) : walk(s, m, kLitBits, kLitCount, kLitSorted, kWalkFirst);
            if (sym < 0) return fail(s, ERR_DATA);
        }
        if (want_dist) {
            if (sym > 29) return fail(s, ERR_DATA);
            uint32_t dist = (uint32_t)sym + 1u;
            if (sym >= 4) {
                const int x = (sym >> 1) - 1;
                dist = ((uint32_t)(2 + (sym & 1)) << x) + 1u + take(s, x);
            }
            const uint32_t len = s.left;
            if (dist > s.op) return fail(s, ERR_DATA);
            flush_literals(s);
            s.tok[s.nt++] = (1u << 31) | (dist << 9) | len;
            s.op += len;
            s.state = ST_DECODE;
        } else if (sym < 256) {
            if (s.op >= s.out_len) return fail(s, ERR_OUTPUT);
            literal(s, (uint32_t)sym);
        } else if (sym == 256) {
            if (overrun(s)) return fail(s, ERR_INPUT);
            s.state = s.last ? ST_END : ST_HEADER;
        } else {
            if (sym > 285) return fail(s, ERR_DATA);
            uint32_t len = (uint32_t)sym - 254u;
            if (sym == 285) len = 258u;
            else if (sym >= 265) {
                const int x = (sym - 261) >> 2;
                len = ((uint32_t)(((sym - 265) & 3) + 4) << x) + 3u + take(s, x);
            }
            if (s.op + len > s.out_len) return fail(s, ERR_OUTPUT);
            s.left = len;
            s.state = ST_DIST;
        }
    } else if (s.state == ST_STORED) {
        uint32_t n = s.left < (uint32_t)kCopy ? s.left : (uint32_t)kCopy;
        for (; n > 0; --n) {
            refill(s);
            literal(s, take(s, 8));
            --s.left;
        }
        if (overrun(s)) return fail(s, ERR_INPUT);
        if (!s.left) s.state = s.last ? ST_END : ST_HEADER;
    }
}

// after ST_END: the literals still pending; returns the block's error code (the output size is part of it)
INF_HD int finish(Lane &s) {
    flush_literals(s);
    if (s.err) return s.err;
    if (overrun(s)) return ERR_INPUT;
    return s.op == s.out_len ? OK : ERR_OUTPUT;
}

// one token into bytes at dst + pos (sequential form of pass 2: the host test, and what lz_kernel must equal)
INF_HD uint32_t expand_token(uint32_t e, uint8_t *dst, uint32_t pos) {
    if (e >> 31) {
        const uint32_t len = e & 0x1ffu, dist = (e >> 9) & 0xffffu;
        for (uint32_t k = 0; k < len; ++k) dst[pos + k] = dst[pos + k - dist];
        return len;
    }
    const uint32_t n = ((e >> 29) & 3u) + 1u;
    for (uint32_t k = 0; k < n; ++k) dst[pos + k] = (uint8_t)(e >> (8 * k));
    return n;
}

// whole block on one thread: tokens, then bytes
template <int STRIDE>
INF_HD int inflate_block(const Mem<STRIDE> &m, const uint8_t *in, uint32_t in_len, uint32_t *tok, uint8_t *out, uint32_t out_len) {
    Lane s;
    lane_init(s, in, in_len, tok, out_len);
    while (s.state != ST_END) {
        if (s.state == ST_HEADER) header(s, m);
        else round(s, m);
    }
    const int err = finish(s);
    if (err) return err;
    uint32_t pos = 0;
    for (uint32_t i = 0; i < s.nt; ++i) pos += expand_token(tok[i], out, pos);
    return pos == out_len ? OK : ERR_OUTPUT;
}

}  // namespace lane
}  // namespace dinflate
