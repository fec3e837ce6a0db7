// DEFLATE (RFC 1951) decoder core for BGZF blocks, written so that ONE warp inflates one block:
//   * lane 0 runs the serial part -- bit reader, Huffman decode through a 10-bit lookup table in shared memory
//     (longer codes fall back to the canonical count / sorted-symbol walk), length / distance extra bits -- and
//     queues up to 32 symbols (literal byte, or match length + distance);
//   * the whole warp then flushes the queue: output positions by a prefix sum over the symbol lengths, every literal
//     written by its own lane (consecutive lanes -> consecutive bytes), every match copied by all lanes, matches taken
//     in stream order because a match may read what an earlier one wrote.
// The same source compiles for the host with one "lane" (tests/test_inflate_host.py checks it against zlib on BGZF
// blocks and on stored / fixed / dynamic streams of every compression level); the CUDA kernel in bamdevice.cu supplies
// the warp primitives.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define INF_HD __host__ __device__ __forceinline__
#else
#define INF_HD inline
#endif

namespace dinflate {

constexpr int kLitBits = 10, kDistBits = 8, kBatch = 32;

enum : int { OK = 0, ERR_DATA = 1, ERR_INPUT = 2, ERR_OUTPUT = 3 };

// A lookup entry has the symbol resolved: bits 0..3 code length (0 = code longer than the table, or unused), bits
// 4..5 kind (1 literal, 2 match length, 3 end of block; distances: 2), bits 8..11 number of extra bits, bits 16..31 the
// literal byte / length base / distance base.  The code-length code of a dynamic block uses kind 0, value = symbol.
struct Tables {                    // per warp, in shared memory (~6.4 KB)
    uint32_t lit_lut[1 << kLitBits];
    uint32_t dist_lut[1 << kDistBits];
    uint16_t lit_sorted[288], dist_sorted[32];
    uint16_t lit_count[16], dist_count[16];
    uint8_t lens[320];                     // 288 literal / length + 30 distance code lengths; [300, 319) scratch of the code-length code
    uint32_t batch[kBatch];                // literal: byte; match: 1 << 31 | distance << 9 | length
};

// Bit reader: `buf` holds at least `cnt` valid bits of the stream (bits above them are a preview of the bytes that
// follow, or of whatever lies behind the input -- the buffers are padded by 16 bytes); refill() loads 16 aligned bytes
// and always leaves cnt >= 56: one check in front of a literal / length symbol (20 bits at most) and one in front of
// a distance (28 bits) are all the decode loop needs.  Reading past the end of the input is detected per batch from
// the bit position.
struct BitReader {
    const uint8_t *p, *in, *end;
    uint64_t buf;
    int cnt;
    INF_HD void init(const uint8_t *src, uint32_t n) { p = in = src; end = src + n; buf = 0; cnt = 0; }
    INF_HD void refill() {
        const uintptr_t a = (uintptr_t)p;
        const uint64_t *w = reinterpret_cast<const uint64_t *>(a & ~(uintptr_t)7);
        const int sh = (int)(a & 7) * 8;
        const uint64_t q0 = w[0], q1 = w[1];
        const uint64_t v = sh ? ((q0 >> sh) | (q1 << (64 - sh))) : q0;
        buf |= v << cnt;
        const int take = (63 - cnt) >> 3;
        p += take;
        cnt += take * 8;
    }
    INF_HD uint32_t peek(int n) const { return (uint32_t)(buf & ((1ull << n) - 1)); }
    INF_HD void drop(int n) { buf >>= n; cnt -= n; }
    INF_HD uint32_t take(int n) { const uint32_t v = peek(n); drop(n); return v; }          // the caller keeps cnt >= n
    INF_HD bool overrun() const { return (int64_t)(p - in) * 8 - cnt > (int64_t)(end - in) * 8; }
};

// RFC 1951 tables: once in constant memory for the device pass, once as plain host constants
#define INF_TABLES(QUAL, P)                                                                                                        \
    QUAL uint16_t P##len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258}; \
    QUAL uint8_t P##len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};      \
    QUAL uint16_t P##dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577}; \
    QUAL uint8_t P##dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13}; \
    QUAL uint8_t P##cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
#ifdef __CUDACC__
INF_TABLES(__device__ __constant__, d_k_)
#endif
INF_TABLES(static const, h_k_)
#ifdef __CUDA_ARCH__
#define INF_K(name) d_k_##name
#else
#define INF_K(name) h_k_##name
#endif

enum : int { KIND_CL = 0, KIND_LIT = 1, KIND_DIST = 2 };

// the resolved form of a symbol (without its code length): kind << 4 | extra << 8 | value << 16; ~0 = invalid symbol
INF_HD uint32_t resolve(int sym, int kind) {
    if (kind == KIND_CL) return (uint32_t)sym << 16;
    if (kind == KIND_LIT) {
        if (sym < 256) return (1u << 4) | ((uint32_t)sym << 16);
        if (sym == 256) return 3u << 4;
        if (sym > 285) return ~0u;
        return (2u << 4) | ((uint32_t)INF_K(len_extra)[sym - 257] << 8) | ((uint32_t)INF_K(len_base)[sym - 257] << 16);
    }
    if (sym > 29) return ~0u;
    return (2u << 4) | ((uint32_t)INF_K(dist_extra)[sym] << 8) | ((uint32_t)INF_K(dist_base)[sym] << 16);
}

INF_HD uint32_t rev_bits(uint32_t v, int n) {
    uint32_t r = 0;
    for (int i = 0; i < n; ++i) { r = (r << 1) | (v & 1); v >>= 1; }
    return r;
}

// Canonical Huffman tables from code lengths: lookup table over `bits` bits + count / sorted arrays for the slow path.
// Single-lane version (a dynamic block has at most 320 symbols; decoding the block's data dwarfs this).
INF_HD bool build(const uint8_t *lens, int n_sym, uint32_t *lut, int bits, uint16_t *count, uint16_t *sorted, int kind) {
    for (int i = 0; i < 16; ++i) count[i] = 0;
    for (int s = 0; s < n_sym; ++s) ++count[lens[s]];
    for (int i = 0; i < (1 << bits); ++i) lut[i] = 0;
    if (count[0] == n_sym) return true;               // no codes at all (legal for the distance tree)
    int left = 1;
    for (int l = 1; l < 16; ++l) { left = (left << 1) - count[l]; if (left < 0) return false; }      // over-subscribed
    uint16_t offs[16];
    offs[1] = 0;
    for (int l = 1; l < 15; ++l) offs[l + 1] = (uint16_t)(offs[l] + count[l]);
    for (int s = 0; s < n_sym; ++s) if (lens[s]) sorted[offs[lens[s]]++] = (uint16_t)s;
    // codes in canonical order
    uint32_t code = 0;
    int idx = 0;
    for (int l = 1; l <= bits; ++l) {
        for (int k = 0; k < count[l]; ++k, ++idx, ++code) {
            const uint32_t rs = resolve(sorted[idx], kind);
            const uint32_t e = rs == ~0u ? 0u : (rs | (uint32_t)l);       // an invalid symbol decodes through the slow path, which reports it
            for (uint32_t r = rev_bits(code, l); r < (1u << bits); r += 1u << l) lut[r] = e;
        }
        code <<= 1;
    }
    return true;
}

// one symbol in its resolved form (see Tables); `bits` = lookup width.  Returns ~0 on an invalid code.
INF_HD uint32_t decode(BitReader &br, const uint32_t *lut, int bits, const uint16_t *count, const uint16_t *sorted, int kind) {
    const uint32_t e = lut[br.peek(bits)];
    if (e & 15) { br.drop((int)(e & 15)); return e; }
    // code longer than the table: canonical walk, one bit at a time
    int code = 0, first = 0, index = 0;
    uint64_t b = br.buf;
    for (int l = 1; l <= 15; ++l) {
        code |= (int)(b & 1);
        b >>= 1;
        const int c = count[l];
        if (code - c < first) { br.drop(l); return resolve(sorted[index + (code - first)], kind); }
        index += c; first += c; first <<= 1; code <<= 1;
    }
    return ~0u;
}

struct State {
    BitReader br;
    uint32_t out_pos, out_len;
    int phase;            // 0 block header expected, 1 inside a Huffman block, 2 inside a stored block, 3 done
    bool last;
    uint32_t stored_left;
    int err;
};

INF_HD void state_init(State &s, const uint8_t *in, uint32_t in_len, uint32_t out_len) {
    s.br.init(in, in_len);
    s.out_pos = 0; s.out_len = out_len; s.phase = 0; s.last = false; s.stored_left = 0; s.err = OK;
}

// Reads a block header and builds the tables of a Huffman block.
INF_HD void begin_block(State &s, Tables &t) {
    BitReader &br = s.br;
    br.refill();
    s.last = br.take(1) != 0;
    const uint32_t type = br.take(2);
    if (type == 0) {
        br.drop(br.cnt & 7);              // to the next byte boundary
        // the bit buffer holds whole bytes now: put them back
        br.p -= br.cnt >> 3; br.buf = 0; br.cnt = 0;
        if (br.p + 4 > br.end) { s.err = ERR_INPUT; return; }
        const uint32_t len = br.p[0] | (br.p[1] << 8), nlen = br.p[2] | (br.p[3] << 8);
        if ((len ^ 0xffffu) != nlen) { s.err = ERR_DATA; return; }
        br.p += 4;
        if (br.p + len > br.end) { s.err = ERR_INPUT; return; }
        if (s.out_pos + len > s.out_len) { s.err = ERR_OUTPUT; return; }
        s.stored_left = len;
        s.phase = 2;
        return;
    }
    if (type == 3) { s.err = ERR_DATA; return; }
    int n_lit = 288, n_dist = 30;
    if (type == 1) {
        for (int i = 0; i < 144; ++i) t.lens[i] = 8;
        for (int i = 144; i < 256; ++i) t.lens[i] = 9;
        for (int i = 256; i < 280; ++i) t.lens[i] = 7;
        for (int i = 280; i < 288; ++i) t.lens[i] = 8;
        for (int i = 0; i < 30; ++i) t.lens[288 + i] = 5;
    } else {
        n_lit = (int)br.take(5) + 257;
        n_dist = (int)br.take(5) + 1;
        const int n_code = (int)br.take(4) + 4;
        if (n_lit > 286 || n_dist > 30) { s.err = ERR_DATA; return; }
        uint8_t *cl = t.lens + 300;        // 19 scratch bytes behind the distance lengths
        for (int i = 0; i < 19; ++i) cl[i] = 0;
        for (int i = 0; i < n_code; ++i) { if (br.cnt < 8) br.refill(); cl[INF_K(cl_order)[i]] = (uint8_t)br.take(3); }
        // the code-length code borrows the distance tables (7-bit lookup)
        if (!build(cl, 19, t.dist_lut, 7, t.dist_count, t.dist_sorted, KIND_CL)) { s.err = ERR_DATA; return; }
        int i = 0;
        while (i < n_lit + n_dist) {
            if (br.cnt < 32) br.refill();
            const uint32_t e = decode(br, t.dist_lut, 7, t.dist_count, t.dist_sorted, KIND_CL);
            if (e == ~0u || br.overrun()) { s.err = ERR_DATA; return; }
            const int sym = (int)(e >> 16);
            if (sym < 16) { t.lens[i++] = (uint8_t)sym; continue; }
            int rep, val = 0;
            if (sym == 16) { if (i == 0) { s.err = ERR_DATA; return; } val = t.lens[i - 1]; rep = 3 + (int)br.take(2); }
            else if (sym == 17) rep = 3 + (int)br.take(3);
            else rep = 11 + (int)br.take(7);
            if (i + rep > n_lit + n_dist) { s.err = ERR_DATA; return; }
            while (rep--) t.lens[i++] = (uint8_t)val;
        }
        if (t.lens[256] == 0) { s.err = ERR_DATA; return; }
        // distance lengths to their own array position
        for (int k = n_dist - 1; k >= 0; --k) t.lens[288 + k] = t.lens[n_lit + k];
        for (int k = n_lit; k < 288; ++k) t.lens[k] = 0;
    }
    if (!build(t.lens, 288, t.lit_lut, kLitBits, t.lit_count, t.lit_sorted, KIND_LIT)) { s.err = ERR_DATA; return; }
    if (!build(t.lens + 288, n_dist, t.dist_lut, kDistBits, t.dist_count, t.dist_sorted, KIND_DIST)) { s.err = ERR_DATA; return; }
    s.phase = 1;
}

// Lane 0: decodes up to kBatch symbols of the current Huffman block into t.batch.  Returns the number queued and, in
// *bytes, the output bytes they stand for.  s.phase leaves 1 at the end-of-block symbol.
INF_HD int decode_batch(State &s, Tables &t, uint32_t *bytes) {
    BitReader &br = s.br;
    int n = 0;
    uint32_t produced = 0;
    const uint32_t room = s.out_len - s.out_pos;
    while (n < kBatch) {
        if (br.cnt < 20) br.refill();          // a literal / length code (<= 15 bits) and the length's extra bits (<= 5)
        const uint32_t e = decode(br, t.lit_lut, kLitBits, t.lit_count, t.lit_sorted, KIND_LIT);
        const uint32_t kind = (e >> 4) & 3;
        if (kind == 1) {
            if (produced + 1 > room) { s.err = ERR_OUTPUT; break; }
            t.batch[n++] = e >> 16;
            ++produced;
            continue;
        }
        if (e == ~0u) { s.err = ERR_DATA; break; }
        if (kind == 3) { s.phase = s.last ? 3 : 0; break; }
        const uint32_t len = (e >> 16) + br.take((int)((e >> 8) & 15));
        if (br.cnt < 28) br.refill();          // a distance code (<= 15 bits) and its extra bits (<= 13)
        const uint32_t d = decode(br, t.dist_lut, kDistBits, t.dist_count, t.dist_sorted, KIND_DIST);
        if (d == ~0u) { s.err = ERR_DATA; break; }
        const uint32_t dist = (d >> 16) + br.take((int)((d >> 8) & 15));
        if (dist > s.out_pos + produced) { s.err = ERR_DATA; break; }
        if (produced + len > room) { s.err = ERR_OUTPUT; break; }
        t.batch[n++] = (1u << 31) | (dist << 9) | len;
        produced += len;
    }
    if (s.err == OK && br.overrun()) s.err = ERR_INPUT;
    *bytes = produced;
    return n;
}

}  // namespace dinflate
