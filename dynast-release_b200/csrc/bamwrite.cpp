// BAM writer: packed records -> alignment records -> BGZF blocks.
//
// Replaces what the reference gets from pysam when call_consensus writes consensus.bam
// (dynast/preprocessing/consensus.py:360-366, 476-494: AlignedSegment objects written one by one, then
// `samtools sort` + `samtools index`), and builds the synthetic BAMs of the ingest benchmarks.
//   * a unit's packed record (full layout: 4-bit sequence and Phred are already in BAM encoding) becomes one BAM
//     alignment record: fixed fields, read name, CIGAR words, sequence, qualities, the caller's pre-encoded aux
//     bytes, and the MD tag rebuilt from the record's MD tokens;
//   * units are written in the caller's order (`h_order`: coordinate order for a sorted file);
//   * all threads encode ranges of units into <= 0xff00-byte payloads that end on record boundaries (as htslib
//     writes them, so that range readers can start at any block) and deflate them (zlib, raw deflate + BGZF
//     header with the BC field, CRC32, ISIZE); the blocks are then written in order, followed by the EOF block;
//   * with `write_index` a .bai goes next to the file: the UCSC binning index plus the 16 kb linear index over the
//     records' virtual offsets (SAM specification section 5), so that pysam.fetch(contig) works on the output.
#include <zlib.h>

#include <atomic>
#include <map>
#include <mutex>
#include <thread>

#include "bamhost.h"

using namespace bamhost;

namespace {

constexpr size_t kPayload = 0xff00;

struct OutBlock {
    std::vector<uint8_t> payload;       // inflated bytes (dropped after compression)
    std::vector<uint8_t> comp;          // the finished BGZF block
    std::vector<uint32_t> rec_at;       // payload offsets of the records that start here
    std::vector<int64_t> rec_unit;      // their units
};

void put16(std::vector<uint8_t> &v, uint16_t x) { v.push_back((uint8_t)x); v.push_back((uint8_t)(x >> 8)); }
void put32(std::vector<uint8_t> &v, uint32_t x) { for (int i = 0; i < 4; ++i) v.push_back((uint8_t)(x >> (8 * i))); }

bool deflate_block(OutBlock &b, int level) {
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
    const size_t bound = deflateBound(&zs, (uLong)b.payload.size());
    b.comp.resize(18 + bound + 8);
    static const uint8_t head[16] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 'B', 'C', 2, 0};
    memcpy(b.comp.data(), head, 16);
    zs.next_in = b.payload.data(); zs.avail_in = (uInt)b.payload.size();
    zs.next_out = b.comp.data() + 18; zs.avail_out = (uInt)bound;
    const int rc = deflate(&zs, Z_FINISH);
    const size_t clen = bound - zs.avail_out;
    deflateEnd(&zs);
    if (rc != Z_STREAM_END || 18 + clen + 8 > 0x10000) return false;
    const uint32_t bsize = (uint32_t)(18 + clen + 8 - 1);
    b.comp[16] = (uint8_t)bsize; b.comp[17] = (uint8_t)(bsize >> 8);
    const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), b.payload.data(), (uInt)b.payload.size());
    const uint32_t isize = (uint32_t)b.payload.size();
    uint8_t *t = b.comp.data() + 18 + clen;
    for (int i = 0; i < 4; ++i) { t[i] = (uint8_t)(crc >> (8 * i)); t[4 + i] = (uint8_t)(isize >> (8 * i)); }
    b.comp.resize(18 + clen + 8);
    return true;
}

int reg2bin(int64_t beg, int64_t end) {
    --end;
    if (beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}

// MD text of a tokenised record (the inverse of tokenize_md for well-formed tags)
void md_text(const uint32_t *tok, uint32_t n_tok, uint32_t m_total, std::string &out) {
    static const char codes[] = "=ACMGRSVTWYHKDBN";
    out.clear();
    uint32_t run_start = 0, del_at = 0;
    bool prev_del = false;
    for (uint32_t i = 0; i < n_tok; ++i) {
        const uint32_t t = tok[i], aoff = t & 0xffff, code = (t >> 16) & 15;
        if ((t >> 20) & 1) {
            // the deleted bases of one D run share their aligned offset; another offset starts another ^ run
            if (!prev_del || aoff != del_at) { out += std::to_string(aoff - run_start); out += '^'; run_start = aoff; del_at = aoff; }
            out += codes[code];
            prev_del = true;
        } else {
            out += std::to_string(aoff - run_start);
            out += codes[code];
            run_start = aoff + 1;
            prev_del = false;
        }
    }
    out += std::to_string(m_total - run_start);
}

}  // namespace

extern "C" int dyn_bam_write(const char *path, const char *header_text, int32_t n_ref, const char *const *ref_names,
                             const int32_t *ref_lens, const void *h_records, const uint32_t *h_rec_off, int64_t n_units,
                             const int64_t *h_order, const char *h_names, const uint32_t *h_name_off, const uint8_t *h_aux,
                             const uint32_t *h_aux_off, const uint8_t *h_mapq, const uint16_t *h_flag, int level, int n_threads,
                             int write_index) {
    dyn_bam_write_opts_t o;
    memset(&o, 0, sizeof(o));
    o.header_text = header_text; o.n_ref = n_ref; o.ref_names = ref_names; o.ref_lens = ref_lens;
    o.h_records = h_records; o.h_rec_off = h_rec_off; o.n_units = n_units; o.h_order = h_order;
    o.h_names = h_names; o.h_name_off = h_name_off; o.h_aux = h_aux; o.h_aux_off = h_aux_off; o.h_mapq = h_mapq; o.h_flag = h_flag;
    o.level = level; o.n_threads = n_threads; o.write_index = write_index;
    return dyn_bam_write_ex(path, &o);
}

namespace {

// swap_gene_tags (dynast/preprocessing/consensus.py:277-286) on the aux bytes of a BAM record: the `gene_tag` tag is
// dropped, GX is set to `gene` and -- when the gene has a name -- GN to it (a tag that is already there keeps its place,
// a new one goes to the end).  Other tags keep their bytes.
void retag(const uint8_t *aux, size_t n, const char *gene_tag, const char *gene, const char *name, std::vector<uint8_t> &out) {
    const uint8_t *a = aux, *end = aux + n;
    bool gx_done = false, gn_done = false;
    auto z_tag = [&](const char *tag, const char *val) {
        out.push_back((uint8_t)tag[0]); out.push_back((uint8_t)tag[1]); out.push_back('Z');
        out.insert(out.end(), val, val + strlen(val)); out.push_back(0);
    };
    while (a + 3 <= end) {
        const uint8_t *t0 = a;
        const char ty = (char)a[2];
        a += 3;
        if (ty == 'A' || ty == 'c' || ty == 'C') a += 1;
        else if (ty == 's' || ty == 'S') a += 2;
        else if (ty == 'i' || ty == 'I' || ty == 'f') a += 4;
        else if (ty == 'Z' || ty == 'H') { while (a < end && *a) ++a; ++a; }
        else if (ty == 'B') { if (a + 5 > end) break; const char sub = (char)a[0]; const uint32_t cnt = rd32(a + 1); a += 5 + (size_t)cnt * ((sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4); }
        else break;
        if (a > end) a = end;
        const bool is_gene_tag = gene_tag && gene_tag[0] == (char)t0[0] && gene_tag[1] == (char)t0[1];
        if (is_gene_tag) continue;
        if (t0[0] == 'G' && t0[1] == 'X') { z_tag("GX", gene); gx_done = true; continue; }
        if (t0[0] == 'G' && t0[1] == 'N' && name && name[0]) { z_tag("GN", name); gn_done = true; continue; }
        out.insert(out.end(), t0, a);
    }
    if (!gx_done) z_tag("GX", gene);
    if (!gn_done && name && name[0]) z_tag("GN", name);
}

}  // namespace

extern "C" int dyn_bam_write_ex(const char *path, const dyn_bam_write_opts_t *opts) {
    if (!path || !opts) { dyn_set_error("dyn_bam_write: null argument"); return DYN_ERR_ARG; }
    const char *header_text = opts->header_text;
    const int32_t n_ref = opts->n_ref;
    const char *const *ref_names = opts->ref_names;
    const int32_t *ref_lens = opts->ref_lens;
    const void *h_records = opts->h_records;
    const uint32_t *h_rec_off = opts->h_rec_off;
    const int64_t n_units = opts->n_units;
    const int64_t *h_order = opts->h_order;
    const char *h_names = opts->h_names;
    const uint32_t *h_name_off = opts->h_name_off;
    const uint8_t *h_aux = opts->h_aux;
    const uint32_t *h_aux_off = opts->h_aux_off;
    const uint8_t *h_mapq = opts->h_mapq;
    const uint16_t *h_flag = opts->h_flag;
    const int level = opts->level, n_threads = opts->n_threads, write_index = opts->write_index;
    const uint8_t *h_raw = opts->h_raw;
    const uint64_t *h_raw_start = opts->h_raw_start;
    const uint32_t *h_raw_len = opts->h_raw_len;
    const char *h_retag = opts->h_retag;
    const uint32_t *h_retag_off = opts->h_retag_off;
    const char *gene_tag = (opts->gene_tag && opts->gene_tag[0] && opts->gene_tag[1]) ? opts->gene_tag : nullptr;
    auto is_raw = [&](int64_t u) { return h_raw && h_raw_start && h_raw_len && h_raw_len[u] > 0; };
    if (!header_text || n_ref < 0 || n_units < 0 || (n_units && !h_raw && (!h_records || !h_rec_off))) {
        dyn_set_error("dyn_bam_write: bad argument");
        return DYN_ERR_ARG;
    }
    const uint8_t *records = static_cast<const uint8_t *>(h_records);
    const int nt = std::max(1, std::min(n_threads, 64));
    // ---- header
    std::vector<OutBlock> head_blocks;
    {
        std::vector<uint8_t> h;
        h.insert(h.end(), {'B', 'A', 'M', 1});
        put32(h, (uint32_t)strlen(header_text));
        h.insert(h.end(), header_text, header_text + strlen(header_text));
        put32(h, (uint32_t)n_ref);
        for (int32_t r = 0; r < n_ref; ++r) {
            const size_t l = strlen(ref_names[r]) + 1;
            put32(h, (uint32_t)l);
            h.insert(h.end(), ref_names[r], ref_names[r] + l);
            put32(h, (uint32_t)ref_lens[r]);
        }
        for (size_t o = 0; o < h.size(); o += kPayload) {
            head_blocks.emplace_back();
            head_blocks.back().payload.assign(h.begin() + o, h.begin() + std::min(h.size(), o + kPayload));
        }
    }
    // ---- records: every thread encodes and deflates ranges of units
    const size_t n_parts = (size_t)std::max<int64_t>(1, std::min<int64_t>((int64_t)nt * 8, n_units / 256 + 1));
    std::vector<std::vector<OutBlock>> parts(n_parts);
    std::atomic<size_t> next(0);
    std::atomic<int> failed(0);
    std::string fail_msg;
    std::mutex fail_mu;
    auto fail = [&](const char *m) { std::lock_guard<std::mutex> l(fail_mu); if (!failed.exchange(1)) fail_msg = m; };
    auto work = [&]() {
        std::string md, name;
        std::vector<uint8_t> rec;
        for (;;) {
            const size_t pi = next.fetch_add(1);
            if (pi >= n_parts || failed.load()) return;
            std::vector<OutBlock> &blocks = parts[pi];
            const int64_t u0 = n_units * (int64_t)pi / (int64_t)n_parts, u1 = n_units * (int64_t)(pi + 1) / (int64_t)n_parts;
            blocks.emplace_back();
            for (int64_t j = u0; j < u1; ++j) {
                const int64_t u = h_order ? h_order[j] : j;
                if (is_raw(u)) {
                    // a record of the input file written through (consensus.py:420-422, 443-444, 476-478), its gene tags
                    // swapped when asked to
                    const uint8_t *r = h_raw + h_raw_start[u];
                    const size_t rn = (size_t)h_raw_len[u];
                    rec.clear();
                    put32(rec, 0);
                    const bool swap = h_retag && h_retag_off && h_retag_off[u + 1] > h_retag_off[u];
                    if (!swap) rec.insert(rec.end(), r, r + rn);
                    else {
                        const uint32_t l_name = r[8], n_cig = rd16(r + 12), l_sq = rd32(r + 16);
                        const size_t fixed = 32 + l_name + 4 * (size_t)n_cig + (l_sq + 1) / 2 + l_sq;
                        if (fixed > rn) { fail("dyn_bam_write: corrupt raw record"); return; }
                        rec.insert(rec.end(), r, r + fixed);
                        const char *g = h_retag + h_retag_off[u];
                        const char *gn = g + strlen(g) + 1;
                        retag(r + fixed, rn - fixed, gene_tag, g, gn < h_retag + h_retag_off[u + 1] ? gn : "", rec);
                    }
                    const uint32_t bs = (uint32_t)(rec.size() - 4);
                    for (int i = 0; i < 4; ++i) rec[i] = (uint8_t)(bs >> (8 * i));
                    if (!blocks.back().payload.empty() && blocks.back().payload.size() + rec.size() > kPayload) blocks.emplace_back();
                    size_t o = 0;
                    blocks.back().rec_at.push_back((uint32_t)blocks.back().payload.size());
                    blocks.back().rec_unit.push_back(u);
                    while (o < rec.size()) {
                        if (blocks.back().payload.size() == kPayload) blocks.emplace_back();
                        const size_t take = std::min(rec.size() - o, kPayload - blocks.back().payload.size());
                        blocks.back().payload.insert(blocks.back().payload.end(), rec.begin() + o, rec.begin() + o + take);
                        o += take;
                    }
                    continue;
                }
                const uint8_t *p = records + ((size_t)h_rec_off[u] << 4);
                const int32_t pos = (int32_t)rd32(p), ref_id = (int32_t)rd32(p + 4);
                const uint32_t l_seq = rd16(p + 8), n_cigar = rd16(p + 10), n_tok = rd16(p + 12);
                const uint16_t pflag = rd16(p + 14);
                if (pflag & (DYN_REC_LEAN | DYN_REC_HAS_MATE)) { fail("dyn_bam_write: lean records and units with a mate cannot be written"); return; }
                const uint8_t *cig = p + 16, *tokb = cig + 4 * (size_t)n_cigar;
                const uint8_t *seq = tokb + 4 * (size_t)n_tok, *qual = seq + ((((size_t)l_seq + 1) / 2 + 3) & ~(size_t)3);
                uint32_t m_total = 0;
                int64_t ref_span = 0;
                for (uint32_t c = 0; c < n_cigar; ++c) {
                    const uint32_t w = rd32(cig + 4 * (size_t)c), op = w & 0xf, len = w >> 4;
                    if (op == 0 || op == 7 || op == 8) { m_total += len; ref_span += len; }
                    else if (op == 2 || op == 3) ref_span += len;
                }
                if (opts->h_md && opts->h_md_off && opts->h_md_off[u + 1] > opts->h_md_off[u])
                    md.assign(opts->h_md + opts->h_md_off[u], opts->h_md + opts->h_md_off[u + 1]);      // the caller's MD text, verbatim
                else md_text(reinterpret_cast<const uint32_t *>(tokb), n_tok, m_total, md);
                if (h_names) name.assign(h_names + h_name_off[u], h_names + h_name_off[u + 1]);
                else name = "r" + std::to_string(u);
                if (name.size() > 254) { fail("dyn_bam_write: read name longer than 254 characters"); return; }
                const size_t aux_len = h_aux ? (size_t)(h_aux_off[u + 1] - h_aux_off[u]) : 0;
                rec.clear();
                put32(rec, 0);                     // block_size, patched below
                put32(rec, (uint32_t)ref_id); put32(rec, (uint32_t)pos);
                rec.push_back((uint8_t)(name.size() + 1));
                rec.push_back(h_mapq ? h_mapq[u] : 255);
                put16(rec, (uint16_t)reg2bin(pos, pos + std::max<int64_t>(ref_span, 1)));
                put16(rec, (uint16_t)n_cigar);
                put16(rec, h_flag ? h_flag[u] : (uint16_t)(pflag & 0x0fff));
                put32(rec, l_seq);
                put32(rec, (uint32_t)-1); put32(rec, (uint32_t)-1); put32(rec, 0);     // no mate, template length 0
                rec.insert(rec.end(), name.begin(), name.end()); rec.push_back(0);
                rec.insert(rec.end(), cig, cig + 4 * (size_t)n_cigar);
                rec.insert(rec.end(), seq, seq + (l_seq + 1) / 2);
                rec.insert(rec.end(), qual, qual + l_seq);
                if (aux_len) rec.insert(rec.end(), h_aux + h_aux_off[u], h_aux + h_aux_off[u] + aux_len);
                rec.insert(rec.end(), {'M', 'D', 'Z'});
                rec.insert(rec.end(), md.begin(), md.end()); rec.push_back(0);
                const uint32_t bs = (uint32_t)(rec.size() - 4);
                for (int i = 0; i < 4; ++i) rec[i] = (uint8_t)(bs >> (8 * i));
                // htslib style: a record never straddles blocks unless it is larger than one
                if (!blocks.back().payload.empty() && blocks.back().payload.size() + rec.size() > kPayload) blocks.emplace_back();
                size_t o = 0;
                blocks.back().rec_at.push_back((uint32_t)blocks.back().payload.size());
                blocks.back().rec_unit.push_back(u);
                while (o < rec.size()) {
                    if (blocks.back().payload.size() == kPayload) blocks.emplace_back();
                    const size_t take = std::min(rec.size() - o, kPayload - blocks.back().payload.size());
                    blocks.back().payload.insert(blocks.back().payload.end(), rec.begin() + o, rec.begin() + o + take);
                    o += take;
                }
            }
            for (OutBlock &b : blocks) {
                if (b.payload.empty()) continue;
                if (!deflate_block(b, level)) { fail("dyn_bam_write: deflate failed"); return; }
                if (!write_index) { std::vector<uint8_t>().swap(b.payload); }
            }
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back(work);
        for (auto &t : th) t.join();
    }
    if (failed.load()) { dyn_set_error("%s", fail_msg.c_str()); return DYN_ERR_FORMAT; }
    for (OutBlock &b : head_blocks)
        if (!deflate_block(b, level)) { dyn_set_error("dyn_bam_write: deflate failed"); return DYN_ERR_FORMAT; }

    // ---- write, collecting the virtual offsets for the index
    FILE *f = fopen(path, "wb");
    if (!f) { dyn_set_error("dyn_bam_write: cannot create %s", path); return DYN_ERR_IO; }
    uint64_t file_off = 0;
    bool io_ok = true;
    for (OutBlock &b : head_blocks) { io_ok &= fwrite(b.comp.data(), 1, b.comp.size(), f) == b.comp.size(); file_off += b.comp.size(); }
    struct Bin { std::vector<std::pair<uint64_t, uint64_t>> chunks; };
    std::vector<std::map<uint32_t, Bin>> bins(write_index ? (size_t)n_ref : 0);
    std::vector<std::vector<uint64_t>> linear(write_index ? (size_t)n_ref : 0);
    std::vector<uint64_t> n_mapped(write_index ? (size_t)n_ref : 0, 0), ref_beg(write_index ? (size_t)n_ref : 0, ~0ull),
        ref_end(write_index ? (size_t)n_ref : 0, 0);
    // a record's end offset = the next record's start; remember the previous record to close its chunk
    int32_t prev_ref = -1; uint32_t prev_bin = 0; uint64_t prev_start = 0; bool have_prev = false;
    auto close_prev = [&](uint64_t voff_end) {
        if (!have_prev || prev_ref < 0) return;
        auto &ch = bins[(size_t)prev_ref][prev_bin].chunks;
        if (!ch.empty() && ch.back().second == prev_start) ch.back().second = voff_end;     // adjacent: merge
        else ch.emplace_back(prev_start, voff_end);
        ref_end[(size_t)prev_ref] = voff_end;
    };
    for (auto &blocks : parts)
        for (OutBlock &b : blocks) {
            if (b.comp.empty()) continue;
            if (write_index) {
                for (size_t k = 0; k < b.rec_at.size(); ++k) {
                    const uint64_t voff = (file_off << 16) | b.rec_at[k];
                    close_prev(voff);
                    int32_t pos, ref_id;
                    uint32_t n_cigar;
                    const uint8_t *cg;
                    if (is_raw(b.rec_unit[k])) {
                        const uint8_t *r = h_raw + h_raw_start[b.rec_unit[k]];
                        ref_id = (int32_t)rd32(r); pos = (int32_t)rd32(r + 4); n_cigar = rd16(r + 12); cg = r + 32 + r[8];
                    } else {
                        const uint8_t *p = records + ((size_t)h_rec_off[b.rec_unit[k]] << 4);
                        pos = (int32_t)rd32(p); ref_id = (int32_t)rd32(p + 4); n_cigar = rd16(p + 10); cg = p + 16;
                    }
                    int64_t span = 0;
                    for (uint32_t c = 0; c < n_cigar; ++c) {
                        const uint32_t w = rd32(cg + 4 * (size_t)c), op = w & 0xf;
                        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) span += w >> 4;
                    }
                    const int64_t end = pos + std::max<int64_t>(span, 1);
                    prev_ref = ref_id; prev_bin = (uint32_t)reg2bin(pos, end); prev_start = voff; have_prev = true;
                    if (ref_id >= 0 && ref_id < n_ref) {
                        ++n_mapped[(size_t)ref_id];
                        ref_beg[(size_t)ref_id] = std::min(ref_beg[(size_t)ref_id], voff);
                        auto &lin = linear[(size_t)ref_id];
                        const size_t w0 = (size_t)(pos >> 14), w1 = (size_t)((end - 1) >> 14);
                        if (lin.size() <= w1) lin.resize(w1 + 1, 0);
                        for (size_t w = w0; w <= w1; ++w) if (lin[w] == 0) lin[w] = voff;
                    }
                }
            }
            io_ok &= fwrite(b.comp.data(), 1, b.comp.size(), f) == b.comp.size();
            file_off += b.comp.size();
        }
    if (write_index) close_prev(file_off << 16);
    static const uint8_t eof_block[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0, 0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    io_ok &= fwrite(eof_block, 1, 28, f) == 28;
    io_ok &= fclose(f) == 0;
    if (!io_ok) { dyn_set_error("dyn_bam_write: write to %s failed", path); return DYN_ERR_IO; }

    if (write_index) {
        std::vector<uint8_t> ix;
        ix.insert(ix.end(), {'B', 'A', 'I', 1});
        put32(ix, (uint32_t)n_ref);
        auto put64 = [&](uint64_t x) { for (int i = 0; i < 8; ++i) ix.push_back((uint8_t)(x >> (8 * i))); };
        for (int32_t r = 0; r < n_ref; ++r) {
            auto &bm = bins[(size_t)r];
            put32(ix, (uint32_t)(bm.size() + (n_mapped[(size_t)r] ? 1 : 0)));
            for (auto &kv : bm) {
                put32(ix, kv.first);
                put32(ix, (uint32_t)kv.second.chunks.size());
                for (auto &c : kv.second.chunks) { put64(c.first); put64(c.second); }
            }
            if (n_mapped[(size_t)r]) {      // the samtools pseudo-bin: file range and mapped / unmapped counts
                put32(ix, 37450); put32(ix, 2);
                put64(ref_beg[(size_t)r]); put64(ref_end[(size_t)r]); put64(n_mapped[(size_t)r]); put64(0);
            }
            auto &lin = linear[(size_t)r];
            for (size_t w = 1; w < lin.size(); ++w) if (lin[w] == 0) lin[w] = lin[w - 1];     // empty windows point at the previous one
            put32(ix, (uint32_t)lin.size());
            for (uint64_t v : lin) put64(v);
        }
        put64(0);       // unplaced unmapped reads
        const std::string ipath = std::string(path) + ".bai";
        FILE *g = fopen(ipath.c_str(), "wb");
        if (!g || fwrite(ix.data(), 1, ix.size(), g) != ix.size() || fclose(g) != 0) { dyn_set_error("dyn_bam_write: cannot write %s", ipath.c_str()); return DYN_ERR_IO; }
    }
    return DYN_OK;
}
