// Streaming, range-split BAM ingest (second half).
//
// Replaces what the reference does with ngs_tools / samtools in dynast/preprocessing/bam.py:609-645 (split_bam: the
// whole BAM is REWRITTEN into 8 x n_threads files) and bam.py:708-726 (one pysam reader per (split, contig)): here a
// reader takes one contiguous range of BGZF blocks of the original file -- BGZF blocks are independent gzip members,
// so a range needs no index and nothing is rewritten -- and hands it out chunk by chunk:
//   * `part` of `n_parts`: the block table is cut by compressed bytes; a part owns every alignment record that
//     STARTS inside its blocks.  htslib-style writers end blocks on record boundaries; when a part's first block
//     does not start with a record the boundary is found by walking candidate offsets until a chain of plausible
//     record headers crosses into the next block (the usual BAM split guesser);
//   * a chunk = a run of blocks of about `chunk_bytes` inflated bytes: inflated by all threads (zlib raw inflate),
//     record starts found per block by the inflating threads (or by one sequential walk when records straddle
//     blocks), records filtered / tokenised / packed by all threads over ranges of records and stitched into one
//     page-locked buffer; the tail of a record that continues in the next chunk is carried over;
//   * a producer thread keeps up to two chunks ready while the caller streams the previous one to the GPU
//     (cudaMemcpyAsync + dyn_tally_reads), so inflate, H2D and the tally overlap;
//   * with `keys` the barcode / UMI strings become 2-bit packed 64-bit keys and the gene tag an index into the
//     caller's gene list -- no string pools, nothing for Python to loop over.
// Read filter and record layout are those of bampack.cpp (bam.py:173-174, 242-248, 286-292).  Paired files keep
// the first-seen-mate table (bam.py:307-312) across chunks but cannot be range-split (n_parts must be 1).
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>
#include <unordered_map>

#include "bamhost.h"

using namespace bamhost;

namespace {

struct PinnedPool {         // page-locked buffers are expensive to create: chunks hand theirs back
    std::mutex mu;
    std::vector<std::pair<uint8_t *, size_t>> free_list;
    std::vector<std::pair<uint8_t *, bool>> all;      // (pointer, pinned)
    uint8_t *take(size_t bytes, size_t *cap) {
        {
            std::lock_guard<std::mutex> l(mu);
            for (size_t i = 0; i < free_list.size(); ++i)
                if (free_list[i].second >= bytes) {
                    uint8_t *p = free_list[i].first;
                    *cap = free_list[i].second;
                    free_list.erase(free_list.begin() + i);
                    return p;
                }
        }
        size_t want = std::max<size_t>(bytes + bytes / 8, 1 << 20);
        want = (want + 4095) & ~(size_t)4095;
        uint8_t *p = nullptr;
        bool pinned = cudaHostAlloc((void **)&p, want, cudaHostAllocPortable) == cudaSuccess;
        if (!pinned) {
            cudaGetLastError();
            if (posix_memalign((void **)&p, 4096, want) != 0) return nullptr;
        }
        std::lock_guard<std::mutex> l(mu);
        all.emplace_back(p, pinned);
        *cap = want;
        return p;
    }
    void give(uint8_t *p, size_t cap) { std::lock_guard<std::mutex> l(mu); free_list.emplace_back(p, cap); }
    ~PinnedPool() { for (auto &b : all) { if (b.second) cudaFreeHost(b.first); else free(b.first); } }
};

// plausible alignment record at p (block_size prefix included)?  *size = bytes it occupies
inline bool sane_record(const uint8_t *p, size_t avail, int32_t n_ref, size_t *size) {
    if (avail < 36) return false;
    const uint32_t bs = rd32(p);
    if (bs < 32 || bs > (1u << 26)) return false;
    const int32_t ref = (int32_t)rd32(p + 4), pos = (int32_t)rd32(p + 8), next_ref = (int32_t)rd32(p + 24), next_pos = (int32_t)rd32(p + 28);
    if (ref < -1 || ref >= n_ref || pos < -1 || next_ref < -1 || next_ref >= n_ref || next_pos < -1) return false;
    const uint32_t l_name = p[12], n_cigar = rd16(p + 16), l_seq = rd32(p + 20);
    if (l_name < 1 || l_seq > (1u << 26)) return false;
    const size_t fixed = 32 + (size_t)l_name + 4 * (size_t)n_cigar + ((size_t)l_seq + 1) / 2 + l_seq;
    if (fixed > bs) return false;
    if (36 + (size_t)l_name <= avail && p[36 + l_name - 1] != 0) return false;
    *size = 4 + (size_t)bs;
    return true;
}

// 2-bit key of an A/C/G/T string with a sentinel bit above it; an optional "-<digits>" suffix (10x GEM well) goes
// to bits 58..63.  Returns ~0 when the string does not fit (other letters, too long).
inline uint64_t acgt_key(const char *s, size_t n, bool allow_suffix) {
    uint64_t suffix = 0;
    if (allow_suffix) {
        size_t d = n;
        while (d > 0 && s[d - 1] >= '0' && s[d - 1] <= '9') --d;
        if (d < n && d > 0 && s[d - 1] == '-') {
            uint64_t v = 0;
            for (size_t i = d; i < n; ++i) v = v * 10 + (uint64_t)(s[i] - '0');
            if (v >= 63) return ~0ull;
            suffix = v + 1;
            n = d - 1;
        }
    }
    if (n > (allow_suffix ? 28u : 31u)) return ~0ull;
    uint64_t k = 1;
    for (size_t i = 0; i < n; ++i) {
        uint64_t c;
        switch (s[i]) {
            case 'A': c = 0; break;
            case 'C': c = 1; break;
            case 'G': c = 2; break;
            case 'T': c = 3; break;
            default: return ~0ull;
        }
        k = (k << 2) | c;
    }
    return k | (suffix << 58);
}

}  // namespace

struct dyn_bam_stream {
    // file
    const uint8_t *raw = nullptr;
    size_t raw_len = 0;
    std::vector<Block> blocks;            // whole file; out_off is unused here (chunks lay their blocks out themselves)
    // options
    std::string bc_tag, umi_tag, gx_tag;
    bool require_gene = false, lean = false, keys = false;
    int n_threads = 8;
    size_t chunk_bytes = 256u << 20;
    std::unordered_map<std::string, int32_t> gene_map;
    // header
    std::string header_text;
    Pool ref_name;
    std::vector<int32_t> ref_len;
    // range state
    size_t next_block = 0, end_block = 0;
    size_t first_skip = 0;                // bytes of the first block that belong to the previous part / the header
    std::vector<uint8_t> carry;           // head of a record that continues in the next chunk
    std::unordered_map<std::string, std::vector<uint8_t>> pending;     // first-seen mates (copies of their BAM records)
    int32_t cur_ref = -2;
    bool any_paired = false;
    int n_parts = 1;
    // producer
    std::thread producer;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<dyn_bam_result *> ready;
    bool finished = false, stop = false;
    int err_code = 0;
    std::string err_msg;
    std::shared_ptr<PinnedPool> pool = std::make_shared<PinnedPool>();
    int64_t records_seen = 0;
    std::vector<uint64_t> lay_start, lay_in_off, lay_in_len, lay_out_len;      // dyn_bam_stream_layout

    ~dyn_bam_stream() {
        {
            std::lock_guard<std::mutex> l(mu);
            stop = true;
        }
        cv.notify_all();
        if (producer.joinable()) producer.join();
        for (dyn_bam_result *r : ready) dyn_bam_result_free(r);
        if (raw && raw_len) munmap(const_cast<uint8_t *>(raw), raw_len);
    }
};

struct dyn_bam_chunk_extra {        // hangs off dyn_bam_result::extra for stream chunks
    std::shared_ptr<PinnedPool> pool;
    size_t blob_cap = 0;
    std::vector<uint64_t> bc_key, umi_key;
    std::vector<int32_t> gene_idx;
};

void dyn_bam_chunk_extra_free(dyn_bam_result *r) {
    dyn_bam_chunk_extra *x = static_cast<dyn_bam_chunk_extra *>(r->extra);
    if (!x) return;
    if (r->blob) x->pool->give(r->blob, x->blob_cap);
    r->blob = nullptr;
    delete x;
    r->extra = nullptr;
}

const void *dyn_bam_chunk_field(const dyn_bam_result *r, int field, size_t *nb) {
    const dyn_bam_chunk_extra *x = static_cast<const dyn_bam_chunk_extra *>(r->extra);
    if (!x) return nullptr;
    switch (field) {
        case 11: *nb = x->bc_key.size() * 8; return x->bc_key.data();
        case 12: *nb = x->umi_key.size() * 8; return x->umi_key.data();
        case 13: *nb = x->gene_idx.size() * 4; return x->gene_idx.data();
        default: return nullptr;
    }
}

namespace {

bool inflate_block(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len) {
    if (out_len == 0) return true;
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, -15) != Z_OK) return false;
    zs.next_in = const_cast<Bytef *>(in); zs.avail_in = (uInt)in_len;
    zs.next_out = out; zs.avail_out = (uInt)out_len;
    const int rc = inflate(&zs, Z_FINISH);
    inflateEnd(&zs);
    return rc == Z_STREAM_END && zs.avail_out == 0;
}

int scan_blocks(dyn_bam_stream *s, const char *path) {
    size_t ip = 0;
    while (ip + 18 <= s->raw_len) {
        const uint8_t *h = s->raw + ip;
        if (h[0] != 31 || h[1] != 139 || h[2] != 8 || !(h[3] & 4)) { dyn_set_error("dyn_bam_stream_open: %s is not BGZF", path); return DYN_ERR_FORMAT; }
        const uint32_t xlen = rd16(h + 10);
        uint32_t bsize = 0;
        bool found = false;
        for (size_t x = 12; x + 4 <= 12 + xlen && ip + x + 4 <= s->raw_len;) {
            const uint32_t slen = rd16(h + x + 2);
            if (h[x] == 'B' && h[x + 1] == 'C' && slen == 2) { bsize = rd16(h + x + 4); found = true; }
            x += 4 + slen;
        }
        if (!found || ip + bsize + 1 > s->raw_len) { dyn_set_error("dyn_bam_stream_open: corrupt BGZF block in %s", path); return DYN_ERR_FORMAT; }
        const size_t blen = (size_t)bsize + 1;
        Block b;
        b.in_off = ip + 12 + xlen; b.in_len = blen - 12 - xlen - 8; b.out_off = 0; b.out_len = rd32(h + blen - 4);
        s->blocks.push_back(b);
        ip += blen;
    }
    return DYN_OK;
}

// Inflates blocks from `b` on into `buf` until it holds `need` bytes past `skip` (or the file ends).
bool inflate_until(dyn_bam_stream *s, size_t &b, std::vector<uint8_t> &buf, size_t need) {
    while (buf.size() < need && b < s->blocks.size()) {
        const Block &bl = s->blocks[b++];
        const size_t o = buf.size();
        buf.resize(o + bl.out_len);
        if (!inflate_block(s->raw + bl.in_off, bl.in_len, buf.data() + o, bl.out_len)) return false;
    }
    return true;
}

// Header: text + reference list; *end_block / *end_off = where the first alignment record starts.
int read_header(dyn_bam_stream *s, size_t *first_block, size_t *first_off) {
    std::vector<uint8_t> buf;
    std::vector<size_t> block_start;     // uncompressed offset at which block i starts
    size_t b = 0;
    auto need = [&](size_t n) -> bool {
        while (buf.size() < n && b < s->blocks.size()) {
            block_start.push_back(buf.size());
            if (!inflate_until(s, b, buf, buf.size() + 1)) return false;
        }
        return buf.size() >= n;
    };
    if (!need(12) || memcmp(buf.data(), "BAM\1", 4) != 0) { dyn_set_error("dyn_bam_stream_open: bad BAM magic"); return DYN_ERR_FORMAT; }
    size_t p = 4;
    const uint32_t l_text = rd32(buf.data() + p); p += 4;
    if (!need(p + l_text + 4)) { dyn_set_error("dyn_bam_stream_open: truncated header"); return DYN_ERR_FORMAT; }
    s->header_text.assign((const char *)buf.data() + p, l_text); p += l_text;
    const uint32_t n_ref = rd32(buf.data() + p); p += 4;
    for (uint32_t r = 0; r < n_ref; ++r) {
        if (!need(p + 4)) { dyn_set_error("dyn_bam_stream_open: truncated reference list"); return DYN_ERR_FORMAT; }
        const uint32_t l_name = rd32(buf.data() + p); p += 4;
        if (!need(p + l_name + 4)) { dyn_set_error("dyn_bam_stream_open: truncated reference list"); return DYN_ERR_FORMAT; }
        s->ref_name.add((const char *)buf.data() + p, l_name ? l_name - 1 : 0); p += l_name;
        s->ref_len.push_back((int32_t)rd32(buf.data() + p)); p += 4;
    }
    // the record section starts at uncompressed offset p: find its block
    block_start.push_back(buf.size());
    size_t i = 0;
    while (i + 1 < block_start.size() && block_start[i + 1] <= p) ++i;
    if (i >= b) { *first_block = b; *first_off = 0; }      // the header ends exactly at the end of what was inflated
    else { *first_block = i; *first_off = p - block_start[i]; }
    return DYN_OK;
}

}  // namespace

// Offset, inside the first len0 bytes of an inflated buffer, of the first alignment record that starts there: the chain of
// plausible records from it must run to the end of the buffer (len0 when none does).  Shared with bamdevice.cu.
size_t dyn_bam_guess_offset(const uint8_t *data, size_t n, size_t len0, int32_t n_ref, bool at_eof) {
    struct View { const uint8_t *p; size_t n; size_t size() const { return n; } const uint8_t *data() const { return p; } } buf{data, n};
    for (size_t o = 0; o < len0; ++o) {
        size_t q = o, n_ok = 0;
        bool good = true;
        while (q < buf.size()) {
            size_t sz = 0;
            if (buf.size() - q < 36) { good = at_eof ? false : true; break; }
            if (!sane_record(buf.data() + q, buf.size() - q, n_ref, &sz)) { good = false; break; }
            if (q + sz > buf.size()) { good = !at_eof; break; }
            q += sz; ++n_ok;
        }
        if (q == buf.size()) good = true;
        if (good && (n_ok >= 4 || (at_eof && q == buf.size() && n_ok >= 1))) return o;
    }
    return len0;
}

namespace {

// Offset inside block `b` of the first record that starts there (a chain of plausible records must reach block b + 2).
bool guess_record_start(dyn_bam_stream *s, size_t b, size_t *off) {
    std::vector<uint8_t> buf;
    size_t bb = b;
    const size_t len0 = s->blocks[b].out_len;
    if (!inflate_until(s, bb, buf, len0 + 1)) return false;       // block b
    const size_t want = buf.size() + (2u << 16);
    if (!inflate_until(s, bb, buf, want)) return false;           // + the next ones
    const bool at_eof = bb >= s->blocks.size();
    const int32_t n_ref = (int32_t)s->ref_len.size();
    *off = dyn_bam_guess_offset(buf.data(), buf.size(), len0, n_ref, at_eof);
    return true;
}

struct Part {                          // what one packing thread produces for its record range
    std::vector<uint8_t> blob;
    std::vector<uint32_t> size16;
    std::vector<int32_t> ref_id, pos, hi, score, gene_idx;
    std::vector<uint16_t> flag;
    std::vector<uint8_t> gx_assigned;
    std::vector<uint64_t> bc_key, umi_key;
    Pool name, barcode, umi, gene;
    std::string error;
    int err_code = 0;
};

struct Packer {
    dyn_bam_stream *s;
    const char *bc, *umi, *gx;
    std::vector<uint32_t> toks;
    explicit Packer(dyn_bam_stream *st) : s(st) {
        bc = s->bc_tag.empty() ? nullptr : s->bc_tag.c_str();
        umi = s->umi_tag.empty() ? nullptr : s->umi_tag.c_str();
        gx = s->gx_tag.empty() ? nullptr : s->gx_tag.c_str();
    }
    // 1 = unit emitted, 0 = filtered / parked, < 0 = error
    int filter(const uint8_t *rec, uint32_t block_size, Aux &a, Part &pt) {
        const int32_t ref_id = (int32_t)rd32(rec);
        const uint32_t l_read_name = rec[8], n_cigar = rd16(rec + 12), l_seq = rd32(rec + 16);
        const uint16_t flag = rd16(rec + 14);
        if (ref_id < 0) return 0;
        if (flag & (0x100 | 0x4 | 0x400)) return 0;                      // bam.py:173
        const size_t fixed = 32 + l_read_name + 4 * (size_t)n_cigar + (l_seq + 1) / 2 + l_seq;
        if (fixed > block_size) { pt.err_code = DYN_ERR_FORMAT; pt.error = "dyn_bam_stream: corrupt record"; return -1; }
        if (!parse_aux(rec + fixed, rec + block_size, bc, umi, gx, a)) { pt.err_code = DYN_ERR_FORMAT; pt.error = "dyn_bam_stream: corrupt aux data"; return -1; }
        if ((bc && !a.has_bc) || (umi && !a.has_umi) || (s->require_gene && gx && !a.has_gx)) return 0;   // bam.py:174, 242-248
        if (bc && a.bc_len == 1 && a.bc[0] == '-') return 0;                                                // bam.py:290
        if (!a.has_hi || !a.has_as) { pt.err_code = DYN_ERR_FORMAT; pt.error = "dyn_bam_stream: alignment without HI / AS tag (KeyError in the reference, bam.py:295-296)"; return -1; }
        if (!a.has_md) { pt.err_code = DYN_ERR_FORMAT; pt.error = "dyn_bam_stream: alignment without MD tag (count.py:119-124)"; return -1; }
        if (l_seq > 0xffff || n_cigar > 0xffff) { pt.err_code = DYN_ERR_FORMAT; pt.error = "dyn_bam_stream: record too long for the packed layout"; return -1; }
        return 1;
    }
    void emit(const uint8_t *rec, const Aux &a, const uint8_t *mate_rec, const Aux *ma, Part &pt) {
        const uint32_t l_read_name = rec[8];
        const size_t before = pt.blob.size();
        pack_record(rec, a, mate_rec ? DYN_REC_HAS_MATE : 0, toks, pt.blob, s->lean && !mate_rec);
        if (mate_rec) pack_record(mate_rec, *ma, 0, toks, pt.blob, false);
        pt.size16.push_back((uint32_t)((pt.blob.size() - before) >> 4));
        pt.ref_id.push_back((int32_t)rd32(rec));
        pt.pos.push_back((int32_t)rd32(rec + 4));
        pt.flag.push_back(rd16(rec + 14));
        pt.hi.push_back((int32_t)a.hi);
        pt.score.push_back((int32_t)a.as);
        pt.gx_assigned.push_back(a.has_gx ? 1 : 0);
        if (s->keys) {
            pt.bc_key.push_back(a.has_bc ? acgt_key(a.bc, a.bc_len, true) : 0ull);
            pt.umi_key.push_back(a.has_umi ? acgt_key(a.umi, a.umi_len, false) : 0ull);
            int32_t gi = -1;
            if (a.has_gx) {
                auto it = s->gene_map.find(std::string(a.gx, a.gx_len));
                gi = it == s->gene_map.end() ? -2 : it->second;
            }
            pt.gene_idx.push_back(gi);
        } else {
            pt.name.add((const char *)rec + 32, l_read_name ? l_read_name - 1 : 0);
            pt.barcode.add(a.has_bc ? a.bc : "NA", a.has_bc ? a.bc_len : 2);
            pt.umi.add(a.has_umi ? a.umi : "NA", a.has_umi ? a.umi_len : 2);
            pt.gene.add(a.has_gx ? a.gx : "", a.has_gx ? a.gx_len : 0);
        }
    }
};

// One chunk: blocks [b0, b1) (+ following blocks while the last record that starts before the range end is
// incomplete, when `last`).  Returns nullptr with s->err_code set on failure.
dyn_bam_result *make_chunk(dyn_bam_stream *s, size_t b0, size_t b1, bool last) {
    PhaseTimer timer;
    const int nt = std::max(1, std::min(s->n_threads, 64));
    // ---- layout: [carry][block b0 (minus first_skip)] ... [block b1-1]
    const size_t skip = s->first_skip;
    s->first_skip = 0;
    std::vector<size_t> at(b1 - b0 + 1);
    size_t total = s->carry.size();
    for (size_t i = b0; i < b1; ++i) { at[i - b0] = total; total += s->blocks[i].out_len; }
    at[b1 - b0] = total;
    const size_t range_end = total;                 // records starting at or after this offset belong to the next part
    std::unique_ptr<uint8_t[]> data_buf(new uint8_t[std::max<size_t>(total, 1) + 16]);
    uint8_t *data = data_buf.get();
    if (!s->carry.empty()) memcpy(data, s->carry.data(), s->carry.size());
    const bool aligned = s->carry.empty() && skip == 0;      // per-block walks may serve as the record chain
    struct BlockRecs { std::vector<uint32_t> at; bool ok = false, paired = false; };
    std::vector<BlockRecs> brecs(b1 - b0);
    std::atomic<size_t> next(0);
    std::atomic<int> failed(0);
    auto inflate_work = [&]() {
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= b1 - b0) return;
            const Block &b = s->blocks[b0 + i];
            uint8_t *bp = data + at[i];
            if (!inflate_block(s->raw + b.in_off, b.in_len, bp, b.out_len)) { failed = 1; return; }
            BlockRecs &br = brecs[i];
            size_t q = 0;
            while (q + 4 <= b.out_len) {
                const uint32_t bs = rd32(bp + q);
                if (bs < 32 || q + 4 + (size_t)bs > b.out_len) break;
                br.paired |= (rd16(bp + q + 4 + 14) & 0x1) != 0;
                br.at.push_back((uint32_t)q);
                q += 4 + (size_t)bs;
            }
            br.ok = q == b.out_len;
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 0; t < std::min<size_t>(nt, b1 - b0); ++t) th.emplace_back(inflate_work);
        for (auto &t : th) t.join();
    }
    if (failed.load()) { s->err_code = DYN_ERR_FORMAT; s->err_msg = "dyn_bam_stream: inflate failed"; return nullptr; }
    timer.mark("inflate");

    // ---- record chain
    std::vector<size_t> rec_at;
    rec_at.reserve(total / 280 + 16);
    bool blockwise = aligned;
    for (size_t i = 0; blockwise && i < brecs.size(); ++i) blockwise = brecs[i].ok;
    size_t q = s->carry.empty() ? skip : 0;
    std::vector<uint8_t> tail_extra;       // the last record completed from the blocks behind the range (`last` only)
    if (blockwise) {
        for (size_t i = 0; i < brecs.size(); ++i) {
            s->any_paired |= brecs[i].paired;
            for (const uint32_t o : brecs[i].at) rec_at.push_back(at[i] + o);
        }
        q = total;
    } else {
        while (q + 4 <= total) {
            const uint32_t bs = rd32(data + q);
            if (bs < 32) { s->err_code = DYN_ERR_FORMAT; s->err_msg = "dyn_bam_stream: corrupt record chain"; return nullptr; }
            if (q + 4 + (size_t)bs > total) break;
            s->any_paired |= (rd16(data + q + 4 + 14) & 0x1) != 0;
            rec_at.push_back(q);
            q += 4 + (size_t)bs;
        }
    }
    std::vector<BlockRecs>().swap(brecs);
    if (s->any_paired && s->n_parts > 1) {
        s->err_code = DYN_ERR_ARG;
        s->err_msg = "dyn_bam_stream: a BAM with paired records cannot be range-split (n_parts must be 1)";
        return nullptr;
    }
    // what is left is the head of a record that continues behind this chunk
    s->carry.assign(data + q, data + total);
    if (last && !s->carry.empty()) {
        // the part owns the record that starts before its end: complete it from the following blocks
        size_t bb = b1;
        tail_extra = s->carry;
        s->carry.clear();
        for (;;) {
            if (tail_extra.size() >= 4) {
                const size_t need = 4 + (size_t)rd32(tail_extra.data());
                if (tail_extra.size() >= need) { tail_extra.resize(need); break; }
            }
            if (bb >= s->blocks.size()) { tail_extra.clear(); break; }      // truncated file: drop the fragment
            if (!inflate_until(s, bb, tail_extra, tail_extra.size() + 1)) { s->err_code = DYN_ERR_FORMAT; s->err_msg = "dyn_bam_stream: inflate failed"; return nullptr; }
        }
    }
    (void)range_end;
    timer.mark("walk");

    // ---- pack
    dyn_bam_result *res = new dyn_bam_result();
    dyn_bam_chunk_extra *extra = new dyn_bam_chunk_extra();
    extra->pool = s->pool;
    res->extra = extra;
    const size_t n_rec = rec_at.size() + (tail_extra.empty() ? 0 : 1);
    res->n_records_seen = (int64_t)n_rec;
    auto rec_ptr = [&](size_t ri) -> const uint8_t * { return ri < rec_at.size() ? data + rec_at[ri] : tail_extra.data(); };
    std::vector<Part> parts;
    if (!s->any_paired) {
        const size_t n_parts = std::max<size_t>(1, std::min<size_t>((size_t)nt * 4, n_rec / 512 + 1));
        parts.resize(n_parts);
        std::atomic<size_t> next_part(0);
        auto work = [&]() {
            Packer pk(s);
            for (;;) {
                const size_t pi = next_part.fetch_add(1);
                if (pi >= n_parts) return;
                Part &pt = parts[pi];
                const size_t r0 = n_rec * pi / n_parts, r1 = n_rec * (pi + 1) / n_parts;
                pt.blob.reserve((r1 - r0) * (s->lean ? 96 : 200));
                for (size_t ri = r0; ri < r1; ++ri) {
                    const uint8_t *p = rec_ptr(ri);
                    Aux a;
                    const int k = pk.filter(p + 4, rd32(p), a, pt);
                    if (k < 0) return;
                    if (k == 0) continue;
                    if (rd16(p + 4 + 14) & 0x1) { pt.err_code = DYN_ERR_FORMAT; pt.error = "dyn_bam_stream: paired record in a file taken for single-end"; return; }
                    pk.emit(p + 4, a, nullptr, nullptr, pt);
                }
            }
        };
        std::vector<std::thread> th;
        for (int t = 0; t < std::min<size_t>(nt, n_parts); ++t) th.emplace_back(work);
        for (auto &t : th) t.join();
    } else {
        // paired files: one sequential pass, the first-seen mate parked under (read name, HI) (bam.py:307-312)
        parts.resize(1);
        Part &pt = parts[0];
        Packer pk(s);
        std::string key;
        for (size_t ri = 0; ri < n_rec; ++ri) {
            const uint8_t *p = rec_ptr(ri);
            const uint8_t *rec = p + 4;
            const uint32_t block_size = rd32(p);
            const int32_t ref_id = (int32_t)rd32(rec);
            if (ref_id != s->cur_ref) { s->pending.clear(); s->cur_ref = ref_id; }     // parse_read_contig runs per contig
            Aux a;
            const int k = pk.filter(rec, block_size, a, pt);
            if (k < 0) break;
            if (k == 0) continue;
            const uint16_t flag = rd16(rec + 14);
            if (!(flag & 0x1)) { pk.emit(rec, a, nullptr, nullptr, pt); continue; }
            const uint32_t l_read_name = rec[8];
            key.assign((const char *)rec + 32, l_read_name ? l_read_name - 1 : 0);
            key.push_back('\0');
            key.append(std::to_string(a.hi));
            auto it = s->pending.find(key);
            if (it == s->pending.end()) { s->pending.emplace(key, std::vector<uint8_t>(p, p + 4 + block_size)); continue; }
            const std::vector<uint8_t> mate = std::move(it->second);
            s->pending.erase(it);
            const uint8_t *mrec = mate.data() + 4;
            const uint32_t m_name = mrec[8], m_cig = rd16(mrec + 12), m_seq = rd32(mrec + 16);
            Aux ma;
            parse_aux(mrec + 32 + m_name + 4 * (size_t)m_cig + (m_seq + 1) / 2 + m_seq, mrec + rd32(mate.data()), pk.bc, pk.umi, pk.gx, ma);
            if (!ma.has_md) { pt.err_code = DYN_ERR_FORMAT; pt.error = "dyn_bam_stream: alignment without MD tag (count.py:119-124)"; break; }
            pk.emit(rec, a, mrec, &ma, pt);
        }
    }
    timer.mark("pack");
    for (const Part &pt : parts)
        if (pt.err_code) { s->err_code = pt.err_code; s->err_msg = pt.error; dyn_bam_result_free(res); return nullptr; }

    // ---- stitch into one page-locked blob
    const size_t n_parts = parts.size();
    size_t units = 0, bytes = 0;
    std::vector<size_t> unit0(n_parts), byte0(n_parts);
    for (size_t i = 0; i < n_parts; ++i) { unit0[i] = units; byte0[i] = bytes; units += parts[i].size16.size(); bytes += parts[i].blob.size(); }
    if ((bytes >> 4) > 0xfffffff0ull) { s->err_code = DYN_ERR_NOMEM; s->err_msg = "dyn_bam_stream: chunk over 64 GiB"; dyn_bam_result_free(res); return nullptr; }
    res->blob_bytes = bytes;
    res->blob = s->pool->take(std::max<size_t>(bytes, 16), &extra->blob_cap);
    if (!res->blob) { s->err_code = DYN_ERR_NOMEM; s->err_msg = "dyn_bam_stream: out of memory"; dyn_bam_result_free(res); return nullptr; }
    res->blob_pinned = true;      // owned by the pool either way
    res->rec_off.assign(units + 1, 0);
    res->ref_id.resize(units); res->pos.resize(units); res->hi.resize(units); res->score.resize(units);
    res->flag.resize(units); res->gx_assigned.resize(units); res->paired.assign(units, 0);
    if (s->keys) { extra->bc_key.resize(units); extra->umi_key.resize(units); extra->gene_idx.resize(units); }
    Pool *dst_pool[4] = {&res->name, &res->barcode, &res->umi, &res->gene};
    std::vector<size_t> pool0[4];
    for (int k = 0; k < 4; ++k) {
        pool0[k].resize(n_parts);
        size_t tot = 0;
        for (size_t i = 0; i < n_parts; ++i) {
            const Pool &sp = k == 0 ? parts[i].name : (k == 1 ? parts[i].barcode : (k == 2 ? parts[i].umi : parts[i].gene));
            pool0[k][i] = tot;
            tot += sp.data.size();
        }
        dst_pool[k]->data.resize(tot);
        dst_pool[k]->off.assign((s->keys ? 0 : units) + 1, 0);
        dst_pool[k]->off.back() = (uint32_t)tot;
    }
    std::atomic<size_t> next_copy(0);
    auto copy = [&]() {
        for (;;) {
            const size_t i = next_copy.fetch_add(1);
            if (i >= n_parts) return;
            const Part &pt = parts[i];
            const size_t u0 = unit0[i], nu = pt.size16.size();
            if (!pt.blob.empty()) memcpy(res->blob + byte0[i], pt.blob.data(), pt.blob.size());
            uint32_t o16 = (uint32_t)(byte0[i] >> 4);
            for (size_t j = 0; j < nu; ++j) { res->rec_off[u0 + j] = o16; o16 += pt.size16[j]; }
            if (!nu) continue;
            memcpy(res->ref_id.data() + u0, pt.ref_id.data(), 4 * nu);
            memcpy(res->pos.data() + u0, pt.pos.data(), 4 * nu);
            memcpy(res->hi.data() + u0, pt.hi.data(), 4 * nu);
            memcpy(res->score.data() + u0, pt.score.data(), 4 * nu);
            memcpy(res->flag.data() + u0, pt.flag.data(), 2 * nu);
            memcpy(res->gx_assigned.data() + u0, pt.gx_assigned.data(), nu);
            if (s->keys) {
                memcpy(extra->bc_key.data() + u0, pt.bc_key.data(), 8 * nu);
                memcpy(extra->umi_key.data() + u0, pt.umi_key.data(), 8 * nu);
                memcpy(extra->gene_idx.data() + u0, pt.gene_idx.data(), 4 * nu);
            } else {
                for (int k = 0; k < 4; ++k) {
                    const Pool &sp = k == 0 ? pt.name : (k == 1 ? pt.barcode : (k == 2 ? pt.umi : pt.gene));
                    if (!sp.data.empty()) memcpy(dst_pool[k]->data.data() + pool0[k][i], sp.data.data(), sp.data.size());
                    for (size_t j = 0; j < nu; ++j) dst_pool[k]->off[u0 + j] = (uint32_t)(pool0[k][i] + sp.off[j]);
                }
            }
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 0; t < std::min<size_t>(nt, n_parts); ++t) th.emplace_back(copy);
        for (auto &t : th) t.join();
    }
    res->rec_off[units] = (uint32_t)(bytes >> 4);
    if (s->any_paired) {
        // which units carry a mate: the HAS_MATE bit of the unit's first record
        for (size_t u = 0; u < units; ++u) res->paired[u] = (res->blob[((size_t)res->rec_off[u] << 4) + 15] & 0x80) ? 1 : 0;
    }
    timer.mark("stitch");
    return res;
}

void produce(dyn_bam_stream *s) {
    while (true) {
        {
            std::unique_lock<std::mutex> l(s->mu);
            s->cv.wait(l, [&] { return s->stop || s->ready.size() < 2; });
            if (s->stop) return;
        }
        if (s->next_block >= s->end_block) break;
        size_t b1 = s->next_block, acc = 0;
        while (b1 < s->end_block && (b1 == s->next_block || acc + s->blocks[b1].out_len <= s->chunk_bytes)) acc += s->blocks[b1++].out_len;
        const bool last = b1 == s->end_block;
        dyn_bam_result *r = make_chunk(s, s->next_block, b1, last);
        s->next_block = b1;
        std::lock_guard<std::mutex> l(s->mu);
        if (!r) { s->finished = true; s->cv.notify_all(); return; }
        s->records_seen += r->n_records_seen;
        s->ready.push_back(r);
        s->cv.notify_all();
    }
    std::lock_guard<std::mutex> l(s->mu);
    s->finished = true;
    s->cv.notify_all();
}

}  // namespace

static int stream_create(const char *path, const dyn_bam_stream_opts_t *o, bool start, dyn_bam_stream_t **out) {
    if (!path || !o || !out) { dyn_set_error("dyn_bam_stream_open: null argument"); return DYN_ERR_ARG; }
    if (o->n_parts < 1 || o->part < 0 || o->part >= o->n_parts) { dyn_set_error("dyn_bam_stream_open: part %d of %d", o->part, o->n_parts); return DYN_ERR_ARG; }
    std::unique_ptr<dyn_bam_stream> s(new dyn_bam_stream());
    if (o->barcode_tag && o->barcode_tag[0]) s->bc_tag = o->barcode_tag;
    if (o->umi_tag && o->umi_tag[0]) s->umi_tag = o->umi_tag;
    if (o->gene_tag && o->gene_tag[0]) s->gx_tag = o->gene_tag;
    s->require_gene = o->require_gene != 0;
    s->lean = o->lean != 0;
    s->keys = o->keys != 0;
    s->n_threads = o->n_threads > 0 ? o->n_threads : 8;
    s->n_parts = o->n_parts;
    if (o->chunk_bytes > 0) s->chunk_bytes = (size_t)o->chunk_bytes;
    for (int64_t i = 0; i < o->n_gene_ids; ++i) s->gene_map.emplace(o->gene_ids[i], (int32_t)i);
    const int fd = open(path, O_RDONLY);
    if (fd < 0) { dyn_set_error("dyn_bam_stream_open: cannot open %s", path); return DYN_ERR_IO; }
    struct stat sb;
    if (fstat(fd, &sb) != 0) { close(fd); dyn_set_error("dyn_bam_stream_open: cannot stat %s", path); return DYN_ERR_IO; }
    s->raw_len = (size_t)sb.st_size;
    if (s->raw_len) {
        void *m = mmap(nullptr, s->raw_len, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) { close(fd); s->raw_len = 0; dyn_set_error("dyn_bam_stream_open: cannot map %s", path); return DYN_ERR_IO; }
        s->raw = static_cast<const uint8_t *>(m);
    }
    close(fd);
    int rc = scan_blocks(s.get(), path);
    if (rc) return rc;
    size_t first_block = 0, first_off = 0;
    if ((rc = read_header(s.get(), &first_block, &first_off))) return rc;
    // ---- this part's block range: the record section cut by compressed bytes
    const size_t nb = s->blocks.size();
    const size_t c0 = first_block < nb ? s->blocks[first_block].in_off : s->raw_len;
    const size_t span = s->raw_len - c0;
    auto block_at = [&](int part) -> size_t {
        if (part <= 0) return first_block;
        if (part >= o->n_parts) return nb;
        const size_t target = c0 + (size_t)((double)span * part / o->n_parts);
        size_t lo = first_block, hi = nb;
        while (lo < hi) { const size_t mid = (lo + hi) >> 1; if (s->blocks[mid].in_off < target) lo = mid + 1; else hi = mid; }
        return lo;
    };
    s->next_block = block_at(o->part);
    s->end_block = block_at(o->part + 1);
    if (o->part == 0) s->first_skip = first_off;
    else {
        // the first record that STARTS in this part (blocks in which none starts belong to the previous part's last record)
        while (s->next_block < s->end_block) {
            size_t off = 0;
            if (!guess_record_start(s.get(), s->next_block, &off)) { dyn_set_error("dyn_bam_stream_open: inflate failed in %s", path); return DYN_ERR_FORMAT; }
            if (off < s->blocks[s->next_block].out_len) { s->first_skip = off; break; }
            ++s->next_block;
        }
    }
    dyn_bam_stream *sp = s.release();
    if (start) sp->producer = std::thread(produce, sp);
    *out = sp;
    return DYN_OK;
}

extern "C" int dyn_bam_stream_open(const char *path, const dyn_bam_stream_opts_t *o, dyn_bam_stream_t **out) {
    return stream_create(path, o, true, out);
}

// The block range of a part without reading it (the device ingest, bamdevice.cu, shares the block table, the header
// parse and the part-boundary guesser): per block its file offset, the offset / length of its deflate payload and its
// inflated size; *first_skip = bytes of the first block that precede the part's first record.  The arrays live in
// *holder (close it with dyn_bam_stream_close).
extern "C" int dyn_bam_stream_layout(const char *path, int part, int n_parts, int64_t *n_blocks, const uint64_t **starts,
                                     const uint64_t **in_offs, const uint64_t **in_lens, const uint64_t **out_lens, uint32_t *first_skip,
                                     dyn_bam_stream_t **holder) {
    dyn_bam_stream_opts_t o;
    memset(&o, 0, sizeof(o));
    o.part = part; o.n_parts = n_parts; o.n_threads = 1;
    dyn_bam_stream_t *s = nullptr;
    const int rc = stream_create(path, &o, false, &s);
    if (rc) return rc;
    for (size_t b = s->next_block; b < s->end_block; ++b) {
        const Block &bl = s->blocks[b];
        // the block starts 12 + xlen bytes before its payload: recover it from the previous block's end
        s->lay_in_off.push_back(bl.in_off);
        s->lay_in_len.push_back(bl.in_len);
        s->lay_out_len.push_back(bl.out_len);
        s->lay_start.push_back(b == 0 ? 0 : s->blocks[b - 1].in_off + s->blocks[b - 1].in_len + 8);
    }
    *n_blocks = (int64_t)s->lay_start.size();
    *starts = s->lay_start.data(); *in_offs = s->lay_in_off.data(); *in_lens = s->lay_in_len.data(); *out_lens = s->lay_out_len.data();
    *first_skip = (uint32_t)s->first_skip;
    *holder = s;
    return DYN_OK;
}

extern "C" int dyn_bam_stream_next(dyn_bam_stream_t *s, dyn_bam_result_t **chunk) {
    if (!s || !chunk) { dyn_set_error("dyn_bam_stream_next: null argument"); return DYN_ERR_ARG; }
    *chunk = nullptr;
    std::unique_lock<std::mutex> l(s->mu);
    s->cv.wait(l, [&] { return !s->ready.empty() || s->finished; });
    if (!s->ready.empty()) {
        *chunk = s->ready.front();
        s->ready.pop_front();
        s->cv.notify_all();
        return DYN_OK;
    }
    if (s->err_code) { dyn_set_error("%s", s->err_msg.c_str()); return s->err_code; }
    return DYN_OK;
}

extern "C" void dyn_bam_stream_close(dyn_bam_stream_t *s) { delete s; }

extern "C" int64_t dyn_bam_stream_n_refs(const dyn_bam_stream_t *s) { return s ? (int64_t)s->ref_len.size() : 0; }

// field: 9 ref_len (int32), 10 header text, 28 / 29 reference names (chars / uint32 offsets)
extern "C" const void *dyn_bam_stream_field(const dyn_bam_stream_t *s, int field, int64_t *n_bytes) {
    if (!s) return nullptr;
    const void *p = nullptr;
    size_t nb = 0;
    switch (field) {
        case 9: p = s->ref_len.data(); nb = s->ref_len.size() * 4; break;
        case 10: p = s->header_text.data(); nb = s->header_text.size(); break;
        case 28: p = s->ref_name.data.data(); nb = s->ref_name.data.size(); break;
        case 29: p = s->ref_name.off.data(); nb = s->ref_name.off.size() * 4; break;
        default: return nullptr;
    }
    if (n_bytes) *n_bytes = (int64_t)nb;
    return p;
}
