// Host BAM ingest: BGZF inflate + record packing (SURVEY.md section 8 row a1/a2 host side, f1).
//
// Replaces what the reference gets from pysam/htslib in dynast/preprocessing/bam.py:253-312 (iterate a
// coordinate-sorted BAM, apply the read filter, pair mates) and writes the packed records of
// include/dynast_b200.h that the tally / coverage / consensus kernels read, plus the per-unit metadata the
// CSV outputs need (read name, HI, AS, barcode / UMI / gene strings).
//   * BGZF blocks are independent gzip members: their boundaries are found from the BSIZE extra field and
//     they are inflated in parallel (std::thread + zlib raw inflate) into one contiguous buffer;
//   * records are then parsed in file order; the read filter is bam.py:173-174, 242-248, 286-292
//     (secondary / unmapped / duplicate dropped, supplementary kept, required tags, barcode '-'); the first
//     seen mate of a pair is parked under (read name, HI) and packed behind its mate (bam.py:307-312); the
//     pending table is reset at every contig change, as parse_read_contig runs once per contig.
// Integer / byte work on the host; no GPU involvement. Output buffers are page-locked when a CUDA context is
// available so that dyn_tally_reads_host can stream them with cudaMemcpyAsync.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <memory>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/dynast_b200.h"

void dyn_set_error(const char *fmt, ...);

#include "bamhost.h"

using namespace bamhost;

extern "C" int dyn_bam_pack(const char *path, const char *barcode_tag, const char *umi_tag, const char *gene_tag,
                            int require_gene, int n_threads, dyn_bam_result_t **out_result) {
    return dyn_bam_pack_ex(path, barcode_tag, umi_tag, gene_tag, require_gene, n_threads, 0u, out_result);
}

extern "C" int dyn_bam_pack_ex(const char *path, const char *barcode_tag, const char *umi_tag, const char *gene_tag,
                               int require_gene, int n_threads, uint32_t flags, dyn_bam_result_t **out_result) {
    if (!path || !out_result) { dyn_set_error("dyn_bam_pack: null argument"); return DYN_ERR_ARG; }
    const bool keep_raw = (flags & DYN_BAM_KEEP_RAW) != 0;
    const uint32_t skip_flags = 0x100u | 0x4u | ((flags & DYN_BAM_KEEP_DUPLICATES) ? 0u : 0x400u);
    if (barcode_tag && !barcode_tag[0]) barcode_tag = nullptr;
    if (umi_tag && !umi_tag[0]) umi_tag = nullptr;
    if (gene_tag && !gene_tag[0]) gene_tag = nullptr;
    PhaseTimer timer;
    // The compressed file is mapped, not read: the block-table scan touches one page per block and the inflating
    // threads fault the rest in as they go (a copy through fread was a quarter of the ingest time on a 16-core host).
    struct FileMap {
        const uint8_t *p = nullptr;
        size_t n = 0;
        std::vector<uint8_t> fallback;
        ~FileMap() { if (p && fallback.empty() && n) munmap(const_cast<uint8_t *>(p), n); }
        const uint8_t *data() const { return p; }
        size_t size() const { return n; }
    } raw;
    {
        const int fd = open(path, O_RDONLY);
        if (fd < 0) { dyn_set_error("dyn_bam_pack: cannot open %s", path); return DYN_ERR_IO; }
        struct stat sb;
        if (fstat(fd, &sb) != 0) { close(fd); dyn_set_error("dyn_bam_pack: cannot stat %s", path); return DYN_ERR_IO; }
        raw.n = (size_t)sb.st_size;
        if (raw.n > 0) {
            void *m = mmap(nullptr, raw.n, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m != MAP_FAILED) {
                raw.p = static_cast<const uint8_t *>(m);
                madvise(m, raw.n, MADV_SEQUENTIAL);
            } else {        // not mappable (pipe, odd file system): read it
                raw.fallback.resize(raw.n);
                size_t got = 0;
                while (got < raw.n) {
                    const ssize_t r = read(fd, raw.fallback.data() + got, raw.n - got);
                    if (r <= 0) break;
                    got += (size_t)r;
                }
                if (got != raw.n) { close(fd); dyn_set_error("dyn_bam_pack: short read on %s", path); return DYN_ERR_IO; }
                raw.p = raw.fallback.data();
            }
        }
        close(fd);
    }

    timer.mark("read");
    // ---- BGZF block table
    std::vector<Block> blocks;
    size_t ip = 0, total_out = 0;
    while (ip + 18 <= raw.size()) {
        const uint8_t *h = raw.data() + ip;
        if (h[0] != 31 || h[1] != 139 || h[2] != 8 || !(h[3] & 4)) { dyn_set_error("dyn_bam_pack: %s is not BGZF", path); return DYN_ERR_FORMAT; }
        const uint32_t xlen = rd16(h + 10);
        uint32_t bsize = 0;
        bool found = false;
        for (size_t x = 12; x + 4 <= 12 + xlen && ip + x + 4 <= raw.size();) {
            const uint32_t slen = rd16(h + x + 2);
            if (h[x] == 'B' && h[x + 1] == 'C' && slen == 2) { bsize = rd16(h + x + 4); found = true; }
            x += 4 + slen;
        }
        if (!found || ip + bsize + 1 > raw.size()) { dyn_set_error("dyn_bam_pack: corrupt BGZF block in %s", path); return DYN_ERR_FORMAT; }
        const size_t blen = (size_t)bsize + 1;
        const uint32_t isize = rd32(h + blen - 4);
        Block b;
        b.in_off = ip + 12 + xlen; b.in_len = blen - 12 - xlen - 8; b.out_off = total_out; b.out_len = isize;
        blocks.push_back(b);
        total_out += isize;
        ip += blen;
    }
    // The inflated stream is NOT zero-initialised (first touch happens in the inflating threads), and the main thread
    // walks the header and the chain of record sizes behind the inflating threads instead of after them: every step
    // of that walk depends on the previous one (~100 ns of memory latency per record), so it is the one serial part.
    std::unique_ptr<uint8_t[]> data_buf(new uint8_t[std::max<size_t>(total_out, 1)]);
    uint8_t *const data = data_buf.get();
    const uint8_t *d = data;
    const size_t n = total_out;
    std::unique_ptr<std::atomic<uint8_t>[]> done(new std::atomic<uint8_t>[blocks.size() + 1]);
    for (size_t i = 0; i <= blocks.size(); ++i) done[i].store(0, std::memory_order_relaxed);
    struct BlockRecs { std::vector<uint32_t> at; bool ok = false, paired = false; };
    std::vector<BlockRecs> block_recs(blocks.size());
    std::atomic<size_t> next(0);
    std::atomic<int> failed(0);
    auto inflate_work = [&]() {
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= blocks.size()) return;
            const Block &b = blocks[i];
            if (b.out_len != 0) {
                z_stream zs;
                memset(&zs, 0, sizeof(zs));
                if (inflateInit2(&zs, -15) != Z_OK) { failed = 1; done[i].store(1, std::memory_order_release); return; }
                zs.next_in = const_cast<Bytef *>(raw.data() + b.in_off); zs.avail_in = (uInt)b.in_len;
                zs.next_out = data + b.out_off; zs.avail_out = (uInt)b.out_len;
                const int rc = inflate(&zs, Z_FINISH);
                inflateEnd(&zs);
                if (rc != Z_STREAM_END || zs.avail_out != 0) failed = 1;
            }
            // htslib-style writers (STAR included) never split a record over BGZF blocks: walk the block's records
            // right away, while it is hot in this core's cache, assuming it starts at a record boundary.  The
            // assumption is checked by induction further down; a block that does not end on a record boundary
            // sends the whole file to the sequential walk.
            {
                BlockRecs &br = block_recs[i];
                const uint8_t *bp = data + b.out_off;
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
            done[i].store(1, std::memory_order_release);
        }
    };
    const int nt_inflate = std::max(1, std::min(n_threads, 64));
    std::vector<std::thread> inflaters;
    for (int t = 0; t < nt_inflate; ++t) inflaters.emplace_back(inflate_work);
    size_t ready_blocks = 0, ready_bytes = 0;
    // blocks until the first `need` bytes of the stream are inflated (or everything is); false when inflate failed
    auto wait_ready = [&](size_t need) -> bool {
        need = std::min(need, total_out);
        while (ready_bytes < need) {
            bool advanced = false;
            while (ready_blocks < blocks.size() && done[ready_blocks].load(std::memory_order_acquire)) {
                ready_bytes = blocks[ready_blocks].out_off + blocks[ready_blocks].out_len;
                ++ready_blocks;
                advanced = true;
            }
            if (failed.load()) return false;
            if (!advanced && ready_bytes < need) std::this_thread::sleep_for(std::chrono::microseconds(20));
        }
        return !failed.load();
    };
    auto join_inflaters = [&]() { for (auto &t : inflaters) if (t.joinable()) t.join(); };
    auto bail = [&](dyn_bam_result *r, int code) { join_inflaters(); delete r; return code; };

    // ---- header
    dyn_bam_result *res = new dyn_bam_result();
    if (!wait_ready(12)) { dyn_set_error("dyn_bam_pack: inflate failed in %s", path); return bail(res, DYN_ERR_FORMAT); }
    if (n < 12 || memcmp(d, "BAM\1", 4) != 0) { dyn_set_error("dyn_bam_pack: bad BAM magic in %s", path); return bail(res, DYN_ERR_FORMAT); }
    size_t p = 4;
    const uint32_t l_text = rd32(d + p); p += 4;
    if (p + l_text + 4 > n) { dyn_set_error("dyn_bam_pack: truncated header"); return bail(res, DYN_ERR_FORMAT); }
    if (!wait_ready(p + l_text + 4)) { dyn_set_error("dyn_bam_pack: inflate failed in %s", path); return bail(res, DYN_ERR_FORMAT); }
    res->header_text.assign((const char *)d + p, l_text); p += l_text;
    const uint32_t n_ref = rd32(d + p); p += 4;
    for (uint32_t r = 0; r < n_ref; ++r) {
        if (p + 4 > n) { dyn_set_error("dyn_bam_pack: truncated reference list"); return bail(res, DYN_ERR_FORMAT); }
        if (!wait_ready(p + 4)) { dyn_set_error("dyn_bam_pack: inflate failed in %s", path); return bail(res, DYN_ERR_FORMAT); }
        const uint32_t l_name = rd32(d + p); p += 4;
        if (p + l_name + 4 > n) { dyn_set_error("dyn_bam_pack: truncated reference list"); return bail(res, DYN_ERR_FORMAT); }
        if (!wait_ready(p + l_name + 4)) { dyn_set_error("dyn_bam_pack: inflate failed in %s", path); return bail(res, DYN_ERR_FORMAT); }
        res->ref_name.add((const char *)d + p, l_name ? l_name - 1 : 0); p += l_name;
        res->ref_len.push_back((int32_t)rd32(d + p)); p += 4;
    }

    // ---- records.  Files without paired records are filtered, tokenised and packed by all threads over ranges of record
    //      indices and stitched together in order.  Paired files keep the sequential loop below: mate pairing
    //      (bam.py:307-312) is a first-seen dictionary over the record order.
    {
        std::vector<size_t> rec_at;
        rec_at.reserve(n / 300);
        bool any_paired = false, chain_ok = true;
        // (a) from the end of the header to the next block boundary, record by record
        size_t b_next = 0;                              // first block that starts at or after p
        while (b_next < blocks.size() && blocks[b_next].out_off < p) ++b_next;
        const size_t stop = b_next < blocks.size() ? blocks[b_next].out_off : n;
        size_t q = p;
        while (q + 4 <= stop) {
            if (!wait_ready(q + 4 + 16)) { chain_ok = false; break; }
            const uint32_t block_size = rd32(d + q);
            if (q + 4 + block_size > n || block_size < 32) { chain_ok = false; break; }
            any_paired |= (rd16(d + q + 4 + 14) & 0x1) != 0;
            rec_at.push_back(q);
            q += 4 + (size_t)block_size;
        }
        // (b) if that landed exactly on the boundary and every later block ends on a record boundary as well, the
        //     per-block walks of the inflating threads are the chain (induction over the blocks)
        //     (the first few blocks decide which way to go: a writer that splits records keeps the sequential walk
        //     running behind the inflating threads, as before)
        bool try_blockwise = chain_ok && q == stop;
        for (size_t i = b_next; try_blockwise && i < std::min(blocks.size(), b_next + 4); ++i) {
            if (!wait_ready(blocks[i].out_off + blocks[i].out_len)) { chain_ok = false; break; }
            try_blockwise = block_recs[i].ok;
        }
        if (chain_ok && !try_blockwise) {
            while (q + 4 <= n) {
                if (!wait_ready(q + 4 + 16)) { chain_ok = false; break; }
                const uint32_t block_size = rd32(d + q);
                if (q + 4 + block_size > n || block_size < 32) { chain_ok = false; break; }
                any_paired |= (rd16(d + q + 4 + 14) & 0x1) != 0;
                rec_at.push_back(q);
                q += 4 + (size_t)block_size;
            }
        }
        join_inflaters();
        bool blockwise = try_blockwise && chain_ok && !failed.load();
        for (size_t i = b_next; blockwise && i < blocks.size(); ++i) blockwise = block_recs[i].ok;
        if (blockwise) {
            for (size_t i = b_next; i < blocks.size(); ++i) {
                any_paired |= block_recs[i].paired;
                const size_t base = blocks[i].out_off;
                for (const uint32_t o : block_recs[i].at) rec_at.push_back(base + o);
            }
        } else if (try_blockwise && chain_ok && !failed.load()) {
            // (c) a later block does not end on a record boundary after all: the sequential walk over the rest
            while (q + 4 <= n) {
                const uint32_t block_size = rd32(d + q);
                if (q + 4 + block_size > n || block_size < 32) { chain_ok = false; break; }
                any_paired |= (rd16(d + q + 4 + 14) & 0x1) != 0;
                rec_at.push_back(q);
                q += 4 + (size_t)block_size;
            }
        }
        std::vector<BlockRecs>().swap(block_recs);
        join_inflaters();
        if (failed.load()) { delete res; dyn_set_error("dyn_bam_pack: inflate failed in %s", path); return DYN_ERR_FORMAT; }
        timer.mark("inflate+walk");
        const int nt = std::max(1, std::min(n_threads, 64));
        if (chain_ok && !any_paired && nt > 1 && rec_at.size() >= 4096) {
            struct Part {
                std::vector<uint8_t> blob;
                std::vector<uint32_t> size16;
                std::vector<int32_t> ref_id, pos, hi, score;
                std::vector<uint16_t> flag;
                std::vector<uint8_t> gx_assigned;
                Pool name, barcode, umi, gene, raw;
                std::vector<int64_t> rec_index;
                std::string error;
                int err_code = 0;
            };
            const size_t n_rec = rec_at.size();
            const size_t n_parts = (size_t)nt * 4;
            std::vector<Part> parts(n_parts);
            std::atomic<size_t> next_part(0);
            auto fail = [](Part &pt, int code, const char *msg) { pt.err_code = code; pt.error = msg; };
            auto work = [&]() {
                std::vector<uint32_t> toks;
                for (;;) {
                    const size_t pi = next_part.fetch_add(1);
                    if (pi >= n_parts) return;
                    Part &pt = parts[pi];
                    const size_t r0 = n_rec * pi / n_parts, r1 = n_rec * (pi + 1) / n_parts;
                    pt.blob.reserve((r1 - r0) * 200);
                    for (size_t ri = r0; ri < r1; ++ri) {
                        const size_t q = rec_at[ri];
                        const uint32_t block_size = rd32(d + q);
                        const uint8_t *rec = d + q + 4;
                        const int32_t ref_id = (int32_t)rd32(rec);
                        const uint32_t l_read_name = rec[8], n_cigar = rd16(rec + 12), l_seq = rd32(rec + 16);
                        const uint16_t flag = rd16(rec + 14);
                        if (ref_id < 0) continue;
                        if (flag & skip_flags) continue;
                        const size_t fixed = 32 + l_read_name + 4 * (size_t)n_cigar + (l_seq + 1) / 2 + l_seq;
                        if (fixed > block_size) { fail(pt, DYN_ERR_FORMAT, "dyn_bam_pack: corrupt record"); return; }
                        Aux a;
                        if (!parse_aux(rec + fixed, rec + block_size, barcode_tag, umi_tag, gene_tag, a)) { fail(pt, DYN_ERR_FORMAT, "dyn_bam_pack: corrupt aux data"); return; }
                        if ((barcode_tag && !a.has_bc) || (umi_tag && !a.has_umi) || (require_gene && gene_tag && !a.has_gx)) continue;
                        if (barcode_tag && a.bc_len == 1 && a.bc[0] == '-') continue;
                        if (!a.has_hi || !a.has_as) { fail(pt, DYN_ERR_FORMAT, "dyn_bam_pack: alignment without HI / AS tag (KeyError in the reference, bam.py:295-296)"); return; }
                        if (!a.has_md) { fail(pt, DYN_ERR_FORMAT, "dyn_bam_pack: alignment without MD tag (count.py:119-124)"); return; }
                        if (l_seq > 0xffff || n_cigar > 0xffff) { fail(pt, DYN_ERR_FORMAT, "dyn_bam_pack: record too long for the packed layout"); return; }
                        const size_t bytes = pack_record(rec, a, 0, toks, pt.blob);
                        pt.size16.push_back((uint32_t)(bytes >> 4));
                        pt.ref_id.push_back(ref_id);
                        pt.pos.push_back((int32_t)rd32(rec + 4));
                        pt.flag.push_back(flag);
                        pt.hi.push_back((int32_t)a.hi);
                        pt.score.push_back((int32_t)a.as);
                        pt.gx_assigned.push_back(a.has_gx ? 1 : 0);
                        pt.name.add((const char *)rec + 32, l_read_name ? l_read_name - 1 : 0);
                        pt.barcode.add(a.has_bc ? a.bc : "NA", a.has_bc ? a.bc_len : 2);
                        pt.umi.add(a.has_umi ? a.umi : "NA", a.has_umi ? a.umi_len : 2);
                        pt.gene.add(a.has_gx ? a.gx : "", a.has_gx ? a.gx_len : 0);
                        pt.rec_index.push_back((int64_t)ri);
                        if (keep_raw) pt.raw.add((const char *)rec, block_size);
                    }
                }
            };
            {
                std::vector<std::thread> th;
                for (int t = 0; t < nt; ++t) th.emplace_back(work);
                for (auto &t : th) t.join();
            }
            timer.mark("pack");
            for (const Part &pt : parts)
                if (pt.err_code) { delete res; dyn_set_error("%s", pt.error.c_str()); return pt.err_code; }
            // stitch: unit counts and byte totals per part, then parallel copies
            size_t units = 0, bytes = 0;
            std::vector<size_t> unit0(n_parts), byte0(n_parts);
            for (size_t i = 0; i < n_parts; ++i) { unit0[i] = units; byte0[i] = bytes; units += parts[i].size16.size(); bytes += parts[i].blob.size(); }
            if ((bytes >> 4) > 0xfffffff0ull) { delete res; dyn_set_error("dyn_bam_pack: more than 64 GiB of packed records: split the BAM by read range"); return DYN_ERR_NOMEM; }
            res->n_records_seen = (int64_t)n_rec;
            res->blob_bytes = bytes;
            const size_t alloc = std::max<size_t>(bytes, 16);
            if (cudaHostAlloc((void **)&res->blob, alloc, cudaHostAllocPortable) == cudaSuccess) res->blob_pinned = true;
            else {
                cudaGetLastError();
                if (posix_memalign((void **)&res->blob, 256, alloc) != 0) { delete res; dyn_set_error("dyn_bam_pack: out of memory"); return DYN_ERR_NOMEM; }
            }
            res->rec_off.assign(units + 1, 0);
            res->ref_id.resize(units); res->pos.resize(units); res->hi.resize(units); res->score.resize(units);
            res->flag.resize(units); res->gx_assigned.resize(units); res->paired.assign(units, 0);
            res->rec_index.resize(units);
            res->raw_mate.off.assign(units + 1, 0);
            Pool *dst_pool[5] = {&res->name, &res->barcode, &res->umi, &res->gene, &res->raw};
            std::vector<size_t> pool0[5];
            for (int k = 0; k < 5; ++k) {
                pool0[k].resize(n_parts);
                size_t tot = 0;
                for (size_t i = 0; i < n_parts; ++i) {
                    const Pool &sp = k == 0 ? parts[i].name : (k == 1 ? parts[i].barcode : (k == 2 ? parts[i].umi : (k == 3 ? parts[i].gene : parts[i].raw)));
                    pool0[k][i] = tot;
                    tot += sp.data.size();
                }
                if (tot > 0xffffffffull) { delete res; dyn_set_error("dyn_bam_pack: string pool over 4 GiB: split the BAM by read range"); return DYN_ERR_NOMEM; }
                dst_pool[k]->data.resize(tot);
                dst_pool[k]->off.assign(units + 1, 0);
                dst_pool[k]->off[units] = (uint32_t)tot;
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
                    if (nu) {
                        memcpy(res->ref_id.data() + u0, pt.ref_id.data(), 4 * nu);
                        memcpy(res->pos.data() + u0, pt.pos.data(), 4 * nu);
                        memcpy(res->hi.data() + u0, pt.hi.data(), 4 * nu);
                        memcpy(res->score.data() + u0, pt.score.data(), 4 * nu);
                        memcpy(res->flag.data() + u0, pt.flag.data(), 2 * nu);
                        memcpy(res->gx_assigned.data() + u0, pt.gx_assigned.data(), nu);
                        memcpy(res->rec_index.data() + u0, pt.rec_index.data(), 8 * nu);
                    }
                    for (int k = 0; k < (keep_raw ? 5 : 4); ++k) {
                        const Pool &sp = k == 0 ? pt.name : (k == 1 ? pt.barcode : (k == 2 ? pt.umi : (k == 3 ? pt.gene : pt.raw)));
                        if (!sp.data.empty()) memcpy(dst_pool[k]->data.data() + pool0[k][i], sp.data.data(), sp.data.size());
                        for (size_t j = 0; j < nu; ++j) dst_pool[k]->off[u0 + j] = (uint32_t)(pool0[k][i] + sp.off[j]);
                    }
                }
            };
            {
                std::vector<std::thread> th;
                for (int t = 0; t < nt; ++t) th.emplace_back(copy);
                for (auto &t : th) t.join();
            }
            res->rec_off[units] = (uint32_t)(bytes >> 4);
            timer.mark("stitch");
            *out_result = res;
            return DYN_OK;
        }
    }
    std::vector<uint8_t> blob;
    blob.reserve(n / 2);
    res->rec_off.push_back(0);
    std::unordered_map<std::string, Pending> pending;
    int32_t cur_ref = -2;
    std::string key;
    std::vector<uint32_t> toks;
    while (p + 4 <= n) {
        const uint32_t block_size = rd32(d + p);
        const uint8_t *rec = d + p + 4;
        if (p + 4 + block_size > n || block_size < 32) { delete res; dyn_set_error("dyn_bam_pack: truncated record"); return DYN_ERR_FORMAT; }
        const size_t rec_pos = p + 4;
        p += 4 + block_size;
        ++res->n_records_seen;
        const int32_t ref_id = (int32_t)rd32(rec);
        const uint32_t l_read_name = rec[8], n_cigar = rd16(rec + 12), l_seq = rd32(rec + 16);
        const uint16_t flag = rd16(rec + 14);
        if (ref_id != cur_ref) { pending.clear(); cur_ref = ref_id; }      // parse_read_contig runs per contig
        if (ref_id < 0) continue;                                           // bam.fetch(contig) never yields these
        if (flag & skip_flags) continue;                                    // bam.py:173: secondary, unmapped, duplicate
        const size_t fixed = 32 + l_read_name + 4 * (size_t)n_cigar + (l_seq + 1) / 2 + l_seq;
        if (fixed > block_size) { delete res; dyn_set_error("dyn_bam_pack: corrupt record"); return DYN_ERR_FORMAT; }
        Aux a;
        if (!parse_aux(rec + fixed, rec + block_size, barcode_tag, umi_tag, gene_tag, a)) { delete res; dyn_set_error("dyn_bam_pack: corrupt aux data"); return DYN_ERR_FORMAT; }
        if ((barcode_tag && !a.has_bc) || (umi_tag && !a.has_umi) || (require_gene && gene_tag && !a.has_gx)) continue;   // bam.py:174, 242-248
        if (barcode_tag && a.bc_len == 1 && a.bc[0] == '-') continue;       // bam.py:290
        if (!a.has_hi || !a.has_as) { delete res; dyn_set_error("dyn_bam_pack: alignment without HI / AS tag (KeyError in the reference, bam.py:295-296)"); return DYN_ERR_FORMAT; }
        if (!a.has_md) { delete res; dyn_set_error("dyn_bam_pack: alignment without MD tag (count.py:119-124)"); return DYN_ERR_FORMAT; }
        if (l_seq > 0xffff || n_cigar > 0xffff) { delete res; dyn_set_error("dyn_bam_pack: record too long for the packed layout"); return DYN_ERR_FORMAT; }
        const char *name = (const char *)rec + 32;
        const size_t name_len = l_read_name ? l_read_name - 1 : 0;
        const uint8_t *mate_rec = nullptr;
        Aux ma;
        if (flag & 0x1) {                                                   // bam.py:307-312
            key.assign(name, name_len);
            key.push_back('\0');
            key.append(std::to_string(a.hi));
            auto it = pending.find(key);
            if (it == pending.end()) { pending.emplace(key, Pending{rec_pos}); continue; }
            mate_rec = d + it->second.rec_pos;
            pending.erase(it);
            const uint32_t m_name = mate_rec[8], m_cig = rd16(mate_rec + 12), m_seq = rd32(mate_rec + 16);
            const uint32_t m_size = rd32(mate_rec - 4);
            parse_aux(mate_rec + 32 + m_name + 4 * (size_t)m_cig + (m_seq + 1) / 2 + m_seq, mate_rec + m_size, barcode_tag, umi_tag, gene_tag, ma);
            if (!ma.has_md) { delete res; dyn_set_error("dyn_bam_pack: alignment without MD tag (count.py:119-124)"); return DYN_ERR_FORMAT; }
        }
        pack_record(rec, a, mate_rec ? DYN_REC_HAS_MATE : 0, toks, blob);
        if (mate_rec) pack_record(mate_rec, ma, 0, toks, blob);
        res->rec_off.push_back((uint32_t)(blob.size() >> 4));
        res->ref_id.push_back(ref_id);
        res->pos.push_back((int32_t)rd32(rec + 4));
        res->flag.push_back(flag);
        res->hi.push_back((int32_t)a.hi);
        res->score.push_back((int32_t)a.as);
        res->gx_assigned.push_back(a.has_gx ? 1 : 0);
        res->paired.push_back(mate_rec ? 1 : 0);
        res->rec_index.push_back(res->n_records_seen - 1);
        if (keep_raw) {
            res->raw.add((const char *)rec, block_size);
            if (mate_rec) res->raw_mate.add((const char *)mate_rec, rd32(mate_rec - 4)); else res->raw_mate.add("", 0);
        }
        res->name.add(name, name_len);
        res->barcode.add(a.has_bc ? a.bc : "NA", a.has_bc ? a.bc_len : 2);
        res->umi.add(a.has_umi ? a.umi : "NA", a.has_umi ? a.umi_len : 2);
        res->gene.add(a.has_gx ? a.gx : "", a.has_gx ? a.gx_len : 0);
        if ((blob.size() >> 4) > 0xfffffff0ull) { delete res; dyn_set_error("dyn_bam_pack: more than 64 GiB of packed records: split the BAM by read range"); return DYN_ERR_NOMEM; }
    }
    res->blob_bytes = blob.size();
    const size_t alloc = std::max<size_t>(blob.size(), 16);
    if (cudaHostAlloc((void **)&res->blob, alloc, cudaHostAllocPortable) == cudaSuccess) res->blob_pinned = true;
    else {
        cudaGetLastError();
        if (posix_memalign((void **)&res->blob, 256, alloc) != 0) { delete res; dyn_set_error("dyn_bam_pack: out of memory"); return DYN_ERR_NOMEM; }
    }
    if (!blob.empty()) memcpy(res->blob, blob.data(), blob.size());
    *out_result = res;
    return DYN_OK;
}

extern "C" void dyn_bam_result_free(dyn_bam_result_t *r) {
    if (!r) return;
    if (r->extra) dyn_bam_chunk_extra_free(r);      // a stream chunk: its blob goes back to the stream's pool
    if (r->blob) { if (r->blob_pinned) cudaFreeHost(r->blob); else free(r->blob); }
    delete r;
}

extern "C" int64_t dyn_bam_n_units(const dyn_bam_result_t *r) { return r ? (int64_t)r->ref_id.size() : 0; }
extern "C" int64_t dyn_bam_n_records_seen(const dyn_bam_result_t *r) { return r ? r->n_records_seen : 0; }
extern "C" int64_t dyn_bam_n_refs(const dyn_bam_result_t *r) { return r ? (int64_t)r->ref_len.size() : 0; }

// field: 0 blob (uint8), 1 rec_off (uint32, n+1), 2 ref_id (int32), 3 pos (int32), 4 flag (uint16), 5 HI (int32),
// 6 AS (int32), 7 gx_assigned (uint8), 8 paired (uint8), 9 ref_len (int32), 10 header text (char)
// string pools (chars / uint32 offsets n+1): 20/21 read name, 22/23 barcode, 24/25 umi, 26/27 gene, 28/29 reference names
extern "C" const void *dyn_bam_field(const dyn_bam_result_t *r, int field, int64_t *n_bytes) {
    if (!r) return nullptr;
    const void *p = nullptr;
    size_t nb = 0;
    auto vec = [&](const auto &v) { p = v.data(); nb = v.size() * sizeof(v[0]); };
    switch (field) {
        case 0: p = r->blob; nb = r->blob_bytes; break;
        case 1: vec(r->rec_off); break;
        case 2: vec(r->ref_id); break;
        case 3: vec(r->pos); break;
        case 4: vec(r->flag); break;
        case 5: vec(r->hi); break;
        case 6: vec(r->score); break;
        case 7: vec(r->gx_assigned); break;
        case 8: vec(r->paired); break;
        case 9: vec(r->ref_len); break;
        case 10: p = r->header_text.data(); nb = r->header_text.size(); break;
        case 11: case 12: case 13: p = dyn_bam_chunk_field(r, field, &nb); break;
        case 14: vec(r->rec_index); break;
        case 30: vec(r->raw.data); break;
        case 31: vec(r->raw.off); break;
        case 32: vec(r->raw_mate.data); break;
        case 33: vec(r->raw_mate.off); break;
        case 20: vec(r->name.data); break;
        case 21: vec(r->name.off); break;
        case 22: vec(r->barcode.data); break;
        case 23: vec(r->barcode.off); break;
        case 24: vec(r->umi.data); break;
        case 25: vec(r->umi.off); break;
        case 26: vec(r->gene.data); break;
        case 27: vec(r->gene.off); break;
        case 28: vec(r->ref_name.data); break;
        case 29: vec(r->ref_name.off); break;
        default: return nullptr;
    }
    if (n_bytes) *n_bytes = (int64_t)nb;
    return p;
}

// First `max_records` alignment records of a BAM: which aux tags occur, which flag bits occur (get_tags_from_bam,
// check_bam_contains_secondary / _unmapped / _duplicate, check_bam_is_paired: dynast/preprocessing/bam.py:447-570).
extern "C" int dyn_bam_peek(const char *path, int64_t max_records, char *tags_out, int32_t tags_capacity, int32_t *n_tags,
                            uint32_t *flag_or, int64_t *n_seen) {
    if (!path || !n_tags || !flag_or || !n_seen) { dyn_set_error("dyn_bam_peek: null argument"); return DYN_ERR_ARG; }
    *n_tags = 0; *flag_or = 0; *n_seen = 0;
    FILE *f = fopen(path, "rb");
    if (!f) { dyn_set_error("dyn_bam_peek: cannot open %s", path); return DYN_ERR_IO; }
    std::vector<uint8_t> data, comp;
    std::vector<uint16_t> tags;
    size_t p = 0;
    bool header_done = false, eof = false;
    auto more = [&]() -> bool {          // inflate one more BGZF block onto `data`
        uint8_t h[18];
        if (fread(h, 1, 18, f) != 18) { eof = true; return false; }
        if (h[0] != 31 || h[1] != 139 || h[2] != 8 || !(h[3] & 4)) return false;
        const uint32_t xlen = rd16(h + 10);
        std::vector<uint8_t> extra(xlen);
        memcpy(extra.data(), h + 12, std::min<size_t>(6, xlen));
        if (xlen > 6 && fread(extra.data() + 6, 1, xlen - 6, f) != xlen - 6) return false;
        uint32_t bsize = 0;
        for (size_t x = 0; x + 4 <= xlen;) {
            const uint32_t slen = rd16(extra.data() + x + 2);
            if (extra[x] == 'B' && extra[x + 1] == 'C' && slen == 2 && x + 6 <= xlen) bsize = rd16(extra.data() + x + 4);
            x += 4 + slen;
        }
        if (bsize < 12 + xlen + 8 - 1) return false;
        const size_t clen = (size_t)bsize + 1 - 12 - xlen - 8;
        comp.resize(clen + 8);
        if (fread(comp.data(), 1, clen + 8, f) != clen + 8) return false;
        const uint32_t isize = rd32(comp.data() + clen + 4);
        const size_t o = data.size();
        data.resize(o + isize);
        if (isize) {
            z_stream zs;
            memset(&zs, 0, sizeof(zs));
            if (inflateInit2(&zs, -15) != Z_OK) return false;
            zs.next_in = comp.data(); zs.avail_in = (uInt)clen; zs.next_out = data.data() + o; zs.avail_out = isize;
            const int rc = inflate(&zs, Z_FINISH);
            inflateEnd(&zs);
            if (rc != Z_STREAM_END) return false;
        }
        return true;
    };
    auto need = [&](size_t n) -> bool { while (data.size() < n) if (!more()) return false; return true; };
    int rc = DYN_OK;
    if (!need(12) || memcmp(data.data(), "BAM\1", 4) != 0) { fclose(f); dyn_set_error("dyn_bam_peek: %s is not a BAM file", path); return DYN_ERR_FORMAT; }
    p = 4;
    const uint32_t l_text = rd32(data.data() + p); p += 4 + l_text;
    if (need(p + 4)) {
        const uint32_t n_ref = rd32(data.data() + p); p += 4;
        header_done = true;
        for (uint32_t r = 0; r < n_ref && header_done; ++r) {
            if (!need(p + 4)) { header_done = false; break; }
            const uint32_t l_name = rd32(data.data() + p);
            p += 4 + l_name + 4;
        }
    }
    while (header_done && *n_seen < max_records) {
        if (!need(p + 4)) break;
        const uint32_t bs = rd32(data.data() + p);
        if (bs < 32 || !need(p + 4 + bs)) break;
        const uint8_t *rec = data.data() + p + 4;
        p += 4 + bs;
        ++*n_seen;
        *flag_or |= rd16(rec + 14);
        const uint32_t l_name = rec[8], n_cigar = rd16(rec + 12), l_seq = rd32(rec + 16);
        const size_t fixed = 32 + l_name + 4 * (size_t)n_cigar + (l_seq + 1) / 2 + l_seq;
        if (fixed > bs) { rc = DYN_ERR_FORMAT; break; }
        const uint8_t *a = rec + fixed, *end = rec + bs;
        while (a + 3 <= end) {
            const uint16_t tag = (uint16_t)(a[0] | (a[1] << 8));
            if (std::find(tags.begin(), tags.end(), tag) == tags.end()) tags.push_back(tag);
            const char ty = (char)a[2];
            a += 3;
            if (ty == 'A' || ty == 'c' || ty == 'C') a += 1;
            else if (ty == 's' || ty == 'S') a += 2;
            else if (ty == 'i' || ty == 'I' || ty == 'f') a += 4;
            else if (ty == 'Z' || ty == 'H') { while (a < end && *a) ++a; ++a; }
            else if (ty == 'B') { if (a + 5 > end) break; const char sub = (char)a[0]; const uint32_t cnt = rd32(a + 1); a += 5 + (size_t)cnt * ((sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4); }
            else break;
        }
        if (p > (64u << 20)) { data.erase(data.begin(), data.begin() + p); p = 0; }
    }
    fclose(f);
    if (rc) { dyn_set_error("dyn_bam_peek: corrupt record in %s", path); return rc; }
    *n_tags = (int32_t)tags.size();
    for (size_t i = 0; i < tags.size() && (int32_t)(2 * i + 1) < tags_capacity; ++i) { tags_out[2 * i] = (char)(tags[i] & 0xff); tags_out[2 * i + 1] = (char)(tags[i] >> 8); }
    (void)eof;
    return DYN_OK;
}
