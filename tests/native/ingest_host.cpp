// TEST SHIM: the device ingest's __host__ __device__ cores (csrc/inflate_core.cuh, csrc/bamrec_core.cuh) compiled for
// the CPU, with the warp of inflate_kernel played by a loop over 32 "lanes" (tests/test_bamdevice_host.py).
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../dynast-release_b200/csrc/bamrec_core.cuh"
#include "../../dynast-release_b200/csrc/inflate_core.cuh"
#include "../../dynast-release_b200/csrc/inflate_lane.cuh"

extern "C" int host_inflate(const uint8_t *in, uint32_t in_len, uint8_t *dst, uint32_t out_len) {
    static dinflate::Tables t;
    dinflate::State s;
    dinflate::state_init(s, in, in_len, out_len);
    if (out_len == 0) s.phase = 3;
    for (;;) {
        int n = 0;
        uint32_t bytes = 0;
        if (s.err == dinflate::OK && s.phase == 0) dinflate::begin_block(s, t);
        if (s.err == dinflate::OK && s.phase == 1) n = dinflate::decode_batch(s, t, &bytes);
        const int err = s.err, phase = s.phase;
        const uint32_t out_pos = s.out_pos;
        if (n > 0) {
            uint32_t pos[32], len[32];
            uint32_t acc = out_pos;
            for (int lane = 0; lane < n; ++lane) {
                const uint32_t e = t.batch[lane];
                len[lane] = (e >> 31) ? (e & 0x1ffu) : 1u;
                pos[lane] = acc;
                acc += len[lane];
            }
            for (int lane = 0; lane < n; ++lane) if (!(t.batch[lane] >> 31)) dst[pos[lane]] = (uint8_t)t.batch[lane];
            // the kernel's two match phases: "ready" matches (source ends before the first match output of the batch, at
            // most 32 bytes) all at once -- played here in REVERSE lane order, which is only right if they really are
            // independent of one another --, then the rest in stream order
            int first = -1;
            for (int lane = 0; lane < n; ++lane) if (t.batch[lane] >> 31) { first = lane; break; }
            bool ready[32] = {false};
            if (first >= 0) {
                for (int lane = n - 1; lane >= 0; --lane) {
                    const uint32_t e = t.batch[lane];
                    if (!(e >> 31)) continue;
                    const uint32_t l = e & 0x1ffu, dist = (e >> 9) & 0xffffu;
                    ready[lane] = l <= 32u && pos[lane] - dist + l <= pos[first];
                    if (ready[lane]) for (uint32_t k = 0; k < l; ++k) dst[pos[lane] + k] = dst[pos[lane] - dist + k];
                }
                for (int i = 0; i < n; ++i) {
                    const uint32_t e = t.batch[i];
                    if (!(e >> 31) || ready[i]) continue;
                    const uint32_t l = e & 0x1ffu, dist = (e >> 9) & 0xffffu;
                    const uint8_t *src = dst + pos[i] - dist;
                    std::vector<uint8_t> tmp(l);
                    for (int lane = 0; lane < 32; ++lane)                 // every lane reads before any lane's write is needed
                        for (uint32_t k = lane; k < l; k += 32) tmp[k] = dist >= l ? src[k] : src[k % dist];
                    memcpy(dst + pos[i], tmp.data(), l);
                }
            }
        }
        if (phase == 2 && err == dinflate::OK) {
            memcpy(dst + out_pos, s.br.p, s.stored_left);
            s.br.p += s.stored_left; s.out_pos += s.stored_left; s.stored_left = 0; s.phase = s.last ? 3 : 0;
        }
        s.out_pos += bytes;
        if (err != dinflate::OK || s.phase == 3) break;
    }
    if (s.err) return s.err;
    return s.out_pos == out_len ? 0 : dinflate::ERR_OUTPUT;
}

// The one-lane-per-block decoder of inflate_lanes_kernel on one "lane" (tables not interleaved).
extern "C" int host_inflate_lane(const uint8_t *in, uint32_t in_len, uint8_t *dst, uint32_t out_len) {
    static uint16_t tables[dinflate::lane::kUnitsAll];
    const dinflate::lane::Mem<1> m{tables};
    std::vector<uint32_t> tok(out_len + 1);
    return dinflate::lane::inflate_block(m, in, in_len, tok.data(), dst, out_len);
}

// Records of one inflated chunk (record chain walked here), parsed and packed by the device cores.
// Returns the number of units, or -(error code).
extern "C" long host_parse_pack(const uint8_t *data, uint32_t n_bytes, uint32_t first_skip, const char *bc, const char *umi, const char *gx,
                                int require_gene, int lean, uint8_t *blob, uint32_t *rec_off, uint64_t *bc_key, uint64_t *umi_key,
                                uint64_t *gx_hash, int32_t *score, int32_t *hi, uint8_t *gx_assigned, long *n_seen) {
    bamrec::Tags tg;
    auto tag2 = [](const char *t) -> uint32_t { return (t && t[0] && t[1]) ? ((uint32_t)(uint8_t)t[0] | ((uint32_t)(uint8_t)t[1] << 8)) : 0u; };
    tg.bc = tag2(bc); tg.umi = tag2(umi); tg.gx = tag2(gx); tg.require_gene = require_gene; tg.lean = lean; tg.full_keys = 0;
    long units = 0;
    uint64_t off16 = 0;
    uint32_t q = first_skip;
    *n_seen = 0;
    while (q + 4 <= n_bytes) {
        const uint32_t bs = bamrec::ld32(data + q);
        if (bs < 32 || q + 4 + bs > n_bytes) return -100;
        ++*n_seen;
        bamrec::Parsed pr;
        bamrec::parse_record(data, q, tg, pr);
        if (pr.error) return -(long)pr.error;
        if (pr.size16) {
            bamrec::pack_record(data, q, pr, lean, blob + (off16 << 4));
            rec_off[units] = (uint32_t)off16;
            bc_key[units] = pr.bc_key; umi_key[units] = pr.umi_key; gx_hash[units] = pr.gx_assigned ? pr.gx_hash : 0;
            score[units] = pr.score; hi[units] = pr.hi; gx_assigned[units] = pr.gx_assigned;
            off16 += pr.size16;
            ++units;
        }
        q += 4 + bs;
    }
    rec_off[units] = (uint32_t)off16;
    return units;
}

extern "C" uint64_t host_fnv1a(const char *s) { return bamrec::fnv1a(reinterpret_cast<const uint8_t *>(s), (uint32_t)strlen(s)); }
