/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the per-read conversion tally on packed records.
 *
 * Input is the same packed-record blob the CUDA tally reads (include/dynast_b200.h); the algorithm is a
 * deliberately plain per-base restatement of
 *   dynast/preprocessing/bam.py:297-429        (aligned-pair walk, content, paired-overlap rules, mismatches)
 *   dynast/preprocessing/conversion.py:417-426 (Phred > quality, SNP filter, 12 + 4 counters)
 * with pysam's get_aligned_pairs / get_reference_sequence semantics restated from the MD tag
 * (see oracle/bamio.py).  It shares no code with the kernel: every base is visited one at a time.
 *
 * Pinned by tests/test_oracle_tally.py: packing the reference's fixture BAMs and running this gives
 * exactly the rows of the fixture alignments.csv / conversions.csv (UMI, Smart-seq paired, NASC).
 * Only tests/, smoke() and bench.py's CPU-baseline legs may call this.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int32_t pos;
    int32_t ref_id;
    uint16_t l_seq, n_cigar, n_tok, flag;
} hdr_t;

#define HAS_MATE 0x8000u
#define BAD_MD 0x4000u
#define LEAN 0x2000u   /* no qual[] stream: the Phred of a mismatching base rides in bits 21..28 of its MD token */
#define ST_OVERFLOW 0x1u
#define ST_BAD 0x2u
#define ST_NON_ACGT 0x4u
#define ST_CAPACITY 0x8u
#define F_NASC 0x1u

static const char SEQ_CODES[] = "=ACMGRSVTWYHKDBN";

static int letter_code(int ch) {
    if (ch >= 'a' && ch <= 'z') ch -= 32;
    for (int i = 0; i < 16; i++)
        if (SEQ_CODES[i] == ch) return i;
    return 0;
}

typedef struct {
    int n_pairs;          /* aligned (M/=/X) pairs */
    int *qpos, *rpos;
    uint8_t *refn;        /* reference base (BAM nibble code) of each aligned pair */
    uint8_t *readn;       /* read base code */
    uint8_t *q;
    int n_del;
    uint8_t *deln;        /* deleted reference bases */
    int bad;
    size_t bytes;         /* record size incl. padding */
    hdr_t h;
} resolved_t;

static size_t record_bytes(const hdr_t *h) {
    size_t b = 16 + 4 * (size_t)h->n_cigar + 4 * (size_t)h->n_tok + ((((size_t)h->l_seq + 1) / 2 + 3) & ~(size_t)3) +
               ((h->flag & LEAN) ? 0 : h->l_seq);
    return (b + 15) & ~(size_t)15;
}

/* Aligned pairs of one record.  The MD tag arrives tokenised by the packer (include/dynast_b200.h): one token
 * per MD letter in MD order -- aligned offset, reference base, deleted-base flag.  An aligned base whose
 * running aligned offset equals the next mismatch token's offset takes that token's reference base, every
 * other aligned base matches; a D op consumes one deleted-base token per base. */
static void resolve(const uint8_t *rec, resolved_t *r) {
    memcpy(&r->h, rec, 16);
    const hdr_t *h = &r->h;
    const uint32_t *cig = (const uint32_t *)(rec + 16);
    const uint32_t *tok = cig + h->n_cigar;
    const uint8_t *seq = (const uint8_t *)(tok + h->n_tok);
    const uint8_t *qual = seq + ((((size_t)h->l_seq + 1) / 2 + 3) & ~(size_t)3);
    r->bytes = record_bytes(h);
    size_t cap = (size_t)h->l_seq + 1, dcap = (size_t)h->n_tok + 1;
    r->qpos = (int *)malloc(sizeof(int) * cap);
    r->rpos = (int *)malloc(sizeof(int) * cap);
    r->refn = (uint8_t *)malloc(cap);
    r->readn = (uint8_t *)malloc(cap);
    r->q = (uint8_t *)malloc(cap);
    r->deln = (uint8_t *)malloc(dcap);
    r->n_pairs = r->n_del = 0;
    r->bad = (h->flag & BAD_MD) != 0;
    int qp = 0, rp = h->pos, ti = 0, aligned = 0;
    for (int ci = 0; ci < h->n_cigar && !r->bad; ci++) {
        int op = cig[ci] & 0xf, len = (int)(cig[ci] >> 4);
        if (op == 0 || op == 7 || op == 8) {
            for (int j = 0; j < len; j++) {
                if (qp >= h->l_seq) { r->bad = 1; break; }
                int rn = (seq[qp >> 1] >> ((qp & 1) ? 0 : 4)) & 0xf;
                int gn = rn;
                /* lean records carry qualities at MD letters only -- the only ones a single-end read needs */
                int q = (h->flag & LEAN) ? 0 : qual[qp];
                if (ti < h->n_tok && !((tok[ti] >> 20) & 1) && (int)(tok[ti] & 0xffff) == aligned) {
                    gn = (int)((tok[ti] >> 16) & 0xf);
                    if (h->flag & LEAN) q = (int)((tok[ti] >> 21) & 0xff);
                    ti++;
                }
                int i = r->n_pairs++;
                r->qpos[i] = qp; r->rpos[i] = rp; r->refn[i] = (uint8_t)gn; r->readn[i] = (uint8_t)rn; r->q[i] = (uint8_t)q;
                qp++; rp++; aligned++;
            }
        } else if (op == 2) {
            for (int j = 0; j < len; j++) {
                if (ti >= h->n_tok || !((tok[ti] >> 20) & 1)) { r->bad = 1; break; }
                r->deln[r->n_del++] = (uint8_t)((tok[ti] >> 16) & 0xf);
                ti++;
            }
            rp += len;
        } else if (op == 3) {
            rp += len;
        } else if (op == 1 || op == 4) {
            qp += len;
            if (qp > h->l_seq) r->bad = 1;
        }
    }
    if (!r->bad && ti != h->n_tok) r->bad = 1;
}

static void resolved_free(resolved_t *r) {
    free(r->qpos); free(r->rpos); free(r->refn); free(r->readn); free(r->q); free(r->deln);
}

static int base_idx(int code) { return code == 1 ? 0 : code == 2 ? 1 : code == 4 ? 2 : code == 8 ? 3 : -1; }

static void add_content(const resolved_t *r, long content[4]) {
    /* Counter(get_reference_sequence().upper()): M/=/X and D reference bases */
    for (int i = 0; i < r->n_pairs; i++) { int b = base_idx(r->refn[i]); if (b >= 0) content[b]++; }
    for (int i = 0; i < r->n_del; i++) { int b = base_idx(r->deln[i]); if (b >= 0) content[b]++; }
}

static int snp_hit(const uint64_t *snps, int64_t n, uint64_t key) {
    int64_t lo = 0, hi = n;
    while (lo < hi) { int64_t m = (lo + hi) >> 1; if (snps[m] < key) lo = m + 1; else hi = m; }
    return lo < n && snps[lo] == key;
}

typedef struct { int pos; int q; int gn; int rn; int alive; } call_t;

/* Process reads [r_begin, r_end). conv_* may be NULL (then only counted). Records of read i are
 * written contiguously starting at conv_off[i] = running total, in walk order. */
int dyn_oracle_tally(const uint8_t *records, const uint32_t *rec_off, int64_t r_begin, int64_t r_end,
                     const uint64_t *snps, int64_t n_snps, int quality, uint32_t flags, uint8_t *counts,
                     uint32_t *conv_off, uint16_t *conv_n, uint64_t *conv_rec, uint32_t *conv_read,
                     int64_t conv_capacity, uint64_t *conv_total, uint32_t *status) {
    uint64_t total = conv_total ? *conv_total : 0;
    uint32_t st = 0;
    for (int64_t r = r_begin; r < r_end; r++) {
        const uint8_t *rec = records + ((size_t)rec_off[r] << 4);
        resolved_t rd, mt;
        resolve(rec, &rd);
        int has_mate = (rd.h.flag & HAS_MATE) != 0;
        if (has_mate && (rd.h.flag & LEAN)) { rd.bad = 1; has_mate = 0; }   /* pairs need every quality */
        long content[4] = {0, 0, 0, 0};
        call_t *calls = NULL;
        int n_calls = 0;
        if (rd.bad) st |= ST_BAD;
        if (has_mate) {
            resolve(rec + rd.bytes, &mt);
            if (mt.bad) st |= ST_BAD;
            add_content(&mt, content);
            long overlap[4] = {0, 0, 0, 0};
            calls = (call_t *)malloc(sizeof(call_t) * (size_t)(mt.n_pairs + 1));
            for (int i = 0; i < mt.n_pairs; i++) {
                int gn = mt.refn[i], rn = mt.readn[i];
                if (gn == 15 || rn == 15) continue;
                int shared = 0;
                for (int j = 0; j < rd.n_pairs; j++) if (rd.rpos[j] == mt.rpos[i]) { shared = 1; break; }
                if (shared) { int b = base_idx(gn); if (b >= 0) overlap[b]++; }
                if (gn != rn && (!(flags & F_NASC) || !shared)) {
                    /* dict assignment: a later call at the same position replaces the value but keeps
                     * the original insertion slot (cannot happen within one read: positions are unique) */
                    calls[n_calls].pos = mt.rpos[i]; calls[n_calls].q = mt.q[i];
                    calls[n_calls].gn = gn; calls[n_calls].rn = rn; calls[n_calls].alive = 1;
                    n_calls++;
                }
            }
            for (int b = 0; b < 4; b++) { content[b] -= overlap[b]; if (content[b] < 0) content[b] = 0; }
        }
        add_content(&rd, content);

        uint32_t cnt[12] = {0};
        uint32_t n_conv = 0;
        uint64_t base = total;
#define EMIT(POS, GN, RN, Q)                                                                           \
    do {                                                                                               \
        if (conv_rec) {                                                                                \
            if ((int64_t)(total) < conv_capacity) {                                                    \
                conv_rec[total] = (uint64_t)(uint32_t)(POS) | ((uint64_t)(((GN) << 4) | (RN)) << 32) | \
                                  ((uint64_t)(Q) << 40);                                               \
                conv_read[total] = (uint32_t)r;                                                        \
            } else st |= ST_CAPACITY;                                                                  \
        }                                                                                              \
        total++; n_conv++;                                                                             \
        if ((int)(Q) > quality) {                                                                      \
            int rb = base_idx(GN), db = base_idx(RN);                                                  \
            if (rb >= 0 && db >= 0) {                                                                  \
                int idx = rb * 3 + (db > rb ? db - 1 : db);                                            \
                uint64_t key = ((uint64_t)idx << 56) | ((uint64_t)(uint32_t)rd.h.ref_id << 32) | (uint32_t)(POS); \
                if (!(n_snps > 0 && snp_hit(snps, n_snps, key))) cnt[idx]++;                           \
            } else st |= ST_NON_ACGT;                                                                  \
        }                                                                                              \
    } while (0)

        for (int i = 0; i < rd.n_pairs; i++) {
            int gn = rd.refn[i], rn = rd.readn[i], q = rd.q[i];
            if (gn == 15 || rn == 15) continue;
            for (int c = 0; c < n_calls; c++) {
                if (calls[c].alive && calls[c].pos == rd.rpos[i]) {
                    if (!(flags & F_NASC) && calls[c].q > q) { rn = calls[c].rn; q = calls[c].q; }
                    calls[c].alive = 0;
                }
            }
            if (gn != rn) EMIT(rd.rpos[i], gn, rn, q);
        }
        for (int c = 0; c < n_calls; c++)
            if (calls[c].alive) EMIT(calls[c].pos, calls[c].gn, calls[c].rn, calls[c].q);
#undef EMIT
        uint8_t *o = counts + (size_t)r * 16;
        for (int i = 0; i < 12; i++) { if (cnt[i] > 255) { st |= ST_OVERFLOW; cnt[i] = 255; } o[i] = (uint8_t)cnt[i]; }
        for (int b = 0; b < 4; b++) { if (content[b] > 255) { st |= ST_OVERFLOW; content[b] = 255; } o[12 + b] = (uint8_t)content[b]; }
        if (conv_off) conv_off[r] = (uint32_t)base;
        if (conv_n) conv_n[r] = (uint16_t)(n_conv > 0xffff ? 0xffff : n_conv);
        resolved_free(&rd);
        if (has_mate) { resolved_free(&mt); free(calls); }
    }
    if (conv_total) *conv_total = total;
    if (status) *status |= st;
    return 0;
}
