// Velocity assignment: spliced / unspliced / ambiguous and the gene of every counting unit.
//
// Replaces dynast/preprocessing/bam.py:177-228 (assign_velocity_type over ngs_tools segment collections: a Python
// loop over every gene of the contig, every transcript of the gene and every exon / intron of the transcript, per
// read) and bam.py:363-383 (read strand from the flags).
//   * the annotation lives in HBM as flat interval tables (dynast_b200/gtf.py AnnotationTables): genes grouped by
//     contig and sorted by (start, end), a running maximum of the gene ends, transcripts per gene, collapsed exons
//     and introns per transcript;
//   * one thread per unit: the unit's aligned blocks come from the CIGAR words alone (runs of consecutive M/=/X
//     positions -- D and N split a run, I does not; the two mates of a pair contribute the union of their
//     positions, bam.py:352), at most kMaxBlocks of them;
//   * a gene is evaluated transcript by transcript with binary searches over the sorted exon / intron lists:
//     any_exon_overlap, any_exon_only (every block inside ONE exon of one transcript), any_intron_overlap;
//   * a unit that carries a gene (the GX tag) is evaluated against that gene only, without strand or range test;
//     otherwise the genes of its contig that contain the alignment span and lie on the read's strand are found
//     by a binary search on the start plus a backward scan bounded by the running maximum, and the unit is
//     assigned iff exactly ONE of them overlaps it (bam.py:206-212; the outcome does not depend on the order).
// HBM traffic: header + CIGAR words of every record in, 5 bytes per unit out; the tables (a few MB for a human
// annotation) stay in L2.
#include "common.cuh"

namespace {

constexpr int kBlock = 128;
constexpr int kMaxBlocks = 48;

struct VelocityTables {
    const int *gene_start, *gene_end, *gene_pmax;
    const uint8_t *gene_strand;
    const int *contig_off;
    int n_ref;
    const int *gene_tr_off, *tr_exon_off, *tr_intron_off;
    const int *exon_start, *exon_end, *intron_start, *intron_end;
};

// first segment of [lo, hi) whose end is > a (segments sorted, disjoint: ends ascend)
__device__ __forceinline__ int first_end_above(const int *seg_end, int lo, int hi, int a) {
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(seg_end + mid) > a) hi = mid; else lo = mid + 1; }
    return lo;
}

__device__ __forceinline__ bool blocks_overlap(const int *bs, const int *be, int nb, const int *s, const int *e, int lo, int hi) {
    if (lo >= hi) return false;
    for (int i = 0; i < nb; ++i) {
        const int j = first_end_above(e, lo, hi, bs[i]);
        if (j < hi && __ldg(s + j) < be[i]) return true;
    }
    return false;
}

__device__ __forceinline__ bool blocks_subset(const int *bs, const int *be, int nb, const int *s, const int *e, int lo, int hi) {
    for (int i = 0; i < nb; ++i) {
        const int j = first_end_above(e, lo, hi, bs[i]);
        if (j >= hi || __ldg(s + j) > bs[i] || __ldg(e + j) < be[i]) return false;
    }
    return true;
}

// aligned blocks of one record appended to (bs, be); returns false when they do not fit
__device__ __forceinline__ bool append_blocks(const uint8_t *rec, int *bs, int *be, int &nb) {
    const uint4 h = *reinterpret_cast<const uint4 *>(rec);
    const uint32_t n_cigar = h.z >> 16;
    const uint32_t *cig = reinterpret_cast<const uint32_t *>(rec + 16);
    int r = (int)h.x;
    int cur = -1;
    for (uint32_t c = 0; c < n_cigar; ++c) {
        const uint32_t w = __ldg(cig + c);
        const int op = (int)(w & 0xf), len = (int)(w >> 4);
        if (op == 0 || op == 7 || op == 8) {
            if (len > 0) {
                if (cur >= 0 && be[cur] == r) be[cur] = r + len;
                else {
                    if (nb >= kMaxBlocks) return false;
                    cur = nb++;
                    bs[cur] = r; be[cur] = r + len;
                }
            }
            r += len;
        } else if (op == 2 || op == 3) r += len;
    }
    return true;
}

__global__ void __launch_bounds__(kBlock) velocity_kernel(const uint8_t *records, const uint32_t *rec_off, long long n_units,
                                                           const int32_t *gx, VelocityTables t, int strand_mode, int strict,
                                                           int32_t *out_gene, uint8_t *out_type, uint32_t *status) {
    const long long u = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (u >= n_units) return;
    const uint8_t *rec = records + ((size_t)rec_off[u] << 4);
    const uint4 h = *reinterpret_cast<const uint4 *>(rec);
    const uint32_t flag = h.w >> 16;
    const int ref_id = (int)h.y;
    int bs[kMaxBlocks], be[kMaxBlocks];
    int nb = 0;
    bool fits = append_blocks(rec, bs, be, nb);
    if (fits && (flag & DYN_REC_HAS_MATE)) {
        const uint32_t n_cigar = h.z >> 16, l_seq = h.z & 0xffff, n_tok = h.w & 0xffff;
        const size_t b = 16 + 4 * (size_t)n_cigar + 4 * (size_t)n_tok + (((((size_t)l_seq + 1) >> 1) + 3) & ~(size_t)3) + l_seq;
        fits = append_blocks(rec + ((b + 15) & ~(size_t)15), bs, be, nb);
        if (fits) {
            // union of the two mates' positions: sort by start, merge runs that overlap or touch
            for (int i = 1; i < nb; ++i) {
                const int s = bs[i], e = be[i];
                int j = i - 1;
                while (j >= 0 && bs[j] > s) { bs[j + 1] = bs[j]; be[j + 1] = be[j]; --j; }
                bs[j + 1] = s; be[j + 1] = e;
            }
            int m = 0;
            for (int i = 1; i < nb; ++i) {
                if (bs[i] <= be[m]) be[m] = max(be[m], be[i]);
                else { ++m; bs[m] = bs[i]; be[m] = be[i]; }
            }
            nb = nb ? m + 1 : 0;
        }
    }
    out_gene[u] = -1; out_type[u] = 0;
    if (!fits) { atomicOr(status, 1u); return; }
    const int a_start = nb ? bs[0] : 0, a_end = nb ? be[nb - 1] : 0;      // an empty collection starts and ends at 0

    // read strand (bam.py:363-380): 0 '+', 1 '-', -1 not tested
    int rs = -1;
    if (strand_mode != 2) {
        const bool reverse = (flag & 0x10) != 0;
        bool minus = reverse;                                    // single end, forward protocol: '-' iff reverse
        if ((flag & 0x1) && (flag & 0x40)) minus = !reverse;     // read 1 of a pair is the antisense mate
        if (strand_mode == 1) minus = !minus;
        rs = minus ? 1 : 0;
    }

    auto evaluate = [&](int g, bool &exon_only, bool &exon_overlap, bool &intron_overlap) {
        exon_only = exon_overlap = intron_overlap = false;
        for (int tr = __ldg(t.gene_tr_off + g); tr < __ldg(t.gene_tr_off + g + 1); ++tr) {
            const int e0 = __ldg(t.tr_exon_off + tr), e1 = __ldg(t.tr_exon_off + tr + 1);
            if (!exon_overlap) exon_overlap = blocks_overlap(bs, be, nb, t.exon_start, t.exon_end, e0, e1);
            if (exon_overlap && !exon_only) exon_only = blocks_subset(bs, be, nb, t.exon_start, t.exon_end, e0, e1);
            if (!intron_overlap)
                intron_overlap = blocks_overlap(bs, be, nb, t.intron_start, t.intron_end, __ldg(t.tr_intron_off + tr),
                                                __ldg(t.tr_intron_off + tr + 1));
            if (exon_only && intron_overlap) break;      // nothing left to learn
        }
    };
    auto type_of = [&](bool exon_only, bool intron_overlap) -> uint8_t {      // bam.py:219-224
        if (exon_only && (!strict || !intron_overlap)) return 1;     // spliced
        if (intron_overlap && (!strict || !exon_only)) return 2;     // unspliced
        return 3;                                                    // ambiguous
    };

    const int given = gx ? gx[u] : -1;
    if (given != -1) {
        // the reference looks the gene up among the genes of the read's contig: KeyError otherwise
        if (given < 0 || ref_id < 0 || ref_id >= t.n_ref || given < __ldg(t.contig_off + ref_id) ||
            given >= __ldg(t.contig_off + ref_id + 1)) {
            atomicOr(status, 2u);
            return;
        }
        bool eo, ev, iv;
        evaluate(given, eo, ev, iv);
        out_gene[u] = given; out_type[u] = type_of(eo, iv);
        return;
    }
    if (ref_id < 0 || ref_id >= t.n_ref) return;
    const int g0 = __ldg(t.contig_off + ref_id), g1 = __ldg(t.contig_off + ref_id + 1);
    int lo = g0, hi = g1;                      // genes [g0, lo) start at or before the alignment
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(t.gene_start + mid) <= a_start) lo = mid + 1; else hi = mid; }
    int found = -1;
    uint8_t found_type = 0;
    for (int g = lo - 1; g >= g0 && __ldg(t.gene_pmax + g) >= a_end; --g) {
        if (__ldg(t.gene_end + g) < a_end) continue;
        if (rs >= 0 && (int)__ldg(t.gene_strand + g) != rs) continue;
        bool eo, ev, iv;
        evaluate(g, eo, ev, iv);
        if (!ev && !iv) continue;
        if (found >= 0) { found = -2; break; }     // a second overlapping gene: the read is dropped (bam.py:208-209)
        found = g; found_type = type_of(eo, iv);
    }
    if (found >= 0) { out_gene[u] = found; out_type[u] = found_type; }
}

}  // namespace

extern "C" int dyn_velocity(dyn_ctx_t *ctx, const void *d_records, const uint32_t *d_rec_off, int64_t n_units,
                            const int32_t *d_gx, const dyn_annotation_t *ann, int strand_mode, int strict_exon_overlap,
                            int32_t *d_out_gene, uint8_t *d_out_type, uint32_t *d_status, void *stream) {
    DYN_ARG(ctx != nullptr, "null context");
    CtxEnter enter_(ctx, stream);
    DYN_ARG(n_units >= 0, "negative n_units");
    if (n_units == 0) return DYN_OK;
    DYN_ARG(d_records && d_rec_off && ann && d_out_gene && d_out_type && d_status, "null buffer");
    DYN_ARG(strand_mode >= 0 && strand_mode <= 2, "strand_mode must be 0 (forward), 1 (reverse) or 2 (unstranded)");
    DYN_ARG(ann->n_ref >= 0 && ann->d_contig_off, "bad annotation");
    DYN_ARG(ann->n_genes == 0 || (ann->d_gene_start && ann->d_gene_end && ann->d_gene_pmax && ann->d_gene_strand && ann->d_gene_tr_off &&
                                  ann->d_tr_exon_off && ann->d_tr_intron_off),
            "null annotation table");
    VelocityTables t;
    t.gene_start = ann->d_gene_start; t.gene_end = ann->d_gene_end; t.gene_pmax = ann->d_gene_pmax;
    t.gene_strand = ann->d_gene_strand; t.contig_off = ann->d_contig_off; t.n_ref = ann->n_ref;
    t.gene_tr_off = ann->d_gene_tr_off; t.tr_exon_off = ann->d_tr_exon_off; t.tr_intron_off = ann->d_tr_intron_off;
    t.exon_start = ann->d_exon_start; t.exon_end = ann->d_exon_end;
    t.intron_start = ann->d_intron_start; t.intron_end = ann->d_intron_end;
    velocity_kernel<<<(unsigned)((n_units + kBlock - 1) / kBlock), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint8_t *>(d_records), d_rec_off, n_units, d_gx, t, strand_mode, strict_exon_overlap ? 1 : 0, d_out_gene,
        d_out_type, d_status);
    DYN_CUDA(cudaGetLastError());
    return DYN_OK;
}
