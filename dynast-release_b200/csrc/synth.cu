// Synthetic scSLAM-seq-like read generator (bench / test utility, NOT part of the timed hot path).
//
// Generates packed read records (include/dynast_b200.h) directly in HBM for the BASELINE configs that
// are too large to build on the host (C2: 50 M reads).  The generative model is SURVEY.md section 8d /
// dynast/benchmarking/simulation.py:22-79: iid uniform reference, per-gene labelled fraction pi,
// labelled reads convert T>C (A>G for reverse-strand genes) per base w.p. p_c, every base is
// substituted w.p. 3*p_e, 0.1 % N calls, Phred 90 % in 36..40 / 10 % in 2..27, CIGAR mix
// 70 % LM / 18 % aMbNcM / 8 % sS(L-s)M / 2 % one 1-3 nt D / 2 % one 1-3 nt I.
// Every decision is a pure function of (seed, read index), so the sizing pass and the fill pass agree
// and any slice can be regenerated (or copied to the host and fed to the CPU oracle).
#include <cub/cub.cuh>

#include "common.cuh"

namespace {

struct SynthParams {
    unsigned long long seed;
    int read_len, n_barcodes, n_genes, umi_len;
    long long n_groups;       // UMI groups (reads are assigned to groups uniformly)
    double p_c, p_e, frac_n;
    const float *bc_cdf;      // [n_barcodes] cumulative weights, last == 1
    const float *gene_cdf;    // [n_genes]
    const float *gene_pi;     // [n_genes]
    const uint8_t *gene_strand;   // [n_genes] 0 '+', 1 '-'
    const int32_t *gene_contig;   // [n_genes]
    long long pos_span;
    int lean;                 // DYN_REC_LEAN layout
};

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}

struct Rng {
    unsigned long long key, ctr;
    __device__ Rng(unsigned long long seed, unsigned long long idx, unsigned long long stream)
        : key(mix64(seed ^ mix64(idx * 0x632be59bd9b4e019ull + stream))), ctr(0) {}
    __device__ __forceinline__ unsigned long long next() { return mix64(key + (ctr++) * 0xd1342543de82ef95ull); }
    __device__ __forceinline__ float uniform() { return (float)(next() >> 40) * (1.0f / 16777216.0f); }
    __device__ __forceinline__ int below(int n) { return (int)((next() >> 32) * (unsigned long long)n >> 32); }
};

__device__ int cdf_search(const float *cdf, int n, float u) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (cdf[mid] < u) lo = mid + 1; else hi = mid;
    }
    return lo;
}

constexpr int kMaxLen = 304;      // read_len limit of the generator
constexpr int kMaxTok = 320;

struct Keys {
    uint32_t barcode, gene;
    unsigned long long umi;
    int score;
    uint8_t strand;
};

// Generates one record into local buffers; returns total bytes (16-byte padded). If `out` != nullptr
// the record is written there.
__device__ size_t gen_record(const SynthParams &P, long long idx, uint8_t *out, Keys *keys) {
    const int L = P.read_len;
    Rng g(P.seed, (unsigned long long)idx, 1);
    // group -> (barcode, gene, umi)
    const unsigned long long grp = g.next() % (unsigned long long)P.n_groups;
    Rng gg(P.seed, grp, 2);
    const int bc = cdf_search(P.bc_cdf, P.n_barcodes, gg.uniform());
    const int gene = cdf_search(P.gene_cdf, P.n_genes, gg.uniform());
    const unsigned long long umi = gg.next() & ((P.umi_len >= 32) ? ~0ull : ((1ull << (2 * P.umi_len)) - 1));
    const bool rev_gene = P.gene_strand[gene] != 0;
    // pos_span < 0 (config C3, consensus input): the reads of a UMI group are fragments of ONE molecule -- they start within
    // 300 nt of the group's anchor, the reference base is a function of (contig, position), the molecule is labelled or not
    // as a whole and carries its conversions at fixed positions; only sequencing errors, N calls and qualities are per read
    const bool clustered = P.pos_span < 0;
    const int contig = P.gene_contig[gene];
    int32_t pos0 = 0;
    if (clustered) {
        Rng gp(P.seed, (unsigned long long)idx, 3);
        pos0 = (int32_t)(mix64(grp ^ 0x5bd1e995ull) % (unsigned long long)(-P.pos_span)) + (int32_t)(gp.next() % 300);
    }
    bool labelled = g.uniform() < P.gene_pi[gene];
    if (clustered) labelled = Rng(P.seed, grp, 4).uniform() < P.gene_pi[gene];
    const unsigned long long ref_key = mix64(P.seed ^ ((unsigned long long)(uint32_t)contig << 40));
    const unsigned long long mol_key = mix64(P.seed ^ mix64(grp + 0x1234567ull));
    long long rp = pos0;
    const float u_kind = g.uniform();
    uint32_t cig[3];
    int n_cig;
    if (u_kind < 0.70f) { cig[0] = (uint32_t)L << 4; n_cig = 1; }
    else if (u_kind < 0.88f) {
        const int a = 10 + g.below(L - 20);
        cig[0] = (uint32_t)a << 4; cig[1] = ((uint32_t)(50 + g.below(4950)) << 4) | 3; cig[2] = (uint32_t)(L - a) << 4; n_cig = 3;
    } else if (u_kind < 0.96f) {
        const int s = 1 + g.below(19);
        cig[0] = ((uint32_t)s << 4) | 4; cig[1] = (uint32_t)(L - s) << 4; n_cig = 2;
    } else if (u_kind < 0.98f) {
        const int a = 10 + g.below(L - 20);
        cig[0] = (uint32_t)a << 4; cig[1] = ((uint32_t)(1 + g.below(3)) << 4) | 2; cig[2] = (uint32_t)(L - a) << 4; n_cig = 3;
    } else {
        const int a = 10 + g.below(L - 20), ins = 1 + g.below(3);
        cig[0] = (uint32_t)a << 4; cig[1] = ((uint32_t)ins << 4) | 1; cig[2] = (uint32_t)(L - a - ins) << 4; n_cig = 3;
    }
    uint8_t seq[(kMaxLen + 1) / 2 + 4];
    uint8_t qual[kMaxLen];
    uint32_t tok[kMaxTok];
    for (int i = 0; i < (L + 1) / 2 + 4; ++i) seq[i] = 0;
    int qp = 0, n_tok = 0, aligned = 0, nm = 0;
    const int conv_from = rev_gene ? 0 : 3, conv_to = rev_gene ? 2 : 1;
    const float pc = (float)P.p_c, pe3 = (float)(3.0 * P.p_e), fn = (float)P.frac_n;
    const uint8_t code[4] = {1, 2, 4, 8};
    for (int ci = 0; ci < n_cig; ++ci) {
        const int op = cig[ci] & 0xf, len = (int)(cig[ci] >> 4);
        if (op == 0) {
            for (int j = 0; j < len; ++j) {
                const unsigned long long r = g.next();
                const int gb = clustered ? (int)(mix64(ref_key + (unsigned long long)rp) & 3) : (int)(r & 3);
                int b = gb;
                const float u1 = clustered ? (float)(mix64(mol_key ^ (unsigned long long)rp) >> 40) * (1.0f / 16777216.0f)
                                           : (float)((r >> 8) & 0xffffff) * (1.0f / 16777216.0f);
                ++rp;
                const float u2 = (float)((r >> 32) & 0xffffff) * (1.0f / 16777216.0f);
                if (labelled && gb == conv_from && u1 < pc) b = conv_to;
                if (u2 < pe3) b = (gb + 1 + (int)(u2 / (float)P.p_e)) & 3;
                const unsigned long long r2 = g.next();
                const float u3 = (float)(r2 & 0xffffff) * (1.0f / 16777216.0f);
                const float u4 = (float)((r2 >> 24) & 0xffffff) * (1.0f / 16777216.0f);
                qual[qp] = (u4 < 0.9f) ? (uint8_t)(36 + (int)((r2 >> 48) % 5)) : (uint8_t)(2 + (int)((r2 >> 48) % 26));
                const bool is_n = u3 < fn;
                const uint8_t nib = is_n ? 15 : code[b];
                seq[qp >> 1] |= nib << ((qp & 1) ? 0 : 4);
                if (is_n || b != gb) {
                    if (n_tok < kMaxTok)
                        tok[n_tok++] = (uint32_t)aligned | ((uint32_t)code[gb] << 16) | (P.lean ? (uint32_t)qual[qp] << 21 : 0u);
                    ++nm;
                }
                ++aligned;
                ++qp;
            }
        } else if (op == 2) {
            for (int j = 0; j < len; ++j) {
                const int db = clustered ? (int)(mix64(ref_key + (unsigned long long)rp) & 3) : (int)(g.next() & 3);
                ++rp;
                if (n_tok < kMaxTok) tok[n_tok++] = (uint32_t)aligned | ((uint32_t)code[db] << 16) | (1u << 20);
            }
        } else if (op == 3) {
            rp += len;
        } else if (op == 1 || op == 4) {
            for (int j = 0; j < len; ++j) {
                const unsigned long long r = g.next();
                seq[qp >> 1] |= code[r & 3] << ((qp & 1) ? 0 : 4);
                qual[qp] = (uint8_t)(30 + (int)((r >> 8) % 10));
                ++qp;
            }
        }
    }
    const int seq_bytes = (((L + 1) >> 1) + 3) & ~3;
    const size_t raw = 16 + 4 * (size_t)n_cig + 4 * (size_t)n_tok + seq_bytes + (P.lean ? 0 : L);
    const size_t bytes = (raw + 15) & ~(size_t)15;
    if (keys) {
        keys->barcode = (uint32_t)bc; keys->gene = (uint32_t)gene; keys->umi = umi;
        keys->score = L - 2 * nm; keys->strand = rev_gene ? 1 : 0;
    }
    if (out) {
        dyn_read_hdr_t h;
        // pos_span > 0: reads laid out in index order over the span (a coordinate-sorted BAM);
        // pos_span < 0: the reads of a UMI group start within 300 nt of the group's anchor (config C3: consensus input)
        if (P.pos_span > 0) h.pos = (int32_t)(((unsigned long long)idx * (unsigned long long)P.pos_span) / 0x100000000ull);
        else h.pos = pos0;
        h.ref_id = contig;
        h.l_seq = (uint16_t)L; h.n_cigar = (uint16_t)n_cig; h.n_tok = (uint16_t)n_tok;
        h.flag = (uint16_t)(((g.next() & 1) ? 16 : 0) | (P.lean ? DYN_REC_LEAN : 0));
        *reinterpret_cast<uint4 *>(out) = *reinterpret_cast<uint4 *>(&h);
        uint8_t *p = out + 16;
        for (int i = 0; i < n_cig; ++i) reinterpret_cast<uint32_t *>(p)[i] = cig[i];
        p += 4 * n_cig;
        for (int i = 0; i < n_tok; ++i) reinterpret_cast<uint32_t *>(p)[i] = tok[i];
        p += 4 * n_tok;
        for (int i = 0; i < seq_bytes; ++i) p[i] = seq[i];
        p += seq_bytes;
        if (!P.lean) {
            for (int i = 0; i < L; ++i) p[i] = qual[i];
            p += L;
        }
        for (size_t i = raw; i < bytes; ++i) *p++ = 0;
    }
    return bytes;
}

__global__ void synth_size_kernel(SynthParams P, long long n, uint32_t *sizes) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    sizes[i] = (uint32_t)(gen_record(P, i, nullptr, nullptr) >> 4);
}

__global__ void synth_fill_kernel(SynthParams P, long long n, const uint32_t *rec_off, uint8_t *records,
                                  uint32_t *barcode, unsigned long long *umi, uint32_t *gene, int32_t *score,
                                  uint8_t *strand) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    Keys k;
    gen_record(P, i, records + ((size_t)rec_off[i] << 4), &k);
    if (barcode) barcode[i] = k.barcode;
    if (umi) umi[i] = k.umi;
    if (gene) gene[i] = k.gene;
    if (score) score[i] = k.score;
    if (strand) strand[i] = k.strand;
}

}  // namespace

static int make_params(const dyn_synth_params_t *sp, SynthParams *P) {
    DYN_ARG(sp != nullptr, "null params");
    DYN_ARG(sp->read_len >= 30 && sp->read_len <= kMaxLen, "read_len must be in [30, 304]");
    DYN_ARG(sp->n_barcodes > 0 && sp->n_genes > 0 && sp->n_groups > 0, "empty tables");
    DYN_ARG(sp->d_bc_cdf && sp->d_gene_cdf && sp->d_gene_pi && sp->d_gene_strand && sp->d_gene_contig, "null table");
    P->seed = sp->seed; P->read_len = sp->read_len; P->n_barcodes = sp->n_barcodes; P->n_genes = sp->n_genes;
    P->umi_len = sp->umi_len; P->n_groups = sp->n_groups; P->p_c = sp->p_c; P->p_e = sp->p_e; P->frac_n = sp->frac_n;
    P->bc_cdf = sp->d_bc_cdf; P->gene_cdf = sp->d_gene_cdf; P->gene_pi = sp->d_gene_pi;
    P->gene_strand = sp->d_gene_strand; P->gene_contig = sp->d_gene_contig;
    P->pos_span = sp->pos_span != 0 ? sp->pos_span : (1ll << 27);
    P->lean = sp->lean != 0;
    return DYN_OK;
}

// Pass 1: d_rec_off[0..n] = exclusive prefix sum of record sizes (16-byte units).
extern "C" int dyn_synth_offsets(dyn_ctx_t *ctx, const dyn_synth_params_t *sp, int64_t n_reads, uint32_t *d_rec_off,
                                 void *stream_) {
    DYN_ARG(ctx && d_rec_off && n_reads > 0, "bad argument");
    CtxEnter enter_(ctx, stream_);
    SynthParams P;
    int rc = make_params(sp, &P);
    if (rc) return rc;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    void *sizes = nullptr;
    if ((rc = dyn_ctx_scratch(ctx, 1, sizeof(uint32_t) * (size_t)(n_reads + 1), &sizes))) return rc;
    uint32_t *d_sizes = static_cast<uint32_t *>(sizes);
    DYN_CUDA(cudaMemsetAsync(d_sizes + n_reads, 0, 4, stream));
    synth_size_kernel<<<(unsigned)((n_reads + 127) / 128), 128, 0, stream>>>(P, n_reads, d_sizes);
    DYN_CUDA(cudaGetLastError());
    size_t tmp_bytes = 0;
    DYN_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_sizes, d_rec_off, (int)(n_reads + 1), stream));
    void *tmp = nullptr;
    if ((rc = dyn_ctx_scratch(ctx, 2, tmp_bytes, &tmp))) return rc;
    DYN_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, d_sizes, d_rec_off, (int)(n_reads + 1), stream));
    return DYN_OK;
}

// Pass 2: write the records and the per-read keys (any key pointer may be NULL).
extern "C" int dyn_synth_fill(dyn_ctx_t *ctx, const dyn_synth_params_t *sp, int64_t n_reads, const uint32_t *d_rec_off,
                              void *d_records, uint32_t *d_barcode, uint64_t *d_umi, uint32_t *d_gene, int32_t *d_score,
                              uint8_t *d_strand, void *stream_) {
    DYN_ARG(ctx && d_rec_off && d_records && n_reads > 0, "bad argument");
    SynthParams P;
    int rc = make_params(sp, &P);
    if (rc) return rc;
    synth_fill_kernel<<<(unsigned)((n_reads + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream_)>>>(
        P, n_reads, d_rec_off, static_cast<uint8_t *>(d_records), d_barcode, reinterpret_cast<unsigned long long *>(d_umi),
        d_gene, d_score, d_strand);
    DYN_CUDA(cudaGetLastError());
    return DYN_OK;
}
