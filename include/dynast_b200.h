/* dynast_b200 -- C-ABI of the B200-native count -> estimate hot path.
 *
 * dynast (the reference) has no FFI layer: its boundary is the Python operator API
 * (dynast/preprocessing/__init__.py:2-31, dynast/estimation/__init__.py:2-10).  The functions below are
 * what those operators bind to in this build (ctypes); every entry point names the
 * reference function whose inner loop it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types.
 *   - pointers named d_* are DEVICE pointers (any allocator: cudaMalloc, torch, ...), h_* are HOST
 *     pointers (pinned for best throughput).  The caller owns every buffer.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Device-pointer entry
 *     points are asynchronous on that stream; host-pointer entry points synchronise before returning.
 *   - return value: 0 = ok, negative = error (dyn_last_error() returns a thread-local message).
 *   - there is no CPU fallback anywhere in this library.
 */
#ifndef DYNAST_B200_H
#define DYNAST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DYN_OK 0
#define DYN_ERR_CUDA (-1)
#define DYN_ERR_ARG (-2)
#define DYN_ERR_NOMEM (-3)
#define DYN_ERR_FORMAT (-4)
#define DYN_ERR_IO (-5)

/* ------------------------------------------------------------------------------------------------
 * Packed read records (what the host BAM decoder writes into pinned memory and the tally reads).
 *
 * Read i occupies bytes [16*rec_off[i], 16*rec_off[i+1]) of the record blob:
 *     dyn_read_hdr_t hdr;                      16 bytes
 *     uint32_t cigar[n_cigar];                 BAM encoding: len << 4 | op   (MIDNSHP=X -> 0..8)
 *     uint32_t tok[n_tok];                     the MD tag, tokenised by the packer (text -> binary, nothing
 *                                              resolved): one token per MD letter, in MD order --
 *                                              bits 0..15 aligned (M/=/X) offset at which it occurs,
 *                                              bits 16..19 reference base (BAM 4-bit code), bit 20 set for a
 *                                              deleted reference base (a letter of a ^XYZ run)
 *     uint8_t  seq[(l_seq + 1) / 2];           BAM 4-bit codes "=ACMGRSVTWYHKDBN", high nibble first;
 *                                              zero-padded to a multiple of 4 bytes
 *     uint8_t  qual[l_seq];                    Phred (absent in a lean record, see DYN_REC_LEAN)
 *     zero padding to a multiple of 16 bytes
 * The packer checks the MD tag against the CIGAR (matched + mismatched bases == M/=/X length, deleted
 * letters == D length) and sets DYN_REC_BAD_MD otherwise (pysam raises on such a record).
 * If hdr.flag has DYN_REC_HAS_MATE, a second record (the first-seen mate of a read pair, same layout)
 * follows inside the same [rec_off[i], rec_off[i+1]) range; the pair is one counting unit
 * (bam.py:307-352) and dyn_tally_reads applies the overlap rules of bam.py:332-346, 393-419.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
    int32_t pos;      /* 0-based leftmost reference position (BAM pos) */
    int32_t ref_id;   /* contig index in the BAM header */
    uint16_t l_seq;   /* query length including soft clips */
    uint16_t n_cigar;
    uint16_t n_tok;   /* number of MD tokens */
    uint16_t flag;    /* BAM flag bits 0..11 | DYN_REC_* */
} dyn_read_hdr_t;

#define DYN_REC_HAS_MATE 0x8000u
#define DYN_REC_BAD_MD 0x4000u
/* Lean record (single-end units only): the qual[] stream is left out and every non-deleted MD token carries, in
 * bits 21..28, the Phred of the query base at its aligned offset -- the only qualities a single-end read's tally
 * needs (bam.py:391-411, conversion.py:424).  Half the bytes per read.  Paired units keep the full layout (the mate
 * rule of bam.py:395-403 compares qualities at positions where only one mate mismatches), and so does anything
 * that goes to dyn_consensus (it sums every base's quality). */
#define DYN_REC_LEAN 0x2000u
#define DYN_TOK_AOFF(tok) ((uint32_t)(tok) & 0xffffu)
#define DYN_TOK_BASE(tok) (((uint32_t)(tok) >> 16) & 0xfu)
#define DYN_TOK_IS_DEL(tok) (((uint32_t)(tok) >> 20) & 1u)
#define DYN_TOK_QUAL(tok) (((uint32_t)(tok) >> 21) & 0xffu)   /* lean records only */

/* Conversion record codes: (ref nibble << 4) | read nibble, both BAM 4-bit codes (A=1,C=2,G=4,T=8).
 * One 64-bit record per mismatch: bits 0..31 genome_i (0-based), 32..39 code, 40..47 Phred,
 * 48..63 reserved (0). */
#define DYN_CONV_POS(rec) ((uint32_t)((rec) & 0xffffffffu))
#define DYN_CONV_CODE(rec) ((uint8_t)(((rec) >> 32) & 0xff))
#define DYN_CONV_QUAL(rec) ((uint8_t)(((rec) >> 40) & 0xff))

/* Column order of the 16 per-read counters == conversion.py:44-63 (CONVERSION_IDX, BASE_IDX):
 * AC AG AT CA CG CT GA GC GT TA TC TG A C G T */
#define DYN_N_COUNTERS 16

/* tally flags */
#define DYN_TALLY_NASC 0x1u        /* --nasc: drop mate conversions inside the pair overlap (bam.py:339,396) */
#define DYN_TALLY_EMIT_CONV 0x2u   /* also emit the per-mismatch records (conversions.csv rows) */
/* optional tuning hint in flags bits 16..31: mean size of a counting unit in 16-byte units (sizes the
 * shared-memory staging buffer; 0 = assume ~91-nt single-end reads). Results never depend on it. */
#define DYN_TALLY_REC16_HINT(units16) ((uint32_t)(units16) << 16)

/* status bits accumulated in *d_status by the tally (0 == clean) */
#define DYN_ST_COUNTER_OVERFLOW 0x1u /* a counter exceeded 255: the reference raises when re-reading uint8 */
#define DYN_ST_BAD_RECORD 0x2u       /* CIGAR / MD / seq disagree (pysam would raise) */
#define DYN_ST_NON_ACGT_CONV 0x4u    /* a quality-passing mismatch involves an IUPAC code (reference: KeyError) */
#define DYN_ST_CONV_CAPACITY 0x8u    /* conversion-record capacity exceeded */

typedef struct dyn_ctx dyn_ctx_t;

const char *dyn_last_error(void);
int dyn_version(void);

/* One context per GPU (per process). Owns scratch buffers and an internal stream pool.
 * Threading / stream contract: a context runs ONE call at a time -- every entry point takes the context's lock for
 * its whole duration, so concurrent host threads serialise -- and its scratch buffers follow stream order: when two
 * consecutive calls that share a scratch buffer come on different streams, the later stream is made to wait
 * (cudaStreamWaitEvent) for everything the earlier call queued.  Use one context per stream for calls that
 * should overlap on the device. */
int dyn_ctx_create(int device, dyn_ctx_t **ctx);
int dyn_ctx_destroy(dyn_ctx_t *ctx);
int dyn_ctx_sm_count(dyn_ctx_t *ctx);

/* ------------------------------------------------------------------------------------------------
 * a1 + a3: per-read conversion tally.
 * Replaces the inner loops of bam.parse_read_contig (bam.py:357-429: CIGAR/MD walk, reference base
 * content, mismatch extraction) and conversion.count_conversions_part (conversion.py:403-432: quality
 * and SNP filter, 12 conversion + 4 content counters).
 *
 *   d_records, d_rec_off      packed records (see above); rec_off has n_reads + 1 entries
 *   d_snps, n_snps            sorted unique uint64 keys (conv_idx << 56 | ref_id << 32 | genome_i) of
 *                             positions to ignore per conversion (conversion.py:395-398); may be NULL/0
 *   quality                   count a mismatch only if Phred > quality (strict, conversion.py:424)
 *   d_counts [n_reads][16] u8 output counters, column order above
 *   d_conv_off [n_reads] u32, d_conv_n [n_reads] u16
 *                             where read i's conversion records live in d_conv_rec / how many
 *                             (all mismatches, NOT quality filtered: conversions.csv semantics, bam.py:405-419)
 *   d_conv_rec [conv_capacity] u64, d_conv_read [conv_capacity] u32
 *                             records and their owning read index; within a read in walk order
 *   d_conv_total              u64: number of records written (device scalar)
 *   d_status                  u32: DYN_ST_* bits (device scalar, OR-accumulated; caller zeroes)
 * The conv outputs may be NULL when DYN_TALLY_EMIT_CONV is not set.
 * ------------------------------------------------------------------------------------------------ */
int dyn_tally_reads(dyn_ctx_t *ctx, const void *d_records, const uint32_t *d_rec_off, int64_t n_reads,
                    const uint64_t *d_snps, int64_t n_snps, int quality, uint32_t flags, uint8_t *d_counts,
                    uint32_t *d_conv_off, uint16_t *d_conv_n, uint64_t *d_conv_rec, uint32_t *d_conv_read,
                    int64_t conv_capacity, uint64_t *d_conv_total, uint32_t *d_status, void *stream);

/* Host-buffer form of the same call (the end-to-end path): records are streamed from pinned host
 * memory in chunks with cudaMemcpyAsync on two streams, results are copied back.  All pointers are
 * host pointers; h_conv_* may be NULL. Returns after everything has landed. */
int dyn_tally_reads_host(dyn_ctx_t *ctx, const void *h_records, const uint32_t *h_rec_off, int64_t n_reads,
                         const uint64_t *h_snps, int64_t n_snps, int quality, uint32_t flags, uint8_t *h_counts,
                         uint32_t *h_conv_off, uint16_t *h_conv_n, uint64_t *h_conv_rec, uint32_t *h_conv_read,
                         int64_t conv_capacity, uint64_t *h_conv_total, uint32_t *h_status);

/* ------------------------------------------------------------------------------------------------
 * a12-a14: per-barcode binomial-mixture EM for p_c.
 * Replaces p_c.expectation_maximization / expectation_maximization_nasc (p_c.py:49-172) as driven per
 * barcode by estimate_p_c (p_c.py:206-227).
 *
 * Barcode b owns histogram triplets [d_off[b], d_off[b+1]) of (k = conversions, n = base content,
 * count = reads); duplicate (k, n) are summed (csr_matrix semantics, p_c.py:219).
 *   d_p_e [B] f64   background rate per barcode
 *   nasc            0: expectation_maximization, 1: expectation_maximization_nasc
 *   d_p_c [B] f64   result; d_em_status [B] i32: 0 ok, 1 'p_c <= p_e' (p_c.py:169), 2 arithmetic
 *                   failure inside the reference (ZeroDivisionError), 3 shape not supported
 *   d_iters [B] i32 iterations used (may be NULL)
 * ------------------------------------------------------------------------------------------------ */
#define DYN_EM_OK 0
#define DYN_EM_PC_LE_PE 1
#define DYN_EM_ZERODIV 2
#define DYN_EM_UNSUPPORTED 3

int dyn_em_pc(dyn_ctx_t *ctx, const int32_t *d_k, const int32_t *d_n, const int64_t *d_count, const int64_t *d_off,
              int64_t n_barcodes, const double *d_p_e, int nasc, double *d_p_c, int32_t *d_em_status,
              int32_t *d_iters, void *stream);

int dyn_em_pc_host(dyn_ctx_t *ctx, const int32_t *h_k, const int32_t *h_n, const int64_t *h_count,
                   const int64_t *h_off, int64_t n_barcodes, const double *h_p_e, int nasc, double *h_p_c,
                   int32_t *h_em_status, int32_t *h_iters);

/* ------------------------------------------------------------------------------------------------
 * a3, file-based entry: conversion counters from mismatch records (the rows of conversions.csv), for
 * callers of count_conversions that hold the reference's CSV interchange rather than packed BAM records.
 * Replaces conversion.count_conversions_part's per-line loop (conversion.py:417-426).  d_counts [n_reads][16]
 * must hold zeros in columns 0..11 and the A/C/G/T content in 12..15; record layout = DYN_CONV_* above;
 * d_conv_contig = contig id of each record for the SNP key (may be NULL when n_snps == 0).
 * ------------------------------------------------------------------------------------------------ */
int dyn_count_records(dyn_ctx_t *ctx, const uint64_t *d_conv_rec, const uint32_t *d_conv_read,
                      const uint32_t *d_conv_contig, int64_t n_records, int64_t n_reads, const uint64_t *d_snps,
                      int64_t n_snps, int quality, uint8_t *d_counts, uint32_t *d_status, void *stream);

/* ------------------------------------------------------------------------------------------------
 * a4: strand complement (conversion.complement_counts, conversion.py:99-125), in place on the 16-byte
 * counter rows: for reads whose gene lies on '-' (d_strand[i] != 0) the 12 conversion columns and the
 * 4 base columns are reversed (AC<->TG, AG<->TC, ..., A<->T, C<->G).  An involution.  The row ORDER
 * change of the reference (forward rows first) is applied by dyn_dedup_umi, not here.
 * ------------------------------------------------------------------------------------------------ */
int dyn_complement_counts(dyn_ctx_t *ctx, uint8_t *d_counts, const uint8_t *d_strand, int64_t n_reads, void *stream);

/* ------------------------------------------------------------------------------------------------
 * a6 helper: per barcode, the smallest frame position of its reads and its read count.  Frame position
 * of read i = (strand << 33) | ((conv_n == 0) << 32) | i  -- the row order of the combined, complemented
 * counts frame (conversion.py:533-541).  The host turns this into the barcode order / split assignment of
 * conversion.py:543-579 (a sequential packing over <= n_barcodes entries).
 * ------------------------------------------------------------------------------------------------ */
int dyn_barcode_first_pos(dyn_ctx_t *ctx, const uint32_t *d_barcode, const uint8_t *d_strand, const uint16_t *d_conv_n,
                          int64_t n_reads, int64_t n_barcodes, uint64_t *d_first_pos, uint32_t *d_n_reads, void *stream);

/* Host half of the same step: HOST arrays in, HOST arrays out; h_ord_of_barcode = rank of the barcode in
 * split-file order, h_split_of_barcode = index of its split file, *h_n_splits = number of splits. */
int dyn_split_plan(const uint64_t *h_first_pos, const uint32_t *h_n_reads, int64_t n_barcodes, int64_t threshold,
                   uint32_t *h_ord_of_barcode, uint32_t *h_split_of_barcode, int64_t *h_n_splits);

/* ------------------------------------------------------------------------------------------------
 * a5 + a6: UMI deduplication and output ordering.
 * Replaces conversion.deduplicate_counts (conversion.py:203-243) as driven by count_conversions
 * (conversion.py:541-606).  Group key of read i = (d_key_hi[i], d_key_lo[i]) = interned (barcode, umi, GX)
 * (d_key_hi may be NULL; *_bits = number of significant low bits).  Priority, on COMPLEMENTED counters:
 * ([has_conversions,] transcriptome, score, conversion_sum) with conversion_sum over the columns of
 * conv_mask (bit j = column j of DYN_N_COUNTERS order); use_conversions adds has_conversions.
 *   d_counts            original-orientation rows (complemented on the fly with d_strand)
 *   d_conv_n            number of mismatch records of the read (0 -> the read sits in the second block)
 *   d_ord_of_barcode    [n_barcodes] rank of the barcode in split-file order; d_split_of_barcode its split
 *                       (both NULL: one split, barcodes unordered -- only valid for a single barcode)
 *   d_final_extra       optional per-read key (final_extra_bits wide) ordering rows inside a (split, strand)
 *                       block before the priority order: the no-UMI path (conversion.drop_multimappers,
 *                       conversion.py:148-200) groups by read id, use_conversions = 0, and sorts survivors
 *                       by (barcode, GX) -- pass the rank of that pair here; NULL for the UMI path
 *   d_out_idx [n_reads] surviving read indices in counts_{convs}.csv row order; *d_n_out their number
 * ------------------------------------------------------------------------------------------------ */
int dyn_dedup_umi(dyn_ctx_t *ctx, const uint64_t *d_key_hi, const uint64_t *d_key_lo, int key_hi_bits, int key_lo_bits,
                  const uint8_t *d_counts, const uint8_t *d_strand, const uint8_t *d_transcriptome,
                  const uint16_t *d_score, const uint16_t *d_conv_n, const uint32_t *d_barcode,
                  const uint32_t *d_ord_of_barcode, const uint32_t *d_split_of_barcode, int64_t n_barcodes,
                  int64_t n_splits, int64_t n_reads, uint32_t conv_mask, int use_conversions,
                  const uint64_t *d_final_extra, int final_extra_bits, uint32_t *d_out_idx, int64_t *d_n_out,
                  void *stream);

/* The UMI-dedup group key of every read from its interned ids: barcode << (umi_bits + gene_bits) | (umi &
 * mask(umi_bits)) << gene_bits | gene -- the (barcode, UMI, GX) triple conversion.py:225 groups on, as the single
 * 64-bit d_key_lo dyn_dedup_umi takes (key_lo_bits = barcode bits + umi_bits + gene_bits). */
int dyn_group_key(dyn_ctx_t *ctx, const uint32_t *d_barcode, const uint64_t *d_umi, const uint32_t *d_gene, int64_t n,
                  int umi_bits, int gene_bits, uint64_t *d_key, void *stream);

/* ------------------------------------------------------------------------------------------------
 * a9 / a11 / a16: per-group reductions over selected reads (d_sel: read indices, NULL = all reads).
 * dyn_group_sums: 16-column sums (aggregation.py:74-79, p_e.py:72-76), number of reads and number of
 * reads with any conversion of conv_mask (alpha.py:72-75) per group.  d_strand != NULL complements.
 * dyn_rates_pe: rates[g][j] = sum_j / sum_base(j) (aggregation.py:82-85) and p_e[g] = NaN-skipping mean of
 * the rates selected by pe_mask (p_e.py:82-93); sums are truncated to uint32 as the reference does.
 * dyn_alpha: alpha[g] = (n_with / n_reads) / pi_c[g], NaN where pi_c <= 0 or missing (alpha.py:76-84).
 * ------------------------------------------------------------------------------------------------ */
int dyn_group_sums(dyn_ctx_t *ctx, const uint8_t *d_counts, const uint8_t *d_strand, const uint32_t *d_group,
                   const uint32_t *d_sel, int64_t n_sel, int64_t n_groups, uint32_t conv_mask, uint64_t *d_sums,
                   uint64_t *d_n_reads, uint64_t *d_n_with, void *stream);
int dyn_rates_pe(dyn_ctx_t *ctx, const uint64_t *d_sums, int64_t n_groups, uint32_t pe_mask, double *d_rates,
                 double *d_p_e, void *stream);
int dyn_alpha(dyn_ctx_t *ctx, const uint64_t *d_n_reads, const uint64_t *d_n_with, const double *d_pi_c,
              int64_t n_groups, double *d_alpha, void *stream);

/* ------------------------------------------------------------------------------------------------
 * a10: (group, gene, k, n) read counts (aggregation.aggregate_counts, aggregation.py:92-119).
 * k = sum of the conv_mask columns, n = sum of the base_mask columns (bit j = base column 12 + j); reads
 * with a non-zero exclude_mask column are skipped (estimate.py:248-255).  Row key =
 * ((group << gene_bits | gene) << 22) | (k << 10) | n.  Rows come out in first-appearance order
 * (groupby(sort=False)) or, with first_appearance_order == 0, sorted by key.  d_first = position (in the
 * selection) of the first read of the row.  Outputs need room for n_sel rows.
 * group_bits / gene_bits = bits of the largest group / gene id: the sort is restricted to them (group_bits == 0:
 * unknown, all 64 bits are sorted; ids that do not fit the stated bits give wrong rows).
 * dyn_kn_csr turns key-sorted rows with gene_bits == 0 into the EM / pi input: triplets (k, n, count),
 * offsets d_off[n_groups + 1] and per-group read totals.
 * ------------------------------------------------------------------------------------------------ */
int dyn_aggregate_kn(dyn_ctx_t *ctx, const uint8_t *d_counts, const uint8_t *d_strand, const uint32_t *d_group,
                     const uint32_t *d_gene, const uint32_t *d_sel, int64_t n_sel, uint32_t conv_mask,
                     uint32_t base_mask, uint32_t exclude_mask, int group_bits, int gene_bits,
                     int first_appearance_order, uint64_t *d_keys, uint32_t *d_count, uint32_t *d_first,
                     int64_t *d_n_rows, void *stream);
int dyn_kn_csr(dyn_ctx_t *ctx, const uint64_t *d_keys, const uint32_t *d_count, const int64_t *d_n_rows,
               int64_t max_rows, int64_t n_groups, int32_t *d_k, int32_t *d_n, int64_t *d_cnt, int64_t *d_off,
               uint64_t *d_total, void *stream);
/* The same for key-sorted rows WITH a gene field: the input of the per-(cell, gene) fit (estimate_pi with group_by =
 * [barcode, GX], pi.py:253-296; estimate.py:282-327 method 'pi_g').  Only the (group, gene) pairs that occur get a CSR
 * row: pair j has key d_pair_key[j] = group << gene_bits | gene, triplets [d_off[j], d_off[j + 1]) and d_total[j]
 * reads; *d_n_pairs pairs in all (device scalar).  Outputs need room for max_rows (+ 1 for d_off) entries. */
int dyn_kn_csr_pairs(dyn_ctx_t *ctx, const uint64_t *d_keys, const uint32_t *d_count, const int64_t *d_n_rows,
                     int64_t max_rows, int32_t *d_k, int32_t *d_n, int64_t *d_cnt, int64_t *d_off, uint64_t *d_pair_key,
                     uint64_t *d_total, int64_t *d_n_pairs, void *stream);

/* ------------------------------------------------------------------------------------------------
 * a15: batched labeled-fraction fit -- posterior means of (alpha, beta, pi_g) of dynast/models/pi.stan
 * (pi.stan:4-34) for every group, replacing pi.fit_stan_mcmc (pi.py:118-187).  Group g owns triplets
 * [d_off[g], d_off[g+1]) of (k, n, count) and its own p_e / p_c.  Deterministic quadrature, see csrc/pi.cu.
 * d_status: 0 ok, 1 empty group, 2 degenerate (p outside (0,1) / zero evidence), 3 more than max_cells.
 * ------------------------------------------------------------------------------------------------ */
int dyn_pi_fit(dyn_ctx_t *ctx, const int32_t *d_k, const int32_t *d_n, const int64_t *d_count, const int64_t *d_off,
               int64_t n_groups, int64_t max_cells, const double *d_p_e, const double *d_p_c, double *d_alpha_beta_pi,
               int32_t *d_status, void *stream);

/* ------------------------------------------------------------------------------------------------
 * a8: UMI consensus.  Replaces consensus.call_consensus_from_reads (dynast/preprocessing/consensus.py:57-200)
 * for every (gene, barcode, UMI) group at once.  Group g consists of the member records
 * d_item_rec16[d_group_off[g] .. d_group_off[g+1]) (offsets into d_records in 16-byte units; the two mates of
 * a pair are two items); the grouping itself (consensus.py:338-430) is done by the caller.
 * Output: one packed record per group (same layout as the input records: pos = leftmost start, CIGAR with
 * M / D / N runs, MD tokens, 4-bit consensus sequence, qualities clipped to 42; flag 0) at
 * d_out_records + 16 * d_out_off[g]; the MD tag TEXT exactly as consensus.py:124-161, 175 builds it at
 * d_out_md + d_out_md_off[g] (bytes, no terminator); NM in d_out_nm[g].  d_out_status[g]: 0 ok, 2 malformed member record,
 * 3 does not fit a BAM record, 4 members on several contigs (the reference raises), 5 output capacity,
 * 6 a member holds a read or reference base other than A/C/G/T/N (the reference raises KeyError on BASE_IDX,
 * consensus.py:84-97).
 * The call synchronises the stream once (it sizes its tuple buffers from a device total).
 * ------------------------------------------------------------------------------------------------ */
int dyn_consensus(dyn_ctx_t *ctx, const void *d_records, const uint32_t *d_item_rec16, const int64_t *d_group_off,
                  int64_t n_groups, int64_t n_items, int quality, void *d_out_records, int64_t out_capacity_bytes,
                  uint32_t *d_out_off, void *d_out_md, int64_t md_capacity_bytes, uint32_t *d_out_md_off,
                  uint32_t *d_out_nm, int32_t *d_out_status, void *stream);

/* ------------------------------------------------------------------------------------------------
 * a7: coverage of positions of interest and grouped key counts.
 * dyn_coverage replaces coverage.calculate_coverage_contig (dynast/preprocessing/coverage.py:101-149):
 * d_cov[i] = number of selected counting units (d_selected[u] != 0, NULL = all) whose aligned (M/=/X)
 * positions include position-of-interest i; d_poi = sorted unique keys (ref_id << 32 | position); a pair
 * counts the union of its mates' positions once (coverage.py:136-142).  d_first[i] = smallest
 * (unit << 32 | position) that touched entry i (all ones if untouched): coverage.csv lists positions in that
 * order (coverage.py:144-149).
 * dyn_unique_counts: distinct 64-bit keys of an array (key ~0 = skip) with their multiplicity and first
 * occurrence, sorted by key -- the counting dictionaries of snp.extract_conversions_part (snp.py:108-133).
 * ------------------------------------------------------------------------------------------------ */
int dyn_coverage(dyn_ctx_t *ctx, const void *d_records, const uint32_t *d_rec_off, int64_t n_units,
                 const uint8_t *d_selected, const uint64_t *d_poi, int64_t n_poi, uint32_t *d_cov, uint64_t *d_first,
                 void *stream);
int dyn_unique_counts(dyn_ctx_t *ctx, const uint64_t *d_keys, int64_t n_keys, uint64_t *d_out_keys, uint32_t *d_out_count,
                      uint32_t *d_out_first, int64_t *d_n_out, void *stream);

/* ------------------------------------------------------------------------------------------------
 * f2: velocity assignment.  Replaces bam.assign_velocity_type (dynast/preprocessing/bam.py:177-228, built on
 * ngs_tools.gtf.SegmentCollection) and the read-strand rule of bam.py:363-380 for every counting unit at once.
 * The annotation is the flat form of (gene_infos, transcript_infos) -- ngs_tools.gtf.genes_and_transcripts_from_gtf,
 * count.py:146-147 -- with 0-based half-open coordinates: genes grouped by contig (BAM header order) and sorted by
 * (start, end) inside a contig (bam.py:177-178), exons and introns of every transcript sorted and collapsed.
 *   d_gx [n_units] i32    gene row of the unit's gene tag; -1 = no tag: the gene is searched among the genes of the
 *                         unit's contig that contain the alignment and lie on the read's strand; any other
 *                         negative value or a gene of another contig = the reference's KeyError (status bit 2).
 *                         NULL = no unit carries a gene
 *   strand_mode           0 forward, 1 reverse, 2 unstranded (--strand)
 *   d_out_gene [n_units]  assigned gene row, -1 when the unit is dropped (bam.py:382-383)
 *   d_out_type [n_units]  0 dropped, 1 spliced, 2 unspliced, 3 ambiguous
 *   d_status              u32, OR-accumulated: 1 a unit has more aligned blocks than the kernel holds (48),
 *                         2 gene tag unknown to the annotation of the unit's contig
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
    int32_t n_ref, n_genes;
    const int32_t *d_contig_off;      /* [n_ref + 1] genes of contig c: rows [d_contig_off[c], d_contig_off[c + 1]) */
    const int32_t *d_gene_start, *d_gene_end;   /* [n_genes] */
    const int32_t *d_gene_pmax;       /* [n_genes] running maximum of d_gene_end inside the contig */
    const uint8_t *d_gene_strand;     /* [n_genes] 0 '+', 1 '-', 2 other */
    const int32_t *d_gene_tr_off;     /* [n_genes + 1] transcripts of a gene */
    const int32_t *d_tr_exon_off, *d_tr_intron_off;   /* [n_transcripts + 1] */
    const int32_t *d_exon_start, *d_exon_end, *d_intron_start, *d_intron_end;
} dyn_annotation_t;

#define DYN_VELOCITY_NONE 0
#define DYN_VELOCITY_SPLICED 1
#define DYN_VELOCITY_UNSPLICED 2
#define DYN_VELOCITY_AMBIGUOUS 3
int dyn_velocity(dyn_ctx_t *ctx, const void *d_records, const uint32_t *d_rec_off, int64_t n_units, const int32_t *d_gx,
                 const dyn_annotation_t *annotation, int strand_mode, int strict_exon_overlap, int32_t *d_out_gene,
                 uint8_t *d_out_type, uint32_t *d_status, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Per-read exchange to the barcode owner (multi-GPU).  The reference runs one process over the whole
 * conversions.csv (conversion.py:529-606); with read-range sharding a UMI group can straddle ranks, so before
 * dyn_dedup_umi every rank sends its reads' fixed-size rows to rank barcode % world.
 *   dyn_owner_partition: d_perm = the local read indices grouped by owner (stable), d_owner_counts[world] = reads
 *                        per owner (the all-to-all split sizes)
 *   dyn_pack_reads:      row i = the columns of read d_perm[i] (d_perm NULL: read i), DYN_READ_ROW_BYTES each:
 *                        counts[16] | umi u64 | barcode u32 | gene u32 | score i16 | conv_n u16 | strand u8 |
 *                        transcriptome u8 | 2 pad
 *   dyn_unpack_reads:    the inverse on the receiving side
 * ------------------------------------------------------------------------------------------------ */
#define DYN_MAX_WORLD 64   /* dyn_owner_partition / dyn_merge_kn take at most 2^31 - 1 rows per call */
#define DYN_READ_ROW_BYTES 40
int dyn_owner_partition(dyn_ctx_t *ctx, const uint32_t *d_barcode, int64_t n, int world, uint32_t *d_perm,
                        uint64_t *d_owner_counts, void *stream);
int dyn_pack_reads(dyn_ctx_t *ctx, const uint32_t *d_perm, int64_t n, const uint8_t *d_counts, const uint32_t *d_barcode,
                   const uint64_t *d_umi, const uint32_t *d_gene, const int16_t *d_score, const uint16_t *d_conv_n,
                   const uint8_t *d_strand, const uint8_t *d_transcriptome, void *d_rows, void *stream);
int dyn_unpack_reads(dyn_ctx_t *ctx, const void *d_rows, int64_t n, uint8_t *d_counts, uint32_t *d_barcode, uint64_t *d_umi,
                     uint32_t *d_gene, int16_t *d_score, uint16_t *d_conv_n, uint8_t *d_strand, uint8_t *d_transcriptome,
                     void *stream);
/* The same exchange fused with the transfer (peer stores over NVLink instead of pack + all-to-all): every rank's receive
 * buffer is mapped into the other ranks of the node (cudaIpc), and the rows are written straight where dedup reads them.
 *   dyn_peer_alloc / dyn_peer_open / dyn_peer_close / dyn_peer_free: a cudaMalloc'ed buffer + its 64-byte IPC handle; the
 *                        handle opened in ANOTHER process of the node gives that process a device pointer to the buffer
 *   dyn_push_plan:       d_tile_base[world * ceil(n / DYN_PUSH_TILE)] u32 (owner-major) = stable position of every
 *                        (owner, tile) run in the order dyn_owner_partition would produce; d_owner_counts[world] as there
 *   dyn_push_reads:      row of read i (layout of dyn_pack_reads) -> peer_rows[owner] + DYN_READ_ROW_BYTES * (dst_base[owner]
 *                        + its stable position), peer_rows / dst_base: HOST arrays [world]; dst_base[o] = rows of lower ranks
 *                        bound for o  -  first position of owner o in my order (may be negative).  The caller orders the
 *                        ranks around the call (distributed.PeerExchange: the count all-gather before, a barrier after).
 * ------------------------------------------------------------------------------------------------ */
#define DYN_PUSH_TILE 256
#define DYN_PUSH_MAX_WORLD 16
#define DYN_PEER_HANDLE_BYTES 64
int dyn_peer_alloc(dyn_ctx_t *ctx, int64_t bytes, void **d_ptr, void *handle_out);
int dyn_peer_open(dyn_ctx_t *ctx, const void *handle, void **d_ptr);
int dyn_peer_close(dyn_ctx_t *ctx, void *d_ptr);
int dyn_peer_free(dyn_ctx_t *ctx, void *d_ptr);
int dyn_push_plan(dyn_ctx_t *ctx, const uint32_t *d_barcode, int64_t n, int world, uint32_t *d_tile_base, uint64_t *d_owner_counts,
                  void *stream);
int dyn_push_reads(dyn_ctx_t *ctx, const uint32_t *d_tile_base, int64_t n, int world, const uint8_t *d_counts,
                   const uint32_t *d_barcode, const uint64_t *d_umi, const uint32_t *d_gene, const int16_t *d_score,
                   const uint16_t *d_conv_n, const uint8_t *d_strand, const uint8_t *d_transcriptome, void *const *peer_rows,
                   const int64_t *dst_base, void *stream);
/* The histogram exchange before the EM when no per-read exchange ran (no UMI stage): every rank holds
 * partial (group, k, n) -> reads rows of its read range (dyn_aggregate_kn keys with gene_bits == 0: group << 22 |
 * k << 10 | n) and sends them to rank group % world.  dyn_merge_kn turns what a rank received -- unsorted, the same
 * key possibly from several ranks -- into dyn_kn_csr's input: group ids divided by world (the owner's local
 * numbering), rows sorted by key, counts of equal keys summed (what csr_matrix does with duplicates, p_c.py:219).
 * group_bits = bits of the largest LOCAL group id (0: all).  Outputs need room for n rows. */
int dyn_merge_kn(dyn_ctx_t *ctx, const uint64_t *d_keys, const uint32_t *d_count, int64_t n, int world, int group_bits,
                 uint64_t *d_out_keys, uint32_t *d_out_count, int64_t *d_n_rows, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Host BAM ingest (what the reference gets from pysam, bam.py:253-312): BGZF inflate on n_threads host
 * threads, read filter (bam.py:173-174, 242-248, 286-292), mate pairing (bam.py:307-312), packing into the
 * record layout above (page-locked memory when available).  The BAM must be coordinate sorted.  tags may be
 * NULL / ""; require_gene = the reference's `not velocity`.  Per-unit metadata is read with dyn_bam_field:
 *   0 blob u8 | 1 rec_off u32[n+1] | 2 ref_id i32 | 3 pos i32 | 4 flag u16 | 5 HI i32 | 6 AS i32 |
 *   7 gx_assigned u8 | 8 paired u8 | 9 ref_len i32 | 10 header text |
 *   string pools (chars, u32 offsets[n+1]): 20/21 read name, 22/23 barcode, 24/25 umi, 26/27 gene, 28/29 ref names
 * ------------------------------------------------------------------------------------------------ */
typedef struct dyn_bam_result dyn_bam_result_t;
int dyn_bam_pack(const char *path, const char *barcode_tag, const char *umi_tag, const char *gene_tag, int require_gene,
                 int n_threads, dyn_bam_result_t **out_result);
/* The same with options (the consensus caller's reader, consensus.py:361-380, filters differently and writes reads
 * through unchanged): DYN_BAM_KEEP_DUPLICATES = duplicate-flagged records are units too (consensus.py:248-249 does not
 * skip them), DYN_BAM_KEEP_RAW = the units' BAM alignment records are kept as they are in the file (fields 30/31: the
 * unit's record, 32/33: its first-seen mate; chars / u32 offsets [n + 1]).  Field 14 (i64) is the ordinal of the unit's
 * record among all records of the file -- enumerate(bam.fetch()) in the reference. */
#define DYN_BAM_KEEP_RAW 0x1u
#define DYN_BAM_KEEP_DUPLICATES 0x2u
int dyn_bam_pack_ex(const char *path, const char *barcode_tag, const char *umi_tag, const char *gene_tag, int require_gene,
                    int n_threads, uint32_t flags, dyn_bam_result_t **out_result);
/* The first max_records records: the aux tags that occur (2 chars each into tags_out, *n_tags of them) and the OR of
 * the flags -- bam.get_tags_from_bam, check_bam_contains_secondary / _unmapped / _duplicate, check_bam_is_paired
 * (bam.py:447-570). */
int dyn_bam_peek(const char *path, int64_t max_records, char *tags_out, int32_t tags_capacity, int32_t *n_tags,
                 uint32_t *flag_or, int64_t *n_seen);
void dyn_bam_result_free(dyn_bam_result_t *result);
int64_t dyn_bam_n_units(const dyn_bam_result_t *result);
int64_t dyn_bam_n_records_seen(const dyn_bam_result_t *result);
int64_t dyn_bam_n_refs(const dyn_bam_result_t *result);
const void *dyn_bam_field(const dyn_bam_result_t *result, int field, int64_t *n_bytes);

/* ------------------------------------------------------------------------------------------------
 * Streaming, range-split ingest (replaces bam.split_bam + one pysam reader per split, bam.py:609-645, 708-726):
 * reader `part` of `n_parts` takes a contiguous range of the file's BGZF blocks (cut by compressed bytes; it owns
 * the records that START there) and returns it chunk by chunk; a producer thread inflates and packs the next chunks
 * while the caller streams the current one to the GPU.  A chunk is a dyn_bam_result_t (dyn_bam_field; free it with
 * dyn_bam_result_free -- its page-locked blob returns to the stream's pool).
 *   lean   single-end units are packed in the DYN_REC_LEAN layout
 *   keys   instead of the string pools (fields 20-27) a chunk carries numeric keys:
 *            11 barcode u64  2-bit packed A/C/G/T under a sentinel bit, a "-<n>" suffix as n + 1 in bits 58..63;
 *                            ~0 = not representable (other letters / longer than 28 nt), 0 = no barcode tag
 *            12 umi u64      likewise (up to 31 nt, no suffix)
 *            13 gene i32     index of the gene tag's value in gene_ids, -1 = no tag, -2 = not in gene_ids
 * Paired BAMs keep the first-seen-mate table across chunks but must be read with n_parts == 1.
 * dyn_bam_stream_next: *chunk == NULL with DYN_OK at the end of the range.
 * dyn_bam_stream_field: 9 ref_len | 10 header text | 28/29 reference names (as dyn_bam_field).
 * ------------------------------------------------------------------------------------------------ */
typedef struct dyn_bam_stream dyn_bam_stream_t;
typedef struct {
    const char *barcode_tag, *umi_tag, *gene_tag;   /* NULL / "" = none */
    int32_t require_gene;                           /* the reference's `not velocity` */
    int32_t n_threads;
    int32_t part, n_parts;
    int64_t chunk_bytes;                            /* inflated bytes per chunk; 0 = 256 MiB */
    int32_t lean, keys;
    const char *const *gene_ids;                    /* [n_gene_ids], used with keys */
    int64_t n_gene_ids;
} dyn_bam_stream_opts_t;
int dyn_bam_stream_open(const char *path, const dyn_bam_stream_opts_t *opts, dyn_bam_stream_t **out_stream);
int dyn_bam_stream_next(dyn_bam_stream_t *stream, dyn_bam_result_t **chunk);
void dyn_bam_stream_close(dyn_bam_stream_t *stream);
int64_t dyn_bam_stream_n_refs(const dyn_bam_stream_t *stream);
const void *dyn_bam_stream_field(const dyn_bam_stream_t *stream, int field, int64_t *n_bytes);

/* ------------------------------------------------------------------------------------------------
 * Device ingest: the same block range, inflated and packed ON THE GPU (csrc/bamdevice.cu) -- BGZF blocks are inflated in
 * two passes (Huffman codes to tokens with one lane per block, tokens to bytes with one warp per block), one thread per
 * alignment record filters / tokenises / packs it; only the compressed bytes cross PCIe.
 * Takes coordinate-sorted single-end BAMs; BGZF blocks need not end on record boundaries (htslib's do, STAR's sorted BAMs
 * do not): the record chain is established per block on the device and checked by induction, a record that starts in one
 * chunk and ends in the next is carried over, the last record of a range is completed from the blocks behind the range.
 * Paired records are refused with DYN_ERR_FORMAT (mate pairing is dyn_bam_stream_*'s).  Options as for
 * dyn_bam_stream_open (keys are always produced, n_threads = host threads that read the file into page-locked memory, 0 = 16;
 * gene_ids gives the gene index).
 * dyn_bam_dev_next fills *chunk with DEVICE pointers to one chunk's units -- packed records, offsets (n_units + 1),
 * keys as in dyn_bam_stream chunks (fields 11-13), AS, reference id, position, HI, flag, gene-tag-present -- valid until
 * the second-next call; chunk->end is 1 on the last chunk of the range and 2 (nothing else set) on calls past it.  The call runs on `stream` and
 * synchronises it (chunk sizes come back from the device).
 * dyn_inflate_blocks: the inflate kernels alone on raw deflate payloads in device memory (the descriptors are read back
 * for the output size, so the call synchronises the stream); d_block_desc = n_blocks x
 * {in_off, in_len, out_off, out_len} (u32), *d_status gets bit (1 << code) of inflate error codes 1 data, 2 input,
 * 3 output size; d_work is a scratch u32.
 * ------------------------------------------------------------------------------------------------ */
typedef struct dyn_bam_dev dyn_bam_dev_t;
typedef struct {
    int64_t n_units, n_records_seen, records_bytes, compressed_bytes, inflated_bytes;
    int32_t end, reserved;
    const void *d_records;
    const uint32_t *d_rec_off;
    const uint64_t *d_barcode_key, *d_umi_key;
    const int32_t *d_gene_idx, *d_score, *d_ref_id, *d_pos, *d_hi;
    const uint16_t *d_flag;
    const uint8_t *d_gx_assigned;
} dyn_bam_dev_chunk_t;
int dyn_bam_dev_open(dyn_ctx_t *ctx, const char *path, const dyn_bam_stream_opts_t *opts, dyn_bam_dev_t **out_dev);
int dyn_bam_dev_next(dyn_bam_dev_t *dev, dyn_bam_dev_chunk_t *chunk, void *stream);
void dyn_bam_dev_close(dyn_bam_dev_t *dev);
int64_t dyn_bam_dev_n_refs(const dyn_bam_dev_t *dev);
const void *dyn_bam_dev_field(const dyn_bam_dev_t *dev, int field, int64_t *n_bytes);
int dyn_inflate_blocks(dyn_ctx_t *ctx, const void *d_in, void *d_out, const uint32_t *d_block_desc, int32_t n_blocks,
                       uint32_t *d_work, uint32_t *d_status, void *stream);

/* ------------------------------------------------------------------------------------------------
 * BAM writer (what call_consensus gets from pysam + samtools sort / index, consensus.py:360-366, 476-494; also builds
 * the synthetic BAMs of the ingest benchmarks).  Unit u -- one packed record in the FULL layout, no mate -- becomes one
 * alignment record: name = h_names[h_name_off[u] .. h_name_off[u + 1]) ("r<u>" when h_names is NULL), MAPQ h_mapq[u]
 * (255 when NULL), flag h_flag[u] (the record's own low 12 bits when NULL), the caller's pre-encoded aux bytes
 * h_aux[h_aux_off[u] .. h_aux_off[u + 1]) followed by the MD tag rebuilt from the record's MD tokens.  Units are written
 * in h_order (NULL: as they come) -- pass the coordinate order for a sorted file.  BGZF blocks end on record boundaries;
 * level = zlib level; write_index != 0 also writes <path>.bai (binning + linear index, SAM specification section 5).
 * Host call, all pointers are host pointers.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
    const char *header_text;
    int32_t n_ref;
    const char *const *ref_names;
    const int32_t *ref_lens;
    const void *h_records;          /* packed units (may be NULL when every unit is raw) */
    const uint32_t *h_rec_off;
    int64_t n_units;
    const int64_t *h_order;
    const char *h_names;
    const uint32_t *h_name_off;
    const uint8_t *h_aux;
    const uint32_t *h_aux_off;
    const uint8_t *h_mapq;
    const uint16_t *h_flag;
    int32_t level, n_threads, write_index, reserved;
    /* dyn_bam_write_ex only: unit u is written from the BAM alignment record h_raw[h_raw_start[u] .. + h_raw_len[u])
     * (fixed fields .. aux, without its block_size) when h_raw_len[u] is not zero -- reads of the input file passed
     * through by call_consensus (consensus.py:420-422) --, and with h_retag[h_retag_off[u] ..) = "gene id\0gene name\0"
     * not empty its gene tags are swapped first (swap_gene_tags, consensus.py:277-286, with `gene_tag`). */
    const uint8_t *h_raw;
    const uint64_t *h_raw_start;
    const uint32_t *h_raw_len;
    const char *h_retag;
    const uint32_t *h_retag_off;
    const char *gene_tag;
    /* MD text of unit u, h_md[h_md_off[u] .. h_md_off[u + 1]), written verbatim instead of being rebuilt from the tokens
     * when not empty (dyn_consensus returns the reference's exact MD string, consensus.py:124-161, 175) */
    const char *h_md;
    const uint32_t *h_md_off;
} dyn_bam_write_opts_t;
int dyn_bam_write_ex(const char *path, const dyn_bam_write_opts_t *opts);
int dyn_bam_write(const char *path, const char *header_text, int32_t n_ref, const char *const *ref_names,
                  const int32_t *ref_lens, const void *h_records, const uint32_t *h_rec_off, int64_t n_units,
                  const int64_t *h_order, const char *h_names, const uint32_t *h_name_off, const uint8_t *h_aux,
                  const uint32_t *h_aux_off, const uint8_t *h_mapq, const uint16_t *h_flag, int level, int n_threads,
                  int write_index);

/* ------------------------------------------------------------------------------------------------
 * Synthetic input generator (bench / test utility; not on the timed path). Generates packed records
 * for the BASELINE configs directly in HBM; every decision is a pure function of (seed, read index).
 * Model: SURVEY.md section 8d (C2) after dynast/benchmarking/simulation.py:22-79.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
    uint64_t seed;
    int32_t read_len;       /* 30..304 */
    int32_t n_barcodes;
    int32_t n_genes;
    int32_t umi_len;        /* nt, 2 bits each in the u64 UMI key */
    int64_t n_groups;       /* number of (barcode, gene, UMI) groups reads are drawn from */
    int64_t pos_span;       /* > 0: reference span over which reads are laid out in index order (0: 2^27);
                             * < 0: reads of a UMI group start within 300 nt of an anchor in [0, -pos_span) */
    double p_c, p_e, frac_n;
    const float *d_bc_cdf;        /* [n_barcodes] cumulative sampling weights */
    const float *d_gene_cdf;      /* [n_genes] */
    const float *d_gene_pi;       /* [n_genes] labelled fraction */
    const uint8_t *d_gene_strand; /* [n_genes] 0 '+', 1 '-' */
    const int32_t *d_gene_contig; /* [n_genes] */
    int32_t lean;                 /* != 0: records in the DYN_REC_LEAN layout (same reads, same qualities at mismatches) */
} dyn_synth_params_t;

/* d_rec_off[0..n_reads] (16-byte units); the blob needs 16 * d_rec_off[n_reads] bytes. */
int dyn_synth_offsets(dyn_ctx_t *ctx, const dyn_synth_params_t *params, int64_t n_reads, uint32_t *d_rec_off,
                      void *stream);
/* Writes the records and per-read keys (any key pointer may be NULL). */
int dyn_synth_fill(dyn_ctx_t *ctx, const dyn_synth_params_t *params, int64_t n_reads, const uint32_t *d_rec_off,
                   void *d_records, uint32_t *d_barcode, uint64_t *d_umi, uint32_t *d_gene, int32_t *d_score,
                   uint8_t *d_strand, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DYNAST_B200_H */
