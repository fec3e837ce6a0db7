"""parse_all_reads and the alignment readers -- same names, arguments and files as dynast/preprocessing/bam.py
(reference).  BAM decode is the C++ ingest (csrc/bampack.cpp), the per-read CIGAR/MD walk is the tally kernel
(csrc/tally.cu, streamed from pinned memory by dyn_tally_reads_host); Python only formats the CSV text."""
import gzip
import pickle
from typing import Dict, List, Optional, Set, Tuple

import numpy as np
import pandas as pd
import torch

from .. import _lib, bamio, device, gtf, pipeline, records

CONVERSION_CSV_COLUMNS = ['read_id', 'index', 'contig', 'genome_i', 'conversion', 'quality']
ALIGNMENT_COLUMNS = ['read_id', 'index', 'barcode', 'umi', 'GX', 'A', 'C', 'G', 'T', 'velocity', 'transcriptome', 'score']


def read_alignments(alignments_path: str, *args, **kwargs) -> pd.DataFrame:
    """bam.read_alignments (bam.py:24-55)."""
    return pd.read_csv(
        alignments_path,
        dtype={'read_id': 'string', 'index': np.uint8, 'barcode': 'category', 'umi': 'category', 'GX': 'category',
               'A': np.uint8, 'C': np.uint8, 'G': np.uint8, 'T': np.uint8, 'velocity': 'category', 'transcriptome': bool,
               'score': np.uint16},
        na_filter=False, *args, **kwargs)


def read_conversions(conversions_path: str, *args, **kwargs) -> pd.DataFrame:
    """bam.read_conversions (bam.py:58-83)."""
    return pd.read_csv(
        conversions_path,
        dtype={'read_id': 'string', 'index': np.uint8, 'contig': 'category', 'genome_i': np.uint32, 'conversion': 'category',
               'quality': np.uint8},
        na_filter=False, *args, **kwargs)


def select_alignments(df_alignments: pd.DataFrame) -> Set[Tuple[str, int]]:
    """bam.select_alignments (bam.py:86-118): preliminary dedup by (transcriptome, score); returns the set of
    (read_id, alignment index) kept.  The priority sort + keep-last runs on the device (dyn_dedup_umi)."""
    from .conversion import _dedup_rows
    df = df_alignments
    umi = bool((df['umi'] != 'NA').all())
    if umi:
        keep = _dedup_rows(df, ['barcode', 'umi', 'GX'])
    else:
        dup = df['read_id'].duplicated(keep=False)
        if dup.any():
            g = df[dup].groupby('read_id', sort=False, observed=True)
            bad = (~g['transcriptome'].any()) & (g['GX'].nunique() > 1)
            df = df[~df['read_id'].isin(set(bad.index[bad]))]
        keep = _dedup_rows(df, ['read_id'])
    sel = df.iloc[keep]
    return set(zip(sel['read_id'].astype(str), sel['index'].astype(int)))


BAM_PEEK_READS = 500000          # dynast/config.py:44
BAM_REQUIRED_TAGS = ['MD']
BAM_READGROUP_TAG, BAM_BARCODE_TAG, BAM_UMI_TAG, BAM_GENE_TAG, BAM_CONSENSUS_READ_COUNT_TAG = 'RG', 'CB', 'UB', 'GX', 'RN'


def get_tags_from_bam(bam_path: str, n_reads: int = 100000, n_threads: int = 8) -> Set[str]:
    """bam.get_tags_from_bam (bam.py:447-466): the aux tags of the first `n_reads` records (dyn_bam_peek)."""
    return bamio.peek(bam_path, n_reads)[0]


def check_bam_tags_exist(bam_path: str, tags: List[str], n_reads: int = 100000, n_threads: int = 8) -> Tuple[bool, List[str]]:
    """bam.check_bam_tags_exist (bam.py:469-497)."""
    found = get_tags_from_bam(bam_path, n_reads)
    missing = [t for t in tags if t not in found]
    return not missing, missing


def check_bam_is_paired(bam_path: str, n_reads: int = 100000, n_threads: int = 8) -> bool:
    """bam.check_bam_is_paired (bam.py:500-517)."""
    return bool(bamio.peek(bam_path, n_reads)[1] & 0x1)


def check_bam_contains_secondary(bam_path: str, n_reads: int = 100000, n_threads: int = 8) -> bool:
    """bam.check_bam_contains_secondary (bam.py:520-537)."""
    return bool(bamio.peek(bam_path, n_reads)[1] & 0x100)


def check_bam_contains_unmapped(bam_path: str, n_reads: int = BAM_PEEK_READS) -> bool:
    """bam.check_bam_contains_unmapped (bam.py:540-553).  The reference asks the .bai for its unmapped counts; here the first
    `n_reads` records are inspected, like the other checks."""
    return bool(bamio.peek(bam_path, n_reads)[1] & 0x4)


def check_bam_contains_duplicate(bam_path: str, n_reads: int = 100000, n_threads: int = 8) -> bool:
    """bam.check_bam_contains_duplicate (bam.py:556-570)."""
    return bool(bamio.peek(bam_path, n_reads)[1] & 0x400)


def check_bam_tags(bam_path: str, umi_tag: Optional[str] = None, barcode_tag: Optional[str] = None, gene_tag: Optional[str] = 'GX',
                   n_threads: int = 8) -> Set[str]:
    """The tag / alignment checks at the top of count() and consensus() (count.py:86-135, consensus.py:52-100): warnings for
    tags that are present but not asked for, an Exception for required tags that never occur, warnings for secondary /
    unmapped / duplicate records.  Returns the tags found."""
    import warnings
    tags, flags, _ = bamio.peek(bam_path, BAM_PEEK_READS)
    required = list(BAM_REQUIRED_TAGS)
    hint = lambda opt, tag: warnings.warn(f"BAM contains reads with {tag} tag. Are you sure you didn't mean to provide `{opt} {tag}`?")
    if barcode_tag:
        required.append(barcode_tag)
    elif BAM_BARCODE_TAG in tags:
        hint('--barcode-tag', BAM_BARCODE_TAG)
    elif BAM_READGROUP_TAG in tags:
        hint('--barcode-tag', BAM_READGROUP_TAG)
    if umi_tag:
        required.append(umi_tag)
    elif BAM_UMI_TAG in tags:
        hint('--umi-tag', BAM_UMI_TAG)
    if gene_tag:
        required.append(gene_tag)
    elif BAM_GENE_TAG in tags:
        hint('--gene-tag', BAM_GENE_TAG)
    missing = set(required) - tags
    if missing:
        raise Exception(f'First {BAM_PEEK_READS} reads in the BAM do not contain the following required tags: '
                        f'{", ".join(sorted(missing))}. ')
    if flags & 0x100:
        warnings.warn('BAM contains secondary alignments, which will be ignored. Only primary alignments are considered.')
    if flags & 0x4:
        warnings.warn('BAM contains unmapped reads, which will be ignored.')
    if flags & 0x400:
        warnings.warn('BAM contains duplicate reads, which will be ignored.')
    return tags


def _check_tally_status(status: int, counts: np.ndarray):
    """DYN_ST_* bits of a quality-0 tally run -> the exception the reference would end in.  The kernel clamps the
    A/C/G/T content at 255 and skips IUPAC mismatches; the reference writes the true values to the CSVs and fails
    when they are re-read (uint8 columns, conversion.py:85-95; CONVERSION_IDX KeyError, conversion.py:421), so a
    silently different file is never produced here."""
    if status & 0x2:
        raise ValueError('malformed alignment record (CIGAR / MD / sequence disagree)')
    if status & 0x8:
        raise ValueError('conversion-record capacity exceeded while parsing the BAM')
    # 0x4 (IUPAC mismatch): the record IS emitted (resolve_event returns it before the counter lookup), so
    # conversions.csv carries the row exactly as the reference writes it; count_conversions raises on it later
    if status & 0x1 and len(counts) and (counts[:, 12:16] == 255).any():
        raise ValueError('a read (or read pair) covers more than 255 reference bases of one kind: alignments.csv '
                         'holds uint8 base contents (conversion.py:85-95)')


_TABLES_CACHE = {}


def annotation_tables(gene_infos: dict, transcript_infos: dict, ref_names) -> gtf.AnnotationTables:
    """Device tables of an annotation (built once per (gene_infos, transcript_infos, contig list))."""
    key = (id(gene_infos), id(transcript_infos), tuple(ref_names))
    hit = _TABLES_CACHE.get(key)
    if hit is None or hit[0] is not gene_infos:
        _TABLES_CACHE.clear()
        hit = _TABLES_CACHE[key] = (gene_infos, gtf.AnnotationTables(gene_infos, transcript_infos or {}, list(ref_names)))
    return hit[1]


def assign_velocity(blob: np.ndarray, rec_off: np.ndarray, gene: np.ndarray, tables: gtf.AnnotationTables,
                    strand: str = 'forward', strict_exon_overlap: bool = False, chunk_bytes: int = 1 << 28):
    """bam.assign_velocity_type (bam.py:177-228, 363-383) for every unit of a packed BAM: (gene ids, assignment
    names, keep mask).  Records are streamed to the device in chunks; the work is dyn_velocity."""
    ctx = _lib.current_context()
    dev = _lib.current_torch_device()
    n = len(rec_off) - 1
    uniq, inv = np.unique(np.asarray(gene, dtype=object).astype(str), return_inverse=True) if n else (np.zeros(0, str), np.zeros(0, int))
    rows = np.array([-1 if g == '' else tables.gene_index.get(g, -2) for g in uniq], dtype=np.int32)
    gx = rows[inv] if n else np.zeros(0, dtype=np.int32)
    out_gene = np.full(n, -1, dtype=np.int32)
    out_type = np.zeros(n, dtype=np.uint8)
    status = 0
    off64 = rec_off.astype(np.int64)
    u0 = 0
    while u0 < n:
        u1 = int(np.searchsorted(off64, off64[u0] + max(chunk_bytes // 16, 1), side='right')) - 1
        u1 = min(max(u1, u0 + 1), n)
        d_rec = torch.from_numpy(np.ascontiguousarray(blob[off64[u0] * 16:off64[u1] * 16])).to(dev)
        d_off = torch.from_numpy((off64[u0:u1 + 1] - off64[u0]).astype(np.int32)).to(dev)
        g, t, st = pipeline.velocity(ctx, d_rec, d_off, tables, gx=torch.from_numpy(gx[u0:u1]).to(dev), strand=strand,
                                     strict_exon_overlap=strict_exon_overlap)
        out_gene[u0:u1] = g.cpu().numpy()
        out_type[u0:u1] = t.cpu().numpy()
        status |= int(st.item())
        u0 = u1
    if status & 2:
        bad = sorted({g for g, r in zip(uniq, rows) if r == -2})
        raise KeyError(f'gene tag value not in the annotation of its contig: {bad[:5] if bad else "(contig mismatch)"}')
    if status & 1:
        raise ValueError('an alignment has more than 48 aligned blocks')
    ids = np.asarray(tables.gene_ids + [''], dtype=object)
    return ids[out_gene], np.asarray(pipeline.VELOCITY_NAMES, dtype=object)[out_type], out_type != 0


def _lens(arr) -> np.ndarray:
    return np.char.str_len(np.asarray(arr, dtype=str)).astype(np.int64)


def parse_all_reads(
    bam_path: str,
    conversions_path: str,
    alignments_path: str,
    index_path: str,
    gene_infos: dict,
    transcript_infos: dict,
    strand: str = 'forward',
    umi_tag: Optional[str] = None,
    barcode_tag: Optional[str] = None,
    gene_tag: str = 'GX',
    barcodes: Optional[List[str]] = None,
    n_threads: int = 8,
    temp_dir: Optional[str] = None,
    nasc: bool = False,
    control: bool = False,
    velocity: bool = True,
    strict_exon_overlap: bool = False,
    return_splits: bool = False,
    velocity_assignments: Optional[Dict[Tuple[str, int], Tuple[str, str]]] = None,
):
    """bam.parse_all_reads (bam.py:648-780): conversions.csv, alignments.csv and the conversions index of a
    coordinate-sorted BAM, byte for byte.

    Velocity assignment (bam.py:177-228, 363-383) runs on the device (dyn_velocity) from `gene_infos` /
    `transcript_infos` (gtf.genes_and_transcripts_from_gtf, or the dictionaries ngs_tools builds); reads it
    leaves unassigned are dropped (bam.py:382-383).  `velocity_assignments` ({(read_id, HI): (GX, velocity)})
    overrides the computation when given.  `n_threads` drives the BGZF inflate; `temp_dir` and `control` are
    accepted for signature compatibility."""
    if velocity and velocity_assignments is None and not gene_infos:
        raise ValueError('velocity=True needs gene_infos and transcript_infos (gtf.genes_and_transcripts_from_gtf)')
    with bamio.PackedBam(bam_path, barcode_tag=barcode_tag, umi_tag=umi_tag, gene_tag=gene_tag, require_gene=not velocity,
                         n_threads=n_threads) as pb:
        n = pb.n_units
        out = device.tally_reads_host(pb.blob, pb.rec_off, quality=0, nasc=nasc, emit_conv=True,
                                      conv_capacity=max(16, int(pb.blob.size)), device=_lib.current_device_index())
        _check_tally_status(out['status'], out['counts'])
        read_id, barcode, umi, gene = pb.read_id, pb.barcode, pb.umi, pb.gene
        hi, score, gx_assigned = pb.hi.copy(), pb.score.copy(), pb.gx_assigned.astype(bool)
        contig = pb.ref_name[pb.ref_id] if n else np.zeros(0, dtype=object)
        computed = None
        if velocity and velocity_assignments is None:
            tables = annotation_tables(gene_infos, transcript_infos, pb.ref_name)
            computed = assign_velocity(pb.blob, pb.rec_off, gene, tables, strand=strand,
                                       strict_exon_overlap=strict_exon_overlap)
    keep = np.ones(n, dtype=bool)
    if barcodes:
        keep &= np.isin(barcode, list(barcodes))
    assignment = np.full(n, 'unassigned', dtype=object)
    gx_out = gene.copy()
    if computed is not None:
        gx_out, assignment, assigned = computed
        keep &= assigned
    elif velocity:
        for i in range(n):
            hit = velocity_assignments.get((read_id[i], int(hi[i]))) if keep[i] else None
            if hit is None:
                keep[i] = False
            else:
                gx_out[i], assignment[i] = hit
    counts = out['counts']
    conv_off, conv_n = out['conv_off'].astype(np.int64), out['conv_n'].astype(np.int64)
    rows = np.flatnonzero(keep)

    # ---- alignments.csv (bam.py:425-429)
    ali = pd.DataFrame({
        'read_id': read_id[rows], 'index': hi[rows], 'barcode': barcode[rows], 'umi': umi[rows], 'GX': gx_out[rows],
        'A': counts[rows, 12], 'C': counts[rows, 13], 'G': counts[rows, 14], 'T': counts[rows, 15],
        'velocity': assignment[rows], 'transcriptome': gx_assigned[rows], 'score': score[rows],
    }, columns=ALIGNMENT_COLUMNS)
    ali.to_csv(alignments_path, index=False)
    ali_len = sum(_lens(ali[c].to_numpy()) for c in ALIGNMENT_COLUMNS) + len(ALIGNMENT_COLUMNS)   # commas + newline
    ali_start = len(','.join(ALIGNMENT_COLUMNS)) + 1 + np.concatenate(([0], np.cumsum(ali_len)[:-1])) if len(rows) else np.zeros(0, np.int64)

    # ---- conversions.csv (bam.py:405-419) and the index (bam.py:422-423)
    cn = conv_n[rows]
    take = np.repeat(conv_off[rows], cn) + (np.arange(int(cn.sum())) - np.repeat(np.cumsum(cn) - cn, cn))
    pos, refl, readl, q = records.decode_conv(out['conv_rec'][take])
    owner = np.repeat(rows, cn)
    conv = pd.DataFrame({
        'read_id': read_id[owner], 'index': hi[owner], 'contig': contig[owner], 'genome_i': pos,
        'conversion': np.char.add(refl.astype(str), readl.astype(str)) if len(pos) else np.zeros(0, dtype=str), 'quality': q,
    }, columns=CONVERSION_CSV_COLUMNS)
    conv.to_csv(conversions_path, index=False)
    line_len = sum(_lens(conv[c].to_numpy()) for c in CONVERSION_CSV_COLUMNS) + len(CONVERSION_CSV_COLUMNS)
    line_start = len(','.join(CONVERSION_CSV_COLUMNS)) + 1 + np.concatenate(([0], np.cumsum(line_len)[:-1])) if len(pos) else np.zeros(0, np.int64)
    first_line = np.cumsum(cn) - cn
    with_conv = np.flatnonzero(cn > 0)
    index = [(int(line_start[first_line[i]]), int(cn[i]), int(ali_start[i])) for i in with_conv]
    with gzip.open(index_path, 'wb') as f:
        pickle.dump(index, f)
    if return_splits:
        return conversions_path, alignments_path, index_path, []
    return conversions_path, alignments_path, index_path
