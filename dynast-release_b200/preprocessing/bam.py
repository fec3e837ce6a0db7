"""parse_all_reads and the alignment readers -- same names, arguments and files as dynast/preprocessing/bam.py
(reference).  BAM decode is the C++ ingest (csrc/bampack.cpp), the per-read CIGAR/MD walk is the tally kernel
(csrc/tally.cu, streamed from pinned memory by dyn_tally_reads_host); Python only formats the CSV text."""
import gzip
import pickle
from typing import Dict, List, Optional, Set, Tuple

import numpy as np
import pandas as pd
import torch

from .. import _lib, bamio, device, pipeline, records

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

    Velocity assignment (bam.py:177-228, 363-383) is SURVEY section 8 row f2 ("next"): with velocity=True the
    (gene, assignment) of every read must be supplied through `velocity_assignments` ({(read_id, HI): (GX,
    velocity)}; reads absent from it are dropped, as bam.py:382-383 does for unassigned reads).
    `n_threads` drives the BGZF inflate; `temp_dir`, `strand`, `transcript_infos`, `control`,
    `strict_exon_overlap` are accepted for signature compatibility."""
    if velocity and velocity_assignments is None:
        raise NotImplementedError('velocity assignment is not part of the B200 hot path yet: pass velocity=False '
                                  '(dynast count --no-splicing) or supply velocity_assignments')
    with bamio.PackedBam(bam_path, barcode_tag=barcode_tag, umi_tag=umi_tag, gene_tag=gene_tag, require_gene=not velocity,
                         n_threads=n_threads) as pb:
        n = pb.n_units
        out = device.tally_reads_host(pb.blob, pb.rec_off, quality=0, nasc=nasc, emit_conv=True,
                                      conv_capacity=max(16, int(pb.blob.size)), device=_lib.current_device_index())
        _check_tally_status(out['status'], out['counts'])
        read_id, barcode, umi, gene = pb.read_id, pb.barcode, pb.umi, pb.gene
        hi, score, gx_assigned = pb.hi.copy(), pb.score.copy(), pb.gx_assigned.astype(bool)
        contig = pb.ref_name[pb.ref_id] if n else np.zeros(0, dtype=object)
    keep = np.ones(n, dtype=bool)
    if barcodes:
        keep &= np.isin(barcode, list(barcodes))
    assignment = np.full(n, 'unassigned', dtype=object)
    gx_out = gene.copy()
    if velocity:
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
