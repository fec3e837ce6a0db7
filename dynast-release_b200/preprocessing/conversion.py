"""count_conversions and friends -- same names, arguments and on-disk outputs as
dynast/preprocessing/conversion.py (reference), computed by the CUDA kernels behind include/dynast_b200.h.

The host side only parses / formats CSV text and interns strings; every per-read computation (quality and
SNP filtered counters, strand complement, dedup priority sorts, output ordering) runs on the device.
"""
import gzip
import pickle
from typing import Dict, FrozenSet, List, Optional, Set

import numpy as np
import pandas as pd
import torch

from .. import pipeline, records
from .. import _lib

CONVERSION_COLUMNS = records.CONVERSION_COLUMNS
BASE_COLUMNS = records.BASE_COLUMNS
COLUMNS = records.COLUMNS
CONVERSION_IDX = {c: i for i, c in enumerate(CONVERSION_COLUMNS)}
BASE_IDX = {b: 12 + i for i, b in enumerate(BASE_COLUMNS)}
_COMP = {'A': 'T', 'C': 'G', 'G': 'C', 'T': 'A'}
CONVERSION_COMPLEMENT = {c: _COMP[c[0]] + _COMP[c[1]] for c in CONVERSION_COLUMNS}
CSV_COLUMNS = ['read_id', 'barcode', 'umi', 'GX'] + COLUMNS + ['velocity', 'transcriptome', 'score']
COUNTS_SPLIT_THRESHOLD = 50000   # dynast/config.py:51
_PANDAS_NA_STRINGS = ['-1.#IND', '1.#QNAN', '1.#IND', '-1.#QNAN', '#N/A N/A', '#N/A', 'N/A', 'n/a', 'NA', '<NA>', '#NA', 'NULL',
                      'null', 'NaN', '-NaN', 'nan', '-nan', 'None', '']


def _device():
    return _lib.current_torch_device()


def read_counts(counts_path: str, *args, **kwargs) -> pd.DataFrame:
    """conversion.read_counts (conversion.py:74-96)."""
    dtypes = {
        'read_id': 'string', 'barcode': 'category', 'umi': 'category', 'GX': 'category', 'velocity': 'category',
        'transcriptome': bool, 'score': np.uint16, **{column: np.uint8 for column in COLUMNS}
    }
    return pd.read_csv(counts_path, dtype=dtypes, na_filter=False, *args, **kwargs)


def subset_counts(df_counts: pd.DataFrame, key: str) -> pd.DataFrame:
    """conversion.subset_counts (conversion.py:128-145)."""
    if key == 'transcriptome':
        df_counts = df_counts[df_counts['transcriptome']]
    if key in ('spliced', 'unspliced'):
        df_counts = df_counts[df_counts['velocity'] == key]
    return df_counts


def split_counts_by_velocity(df_counts: pd.DataFrame) -> Dict[str, pd.DataFrame]:
    """conversion.split_counts_by_velocity (conversion.py:284-300)."""
    return {v: part.reset_index(drop=True) for v, part in df_counts.groupby('velocity', sort=False, observed=True)}


def _strand_array(gx: pd.Series, gene_infos: dict) -> np.ndarray:
    codes, uniques = pd.factorize(gx.astype(object), sort=False)
    rev = np.array([gene_infos[g]['strand'] != '+' for g in uniques], dtype=np.uint8)
    return rev[codes] if len(uniques) else np.zeros(0, dtype=np.uint8)


def _counts_to_device(df: pd.DataFrame) -> torch.Tensor:
    arr = np.ascontiguousarray(df[COLUMNS].to_numpy(dtype=np.int64))
    if arr.size and (arr.max() > 255 or arr.min() < 0):
        raise ValueError('counter outside uint8 range (the reference reads these columns as uint8)')
    return torch.from_numpy(arr.astype(np.uint8)).to(_device())


def complement_counts(df_counts: pd.DataFrame, gene_infos: dict) -> pd.DataFrame:
    """conversion.complement_counts (conversion.py:99-125): counters of reads on '-' genes are complemented on
    the device; rows come back forward-strand first, then reverse (original index kept)."""
    other = [c for c in df_counts.columns if c not in COLUMNS]
    strand = _strand_array(df_counts['GX'], gene_infos)
    ctx = _lib.current_context()
    counts = _counts_to_device(df_counts)
    pipeline.complement_counts(ctx, counts, torch.from_numpy(strand).to(counts.device))
    out = df_counts[other].copy()
    host = counts.cpu().numpy()
    for j, c in enumerate(COLUMNS):
        out[c] = host[:, j].astype(df_counts[c].dtype, copy=False)
    out = out[other + COLUMNS]
    fwd = strand == 0
    return pd.concat((out[fwd], out[~fwd]), verify_integrity=True)


def _intern(values) -> (np.ndarray, int):
    codes, uniques = pd.factorize(pd.Series(values).astype(object), sort=False)
    return codes.astype(np.uint32), len(uniques)


def _lex_rank(values) -> (np.ndarray, int):
    """rank of each value among the sorted distinct values (order of a pandas categorical read from CSV)."""
    codes, uniques = pd.factorize(pd.Series(values).astype(object), sort=True)
    return codes.astype(np.uint64), len(uniques)


def _group_key(parts):
    """Pack interned ids (most significant first) into (hi, lo, hi_bits, lo_bits) uint64 arrays."""
    lo = np.zeros(len(parts[0][0]), dtype=np.uint64)
    hi = None
    lo_bits = 0
    hi_bits = 0
    for ids, n in reversed(parts):
        b = max(1, int(max(n - 1, 0)).bit_length())
        if hi is None and lo_bits + b <= 64:
            lo |= ids.astype(np.uint64) << np.uint64(lo_bits)
            lo_bits += b
        else:
            if hi is None:
                hi = np.zeros(len(lo), dtype=np.uint64)
            if hi_bits + b > 64:
                raise ValueError('group key wider than 128 bits')
            hi |= ids.astype(np.uint64) << np.uint64(hi_bits)
            hi_bits += b
    return hi, lo, hi_bits, lo_bits


def _dedup_frame(df: pd.DataFrame, strand: np.ndarray, conv_n: np.ndarray, conversions, use_conversions: bool, umi: bool,
                 counts_complemented: bool, split_threshold: int, barcode_ids=None, plan=None):
    """Runs dyn_dedup_umi for a counts frame. Returns surviving row positions in output order.
    barcode_ids = (ids, n) and plan = pipeline.plan_splits(...) when the splits were cut from a larger frame."""
    ctx = _lib.current_context()
    dev = _device()
    n = len(df)
    counts = _counts_to_device(df)
    strand_d = torch.from_numpy(np.ascontiguousarray(strand, dtype=np.uint8)).to(dev)
    if counts_complemented:   # the kernel wants original orientation and complements on the fly
        pipeline.complement_counts(ctx, counts, strand_d)
    bc, n_bc = barcode_ids if barcode_ids is not None else _intern(df['barcode'])
    score = df['score'].to_numpy(dtype=np.int64)
    if n and (score.min() < 0 or score.max() > 65535):
        raise ValueError('alignment score outside uint16 range (the reference reads it as uint16)')
    extra = None
    extra_bits = 0
    if umi:
        um, n_um = _intern(df['umi'])
        gx, n_gx = _intern(df['GX'])
        hi, lo, hi_bits, lo_bits = _group_key([(bc, n_bc), (um, n_um), (gx, n_gx)])
    else:
        rid, n_rid = _intern(df['read_id'])
        hi, lo, hi_bits, lo_bits = _group_key([(rid, n_rid)])
        rb, nb = _lex_rank(df['barcode'])
        rg, ng = _lex_rank(df['GX'])
        extra = rb * np.uint64(max(ng, 1)) + rg
        extra_bits = max(1, int(max(nb * max(ng, 1) - 1, 0)).bit_length())
        use_conversions = False
    to = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a).view(dt) if a.dtype.itemsize == np.dtype(dt).itemsize
                                        else np.ascontiguousarray(a).astype(dt)).to(dev)
    out = pipeline.dedup(
        ctx, to(lo, np.int64), counts, strand_d, to(df['transcriptome'].to_numpy(dtype=bool).astype(np.uint8), np.uint8),
        to(score.astype(np.uint16), np.int16), to(conv_n.astype(np.uint16), np.int16), to(bc, np.int32), n_bc,
        conversions=conversions, use_conversions=use_conversions, key_hi=None if hi is None else to(hi, np.int64),
        key_lo_bits=lo_bits, key_hi_bits=hi_bits, final_extra=None if extra is None else to(extra, np.int64),
        final_extra_bits=extra_bits, split_threshold=split_threshold, plan=plan)
    return out.cpu().numpy().astype(np.int64)


def deduplicate_counts(df_counts: pd.DataFrame, conversions: Optional[FrozenSet[str]] = None,
                       use_conversions: bool = True) -> pd.DataFrame:
    """conversion.deduplicate_counts (conversion.py:203-243) on an already complemented frame: one split,
    input row order is the tie-break, survivors come back in ascending priority order."""
    n = len(df_counts)
    # one split, frame order = row order: every read sits in the same (strand, no-mismatch) block
    keep = _dedup_frame_plain(df_counts, conversions, use_conversions)
    return df_counts.iloc[keep].reset_index(drop=True)


def _dedup_rows(df, key_cols, conversions=None, use_conversions=False, with_counts=False):
    """One-split dedup of a frame on the device: priority ([has_conversions,] transcriptome, score,
    conversion_sum), ties -> last row; returns surviving row positions in ascending priority order."""
    ctx = _lib.current_context()
    dev = _device()
    n = len(df)
    counts = _counts_to_device(df) if with_counts else torch.zeros((n, 16), dtype=torch.uint8, device=dev)
    hi, lo, hi_bits, lo_bits = _group_key([_intern(df[c]) for c in key_cols])
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a).view(dt)).to(dev)
    out = pipeline.dedup(
        ctx, t(lo, np.int64), counts, torch.zeros(n, dtype=torch.uint8, device=dev),
        t(df['transcriptome'].to_numpy(dtype=bool).astype(np.uint8), np.uint8),
        t(df['score'].to_numpy(dtype=np.int64).astype(np.uint16), np.int16), torch.ones(n, dtype=torch.int16, device=dev),
        None, 0, conversions=conversions, use_conversions=use_conversions, key_hi=None if hi is None else t(hi, np.int64),
        key_lo_bits=lo_bits, key_hi_bits=hi_bits)
    return out.cpu().numpy().astype(np.int64)


def _dedup_frame_plain(df, conversions, use_conversions):
    return _dedup_rows(df, ['barcode', 'umi', 'GX'], conversions, use_conversions, with_counts=True)


def drop_multimappers(df_counts: pd.DataFrame, conversions: Optional[FrozenSet[str]] = None) -> pd.DataFrame:
    """conversion.drop_multimappers (conversion.py:148-200) on an already complemented frame."""
    df = _multimapper_rules(df_counts)
    strand = np.zeros(len(df), dtype=np.uint8)
    keep = _dedup_frame(df, strand, np.ones(len(df), dtype=np.uint16), conversions, False, umi=False,
                        counts_complemented=False, split_threshold=1 << 62)
    return df.iloc[keep].reset_index(drop=True)


def _multimapper_rules(df: pd.DataFrame) -> pd.DataFrame:
    """Rules 2 and 3 of drop_multimappers concern read ids that occur more than once.  BAM parsing keeps only
    primary alignments (bam.py:173), so this is normally empty; the few affected rows are fixed up here."""
    dup = df['read_id'].duplicated(keep=False)
    if not dup.any():
        return df
    sub = df[dup]
    g = sub.groupby('read_id', sort=False, observed=True)
    none_tx = ~g['transcriptome'].any()
    multigene = g['GX'].nunique() > 1
    multivel = g['velocity'].nunique() > 1
    drop_ids = set(none_tx.index[none_tx & multigene])
    amb_ids = set(none_tx.index[none_tx & multivel])
    df = df.copy()
    if amb_ids:
        if hasattr(df['velocity'], 'cat') and 'ambiguous' not in df['velocity'].cat.categories:
            df['velocity'] = df['velocity'].cat.add_categories('ambiguous')
        df.loc[df['read_id'].isin(amb_ids), 'velocity'] = 'ambiguous'
    if drop_ids:
        df = df[~df['read_id'].isin(drop_ids)]
    return df


def _line_starts(path: str) -> np.ndarray:
    with open(path, 'rb') as f:
        data = np.frombuffer(f.read(), dtype=np.uint8)
    nl = np.flatnonzero(data == 10)
    return np.concatenate(([0], nl + 1))


def count_conversions(
    conversions_path: str,
    alignments_path: str,
    index_path: str,
    counts_path: str,
    gene_infos: dict,
    barcodes: Optional[List[str]] = None,
    snps: Optional[Dict[str, Dict[str, Set[int]]]] = None,
    quality: int = 27,
    conversions: Optional[FrozenSet[str]] = None,
    dedup_use_conversions: bool = True,
    n_threads: int = 8,
    temp_dir: Optional[str] = None
) -> str:
    """conversion.count_conversions (conversion.py:441-608): same inputs, same counts_{convs}.csv bytes.
    n_threads / temp_dir are accepted for signature compatibility (no process pool, no temp files)."""
    with gzip.open(index_path, 'rb') as f:
        index = pickle.load(f)
    ali = pd.read_csv(alignments_path, dtype={'read_id': 'string', 'barcode': object, 'umi': object, 'GX': object,
                                              'velocity': object, 'transcriptome': bool}, na_filter=False)
    n_rows = len(ali)
    ali_starts = _line_starts(alignments_path)[1:n_rows + 1]
    idx = np.asarray(index, dtype=np.int64).reshape(-1, 3)
    row_of = np.searchsorted(ali_starts, idx[:, 2])
    if len(idx) and not np.array_equal(ali_starts[row_of], idx[:, 2]):
        raise ValueError('conversions index does not point at alignment rows')
    has_conv = np.zeros(n_rows, dtype=bool)
    has_conv[row_of] = True
    order = np.concatenate((row_of, np.flatnonzero(~has_conv)))      # conversion.py:533-537
    conv_n = np.concatenate((idx[:, 1], np.zeros(n_rows - len(idx), dtype=np.int64)))
    frame = ali.iloc[order].reset_index(drop=True)

    # mismatch records of the indexed reads
    dev = _device()
    ctx = _lib.current_context()
    conv = pd.read_csv(conversions_path, dtype={'read_id': 'string', 'contig': object, 'conversion': object}, na_filter=False)
    conv_starts = _line_starts(conversions_path)[1:len(conv) + 1]
    first_line = np.searchsorted(conv_starts, idx[:, 0])
    if len(idx) and not np.array_equal(conv_starts[first_line], idx[:, 0]):
        raise ValueError('conversions index does not point at conversion rows')
    line_read = np.repeat(np.arange(len(idx), dtype=np.int64), idx[:, 1])
    line_no = (np.repeat(first_line, idx[:, 1]) + (np.arange(len(line_read)) - np.repeat(np.cumsum(idx[:, 1]) - idx[:, 1], idx[:, 1])))
    sel = conv.iloc[line_no]
    code = {c: i for i, c in enumerate(records.SEQ_CODES)}
    cv = sel['conversion'].to_numpy()
    gn = np.array([code.get(c[0], 0) for c in cv], dtype=np.uint64)
    rn = np.array([code.get(c[1], 0) for c in cv], dtype=np.uint64)
    rec = (sel['genome_i'].to_numpy(dtype=np.uint64) | ((gn << np.uint64(4) | rn) << np.uint64(32)) |
           (sel['quality'].to_numpy(dtype=np.uint64) << np.uint64(40)))
    contig_ids, contig_names = pd.factorize(sel['contig'].astype(object))
    snp_table = records.snp_keys(snps, {c: i for i, c in enumerate(contig_names)})

    if barcodes:
        keep_bc = frame['barcode'].isin(set(barcodes)).to_numpy()
    else:
        keep_bc = np.ones(len(frame), dtype=bool)

    counts = np.zeros((len(frame), 16), dtype=np.uint8)
    base = frame[BASE_COLUMNS].to_numpy(dtype=np.int64)
    if base.size and base.max() > 255:
        raise ValueError('base content above 255 (the reference reads these columns as uint8)')
    counts[:, 12:] = base
    counts_d = torch.from_numpy(counts).to(dev)
    if len(rec):
        st = pipeline.count_records(
            ctx, torch.from_numpy(rec.view(np.int64)).to(dev), torch.from_numpy(line_read.astype(np.int32)).to(dev),
            torch.from_numpy(contig_ids.astype(np.int32)).to(dev), counts_d, quality=quality,
            snps=torch.from_numpy(snp_table.view(np.int64)).to(dev) if len(snp_table) else None)
        status = int(st.item())
        if status & 0x4:
            raise KeyError('conversion with a base outside A/C/G/T passed the quality filter')   # CONVERSION_IDX[...] in the reference
        if status & 0x1:
            raise ValueError('conversion counter above 255 (the reference reads these columns as uint8)')
    frame = frame[keep_bc].reset_index(drop=True)
    conv_n = conv_n[keep_bc]
    counts_d = counts_d[torch.from_numpy(np.flatnonzero(keep_bc)).to(dev)] if not keep_bc.all() else counts_d
    frame[COLUMNS] = counts_d.cpu().numpy()

    umi = bool((frame['umi'] != 'NA').all())
    strand = _strand_array(frame['GX'], gene_infos)
    # splits are cut from the whole frame (conversion.py:543-579), before any per-split rule drops rows
    bc, n_bc = _intern(frame['barcode'])
    tdev = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a).astype(dt)).to(dev)
    plan = pipeline.plan_splits(ctx, tdev(bc, np.int32), tdev(strand, np.uint8), tdev(conv_n.astype(np.uint16), np.int16),
                                n_bc, COUNTS_SPLIT_THRESHOLD) if len(frame) else None
    if not umi:   # conversion.py:542 -> drop_multimappers path
        frame['_row'] = np.arange(len(frame))
        frame = _multimapper_rules(frame)
        rows = frame['_row'].to_numpy()
        conv_n, strand, bc = conv_n[rows], strand[rows], bc[rows]
        frame = frame.drop(columns='_row').reset_index(drop=True)
    keep = _dedup_frame(frame, strand, conv_n, conversions, dedup_use_conversions, umi, counts_complemented=False,
                        split_threshold=COUNTS_SPLIT_THRESHOLD, barcode_ids=(bc, n_bc), plan=plan)
    out = frame.iloc[keep][CSV_COLUMNS[1:]].copy()
    # The reference re-reads every split with pandas' default NA parsing before the final write
    # (conversion.py:604-605), so strings such as the 'NA' umi placeholder come out as empty cells.
    for col in ('barcode', 'umi', 'GX', 'velocity'):
        out[col] = out[col].where(~out[col].isin(_PANDAS_NA_STRINGS), '')
    out.to_csv(counts_path, header=True, index=False)
    return counts_path
