"""SNP detection -- same names, arguments and snps.csv as dynast/preprocessing/snp.py (reference)."""
import copy
import ctypes as C
import gzip
import pickle
from typing import Dict, FrozenSet, List, Optional, Set, Tuple

import numpy as np
import pandas as pd
import torch

from .. import _lib, records
from .._lib import check, ptr
from .conversion import CONVERSION_COLUMNS, CONVERSION_COMPLEMENT, _line_starts

SNP_COLUMNS = ['contig', 'genome_i', 'conversion']


def read_snps(snps_path: str) -> Dict[str, Dict[str, Set[int]]]:
    """snp.read_snps (snp.py:17-37)."""
    df = pd.read_csv(snps_path, dtype={'contig': 'category', 'genome_i': np.uint32, 'conversion': 'category'})
    snps = {}
    for (conversion, contig), indices in dict(df.groupby(['conversion', 'contig'], sort=False, observed=True)['genome_i'].agg(set)).items():
        snps.setdefault(conversion, {})[contig] = indices
    return snps


def read_snp_csv(snp_csv: str) -> Dict[str, Dict[str, Set[int]]]:
    """snp.read_snp_csv (snp.py:40-77)."""
    with open(snp_csv, 'r') as f:
        columns = f.readline().strip().split(',')
    if columns == SNP_COLUMNS:
        return read_snps(snp_csv)
    if len(columns) == 2:
        header = False
        try:
            int(columns[-1])
        except ValueError:
            header = True
        df = pd.read_csv(snp_csv, names=['contig', 'genome_i'], skiprows=1 if header else None,
                         dtype={'contig': 'category', 'genome_i': np.uint32})
        _snps = dict(df.groupby('contig', sort=False, observed=True).agg(set)['genome_i'])
        return {conversion: copy.deepcopy(_snps) for conversion in CONVERSION_COMPLEMENT.keys()}
    raise Exception(f'Failed to parse {snp_csv}')


def unique_counts(keys: np.ndarray):
    """dyn_unique_counts on a host uint64 array -> (keys, count, first) sorted by key."""
    ctx = _lib.current_context()
    dev = _lib.current_torch_device()
    n = len(keys)
    k_d = torch.from_numpy(np.ascontiguousarray(keys).view(np.int64)).to(dev)
    ok = torch.zeros(max(n, 1), dtype=torch.int64, device=dev)
    oc = torch.zeros(max(n, 1), dtype=torch.int32, device=dev)
    of = torch.zeros(max(n, 1), dtype=torch.int32, device=dev)
    n_out = torch.zeros(1, dtype=torch.int64, device=dev)
    check(ctx.lib.dyn_unique_counts(ctx.handle, ptr(k_d), n, ptr(ok), ptr(oc), ptr(of), ptr(n_out),
                                    C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    m = int(n_out.item())
    return ok[:m].cpu().numpy().view(np.uint64), oc[:m].cpu().numpy().view(np.uint32), of[:m].cpu().numpy().view(np.uint32)


def extract_conversions(
    conversions_path: str,
    index_path: str,
    alignments: Optional[Set[Tuple[str, int]]] = None,
    conversions: Optional[FrozenSet[str]] = None,
    quality: int = 27,
    n_threads: int = 8
) -> Dict[str, Dict[str, Dict[int, int]]]:
    """snp.extract_conversions (snp.py:80-194): {conversion: {contig: {genome_i: count}}} over the indexed lines
    of selected alignments with Phred > quality; dictionaries in first-appearance order, as the reference's are."""
    with gzip.open(index_path, 'rb') as f:
        index = pickle.load(f)
    conv = pd.read_csv(conversions_path, dtype={'read_id': 'string', 'contig': object, 'conversion': object}, na_filter=False)
    idx = np.asarray(index, dtype=np.int64).reshape(-1, 3)
    starts = _line_starts(conversions_path)[1:len(conv) + 1]
    first_line = np.searchsorted(starts, idx[:, 0])
    line_no = np.repeat(first_line, idx[:, 1]) + (np.arange(int(idx[:, 1].sum())) - np.repeat(np.cumsum(idx[:, 1]) - idx[:, 1], idx[:, 1]))
    sel = conv.iloc[line_no]
    ok = sel['quality'].to_numpy(dtype=np.int64) > quality
    if alignments:
        rid, hi = sel['read_id'].to_numpy(), sel['index'].to_numpy()
        ok &= np.fromiter(((rid[i], int(hi[i])) in alignments for i in range(len(sel))), dtype=bool, count=len(sel))
    cv = sel['conversion'].to_numpy()
    cnames, cidx = np.unique(cv, return_inverse=True) if len(cv) else (np.zeros(0, dtype=object), np.zeros(0, dtype=np.int64))
    if conversions:
        ok &= np.isin(cv, list(conversions))
    contig_ids, contig_names = pd.factorize(sel['contig'].astype(object))
    keys = (cidx.astype(np.uint64) << np.uint64(56)) | (contig_ids.astype(np.uint64) << np.uint64(32)) | sel['genome_i'].to_numpy(dtype=np.uint64)
    keys = np.where(ok, keys, np.uint64(0xffffffffffffffff))
    uk, cnt, first = unique_counts(keys)
    convs: Dict[str, Dict[str, Dict[int, int]]] = {}
    for j in np.argsort(first, kind='stable'):
        k = int(uk[j])
        convs.setdefault(str(cnames[k >> 56]), {}).setdefault(contig_names[(k >> 32) & 0xffffff], {})[k & 0xffffffff] = int(cnt[j])
    return convs


def detect_snps(
    conversions_path: str,
    index_path: str,
    coverage: Dict[str, Dict[int, int]],
    snps_path: str,
    alignments: Optional[Set[Tuple[str, int]]] = None,
    conversions: Optional[FrozenSet[str]] = None,
    quality: int = 27,
    threshold: float = 0.5,
    min_coverage: int = 1,
    n_threads: int = 8,
) -> str:
    """snp.detect_snps (snp.py:197-250): positions with conversions / coverage > threshold."""
    convs = extract_conversions(conversions_path, index_path, alignments=alignments, conversions=conversions, quality=quality,
                                n_threads=n_threads)
    with open(snps_path, 'w') as f:
        f.write(f'{",".join(SNP_COLUMNS)}\n')
        for conversion, _convs in convs.items():
            for contig, by_pos in _convs.items():
                cov = coverage.get(contig, {})
                for genome_i, count in by_pos.items():
                    fraction = count / cov.get(genome_i, 0)   # ZeroDivisionError where the reference raises it
                    if cov[genome_i] >= min_coverage and fraction > threshold:
                        f.write(f'{contig},{genome_i},{conversion}\n')
    return snps_path
