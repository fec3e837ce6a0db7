"""Shared host helper: aggregate rows -> CSR triplets per group (groupby(sort=False, observed=True) order)."""
from typing import List, Optional

import numpy as np
import pandas as pd


def group_codes(df: pd.DataFrame, group_by: Optional[List[str]]):
    """Returns (codes int64 [rows], keys list) with keys in first-appearance order; a key is a str for one
    column and a tuple for several, like the keys of DataFrameGroupBy.indices."""
    if group_by is None:
        return np.zeros(len(df), dtype=np.int64), [None]
    cols = list(group_by)
    if len(cols) == 1:
        codes, uniques = pd.factorize(df[cols[0]].astype(object), sort=False)
        return codes.astype(np.int64), list(uniques)
    codes, uniques = pd.factorize(pd.MultiIndex.from_frame(df[cols].astype(object)), sort=False)
    return codes.astype(np.int64), [tuple(u) for u in uniques]


def csr_from_rows(codes: np.ndarray, n_groups: int, k, n, count):
    """Stable group-major CSR of (k, n, count) rows."""
    order = np.argsort(codes, kind='stable')
    off = np.zeros(n_groups + 1, dtype=np.int64)
    np.cumsum(np.bincount(codes, minlength=n_groups), out=off[1:])
    return (np.ascontiguousarray(np.asarray(k)[order], dtype=np.int32), np.ascontiguousarray(np.asarray(n)[order], dtype=np.int32),
            np.ascontiguousarray(np.asarray(count)[order], dtype=np.int64), off, order)


def lookup(mapping, keys, n_cols):
    """p_e / p_c / pi_c given as a float or as a {key: value} dict -> float64 array aligned with keys (NaN if absent)."""
    if not isinstance(mapping, dict):
        return np.full(len(keys), float(mapping), dtype=np.float64)
    return np.array([mapping.get(k, np.nan) for k in keys], dtype=np.float64)
