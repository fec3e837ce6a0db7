"""calculate_mutation_rates / aggregate_counts -- same names, arguments and CSV outputs as
dynast/preprocessing/aggregation.py (reference); the group-bys run on the device (csrc/aggregate.cu)."""
from typing import FrozenSet, List, Optional

import numpy as np
import pandas as pd
import torch

from .. import pipeline
from .. import _lib
from .conversion import BASE_COLUMNS, COLUMNS, CONVERSION_COLUMNS, _counts_to_device, _device


def read_rates(rates_path: str) -> pd.DataFrame:
    """aggregation.read_rates (aggregation.py:10-25)."""
    df = pd.read_csv(rates_path, index_col=None)
    if df.shape[0] == 1:
        return df.iloc[0]
    return df


def read_aggregates(aggregates_path: str) -> pd.DataFrame:
    """aggregation.read_aggregates (aggregation.py:28-42)."""
    return pd.read_csv(aggregates_path, dtype={'barcode': 'category', 'GX': 'category', 'conversion': np.uint8,
                                               'base': np.uint8, 'count': np.uint16}, na_filter=False)


def merge_aggregates(*dfs: pd.DataFrame) -> pd.DataFrame:
    """aggregation.merge_aggregates (aggregation.py:45-58)."""
    df = pd.concat(dfs).groupby(['barcode', 'GX', 'conversion', 'base'], sort=False, observed=True).sum().reset_index()
    df['conversion'] = df['conversion'].astype(np.uint8)
    df['base'] = df['base'].astype(np.uint8)
    df['count'] = df['count'].astype(np.uint16)
    return df


def _group_ids(df: pd.DataFrame, group_by):
    """first-appearance group ids (groupby(sort=False) order) -> (ids uint32, key frame of the groups)."""
    if group_by is None:
        return np.zeros(len(df), dtype=np.uint32), None
    cols = [group_by] if isinstance(group_by, str) else list(group_by)
    if len(cols) == 1:
        codes, uniques = pd.factorize(df[cols[0]].astype(object), sort=False)
        return codes.astype(np.uint32), pd.DataFrame({cols[0]: uniques})
    tup = pd.MultiIndex.from_frame(df[cols].astype(object))
    codes, uniques = pd.factorize(tup, sort=False)
    return codes.astype(np.uint32), pd.DataFrame(list(uniques), columns=cols)


def group_sums(df_counts: pd.DataFrame, group_by, conversions=None):
    """Device reduction shared by rates / p_e / alpha: returns (keys frame, sums [G,16], n_reads, n_with)."""
    ctx = _lib.current_context()
    ids, keys = _group_ids(df_counts, group_by)
    G = 1 if keys is None else len(keys)
    counts = _counts_to_device(df_counts)
    sums, n_reads, n_with = pipeline.group_sums(ctx, counts, None, torch.from_numpy(ids.view(np.int32)).to(counts.device), G,
                                                conversions=conversions)
    return keys, sums, n_reads, n_with


def calculate_mutation_rates(df_counts: pd.DataFrame, rates_path: str, group_by: Optional[List[str]] = None) -> str:
    """aggregation.calculate_mutation_rates (aggregation.py:61-89); df_counts is the complemented frame."""
    ctx = _lib.current_context()
    keys, sums, _, _ = group_sums(df_counts, group_by)
    rates, _ = pipeline.rates_pe(ctx, sums)
    df_rates = pd.DataFrame(rates.cpu().numpy(), columns=CONVERSION_COLUMNS)
    if keys is not None:
        df_rates = pd.concat((keys.reset_index(drop=True), df_rates), axis=1)
    df_rates.to_csv(rates_path, index=False)
    return rates_path


def aggregate_counts(df_counts: pd.DataFrame, aggregates_path: str, conversions: FrozenSet[str] = frozenset({'TC'})) -> str:
    """aggregation.aggregate_counts (aggregation.py:92-119): rows (barcode, GX, conversion=k, base=n, count) in
    first-appearance order."""
    ctx = _lib.current_context()
    bc, bc_names = pd.factorize(df_counts['barcode'].astype(object), sort=False)
    gx, gx_names = pd.factorize(df_counts['GX'].astype(object), sort=False)
    counts = _counts_to_device(df_counts)
    dev = counts.device
    rows = pipeline.aggregate_kn(ctx, counts, None, torch.from_numpy(bc.astype(np.int32)).to(dev),
                                 torch.from_numpy(gx.astype(np.int32)).to(dev), len(bc_names), len(gx_names), conversions)
    out = pd.DataFrame({
        'barcode': np.asarray(bc_names, dtype=object)[rows['group'].cpu().numpy()] if len(bc_names) else [],
        'GX': np.asarray(gx_names, dtype=object)[rows['gene'].cpu().numpy()] if len(gx_names) else [],
        'conversion': rows['k'].cpu().numpy(),
        'base': rows['n'].cpu().numpy(),
        'count': rows['count'].cpu().numpy(),
    })
    out.to_csv(aggregates_path, index=False)
    return aggregates_path
