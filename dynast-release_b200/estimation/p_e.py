"""Background rate p_e -- same names, arguments and outputs as dynast/estimation/p_e.py (reference)."""
from typing import FrozenSet, List, Optional

import pandas as pd
import torch

from .. import pipeline
from .. import _lib
from ..preprocessing.aggregation import group_sums
from ..preprocessing.conversion import BASE_COLUMNS, CONVERSION_COLUMNS


def _flatten(conversions):
    out = []
    for c in conversions:
        if isinstance(c, str):
            out.append(c)
        else:
            out.extend(_flatten(c))
    return out


def read_p_e(p_e_path: str, group_by: Optional[List[str]] = None):
    """p_e.read_p_e (p_e.py:11-26)."""
    if group_by is None:
        with open(p_e_path, 'r') as f:
            return float(f.read())
    df = pd.read_csv(p_e_path, dtype={key: 'category' for key in group_by})
    return dict(df.set_index(group_by)['p_e'])


def estimate_p_e_control(df_counts: pd.DataFrame, p_e_path: str,
                         conversions: FrozenSet[FrozenSet[str]] = frozenset({frozenset({'TC'})})) -> str:
    """p_e.estimate_p_e_control (p_e.py:29-49): sum of the conversion columns / sum of their base columns."""
    flattened = _flatten(conversions)
    bases = list(set(f[0] for f in flattened))
    _, sums, _, _ = group_sums(df_counts, None)
    s = sums.cpu().numpy()[0]
    num = sum(int(s[CONVERSION_COLUMNS.index(c)]) for c in flattened)
    den = sum(int(s[12 + BASE_COLUMNS.index(b)]) for b in bases)
    p_e = num / den
    with open(p_e_path, 'w') as f:
        f.write(str(p_e))
    return p_e_path


def estimate_p_e(df_counts: pd.DataFrame, p_e_path: str,
                 conversions: FrozenSet[FrozenSet[str]] = frozenset({frozenset({'TC'})}),
                 group_by: Optional[List[str]] = None) -> str:
    """p_e.estimate_p_e (p_e.py:52-108): NaN-skipping mean of the rates of the conversions that do not start
    with a labelled base (all non-induced conversions when every base is labelled)."""
    flattened = _flatten(conversions)
    bases = sorted(set(f[0] for f in flattened))
    columns = [c for c in CONVERSION_COLUMNS if c[0] not in bases]
    if bases == BASE_COLUMNS:
        columns = [c for c in CONVERSION_COLUMNS if c not in flattened]
    ctx = _lib.current_context()
    keys, sums, _, _ = group_sums(df_counts, group_by)
    _, p_e = pipeline.rates_pe(ctx, sums, pe_conversions=columns)
    p_e = p_e.cpu().numpy()
    if group_by is not None:
        out = keys.reset_index(drop=True).copy()
        out['p_e'] = p_e
        out.to_csv(p_e_path, header=list(group_by) + ['p_e'], index=False)
    else:
        with open(p_e_path, 'w') as f:
            f.write(str(float(p_e[0])))
    return p_e_path


def estimate_p_e_nasc(df_rates: pd.DataFrame, p_e_path: str, group_by: Optional[List[str]] = None) -> str:
    """p_e.estimate_p_e_nasc (p_e.py:111-134): (CT + GA) / 2 of already computed rates (two columns of a
    per-barcode table: no kernel involved)."""
    if group_by is not None:
        df_rates = df_rates.set_index(group_by)
    p_e = (df_rates['CT'] + df_rates['GA']) / 2
    if group_by is not None:
        p_e.reset_index().to_csv(p_e_path, header=group_by + ['p_e'], index=False)
    else:
        with open(p_e_path, 'w') as f:
            f.write(str(p_e))
    return p_e_path
