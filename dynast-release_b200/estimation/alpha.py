"""Detection rate alpha -- same names, arguments and outputs as dynast/estimation/alpha.py (reference)."""
from typing import Dict, FrozenSet, List, Optional, Tuple, Union

import numpy as np
import pandas as pd
import torch

from .. import pipeline
from .. import _lib
from ..preprocessing.aggregation import group_sums
from ._groups import group_codes, lookup


def read_alpha(alpha_path: str, group_by: Optional[List[str]] = None):
    """alpha.read_alpha (alpha.py:9-23)."""
    if group_by is None:
        with open(alpha_path, 'r') as f:
            return float(f.read())
    df = pd.read_csv(alpha_path, dtype={key: 'category' for key in group_by})
    return dict(df.set_index(group_by)['alpha'])


def estimate_alpha(
    df_counts: pd.DataFrame,
    pi_c: Union[float, Dict[str, float], Dict[Tuple[str, ...], float]],
    alpha_path: str,
    conversions: FrozenSet[str] = frozenset({'TC'}),
    group_by: Optional[List[str]] = None,
    pi_c_group_by: Optional[List[str]] = None,
) -> str:
    """alpha.estimate_alpha (alpha.py:26-95): alpha = (reads with any of `conversions` / reads) / pi_c."""
    df = df_counts
    if pi_c_group_by is not None:
        pcodes, pkeys = group_codes(df, pi_c_group_by)
        pic_rows = lookup(pi_c, pkeys, len(pi_c_group_by))[pcodes] if len(df) else np.zeros(0)
    else:
        pic_rows = np.full(len(df), float(pi_c))
    keep = ~np.isnan(pic_rows)
    df = df[keep]
    pic_rows = pic_rows[keep]
    ctx = _lib.current_context()
    if group_by is None:
        if not isinstance(pi_c, float):
            raise Exception('`pi_c` and `p_c` must be a float when `group_by` is not provided')
        if pi_c <= 0:
            raise Exception(f'Estimated `pi_c` must be positive, but got {pi_c}')
    keys, sums, n_reads, n_with = group_sums(df, group_by, conversions=conversions)
    codes, gkeys = group_codes(df, group_by)
    G = len(gkeys)
    pic = np.full(G, np.nan)
    if len(df):
        lo = np.full(G, np.inf); hi = np.full(G, -np.inf)
        np.minimum.at(lo, codes, pic_rows); np.maximum.at(hi, codes, pic_rows)
        if np.any(lo != hi):
            raise Exception('`pi_c` for each aggregate group must be a constant')
        pic = lo
    alpha = pipeline.alpha(ctx, n_reads, n_with, torch.from_numpy(pic).to(n_reads.device)).cpu().numpy()
    with open(alpha_path, 'w') as f:
        if group_by is None:
            f.write(str(float(alpha[0])))
        else:
            f.write(f'{",".join(group_by)},alpha\n')
            alphas = {gkeys[g]: float(alpha[g]) for g in range(G) if pic[g] > 0}
            for key in sorted(alphas.keys()):
                formatted_key = key if isinstance(key, str) else ','.join(key)
                f.write(f'{formatted_key},{alphas[key]}\n')
    return alpha_path
