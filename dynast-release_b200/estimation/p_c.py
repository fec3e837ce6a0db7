"""Induced rate p_c -- same names, arguments and outputs as dynast/estimation/p_c.py (reference); the EM runs
in csrc/em.cu (one CTA per barcode, bit-identical arithmetic)."""
from typing import Dict, List, Optional, Tuple, Union

import numpy as np
import pandas as pd
import torch

from .. import _lib

from .. import device
from ._groups import csr_from_rows, group_codes, lookup


def read_p_c(p_c_path: str, group_by: Optional[List[str]] = None):
    """p_c.read_p_c (p_c.py:13-29)."""
    if group_by is None:
        with open(p_c_path, 'r') as f:
            return float(f.read())
    df = pd.read_csv(p_c_path, dtype={key: 'category' for key in group_by})
    return dict(df.set_index(group_by)['p_c'])


def estimate_p_c(
    df_aggregates: pd.DataFrame,
    p_e: Union[float, Dict[str, float], Dict[Tuple[str, ...], float]],
    p_c_path: str,
    group_by: Optional[List[str]] = None,
    threshold: int = 1000,
    n_threads: int = 8,
    nasc: bool = False
) -> str:
    """p_c.estimate_p_c (p_c.py:175-247). Groups below `threshold` reads are skipped; groups whose EM fails
    (p_c <= p_e, arithmetic error inside the reference's binomial_pmf) are left out of the CSV, as the
    reference does by swallowing the exception (p_c.py:222-227).  n_threads is accepted and unused."""
    codes, keys = group_codes(df_aggregates, group_by)
    k, n, c, off, _ = csr_from_rows(codes, len(keys), df_aggregates['conversion'].to_numpy(),
                                    df_aggregates['base'].to_numpy(), df_aggregates['count'].to_numpy(dtype=np.int64))
    totals = np.add.reduceat(c, off[:-1]) if len(c) else np.zeros(len(keys), dtype=np.int64)
    totals = np.where(off[1:] > off[:-1], totals, 0)
    pe = lookup(p_e, keys, 1)
    if group_by is None:
        p_c, status, _ = device.em_pc_host(k, n, c, off, pe, nasc=nasc, device=_lib.current_device_index())
        if status[0] == 1:
            raise ValueError('p_c <= p_e')
        if status[0] != 0:
            raise ZeroDivisionError('division by zero')
        with open(p_c_path, 'w') as f:
            f.write(str(float(p_c[0])))
        return p_c_path
    run = (totals >= threshold) & ~np.isnan(pe)
    if np.isnan(pe[totals >= threshold]).any():
        missing = [keys[i] for i in np.flatnonzero((totals >= threshold) & np.isnan(pe))]
        raise KeyError(missing[0])   # p_e[key] in the reference
    sel = np.flatnonzero(run)
    # gather the selected groups into a compact CSR
    lens = (off[1:] - off[:-1])[sel]
    sub_off = np.zeros(len(sel) + 1, dtype=np.int64)
    np.cumsum(lens, out=sub_off[1:])
    take = np.concatenate([np.arange(off[g], off[g + 1]) for g in sel]) if len(sel) else np.zeros(0, dtype=np.int64)
    p_c, status, _ = device.em_pc_host(k[take], n[take], c[take], sub_off, pe[sel], nasc=nasc,
                                       device=_lib.current_device_index())
    p_cs = {keys[g]: float(p_c[i]) for i, g in enumerate(sel) if status[i] == 0}
    with open(p_c_path, 'w') as f:
        f.write(f'{",".join(group_by)},p_c\n')
        for key in sorted(p_cs.keys()):
            formatted_key = key if isinstance(key, str) else ','.join(key)
            f.write(f'{formatted_key},{p_cs[key]}\n')
    return p_c_path
