"""Labeled fraction pi -- same names, arguments and CSV outputs as dynast/estimation/pi.py (reference).

The reference samples the posterior of dynast/models/pi.stan with NUTS (pystan 2.19) and reports the means of
alpha, beta and pi_g.  Here the same three posterior means are computed deterministically on the device
(csrc/pi.cu); `seed`, `model` and `n_threads` are accepted for signature compatibility and unused."""
from typing import Dict, List, Optional, Tuple, Union

import numpy as np
import pandas as pd
import torch

from .. import pipeline
from .. import _lib
from ._groups import csr_from_rows, group_codes, lookup


def read_pi(pi_path: str, group_by: Optional[List[str]] = None):
    """pi.read_pi (pi.py:13-36)."""
    if group_by is None:
        with open(pi_path, 'r') as f:
            return tuple(float(s) for s in f.readline().strip().split(','))
    df = pd.read_csv(pi_path, dtype={key: 'category' for key in group_by})
    df.set_index(group_by, inplace=True)
    return dict(df['alpha']), dict(df['beta']), dict(df['pi'])


def beta_mean(alpha: float, beta: float) -> float:
    """pi.beta_mean (pi.py:58-69)."""
    return alpha / (alpha + beta)


def beta_mode(alpha: float, beta: float) -> float:
    """pi.beta_mode (pi.py:72-98)."""
    pi = float('nan')
    if alpha > 1 and beta > 1:
        pi = (alpha - 1) / (alpha + beta - 2)
    elif alpha > 1 and beta <= 1:
        pi = 1
    elif alpha <= 1 and beta > 1:
        pi = 0
    return pi


def guess_beta_parameters(guess: float, strength: int = 5) -> Tuple[float, float]:
    """pi.guess_beta_parameters (pi.py:101-115)."""
    alpha = max(strength * guess, 0.1)
    beta = max(strength - alpha, 0.1)
    return alpha, beta


def fit_batch(k, n, count, off, p_e, p_c):
    """Posterior means (alpha, beta, pi) for CSR groups on the current device. Returns (values [G,3], status [G])."""
    dev = _lib.current_torch_device()
    ctx = _lib.current_context()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    out, status = pipeline.pi_fit(ctx, t(k.astype(np.int32)), t(n.astype(np.int32)), t(count.astype(np.int64)),
                                  t(off.astype(np.int64)), t(np.asarray(p_e, dtype=np.float64)),
                                  t(np.asarray(p_c, dtype=np.float64)))
    return out.cpu().numpy(), status.cpu().numpy()


def estimate_pi(
    df_aggregates: pd.DataFrame,
    p_e,
    p_c,
    pi_path: str,
    group_by: Optional[List[str]] = None,
    p_group_by: Optional[List[str]] = None,
    n_threads: int = 8,
    threshold: int = 16,
    seed: Optional[int] = None,
    nasc: bool = False,
    model=None,
) -> str:
    """pi.estimate_pi (pi.py:190-324)."""
    df_full = df_aggregates[(df_aggregates[['base', 'count']] > 0).all(axis=1)].copy()
    if p_group_by is not None:
        pcodes, pkeys = group_codes(df_full, p_group_by)
        df_full['p_e'] = lookup(p_e, pkeys, len(p_group_by))[pcodes] if len(df_full) else []
        df_full['p_c'] = lookup(p_c, pkeys, len(p_group_by))[pcodes] if len(df_full) else []
    else:
        df_full['p_e'] = p_e
        df_full['p_c'] = p_c
    df_full.dropna(subset=['p_c'], inplace=True)
    codes, keys = group_codes(df_full, group_by)
    k, n, c, off, order = csr_from_rows(codes, len(keys), df_full['conversion'].to_numpy(), df_full['base'].to_numpy(),
                                        df_full['count'].to_numpy(dtype=np.int64))
    p_es = df_full['p_e'].to_numpy(dtype=np.float64)[order]
    p_cs = df_full['p_c'].to_numpy(dtype=np.float64)[order]
    G = len(keys)
    if group_by is None:
        if not isinstance(p_e, float) or not isinstance(p_c, float):
            raise Exception('`p_e` and `p_c` must be floats when `group_by` is not provided')
        vals, status = fit_batch(k, n, c, off, [p_e], [p_c])
        if status[0] != 0:
            raise RuntimeError('pi fit failed')
        alpha, beta, pi = (float(x) for x in vals[0])
        pi = beta_mode(alpha, beta) if nasc else pi
        with open(pi_path, 'w') as f:
            f.write(f'{alpha},{beta},{pi}')
        return pi_path
    nonempty = off[1:] > off[:-1]
    first = off[:-1].clip(max=max(len(k) - 1, 0))
    for name, arr in (('p_e', p_es), ('p_c', p_cs)):
        if len(arr):
            lo = np.minimum.reduceat(arr, off[:-1][nonempty]) if nonempty.any() else []
            hi = np.maximum.reduceat(arr, off[:-1][nonempty]) if nonempty.any() else []
            if np.any(np.asarray(lo) != np.asarray(hi)):
                raise Exception('`p_e` and `p_c` for each aggregate group must be a constant')
    totals = np.where(nonempty, np.add.reduceat(c, first) if len(c) else 0, 0)
    sel = np.flatnonzero(nonempty & (totals >= threshold))
    lens = (off[1:] - off[:-1])[sel]
    sub_off = np.zeros(len(sel) + 1, dtype=np.int64)
    np.cumsum(lens, out=sub_off[1:])
    take = np.concatenate([np.arange(off[g], off[g + 1]) for g in sel]) if len(sel) else np.zeros(0, dtype=np.int64)
    pis = {}
    if len(sel):
        vals, status = fit_batch(k[take], n[take], c[take], sub_off, p_es[off[sel]], p_cs[off[sel]])
        for i, g in enumerate(sel):
            if status[i] != 0:
                continue   # the reference drops groups whose sampler raises (pi.py:298-304)
            alpha, beta, pi = (float(x) for x in vals[i])
            pis[keys[g]] = (alpha, beta, beta_mode(alpha, beta) if nasc else pi)
    with open(pi_path, 'w') as f:
        f.write(f'{",".join(group_by)},alpha,beta,pi\n')
        for key in sorted(pis.keys()):
            alpha, beta, pi = pis[key]
            formatted_key = key if isinstance(key, str) else ','.join(key)
            f.write(f'{formatted_key},{alpha},{beta},{pi}\n')
    return pi_path
