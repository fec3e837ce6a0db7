"""`dynast estimate` -- thin orchestrator with the stage order, options and output files of
dynast/estimate.py:15-429 (reference), calling the B200 operators of dynast_b200.estimation.

`--downsample*` is the reference's host-side row sampling (utils.py:307-347), reproduced with the same numpy generator.
The AnnData object is written only when `anndata` is importable; the layers / obs columns are returned either way."""
import datetime as dt
import gzip
import json
import os
import pickle
import sys
import time
from typing import Dict, List, Optional, Union

import pandas as pd

from . import constants, estimation, preprocessing, utils


def _read_pickle(path):
    with gzip.open(path, 'rb') as f:
        return pickle.load(f)


def estimate(
    count_dirs: List[str],
    out_dir: str,
    reads: Union[str, List[str]] = 'complete',
    barcodes: Optional[List[List[str]]] = None,
    groups: Optional[List[Dict[str, str]]] = None,
    ignore_groups_for_est: bool = True,
    genes: Optional[List[str]] = None,
    downsample: Optional[Union[int, float]] = None,
    downsample_mode: str = 'uniform',
    cell_threshold: int = 1000,
    cell_gene_threshold: int = 16,
    control_p_e: Optional[float] = None,
    control: bool = False,
    method: str = 'alpha',
    n_threads: int = 8,
    temp_dir: Optional[str] = None,
    nasc: bool = False,
    by_name: bool = False,
    seed: Optional[int] = None,
):
    """estimate.estimate (estimate.py:15-429). Returns the results object (AnnData or utils.Results; None for --control)."""
    t0 = time.time()
    start = dt.datetime.now()
    os.makedirs(out_dir, exist_ok=True)
    stats_path = os.path.join(out_dir, f'{constants.STATS_PREFIX}_{dt.datetime.strftime(start, "%Y%m%d_%H%M%S_%f")}.json')
    conversions = _read_pickle(os.path.join(count_dirs[0], constants.CONVS_FILENAME))
    for count_dir in count_dirs[1:]:
        if _read_pickle(os.path.join(count_dir, constants.CONVS_FILENAME)) != conversions:
            raise Exception(f"Conversions for {count_dir} doesn't match conversions for {count_dirs[0]}.")
    all_conversions = sorted(utils.flatten_iter(conversions))
    gene_infos = _read_pickle(os.path.join(count_dirs[0], constants.GENES_FILENAME))

    dfs = []
    for i, count_dir in enumerate(count_dirs):
        df = preprocessing.read_counts(os.path.join(count_dir, f'{constants.COUNTS_PREFIX}_{"_".join(all_conversions)}.csv'))
        if barcodes:
            df = df[df['barcode'].isin(barcodes[i])]
        if len(count_dirs) > 1:
            df['barcode'] = df['barcode'].astype(str) + f'-{i}'
        dfs.append(df)
    df_unc = pd.concat(dfs, ignore_index=True) if len(count_dirs) > 1 else dfs[0]
    df_unc['barcode'] = df_unc['barcode'].astype('category')
    df_unc = df_unc.drop(columns=['umi'])

    if isinstance(groups, list):
        groups = groups[0] if len(groups) == 1 else {f'{b}-{i}': g for i, gs in enumerate(groups) for b, g in gs.items()}
    group_cells = {}
    if groups:
        present = set(df_unc['barcode'])
        groups = {b: g for b, g in groups.items() if b in present}
        for b, g in groups.items():
            group_cells.setdefault(g, []).append(b)
        df_unc = df_unc[df_unc['barcode'].isin(groups.keys())].reset_index(drop=True)
        df_unc['group'] = df_unc['barcode'].map(groups).astype('category')

    if downsample:          # estimate.py:129-152
        is_count = int(downsample) == downsample
        by = {'cell': ['barcode'], 'group': ['group']}.get(downsample_mode)
        df_unc = utils.downsample_counts(df_unc, proportion=None if is_count else downsample, count=int(downsample) if is_count else None,
                                         seed=seed, group_by=by)

    transcriptome_any = bool(df_unc['transcriptome'].any())
    transcriptome_all = bool(df_unc['transcriptome'].all())
    if reads != 'complete':
        reads = [reads] if isinstance(reads, str) else list(reads)
        if 'transcriptome' in reads and not transcriptome_any:
            raise Exception('No reads are assigned to `transcriptome`, so estimation is not supported for this read group.')
        if 'total' in reads and transcriptome_all:
            raise Exception('All reads are assigned to `transcriptome`, so estimation is not supported for `total` read group.')
        for key in set(reads).intersection(('spliced', 'unspliced')):
            if not (df_unc['velocity'] == key).any():
                raise Exception(f'No reads are assigned to `{key}`, so estimation is not supported for this read group.')
    else:
        reads = (['transcriptome'] if transcriptome_any else []) + ([] if transcriptome_all else ['total'])
        reads += list(set(df_unc['velocity'].unique()) - set(constants.VELOCITY_BLACKLIST))

    df_counts = preprocessing.complement_counts(df_unc, gene_infos)
    if by_name:
        df_counts['GX'] = df_counts['GX'].apply(lambda gx: gene_infos[gx].get('gene_name') or gx)
    if genes:
        df_counts = df_counts[df_counts['GX'].isin(genes)]

    # ---- p_e (estimate.py:194-234)
    p_key = 'group' if groups else 'barcode'
    p_e_path = os.path.join(out_dir, constants.P_E_FILENAME)
    if control_p_e:
        df_b = df_counts[[p_key]].drop_duplicates().reset_index(drop=True)
        df_b['p_e'] = control_p_e
        df_b.to_csv(p_e_path, header=[p_key, 'p_e'], index=False)
    elif control:
        estimation.estimate_p_e_control(df_counts, p_e_path, conversions=conversions)
    elif nasc:
        rates = preprocessing.read_rates(os.path.join(count_dirs[0], f'{constants.RATES_PREFIX}.csv'))
        if groups:
            rates['group'] = rates['barcode'].map(groups)
        estimation.estimate_p_e_nasc(rates, p_e_path, group_by=[p_key])
    else:
        estimation.estimate_p_e(df_counts, p_e_path, conversions=conversions, group_by=[p_key])
    if control:
        return None
    p_es = estimation.read_p_e(p_e_path, group_by=[p_key])

    # ---- aggregates (estimate.py:236-255)
    aggregates_paths = {}
    for key in set(reads).union(['transcriptome'] if transcriptome_all else ['total']):
        df = preprocessing.subset_counts(df_counts, key)
        for convs in conversions:
            convs = sorted(convs)
            other = list(set(all_conversions) - set(convs))
            aggregates_paths.setdefault(key, {})[tuple(convs)] = preprocessing.aggregate_counts(
                df[(df[other] == 0).all(axis=1)] if other else df, os.path.join(out_dir, f'A_{key}_{"_".join(convs)}.csv'),
                conversions=convs)

    def read_agg(key, convs):
        df_a = preprocessing.read_aggregates(aggregates_paths[key][tuple(convs)])
        if groups:
            df_a['group'] = df_a['barcode'].map(groups).astype('category')
        return df_a

    # ---- p_c (estimate.py:257-275)
    p_cs = {}
    for convs in conversions:
        convs = sorted(convs)
        p_c_path = os.path.join(out_dir, f'{constants.P_C_PREFIX}_{"_".join(convs)}.csv')
        estimation.estimate_p_c(read_agg('transcriptome' if transcriptome_all else 'total', convs), p_es, p_c_path,
                                group_by=[p_key], threshold=cell_threshold, n_threads=n_threads, nasc=nasc)
        p_cs[tuple(convs)] = estimation.read_p_c(p_c_path, group_by=[p_key])

    # ---- pi / alpha (estimate.py:280-398)
    pis = alphas = None
    if method == 'pi_g':
        pi_key = 'barcode' if ignore_groups_for_est or not groups else 'group'
        pis = {}
        for key in reads:
            for convs in conversions:
                convs = sorted(convs)
                pi_path = os.path.join(out_dir, f'pi_{key}_{"_".join(convs)}.csv')
                estimation.estimate_pi(read_agg(key, convs), p_es, p_cs[tuple(convs)], pi_path, group_by=[pi_key, 'GX'],
                                       p_group_by=[p_key], n_threads=n_threads, threshold=cell_gene_threshold, seed=seed, nasc=nasc)
                _, _, pi = estimation.read_pi(pi_path, group_by=[pi_key, 'GX'])
                pis.setdefault(key, {})[tuple(convs)] = pi
        if groups and not ignore_groups_for_est:
            pis = {key: {convs: {(b, gx): v for (g, gx), v in pis[key][convs].items() for b in group_cells[g]}
                         for convs in pis[key]} for key in pis}
    elif method == 'alpha':
        pi_c_key = 'group' if groups else 'barcode'
        alpha_key = 'group' if groups and not ignore_groups_for_est else 'barcode'
        alphas = {}
        for key in reads:
            for convs in conversions:
                convs = sorted(convs)
                pi_c_path = os.path.join(out_dir, f'pi_{key}_{"_".join(convs)}.csv')
                estimation.estimate_pi(read_agg(key, convs), p_es, p_cs[tuple(convs)], pi_c_path, group_by=[pi_c_key],
                                       p_group_by=[p_key], n_threads=n_threads, threshold=cell_gene_threshold, seed=seed, nasc=nasc)
                _, _, pi_c = estimation.read_pi(pi_c_path, group_by=[pi_c_key])
                alpha_path = os.path.join(out_dir, f'alpha_{key}_{"_".join(convs)}.csv')
                estimation.estimate_alpha(preprocessing.subset_counts(df_counts, key), pi_c, alpha_path, conversions=convs,
                                          group_by=[alpha_key], pi_c_group_by=[pi_c_key])
                alphas.setdefault(key, {})[tuple(convs)] = estimation.read_alpha(alpha_path, group_by=[alpha_key])
        if groups and not ignore_groups_for_est:
            alphas = {key: {convs: {b: v for g, v in alphas[key][convs].items() for b in group_cells[g]}
                            for convs in alphas[key]} for key in alphas}
    else:
        raise Exception(f'Unrecognized method {method}')

    # ---- AnnData (estimate.py:400-427)
    result = utils.results_to_adata(df_counts, conversions, gene_infos=gene_infos if not by_name else None, pis=pis, alphas=alphas)
    obs = result.obs
    if groups:
        obs['group'] = obs.index.map(groups).astype('category')
    obs.reset_index(inplace=True)
    obs['p_e'] = obs[p_key].map(p_es).astype(float)
    for convs in conversions:
        convs = sorted(convs)
        obs[f'p_c_{"_".join(convs)}'] = obs[p_key].map(p_cs[tuple(convs)]).astype(float)
    obs.set_index('barcode', inplace=True)
    if hasattr(result, 'write'):
        result.write(os.path.join(out_dir, constants.ADATA_FILENAME), compression='gzip')
    with open(stats_path, 'w') as f:
        json.dump({'start_time': start.isoformat(), 'end_time': dt.datetime.now().isoformat(), 'elapsed': time.time() - t0,
                   'call': ' '.join(sys.argv), 'version': 'dynast_b200'}, f, indent=4)
    return result
