"""Output assembly -- the (barcode x gene) matrices of dynast/utils.py:350-631 (reference).

`counts_to_matrix` is the only one that touches every read: its group-by runs on the device (dyn_unique_counts on
(barcode, gene) keys).  The split_* functions are element-wise arithmetic on already reduced sparse matrices and
stay vectorised scipy.  `results_to_adata` returns an `anndata.AnnData` when anndata is importable and otherwise a
`Results` record with the same X / obs / var / layers contents (anndata and h5py are not in this image)."""
from typing import Dict, FrozenSet, List, Optional, Tuple, Union

import numpy as np
import pandas as pd
from scipy import sparse

VELOCITY_BLACKLIST = ['unassigned', 'ambiguous']   # dynast/config.py:48


def flatten_iter(it):
    for x in it:
        if isinstance(x, (list, tuple, set, frozenset)):
            yield from flatten_iter(x)
        else:
            yield x


def downsample_counts(df_counts: pd.DataFrame, proportion: Optional[float] = None, count: Optional[int] = None,
                      seed: Optional[int] = None, group_by: Optional[List[str]] = None) -> pd.DataFrame:
    """utils.downsample_counts (utils.py:307-347): rows kept are drawn with numpy's Generator.choice(shuffle=False,
    replace=False), seeded like the reference, so that the same seed keeps the same rows."""
    rng = np.random.default_rng(seed)
    if not group_by:
        if bool(proportion) == bool(count):
            raise Exception('Only one of `proportion` or `count` must be provided.')
        n_keep = int(df_counts.shape[0] * proportion) if proportion is not None else count
        return df_counts.iloc[rng.choice(df_counts.shape[0], n_keep, shuffle=False, replace=False)]
    if not count:
        raise Exception('`count` must be provided when using `group_by`')
    parts = []
    for _, df_group in df_counts.groupby(group_by, sort=False, observed=True):
        if df_group.shape[0] > count:
            df_group = df_group.iloc[rng.choice(df_group.shape[0], count, shuffle=False, replace=False)]
        parts.append(df_group)
    return pd.concat(parts)


def counts_to_matrix(df_counts: pd.DataFrame, barcodes: List[str], features: List[str], barcode_column: str = 'barcode',
                     feature_column: str = 'GX') -> sparse.csr_matrix:
    """utils.counts_to_matrix (utils.py:350-380): reads per (barcode, feature) as a float32 CSR matrix with rows /
    columns in the order of `barcodes` / `features`."""
    from .preprocessing.snp import unique_counts
    barcode_indices = {b: i for i, b in enumerate(barcodes)}
    feature_indices = {f: i for i, f in enumerate(features)}
    shape = (len(barcodes), len(features))
    if len(df_counts) == 0:
        return sparse.csr_matrix(shape, dtype=np.float32)
    bc_codes, bc_names = pd.factorize(df_counts[barcode_column].astype(object), sort=False)
    ft_codes, ft_names = pd.factorize(df_counts[feature_column].astype(object), sort=False)
    rows = np.array([barcode_indices[b] for b in bc_names], dtype=np.uint64)[bc_codes]   # KeyError like the reference
    cols = np.array([feature_indices[f] for f in ft_names], dtype=np.uint64)[ft_codes]
    keys, count, _ = unique_counts((rows << np.uint64(32)) | cols)
    return sparse.csr_matrix((count.astype(np.float32), ((keys >> np.uint64(32)).astype(np.int64), (keys & np.uint64(0xffffffff)).astype(np.int64))),
                             shape=shape, dtype=np.float32)


def split_counts(df_counts: pd.DataFrame, barcodes: List[str], features: List[str], barcode_column: str = 'barcode',
                 feature_column: str = 'GX', conversions: FrozenSet[str] = frozenset({'TC'})):
    """utils.split_counts (utils.py:383-418): (reads without any of `conversions`, reads with at least one)."""
    convs = list(conversions)
    unlabeled = counts_to_matrix(df_counts[(df_counts[convs] == 0).all(axis=1)], barcodes, features, barcode_column, feature_column)
    labeled = counts_to_matrix(df_counts[(df_counts[convs] > 0).any(axis=1)], barcodes, features, barcode_column, feature_column)
    return unlabeled, labeled


def split_matrix_pi(matrix, pis: Dict[Tuple[str, str], float], barcodes: List[str], features: List[str]):
    """utils.split_matrix_pi (utils.py:421-455): (pi matrix, unlabeled = x (1 - pi), labeled = x pi) at the cells
    that have a pi estimate."""
    barcode_indices = {b: i for i, b in enumerate(barcodes)}
    feature_indices = {f: i for i, f in enumerate(features)}
    rows, cols, vals = [], [], []
    for (barcode, gx), pi in pis.items():
        if barcode not in barcode_indices or gx not in feature_indices:
            continue
        try:
            pi = float(pi)
        except ValueError:
            continue
        rows.append(barcode_indices[barcode]); cols.append(feature_indices[gx]); vals.append(pi)
    shape = (len(barcodes), len(features))
    rows, cols = np.asarray(rows, dtype=np.int64), np.asarray(cols, dtype=np.int64)
    pi = np.asarray(vals, dtype=np.float64)
    m = sparse.csr_matrix(matrix) if not sparse.issparse(matrix) else matrix.tocsr()
    val = np.asarray(m[rows, cols]).reshape(-1).astype(np.float64) if len(rows) else np.zeros(0)
    mk = lambda v: sparse.csr_matrix((v.astype(np.float32), (rows, cols)), shape=shape, dtype=np.float32)
    out = mk(pi), mk(val * (1 - pi)), mk(val * pi)
    for o in out:
        o.eliminate_zeros()     # lil_matrix assignment of 0 stores nothing
    return out


def split_matrix_alpha(unlabeled_matrix, labeled_matrix, alphas: Dict[str, float], barcodes: List[str]):
    """utils.split_matrix_alpha (utils.py:458-495): labeled_est = min(labeled / clip(alpha, 0, 1), total) per
    barcode row, unlabeled_est = total - labeled_est; rows without an alpha stay empty."""
    dense = not sparse.issparse(unlabeled_matrix)
    U = np.asarray(unlabeled_matrix, dtype=np.float64) if dense else unlabeled_matrix.toarray().astype(np.float64)
    L = np.asarray(labeled_matrix, dtype=np.float64) if dense else labeled_matrix.toarray().astype(np.float64)
    total = U + L
    est_l = np.zeros(total.shape, dtype=np.float32)     # the reference stores into float32 lil matrices,
    est_u = np.zeros(total.shape, dtype=np.float32)     # so the subtraction below happens in float32
    total32 = total.astype(np.float32)
    barcode_indices = {b: i for i, b in enumerate(barcodes)}
    for barcode, alpha in alphas.items():
        if barcode not in barcode_indices:
            continue
        try:
            alpha = np.clip(float(alpha), 0, 1)
        except ValueError:
            continue
        row = barcode_indices[barcode]
        with np.errstate(divide='ignore', invalid='ignore'):
            est_l[row] = np.minimum(L[row] / alpha, total[row]).astype(np.float32)
        est_u[row] = total32[row] - est_l[row]
    return sparse.csr_matrix(est_u), sparse.csr_matrix(est_l)


class Results:
    """Stand-in for anndata.AnnData when anndata is not installed: same X / obs / var / layers."""

    def __init__(self, X, obs, var, layers):
        self.X, self.obs, self.var, self.layers = X, obs, var, layers

    def to_anndata(self):
        import anndata
        return anndata.AnnData(X=self.X, obs=self.obs, var=self.var, layers=self.layers)


def results_to_adata(df_counts: pd.DataFrame, conversions=frozenset({frozenset({'TC'})}), gene_infos: Optional[dict] = None,
                     pis: Optional[dict] = None, alphas: Optional[dict] = None):
    """utils.results_to_adata (utils.py:498-631): X = transcriptome reads; layers X_n_/X_l_, total, unlabeled_/
    labeled_, velocity layers, *_est, *_pi_g; obs *_alpha."""
    if pis is not None and alphas is not None:
        raise Exception('Only one of `pis` or `alphas` may be provided.')
    pis = pis or {}
    alphas = alphas or {}
    all_conversions = sorted(flatten_iter(conversions))
    transcriptome_exists = df_counts['transcriptome'].any()
    transcriptome_only = df_counts['transcriptome'].all()
    velocities = df_counts['velocity'].unique()
    barcodes = sorted(df_counts['barcode'].unique())
    features = sorted(df_counts['GX'].unique())
    add_names = gene_infos is not None
    obs = pd.DataFrame(index=pd.Series(barcodes, name='barcode'))
    var = pd.DataFrame(index=pd.Series(features, name='gene_id' if add_names else 'gene_name'))
    if add_names:
        var['gene_name'] = pd.Categorical([gene_infos.get(f, {}).get('gene_name') for f in features])
    df_tx = df_counts[df_counts['transcriptome']]
    matrix = counts_to_matrix(df_tx, barcodes, features)
    layers = {}

    def add_group(df, prefix_n, prefix_l, pi_key, pi_layer, alpha_obs):
        for convs in conversions:
            convs = sorted(convs)
            other = list(set(all_conversions) - set(convs))
            join = '_'.join(convs)
            layers[f'{prefix_n}_{join}'], layers[f'{prefix_l}_{join}'] = split_counts(
                df[(df[other] == 0).all(axis=1)], barcodes, features, conversions=convs)
            pi = pis.get(pi_key, {}).get(tuple(convs))
            if pi is not None:
                layers[f'{pi_layer}_{join}_pi_g'], layers[f'{prefix_n}_{join}_est'], layers[f'{prefix_l}_{join}_est'] = split_matrix_pi(
                    layers[f'{prefix_n}_{join}'] + layers[f'{prefix_l}_{join}'], pi, barcodes, features)
            alpha = alphas.get(pi_key, {}).get(tuple(convs))
            if alpha is not None:
                obs[f'{alpha_obs}_{join}_alpha'] = obs.index.map(alpha)
                layers[f'{prefix_n}_{join}_est'], layers[f'{prefix_l}_{join}_est'] = split_matrix_alpha(
                    layers[f'{prefix_n}_{join}'], layers[f'{prefix_l}_{join}'], alpha, barcodes)

    if transcriptome_exists:
        add_group(df_tx, 'X_n', 'X_l', 'transcriptome', 'X', 'X')
    if not transcriptome_only:
        layers['total'] = counts_to_matrix(df_counts, barcodes, features)
        add_group(df_counts, 'unlabeled', 'labeled', 'total', 'total', 'total')
    for key in velocities:
        if key == 'unassigned':
            continue
        df_v = df_counts[df_counts['velocity'] == key]
        layers[key] = counts_to_matrix(df_v, barcodes, features)
        if key in VELOCITY_BLACKLIST:
            continue
        add_group(df_v, f'{key[0]}n', f'{key[0]}l', key, key, key)
    try:
        import anndata
        return anndata.AnnData(X=matrix, obs=obs, var=var, layers=layers)
    except ImportError:
        return Results(matrix, obs, var, layers)
