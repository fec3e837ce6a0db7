"""GPU: the per-(cell, gene) labeled-fraction fit (estimate(method='pi_g'): estimate.py:282-327, pi.py:253-296) -- on the
device path (pipeline.count_to_estimate(method='pi_g')) and through the operators on the reference's fixture."""
import gzip
import io
import os
import pickle

import numpy as np
import pandas as pd
import pytest
import torch

from helpers import gene_infos, materialize, ref_text

pytestmark = pytest.mark.gpu


def test_device_path_pi_g(gpu_ctx, oracle_lib):
    from dynast_b200 import pipeline
    from oracle import pi_oracle
    b = pipeline.SynthBatch(400_000, seed=17, n_barcodes=12, n_genes=40, device=0, reads_per_umi=1.3)
    out = pipeline.TallyBuffers(b.n_reads, 3 * b.n_reads, device=0)
    res = pipeline.count_to_estimate(gpu_ctx, b.records, b.rec_off, b.barcode, b.umi, b.gene, b.score, b.strand, 12, 40, out,
                                     cell_threshold=500, method='pi_g', cell_gene_threshold=16)
    torch.cuda.synchronize()
    sel = res['selected'].long()
    counts = out.counts[sel].cpu().numpy().astype(np.int64)
    strand = b.strand[sel].cpu().numpy()
    comp = counts.copy()
    comp[strand != 0, :12] = counts[strand != 0, :12][:, ::-1]
    comp[strand != 0, 12:] = counts[strand != 0, 12:][:, ::-1]
    df = pd.DataFrame({'barcode': b.barcode[sel].cpu().numpy(), 'gene': b.gene[sel].cpu().numpy(), 'k': comp[:, 10], 'n': comp[:, 15]})
    pairs = df.groupby(['barcode', 'gene']).size()
    gene_bits = 6
    got_pair = res['pi_g_pair'].cpu().numpy()
    assert list(got_pair) == [int(bc) << gene_bits | int(g) for bc, g in pairs.index]          # every pair that occurs, sorted
    assert (res['pi_g_reads'].cpu().numpy() == pairs.to_numpy()).all()
    st = res['pi_g_status'].cpu().numpy()
    em_ok = res['em_status'].cpu().numpy() == 0
    want_fit = np.array([em_ok[bc] and cnt >= 16 for (bc, g), cnt in pairs.items()])
    assert ((st >= 0) == want_fit).all() and want_fit.sum() > 50
    # each pair's fit equals a direct fit of its own histogram, and the QUADPACK oracle on a few of them
    gk, gn, gc, goff = (t.cpu().numpy() for t in res['pi_g_hist'])
    vals = res['pi_g'].cpu().numpy()
    p_e, p_c = res['p_e'].cpu().numpy(), res['p_c'].cpu().numpy()
    checked = 0
    for j in np.flatnonzero(st == 0)[::17][:6]:
        bc, g = got_pair[j] >> gene_bits, got_pair[j] & ((1 << gene_bits) - 1)
        sub = df[(df.barcode == bc) & (df.gene == g)]
        h = sub.groupby(['k', 'n']).size()
        assert sorted(zip(gk[goff[j]:goff[j + 1]], gn[goff[j]:goff[j + 1]], gc[goff[j]:goff[j + 1]])) == sorted((k, n, c) for (k, n), c in h.items())
        want = pi_oracle.posterior_means([k for k, n in h.index], [n for k, n in h.index], h.to_numpy(), p_e[bc], p_c[bc], n_gl=24)
        np.testing.assert_allclose(vals[j], want, rtol=2e-5)
        checked += 1
    assert checked >= 3


def test_estimate_pi_g_on_fixture(gpu_ctx, tmp_path):
    """estimate(method='pi_g') through the operators on the reference's count directory: file layout, groups and --
    within the sampler's Monte-Carlo error -- values of the committed pi_total_TC.csv (barcode, GX rows)."""
    from dynast_b200 import estimate
    cdir = tmp_path / 'count'
    cdir.mkdir()
    materialize(cdir, 'ref', 'SRR11683995_count', 'counts_TC.csv')
    materialize(cdir, 'ref', 'SRR11683995_count', 'rates.csv')
    with gzip.open(cdir / 'convs.pkl.gz', 'wb') as f:
        pickle.dump(frozenset({frozenset({'TC'})}), f)
    with gzip.open(cdir / 'genes.pkl.gz', 'wb') as f:
        pickle.dump(gene_infos(), f)
    out = str(tmp_path / 'estimate')
    res = estimate.estimate([str(cdir)], out, reads='total', cell_threshold=0, cell_gene_threshold=0, method='pi_g')
    got = pd.read_csv(os.path.join(out, 'pi_total_TC.csv'))
    want = pd.read_csv(io.StringIO(ref_text('SRR11683995_estimate', 'pi_total_TC.csv')))
    assert list(got.columns) == ['barcode', 'GX', 'alpha', 'beta', 'pi']
    m = got.merge(want, on=['barcode', 'GX'], suffixes=('', '_ref'))
    assert len(m) == len(want) == len(got)                  # the same (cell, gene) groups
    assert np.abs(m['pi'] - m['pi_ref']).mean() < 0.03
    assert {'X_l_TC', 'X_n_TC'} <= set(res.layers) and any(k.endswith('_pi_g') or 'est' in k for k in res.layers)


def test_device_path_dual_conversion_groups(gpu_ctx):
    """`--conversion TC --conversion AG`: the histogram of each labelling group only holds reads WITHOUT a conversion of
    the other group (estimate.py:248-255); dedup and p_e use the flattened set (count.py:306, p_e.py:82-89)."""
    from dynast_b200 import pipeline
    b = pipeline.SynthBatch(300_000, seed=23, n_barcodes=10, n_genes=30, device=0, reads_per_umi=1.2)
    out = pipeline.TallyBuffers(b.n_reads, 3 * b.n_reads, device=0)
    res = pipeline.count_to_estimate(gpu_ctx, b.records, b.rec_off, b.barcode, b.umi, b.gene, b.score, b.strand, 10, 30, out,
                                     conversions=('AG', 'TC'), conversion_groups=[('TC',), ('AG',)], cell_threshold=500)
    one = pipeline.count_to_estimate(gpu_ctx, b.records, b.rec_off, b.barcode, b.umi, b.gene, b.score, b.strand, 10, 30, out,
                                     conversions=('AG', 'TC'), cell_threshold=500, with_pi=False)
    torch.cuda.synchronize()
    assert torch.equal(res['selected'], one['selected'])
    np.testing.assert_array_equal(res['p_e'].cpu().numpy(), one['p_e'].cpu().numpy())
    sel = res['selected'].long()
    counts = out.counts[sel].cpu().numpy().astype(np.int64)
    strand = b.strand[sel].cpu().numpy()
    comp = counts.copy()
    comp[strand != 0, :12] = counts[strand != 0, :12][:, ::-1]
    comp[strand != 0, 12:] = counts[strand != 0, 12:][:, ::-1]
    df = pd.DataFrame(comp, columns=pipeline.CONVERSION_COLUMNS + pipeline.BASE_COLUMNS)
    df['barcode'] = b.barcode[sel].cpu().numpy()
    assert set(res['groups']) == {'TC', 'AG'}
    for conv, other, base in (('TC', 'AG', 'T'), ('AG', 'TC', 'A')):
        g = res['groups'][conv]
        k, n, c, off = (t.cpu().numpy() for t in g['hist'])
        sub = df[df[other] == 0]
        want = sub.groupby(['barcode', conv, base]).size()
        got = {(bc, int(k[j]), int(n[j])): int(c[j]) for bc in range(10) for j in range(off[bc], off[bc + 1])}
        assert got == {(int(bc), int(kk), int(nn)): int(v) for (bc, kk, nn), v in want.items()}
        new = df[df[conv] > 0].groupby('barcode').size().reindex(range(10), fill_value=0).to_numpy()
        total = df.groupby('barcode').size().reindex(range(10), fill_value=0).to_numpy()
        ok = (g['em_status'].cpu().numpy() == 0) & (g['pi_status'].cpu().numpy() == 0)
        assert ok.sum() >= 5
        pi_c = g['pi'].cpu().numpy()[:, 2]
        np.testing.assert_allclose(g['alpha'].cpu().numpy()[ok], (new / total / pi_c)[ok], rtol=1e-9)
