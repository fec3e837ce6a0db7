"""GPU: the two command-level orchestrators (count, estimate) run the whole chain on the reference's fixture inputs
and reproduce its output directory (the reference itself has no end-to-end test of count() / estimate())."""
import gzip
import io
import os
import pickle
import shutil

import numpy as np
import pandas as pd
import pytest

from helpers import gene_infos, materialize, ref_path, ref_rows, ref_text

pytestmark = pytest.mark.gpu


def test_count_smartseq_no_splicing(gpu_ctx, tmp_path):
    from dynast_b200 import count
    out = str(tmp_path / 'count')
    res = count.count(ref_path('smartseq_align', 'Aligned.sortedByCoord.out.bam'), gene_infos(), out, strand='unstranded',
                      umi_tag=None, barcode_tag='RG', gene_tag='GX', velocity=False, n_threads=2)
    for name in ('conversions.csv', 'alignments.csv'):
        assert open(os.path.join(out, name)).read() == ref_text('smartseq_count', name), name
    # count() passes conversions={'TC'} to count_conversions (count.py:306-318), which changes the tie order of the
    # no-UMI path relative to the committed fixture (written with conversions=None): compared byte for byte with what the
    # REFERENCE's count_conversions writes for conversions=['TC'] (oracle/make_golden.py::smartseq_counts_tc)
    from helpers import golden_text
    assert open(os.path.join(out, 'counts_TC.csv')).read() == golden_text('smartseq_count_tc', 'counts_TC.csv')
    assert sorted(open(os.path.join(out, 'counts_TC.csv')).read().splitlines()) == sorted(ref_text('smartseq_count', 'counts_TC.csv').splitlines())
    a = pd.read_csv(os.path.join(out, 'rates.csv')).sort_values('barcode').reset_index(drop=True)
    b = pd.read_csv(io.StringIO(ref_text('smartseq_count', 'rates.csv'))).sort_values('barcode').reset_index(drop=True)
    pd.testing.assert_frame_equal(a, b, check_exact=True)
    with gzip.open(os.path.join(out, 'convs.pkl.gz'), 'rb') as f:
        assert pickle.load(f) == frozenset({frozenset({'TC'})})
    assert any(n.startswith('run_info_') for n in os.listdir(out))
    counts = pd.read_csv(os.path.join(out, 'counts_TC.csv'))
    assert res.X.shape == (counts['barcode'].nunique(), counts['GX'].nunique())
    assert float(res.X.sum()) == float(counts['transcriptome'].sum())
    assert {'X_n_TC', 'X_l_TC'} <= set(res.layers)
    # second call: the parse stage is skipped (outputs exist, count.py:138-171) and the result is the same
    before = os.path.getmtime(os.path.join(out, 'alignments.csv'))
    count.count(ref_path('smartseq_align', 'Aligned.sortedByCoord.out.bam'), gene_infos(), out, strand='unstranded',
                umi_tag=None, barcode_tag='RG', velocity=False)
    assert os.path.getmtime(os.path.join(out, 'alignments.csv')) == before


def test_count_control_with_snp_detection(gpu_ctx, tmp_path):
    """UMI fixture, --control --snp-threshold 0.5: coverage.csv and snps.csv byte-equal, counts use the SNP mask."""
    from dynast_b200 import count
    from velocity_cases import FIXTURE_GTF
    out = str(tmp_path / 'control')
    # the default path: GTF in, velocity assigned on the device
    res = count.count(ref_path('SRR11683995_align', 'Aligned.sortedByCoord.out.bam'), FIXTURE_GTF, out, umi_tag='UB',
                      barcode_tag='CB', control=True, snp_threshold=0.5)
    assert res is None
    assert open(os.path.join(out, 'alignments.csv')).read() == ref_text('SRR11683995_count', 'alignments.csv')
    with gzip.open(os.path.join(out, 'genes.pkl.gz'), 'rb') as f:
        assert len(pickle.load(f)) == len(gene_infos())
    for name in ('coverage.csv', 'snps.csv', 'counts_TC.csv'):
        assert open(os.path.join(out, name)).read() == ref_text('SRR11683995_count_control', name), name


def test_estimate_alpha_from_fixture_count_dir(gpu_ctx, tmp_path):
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
    res = estimate.estimate([str(cdir)], out, reads='total', cell_threshold=0, cell_gene_threshold=0, method='alpha')
    est = 'SRR11683995_estimate_alpha'
    assert open(os.path.join(out, 'p_e.csv')).read() == ref_text(est, 'p_e.csv')
    a = pd.read_csv(os.path.join(out, 'A_total_TC.csv')).sort_values(['barcode', 'GX', 'conversion', 'base']).reset_index(drop=True)
    b = pd.read_csv(io.StringIO(ref_text(est, 'A_total_TC.csv'))).sort_values(['barcode', 'GX', 'conversion', 'base']).reset_index(drop=True)
    pd.testing.assert_frame_equal(a, b)
    pc = pd.read_csv(os.path.join(out, 'p_c_TC.csv'))
    want = pd.read_csv(io.StringIO(ref_text(est, 'p_c_TC.csv')))
    assert list(pc['barcode']) == list(want['barcode'])
    np.testing.assert_allclose(pc['p_c'], want['p_c'], rtol=2.5e-6)
    pi = pd.read_csv(os.path.join(out, 'pi_total_TC.csv'))
    wpi = pd.read_csv(io.StringIO(ref_text(est, 'pi_total_TC.csv')))
    assert list(pi['barcode']) == list(wpi['barcode'])
    assert np.abs(pi['pi'] - wpi['pi']).mean() < 0.02
    al = pd.read_csv(os.path.join(out, 'alpha_total_TC.csv'))
    wal = pd.read_csv(io.StringIO(ref_text(est, 'alpha_total_TC.csv')))
    assert list(al['barcode']) == list(wal['barcode'])
    np.testing.assert_allclose(al['alpha'], wal['alpha'], rtol=0.08)
    assert 'p_e' in res.obs.columns and 'p_c_TC' in res.obs.columns and 'total_TC_alpha' in res.obs.columns
    assert {'labeled_TC_est', 'unlabeled_TC_est'} <= set(res.layers)
