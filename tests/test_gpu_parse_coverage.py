"""GPU: BAM -> conversions.csv / alignments.csv / index (parse_all_reads), coverage.csv, snps.csv and
select_alignments against the reference's fixtures, byte for byte.
Mirrors tests/preprocessing/test_bam.py:40-152, test_coverage.py:15-45, test_snp.py:21-61 of the reference."""
import gzip
import io
import pickle

import pandas as pd
import pytest

from helpers import gene_infos, materialize, ref_path, ref_rows, ref_text

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('align,count,kw,nasc', [
    ('SRR11683995_align', 'SRR11683995_count', dict(umi_tag='UB', barcode_tag='CB', velocity=True), False),
    ('smartseq_align', 'smartseq_count', dict(umi_tag=None, barcode_tag='RG', velocity=False), False),
    ('nasc_align', 'nasc_count', dict(umi_tag=None, barcode_tag='RG', velocity=False), True),
])
def test_parse_all_reads_bytes(gpu_ctx, tmp_path, align, count, kw, nasc):
    from dynast_b200 import gtf
    from dynast_b200.preprocessing import bam
    from velocity_cases import FIXTURE_GTF
    genes, transcripts = gene_infos(), None
    if kw['velocity']:   # velocity assignment on the device from the annotation oracle/make_velocity_gtf.py synthesised
        genes, transcripts = gtf.genes_and_transcripts_from_gtf(FIXTURE_GTF)
    c, a, i = (str(tmp_path / x) for x in ('conversions.csv', 'alignments.csv', 'conversions.idx'))
    got = bam.parse_all_reads(ref_path(align, 'Aligned.sortedByCoord.out.bam'), c, a, i, genes, transcripts, strand='forward',
                              gene_tag='GX', barcodes=None, n_threads=1, temp_dir=str(tmp_path), nasc=nasc, **kw)
    assert got == (c, a, i)
    assert open(c).read() == ref_text(count, 'conversions.csv')
    assert open(a).read() == ref_text(count, 'alignments.csv')
    with gzip.open(ref_path(count, 'conversions.idx'), 'rb') as f, gzip.open(i, 'rb') as g:
        assert pickle.load(f) == pickle.load(g)


def test_parse_velocity_injected_assignments(gpu_ctx, tmp_path):
    """`velocity_assignments` overrides the device assignment (what round 1 tested); without an annotation and
    without assignments the call refuses."""
    from dynast_b200.preprocessing import bam
    path = ref_path('SRR11683995_align', 'Aligned.sortedByCoord.out.bam')
    with pytest.raises(ValueError):
        bam.parse_all_reads(path, 'c', 'a', 'i', {}, None, umi_tag='UB', barcode_tag='CB')
    lookup = {(r['read_id'], int(r['index'])): (r['GX'], r['velocity']) for r in ref_rows('SRR11683995_count', 'alignments.csv')}
    c, a, i = (str(tmp_path / x) for x in ('conversions.csv', 'alignments.csv', 'conversions.idx'))
    bam.parse_all_reads(path, c, a, i, gene_infos(), None, umi_tag='UB', barcode_tag='CB', velocity_assignments=lookup)
    assert open(a).read() == ref_text('SRR11683995_count', 'alignments.csv')


def test_parse_velocity_unknown_gene_tag_raises(gpu_ctx, tmp_path):
    """A GX value absent from the annotation of the read's contig is a KeyError in the reference (bam.py:186, 203)."""
    from dynast_b200 import gtf
    from dynast_b200.preprocessing import bam
    from velocity_cases import FIXTURE_GTF
    genes, transcripts = gtf.genes_and_transcripts_from_gtf(FIXTURE_GTF)
    tagged = next(r['GX'] for r in ref_rows('SRR11683995_count', 'alignments.csv') if r['transcriptome'] == 'True')
    genes = {g: v for g, v in genes.items() if g != tagged}
    with pytest.raises(KeyError):
        bam.parse_all_reads(ref_path('SRR11683995_align', 'Aligned.sortedByCoord.out.bam'), str(tmp_path / 'c'), str(tmp_path / 'a'),
                            str(tmp_path / 'i'), genes, transcripts, umi_tag='UB', barcode_tag='CB')


def _selected(tmp_path):
    from dynast_b200.preprocessing import bam
    ali = materialize(tmp_path, 'ref', 'SRR11683995_count', 'alignments.csv')
    return bam.select_alignments(bam.read_alignments(ali))


def test_select_alignments(gpu_ctx, tmp_path):
    """tests/preprocessing/test_bam.py:31-38 of the reference + equality with a pandas restatement."""
    sel = _selected(tmp_path)
    df = pd.read_csv(io.StringIO(ref_text('SRR11683995_count', 'alignments.csv')), na_filter=False)
    want = df.sort_values(['transcriptome', 'score']).drop_duplicates(subset=['barcode', 'umi', 'GX'], keep='last')
    assert sel == set(zip(want['read_id'], want['index']))


def test_calculate_coverage_bytes(gpu_ctx, tmp_path):
    from dynast_b200.preprocessing import bam, coverage
    alignments = _selected(tmp_path)
    dfc = bam.read_conversions(materialize(tmp_path, 'ref', 'SRR11683995_count', 'conversions.csv'))
    dfc = dfc[[key in alignments for key in zip(dfc['read_id'], dfc['index'])]]
    dfc = dfc.loc[dfc['conversion'].isin(('TC', 'AG')), ['contig', 'genome_i']]
    conversions = {contig: set(part['genome_i']) for contig, part in dfc.groupby('contig', sort=False, observed=True)}
    out = str(tmp_path / 'coverage.csv')
    assert coverage.calculate_coverage(ref_path('SRR11683995_align', 'Aligned.sortedByCoord.out.bam'), conversions, out,
                                       alignments=alignments, umi_tag='UB', barcode_tag='CB', barcodes=None,
                                       temp_dir=str(tmp_path), velocity=True) == out
    assert open(out).read() == ref_text('SRR11683995_count_control', 'coverage.csv')
    cov = coverage.read_coverage(out)
    assert sum(len(v) for v in cov.values()) == 137


def test_extract_conversions_and_detect_snps(gpu_ctx, tmp_path):
    from dynast_b200.preprocessing import coverage, snp
    alignments = _selected(tmp_path)
    conv = materialize(tmp_path, 'ref', 'SRR11683995_count', 'conversions.csv')
    idx = materialize(tmp_path, 'ref', 'SRR11683995_count', 'conversions.idx')
    got = snp.extract_conversions(conv, idx, alignments=alignments, quality=35, n_threads=2)
    df = pd.read_csv(conv)
    dfi = df.set_index(['read_id', 'index'])
    dfs = dfi[(dfi['quality'] > 35) & (dfi.index.isin(alignments))].reset_index()
    truth = {}
    for (contig, genome_i, conversion), count in dict(dfs.groupby(['contig', 'genome_i', 'conversion']).size()).items():
        truth.setdefault(conversion, {}).setdefault(str(contig), {})[genome_i] = count
    assert {c: {str(k): v for k, v in d.items()} for c, d in got.items()} == truth
    cov = coverage.read_coverage(materialize(tmp_path, 'ref', 'SRR11683995_count_control', 'coverage.csv'))
    out = str(tmp_path / 'snps.csv')
    assert snp.detect_snps(conv, idx, cov, out, alignments=alignments, conversions=frozenset({'TC', 'AG'}), quality=27,
                           threshold=0.5, n_threads=2) == out
    assert open(out).read() == ref_text('SRR11683995_count_control', 'snps.csv')
    assert snp.read_snps(out) == snp.read_snp_csv(out)


def test_consensus_groups_from_bam(gpu_ctx, tmp_path):
    """Grouping + device consensus on the UMI fixture BAM equals the oracle run on the same groups."""
    import numpy as np
    from oracle import bamio as obam
    from oracle.consensus_oracle import call_consensus
    from dynast_b200.preprocessing import consensus
    from test_gpu_consensus import decode
    path = ref_path('SRR11683995_align', 'Aligned.sortedByCoord.out.bam')
    infos = gene_infos()
    # the fixture holds no UMI group with two reads: group by gene alone so that real multi-read groups form
    res = consensus.call_consensus_groups(path, infos, strand='forward', umi_tag=None, barcode_tag=None, gene_tag='GX',
                                          quality=27)
    _, refs, reads = obam.read_bam(path)
    groups = {}
    for r in reads:
        if r.is_secondary or r.is_unmapped or r.is_duplicate or not r.has_tag('GX'):
            continue
        groups.setdefault((r.get_tag('GX'), 'NA', 'NA'), []).append(r)
    multi = {k: v for k, v in groups.items() if len(v) >= 2}
    assert len(res['gene']) == len(multi) > 0
    assert (res['status'] == 0).all()
    out = dict(records=res['records'], off=res['off'], nm=res['nm'], md=res['md'], md_off=res['md_off'])
    for g in range(len(res['gene'])):
        members = multi[(res['gene'][g], res['barcode'][g], res['umi'][g])]
        want = call_consensus(members, quality=27)
        got = decode(out, g)
        assert (got['start'], got['cigar'], got['seq'], got['qual'], got['md'], got['nm']) == (
            want['start'], want['cigar'], want['seq'], want['qual'], want['md'], want['nm'])
        assert res['n_reads'][g] == len(members)
        assert res['AS'][g] == sum(m.get_tag('AS') for m in members)
