"""CPU: the velocity oracle (oracle/velocity_oracle.py: bam.py:177-228, 363-383 + the ngs_tools segment semantics) is
pinned to the GX / velocity columns of the reference's alignments.csv fixture through the synthesised annotation
(oracle/make_velocity_gtf.py), and the product's GTF parser / device tables agree with the oracle's."""
import numpy as np
import pytest

from helpers import ref_path, ref_text
from velocity_cases import FIXTURE_GTF, random_gtf


def test_velocity_oracle_reproduces_fixture_alignments(tmp_path):
    from oracle import parse_oracle, velocity_oracle
    annotation = velocity_oracle.parse_gtf(FIXTURE_GTF)
    c, a = str(tmp_path / 'c.csv'), str(tmp_path / 'a.csv')
    parse_oracle.parse_all_reads(ref_path('SRR11683995_align', 'Aligned.sortedByCoord.out.bam'), c, a, umi_tag='UB',
                                 barcode_tag='CB', velocity=True, annotation=annotation, strand='forward')
    assert open(a).read() == ref_text('SRR11683995_count', 'alignments.csv')
    assert open(c).read() == ref_text('SRR11683995_count', 'conversions.csv')
    # the protocol strand matters: the reads located by position would be lost on the other strand
    parse_oracle.parse_all_reads(ref_path('SRR11683995_align', 'Aligned.sortedByCoord.out.bam'), c, a, umi_tag='UB',
                                 barcode_tag='CB', velocity=True, annotation=annotation, strand='reverse')
    assert open(a).read() != ref_text('SRR11683995_count', 'alignments.csv')


def test_segment_collection_rules():
    from oracle.velocity_oracle import Segment, SegmentCollection
    c = SegmentCollection([Segment(10, 20), Segment(5, 12), Segment(20, 30), Segment(40, 50)])
    assert [tuple(s) for s in c] == [(5, 20), (20, 30), (40, 50)]            # overlapping merged, touching kept apart
    assert [tuple(s) for s in c.invert(Segment(0, 60))] == [(0, 5), (30, 40), (50, 60)]
    r = SegmentCollection.from_positions([7, 8, 9, 30, 31, 8, 9, 10])       # mates concatenated: duplicates, unsorted
    assert [tuple(s) for s in r] == [(7, 11), (30, 32)]
    assert r.is_overlapping(c) and not r.is_subset(c)
    assert SegmentCollection.from_positions([41, 42]).is_subset(c)
    assert not SegmentCollection.from_positions([19, 20]).is_subset(c)      # spans two touching segments
    assert SegmentCollection([]).start == 0 and SegmentCollection([]).end == 0


@pytest.mark.parametrize('which', ['fixture', 'random'])
def test_product_gtf_parser_matches_oracle(tmp_path, which):
    from dynast_b200 import gtf
    from oracle import velocity_oracle
    path = FIXTURE_GTF
    if which == 'random':
        path = str(tmp_path / 'r.gtf')
        random_gtf(path, seed=11)
    g1, t1 = gtf.genes_and_transcripts_from_gtf(path)
    g2, t2 = velocity_oracle.parse_gtf(path)
    assert list(g1) == list(g2) and list(t1) == list(t2)
    for g in g1:
        assert tuple(g1[g]['segment']) == tuple(g2[g]['segment'])
        assert (g1[g]['chromosome'], g1[g]['strand'], g1[g]['gene_name'], g1[g]['transcripts']) == (
            g2[g]['chromosome'], g2[g]['strand'], g2[g]['gene_name'], g2[g]['transcripts'])
    for t in t1:
        assert [tuple(s) for s in t1[t]['exons']] == [tuple(s) for s in t2[t]['exons']]
        assert [tuple(s) for s in t1[t]['introns']] == [tuple(s) for s in t2[t]['introns']]
        assert tuple(t1[t]['segment']) == tuple(t2[t]['segment'])


def test_annotation_tables_layout(tmp_path):
    from dynast_b200 import gtf
    from oracle import velocity_oracle
    path = str(tmp_path / 'r.gtf')
    contigs = random_gtf(path, seed=5)
    genes, transcripts = gtf.genes_and_transcripts_from_gtf(path)
    header = ['zz_unused'] + contigs[::-1]              # BAM header order differs from the GTF's
    tab = gtf.AnnotationTables(genes, transcripts, header)
    assert tab.n_genes == len(genes)
    by_contig = velocity_oracle.by_contig(velocity_oracle.parse_gtf(path)[0])
    for c, name in enumerate(header):
        rows = tab.gene_ids[tab.contig_off[c]:tab.contig_off[c + 1]]
        assert rows == (velocity_oracle.gene_order(by_contig[name]) if name in by_contig else [])
        ends = tab.gene_end[tab.contig_off[c]:tab.contig_off[c + 1]]
        assert (tab.gene_pmax[tab.contig_off[c]:tab.contig_off[c + 1]] == np.maximum.accumulate(ends)).all() if len(ends) else True
    g = tab.gene_ids[3]
    trs = genes[g]['transcripts']
    assert tab.gene_tr_off[4] - tab.gene_tr_off[3] == len(trs)
    t0 = tab.gene_tr_off[3]
    ex = [(int(tab.exon_start[i]), int(tab.exon_end[i])) for i in range(tab.tr_exon_off[t0], tab.tr_exon_off[t0 + 1])]
    assert ex == [tuple(s) for s in transcripts[trs[0]]['exons']]
