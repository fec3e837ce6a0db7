"""GPU: dyn_velocity (csrc/velocity.cu) against the velocity oracle (bam.py:177-228, 363-383) on random annotations
and alignments: overlapping genes on both strands, touching / overlapping exons, spliced / deleted / inserted /
clipped CIGARs, read pairs (union of the mates' positions), reads with and without a gene tag, the three protocol
strands and --strict-exon-overlap.  The fixture pin is tests/test_gpu_parse_coverage.py::test_parse_all_reads_bytes."""
import numpy as np
import pytest
import torch

from velocity_cases import md_for, random_alignments, random_gtf, reference_positions

pytestmark = pytest.mark.gpu

NAMES = {None: 0, 'spliced': 1, 'unspliced': 2, 'ambiguous': 3}


def _pack(alignments):
    from dynast_b200 import records
    units = []
    for al in alignments:
        recs = []
        for pos, cigar, flag in al['units']:
            l_seq = sum(n for op, n in cigar if op in (0, 1, 4, 7, 8))
            recs.append(records.pack_one(pos, al['ref_id'], flag, cigar, 'A' * l_seq, [30] * l_seq, md_for(cigar)))
        units.append(recs)
    return records.pack_units(units)


@pytest.mark.parametrize('strand', ['forward', 'reverse', 'unstranded'])
@pytest.mark.parametrize('strict', [False, True])
def test_velocity_kernel_matches_oracle(gpu_ctx, tmp_path, strand, strict):
    from dynast_b200 import gtf, pipeline
    from oracle import velocity_oracle as vo
    path = str(tmp_path / 'a.gtf')
    contigs = random_gtf(path, seed=3)
    genes, transcripts = gtf.genes_and_transcripts_from_gtf(path)
    header = contigs[::-1] + ['unused']
    tables = gtf.AnnotationTables(genes, transcripts, header)
    als = random_alignments(genes, header, 6000, seed=4 + strict)
    blob, off = _pack(als)
    gx = np.array([tables.gene_index[a['gx']] if a['gx'] else -1 for a in als], dtype=np.int32)
    dev = torch.device('cuda', 0)
    g, t, st = pipeline.velocity(gpu_ctx, torch.from_numpy(blob).to(dev), torch.from_numpy(off.view(np.int32)).to(dev), tables,
                                 gx=torch.from_numpy(gx).to(dev), strand=strand, strict_exon_overlap=strict)
    assert int(st.item()) == 0
    g, t = g.cpu().numpy(), t.cpu().numpy()
    o_genes, o_tr = vo.parse_gtf(path)
    per_contig = vo.by_contig(o_genes)
    order = {c: vo.gene_order(gi) for c, gi in per_contig.items()}
    seen = set()
    for i, a in enumerate(als):
        positions = reference_positions(a['units'][0][0], a['units'][0][1])
        if len(a['units']) == 2:
            positions = positions + reference_positions(a['units'][1][0], a['units'][1][1])
        c = header[a['ref_id']]
        want_g, want_t = vo.assign_velocity_type(positions, per_contig[c], o_tr, order[c], gx=a['gx'],
                                                 strand=vo.read_strand(a['units'][0][2], strand), strict_exon_overlap=strict)
        if not want_t or not want_g:
            want_g, want_t = None, None
        got_g = tables.gene_ids[g[i]] if g[i] >= 0 else None
        assert (got_g, int(t[i])) == (want_g, NAMES[want_t]), (i, a)
        seen.add((want_t, bool(a['gx']), len(a['units'])))
    # every outcome was exercised with and without a gene tag, single and paired
    assert {(k, tag, n) for k in ('spliced', 'unspliced', 'ambiguous') for tag in (True, False) for n in (1, 2)} <= seen
    assert (None, False, 1) in seen


def test_velocity_status_bits(gpu_ctx, tmp_path):
    from dynast_b200 import gtf, pipeline, records
    path = str(tmp_path / 'a.gtf')
    contigs = random_gtf(path, seed=9, n_contigs=2, genes_per_contig=5)
    genes, transcripts = gtf.genes_and_transcripts_from_gtf(path)
    tables = gtf.AnnotationTables(genes, transcripts, contigs)
    dev = torch.device('cuda', 0)
    # 60 aligned blocks: more than the kernel holds
    cigar = []
    for _ in range(60):
        cigar += [(0, 2), (3, 5)]
    cigar.pop()
    rec = records.pack_one(100, 0, 0, cigar, 'A' * 120, [30] * 120, '120')
    blob, off = records.pack_units([[rec]])
    _, t, st = pipeline.velocity(gpu_ctx, torch.from_numpy(blob).to(dev), torch.from_numpy(off.view(np.int32)).to(dev), tables)
    assert int(st.item()) == 1 and int(t[0]) == 0
    # a gene of the other contig
    other = next(i for i, gid in enumerate(tables.gene_ids) if genes[gid]['chromosome'] == contigs[1])
    rec = records.pack_one(100, 0, 0, [(0, 20)], 'A' * 20, [30] * 20, '20')
    blob, off = records.pack_units([[rec]])
    _, _, st = pipeline.velocity(gpu_ctx, torch.from_numpy(blob).to(dev), torch.from_numpy(off.view(np.int32)).to(dev), tables,
                                 gx=torch.tensor([other], dtype=torch.int32, device=dev))
    assert int(st.item()) == 2
