"""CPU host logic (device stage replaced by the oracle) and GPU: preprocessing.call_consensus / consensus() -- the whole consensus caller (dynast/preprocessing/consensus.py:227-494)
against its line-by-line restatement (oracle/consensus_oracle.py::call_consensus_bam): which reads are grouped, passed
through or gene-tag-swapped, the periodic-flush quirk, consensus records (sequence, CIGAR, MD / NM, tags, names, strand),
the coordinate order and the index of the BAM that comes out; then count() on that BAM picks dedup mode `exon` by
itself (count.py:295-304)."""
import os

import numpy as np
import pytest

from helpers import gene_infos, ref_path



def _canon_read(r, tags):
    return ('read', r.query_name, r.flag, r.reference_id, r.reference_start, tuple(r.cigar), r.query_sequence, tuple(r.query_qualities),
            tuple(sorted((k, str(v)) for k, v in tags.items())))


def _canon_expected(o):
    from oracle import bamio
    if o['kind'] == 'read':
        return _canon_read(o['read'], o['tags'])
    cigar = tuple((bamio.CIGAR_OPS.index(op), n) for op, n in o['cigar'])
    return ('cons', o['name'], 16 if o['reverse'] else 0, o['reference_id'], o['start'], cigar, o['seq'], tuple(o['qual']),
            tuple(sorted((k, str(v)) for k, v in o['tags'].items())))


def _canon_written(r):
    if 'RN' in r.tags:
        return ('cons', r.query_name, r.flag, r.reference_id, r.reference_start, tuple(r.cigar), r.query_sequence, tuple(r.query_qualities),
                tuple(sorted((k, str(v)) for k, v in r.tags.items())))
    return _canon_read(r, r.tags)


def _compare(bam_in, out, kw, infos):
    from oracle import bamio, consensus_oracle
    _, refs, reads = bamio.read_bam(bam_in)
    want = [_canon_expected(o) for o in consensus_oracle.call_consensus_bam(reads, infos, **kw)]
    text, refs2, got_reads = bamio.read_bam(out)
    got = [_canon_written(r) for r in got_reads]
    assert refs2 == refs and 'SO:coordinate' in text
    assert sorted(got) == sorted(want)
    keys = [(r.reference_id, r.reference_start) for r in got_reads]
    assert keys == sorted(keys)                      # coordinate sorted
    assert os.path.getsize(out + '.bai') > 8
    return want, got_reads


@pytest.mark.parametrize('kw', [
    dict(umi_tag='UB', barcode_tag='CB', gene_tag='GX', strand='forward'),
    dict(umi_tag=None, barcode_tag=None, gene_tag='GX', strand='forward'),        # grouped by gene alone: deep groups
    dict(umi_tag=None, barcode_tag='CB', gene_tag='GX', strand='reverse', quality=35),
    dict(umi_tag='UB', barcode_tag='CB', gene_tag=None, strand='forward'),         # genes by position only
])
def test_call_consensus_matches_oracle_on_umi_fixture(backend, tmp_path, kw):
    from dynast_b200.preprocessing import call_consensus
    path = ref_path('SRR11683995_align', 'Aligned.sortedByCoord.out.bam')
    out = str(tmp_path / 'consensus.bam')
    infos = gene_infos()
    assert call_consensus(path, out, infos, n_threads=2, **kw) == out
    want, _ = _compare(path, out, kw, infos)
    kinds = {w[0] for w in want}
    assert 'read' in kinds


def test_call_consensus_paired_pseudo_umi(backend, tmp_path):
    """Smart-seq: no UMI tag, pairs grouped by (mate start, read end) (consensus.py:392-395); members = read, then mate."""
    from dynast_b200.preprocessing import call_consensus
    path = ref_path('smartseq_align', 'Aligned.sortedByCoord.out.bam')
    out = str(tmp_path / 'consensus.bam')
    kw = dict(umi_tag=None, barcode_tag='RG', gene_tag='GX', strand='unstranded')
    call_consensus(path, out, gene_infos(), n_threads=2, **kw)
    want, _ = _compare(path, out, kw, gene_infos())
    assert sum(w[0] == 'cons' for w in want) > 10


def _many_singletons(tmp_path, copies=40):
    """The UMI fixture's gene-tagged reads `copies` times, every copy with its own read name and UMI: > 10 000 records
    whose groups stay singletons, so that the flush at record 10 000 (consensus.py:432-452) has singletons to duplicate."""
    from oracle import bamio
    _, refs, reads = bamio.read_bam(ref_path('SRR11683995_align', 'Aligned.sortedByCoord.out.bam'))
    recs = []
    for r in reads:
        if r.reference_id < 0 or not (r.has_tag('CB') and r.has_tag('UB') and r.has_tag('GX')) or r.get_tag('CB') == '-':
            continue
        for k in range(copies):
            tags = []
            for t, v in r.tags.items():
                if t == 'UB':
                    v = v[:-3] + 'ACGT'[k % 4] + 'ACGT'[(k // 4) % 4] + 'ACGT'[(k // 16) % 4]
                tags.append((t, 'Z' if isinstance(v, str) else ('i' if isinstance(v, int) else 'A'), v))
            recs.append(bamio.encode_record(f'{r.query_name}_{k}', r.flag, r.reference_id, r.reference_start, r.mapping_quality, r.cigar,
                                            r.query_sequence, bytes(r.query_qualities), tags))
    out = str(tmp_path / 'many.bam')
    bamio.write_bam(out, refs, recs)
    return out


def test_call_consensus_flush_quirk(backend, tmp_path):
    from dynast_b200.preprocessing import call_consensus
    from oracle import bamio, consensus_oracle
    kw = dict(umi_tag='UB', barcode_tag='CB', gene_tag='GX', strand='forward')
    path = _many_singletons(tmp_path)          # record 10 000 is a grouped read: the flush of consensus.py:432 runs
    out = str(tmp_path / 'consensus.bam')
    call_consensus(path, out, gene_infos(), n_threads=3, **kw)
    want, got = _compare(path, out, kw, gene_infos())
    one_read_consensus = [w for w in want if w[0] == 'cons' and dict(w[8])['RN'] == '1']
    assert len(one_read_consensus) > 0          # the reference writes these singletons twice (no `continue` at :444)


@pytest.mark.gpu
def test_consensus_command_then_count_detects_consensus_bam(gpu_ctx, tmp_path):
    from dynast_b200 import bamio, consensus as consensus_cmd, count
    path = ref_path('SRR11683995_align', 'Aligned.sortedByCoord.out.bam')
    out_dir = str(tmp_path / 'cons')
    with pytest.warns(None) if False else _no_ctx():
        bam = consensus_cmd.consensus(path, gene_infos(), out_dir, umi_tag=None, barcode_tag='CB', gene_tag='GX')
    assert bam == os.path.join(out_dir, 'consensus.bam') and any(n.startswith('run_info_') for n in os.listdir(out_dir))
    tags, flags, n = bamio.peek(bam)
    assert 'RN' in tags and {'MD', 'NM', 'AS', 'HI', 'CB'} <= tags
    # count() on the consensus BAM: RN tag -> dedup mode `exon` (no UMI tag was used above, so give the barcode as UMI-less run)
    res_dir = str(tmp_path / 'count')
    count.count(bam, gene_infos(), res_dir, barcode_tag='CB', umi_tag=None, gene_tag='GX', velocity=False, n_threads=2)
    assert os.path.exists(os.path.join(res_dir, 'counts_TC.csv'))
    with pytest.raises(Exception, match='required tags'):
        consensus_cmd.consensus(path, gene_infos(), str(tmp_path / 'x'), umi_tag='ZZ', barcode_tag='CB')


class _no_ctx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
