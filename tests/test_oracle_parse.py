"""CPU: the parse/tally oracles are pinned to the reference's fixture outputs (byte equality, as in
tests/preprocessing/test_bam.py:40-152 of the reference)."""
import csv
import gzip
import io
import os
import pickle

import pytest

from helpers import ref_path, ref_rows, ref_text

CASES = [
    ('SRR11683995_align', 'SRR11683995_count', dict(umi_tag='UB', barcode_tag='CB', velocity=True), False),
    ('smartseq_align', 'smartseq_count', dict(umi_tag=None, barcode_tag='RG', velocity=False), False),
    ('nasc_align', 'nasc_count', dict(umi_tag=None, barcode_tag='RG', velocity=False), True),
]


@pytest.mark.parametrize('align,count,kw,nasc', CASES)
def test_parse_oracle_bytes_equal(tmp_path, align, count, kw, nasc):
    from oracle import parse_oracle
    lookup = None
    if kw['velocity']:   # velocity needs the missing transcripts.pkl.gz blob: injected from the fixture
        lookup = {(r['read_id'], int(r['index'])): (r['GX'], r['velocity']) for r in ref_rows(count, 'alignments.csv')}
    c, a = str(tmp_path / 'c.csv'), str(tmp_path / 'a.csv')
    index = parse_oracle.parse_all_reads(ref_path(align, 'Aligned.sortedByCoord.out.bam'), c, a, nasc=nasc,
                                         velocity_lookup=lookup, **kw)
    assert open(c).read() == ref_text(count, 'conversions.csv')
    assert open(a).read() == ref_text(count, 'alignments.csv')
    with gzip.open(ref_path(count, 'conversions.idx'), 'rb') as f:
        assert pickle.load(f) == index


@pytest.mark.parametrize('align,count,kw,nasc', CASES)
def test_tally_oracle_matches_fixture_rows(oracle_lib, align, count, kw, nasc):
    """Packed records -> C oracle == rows of the fixture alignments.csv / conversions.csv."""
    from oracle import pack
    from dynast_b200 import records as rec
    blob, off, meta, refs = pack.units_from_bam(ref_path(align, 'Aligned.sortedByCoord.out.bam'), rec.pack_one,
                                                rec.pack_units, **kw)
    out = pack.run_tally(blob, off, quality=27, nasc=nasc)
    assert out['status'] == 0
    ali = {(r['read_id'], int(r['index'])): r for r in ref_rows(count, 'alignments.csv')}
    convs = {}
    for r in ref_rows(count, 'conversions.csv'):
        convs.setdefault((r['read_id'], int(r['index'])), []).append(
            (int(r['genome_i']), r['conversion'], int(r['quality']), r['contig']))
    pos, refl, readl, q = rec.decode_conv(out['conv_rec'])
    n_cmp = 0
    for i, m in enumerate(meta):
        key = (m['read_id'], m['index'])
        if key not in ali:
            continue   # read dropped by velocity assignment in the fixture run
        n_cmp += 1
        row = ali[key]
        assert [int(row[b]) for b in 'ACGT'] == [int(x) for x in out['counts'][i, 12:16]]
        o, n = int(out['conv_off'][i]), int(out['conv_n'][i])
        got = [(int(pos[j]), refl[j] + readl[j], int(q[j]), m['contig']) for j in range(o, o + n)]
        assert got == convs.get(key, [])
        want = [0] * 12
        for _, c, qq, _ in convs.get(key, []):
            if qq > 27:
                want[rec.CONVERSION_COLUMNS.index(c)] += 1
        assert want == [int(x) for x in out['counts'][i, :12]]
        assert (row['barcode'], row['umi'], int(row['score']), row['transcriptome']) == (
            m['barcode'], m['umi'], m['score'], str(m['gx_assigned']))
    assert n_cmp == len(ali)


def _plain_tally(pos, cigar, seq, qual, md, quality):
    """The reference's per-read rules (bam.py:357-411, conversion.py:417-426) written out on (CIGAR, MD, SEQ)
    text, independent of the packed format and of the C walker."""
    import re
    aligned, q, r = [], 0, pos
    for op, n in cigar:
        if op in (0, 7, 8):
            aligned += [(q + i, r + i) for i in range(n)]
            q += n
            r += n
        elif op in (1, 4):
            q += n
        elif op in (2, 3):
            r += n
    ref_bases, deleted, ai = [], [], 0
    for num, dele, mm in re.findall(r'(\d+)|(\^[A-Z]+)|([A-Z])', md):
        if num:
            for _ in range(int(num)):
                ref_bases.append(seq[aligned[ai][0]])
                ai += 1
        elif dele:
            deleted += list(dele[1:])
        else:
            ref_bases.append(mm)
            ai += 1
    assert ai == len(aligned)
    content = [sum(1 for b in ref_bases + deleted if b == c) for c in 'ACGT']
    names = [a + b for a in 'ACGT' for b in 'ACGT' if a != b]
    counts = [0] * 12
    recs = []
    for (qp, rp), g in zip(aligned, ref_bases):
        b = seq[qp]
        if g == 'N' or b == 'N' or g == b:
            continue
        recs.append((rp, g + b, qual[qp]))
        if qual[qp] > quality:
            counts[names.index(g + b)] += 1
    return counts + content, recs


@pytest.mark.parametrize('read_len,p_e,seed', [(91, 5e-4, 31), (40, 0.03, 33), (17, 0.05, 34)])
def test_tally_oracle_complex_cigars(oracle_lib, read_len, p_e, seed):
    """C oracle == the rules above on reads with clips at both ends, =/X pieces, several I / D / N per read."""
    from oracle import pack
    from dynast_b200 import records as rec, synth
    d = synth.make_reads(400, seed=seed, read_len=read_len, p_e=p_e, frac_n=0.01, complex_cigars=True)
    out = pack.run_tally(d['blob'], d['rec_off'], quality=20)
    assert out['status'] == 0
    pos, refl, readl, q = rec.decode_conv(out['conv_rec'])
    n_ops = set()
    for i, (p0, cigar, seq, qual, md) in enumerate(d['raw']):
        want, recs = _plain_tally(p0, cigar, seq, qual, md, 20)
        assert want == [int(x) for x in out['counts'][i]], (cigar, md)
        o, n = int(out['conv_off'][i]), int(out['conv_n'][i])
        assert [(int(pos[j]), refl[j] + readl[j], int(q[j])) for j in range(o, o + n)] == recs
        n_ops.update(op for op, _ in cigar)
    assert n_ops == {0, 1, 2, 3, 4, 5, 7, 8}
