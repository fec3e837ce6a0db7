"""TEST INFRASTRUCTURE ONLY -- synthesise the annotation that pins velocity assignment to the reference fixture.

The reference's velocity test (tests/preprocessing/test_bam.py:40-68) needs ``transcripts.pkl.gz``, a missing blob,
and the full Ensembl GTF the fixtures were made with is not in the repository.  What IS committed upstream:
``genes.pkl.gz`` (gene segments, strands, contigs -> tests/golden/gene_infos.json) and the outcome of the assignment
for every read of the UMI fixture (``SRR11683995_count/alignments.csv``: columns GX, velocity; reads that are
absent were dropped).  This script constructs exon structures that are consistent with every one of those rows
under the rules of bam.py:180-228 and writes them as ``tests/golden/velocity_fixture.gtf``:

  * a `spliced` read gets a transcript whose exons are exactly its aligned blocks (so it is a subset of them and
    the transcript's introns are the read's own N gaps);
  * an `unspliced` read must overlap an intron of some transcript of its gene and be inside no transcript's
    exons: if no existing intron does, a transcript with two 1-base exons around a 1-base intron inside the read
    is added at a base no `ambiguous` read of the gene covers;
  * an `ambiguous` read must overlap an exon, be inside no transcript's exons and overlap no intron: if nothing
    overlaps it yet, a transcript with a 1-base exon at its first base is added;
  * reads without a GX tag are located by position and strand among the REAL gene segments, as in the reference.

The result is then checked by running oracle/parse_oracle.py with oracle/velocity_oracle.py on the BAM and
comparing alignments.csv byte for byte; the script fails if any row (or any dropped read) disagrees.

    python -m oracle.make_velocity_gtf        # rewrites tests/golden/velocity_fixture.gtf
"""
import gzip
import io
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
OUT = os.path.join(GOLDEN, 'velocity_fixture.gtf')


def main():
    sys.path.insert(0, ROOT)
    import csv
    from oracle import bamio, parse_oracle
    from oracle.velocity_oracle import SegmentCollection

    bam = os.path.join(GOLDEN, 'ref', 'SRR11683995_align', 'Aligned.sortedByCoord.out.bam')
    with gzip.open(os.path.join(GOLDEN, 'ref', 'SRR11683995_count', 'alignments.csv.gz'), 'rt') as f:
        want_text = f.read()
    want = {(r['read_id'], int(r['index'])): (r['GX'], r['velocity']) for r in csv.DictReader(io.StringIO(want_text))}
    with open(os.path.join(GOLDEN, 'gene_infos.json')) as f:
        genes = json.load(f)

    _, refs, reads = bamio.read_bam(bam)
    spliced, unspliced, ambiguous = {}, {}, {}
    for r in reads:
        if r.is_secondary or r.is_unmapped or r.is_duplicate or not r.has_tag('UB') or not r.has_tag('CB'):
            continue
        if r.get_tag('CB') == '-':
            continue
        hit = want.get((r.query_name, r.get_tag('HI')))
        if hit is None:
            continue
        gx, kind = hit
        assert genes[gx]['chromosome'] == r.reference_name, (gx, r.reference_name)
        blocks = [(s.start, s.end) for s in SegmentCollection.from_positions(r.get_reference_positions())]
        {'spliced': spliced, 'unspliced': unspliced, 'ambiguous': ambiguous}[kind].setdefault(gx, []).append(blocks)

    transcripts = {}     # gene -> list of exon lists

    def covers(blocks, p):
        return any(a <= p < b for a, b in blocks)

    for gx, rs in spliced.items():
        seen = set()
        for blocks in rs:
            if tuple(blocks) not in seen:
                seen.add(tuple(blocks))
                transcripts.setdefault(gx, []).append(list(blocks))

    def introns_of(exons):
        return [(exons[i][1], exons[i + 1][0]) for i in range(len(exons) - 1) if exons[i][1] < exons[i + 1][0]]

    def overlaps(blocks, segs):
        return any(a < d and b > c for a, b in blocks for c, d in segs)

    def subset(blocks, segs):
        return all(any(a >= c and b <= d for c, d in segs) for a, b in blocks)

    for gx, rs in unspliced.items():
        for blocks in rs:
            ts = transcripts.setdefault(gx, [])
            assert not any(subset(blocks, t) for t in ts), ('unspliced read inside exons', gx, blocks)
            if any(overlaps(blocks, introns_of(t)) for t in ts):
                continue
            amb = ambiguous.get(gx, [])
            for a, b in blocks:
                for p in range(a + 1, b - 1):
                    if not any(covers(x, p) for x in amb):
                        ts.append([(p - 1, p), (p + 1, p + 2)])
                        break
                else:
                    continue
                break
            else:
                raise SystemExit(f'no free base for an intron in {gx} {blocks}')
    for gx, rs in ambiguous.items():
        for blocks in rs:
            ts = transcripts.setdefault(gx, [])
            assert not any(subset(blocks, t) for t in ts), ('ambiguous read inside exons', gx, blocks)
            assert not any(overlaps(blocks, introns_of(t)) for t in ts), ('ambiguous read over an intron', gx, blocks)
            if not any(overlaps(blocks, t) for t in ts):
                ts.append([(blocks[0][0], blocks[0][0] + 1)])

    lines = ['#!genome-build synthetic: exon structure constructed by oracle/make_velocity_gtf.py from genes.pkl.gz and '
             'SRR11683995_count/alignments.csv']
    order = sorted(genes, key=lambda g: (genes[g]['chromosome'], genes[g]['start'], genes[g]['end'], g))
    for gx in order:
        gi = genes[gx]
        attr = f'gene_id "{gx}"; gene_name "{gi.get("gene_name", "")}";'
        lines.append(f'{gi["chromosome"]}\tsynthetic\tgene\t{gi["start"] + 1}\t{gi["end"]}\t.\t{gi["strand"]}\t.\t{attr}')
        for ti, exons in enumerate(transcripts.get(gx, [])):
            tid = f'{gx}.T{ti + 1}'
            tattr = f'{attr} transcript_id "{tid}"; transcript_name "{tid}";'
            lines.append(f'{gi["chromosome"]}\tsynthetic\ttranscript\t{exons[0][0] + 1}\t{exons[-1][1]}\t.\t{gi["strand"]}\t.\t{tattr}')
            for a, b in exons:
                lines.append(f'{gi["chromosome"]}\tsynthetic\texon\t{a + 1}\t{b}\t.\t{gi["strand"]}\t.\t{tattr}')
    text = '\n'.join(lines) + '\n'

    # ---- check: the oracle must reproduce the fixture from the BAM + this annotation
    from oracle import velocity_oracle
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, 'a.gtf')
        with open(path, 'w') as f:
            f.write(text)
        annotation = velocity_oracle.parse_gtf(path)
        c, a = os.path.join(tmp, 'c.csv'), os.path.join(tmp, 'a.csv')
        parse_oracle.parse_all_reads(bam, c, a, umi_tag='UB', barcode_tag='CB', velocity=True, annotation=annotation)
        got_text = open(a).read()
    if got_text != want_text:
        got = {(r['read_id'], int(r['index'])): (r['GX'], r['velocity']) for r in csv.DictReader(io.StringIO(got_text))}
        for k in sorted(set(got) | set(want)):
            if got.get(k) != want.get(k):
                print('MISMATCH', k, 'got', got.get(k), 'want', want.get(k))
        raise SystemExit('the synthesised annotation does not reproduce alignments.csv')
    with open(OUT, 'w') as f:
        f.write(text)
    n_t = sum(len(v) for v in transcripts.values())
    print(f'{OUT}: {len(order)} genes, {n_t} transcripts; alignments.csv reproduced byte for byte ({len(want)} rows)')


if __name__ == '__main__':
    main()
