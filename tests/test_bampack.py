"""CPU: the C++ BAM ingest (dyn_bam_pack) produces exactly the packed records and metadata that the pure-Python
reader + packer of the oracle produce for the reference's three fixture BAMs (filter, mate pairing, order)."""
import numpy as np
import pytest

from helpers import ref_path

CASES = [
    ('SRR11683995_align', dict(umi_tag='UB', barcode_tag='CB', velocity=True)),
    ('smartseq_align', dict(umi_tag=None, barcode_tag='RG', velocity=False)),
    ('nasc_align', dict(umi_tag=None, barcode_tag='RG', velocity=False)),
]


@pytest.mark.parametrize('align,kw', CASES)
def test_bam_pack_equals_python_packer(align, kw):
    import build
    build.build_product()
    from oracle import pack
    from dynast_b200 import bamio, records as rec
    path = ref_path(align, 'Aligned.sortedByCoord.out.bam')
    blob, off, meta, refs = pack.units_from_bam(path, rec.pack_one, rec.pack_units, **kw)
    with bamio.PackedBam(path, barcode_tag=kw['barcode_tag'], umi_tag=kw['umi_tag'], gene_tag='GX',
                         require_gene=not kw['velocity'], n_threads=3) as pb:
        assert pb.n_units == len(meta) > 100
        assert np.array_equal(pb.rec_off, off)
        assert np.array_equal(pb.blob, blob)
        assert list(pb.read_id) == [m['read_id'] for m in meta]
        assert list(pb.hi) == [m['index'] for m in meta]
        assert list(pb.barcode) == [m['barcode'] for m in meta]
        assert list(pb.umi) == [m['umi'] for m in meta]
        assert list(pb.gene) == [m['GX'] for m in meta]
        assert list(pb.gx_assigned) == [int(m['gx_assigned']) for m in meta]
        assert list(pb.score) == [m['score'] for m in meta]
        assert list(pb.ref_id) == [m['ref_id'] for m in meta]
        assert list(pb.ref_name) == [r[0] for r in refs]
        assert pb.n_records_seen >= pb.n_units


def test_bam_pack_errors(tmp_path):
    from dynast_b200 import _lib, bamio
    with pytest.raises(_lib.DynastB200Error):
        bamio.PackedBam(str(tmp_path / 'missing.bam'))
    bad = tmp_path / 'bad.bam'
    bad.write_bytes(b'not a bam file at all, just bytes')
    with pytest.raises(_lib.DynastB200Error):
        bamio.PackedBam(str(bad))


@pytest.mark.parametrize('aligned', [True, False])
def test_bam_pack_parallel_paths(tmp_path, aligned):
    """Single-end files of >= 4 096 records take the parallel ingest (records packed by all threads, stitched in order;
    the record chain from per-block walks when BGZF blocks end on record boundaries, else from the walk behind the
    inflating threads): every field must equal the one-thread sequential result, and the oracle's packer."""
    import os
    import subprocess
    import sys
    import build
    build.build_product()
    from oracle import pack
    from dynast_b200 import bamio, records as rec
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path / 'big.bam')
    subprocess.run([sys.executable, os.path.join(root, 'tools', 'make_bam.py'),
                    ref_path('SRR11683995_align', 'Aligned.sortedByCoord.out.bam'), out, '5'] + (['aligned'] if aligned else []),
                   check=True, capture_output=True)
    fields = ('rec_off', 'blob', 'read_id', 'hi', 'barcode', 'umi', 'gene', 'gx_assigned', 'score', 'ref_id', 'pos', 'flag', 'paired')
    got = {}
    for nt in (1, 4):
        with bamio.PackedBam(out, barcode_tag='CB', umi_tag='UB', gene_tag='GX', require_gene=False, n_threads=nt) as pb:
            assert pb.n_records_seen == 1780 * 5 and pb.n_units % 5 == 0 and pb.n_units > 1000
            got[nt] = {f: np.array(getattr(pb, f)) for f in fields}
    for f in fields:
        assert np.array_equal(got[1][f], got[4][f]), f
    blob, off, meta, refs = pack.units_from_bam(out, rec.pack_one, rec.pack_units, umi_tag='UB', barcode_tag='CB', velocity=True)
    assert np.array_equal(got[4]['blob'], blob) and np.array_equal(got[4]['rec_off'], off)
    assert list(got[4]['read_id']) == [m['read_id'] for m in meta] and len(meta) == len(off) - 1


def _crafted_md_bam(path):
    """Three alignments whose MD tags put a mismatch right behind a deletion: with the "0" the SAM specification asks for,
    without it (what dynast's consensus writes, consensus.py:116-176: "10^TA" + "A" -> "10^TAA"), and a tag that really
    disagrees with its CIGAR."""
    from oracle import bamio as obam
    tags = lambda md: [('NH', 'C', 1), ('HI', 'C', 1), ('AS', 'C', 40), ('CB', 'Z', 'ACGTACGTACGTACGT'), ('UB', 'Z', 'ACGTACGTACGT'),
                       ('GX', 'Z', 'ENSG1'), ('MD', 'Z', md)]
    seq = 'ACGTACGTACCCGTACGTACGT'[:21]
    qual = bytes([38] * 21)
    recs = [obam.encode_record('r%d' % i, 0, 0, 1000 + i, 255, [(0, 10), (2, 2), (0, 11)], seq, qual, tags(md))
            for i, md in enumerate(['10^TA0A10', '10^TAA10', '10^T11', '9C0^TA11', '9C0^T0A11'])]
    obam.write_bam(path, [('chr1', 100000)], iter(recs))


def test_mismatch_glued_to_a_deletion_in_md(tmp_path):
    """"10^TAA10" under CIGAR 10M2D11M: two deleted bases, then a mismatch -- the same tokens as the spec-conformant
    "10^TA0A10", in the Python packer and in the C++ one; a run SHORTER than the D operation stays an error."""
    import build
    build.build_product()
    from dynast_b200 import bamio, records as rec
    cig = [(0, 10), (2, 2), (0, 11)]
    good, ok1 = rec.tokenize_md('10^TA0A10', cig)
    glued, ok2 = rec.tokenize_md('10^TAA10', cig)
    assert ok1 and ok2 and good == glued == [10 | (8 << 16) | (1 << 20), 10 | (1 << 16) | (1 << 20), 10 | (1 << 16)]
    assert rec.tokenize_md('10^T11', cig)[1] is False
    # a mismatch BEFORE the deletion: the reference keeps writing "0" in front of every deleted base
    spec, ok3 = rec.tokenize_md('9C0^TA11', cig)
    quirk, ok4 = rec.tokenize_md('9C0^T0A11', cig)
    assert ok3 and ok4 and spec == quirk == [9 | (2 << 16), 10 | (8 << 16) | (1 << 20), 10 | (1 << 16) | (1 << 20)]
    assert rec.tokenize_md('9C0^T3A8', cig)[1] is False
    path = str(tmp_path / 'md.bam')
    _crafted_md_bam(path)
    with bamio.PackedBam(path, barcode_tag='CB', umi_tag='UB', gene_tag='GX', require_gene=True, n_threads=1) as pb:
        off = pb.rec_off.astype(np.int64) * 16
        units = [bytes(pb.blob[off[i]:off[i + 1]]) for i in range(pb.n_units)]
    assert len(units) == 5
    flags = [int.from_bytes(u[14:16], 'little') for u in units]
    assert [bool(f & rec.BAD_MD) for f in flags] == [False, False, True, False, False]
    assert units[0][8:] == units[1][8:] and units[3][8:] == units[4][8:]                 # same records but for the position
