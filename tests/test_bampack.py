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
