"""CPU: the BAM writer (dyn_bam_write, csrc/bamwrite.cpp): packed records -> BAM + .bai; read back with the oracle's
pure-Python reader (pysam semantics) and with the streaming ingest."""
import struct

import numpy as np
import pytest

from helpers import ref_path


@pytest.fixture(scope='module')
def lib():
    import build
    build.build_product()
    from dynast_b200 import _lib
    return _lib.load()


def _tags(meta):
    out = b''
    out += b'NHC' + bytes([1]) + b'HIC' + bytes([meta['index']]) + b'ASC' + bytes([meta['score']])
    out += b'CBZ' + meta['barcode'].encode() + b'\0' + b'UBZ' + meta['umi'].encode() + b'\0'
    if meta['gx_assigned']:
        out += b'GXZ' + meta['GX'].encode() + b'\0'
    return out


@pytest.mark.parametrize('shuffle', [False, True])
def test_written_bam_reads_back(lib, tmp_path, shuffle):
    from oracle import bamio as obam, pack
    from dynast_b200 import bamio, records as rec
    src = ref_path('SRR11683995_align', 'Aligned.sortedByCoord.out.bam')
    blob, off, meta, refs = pack.units_from_bam(src, rec.pack_one, rec.pack_units, umi_tag='UB', barcode_tag='CB')
    n = len(meta)
    _, _, reads = obam.read_bam(src)
    by_name = {(r.query_name, r.get_tag('HI')): r for r in reads if not r.is_secondary}
    order = None
    seq_order = list(range(n))
    if shuffle:      # written through a permutation: reversed file order
        order = np.arange(n - 1, -1, -1)
        seq_order = list(order)
    out = str(tmp_path / 'w.bam')
    bamio.write_bam(out, refs, blob, off, order=order, names=[m['read_id'] for m in meta], aux=[_tags(m) for m in meta],
                    mapq=np.full(n, 255, np.uint8), level=6, n_threads=3, index=not shuffle)
    text, refs2, got = obam.read_bam(out)
    assert refs2 == refs and len(got) == n and '@SQ' in text
    for j, u in enumerate(seq_order):
        want, g = by_name[(meta[u]['read_id'], meta[u]['index'])], got[j]
        assert (g.query_name, g.reference_id, g.reference_start, g.cigar, g.query_sequence, g.query_qualities, g.flag & 0xfff) == (
            want.query_name, want.reference_id, want.reference_start, want.cigar, want.query_sequence, want.query_qualities,
            want.flag & 0xfff)
        assert g.get_tag('MD') == want.get_tag('MD') and g.get_tag('CB') == want.get_tag('CB')
        assert g.get_reference_sequence() == want.get_reference_sequence()
    if shuffle:
        return
    # the ingest reads the written file back into the very same packed records, whole and range-split
    with bamio.PackedBam(out, barcode_tag='CB', umi_tag='UB', n_threads=2) as pb:
        assert np.array_equal(pb.blob, blob) and np.array_equal(pb.rec_off, off)
    parts = []
    for part in range(3):
        with bamio.BamStream(out, barcode_tag='CB', umi_tag='UB', part=part, n_parts=3, n_threads=2) as st:
            for chunk in st:
                parts.append(bytes(chunk.blob))
                chunk.close()
    assert b''.join(parts) == bytes(blob)
    # ---- .bai: magic, one entry per reference, chunks and linear offsets point at record starts of that reference
    data = open(out + '.bai', 'rb').read()
    assert data[:4] == b'BAI\x01'
    n_ref = struct.unpack_from('<i', data, 4)[0]
    assert n_ref == len(refs)
    p = 8
    import gzip, zlib
    raw = open(out, 'rb').read()

    def record_at(voff):
        block, within = voff >> 16, voff & 0xffff
        xlen = struct.unpack_from('<H', raw, block + 10)[0]
        bsize = struct.unpack_from('<H', raw, block + 16)[0] + 1
        payload = zlib.decompress(raw[block + 12 + xlen:block + bsize - 8], -15)
        bs, ref_id, pos = struct.unpack_from('<iii', payload, within)
        return ref_id, pos

    seen_refs = set()
    for r in range(n_ref):
        n_bin = struct.unpack_from('<i', data, p)[0]
        p += 4
        for _ in range(n_bin):
            b, n_chunk = struct.unpack_from('<Ii', data, p)
            p += 8
            for c in range(n_chunk):
                beg, end = struct.unpack_from('<QQ', data, p)
                p += 16
                if b != 37450:
                    assert record_at(beg)[0] == r and end > beg
                    seen_refs.add(r)
        n_intv = struct.unpack_from('<i', data, p)[0]
        p += 4
        for w in range(n_intv):
            v = struct.unpack_from('<Q', data, p)[0]
            p += 8
            if v:
                ref_id, pos = record_at(v)
                assert ref_id == r and pos < (w + 1) << 14
    assert struct.unpack_from('<Q', data, p)[0] == 0 and p + 8 == len(data)
    assert seen_refs == {m['ref_id'] for m in meta}


def test_writer_rejects_lean_records(lib, tmp_path):
    from dynast_b200 import _lib, bamio, records as rec
    a = rec.pack_one(100, 0, 0, [(0, 20)], 'A' * 20, [30] * 20, '20', lean=True)
    blob, off = rec.pack_units([[a]])
    with pytest.raises(_lib.DynastB200Error):
        bamio.write_bam(str(tmp_path / 'x.bam'), [('c', 1000)], blob, off)
