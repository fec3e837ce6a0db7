"""CPU: the streaming, range-split BAM reader (dyn_bam_stream_*, csrc/bamstream.cpp) returns exactly what the
whole-file packer returns -- chunk by chunk, part by part, for block-aligned and non-aligned files, single-end and
paired -- plus the numeric keys and the lean record layout."""
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import ref_path

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIELDS = ('ref_id', 'pos', 'flag', 'hi', 'score', 'gx_assigned', 'paired')
POOLS = ('read_id', 'barcode', 'umi', 'gene')


@pytest.fixture(scope='module')
def lib():
    import build
    build.build_product()
    from dynast_b200 import _lib
    return _lib.load()


def _big(tmp_path, aligned, copies=5):
    out = str(tmp_path / f'big{int(aligned)}.bam')
    subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'make_bam.py'),
                    ref_path('SRR11683995_align', 'Aligned.sortedByCoord.out.bam'), out, str(copies)] + (['aligned'] if aligned else []),
                   check=True, capture_output=True)
    return out


def _whole(path, **kw):
    from dynast_b200 import bamio
    with bamio.PackedBam(path, n_threads=2, **kw) as pb:
        d = {f: np.array(getattr(pb, f)) for f in FIELDS + POOLS}
        d['records'] = [bytes(pb.blob[16 * int(a):16 * int(b)]) for a, b in zip(pb.rec_off[:-1], pb.rec_off[1:])]
        d['seen'] = pb.n_records_seen
        d['ref_name'] = list(pb.ref_name)
    return d


def _streamed(path, n_parts=1, chunk_bytes=0, pools=True, **kw):
    from dynast_b200 import bamio
    out = {f: [] for f in FIELDS + (POOLS if pools else ())}
    out['records'] = []
    seen = 0
    n_chunks = 0
    for part in range(n_parts):
        with bamio.BamStream(path, n_threads=3, part=part, n_parts=n_parts, chunk_bytes=chunk_bytes, **kw) as st:
            ref_name = list(st.ref_name)
            for chunk in st:
                n_chunks += 1
                seen += chunk.n_records_seen
                for f in FIELDS + (POOLS if pools else ()):
                    out[f].extend(list(getattr(chunk, f)))
                out['records'] += [bytes(chunk.blob[16 * int(a):16 * int(b)]) for a, b in zip(chunk.rec_off[:-1], chunk.rec_off[1:])]
                for extra in ('barcode_key', 'umi_key', 'gene_idx'):
                    out.setdefault(extra, []).extend(list(getattr(chunk, extra)))
                chunk.close()
    out['seen'], out['ref_name'], out['n_chunks'] = seen, ref_name, n_chunks
    return out


def _same(a, b, pools=True):
    assert a['records'] == b['records']
    for f in FIELDS + (POOLS if pools else ()):
        assert list(a[f]) == list(b[f]), f
    assert a['seen'] == b['seen'] and a['ref_name'] == b['ref_name']


@pytest.mark.parametrize('align,kw', [
    ('SRR11683995_align', dict(umi_tag='UB', barcode_tag='CB', gene_tag='GX', require_gene=False)),
    ('smartseq_align', dict(umi_tag=None, barcode_tag='RG', gene_tag='GX', require_gene=True)),
    ('nasc_align', dict(umi_tag=None, barcode_tag='RG', gene_tag='GX', require_gene=True)),
])
@pytest.mark.parametrize('chunk_bytes', [0, 40_000])
def test_stream_equals_whole_file_on_fixtures(lib, align, kw, chunk_bytes):
    """Paired fixtures included: the first-seen-mate table survives chunk boundaries."""
    path = ref_path(align, 'Aligned.sortedByCoord.out.bam')
    whole = _whole(path, **kw)
    got = _streamed(path, chunk_bytes=chunk_bytes, **kw)
    _same(whole, got)
    if chunk_bytes:
        assert got['n_chunks'] > 3


@pytest.mark.parametrize('aligned', [True, False])
@pytest.mark.parametrize('n_parts,chunk_bytes', [(1, 150_000), (3, 0), (7, 100_000)])
def test_stream_parts_and_chunks_cover_the_file_once(lib, tmp_path, aligned, n_parts, chunk_bytes):
    """Block ranges: every record belongs to exactly one part, in file order -- also when records straddle BGZF blocks
    (the part boundary is then found by the record-chain guesser) and when a record straddles two parts."""
    path = _big(tmp_path, aligned)
    kw = dict(umi_tag='UB', barcode_tag='CB', gene_tag='GX', require_gene=False)
    whole = _whole(path, **kw)
    assert whole['seen'] == 1780 * 5
    got = _streamed(path, n_parts=n_parts, chunk_bytes=chunk_bytes, **kw)
    _same(whole, got)


def test_stream_keys_and_lean(lib, tmp_path, oracle_lib):
    from dynast_b200 import bamio
    from oracle import pack
    path = _big(tmp_path, True, copies=2)
    kw = dict(umi_tag='UB', barcode_tag='CB', gene_tag='GX', require_gene=False)
    whole = _whole(path, **kw)
    genes = sorted(set(g for g in whole['gene'] if g))
    dropped = genes.pop()                      # a gene the caller's list does not know
    got = _streamed(path, n_parts=2, chunk_bytes=200_000, pools=False, lean=True, gene_ids=genes, **kw)
    assert [bamio.decode_key(k) for k in got['barcode_key']] == list(whole['barcode'])
    assert [bamio.decode_key(k, suffix=False) for k in got['umi_key']] == list(whole['umi'])
    want_idx = [-1 if g == '' else (-2 if g == dropped else genes.index(g)) for g in whole['gene']]
    assert list(got['gene_idx']) == want_idx
    for f in FIELDS:
        assert list(got[f]) == list(whole[f]), f
    # lean records: same tally as the full records (conversion.py:417-426 needs qualities at mismatches only)
    def tally(records):
        blob = np.frombuffer(b''.join(records), dtype=np.uint8)
        off = np.concatenate(([0], np.cumsum([len(r) // 16 for r in records]))).astype(np.uint32)
        return pack.run_tally(blob, off, quality=27)
    full, lean = tally(whole['records']), tally(got['records'])
    assert full['status'] == 0 and lean['status'] == 0
    assert sum(len(r) for r in got['records']) < 0.62 * sum(len(r) for r in whole['records'])
    for k in ('counts', 'conv_off', 'conv_n', 'conv_rec', 'conv_read'):
        assert np.array_equal(full[k], lean[k]), k
    # and the lean bytes are what the Python packer writes
    from dynast_b200 import records as rec
    blob, off, meta, _ = pack.units_from_bam(path, lambda *a: rec.pack_one(*a, lean=True), rec.pack_units, umi_tag='UB', barcode_tag='CB')
    assert [bytes(blob[16 * int(a):16 * int(b)]) for a, b in zip(off[:-1], off[1:])] == got['records']


def test_key_encoding_round_trip(lib):
    from dynast_b200 import bamio
    assert bamio.decode_key(0) == 'NA'
    assert bamio.decode_key((1 << 8) | 0b00011011) == 'ACGT'
    assert bamio.decode_key(((1 << 4) | 0b1100) | (2 << 58)) == 'TA-1'


def test_paired_bam_cannot_be_range_split(lib):
    from dynast_b200 import _lib, bamio
    path = ref_path('smartseq_align', 'Aligned.sortedByCoord.out.bam')
    with pytest.raises(_lib.DynastB200Error):
        with bamio.BamStream(path, barcode_tag='RG', part=0, n_parts=2, chunk_bytes=30_000) as st:
            for chunk in st:
                chunk.close()
