"""GPU: the device ingest (csrc/bamdevice.cu) -- inflate_kernel against zlib, and whole chunks (inflate -> record index
-> parse -> pack, all on the GPU) against the host reader (dyn_bam_stream with keys) on the same files."""
import os
import struct
import subprocess
import sys
import zlib

import numpy as np
import pytest
import torch

from helpers import ref_path

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _inflate_on_gpu(ctx, payloads):
    """payloads: list of (raw deflate bytes, inflated size) -> (list of inflated bytes, status)"""
    from dynast_b200._lib import check, ptr
    desc = np.zeros((len(payloads), 4), dtype=np.uint32)
    in_off = out_off = 0
    for i, (comp, n) in enumerate(payloads):
        desc[i] = (in_off, len(comp), out_off, n)
        in_off += (len(comp) + 15) & ~15
        out_off += n
    src = np.zeros(in_off + 16, dtype=np.uint8)
    for (comp, _), d in zip(payloads, desc):
        src[d[0]:d[0] + len(comp)] = np.frombuffer(comp, dtype=np.uint8)
    dev = torch.device('cuda', 0)
    d_in, d_desc = torch.from_numpy(src).to(dev), torch.from_numpy(desc.view(np.int32)).to(dev)
    d_out = torch.zeros(out_off + 16, dtype=torch.uint8, device=dev)
    scratch = torch.zeros(2, dtype=torch.int32, device=dev)
    check(ctx.lib.dyn_inflate_blocks(ctx.handle, ptr(d_in), ptr(d_out), ptr(d_desc), len(payloads), ptr(scratch),
                                     C_void(scratch.data_ptr() + 4), None))
    torch.cuda.synchronize()
    out = d_out.cpu().numpy()
    return [bytes(out[d[2]:d[2] + d[3]]) for d in desc], int(scratch[1].item())


def C_void(p):
    import ctypes
    return ctypes.c_void_p(p)


def _raw_deflate(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY):
    c = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
    return c.compress(data) + c.flush()


def test_inflate_kernel_matches_zlib(gpu_ctx):
    from test_bamdevice_host import bgzf_payloads
    payloads = []
    for align in ('SRR11683995_align', 'smartseq_align', 'nasc_align'):
        payloads += list(bgzf_payloads(ref_path(align, 'Aligned.sortedByCoord.out.bam')))
    rng = np.random.default_rng(0)
    texts = [b'', b'a', b'abc' * 20000, bytes(rng.integers(0, 256, 60000, dtype=np.uint8)),
             bytes(rng.integers(65, 69, 65000, dtype=np.uint8)), (b'ACGT' * 7 + b'\x00\x01\x02' * 5) * 700,
             bytes(rng.choice(np.arange(256, dtype=np.uint8), 65280, p=np.r_[[0.5], np.full(255, 0.5 / 255)])), b'x' * 65280]
    for data in texts:
        for level in (0, 1, 6, 9):
            for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE):
                payloads.append((_raw_deflate(data, level, strategy), len(data)))
    assert len(payloads) > 200       # more blocks than one wave of warps on a small grid
    got, status = _inflate_on_gpu(gpu_ctx, payloads)
    assert status == 0
    for (comp, n), g in zip(payloads, got):
        assert g == zlib.decompress(comp, -15)


def test_inflate_kernel_reports_corrupt_blocks(gpu_ctx):
    data = bytes(np.random.default_rng(1).integers(65, 70, 30000, dtype=np.uint8))
    comp = _raw_deflate(data)
    good = (comp, len(data))
    for bad in ((comp, len(data) - 10), (comp[:len(comp) // 2], len(data))):
        got, status = _inflate_on_gpu(gpu_ctx, [good, bad, good])
        assert status != 0
        assert got[0] == data and got[2] == data            # the neighbours are not disturbed


def _host_chunks(path, lean, gene_ids, n_parts=1, **kw):
    from dynast_b200 import bamio
    out = dict(records=[], barcode_key=[], umi_key=[], gene_idx=[], score=[], ref_id=[], pos=[], hi=[], flag=[], gx_assigned=[])
    seen = 0
    for part in range(n_parts):
        with bamio.BamStream(path, keys=True, lean=lean, gene_ids=gene_ids, part=part, n_parts=n_parts, n_threads=4, **kw) as st:
            for ch in st:
                seen += ch.n_records_seen
                out['records'] += [bytes(ch.blob[16 * int(a):16 * int(b)]) for a, b in zip(ch.rec_off[:-1], ch.rec_off[1:])]
                for k in out:
                    if k != 'records':
                        out[k] += [int(x) for x in getattr(ch, k)]
                ch.close()
    return out, seen


def _device_chunks(ctx, path, lean, gene_ids, n_parts=1, chunk_bytes=0, **kw):
    from dynast_b200 import bamio
    out = dict(records=[], barcode_key=[], umi_key=[], gene_idx=[], score=[], ref_id=[], pos=[], hi=[], flag=[], gx_assigned=[])
    seen = n_chunks = 0
    for part in range(n_parts):
        with bamio.DeviceBamStream(ctx, path, lean=lean, gene_ids=gene_ids, part=part, n_parts=n_parts, chunk_bytes=chunk_bytes, **kw) as st:
            for ch in st:
                n_chunks += 1
                seen += ch['n_records_seen']
                if not ch['n_units']:
                    continue
                blob = ch['records'].cpu().numpy()
                off = ch['rec_off'].cpu().numpy().view(np.uint32)
                out['records'] += [bytes(blob[16 * int(a):16 * int(b)]) for a, b in zip(off[:-1], off[1:])]
                out['barcode_key'] += [int(x) for x in ch['barcode_key'].cpu().numpy().view(np.uint64)]
                out['umi_key'] += [int(x) for x in ch['umi_key'].cpu().numpy().view(np.uint64)]
                out['flag'] += [int(x) for x in ch['flag'].cpu().numpy().view(np.uint16)]
                for k in ('gene_idx', 'score', 'ref_id', 'pos', 'hi', 'gx_assigned'):
                    out[k] += [int(x) for x in ch[k].cpu().numpy()]
    return out, seen, n_chunks


def _genes(path):
    from dynast_b200 import bamio
    with bamio.PackedBam(path, barcode_tag='CB', umi_tag='UB', n_threads=2) as pb:
        g = sorted(set(x for x in pb.gene if x))
    return g[:-1]          # one gene the list does not know (-> index -2)


@pytest.mark.parametrize('lean', [True, False])
def test_device_ingest_equals_host_reader_on_fixture(gpu_ctx, lean):
    path = ref_path('SRR11683995_align', 'Aligned.sortedByCoord.out.bam')
    genes = _genes(path)
    kw = dict(barcode_tag='CB', umi_tag='UB', gene_tag='GX', require_gene=False)
    want, seen = _host_chunks(path, lean, genes, **kw)
    got, seen_d, _ = _device_chunks(gpu_ctx, path, lean, genes, **kw)
    assert seen_d == seen == 1780
    assert -2 in want['gene_idx'] and -1 in want['gene_idx']
    for k in want:
        assert got[k] == want[k], k


@pytest.mark.parametrize('blocks', ['aligned', 'straddling'])
@pytest.mark.parametrize('n_parts,chunk_bytes', [(1, 0), (1, 300_000), (3, 200_000), (5, 70_000)])
def test_device_ingest_chunks_and_parts(gpu_ctx, tmp_path, n_parts, chunk_bytes, blocks):
    """Chunks and range parts of a file whose BGZF blocks end on record boundaries (htslib) and of one cut every 64 KB
    wherever that falls (STAR's sorted BAMs): records that straddle blocks, chunks and parts are taken once, by the part /
    chunk they start in."""
    out = str(tmp_path / 'big.bam')
    subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'make_bam.py'), ref_path('SRR11683995_align', 'Aligned.sortedByCoord.out.bam'),
                    out, '40'] + (['aligned'] if blocks == 'aligned' else []), check=True, capture_output=True)
    genes = _genes(out)
    kw = dict(barcode_tag='CB', umi_tag='UB', gene_tag='GX', require_gene=True)
    want, seen = _host_chunks(out, True, genes, **kw)
    got, seen_d, n_chunks = _device_chunks(gpu_ctx, out, True, genes, n_parts=n_parts, chunk_bytes=chunk_bytes, **kw)
    assert seen_d == seen == 1780 * 40
    if chunk_bytes:
        assert n_chunks > 2 * n_parts
    for k in want:
        assert got[k] == want[k], k


def test_device_ingest_refuses_what_it_cannot_take(gpu_ctx, tmp_path):
    from dynast_b200 import _lib, bamio
    with pytest.raises(_lib.DynastB200Error, match='paired'):
        with bamio.DeviceBamStream(gpu_ctx, ref_path('smartseq_align', 'Aligned.sortedByCoord.out.bam'), barcode_tag='RG') as st:
            list(st)
    # a file cut off in the middle of a record
    out = str(tmp_path / 'cut.bam')
    subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'make_bam.py'), ref_path('SRR11683995_align', 'Aligned.sortedByCoord.out.bam'),
                    out, '5'], check=True, capture_output=True)
    raw = open(out, 'rb').read()
    o, starts = 0, []
    while o < len(raw):
        starts.append(o)
        o += int.from_bytes(raw[o + 16:o + 18], 'little') + 1

    def left_over(n_blocks):            # bytes behind the last whole record when only the first n_blocks are kept
        data = b''.join(zlib.decompress(raw[a + 12 + int.from_bytes(raw[a + 10:a + 12], 'little'):b - 8], -15) for a, b in zip(starts[:n_blocks], starts[1:]))
        p = 8 + struct.unpack_from('<i', data, 4)[0]
        n_ref = struct.unpack_from('<i', data, p)[0]
        p += 4
        for _ in range(n_ref):
            p += 8 + struct.unpack_from('<i', data, p)[0]
        while p + 4 <= len(data) and p + 4 + struct.unpack_from('<i', data, p)[0] <= len(data):
            p += 4 + struct.unpack_from('<i', data, p)[0]
        return len(data) - p

    keep = next(k for k in range(len(starts) // 2, len(starts) - 1) if left_over(k) > 0)
    open(out, 'wb').write(raw[:starts[keep]])                       # whole blocks, but the last record is incomplete
    with pytest.raises(_lib.DynastB200Error, match='cut off|straddle'):
        with bamio.DeviceBamStream(gpu_ctx, out, barcode_tag='CB', umi_tag='UB') as st:
            list(st)


@pytest.mark.parametrize('lean', [True, False])
def test_device_packer_reads_a_mismatch_glued_to_a_deletion(gpu_ctx, tmp_path, lean):
    """dynast's consensus writes "10^TAA10" for two deleted bases followed by a mismatch (no "0" in between): the device
    tokenizer takes the D operation's length from the CIGAR, like the host packers (tests/test_bampack.py)."""
    from test_bampack import _crafted_md_bam
    path = str(tmp_path / 'md.bam')
    _crafted_md_bam(path)
    kw = dict(barcode_tag='CB', umi_tag='UB', gene_tag='GX', require_gene=True)
    want, seen = _host_chunks(path, lean, ['ENSG1'], **kw)
    got, seen_d, _ = _device_chunks(gpu_ctx, path, lean, ['ENSG1'], **kw)
    assert seen == seen_d == 5
    for k in want:
        assert got[k] == want[k], k
    flags = [int.from_bytes(r[14:16], 'little') for r in got['records']]
    assert [bool(f & 0x4000) for f in flags] == [False, False, True, False, False]
