"""CPU: the device ingest's cores compiled for the host (tests/native/ingest_host.cpp): the DEFLATE decoder of
csrc/inflate_core.cuh against zlib, and the per-record parse / pack of csrc/bamrec_core.cuh against the host packer
(dyn_bam_stream with keys).  The same source runs in inflate_kernel / parse_kernel / pack_kernel (csrc/bamdevice.cu);
tests/test_gpu_bamdevice.py checks those on the GPU."""
import ctypes as C
import gzip
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

from helpers import ref_path

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def shim(tmp_path_factory):
    out = str(tmp_path_factory.mktemp('native') / 'ingest_host.so')
    subprocess.run(['g++', '-O1', '-std=c++17'] + os.environ.get('DYN_TEST_CXXFLAGS', '').split() + [ '-shared', '-fPIC', '-o', out, os.path.join(ROOT, 'tests', 'native', 'ingest_host.cpp')],
                   check=True)
    lib = C.CDLL(out)
    lib.host_inflate.restype = C.c_int
    lib.host_inflate_lane.restype = C.c_int
    lib.host_parse_pack.restype = C.c_long
    lib.host_fnv1a.restype = C.c_uint64
    lib.host_fnv1a.argtypes = [C.c_char_p]
    return lib


DECODERS = ['host_inflate_lane', 'host_inflate']          # one lane per block (the kernel in use), one warp per block


def _inflate(shim, comp: bytes, n_out: int, shift: int = 0, fn: str = 'host_inflate_lane'):
    out = np.full(n_out + 64, 0x5a, dtype=np.uint8)
    src = np.full(len(comp) + 48, 0xa5, dtype=np.uint8)     # the bit readers load aligned words: padded input (not zeros:
    src[shift:shift + len(comp)] = np.frombuffer(comp, dtype=np.uint8)      # nothing behind the payload may matter)
    rc = getattr(shim, fn)(C.c_void_p(src.ctypes.data + shift), C.c_uint32(len(comp)), C.c_void_p(out.ctypes.data), C.c_uint32(n_out))
    assert bool((out[n_out:] == 0x5a).all()), 'wrote behind the output block'
    return rc, bytes(out[:n_out])


def _raw_deflate(data: bytes, level=6, strategy=zlib.Z_DEFAULT_STRATEGY):
    c = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
    return c.compress(data) + c.flush()


def bgzf_payloads(path):
    raw = open(path, 'rb').read()
    o = 0
    while o < len(raw):
        xlen = struct.unpack_from('<H', raw, o + 10)[0]
        bsize = struct.unpack_from('<H', raw, o + 16)[0] + 1
        isize = struct.unpack_from('<I', raw, o + bsize - 4)[0]
        yield raw[o + 12 + xlen:o + bsize - 8], isize
        o += bsize


@pytest.mark.parametrize('fn', DECODERS)
@pytest.mark.parametrize('align', ['SRR11683995_align', 'smartseq_align', 'nasc_align'])
def test_inflate_core_on_bgzf_blocks(shim, align, fn):
    n = 0
    for comp, isize in bgzf_payloads(ref_path(align, 'Aligned.sortedByCoord.out.bam')):
        rc, got = _inflate(shim, comp, isize, shift=n % 8, fn=fn)          # payloads start at any byte alignment
        assert rc == 0
        assert got == zlib.decompress(comp, -15)
        n += 1
    assert n > 3


def _fibonacci_text(rng, n):
    """Byte frequencies that fall like Fibonacci numbers: the Huffman code of such a text is as deep as it gets."""
    w = np.array([1, 1] + [0] * 30, dtype=np.float64)
    for i in range(2, 32):
        w[i] = w[i - 1] + w[i - 2]
    return bytes(rng.choice(np.arange(32, dtype=np.uint8), n, p=w / w.sum()))


@pytest.mark.parametrize('fn', DECODERS)
def test_inflate_core_stream_kinds(shim, fn):
    rng = np.random.default_rng(0)
    texts = [
        b'',
        b'a',
        b'abc' * 20000,                                         # long matches, overlapping copies (distance 3 < length)
        bytes(rng.integers(0, 256, 60000, dtype=np.uint8)),     # incompressible: stored blocks inside a zlib stream
        bytes(rng.integers(65, 69, 65000, dtype=np.uint8)),     # 2 bits of entropy per byte: short codes
        (b'ACGT' * 7 + b'\x00\x01\x02' * 5) * 700,
        bytes(rng.choice(np.arange(256, dtype=np.uint8), 65280, p=np.r_[[0.5], np.full(255, 0.5 / 255)])),   # long codes (> 10 bits)
        b'x' * 65280,                                            # distance 1 runs
        _fibonacci_text(rng, 65000),                            # code lengths up to zlib's limit of 15 bits
        bytes(rng.integers(0, 256, 30000, dtype=np.uint8)) * 2 + b'tail',   # a 30 000-byte match distance, 258-byte matches
    ]
    for data in texts:
        for level in (0, 1, 6, 9):
            for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE):
                comp = _raw_deflate(data, level, strategy)
                for shift in (0, 3):
                    rc, got = _inflate(shim, comp, len(data), shift=shift, fn=fn)
                    assert rc == 0 and got == data, (len(data), level, strategy, shift)


@pytest.mark.parametrize('fn', DECODERS)
def test_inflate_core_rejects_corrupt_input(shim, fn):
    data = bytes(np.random.default_rng(1).integers(65, 70, 30000, dtype=np.uint8))
    comp = _raw_deflate(data)
    assert _inflate(shim, comp, len(data) - 10, fn=fn)[0] != 0          # output larger than announced
    assert _inflate(shim, comp, len(data) + 10, fn=fn)[0] != 0          # ... and smaller
    assert _inflate(shim, comp[:len(comp) // 2], len(data), fn=fn)[0] != 0      # truncated input
    bad = bytearray(comp)
    bad[0] |= 0x06                                                # block type 3
    assert _inflate(shim, bytes(bad), len(data), fn=fn)[0] != 0
    # random garbage must end (with an error or by chance without), never hang or write out of bounds
    rng = np.random.default_rng(5)
    for _ in range(200):
        junk = bytes(rng.integers(0, 256, int(rng.integers(1, 400)), dtype=np.uint8))
        _inflate(shim, junk, int(rng.integers(0, 3000)), fn=fn)


@pytest.mark.parametrize('lean', [1, 0])
def test_record_cores_equal_host_packer(shim, lean):
    import build
    build.build_product()
    from dynast_b200 import bamio
    path = ref_path('SRR11683995_align', 'Aligned.sortedByCoord.out.bam')
    with gzip.open(path, 'rb') as f:
        data = f.read()
    l_text = struct.unpack_from('<i', data, 4)[0]
    o = 8 + l_text
    n_ref = struct.unpack_from('<i', data, o)[0]
    o += 4
    for _ in range(n_ref):
        o += 8 + struct.unpack_from('<i', data, o)[0]
    buf = np.frombuffer(data + b'\0' * 16, dtype=np.uint8)
    cap = len(data)
    blob = np.zeros(cap, np.uint8); rec_off = np.zeros(cap // 36 + 2, np.uint32)
    bc = np.zeros(cap // 36 + 2, np.uint64); umi = bc.copy(); gxh = bc.copy()
    score = np.zeros(cap // 36 + 2, np.int32); hi = score.copy(); gxa = np.zeros(cap // 36 + 2, np.uint8)
    seen = C.c_long(0)
    p = lambda a: C.c_void_p(a.ctypes.data)
    n = shim.host_parse_pack(p(buf), C.c_uint32(len(data)), C.c_uint32(o), b'CB', b'UB', b'GX', 0, lean, p(blob), p(rec_off), p(bc), p(umi),
                             p(gxh), p(score), p(hi), p(gxa), C.byref(seen))
    assert n > 100 and seen.value == 1780
    with bamio.BamStream(path, barcode_tag='CB', umi_tag='UB', gene_tag='GX', keys=True, lean=bool(lean), n_threads=2) as st:
        chunks = list(st)
        assert len(chunks) == 1
        ch = chunks[0]
        assert ch.n_units == n
        assert np.array_equal(ch.rec_off, rec_off[:n + 1])
        assert np.array_equal(ch.blob, blob[:int(rec_off[n]) * 16])
        assert np.array_equal(ch.barcode_key, bc[:n]) and np.array_equal(ch.umi_key, umi[:n])
        assert np.array_equal(ch.score, score[:n]) and np.array_equal(ch.hi, hi[:n]) and np.array_equal(ch.gx_assigned, gxa[:n])
        ch.close()
    # gene hashes: the ids the string path reports
    with bamio.PackedBam(path, barcode_tag='CB', umi_tag='UB', gene_tag='GX', n_threads=1) as pb:
        want = [shim.host_fnv1a(g.encode()) if g else 0 for g in pb.gene]
    assert [int(x) for x in gxh[:n]] == want
