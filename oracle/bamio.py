"""TEST INFRASTRUCTURE ONLY -- a small pure-Python BAM reader/writer with pysam-like accessors.

The reference decodes BAM through ``pysam`` (htslib), which is not installable here
(requirements.txt:7 ``pysam>=0.16.0.1``, un-vendored).  This module restates the published
SAM/BAM specification (SAMv1 section 4.2) and the documented semantics of the ``pysam.AlignedSegment``
accessors the hot path calls (call sites: dynast/preprocessing/bam.py:253-429, coverage.py:101-149,
consensus.py:73-97):

* ``get_aligned_pairs(matches_only=True, with_seq=True)`` -> (qpos, refpos, refbase) for M/=/X only;
  refbase comes from the MD tag and is lower-case where the read mismatches.
* ``get_aligned_pairs(matches_only=False, with_seq=True)`` additionally yields (qpos, None, None) for
  I/S, (None, refpos, base) for D and (None, refpos, None) for N.
* ``get_reference_sequence()`` -> reference bases over M/=/X **and** D (no N, no I, no S).
* ``get_reference_positions()`` -> refpos of M/=/X only.
* ``query_sequence`` / ``query_qualities`` include soft-clipped bases; ``reference_end`` is
  pos + reference-consuming ops (M, D, N, =, X).

Pinned by reproducing the reference's fixture ``alignments.csv`` / ``conversions.csv`` /
``conversions.idx`` byte for byte (tests/test_oracle_parse.py).  The writer exists to build synthetic
BAMs for tests and the bench; BGZF blocks are plain gzip members with the ``BC`` extra field.
"""
import gzip
import struct
import zlib
from typing import Dict, Iterator, List, Optional, Tuple

SEQ_CODES = '=ACMGRSVTWYHKDBN'
CIGAR_OPS = 'MIDNSHP=X'
CIGAR_CONSUMES_QUERY = (1, 1, 0, 0, 1, 0, 0, 1, 1)
CIGAR_CONSUMES_REF = (1, 0, 1, 1, 0, 0, 0, 1, 1)

_TAG_FMT = {'c': 'b', 'C': 'B', 's': 'h', 'S': 'H', 'i': 'i', 'I': 'I', 'f': 'f'}


class Read:
    """One BAM alignment record (all fields decoded eagerly; fine for test-sized inputs)."""

    __slots__ = (
        'query_name', 'flag', 'reference_id', 'reference_name', 'reference_start', 'mapping_quality', 'cigar',
        'next_reference_id', 'next_reference_start', 'template_length', 'query_sequence', 'query_qualities', 'tags',
        '_pairs'
    )

    def __init__(self):
        self._pairs = None

    # ---- flags
    @property
    def is_paired(self):
        return bool(self.flag & 0x1)

    @property
    def is_proper_pair(self):
        return bool(self.flag & 0x2)

    @property
    def is_unmapped(self):
        return bool(self.flag & 0x4)

    @property
    def is_reverse(self):
        return bool(self.flag & 0x10)

    @property
    def is_read1(self):
        return bool(self.flag & 0x40)

    @property
    def is_secondary(self):
        return bool(self.flag & 0x100)

    @property
    def is_duplicate(self):
        return bool(self.flag & 0x400)

    # ---- tags
    def has_tag(self, tag):
        return tag in self.tags

    def get_tag(self, tag):
        return self.tags[tag]

    # ---- geometry
    @property
    def reference_end(self):
        return self.reference_start + sum(n for op, n in self.cigar if CIGAR_CONSUMES_REF[op])

    def _resolve(self):
        """Walk CIGAR with the MD tag expanded -> list of (qpos|None, refpos|None, refbase|None)."""
        if self._pairs is not None:
            return self._pairs
        md = self.tags['MD']
        seq = self.query_sequence
        # Expand MD into per-reference-position entries over the M/=/X + D span:
        # None = match (reference equals read), letter = mismatch/deleted reference base.
        md_items = []  # list of ('=', count) | ('X', base) | ('D', base)
        i = 0
        while i < len(md):
            c = md[i]
            if c.isdigit():
                j = i
                while j < len(md) and md[j].isdigit():
                    j += 1
                if int(md[i:j]) > 0:
                    md_items.append(['=', int(md[i:j])])
                i = j
            elif c == '^':
                j = i + 1
                while j < len(md) and md[j].isalpha():
                    md_items.append(['D', md[j]])
                    j += 1
                i = j
            else:
                md_items.append(['X', c])
                i += 1
        mi = 0
        pairs = []
        qpos = 0
        rpos = self.reference_start
        for op, n in self.cigar:
            if op in (0, 7, 8):  # M = X
                for _ in range(n):
                    item = md_items[mi]
                    if item[0] == '=':
                        base = seq[qpos]
                        item[1] -= 1
                        if item[1] == 0:
                            mi += 1
                    else:
                        assert item[0] == 'X', 'MD/CIGAR mismatch'
                        base = item[1].lower()
                        mi += 1
                    pairs.append((qpos, rpos, base))
                    qpos += 1
                    rpos += 1
            elif op == 2:  # D
                for _ in range(n):
                    item = md_items[mi]
                    assert item[0] == 'D', 'MD/CIGAR mismatch'
                    pairs.append((None, rpos, item[1]))
                    mi += 1
                    rpos += 1
            elif op == 3:  # N
                for _ in range(n):
                    pairs.append((None, rpos, None))
                    rpos += 1
            elif op in (1, 4):  # I S
                for _ in range(n):
                    pairs.append((qpos, None, None))
                    qpos += 1
            # H, P: nothing
        self._pairs = pairs
        return pairs

    def get_aligned_pairs(self, matches_only=False, with_seq=False):
        pairs = self._resolve()
        if matches_only:
            pairs = [p for p in pairs if p[0] is not None and p[1] is not None]
        if with_seq:
            return pairs
        return [(q, r) for q, r, _ in pairs]

    def get_reference_sequence(self):
        return ''.join(b for q, r, b in self._resolve() if r is not None and b is not None)

    def get_reference_positions(self):
        return [r for q, r, b in self._resolve() if q is not None and r is not None]


def _parse_tags(buf: bytes) -> Dict[str, object]:
    tags = {}
    i = 0
    n = len(buf)
    while i < n:
        tag = buf[i:i + 2].decode()
        typ = chr(buf[i + 2])
        i += 3
        if typ == 'A':
            tags[tag] = chr(buf[i])
            i += 1
        elif typ in _TAG_FMT:
            fmt = '<' + _TAG_FMT[typ]
            tags[tag] = struct.unpack_from(fmt, buf, i)[0]
            i += struct.calcsize(fmt)
        elif typ in 'ZH':
            j = buf.index(b'\0', i)
            tags[tag] = buf[i:j].decode()
            i = j + 1
        elif typ == 'B':
            sub = chr(buf[i])
            cnt = struct.unpack_from('<I', buf, i + 1)[0]
            fmt = '<%d%s' % (cnt, _TAG_FMT[sub])
            tags[tag] = list(struct.unpack_from(fmt, buf, i + 5))
            i += 5 + struct.calcsize(fmt)
        else:
            raise ValueError(f'unknown tag type {typ}')
    return tags


def read_bam(path: str) -> Tuple[str, List[Tuple[str, int]], List[Read]]:
    """Decode a whole (small) BAM. Returns (header text, [(contig, length)], reads in file order)."""
    with gzip.open(path, 'rb') as f:
        data = f.read()
    assert data[:4] == b'BAM\1'
    l_text = struct.unpack_from('<i', data, 4)[0]
    text = data[8:8 + l_text].decode(errors='replace')
    off = 8 + l_text
    n_ref = struct.unpack_from('<i', data, off)[0]
    off += 4
    refs = []
    for _ in range(n_ref):
        l_name = struct.unpack_from('<i', data, off)[0]
        name = data[off + 4:off + 4 + l_name - 1].decode()
        l_ref = struct.unpack_from('<i', data, off + 4 + l_name)[0]
        refs.append((name, l_ref))
        off += 8 + l_name
    reads = []
    n = len(data)
    while off < n:
        block_size = struct.unpack_from('<i', data, off)[0]
        rec = data[off + 4:off + 4 + block_size]
        off += 4 + block_size
        (ref_id, pos, l_read_name, mapq, _bin, n_cigar, flag, l_seq, next_ref, next_pos,
         tlen) = struct.unpack_from('<iiBBHHHiiii', rec, 0)
        r = Read()
        p = 32
        r.query_name = rec[p:p + l_read_name - 1].decode()
        p += l_read_name
        r.cigar = [(c & 0xf, c >> 4) for c in struct.unpack_from('<%dI' % n_cigar, rec, p)]
        p += 4 * n_cigar
        nb = (l_seq + 1) // 2
        sb = rec[p:p + nb]
        p += nb
        r.query_sequence = ''.join(SEQ_CODES[sb[i >> 1] >> 4] if (i & 1) == 0 else SEQ_CODES[sb[i >> 1] & 0xf]
                                   for i in range(l_seq))
        r.query_qualities = list(rec[p:p + l_seq])
        p += l_seq
        r.tags = _parse_tags(rec[p:])
        r.flag = flag
        r.reference_id = ref_id
        r.reference_name = refs[ref_id][0] if ref_id >= 0 else None
        r.reference_start = pos
        r.mapping_quality = mapq
        r.next_reference_id = next_ref
        r.next_reference_start = next_pos
        r.template_length = tlen
        reads.append(r)
    return text, refs, reads


# ------------------------------------------------------------------------------------------------
# Writer (synthetic BAMs for tests / bench)
# ------------------------------------------------------------------------------------------------
def _bgzf_block(payload: bytes, level: int = 1) -> bytes:
    comp = zlib.compressobj(level, zlib.DEFLATED, -15)
    cdata = comp.compress(payload) + comp.flush()
    bsize = len(cdata) + 25
    return (b'\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00' + struct.pack('<H', bsize) + cdata +
            struct.pack('<II', zlib.crc32(payload) & 0xffffffff, len(payload)))


BGZF_EOF = bytes.fromhex('1f8b08040000000000ff0600424302001b0003000000000000000000')


def encode_record(
    name: str,
    flag: int,
    ref_id: int,
    pos: int,
    mapq: int,
    cigar: List[Tuple[int, int]],
    seq: str,
    qual: bytes,
    tags: List[Tuple[str, str, object]],
    next_ref: int = -1,
    next_pos: int = -1,
    tlen: int = 0
) -> bytes:
    """tags: list of (tag, type char in 'AcCsSiIZ', value)."""
    code = {c: i for i, c in enumerate(SEQ_CODES)}
    l_seq = len(seq)
    sb = bytearray((l_seq + 1) // 2)
    for i, c in enumerate(seq):
        sb[i >> 1] |= code[c] << (4 if (i & 1) == 0 else 0)
    tb = bytearray()
    for tag, typ, val in tags:
        tb += tag.encode() + typ.encode()
        if typ == 'Z':
            tb += val.encode() + b'\0'
        elif typ == 'A':
            tb += val.encode()
        else:
            tb += struct.pack('<' + _TAG_FMT[typ], val)
    nm = name.encode() + b'\0'
    end = pos + sum(n for op, n in cigar if CIGAR_CONSUMES_REF[op])
    body = struct.pack(
        '<iiBBHHHiiii', ref_id, pos, len(nm), mapq, _reg2bin(pos, end), len(cigar), flag, l_seq, next_ref, next_pos,
        tlen
    ) + nm + struct.pack('<%dI' % len(cigar), *[(n << 4) | op for op, n in cigar]) + bytes(sb) + bytes(qual) + bytes(tb)
    return struct.pack('<i', len(body)) + body


def _reg2bin(beg, end):
    end -= 1
    if beg >> 14 == end >> 14:
        return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17:
        return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20:
        return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23:
        return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26:
        return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


def write_bam(path: str, refs: List[Tuple[str, int]], records: Iterator[bytes], header_text: Optional[str] = None):
    """Write a BAM from pre-encoded records (see `encode_record`)."""
    if header_text is None:
        header_text = '@HD\tVN:1.4\tSO:coordinate\n' + ''.join(f'@SQ\tSN:{n}\tLN:{l}\n' for n, l in refs)
    hdr = bytearray(b'BAM\1' + struct.pack('<i', len(header_text)) + header_text.encode() +
                    struct.pack('<i', len(refs)))
    for n, l in refs:
        nb = n.encode() + b'\0'
        hdr += struct.pack('<i', len(nb)) + nb + struct.pack('<i', l)
    with open(path, 'wb') as f:
        buf = bytearray(hdr)
        for rec in records:
            if len(buf) + len(rec) > 0xff00 and buf:
                f.write(_bgzf_block(bytes(buf)))
                buf = bytearray()
            buf += rec
            while len(buf) > 0xff00:
                f.write(_bgzf_block(bytes(buf[:0xff00])))
                buf = buf[0xff00:]
        if buf:
            f.write(_bgzf_block(bytes(buf)))
        f.write(BGZF_EOF)
