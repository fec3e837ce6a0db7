"""Packed read records: the interchange format between the host BAM decoder and the tally kernel
(layout in include/dynast_b200.h).  The packer here is the slow, obviously-correct Python one used by
tests and small inputs; the production packer is the C++ BGZF/BAM decoder in csrc/bampack.cpp."""
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

HAS_MATE = 0x8000
SEQ_CODES = '=ACMGRSVTWYHKDBN'
_CODE = {c: i for i, c in enumerate(SEQ_CODES)}
_CODE.update({c.lower(): i for i, c in enumerate(SEQ_CODES)})

CONVERSION_COLUMNS = ['AC', 'AG', 'AT', 'CA', 'CG', 'CT', 'GA', 'GC', 'GT', 'TA', 'TC', 'TG']
BASE_COLUMNS = ['A', 'C', 'G', 'T']
COLUMNS = CONVERSION_COLUMNS + BASE_COLUMNS


def pack_one(pos: int, ref_id: int, flag: int, cigar: Sequence[Tuple[int, int]], seq: str, qual: Sequence[int],
             md: str) -> bytes:
    """One record: 16-byte header + cigar + 4-bit seq (padded to 4) + qual + MD, padded to 16 bytes."""
    l_seq = len(seq)
    hdr = np.zeros(1, dtype=[('pos', '<i4'), ('ref', '<i4'), ('l', '<u2'), ('nc', '<u2'), ('ml', '<u2'), ('fl', '<u2')])
    mdb = md.encode()
    hdr['pos'], hdr['ref'], hdr['l'], hdr['nc'], hdr['ml'], hdr['fl'] = pos, ref_id, l_seq, len(cigar), len(mdb), flag
    cig = np.array([(n << 4) | op for op, n in cigar], dtype='<u4').tobytes()
    nb = (l_seq + 1) // 2
    sb = bytearray((nb + 3) & ~3)
    for i, c in enumerate(seq):
        sb[i >> 1] |= _CODE[c] << (4 if (i & 1) == 0 else 0)
    body = hdr.tobytes() + cig + bytes(sb) + np.asarray(qual, dtype=np.uint8).tobytes() + mdb
    return body + b'\0' * (-len(body) % 16)


def pack_units(units: Iterable[List[bytes]]) -> Tuple[np.ndarray, np.ndarray]:
    """units: each a list of 1 (single read) or 2 (read, first-seen mate) pre-packed records.
    Returns (blob uint8 [16-byte aligned storage], rec_off uint32 [n+1] in 16-byte units)."""
    chunks = []
    offs = [0]
    total = 0
    for u in units:
        if len(u) == 2:
            first = bytearray(u[0])
            fl = int.from_bytes(first[14:16], 'little') | HAS_MATE
            first[14:16] = fl.to_bytes(2, 'little')
            u = [bytes(first), u[1]]
        for rec in u:
            chunks.append(rec)
            total += len(rec)
        offs.append(total // 16)
    blob = aligned_empty(max(total, 16))
    blob[:total] = np.frombuffer(b''.join(chunks), dtype=np.uint8) if total else 0
    return blob[:total] if total else blob[:0], np.asarray(offs, dtype=np.uint32)


def aligned_empty(nbytes: int, align: int = 256) -> np.ndarray:
    raw = np.empty(nbytes + align, dtype=np.uint8)
    shift = (-raw.ctypes.data) % align
    return raw[shift:shift + nbytes]


def decode_conv(rec: np.ndarray):
    """Split uint64 conversion records into (genome_i, ref letter, read letter, quality) arrays."""
    pos = (rec & 0xffffffff).astype(np.int64)
    code = ((rec >> np.uint64(32)) & np.uint64(0xff)).astype(np.int64)
    qual = ((rec >> np.uint64(40)) & np.uint64(0xff)).astype(np.int64)
    letters = np.array(list(SEQ_CODES))
    return pos, letters[code >> 4], letters[code & 15], qual


def snp_keys(snps: Optional[dict], contig_ids: dict) -> np.ndarray:
    """{conversion: {contig: set(pos)}} (snp.read_snps layout, snp.py:17-37) -> sorted uint64 keys
    (conv_idx << 56 | ref_id << 32 | pos) for the tally's SNP filter. Contigs absent from the BAM
    header can never match a read and are dropped."""
    keys = []
    for conv, by_contig in (snps or {}).items():
        if conv not in CONVERSION_COLUMNS:
            continue
        ci = CONVERSION_COLUMNS.index(conv)
        for contig, positions in by_contig.items():
            if contig not in contig_ids:
                continue
            rid = contig_ids[contig]
            for p in positions:
                keys.append((ci << 56) | (rid << 32) | int(p))
    return np.unique(np.asarray(keys, dtype=np.uint64))
