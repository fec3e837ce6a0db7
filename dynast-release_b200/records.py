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


BAD_MD = 0x4000


def tokenize_md(md: str, cigar: Sequence[Tuple[int, int]]):
    """MD tag text -> (tokens, consistent): one uint32 per MD letter, in MD order -- aligned (M/=/X) offset in
    bits 0..15, reference base (BAM 4-bit code) in bits 16..19, bit 20 for a deleted base (^XYZ run).
    `consistent` is False when the tag disagrees with the CIGAR (matched + mismatched bases != M/=/X length or
    deleted letters != D length); pysam raises on such records.  One deviation from a purely lexical reading, below."""
    toks = []
    aoff = 0
    val = 0
    in_del = False
    n_del = 0
    ok = True
    # Behind a '^', as many letters as the CIGAR's next D operation has are deleted bases (zero counts between them skipped),
    # letters behind them are mismatches: dynast's consensus writes two deleted bases + a mismatch as "148^TAA" (no "0") and
    # a mismatch + three deleted bases as "56C0^G0T0C" (consensus.py:116-176).
    deletions = iter(n for op, n in cigar if op == 2)
    del_left = 0
    for ch in md:
        if ch.isdigit():
            val = val * 10 + int(ch)
            in_del = False
        elif ch == '^':
            aoff += val
            val = 0
            in_del = True
            del_left = next(deletions, 0)
        else:
            code = _CODE.get(ch, 0)
            if del_left > 0:
                if val:
                    ok = False              # matched bases inside a deletion
                val = 0
                toks.append((min(aoff, 0xffff)) | (code << 16) | (1 << 20))
                n_del += 1
                del_left -= 1
            else:
                aoff += val
                val = 0
                if aoff > 0xffff:
                    ok = False
                toks.append((aoff & 0xffff) | (code << 16))
                aoff += 1
    aoff += val
    m_total = sum(n for op, n in cigar if op in (0, 7, 8))
    d_total = sum(n for op, n in cigar if op == 2)
    if aoff != m_total or n_del != d_total:
        ok = False
    return toks, ok


LEAN = 0x2000


def fold_token_quals(toks: List[int], cigar: Sequence[Tuple[int, int]], qual: Sequence[int]) -> List[int]:
    """Lean layout: the Phred of the query base at every non-deleted token's aligned offset goes into bits 21..28."""
    aligned_to_query = {}
    done = q = 0
    for op, n in cigar:
        if op in (0, 7, 8):
            for i in range(n):
                aligned_to_query[done + i] = q + i
            done += n
            q += n
        elif op in (1, 4):
            q += n
    out = []
    for t in toks:
        if not (t >> 20) & 1:
            qpos = aligned_to_query.get(t & 0xffff)
            t |= (int(qual[qpos]) if qpos is not None and qpos < len(qual) else 0) << 21
        out.append(t)
    return out


def pack_one(pos: int, ref_id: int, flag: int, cigar: Sequence[Tuple[int, int]], seq: str, qual: Sequence[int],
             md: str, lean: bool = False) -> bytes:
    """One record: 16-byte header + cigar + MD tokens + 4-bit seq (padded to 4) + qual, padded to 16 bytes.
    lean: the DYN_REC_LEAN layout (no qual stream, mismatch qualities inside the tokens; single-end units only)."""
    l_seq = len(seq)
    toks, ok = tokenize_md(md, cigar)
    if len(toks) > 0xffff:
        toks, ok = [], False
    if lean:
        toks = fold_token_quals(toks, cigar, qual)
        flag |= LEAN
    hdr = np.zeros(1, dtype=[('pos', '<i4'), ('ref', '<i4'), ('l', '<u2'), ('nc', '<u2'), ('nt', '<u2'), ('fl', '<u2')])
    hdr['pos'], hdr['ref'], hdr['l'], hdr['nc'], hdr['nt'] = pos, ref_id, l_seq, len(cigar), len(toks)
    hdr['fl'] = flag | (0 if ok else BAD_MD)
    cig = np.array([(n << 4) | op for op, n in cigar], dtype='<u4').tobytes()
    tok = np.array(toks, dtype='<u4').tobytes()
    nb = (l_seq + 1) // 2
    sb = bytearray((nb + 3) & ~3)
    for i, c in enumerate(seq):
        sb[i >> 1] |= _CODE[c] << (4 if (i & 1) == 0 else 0)
    body = hdr.tobytes() + cig + tok + bytes(sb) + (b'' if lean else np.asarray(qual, dtype=np.uint8).tobytes())
    return body + b'\0' * (-len(body) % 16)


def md_from_tokens(tokens: Sequence[int], cigar: Sequence[Tuple[int, int]]) -> str:
    """Canonical MD text of a tokenised record (the inverse of tokenize_md for well-formed tags)."""
    out = []
    run_start = 0
    i = 0
    m_total = sum(n for op, n in cigar if op in (0, 7, 8))
    prev_del = False
    del_at = 0
    for t in tokens:
        aoff, code, is_del = t & 0xffff, (t >> 16) & 15, (t >> 20) & 1
        if is_del:
            if not prev_del or aoff != del_at:      # another aligned offset: another ^ run
                out.append(str(aoff - run_start))
                out.append('^')
                run_start = del_at = aoff
            out.append(SEQ_CODES[code])
            prev_del = True
        else:
            out.append(str(aoff - run_start))
            out.append(SEQ_CODES[code])
            run_start = aoff + 1
            prev_del = False
    out.append(str(m_total - run_start))
    return ''.join(out)


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
