"""TEST INFRASTRUCTURE ONLY -- numpy generator of C2-like packed read records for the CPU legs of bench.py
(`--impl reference`): the reference arm must not load the product library, so its input cannot come from
csrc/synth.cu.  Same generative model: 91-nt single-end reads, CIGAR mix 70 % LM / 18 %
aM bN cM / 8 % sS(L-s)M / 2 % one 1-3 nt D / 2 % one 1-3 nt I, ~1.2 MD letters per read (labelled reads convert T>C
w.p. 0.05 per T, every base is substituted w.p. 1.5e-3, 0.1 % N calls), Phred 90 % in 36..40 / 10 % in 2..27.
Record layout: include/dynast_b200.h (full, or lean with the mismatch Phred inside the MD tokens).
"""
import numpy as np

CODE = np.array([1, 2, 4, 8], dtype=np.uint8)


def make_records(n: int, seed: int = 2001, read_len: int = 91, lean: bool = True, max_tok: int = 12):
    """Returns (blob uint8, rec_off uint32 [n + 1] in 16-byte units) of n single-end reads."""
    rng = np.random.default_rng(seed)
    L = read_len
    kind = rng.choice(5, n, p=[0.70, 0.18, 0.08, 0.02, 0.02])
    a = 10 + rng.integers(0, L - 20, n)
    gap = np.where(kind == 1, 50 + rng.integers(0, 4950, n), 1 + rng.integers(0, 3, n))
    n_cig = np.where(kind == 0, 1, np.where(kind == 2, 2, 3))
    cig = np.zeros((n, 3), dtype=np.uint32)
    cig[:, 0] = np.where(kind == 0, L << 4, np.where(kind == 2, ((1 + a % 19) << 4) | 4, a << 4))
    soft = 1 + a % 19
    cig[:, 1] = np.where(kind == 1, (gap << 4) | 3, np.where(kind == 2, (L - soft) << 4, np.where(kind == 3, (gap << 4) | 2, (gap << 4) | 1)))
    cig[:, 2] = np.where(kind == 4, (L - a - gap) << 4, (L - a) << 4)
    m_total = np.where(kind == 2, L - soft, np.where(kind == 4, L - gap, L))
    n_del = np.where(kind == 3, gap, 0)
    # read bases, qualities
    bases = rng.integers(0, 4, (n, L), dtype=np.uint8)
    is_n = rng.random((n, L)) < 0.001
    nib = np.where(is_n, 15, CODE[bases]).astype(np.uint8)
    pad = (-L) % 8
    nibp = np.concatenate([nib, np.zeros((n, pad + (1 if (L + pad) % 2 else 0)), np.uint8)], axis=1)
    seq = (nibp[:, 0::2] << 4) | nibp[:, 1::2]
    seq_bytes = (((L + 1) // 2) + 3) & ~3
    seq = seq[:, :seq_bytes]
    qual = np.where(rng.random((n, L)) < 0.9, rng.integers(36, 41, (n, L)), rng.integers(2, 28, (n, L))).astype(np.uint8)
    # MD letters: mismatches at distinct ascending aligned offsets
    labelled = rng.random(n) < rng.beta(2, 5, n)
    lam = 0.0015 * 3 * m_total / 3 + 0.001 * m_total + np.where(labelled, 0.05 * m_total / 4, 0.0)
    n_mm = np.minimum(rng.poisson(lam), max_tok - 3)
    raw = np.sort(rng.integers(0, 1 << 30, (n, max_tok - 3)), axis=1)
    aoff = (raw % np.maximum(m_total - (max_tok - 3), 1)[:, None]) + np.arange(max_tok - 3)[None, :]
    aoff = np.sort(aoff, axis=1)
    aoff = np.minimum(aoff + 0, (m_total - 1)[:, None])
    # strictly ascending after the clamp
    aoff = np.minimum.accumulate((aoff - np.arange(max_tok - 3)[None, :])[:, ::-1], axis=1)[:, ::-1] + np.arange(max_tok - 3)[None, :]
    aoff = np.clip(aoff, 0, None)
    # query position of an aligned offset (for the lean layout's Phred): leading soft clip / insertion shift
    lead = np.where(kind == 2, soft, 0)
    qpos = aoff + lead[:, None] + np.where((kind == 4)[:, None] & (aoff >= a[:, None]), gap[:, None], 0)
    qpos = np.minimum(qpos, L - 1)
    read_code = np.take_along_axis(nib, qpos, axis=1)
    ref_idx = (np.take_along_axis(bases, qpos, axis=1) + 1 + rng.integers(0, 3, qpos.shape)) % 4
    tok_mm = aoff.astype(np.uint32) | (CODE[ref_idx].astype(np.uint32) << 16)
    if lean:
        tok_mm |= np.take_along_axis(qual, qpos, axis=1).astype(np.uint32) << 21
    del read_code
    n_tok = n_mm + n_del
    raw_bytes = 16 + 4 * n_cig + 4 * n_tok + seq_bytes + (0 if lean else L)
    size16 = (raw_bytes + 15) // 16
    rec_off = np.zeros(n + 1, dtype=np.uint32)
    np.cumsum(size16, out=rec_off[1:])
    blob = np.zeros(int(rec_off[-1]) * 16, dtype=np.uint8)
    base = rec_off[:-1].astype(np.int64) * 16
    hdr = np.zeros((n, 4), dtype=np.uint32)
    hdr[:, 0] = np.sort(rng.integers(0, 1 << 27, n)).astype(np.uint32)
    hdr[:, 1] = rng.integers(0, 25, n)
    hdr[:, 2] = L | (n_cig.astype(np.uint32) << 16)
    hdr[:, 3] = n_tok.astype(np.uint32) | ((np.where(rng.random(n) < 0.5, 16, 0) | (0x2000 if lean else 0)).astype(np.uint32) << 16)
    b32 = blob.view(np.uint32)
    w = base // 4
    for j in range(4):
        b32[w + j] = hdr[:, j]
    for j in range(3):
        m = n_cig > j
        b32[w[m] + 4 + j] = cig[m, j]
    # tokens in MD order: mismatches before the deletion point, the deleted bases, the rest
    tw = w + 4 + n_cig
    before = (aoff < a[:, None]) | (kind != 3)[:, None]
    n_before = np.minimum((before & (np.arange(max_tok - 3)[None, :] < n_mm[:, None])).sum(axis=1), n_mm)
    for j in range(max_tok - 3):
        m = n_mm > j
        pos = np.where(j < n_before, j, j + n_del)
        b32[tw[m] + pos[m]] = tok_mm[m, j]
    for j in range(3):
        m = n_del > j
        b32[tw[m] + n_before[m] + j] = a[m].astype(np.uint32) | (np.uint32(CODE[j % 4]) << 16) | (1 << 20)
    sb = base + 16 + 4 * n_cig + 4 * n_tok
    cols = np.arange(seq_bytes)
    blob[(sb[:, None] + cols[None, :]).ravel()] = seq.ravel()
    if not lean:
        qb = sb + seq_bytes
        blob[(qb[:, None] + np.arange(L)[None, :]).ravel()] = qual.ravel()
    return blob, rec_off
