"""TEST INFRASTRUCTURE ONLY -- numpy stand-ins for the device stages behind dynast_b200.pipeline, so that the
host logic of the operator layer (CSV parsing, interning, split planning, key packing, output formatting) is
covered by the CPU test-suite.  They restate what each C-ABI call is documented to compute
(include/dynast_b200.h); the `-m gpu` tests run the same operator tests against the real kernels.
Installed by the `fake_gpu` fixture (tests/conftest.py); never imported by the product."""
import numpy as np
import torch

CONV = ['AC', 'AG', 'AT', 'CA', 'CG', 'CT', 'GA', 'GC', 'GT', 'TA', 'TC', 'TG']


def _np(t):
    return None if t is None else t.cpu().numpy()


def _complement(counts, strand):
    out = counts.copy()
    rev = strand != 0
    out[rev, :12] = counts[rev, :12][:, ::-1]
    out[rev, 12:] = counts[rev, 12:][:, ::-1]
    return out


def _mask_cols(mask, n=12):
    return [j for j in range(n) if (mask >> j) & 1]


def count_records(ctx, conv_rec, conv_read, conv_contig, counts, quality=27, snps=None):
    rec = _np(conv_rec).view(np.uint64)
    read = _np(conv_read).astype(np.int64)
    contig = _np(conv_contig).astype(np.uint64) if conv_contig is not None else np.zeros(len(rec), dtype=np.uint64)
    snp = set(_np(snps).view(np.uint64).tolist()) if snps is not None else set()
    c = counts.numpy()
    status = 0
    for i in range(len(rec)):
        r = int(rec[i])
        q, code = (r >> 40) & 0xff, (r >> 32) & 0xff
        if q <= quality:
            continue
        gn, rn = code >> 4, code & 15
        if bin(gn).count('1') != 1 or bin(rn).count('1') != 1 or gn == rn:
            status |= 4
            continue
        g, d = gn.bit_length() - 1, rn.bit_length() - 1
        idx = g * 3 + (d - 1 if d > g else d)
        if ((idx << 56) | (int(contig[i]) << 32) | (r & 0xffffffff)) in snp:
            continue
        if c[read[i], idx] == 255:
            status |= 1
        else:
            c[read[i], idx] += 1
    return torch.tensor([status], dtype=torch.int32)


def complement_counts(ctx, counts, strand):
    counts.copy_(torch.from_numpy(_complement(counts.numpy(), strand.numpy())))
    return counts


def dedup(ctx, key_lo, counts, strand, transcriptome, score, conv_n, barcode, n_barcodes, conversions=None,
          use_conversions=True, key_hi=None, key_lo_bits=64, key_hi_bits=0, final_extra=None, final_extra_bits=0,
          split_threshold=50000, plan=None):
    from dynast_b200 import pipeline
    n = counts.shape[0]
    strand_np = _np(strand).astype(np.int64)
    noconv = (_np(conv_n).view(np.uint16) == 0).astype(np.int64)
    idx = np.arange(n, dtype=np.int64)
    if barcode is None:
        ord_of_read = np.zeros(n, dtype=np.int64)
        split_of_read = np.zeros(n, dtype=np.int64)
    else:
        bc = _np(barcode).astype(np.int64)
        if plan is None:
            plan = plan_splits(ctx, barcode, strand, conv_n, n_barcodes, split_threshold)
        ord_of, split_of = _np(plan[0]).view(np.uint32), _np(plan[1]).view(np.uint32)
        ord_of_read, split_of_read = ord_of[bc].astype(np.int64), split_of[bc].astype(np.int64)
    comp = _complement(counts.numpy(), strand_np).astype(np.int64)
    mask = pipeline.conv_mask(conversions)
    conv_sum = comp[:, _mask_cols(mask)].sum(axis=1)
    has = ((conv_sum > 0) & bool(use_conversions and conversions is not None)).astype(np.int64)
    prio = (has << 29) | ((_np(transcriptome) != 0).astype(np.int64) << 28) | (_np(score).view(np.uint16).astype(np.int64) << 12) | conv_sum
    # split-file order, then stable priority sort
    o1 = np.lexsort((idx, noconv, strand_np, ord_of_read))
    order = o1[np.argsort(prio[o1], kind='stable')]
    lo = _np(key_lo).view(np.uint64)
    hi = _np(key_hi).view(np.uint64) if key_hi is not None else np.zeros(n, dtype=np.uint64)
    seen = {}
    for r in order:
        seen[(int(hi[r]), int(lo[r]))] = r      # last in sorted order wins
    win = np.zeros(n, dtype=bool)
    win[list(seen.values())] = True
    winners = order[win[order]]
    extra = _np(final_extra).view(np.uint64) if final_extra is not None else np.zeros(n, dtype=np.uint64)
    fin = np.lexsort((np.arange(len(winners)), extra[winners], strand_np[winners], split_of_read[winners]))
    return torch.from_numpy(winners[fin].astype(np.int32))


def plan_splits(ctx, barcode, strand, conv_n, n_barcodes, split_threshold=50000):
    from dynast_b200 import pipeline
    bc = _np(barcode).astype(np.int64)
    n = len(bc)
    pos = ((_np(strand).astype(np.int64) != 0).astype(np.int64) << 33) | ((_np(conv_n).view(np.uint16) == 0).astype(np.int64) << 32) | np.arange(n)
    first = np.full(n_barcodes, np.iinfo(np.int64).max, dtype=np.int64)
    np.minimum.at(first, bc, pos)
    first = np.where(first == np.iinfo(np.int64).max, -1, first)      # 0xff.. pattern of the kernel
    ord_of, split_of, n_splits = pipeline.split_plan(first, np.bincount(bc, minlength=n_barcodes), split_threshold)
    return torch.from_numpy(ord_of.view(np.int32)), torch.from_numpy(split_of.view(np.int32)), n_splits


def group_sums(ctx, counts, strand, group, n_groups, sel=None, conversions=None):
    from dynast_b200 import pipeline
    c = counts.numpy().astype(np.int64)
    if strand is not None:
        c = _complement(counts.numpy(), strand.numpy()).astype(np.int64)
    g = _np(group).astype(np.int64) if group is not None else np.zeros(len(c), dtype=np.int64)
    if sel is not None:
        s = _np(sel).astype(np.int64)
        c, g = c[s], g[s]
    sums = np.zeros((n_groups, 16), dtype=np.int64)
    np.add.at(sums, g, c)
    n_reads = np.bincount(g, minlength=n_groups).astype(np.int64)
    k = c[:, _mask_cols(pipeline.conv_mask(conversions))].sum(axis=1)
    n_with = np.bincount(g, weights=(k > 0), minlength=n_groups).astype(np.int64)
    return torch.from_numpy(sums), torch.from_numpy(n_reads), torch.from_numpy(n_with)


def rates_pe(ctx, sums, pe_conversions=None):
    from dynast_b200 import pipeline
    s = (sums.numpy() & 0xffffffff).astype(np.float64)
    with np.errstate(divide='ignore', invalid='ignore'):
        rates = s[:, :12] / s[:, 12 + np.arange(12) // 3]
    cols = _mask_cols(pipeline.conv_mask(pe_conversions)) if pe_conversions is not None else []
    p_e = np.full(len(s), np.nan)
    for g in range(len(s)):
        acc, cnt = 0.0, 0
        for j in cols:
            if not np.isnan(rates[g, j]):
                acc += rates[g, j]; cnt += 1
        if cnt:
            p_e[g] = acc / cnt
    return torch.from_numpy(rates), torch.from_numpy(p_e)


def alpha(ctx, n_reads, n_with, pi_c):
    nr, nw, p = n_reads.numpy().astype(np.float64), n_with.numpy().astype(np.float64), pi_c.numpy()
    with np.errstate(divide='ignore', invalid='ignore'):
        out = np.where((p > 0) & (nr > 0), (nw / nr) / p, np.nan)
    return torch.from_numpy(out)


def aggregate_kn(ctx, counts, strand, group, gene, n_groups, n_genes, conversions, sel=None, exclude=None,
                 first_appearance=True):
    from dynast_b200 import pipeline
    c = counts.numpy().astype(np.int64)
    if strand is not None:
        c = _complement(counts.numpy(), strand.numpy()).astype(np.int64)
    g = _np(group).astype(np.int64) if group is not None else np.zeros(len(c), dtype=np.int64)
    gx = _np(gene).astype(np.int64) if gene is not None else np.zeros(len(c), dtype=np.int64)
    if sel is not None:
        s = _np(sel).astype(np.int64)
        c, g, gx = c[s], g[s], gx[s]
    k = c[:, _mask_cols(pipeline.conv_mask(conversions))].sum(axis=1)
    n = c[:, [12 + j for j in _mask_cols(pipeline.base_mask(conversions), 4)]].sum(axis=1)
    keep = np.ones(len(c), dtype=bool)
    if exclude:
        keep = c[:, _mask_cols(pipeline.conv_mask(exclude))].sum(axis=1) == 0
    rows = {}
    for i in np.flatnonzero(keep):
        key = (int(g[i]), int(gx[i]), int(k[i]), int(n[i]))
        if key in rows:
            rows[key][0] += 1
        else:
            rows[key] = [1, i]
    keys = list(rows) if first_appearance else sorted(rows)
    arr = np.array(keys, dtype=np.int64).reshape(-1, 4)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64))
    return dict(group=t(arr[:, 0]), gene=t(arr[:, 1]), k=t(arr[:, 2]), n=t(arr[:, 3]),
                count=t([rows[x][0] for x in keys]), first=t([rows[x][1] for x in keys]))


_ROW = np.dtype([('counts', 'u1', 16), ('umi', '<i8'), ('barcode', '<i4'), ('gene', '<i4'), ('score', '<i2'), ('conv_n', '<i2'),
                 ('strand', 'u1'), ('transcriptome', 'u1'), ('pad', 'u1', 2)])
assert _ROW.itemsize == 40


def owner_partition(ctx, barcode, world):
    """numpy stand-in for dyn_owner_partition."""
    owner = _np(barcode).astype(np.int64) % world
    perm = np.argsort(owner, kind='stable').astype(np.int32)
    return torch.from_numpy(perm), torch.from_numpy(np.bincount(owner, minlength=world).astype(np.int64))


def pack_reads(ctx, perm, counts, barcode, umi, gene, score16, conv_n, strand, transcriptome=None):
    """numpy stand-in for dyn_pack_reads (same 40-byte row layout)."""
    n = barcode.numel()
    idx = _np(perm).astype(np.int64) if perm is not None else np.arange(n)
    rows = np.zeros(n, dtype=_ROW)
    rows['counts'] = _np(counts)[idx]
    rows['umi'], rows['barcode'], rows['gene'] = _np(umi)[idx], _np(barcode)[idx], _np(gene)[idx]
    rows['score'], rows['conv_n'], rows['strand'] = _np(score16)[idx], _np(conv_n)[idx], _np(strand)[idx]
    rows['transcriptome'] = _np(transcriptome)[idx] if transcriptome is not None else 1
    return torch.from_numpy(rows.view(np.uint8).reshape(n, 40))


def unpack_reads(ctx, rows):
    """numpy stand-in for dyn_unpack_reads."""
    r = np.ascontiguousarray(_np(rows)).reshape(-1).view(_ROW)
    return {k: torch.from_numpy(np.ascontiguousarray(r[k])) for k in ('counts', 'barcode', 'umi', 'gene', 'score', 'conv_n', 'strand',
                                                                      'transcriptome')}


def merge_kn(ctx, keys, count32, world=1, group_bits=0):
    """numpy stand-in for dyn_merge_kn."""
    k = _np(keys).astype(np.int64)
    k = (((k >> 22) // world) << 22) | (k & 0x3fffff)
    uk, inv = np.unique(k, return_inverse=True)
    c = np.bincount(inv, weights=_np(count32).astype(np.float64), minlength=len(uk)).astype(np.int32)
    out_k = np.zeros(len(k), np.int64); out_c = np.zeros(len(k), np.int32)
    out_k[:len(uk)] = uk; out_c[:len(uk)] = c
    return torch.from_numpy(out_k), torch.from_numpy(out_c), torch.tensor([len(uk)], dtype=torch.int64)


def kn_csr(ctx, keys, count32, n_rows, n_groups):
    """numpy stand-in for dyn_kn_csr."""
    m = int(n_rows[0])
    k = _np(keys)[:m].astype(np.int64)
    g = k >> 22
    off = np.zeros(n_groups + 1, np.int64)
    off[1:] = np.cumsum(np.bincount(g, minlength=n_groups))
    c = _np(count32)[:m].astype(np.int64)
    total = np.bincount(g, weights=c.astype(np.float64), minlength=n_groups).astype(np.int64)
    t = lambda a, d: torch.from_numpy(np.ascontiguousarray(a.astype(d)))
    return t((k >> 10) & 0xfff, np.int32), t(k & 0x3ff, np.int32), t(c, np.int64), t(off, np.int64), t(total, np.int64)


def consensus(ctx, records, item_rec16, group_off, quality=27, out_capacity_bytes=None):
    """Stand-in for dyn_consensus: members decoded from their packed records, the oracle's call_consensus
    (consensus.py:57-200), the result packed again."""
    from dynast_b200 import records as rec
    from oracle import bamio, consensus_oracle
    blob, items, goff = _np(records), _np(item_rec16).astype(np.int64), _np(group_off).astype(np.int64)

    def unpack(at16):
        o = int(at16) * 16
        pos, ref_id = int(np.frombuffer(blob[o:o + 4].tobytes(), '<i4')[0]), int(np.frombuffer(blob[o + 4:o + 8].tobytes(), '<i4')[0])
        l_seq, n_cig, n_tok, flag = (int(x) for x in np.frombuffer(blob[o + 8:o + 16].tobytes(), '<u2'))
        cig = np.frombuffer(blob[o + 16:o + 16 + 4 * n_cig].tobytes(), '<u4')
        cigar = [(int(c & 0xf), int(c >> 4)) for c in cig]
        toks = [int(x) for x in np.frombuffer(blob[o + 16 + 4 * n_cig:o + 16 + 4 * n_cig + 4 * n_tok].tobytes(), '<u4')]
        so = o + 16 + 4 * n_cig + 4 * n_tok
        sb = blob[so:so + (l_seq + 1) // 2]
        seq = ''.join(rec.SEQ_CODES[sb[i >> 1] >> 4] if (i & 1) == 0 else rec.SEQ_CODES[sb[i >> 1] & 0xf] for i in range(l_seq))
        qo = so + ((((l_seq + 1) // 2) + 3) & ~3)
        return consensus_oracle.make_read('r', pos, cigar, seq, [int(x) for x in blob[qo:qo + l_seq]], rec.md_from_tokens(toks, cigar),
                                          ref_id=ref_id, flag=flag & 0xfff)

    packed, nm, status, mds = [], [], [], []
    for g in range(len(goff) - 1):
        members = [unpack(items[i]) for i in range(goff[g], goff[g + 1])]
        if len({m.reference_id for m in members}) > 1:
            packed.append(rec.pack_one(0, 0, 0, [], '', [], '0')); nm.append(0); status.append(4); mds.append(b'')
            continue
        c = consensus_oracle.call_consensus(members, quality=quality)
        cigar = [(bamio.CIGAR_OPS.index(op), n) for op, n in c['cigar']]
        packed.append(rec.pack_one(c['start'], members[0].reference_id, 0, cigar, c['seq'], c['qual'], c['md']))
        nm.append(c['nm']); status.append(0); mds.append(c['md'].encode())
    out_blob, off = rec.pack_units([[p] for p in packed]) if packed else (np.zeros(0, np.uint8), np.zeros(1, np.uint32))
    t = lambda a, d: torch.from_numpy(np.ascontiguousarray(np.asarray(a).astype(d)))
    md_off = np.zeros(len(packed) + 1, dtype=np.int32)
    np.cumsum([len(m) for m in mds], out=md_off[1:])
    return dict(records=t(out_blob, np.uint8), off=t(off.view(np.int32), np.int32), md=t(np.frombuffer(b''.join(mds), np.uint8), np.uint8),
                md_off=t(md_off, np.int32), nm=t(nm, np.int32), status=t(status, np.int32))


def install(monkeypatch):
    """Route the operator layer to the numpy stand-ins (torch CPU tensors)."""
    from dynast_b200 import _lib, device, pipeline
    import oracle
    from helpers import oracle_em_batch

    class _Ctx:
        handle = None
        lib = None

    monkeypatch.setattr(_lib, 'current_context', lambda: _Ctx())
    monkeypatch.setattr(_lib, 'current_torch_device', lambda: torch.device('cpu'))
    monkeypatch.setattr(_lib, 'current_device_index', lambda: 0)
    for name in ('count_records', 'complement_counts', 'plan_splits', 'dedup', 'group_sums', 'rates_pe', 'alpha', 'aggregate_kn',
                 'owner_partition', 'pack_reads', 'unpack_reads', 'merge_kn', 'kn_csr', 'consensus'):
        monkeypatch.setattr(pipeline, name, globals()[name])

    def em_pc_host(k, n, count, off, p_e, nasc=False, device=0):
        p_c, status = oracle_em_batch(oracle.lib(), k, n, count, off, p_e, nasc=nasc)
        return p_c, status, np.zeros(len(p_e), dtype=np.int32)

    monkeypatch.setattr(device, 'em_pc_host', em_pc_host)
    monkeypatch.setattr(torch.cuda, 'synchronize', lambda *a, **k: None)
    from dynast_b200.preprocessing import snp
    monkeypatch.setattr(snp, 'unique_counts', unique_counts_host)


def unique_counts_host(keys):
    """numpy stand-in for dyn_unique_counts (preprocessing.snp.unique_counts)."""
    keys = np.asarray(keys, dtype=np.uint64)
    valid = keys != np.uint64(0xffffffffffffffff)
    uk, first, cnt = np.unique(keys[valid], return_index=True, return_counts=True)
    pos = np.flatnonzero(valid)[first] if len(uk) else np.zeros(0, dtype=np.int64)
    return uk, cnt.astype(np.uint32), pos.astype(np.uint32)
