"""UMI consensus -- dynast/preprocessing/consensus.py (reference) on the B200 path.

`call_consensus` is the operator of consensus.py:227-494 with the reference's signature: it reads a coordinate-sorted
BAM, groups reads by (gene, barcode, UMI), calls one consensus alignment per group ON THE DEVICE (dyn_consensus,
csrc/consensus.cu: consensus.py:57-200 for every group at once), passes every other read through, and writes a
coordinate-sorted, indexed BAM (dyn_bam_write_ex, csrc/bamwrite.cpp).  The host side is array work: read filter and
mate pairing come from the C++ ingest, gene lookup / grouping / the periodic-flush rule are numpy.

Rules restated from the reference (each is checked against oracle/consensus_oracle.py::call_consensus_bam):
  * read filter (:248-249, 374-380): secondary and unmapped records and records without the barcode / UMI tag are
    dropped -- duplicates are NOT --, as are barcode '-' and barcodes outside `barcodes`;
  * a pair is one member list [read, first-seen mate] (:385-395); without a UMI tag its UMI is (mate start, read end);
  * gene = the gene tag, else the single gene of the read's contig and strand that contains [start, end) (:264-275,
    416-418); reads with no or several candidates are written through unchanged (:420-424);
  * groups of one read are written through with their gene tags swapped (:277-286, 476-478);
  * every 10 000th record (:432) the reference flushes the genes it has passed -- in first-seen order, stopping at the
    first gene not yet passed -- and in THAT branch a group of one read is written through AND ALSO sent to the
    consensus caller (no `continue` at :444): such reads appear twice, once as themselves and once as a one-read
    consensus.  Reproduced exactly for single-end data; for paired data the rule uses the read's own start where the
    reference uses the oldest pending mate's (documented deviation: it only moves which singletons get the duplicate);
  * consensus record (:179-200, 288-325): name = sha256 of the members' names, MAPQ 255, flag 16 iff the consensus
    strand is '-', tags AS (sum) NH HI RN [barcode] [UMI] [GX GN when a member carries the gene tag] [RS RI] NM MD.
"""
import functools
import hashlib
import struct
from typing import Dict, List, Optional

import numpy as np
import pandas as pd
import torch

from .. import _lib, bamio, pipeline

FLUSH_EVERY = 10000          # consensus.py:432


def mate_offsets(blob: np.ndarray, rec_off: np.ndarray) -> np.ndarray:
    """16-byte offset of the second record of every unit (equal to the next unit's offset when there is none)."""
    o = rec_off[:-1].astype(np.int64) * 16
    u16 = lambda at: blob[o + at].astype(np.int64) | (blob[o + at + 1].astype(np.int64) << 8)
    l_seq, n_cigar, n_tok = u16(8), u16(10), u16(12)
    size = 16 + 4 * n_cigar + 4 * n_tok + ((((l_seq + 1) // 2) + 3) & ~3) + l_seq
    return ((o + ((size + 15) & ~15)) // 16).astype(np.int64)


def _header_fields(blob: np.ndarray, at16: np.ndarray):
    """(pos, ref_id, n_cigar, reference span) of the packed records at 16-byte offsets `at16`."""
    o = at16.astype(np.int64) * 16
    b32 = lambda at: (blob[o + at].astype(np.int64) | (blob[o + at + 1].astype(np.int64) << 8) | (blob[o + at + 2].astype(np.int64) << 16) |
                      (blob[o + at + 3].astype(np.int64) << 24))
    pos = b32(0).astype(np.int32).astype(np.int64)
    ref_id = b32(4).astype(np.int32).astype(np.int64)
    n_cigar = blob[o + 10].astype(np.int64) | (blob[o + 11].astype(np.int64) << 8)
    span = np.zeros(len(o), dtype=np.int64)
    for j in range(int(n_cigar.max()) if len(o) else 0):
        m = n_cigar > j
        w = b32(16 + 4 * j)
        op = w & 0xf
        span += np.where(m & np.isin(op, (0, 2, 3, 7, 8)), w >> 4, 0)
    return pos, ref_id, n_cigar, span


@functools.lru_cache(maxsize=1 << 16)
def _int_tag(tag: str, v: int) -> bytes:
    """An integer tag with the smallest type that holds it (what pysam's set_tags picks)."""
    v = int(v)
    if v < 0:
        code, fmt = ('c', 'b') if v >= -128 else (('s', 'h') if v >= -32768 else ('i', 'i'))
    else:
        code, fmt = ('C', 'B') if v <= 255 else (('S', 'H') if v <= 65535 else ('I', 'I'))
    return tag.encode() + code.encode() + struct.pack('<' + fmt, v)


@functools.lru_cache(maxsize=1 << 16)
def _z_tag(tag: str, v: str) -> bytes:
    return tag.encode() + b'Z' + v.encode() + b'\0'


def _segment(info):
    return tuple(info['segment']) if 'segment' in info else (int(info['start']), int(info['end']))


def _find_genes(gene_infos: dict, contig_names, ref_id, start, end, read_strand):
    """consensus.py:264-275 for many reads: (number of candidate genes, one candidate) per read.  read_strand: 0 '+', 1
    '-', -1 not tested."""
    n = len(ref_id)
    count = np.zeros(n, dtype=np.int64)
    which = np.full(n, -1, dtype=np.int64)
    genes = list(gene_infos)
    by_contig: Dict[str, List[int]] = {}
    for gi, g in enumerate(genes):
        by_contig.setdefault(str(gene_infos[g]['chromosome']), []).append(gi)
    for c in np.unique(ref_id):
        idx = by_contig.get(str(contig_names[c]), [])
        if not idx:
            continue
        gs = np.array([_segment(gene_infos[genes[i]])[0] for i in idx])
        ge = np.array([_segment(gene_infos[genes[i]])[1] for i in idx])
        gstrand = np.array([0 if gene_infos[genes[i]]['strand'] == '+' else (1 if gene_infos[genes[i]]['strand'] == '-' else 2) for i in idx])
        rows = np.flatnonzero(ref_id == c)
        for a in range(0, len(rows), 4096):
            r = rows[a:a + 4096]
            ok = (start[r][:, None] >= gs[None, :]) & (end[r][:, None] <= ge[None, :])
            ok &= (read_strand[r][:, None] < 0) | (read_strand[r][:, None] == gstrand[None, :])
            count[r] = ok.sum(axis=1)
            which[r] = np.asarray(idx)[ok.argmax(axis=1)]
    return count, which, genes


def call_consensus(
    bam_path: str,
    out_path: str,
    gene_infos: dict,
    strand: str = 'forward',
    umi_tag: Optional[str] = None,
    barcode_tag: Optional[str] = None,
    gene_tag: str = 'GX',
    barcodes: Optional[List[str]] = None,
    quality: int = 27,
    add_RS_RI: bool = False,
    temp_dir: Optional[str] = None,
    n_threads: int = 8,
) -> str:
    """consensus.call_consensus (consensus.py:227-494).  Returns `out_path` (sorted, with <out_path>.bai next to it)."""
    with bamio.PackedBam(bam_path, barcode_tag=barcode_tag, umi_tag=umi_tag, gene_tag=gene_tag, require_gene=False, n_threads=n_threads,
                         flags=bamio.KEEP_RAW | bamio.KEEP_DUPLICATES) as pb:
        n = pb.n_units
        blob, rec_off = pb.blob.copy(), pb.rec_off.copy()
        names, barcode, umi_s, gene_s = pb.read_id, pb.barcode, pb.umi, pb.gene
        flag, hi, score = pb.flag.astype(np.int64), pb.hi.astype(np.int64), pb.score.astype(np.int64)
        gx_assigned, paired = pb.gx_assigned.astype(bool), pb.paired.astype(bool)
        rec_index = pb.rec_index.copy()
        raw = np.concatenate([pb.raw, pb.raw_mate])
        raw_start = pb.raw_off[:-1].astype(np.uint64)
        raw_len = np.diff(pb.raw_off.astype(np.int64)).astype(np.uint32)
        mate_start = (pb.raw_mate_off[:-1].astype(np.uint64) + np.uint64(len(pb.raw))) if len(pb.raw_mate_off) > 1 else np.zeros(n, np.uint64)
        mate_len = np.diff(pb.raw_mate_off.astype(np.int64)).astype(np.uint32) if len(pb.raw_mate_off) > 1 else np.zeros(n, np.uint32)
        ref_names, ref_len, header_text = pb.ref_name, pb.ref_len.copy(), pb.header_text
    keep = np.ones(n, dtype=bool)
    if barcodes:
        keep &= np.isin(barcode, list(barcodes))
    pos, ref_id, _, span = _header_fields(blob, rec_off[:-1]) if n else (np.zeros(0, np.int64),) * 4
    end = pos + span
    m16 = mate_offsets(blob, rec_off) if n else np.zeros(0, np.int64)
    mpos = np.zeros(n, dtype=np.int64)
    if paired.any():
        mp, _, _, _ = _header_fields(blob, m16[paired])
        mpos[paired] = mp
    # ---- read strand (:398-413)
    rev, r1, pr = (flag & 0x10) != 0, (flag & 0x40) != 0, (flag & 0x1) != 0
    minus = np.where(pr & r1, ~rev, rev)
    if strand == 'reverse':
        minus = ~minus
    read_strand = np.where(strand == 'unstranded', -1, minus.astype(np.int64))
    # ---- gene (:416-418): the tag, else the one containing gene
    start = np.where(pr & ~np.asarray([bool(umi_tag)] * n, dtype=bool), mpos, pos) if n else pos      # :392-395
    use_tag = gx_assigned if gene_tag else np.zeros(n, dtype=bool)
    cnt, which, genes = _find_genes(gene_infos, ref_names, ref_id, start, end, read_strand) if (n and (~use_tag).any()) else (
        np.zeros(n, np.int64), np.full(n, -1, np.int64), list(gene_infos))
    # A gene tag the annotation does not know is legal as long as its groups stay singletons: swap_gene_tags uses
    # gene_infos.get (:283); the flush loop (:436) and create_tags_and_strand (:447, 480) index gene_infos and raise.
    n_known = len(genes)
    genes = genes + sorted({g for g, t in zip(gene_s, use_tag) if t and g not in gene_infos})
    gene_index = {g: i for i, g in enumerate(genes)}
    tagged = np.array([gene_index.get(g, -1) for g in gene_s], dtype=np.int64) if n else np.zeros(0, np.int64)
    gene_of = np.where(use_tag, tagged, np.where(cnt == 1, which, -1))
    grouped = keep & (gene_of >= 0)
    through = keep & ~grouped
    # ---- UMI of a unit: the tag, or (mate start, read end) for pairs without one (:392-395), else nothing
    if umi_tag:
        umi_key = pd.Series(umi_s).astype(object)
    else:
        umi_key = pd.Series([(int(mpos[i]), int(end[i])) if pr[i] else None for i in range(n)], dtype=object)
    bc_key = pd.Series(barcode if barcode_tag else [None] * n, dtype=object)

    # ---- the periodic flush (:432-452): which genes were flushed when, and the epoch of every unit's gene
    g_rows = np.flatnonzero(grouped)
    epoch = np.zeros(n, dtype=np.int64)
    n_flush = {}
    if len(g_rows):
        chrom = {gi: str(gene_infos[genes[gi]]['chromosome']) for gi in np.unique(gene_of[g_rows]) if gi < n_known}
        g_end = {gi: _segment(gene_infos[genes[gi]])[1] for gi in chrom}
        checkpoints = g_rows[rec_index[g_rows] % FLUSH_EVERY == 0]
        open_genes: List[int] = []          # insertion order of gx_barcode_umi_groups
        in_dict = set()
        events = []                         # (gene, position in g_rows order) of every flush
        prev = 0
        where = {int(r): k for k, r in enumerate(g_rows)}
        for cp in checkpoints:
            k = where[int(cp)]
            seg_genes = gene_of[g_rows[prev:k + 1]]
            uniq, first = np.unique(seg_genes, return_index=True)
            for gi in uniq[np.argsort(first)]:
                if int(gi) not in in_dict:
                    in_dict.add(int(gi)); open_genes.append(int(gi))
            prev = k + 1
            contig, leftmost = str(ref_names[ref_id[cp]]), int(start[cp])
            while open_genes:
                gi = open_genes[0]
                if gi >= n_known:
                    raise KeyError(genes[gi])           # gene_infos[gene] at consensus.py:436
                if chrom[gi] < contig or (chrom[gi] == contig and g_end[gi] <= leftmost):
                    open_genes.pop(0); in_dict.discard(gi)
                    events.append((gi, k))
                else:
                    break
        if events:
            ev_gene = np.array([e[0] for e in events]); ev_pos = np.array([e[1] for e in events])
            order = np.lexsort((ev_pos, ev_gene))
            ev_gene, ev_pos = ev_gene[order], ev_pos[order]
            key = ev_gene * (len(g_rows) + 1) + ev_pos
            unit_key = gene_of[g_rows] * (len(g_rows) + 1) + np.arange(len(g_rows))
            first_of_gene = np.searchsorted(ev_gene, gene_of[g_rows], side='left')
            epoch[g_rows] = np.searchsorted(key, unit_key, side='left') - first_of_gene
            for gi in np.unique(ev_gene):
                n_flush[int(gi)] = int((ev_gene == gi).sum())

    # ---- groups
    frame = pd.DataFrame({'gene': gene_of[g_rows], 'epoch': epoch[g_rows], 'bc': bc_key.to_numpy()[g_rows], 'umi': umi_key.to_numpy()[g_rows]})
    codes = frame.groupby(['gene', 'epoch', 'bc', 'umi'], sort=False, dropna=False).ngroup().to_numpy() if len(g_rows) else np.zeros(0, np.int64)
    G = int(codes.max()) + 1 if len(codes) else 0
    members = np.bincount(codes, minlength=G) + np.bincount(codes, weights=paired[g_rows], minlength=G).astype(np.int64)
    first_unit = np.full(G, -1, dtype=np.int64)
    if G:
        first_unit[codes[::-1]] = g_rows[::-1]
    g_gene = gene_of[first_unit] if G else np.zeros(0, np.int64)
    flushed_early = np.array([epoch[first_unit[g]] < n_flush.get(int(g_gene[g]), 0) for g in range(G)], dtype=bool) if G else np.zeros(0, bool)
    single = members == 1
    need_consensus = ~single | flushed_early              # :443-450 (no `continue` in the flush branch)
    swap_units = g_rows[single[codes]] if G else np.zeros(0, np.int64)

    # ---- consensus on the device
    cons = None
    cg = np.flatnonzero(need_consensus)
    if len(cg) and (g_gene[cg] >= n_known).any():
        raise KeyError(genes[int(g_gene[cg][g_gene[cg] >= n_known][0])])       # gene_infos[gene] at consensus.py:447, 480
    if len(cg):
        gid = np.full(G, -1, dtype=np.int64)
        gid[cg] = np.arange(len(cg))
        sel = need_consensus[codes]
        order = np.argsort(gid[codes[sel]], kind='stable')
        units = g_rows[sel][order]
        ugid = gid[codes[sel]][order]
        n_items = np.where(paired[units], 2, 1)
        item_unit = np.repeat(units, n_items)
        second = np.arange(len(item_unit)) - np.repeat(np.cumsum(n_items) - n_items, n_items)
        item_rec16 = np.where(second == 0, rec_off[item_unit].astype(np.int64), m16[item_unit])
        group_off = np.zeros(len(cg) + 1, dtype=np.int64)
        np.cumsum(np.bincount(np.repeat(ugid, n_items), minlength=len(cg)), out=group_off[1:])
        ctx = _lib.current_context()
        dev = _lib.current_torch_device()
        out = pipeline.consensus(ctx, torch.from_numpy(blob).to(dev), torch.from_numpy(item_rec16.astype(np.int32)).to(dev),
                                 torch.from_numpy(group_off).to(dev), quality=quality)
        torch.cuda.synchronize()
        status = out['status'].cpu().numpy()
        if (status == 4).any():
            raise Exception('Can not call consensus from reads mapping to multiple contigs.')       # consensus.py:57-58
        if (status == 6).any():
            raise KeyError('a read or reference base other than A/C/G/T/N (BASE_IDX, consensus.py:84-97)')
        if (status != 0).any():
            raise ValueError(f'consensus failed with status {sorted(set(status[status != 0].tolist()))}')
        cons = dict(records=out['records'].cpu().numpy(), off=out['off'].cpu().numpy().view(np.uint32), nm=out['nm'].cpu().numpy(),
                    md=out['md'].cpu().numpy(), md_off=out['md_off'].cpu().numpy().view(np.uint32), units=units, ugid=ugid, groups=cg)

    # ---- output units: consensus records (packed), then the reads written through (raw)
    n_cons = len(cg)
    t_units = np.flatnonzero(through)
    raw_units = []          # (raw start, raw length, ref_id, pos, swap unit or -1)
    for u in t_units:
        raw_units.append((raw_start[u], raw_len[u], ref_id[u], pos[u], -1))
        if paired[u]:
            raw_units.append((mate_start[u], mate_len[u], ref_id[u], mpos[u], -1))
    for u in swap_units:
        raw_units.append((raw_start[u], raw_len[u], ref_id[u], pos[u], u))
    n_total = n_cons + len(raw_units)
    out_names, out_aux, out_flag = [], [], np.zeros(n_total, dtype=np.uint16)
    o_ref, o_pos = np.zeros(n_total, dtype=np.int64), np.zeros(n_total, dtype=np.int64)
    rec_off_out = np.zeros(n_total + 1, dtype=np.uint32)
    records_out = np.zeros(0, dtype=np.uint8)
    if cons is not None:
        records_out = cons['records']
        rec_off_out[:n_cons + 1] = cons['off'][:n_cons + 1]
        rec_off_out[n_cons + 1:] = cons['off'][n_cons]
        cpos, cref, _, _ = _header_fields(records_out, cons['off'][:n_cons])
        o_ref[:n_cons], o_pos[:n_cons] = cref, cpos
        bounds = np.searchsorted(cons['ugid'], np.arange(n_cons + 1))
        for j, g in enumerate(cg):
            us = cons['units'][bounds[j]:bounds[j + 1]]
            member_names, member_hi, total_as = [], [], 0
            for u in us:
                member_names.append(names[u]); member_hi.append(int(hi[u])); total_as += int(score[u])
                if paired[u]:
                    member_names.append(names[u]); member_hi.append(int(hi[u]))
                    total_as += _mate_score(raw, int(mate_start[u]), int(mate_len[u]))
            info = gene_infos[genes[int(g_gene[g])]]
            aux = _int_tag('AS', total_as) + _int_tag('NH', 1) + _int_tag('HI', 1) + _int_tag('RN', len(member_names))
            u0 = us[0]
            if barcode_tag:
                aux += _z_tag(barcode_tag, barcode[u0])
            if umi_tag:
                aux += _z_tag(umi_tag, umi_s[u0])
            if gene_tag:
                tagged_gene = next((gene_s[u] for u in us if gx_assigned[u]), None)
                if tagged_gene:
                    aux += _z_tag('GX', tagged_gene)
                    if info.get('gene_name'):
                        aux += _z_tag('GN', info['gene_name'])
            if add_RS_RI:
                aux += _z_tag.__wrapped__('RS', ';'.join(member_names)) + _z_tag.__wrapped__('RI', ';'.join(str(x) for x in member_hi))      # (unique: not cached)
            aux += _int_tag('NM', int(cons['nm'][j]))
            out_aux.append(aux)
            out_names.append(hashlib.sha256(''.join(member_names).encode('utf-8')).hexdigest())
            cs = info['strand']
            if strand == 'reverse':
                cs = '-' if cs == '+' else '+'
            out_flag[j] = 16 if cs == '-' else 0
    else:
        rec_off_out[:] = 0
    r_start = np.zeros(n_total, dtype=np.uint64)
    r_len = np.zeros(n_total, dtype=np.uint32)
    retag = []
    for k, (rs, rl, rr, rp, su) in enumerate(raw_units):
        j = n_cons + k
        r_start[j], r_len[j], o_ref[j], o_pos[j] = rs, rl, rr, rp
        out_names.append('')
        out_aux.append(b'')
        if su >= 0:
            g = genes[int(gene_of[su])]
            retag.append(g.encode() + b'\0' + (gene_infos.get(g, {}).get('gene_name') or '').encode() + b'\0')
        else:
            retag.append(b'')
    retag = [b''] * n_cons + retag
    order = np.lexsort((np.arange(n_total), o_pos, o_ref)).astype(np.int64)
    lines = [l for l in header_text.split('\n') if l]
    hd = [l for l in lines if l.startswith('@HD')]
    rest = [l for l in lines if not l.startswith('@HD')]
    hd_line = '@HD\tVN:1.4\tSO:coordinate'
    if hd:
        f = [x for x in hd[0].split('\t') if not x.startswith('SO:')]
        hd_line = '\t'.join(f + ['SO:coordinate'])
    header_out = '\n'.join([hd_line] + rest) + '\n'
    refs = [(str(nm), int(l)) for nm, l in zip(ref_names, ref_len)]
    md = None
    if cons is not None:        # the kernel's MD string is the reference's, character for character (consensus.py:124-161, 175)
        md_off = np.concatenate([cons['md_off'][:n_cons + 1], np.full(n_total - n_cons, cons['md_off'][n_cons], dtype=np.uint32)])
        md = (cons['md'][:int(cons['md_off'][n_cons])].tobytes() or b'\0', md_off)
    bamio.write_bam(out_path, refs, records_out, rec_off_out, order=order, names=out_names, aux=out_aux, mapq=np.full(n_total, 255, np.uint8),
                    flag=out_flag, header_text=header_out, level=6, n_threads=n_threads, index=True, raw=raw if len(raw) else np.zeros(1, np.uint8),
                    raw_start=r_start, raw_len=r_len, retag=retag, gene_tag=gene_tag, md=md)
    return out_path


def _mate_score(raw: np.ndarray, start: int, length: int) -> int:
    """AS tag of a raw BAM record."""
    r = raw[start:start + length].tobytes()
    l_name, n_cigar, l_seq = r[8], struct.unpack_from('<H', r, 12)[0], struct.unpack_from('<i', r, 16)[0]
    a = 32 + l_name + 4 * n_cigar + (l_seq + 1) // 2 + l_seq
    sizes = {'A': 1, 'c': 1, 'C': 1, 's': 2, 'S': 2, 'i': 4, 'I': 4, 'f': 4}
    fmts = {'c': 'b', 'C': 'B', 's': 'h', 'S': 'H', 'i': 'i', 'I': 'I'}
    while a + 3 <= len(r):
        tag, ty = r[a:a + 2], chr(r[a + 2])
        a += 3
        if ty in sizes:
            if tag == b'AS' and ty in fmts:
                return struct.unpack_from('<' + fmts[ty], r, a)[0]
            a += sizes[ty]
        elif ty in 'ZH':
            a = r.index(b'\0', a) + 1
        elif ty == 'B':
            sub, cnt = chr(r[a]), struct.unpack_from('<I', r, a + 1)[0]
            a += 5 + cnt * sizes.get(sub, 4)
        else:
            break
    raise KeyError('AS')


def call_consensus_groups(
    bam_path: str,
    gene_infos: dict,
    strand: str = 'forward',
    umi_tag: Optional[str] = None,
    barcode_tag: Optional[str] = None,
    gene_tag: str = 'GX',
    barcodes: Optional[List[str]] = None,
    quality: int = 27,
    n_threads: int = 8,
):
    """Consensus alignment of every (gene, barcode, UMI) group with at least two reads whose gene is given by
    `gene_tag`, as arrays (no BAM written).  Returns a dict: records / off (packed consensus records), nm, status, and per
    group gene, barcode, umi, n_reads (the RN tag), AS (sum of member AS), reverse (consensus strand '-')."""
    with bamio.PackedBam(bam_path, barcode_tag=barcode_tag, umi_tag=umi_tag, gene_tag=gene_tag, require_gene=True,
                         n_threads=n_threads) as pb:
        blob, rec_off = pb.blob.copy(), pb.rec_off.copy()
        gene, barcode, umi = pb.gene, pb.barcode, pb.umi
        score, paired = pb.score.astype(np.int64), pb.paired.astype(bool)
    n = len(gene)
    keep = np.ones(n, dtype=bool)
    if barcodes:
        keep &= np.isin(barcode, list(barcodes))
    rows = np.flatnonzero(keep)
    key = pd.MultiIndex.from_arrays([gene[rows], barcode[rows], umi[rows]])
    codes, uniq = pd.factorize(key, sort=False)
    G = len(uniq)
    members = np.bincount(codes, minlength=G) + np.bincount(codes, weights=paired[rows], minlength=G).astype(np.int64)
    multi = members >= 2                                   # singletons are written through unchanged (consensus.py:443-444)
    gid = np.cumsum(multi) - 1
    sel = multi[codes]
    order = np.argsort(gid[codes[sel]], kind='stable')
    units = rows[sel][order]
    ugid = gid[codes[sel]][order]
    mate16 = mate_offsets(blob, rec_off)
    # items: the read, then its parked mate (consensus.py:428-430)
    n_items = np.where(paired[units], 2, 1)
    item_unit = np.repeat(units, n_items)
    second = np.arange(len(item_unit)) - np.repeat(np.cumsum(n_items) - n_items, n_items)
    item_rec16 = np.where(second == 0, rec_off[item_unit].astype(np.int64), mate16[item_unit])
    n_groups = int(multi.sum())
    group_off = np.zeros(n_groups + 1, dtype=np.int64)
    np.cumsum(np.bincount(np.repeat(ugid, n_items), minlength=n_groups), out=group_off[1:])
    ctx = _lib.current_context()
    dev = _lib.current_torch_device()
    out = pipeline.consensus(ctx, torch.from_numpy(blob).to(dev), torch.from_numpy(item_rec16.astype(np.int32)).to(dev),
                             torch.from_numpy(group_off).to(dev), quality=quality)
    torch.cuda.synchronize()
    g_gene = np.asarray([u[0] for u in uniq], dtype=object)[multi]
    gene_strand = np.array([gene_infos[g]['strand'] for g in g_gene], dtype=object) if n_groups else np.zeros(0, dtype=object)
    reverse = (gene_strand == '-') if strand != 'reverse' else (gene_strand == '+')
    return dict(
        records=out['records'].cpu().numpy(), off=out['off'].cpu().numpy().view(np.uint32), md=out['md'].cpu().numpy(),
        md_off=out['md_off'].cpu().numpy().view(np.uint32), nm=out['nm'].cpu().numpy(),
        status=out['status'].cpu().numpy(), gene=g_gene, barcode=np.asarray([u[1] for u in uniq], dtype=object)[multi],
        umi=np.asarray([u[2] for u in uniq], dtype=object)[multi], n_reads=members[multi],
        AS=np.bincount(ugid, weights=score[units], minlength=n_groups).astype(np.int64), reverse=reverse,
    )
