"""UMI consensus -- device side of dynast/preprocessing/consensus.py (reference).

`call_consensus_groups` covers consensus.py:57-200 (call_consensus_from_reads) for every (gene, barcode, UMI)
group of a BAM at once and the tag arithmetic of :288-316; it returns the consensus alignments as packed records
(the layout the tally reads) plus their tags.  Writing them back as a sorted, indexed consensus.bam together
with the pass-through reads (consensus.py:360-494) is SURVEY section 8 row f4 ("next")."""
from typing import Dict, List, Optional

import numpy as np
import pandas as pd
import torch

from .. import _lib, bamio, pipeline


def mate_offsets(blob: np.ndarray, rec_off: np.ndarray) -> np.ndarray:
    """16-byte offset of the second record of every unit (equal to the next unit's offset when there is none)."""
    o = rec_off[:-1].astype(np.int64) * 16
    u16 = lambda at: blob[o + at].astype(np.int64) | (blob[o + at + 1].astype(np.int64) << 8)
    l_seq, n_cigar, n_tok = u16(8), u16(10), u16(12)
    size = 16 + 4 * n_cigar + 4 * n_tok + ((((l_seq + 1) // 2) + 3) & ~3) + l_seq
    return ((o + ((size + 15) & ~15)) // 16).astype(np.int64)


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
    `gene_tag`.  Returns a dict: records / off (packed consensus records), nm, status, and per group gene,
    barcode, umi, n_reads (the RN tag), AS (sum of member AS), reverse (consensus strand '-')."""
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
