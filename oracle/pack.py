"""TEST INFRASTRUCTURE ONLY -- build packed read records from a BAM with the pure-Python reader,
applying the reference's read filter and mate pairing (bam.py:173-174, 242-248, 286-312), so that tests
can feed identical inputs to the C oracle and to the CUDA tally, and can check the C++ BAM packer."""
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import bamio


def _md_of(read):
    return read.tags['MD']


def pack_read(read: bamio.Read, pack_one) -> bytes:
    return pack_one(read.reference_start, read.reference_id, read.flag & 0x0fff, read.cigar, read.query_sequence,
                    read.query_qualities, _md_of(read))


def units_from_bam(
    bam_path: str,
    pack_one,
    pack_units,
    umi_tag: Optional[str] = None,
    barcode_tag: Optional[str] = None,
    gene_tag: str = 'GX',
    barcodes: Optional[set] = None,
    velocity: bool = True,
):
    """Returns (blob, rec_off, meta) where meta[i] = dict(read_id, index, barcode, umi, GX, gx_assigned,
    score, contig) for counting unit i, in the reference's output order (contig by contig, file order;
    a pair is emitted when its second mate is seen)."""
    _, refs, reads = bamio.read_bam(bam_path)
    required = []
    if not velocity:
        required.append(gene_tag)
    if umi_tag:
        required.append(umi_tag)
    if barcode_tag:
        required.append(barcode_tag)
    per_contig: Dict[int, List[bamio.Read]] = {}
    for r in reads:
        if r.reference_id >= 0:
            per_contig.setdefault(r.reference_id, []).append(r)
    units, meta = [], []
    for ref_id in sorted(per_contig):
        pending = {}
        for read in per_contig[ref_id]:
            if read.is_secondary or read.is_unmapped or read.is_duplicate:
                continue
            if any(not read.has_tag(t) for t in required):
                continue
            barcode = read.get_tag(barcode_tag) if barcode_tag else 'NA'
            if barcode == '-' or (barcodes and barcode not in barcodes):
                continue
            hi = read.get_tag('HI')
            unit = [pack_read(read, pack_one)]
            if read.is_paired:
                key = (read.query_name, hi)
                if key not in pending:
                    pending[key] = read
                    continue
                unit.append(pack_read(pending.pop(key), pack_one))
            units.append(unit)
            gx_assigned = read.has_tag(gene_tag)
            meta.append(
                dict(
                    read_id=read.query_name,
                    index=hi,
                    barcode=barcode,
                    umi=read.get_tag(umi_tag) if umi_tag else 'NA',
                    GX=read.get_tag(gene_tag) if gx_assigned else '',
                    gx_assigned=gx_assigned,
                    score=read.get_tag('AS'),
                    contig=refs[ref_id][0],
                    ref_id=ref_id,
                )
            )
    blob, rec_off = pack_units(units)
    return blob, rec_off, meta, refs


def run_tally(blob, rec_off, quality=27, nasc=False, snps=None, emit=True):
    """Run the C oracle tally on packed records. Returns dict of numpy outputs."""
    import ctypes as C
    from . import lib
    n = len(rec_off) - 1
    counts = np.zeros((n, 16), dtype=np.uint8)
    conv_off = np.zeros(n, dtype=np.uint32)
    conv_n = np.zeros(n, dtype=np.uint16)
    cap = int(len(blob)) + 16
    conv_rec = np.zeros(cap, dtype=np.uint64)
    conv_read = np.zeros(cap, dtype=np.uint32)
    total = C.c_uint64(0)
    status = C.c_uint32(0)
    snps = np.ascontiguousarray(snps if snps is not None else np.zeros(0, dtype=np.uint64), dtype=np.uint64)
    blob = np.ascontiguousarray(blob)
    lib().dyn_oracle_tally(
        blob.ctypes.data, rec_off.ctypes.data, 0, n, snps.ctypes.data if len(snps) else None, len(snps), quality,
        1 if nasc else 0, counts.ctypes.data, conv_off.ctypes.data, conv_n.ctypes.data,
        conv_rec.ctypes.data if emit else None, conv_read.ctypes.data if emit else None, cap, C.byref(total),
        C.byref(status)
    )
    t = int(total.value)
    return dict(counts=counts, conv_off=conv_off, conv_n=conv_n, conv_rec=conv_rec[:t], conv_read=conv_read[:t],
                total=t, status=int(status.value))
