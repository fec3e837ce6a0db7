"""Device-resident driver of the count -> estimate hot path (torch tensors own the HBM; the work is
done by the C-ABI kernels on the current torch stream)."""
import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import Context, SynthParams, check, ptr

TALLY_NASC = 0x1
TALLY_EMIT_CONV = 0x2


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class SynthBatch:
    """Synthetic packed reads generated on the device (csrc/synth.cu; model of SURVEY.md 8d, config C2)."""

    def __init__(self, n_reads: int, seed: int = 2001, read_len: int = 91, n_barcodes: int = 10000,
                 n_genes: int = 20000, n_contigs: int = 25, p_c: float = 0.05, p_e: float = 5e-4,
                 frac_n: float = 0.001, umi_len: int = 12, reads_per_umi: float = 2.0, device: int = 0,
                 barcode_base: int = 0, with_keys: bool = True, clustered: bool = False, lean: bool = False):
        self.ctx = Context.get(device)
        dev = torch.device('cuda', device)
        rng = np.random.default_rng(seed)
        w = rng.lognormal(np.log(5000), 0.6, n_barcodes)
        gw = 1.0 / np.arange(1, n_genes + 1)**1.1
        self.tables = dict(
            bc_cdf=torch.from_numpy(np.minimum(np.cumsum(w / w.sum()), 1.0).astype(np.float32)).to(dev),
            gene_cdf=torch.from_numpy(np.minimum(np.cumsum(gw / gw.sum()), 1.0).astype(np.float32)).to(dev),
            gene_pi=torch.from_numpy(rng.beta(2, 5, n_genes).astype(np.float32)).to(dev),
            gene_strand=torch.from_numpy(rng.integers(0, 2, n_genes).astype(np.uint8)).to(dev),
            gene_contig=torch.from_numpy(rng.integers(0, n_contigs, n_genes).astype(np.int32)).to(dev),
        )
        self.tables['bc_cdf'][-1] = 1.0
        self.tables['gene_cdf'][-1] = 1.0
        sp = SynthParams()
        sp.seed, sp.read_len, sp.n_barcodes, sp.n_genes, sp.umi_len = seed, read_len, n_barcodes, n_genes, umi_len
        sp.n_groups = max(1, int(n_reads / reads_per_umi))
        sp.pos_span = -(1 << 27) if clustered else 1 << 27
        sp.p_c, sp.p_e, sp.frac_n = p_c, p_e, frac_n
        sp.lean = int(lean)
        self.lean = bool(lean)
        sp.d_bc_cdf = self.tables['bc_cdf'].data_ptr()
        sp.d_gene_cdf = self.tables['gene_cdf'].data_ptr()
        sp.d_gene_pi = self.tables['gene_pi'].data_ptr()
        sp.d_gene_strand = self.tables['gene_strand'].data_ptr()
        sp.d_gene_contig = self.tables['gene_contig'].data_ptr()
        self.params = sp
        self.n_reads = n_reads
        self.n_barcodes, self.n_genes = n_barcodes, n_genes
        lib = self.ctx.lib
        with torch.cuda.device(dev):
            self.rec_off = torch.empty(n_reads + 1, dtype=torch.int32, device=dev)   # bit pattern: uint32
            check(lib.dyn_synth_offsets(self.ctx.handle, C.byref(sp), n_reads, ptr(self.rec_off), _stream()))
            total16 = int(self.rec_off[-1].item()) & 0xffffffff
            self.n_bytes = total16 * 16
            self.records = torch.empty(self.n_bytes, dtype=torch.uint8, device=dev)
            if with_keys:
                self.barcode = torch.empty(n_reads, dtype=torch.int32, device=dev)
                self.umi = torch.empty(n_reads, dtype=torch.int64, device=dev)
                self.gene = torch.empty(n_reads, dtype=torch.int32, device=dev)
                self.score = torch.empty(n_reads, dtype=torch.int32, device=dev)
                self.strand = torch.empty(n_reads, dtype=torch.uint8, device=dev)
            else:
                self.barcode = self.umi = self.gene = self.score = self.strand = None
            check(
                lib.dyn_synth_fill(
                    self.ctx.handle, C.byref(sp), n_reads, ptr(self.rec_off), ptr(self.records), ptr(self.barcode),
                    ptr(self.umi), ptr(self.gene), ptr(self.score), ptr(self.strand), _stream()
                )
            )
            if barcode_base and with_keys:
                self.barcode += barcode_base
            torch.cuda.current_stream().synchronize()

    def algorithmic_input_bytes(self) -> int:
        """Unpadded record bytes (16 + 4 n_cigar + 4 n_tok + ceil(l/2) [+ l: Phred, absent from lean records] per
        read) + 4-byte offsets."""
        off = (self.rec_off[:-1].to(torch.int64) & 0xffffffff) * 16
        total = 0
        step = 1 << 24
        for i in range(0, self.n_reads, step):
            o = off[i:i + step]
            lo = self.records[o + 8].to(torch.int64) | (self.records[o + 9].to(torch.int64) << 8)
            nc = self.records[o + 10].to(torch.int64) | (self.records[o + 11].to(torch.int64) << 8)
            ml = self.records[o + 12].to(torch.int64) | (self.records[o + 13].to(torch.int64) << 8)
            total += int((16 + 4 * nc + (lo + 1) // 2 + (0 if self.lean else lo) + 4 * ml).sum().item())
        return total + 4 * (self.n_reads + 1)


class TallyBuffers:
    """Outputs of dyn_tally_reads for up to n_reads reads."""

    def __init__(self, n_reads: int, conv_capacity: int, device: int = 0):
        dev = torch.device('cuda', device)
        self.counts = torch.empty((n_reads, 16), dtype=torch.uint8, device=dev)
        self.conv_off = torch.empty(n_reads, dtype=torch.int32, device=dev)
        self.conv_n = torch.empty(n_reads, dtype=torch.int16, device=dev)
        self.conv_rec = torch.empty(conv_capacity, dtype=torch.int64, device=dev)
        self.conv_read = torch.empty(conv_capacity, dtype=torch.int32, device=dev)
        self.scalars = torch.zeros(4, dtype=torch.int64, device=dev)   # [0] conv_total, [1] status
        self.conv_capacity = conv_capacity


def tally_reads(ctx: Context, records: torch.Tensor, rec_off: torch.Tensor, out: TallyBuffers, quality: int = 27,
                nasc: bool = False, emit_conv: bool = True, snps: Optional[torch.Tensor] = None):
    """Asynchronous on the current torch stream (dyn_tally_reads)."""
    n = rec_off.numel() - 1
    out.scalars.zero_()
    flags = (TALLY_NASC if nasc else 0) | (TALLY_EMIT_CONV if emit_conv else 0)
    if n > 0:   # DYN_TALLY_REC16_HINT: mean unit size in 16-byte units (sizes the kernel's staging buffer)
        flags |= min(0xffff, -(-(records.numel() // 16) // n)) << 16
    check(
        ctx.lib.dyn_tally_reads(
            ctx.handle, ptr(records), ptr(rec_off), n, ptr(snps), 0 if snps is None else snps.numel(), int(quality),
            flags, ptr(out.counts), ptr(out.conv_off), ptr(out.conv_n), ptr(out.conv_rec), ptr(out.conv_read),
            out.conv_capacity, C.c_void_p(out.scalars.data_ptr()), C.c_void_p(out.scalars.data_ptr() + 8), _stream()
        )
    )


def em_pc(ctx: Context, k: torch.Tensor, n: torch.Tensor, count: torch.Tensor, off: torch.Tensor, p_e: torch.Tensor,
          nasc: bool = False):
    """dyn_em_pc on device tensors (int32, int32, int64, int64, float64). Returns (p_c, status, iters) tensors."""
    B = off.numel() - 1
    p_c = torch.empty(B, dtype=torch.float64, device=k.device)
    status = torch.empty(B, dtype=torch.int32, device=k.device)
    iters = torch.empty(B, dtype=torch.int32, device=k.device)
    check(
        ctx.lib.dyn_em_pc(ctx.handle, ptr(k), ptr(n), ptr(count), ptr(off), B, ptr(p_e), 1 if nasc else 0, ptr(p_c),
                          ptr(status), ptr(iters), _stream())
    )
    return p_c, status, iters


# --------------------------------------------------------------------------------------------------
# count -> estimate stages after the tally (all asynchronous on the current torch stream)
# --------------------------------------------------------------------------------------------------
CONVERSION_COLUMNS = ['AC', 'AG', 'AT', 'CA', 'CG', 'CT', 'GA', 'GC', 'GT', 'TA', 'TC', 'TG']
BASE_COLUMNS = ['A', 'C', 'G', 'T']


def conv_mask(conversions) -> int:
    """12-bit column mask of an iterable of conversions ('TC', ...); None -> all columns."""
    if conversions is None:
        return 0xfff
    m = 0
    for c in conversions:
        m |= 1 << CONVERSION_COLUMNS.index(c)
    return m


def base_mask(conversions) -> int:
    m = 0
    for c in conversions:
        m |= 1 << BASE_COLUMNS.index(c[0])
    return m


def _bits(max_value: int) -> int:
    return max(1, int(max_value).bit_length())


def count_records(ctx: Context, conv_rec, conv_read, conv_contig, counts, quality=27, snps=None):
    """dyn_count_records: adds the quality/SNP-filtered mismatch records to counts[:, :12] in place.
    Returns the status tensor (int32 scalar on the device)."""
    status = torch.zeros(1, dtype=torch.int32, device=counts.device)
    check(
        ctx.lib.dyn_count_records(ctx.handle, ptr(conv_rec), ptr(conv_read), ptr(conv_contig), conv_rec.numel(),
                                  counts.shape[0], ptr(snps), 0 if snps is None else snps.numel(), int(quality),
                                  ptr(counts), ptr(status), _stream())
    )
    return status


def complement_counts(ctx: Context, counts: torch.Tensor, strand: torch.Tensor):
    """dyn_complement_counts, in place."""
    check(ctx.lib.dyn_complement_counts(ctx.handle, ptr(counts), ptr(strand), counts.shape[0], _stream()))
    return counts


def barcode_first_pos(ctx: Context, barcode, strand, conv_n, n_barcodes: int):
    dev = barcode.device
    first = torch.empty(n_barcodes, dtype=torch.int64, device=dev)
    n_reads = torch.empty(n_barcodes, dtype=torch.int32, device=dev)
    check(ctx.lib.dyn_barcode_first_pos(ctx.handle, ptr(barcode), ptr(strand), ptr(conv_n), barcode.numel(), n_barcodes,
                                        ptr(first), ptr(n_reads), _stream()))
    return first, n_reads


def split_plan(first_pos: np.ndarray, n_reads: np.ndarray, threshold: int = 50000):
    """Barcode order and split assignment of conversion.py:543-579 (dyn_split_plan: sequential packing on the host,
    O(barcodes)).  Returns (ord_of_barcode, split_of_barcode, n_splits)."""
    B = len(first_pos)
    first = np.ascontiguousarray(first_pos).view(np.uint64)
    cnt = np.ascontiguousarray(n_reads, dtype=np.uint32)
    ord_of = np.zeros(B, dtype=np.uint32)
    split_of = np.zeros(B, dtype=np.uint32)
    n_splits = C.c_int64(1)
    check(_lib.load().dyn_split_plan(ptr(first), ptr(cnt), B, int(min(threshold, 2**62)), ptr(ord_of), ptr(split_of),
                                     C.cast(C.byref(n_splits), C.c_void_p)))
    return ord_of, split_of, int(n_splits.value)


def plan_splits(ctx: Context, barcode, strand, conv_n, n_barcodes: int, split_threshold=50000):
    """Barcode order / split assignment (conversion.py:543-579) of a counts frame on the device.
    Returns (ord_of_barcode, split_of_barcode, n_splits) with the two tables on the device."""
    first, n_reads = barcode_first_pos(ctx, barcode, strand, conv_n, n_barcodes)
    ord_of, split_of, n_splits = split_plan(first.cpu().numpy(), n_reads.cpu().numpy(), split_threshold)
    dev = barcode.device
    return (torch.from_numpy(ord_of.view(np.int32)).to(dev), torch.from_numpy(split_of.view(np.int32)).to(dev), n_splits)


def dedup(ctx: Context, key_lo, counts, strand, transcriptome, score, conv_n, barcode, n_barcodes, conversions=None,
          use_conversions=True, key_hi=None, key_lo_bits=64, key_hi_bits=0, final_extra=None, final_extra_bits=0,
          split_threshold=50000, plan=None):
    """a5 + a6 on device tensors. Returns the surviving read indices (int32 tensor) in output row order.
    `plan` = plan_splits(...) of the frame the splits were cut from (defaults to this frame)."""
    n = counts.shape[0]
    dev = counts.device
    if barcode is None:   # one split, input order is the tie-break (conversion.deduplicate_counts on a frame)
        ord_of_d = split_of_d = None
        n_splits, n_barcodes = 1, 0
    else:
        ord_of_d, split_of_d, n_splits = plan if plan is not None else plan_splits(ctx, barcode, strand, conv_n, n_barcodes,
                                                                                   split_threshold)
    out = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    n_out = torch.zeros(1, dtype=torch.int64, device=dev)
    use = 1 if (use_conversions and conversions is not None) else 0
    check(
        ctx.lib.dyn_dedup_umi(ctx.handle, ptr(key_hi), ptr(key_lo), int(key_hi_bits), int(key_lo_bits), ptr(counts),
                              ptr(strand), ptr(transcriptome), ptr(score), ptr(conv_n), ptr(barcode), ptr(ord_of_d),
                              ptr(split_of_d), n_barcodes, n_splits, n, conv_mask(conversions), use, ptr(final_extra),
                              int(final_extra_bits), ptr(out), ptr(n_out), _stream())
    )
    return out[:int(n_out.item())]


def group_sums(ctx: Context, counts, strand, group, n_groups: int, sel=None, conversions=None):
    """dyn_group_sums -> (sums [G,16] int64, n_reads [G] int64, n_with [G] int64)."""
    dev = counts.device
    sums = torch.empty((n_groups, 16), dtype=torch.int64, device=dev)
    n_reads = torch.empty(n_groups, dtype=torch.int64, device=dev)
    n_with = torch.empty(n_groups, dtype=torch.int64, device=dev)
    n_sel = counts.shape[0] if sel is None else sel.numel()
    check(ctx.lib.dyn_group_sums(ctx.handle, ptr(counts), ptr(strand), ptr(group), ptr(sel), n_sel, n_groups,
                                 conv_mask(conversions), ptr(sums), ptr(n_reads), ptr(n_with), _stream()))
    return sums, n_reads, n_with


def rates_pe(ctx: Context, sums, pe_conversions=None):
    """dyn_rates_pe -> (rates [G,12] float64, p_e [G] float64)."""
    G = sums.shape[0]
    rates = torch.empty((G, 12), dtype=torch.float64, device=sums.device)
    p_e = torch.empty(G, dtype=torch.float64, device=sums.device)
    mask = 0 if pe_conversions is None else conv_mask(pe_conversions)
    check(ctx.lib.dyn_rates_pe(ctx.handle, ptr(sums), G, mask, ptr(rates), ptr(p_e), _stream()))
    return rates, p_e


def alpha(ctx: Context, n_reads, n_with, pi_c):
    out = torch.empty(n_reads.numel(), dtype=torch.float64, device=n_reads.device)
    check(ctx.lib.dyn_alpha(ctx.handle, ptr(n_reads), ptr(n_with), ptr(pi_c), n_reads.numel(), ptr(out), _stream()))
    return out


def aggregate_kn(ctx: Context, counts, strand, group, gene, n_groups: int, n_genes: int, conversions, sel=None,
                 exclude=None, first_appearance=True):
    """dyn_aggregate_kn -> (group, gene, k, n, count, first) int64 tensors, one row per distinct key."""
    dev = counts.device
    n_sel = counts.shape[0] if sel is None else sel.numel()
    gene_bits = _bits(max(n_genes - 1, 0)) if gene is not None else 0
    group_bits = _bits(max(n_groups - 1, 0))
    keys = torch.empty(max(n_sel, 1), dtype=torch.int64, device=dev)
    cnt = torch.empty(max(n_sel, 1), dtype=torch.int32, device=dev)
    first = torch.empty(max(n_sel, 1), dtype=torch.int32, device=dev)
    n_rows = torch.zeros(1, dtype=torch.int64, device=dev)
    check(
        ctx.lib.dyn_aggregate_kn(ctx.handle, ptr(counts), ptr(strand), ptr(group), ptr(gene), ptr(sel), n_sel,
                                 conv_mask(conversions), base_mask(conversions),
                                 0 if not exclude else conv_mask(exclude), group_bits, gene_bits,
                                 1 if first_appearance else 0, ptr(keys), ptr(cnt), ptr(first), ptr(n_rows), _stream())
    )
    m = int(n_rows.item())
    k64 = keys[:m]
    return dict(
        group=(k64 >> (22 + gene_bits)), gene=((k64 >> 22) & ((1 << gene_bits) - 1)) if gene_bits else torch.zeros_like(k64),
        k=(k64 >> 10) & 0xfff, n=k64 & 0x3ff, count=cnt[:m].to(torch.int64), first=first[:m].to(torch.int64),
        keys=k64, count32=cnt[:m], n_rows=n_rows,
    )


def kn_csr(ctx: Context, keys, count32, n_rows, n_groups: int):
    """dyn_kn_csr on key-sorted rows without gene bits -> (k, n, count, off, total)."""
    dev = keys.device
    m = keys.numel()
    k = torch.empty(max(m, 1), dtype=torch.int32, device=dev)
    n = torch.empty(max(m, 1), dtype=torch.int32, device=dev)
    c = torch.empty(max(m, 1), dtype=torch.int64, device=dev)
    off = torch.empty(n_groups + 1, dtype=torch.int64, device=dev)
    total = torch.empty(max(n_groups, 1), dtype=torch.int64, device=dev)
    check(ctx.lib.dyn_kn_csr(ctx.handle, ptr(keys), ptr(count32), ptr(n_rows), m, n_groups, ptr(k), ptr(n), ptr(c),
                             ptr(off), ptr(total), _stream()))
    return k[:m], n[:m], c[:m], off, total[:n_groups]


def kn_csr_pairs(ctx: Context, keys, count32, n_rows):
    """dyn_kn_csr_pairs on key-sorted rows WITH gene bits -> (k, n, count, off [P + 1], pair_key [P], total [P]): one CSR
    row per (group, gene) pair that occurs -- the groups of the per-(cell, gene) fit (pi.py:253-296)."""
    dev = keys.device
    m = keys.numel()
    k = torch.empty(max(m, 1), dtype=torch.int32, device=dev)
    n = torch.empty(max(m, 1), dtype=torch.int32, device=dev)
    c = torch.empty(max(m, 1), dtype=torch.int64, device=dev)
    off = torch.zeros(m + 1, dtype=torch.int64, device=dev)
    pair_key = torch.empty(max(m, 1), dtype=torch.int64, device=dev)
    total = torch.empty(max(m, 1), dtype=torch.int64, device=dev)
    n_pairs = torch.zeros(1, dtype=torch.int64, device=dev)
    check(ctx.lib.dyn_kn_csr_pairs(ctx.handle, ptr(keys), ptr(count32), ptr(n_rows), m, ptr(k), ptr(n), ptr(c), ptr(off), ptr(pair_key),
                                   ptr(total), ptr(n_pairs), _stream()))
    P = int(n_pairs.item())
    return k[:m], n[:m], c[:m], off[:P + 1], pair_key[:P], total[:P]


READ_ROW_BYTES = 40


VELOCITY_NAMES = ('unassigned', 'spliced', 'unspliced', 'ambiguous')
STRAND_MODES = {'forward': 0, 'reverse': 1, 'unstranded': 2}


def velocity(ctx: Context, records: torch.Tensor, rec_off: torch.Tensor, tables, gx: Optional[torch.Tensor] = None,
             strand: str = 'forward', strict_exon_overlap: bool = False):
    """dyn_velocity: (gene row int32 [n], type uint8 [n] -- 0 dropped, 1 spliced, 2 unspliced, 3 ambiguous -- and the
    status word) of every unit.  `tables` = gtf.AnnotationTables; gx = gene row per unit, -1 where the read carries
    no gene tag."""
    from ._lib import Annotation
    n = rec_off.numel() - 1
    dev = records.device
    d = tables.to_device(dev)
    ann = Annotation(n_ref=len(tables.contig_off) - 1, n_genes=tables.n_genes,
                     **{'d_' + k: d[k].data_ptr() for k in tables.ARRAYS})
    gene = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    kind = torch.empty(max(n, 1), dtype=torch.uint8, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    check(ctx.lib.dyn_velocity(ctx.handle, ptr(records), ptr(rec_off), n, ptr(gx), C.byref(ann), STRAND_MODES[strand],
                               int(bool(strict_exon_overlap)), ptr(gene), ptr(kind), ptr(status), _stream()))
    return gene[:n], kind[:n], status


def owner_partition(ctx: Context, barcode: torch.Tensor, world: int):
    """dyn_owner_partition -> (perm int32 [n] (bit pattern uint32): local reads grouped by barcode % world, stable;
    counts int64 [world])."""
    n = barcode.numel()
    perm = torch.empty(max(n, 1), dtype=torch.int32, device=barcode.device)
    counts = torch.empty(world, dtype=torch.int64, device=barcode.device)
    check(ctx.lib.dyn_owner_partition(ctx.handle, ptr(barcode), n, int(world), ptr(perm), ptr(counts), _stream()))
    return perm[:n], counts


def pack_reads(ctx: Context, perm, counts, barcode, umi, gene, score16, conv_n, strand, transcriptome=None):
    """dyn_pack_reads -> uint8 [n, 40]: one row per read in `perm` order (perm None: as is)."""
    n = barcode.numel()
    rows = torch.empty((n, READ_ROW_BYTES), dtype=torch.uint8, device=barcode.device)
    check(ctx.lib.dyn_pack_reads(ctx.handle, ptr(perm), n, ptr(counts), ptr(barcode), ptr(umi), ptr(gene), ptr(score16),
                                 ptr(conv_n), ptr(strand), ptr(transcriptome), ptr(rows), _stream()))
    return rows


class DevicePointer:
    """A raw device allocation dressed up for ptr(): what dyn_peer_alloc returns (not a torch tensor)."""

    def __init__(self, address: int, shape, device):
        self.address, self.shape, self.device = int(address), tuple(shape), device

    def data_ptr(self) -> int:
        return self.address


def unpack_reads(ctx: Context, rows):
    """dyn_unpack_reads -> dict(counts uint8 [n,16], barcode int32, umi int64, gene int32, score int16, conv_n int16
    (bit pattern uint16), strand uint8, transcriptome uint8)."""
    n = rows.shape[0]
    dev = rows.device
    m = max(n, 1)
    out = dict(counts=torch.empty((m, 16), dtype=torch.uint8, device=dev), barcode=torch.empty(m, dtype=torch.int32, device=dev),
               umi=torch.empty(m, dtype=torch.int64, device=dev), gene=torch.empty(m, dtype=torch.int32, device=dev),
               score=torch.empty(m, dtype=torch.int16, device=dev), conv_n=torch.empty(m, dtype=torch.int16, device=dev),
               strand=torch.empty(m, dtype=torch.uint8, device=dev), transcriptome=torch.empty(m, dtype=torch.uint8, device=dev))
    check(ctx.lib.dyn_unpack_reads(ctx.handle, ptr(rows), n, ptr(out['counts']), ptr(out['barcode']), ptr(out['umi']),
                                   ptr(out['gene']), ptr(out['score']), ptr(out['conv_n']), ptr(out['strand']),
                                   ptr(out['transcriptome']), _stream()))
    return {k: v[:n] for k, v in out.items()}


def merge_kn(ctx: Context, keys: torch.Tensor, count32: torch.Tensor, world: int = 1, group_bits: int = 0):
    """dyn_merge_kn: weighted (group << 22 | k << 10 | n) rows, possibly repeated -> (keys int64 [m] sorted, count int32 [m],
    n_rows int64 [1] device scalar); group ids are divided by `world`.  The outputs have room for every input row; the
    first n_rows[0] are valid."""
    m = keys.numel()
    out_k = torch.empty(max(m, 1), dtype=torch.int64, device=keys.device)
    out_c = torch.empty(max(m, 1), dtype=torch.int32, device=keys.device)
    n_rows = torch.zeros(1, dtype=torch.int64, device=keys.device)
    check(ctx.lib.dyn_merge_kn(ctx.handle, ptr(keys), ptr(count32), m, int(world), int(group_bits), ptr(out_k), ptr(out_c),
                               ptr(n_rows), _stream()))
    return out_k[:m], out_c[:m], n_rows


def pi_fit(ctx: Context, k, n, count, off, p_e, p_c, max_cells=None):
    """dyn_pi_fit -> (alpha_beta_pi [G,3] float64, status [G] int32)."""
    G = off.numel() - 1
    dev = off.device
    out = torch.empty((G, 3), dtype=torch.float64, device=dev)
    status = torch.empty(G, dtype=torch.int32, device=dev)
    if max_cells is None:
        max_cells = int((off[1:] - off[:-1]).max().item()) if G else 1
    check(ctx.lib.dyn_pi_fit(ctx.handle, ptr(k), ptr(n), ptr(count), ptr(off), G, int(max_cells), ptr(p_e), ptr(p_c),
                             ptr(out), ptr(status), _stream()))
    return out, status


def consensus(ctx: Context, records: torch.Tensor, item_rec16: torch.Tensor, group_off: torch.Tensor, quality: int = 27,
              out_capacity_bytes: Optional[int] = None):
    """dyn_consensus -> dict(records uint8 tensor, off int32 [G+1] (16-byte units), md uint8 tensor (MD text),
    md_off int32 [G+1] (bytes), nm int32 [G], status int32 [G])."""
    G = group_off.numel() - 1
    dev = records.device
    if out_capacity_bytes is None:   # a consensus record is never larger than its members together (+ header slack)
        out_capacity_bytes = int(records.numel()) + 64 * G + 1024
    out = torch.empty(out_capacity_bytes, dtype=torch.uint8, device=dev)
    md = torch.empty(out_capacity_bytes, dtype=torch.uint8, device=dev)
    off = torch.zeros(G + 1, dtype=torch.int32, device=dev)
    md_off = torch.zeros(G + 1, dtype=torch.int32, device=dev)
    nm = torch.zeros(max(G, 1), dtype=torch.int32, device=dev)
    status = torch.zeros(max(G, 1), dtype=torch.int32, device=dev)
    check(ctx.lib.dyn_consensus(ctx.handle, ptr(records), ptr(item_rec16), ptr(group_off), G, item_rec16.numel(), int(quality),
                                ptr(out), out_capacity_bytes, ptr(off), ptr(md), out_capacity_bytes, ptr(md_off), ptr(nm),
                                ptr(status), _stream()))
    return dict(records=out, off=off, md=md, md_off=md_off, nm=nm[:G], status=status[:G])


# --------------------------------------------------------------------------------------------------
# The whole device-resident hot path for one read range: tally -> dedup -> p_e -> (k, n) histograms -> EM -> pi
# --------------------------------------------------------------------------------------------------
def count_to_estimate(ctx: Context, records, rec_off, barcode, umi, gene, score, strand, n_barcodes: int, n_genes: int,
                      out: TallyBuffers, conversions=('TC',), quality: int = 27, umi_bits: int = 24,
                      cell_threshold: int = 1000, with_pi: bool = True, timings: Optional[dict] = None, group=None,
                      exchange: bool = False, tallied=None, transcriptome=None, method: str = 'alpha',
                      cell_gene_threshold: int = 16, conversion_groups=None):
    """Runs every stage of the path on device tensors (the packed records of one GPU's read range plus the
    interned barcode / umi / gene ids of its reads).  Mirrors count() then estimate(method='alpha') of the
    reference: count_conversions (count.py:306) -> estimate_p_e (estimate.py:197-234) -> aggregate_counts
    (:243-255) -> estimate_p_c (:259-275) -> estimate_pi per barcode + estimate_alpha (:328-395).
    Returns a dict of device tensors.  `timings`, if given, receives per-stage milliseconds (CUDA events).
    exchange=True (multi-GPU, read-range sharding): `barcode` holds GLOBAL barcode ids and n_barcodes the global
    count; after the tally every read's row goes to rank barcode % world (distributed.exchange_reads) and the rest
    of the path runs on the barcodes this rank owns (local id = global id // world).
    method = 'pi_g' adds the per-(cell, gene) fit of estimate(method='pi_g') (estimate.py:282-327, pi.py:253-296): every
    (barcode, gene) pair with at least `cell_gene_threshold` reads gets its own (alpha, beta, pi) from the barcode's p_e /
    p_c -- result keys `pi_g` [P, 3], `pi_g_status`, `pi_g_pair` (barcode << gene_bits | gene), `pi_g_reads`.
    conversion_groups (dual labelling, e.g. [('TC',), ('AG',)] for `--conversion TC --conversion AG`): `conversions` is then
    the flattened set (dedup priority and the p_e columns use all of them, count.py:306, p_e.py:82-89) and histograms, EM,
    pi and alpha run once per group on the reads WITHOUT any conversion of the other groups (estimate.py:248-255); the
    per-group results sit in res['groups'][name] and the top-level keys hold the first group's.
    tallied = (counts [n, 16] uint8, conv_n [n] int16): the tally already ran (tally_bam streams it chunk by chunk while
    the BAM is being read) and records / rec_off / out are not used; transcriptome = the reads' gene-tag-present flag."""
    dev = barcode.device
    stream = torch.cuda.current_stream()
    marks = []

    def mark(name):
        if timings is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record(stream)
            marks.append((name, e))

    mark('start')
    if tallied is None:
        tally_reads(ctx, records, rec_off, out, quality=quality, emit_conv=True)
        counts, conv_n = out.counts, out.conv_n
    else:
        counts, conv_n = tallied
    mark('tally')
    score16 = score.to(torch.int16)
    if exchange:
        from . import distributed
        import torch.distributed as dist
        world = dist.get_world_size(group)
        got = distributed.exchange_reads(ctx, barcode, umi, gene, score16, conv_n, strand, counts, world, group=group,
                                         transcriptome=transcriptome)
        counts, conv_n, score16, strand, umi, gene = got['counts'], got['conv_n'], got['score'], got['strand'], got['umi'], got['gene']
        transcriptome = got['transcriptome']
        barcode = torch.div(got['barcode'], world, rounding_mode='floor').to(torch.int32)
        n_barcodes = (n_barcodes + world - 1) // world
        mark('exchange')
    gene_bits = _bits(max(n_genes - 1, 0))
    bc_bits = _bits(max(n_barcodes - 1, 0))
    n = barcode.numel()
    key = torch.empty(n, dtype=torch.int64, device=dev)
    check(ctx.lib.dyn_group_key(ctx.handle, ptr(barcode), ptr(umi), ptr(gene), n, int(umi_bits), int(gene_bits), ptr(key), _stream()))
    if transcriptome is None:
        transcriptome = torch.ones(n, dtype=torch.uint8, device=dev)
    sel = dedup(ctx, key, counts, strand, transcriptome, score16, conv_n, barcode, n_barcodes,
                conversions=conversions, use_conversions=True, key_lo_bits=bc_bits + umi_bits + gene_bits)
    mark('dedup')
    pe_cols = [c for c in CONVERSION_COLUMNS if c[0] not in {x[0] for x in conversions}]
    if not pe_cols:      # the labelled conversions start with all four bases (p_e.py:84-91): every non-induced column
        pe_cols = [c for c in CONVERSION_COLUMNS if c not in set(conversions)]
    group_list = [tuple(g) for g in conversion_groups] if conversion_groups else [tuple(conversions)]
    all_convs = tuple(conversions)
    first = group_list[0]
    # n_with (alpha's "new" reads, alpha.py:75) counts reads with a conversion of THIS group; the sums are all 16 columns
    sums, n_reads, n_with = group_sums(ctx, counts, strand, barcode, n_barcodes, sel=sel, conversions=first)
    rates, p_e = rates_pe(ctx, sums, pe_conversions=pe_cols)
    mark('p_e')
    others = tuple(c for c in all_convs if c not in first)
    rows = aggregate_kn(ctx, counts, strand, barcode, None, n_barcodes, 0, first, sel=sel, first_appearance=False,
                        exclude=others if conversion_groups else None)
    k, nn, c, off, total = kn_csr(ctx, rows['keys'], rows['count32'], rows['n_rows'], n_barcodes)
    mark('aggregate')
    # estimate_p_c skips barcodes below the read threshold (p_c.py:215-217): give them an empty histogram
    lens = off[1:] - off[:-1]
    max_len = int(lens.max().item()) if lens.numel() else 1
    keep = total >= cell_threshold
    off_kept = torch.zeros_like(off)
    off_kept[1:] = torch.cumsum(torch.where(keep, lens, torch.zeros_like(lens)), 0)
    row_keep = torch.repeat_interleave(keep, lens)
    p_c, em_status, iters = em_pc(ctx, k[row_keep].contiguous(), nn[row_keep].contiguous(), c[row_keep].contiguous(), off_kept, p_e)
    em_status = torch.where(keep, em_status, torch.full_like(em_status, -1))
    mark('em')
    res = dict(selected=sel, sums=sums, rates=rates, p_e=p_e, p_c=p_c, em_status=em_status, em_iters=iters, n_reads=n_reads,
               n_with=n_with, hist=(k, nn, c, off), n_local_reads=n)
    if with_pi:
        ok = em_status == 0
        # barcodes without a usable p_c keep their rows but get p_c = NaN: dyn_pi_fit reports them at once
        # (status 2) and nothing has to be compacted
        pc_in = torch.where(ok, p_c, torch.full_like(p_c, float('nan')))
        pe_safe = torch.where(torch.isnan(p_e), torch.zeros_like(p_e), p_e)
        abp, pi_status = pi_fit(ctx, k, nn, c, off, pe_safe, pc_in, max_cells=max_len)
        pi_c = torch.where(ok & (pi_status == 0), abp[:, 2], torch.full_like(p_c, float('nan')))
        res.update(pi=abp, pi_status=pi_status, alpha=alpha(ctx, n_reads, n_with, pi_c))
        mark('pi')
    if conversion_groups:
        name = lambda g: '_'.join(sorted(g))
        res['groups'] = {name(first): {k_: res[k_] for k_ in ('p_c', 'em_status', 'hist', 'pi', 'pi_status', 'alpha') if k_ in res}}
        for g in group_list[1:]:
            og = tuple(c_ for c_ in all_convs if c_ not in g)
            r2 = aggregate_kn(ctx, counts, strand, barcode, None, n_barcodes, 0, g, sel=sel, first_appearance=False, exclude=og)
            k2, n2, c2, off2, total2 = kn_csr(ctx, r2['keys'], r2['count32'], r2['n_rows'], n_barcodes)
            lens2 = off2[1:] - off2[:-1]
            keep2 = total2 >= cell_threshold
            offk = torch.zeros_like(off2)
            offk[1:] = torch.cumsum(torch.where(keep2, lens2, torch.zeros_like(lens2)), 0)
            rk = torch.repeat_interleave(keep2, lens2)
            pc2, st2, _ = em_pc(ctx, k2[rk].contiguous(), n2[rk].contiguous(), c2[rk].contiguous(), offk, p_e)
            st2 = torch.where(keep2, st2, torch.full_like(st2, -1))
            entry = dict(p_c=pc2, em_status=st2, hist=(k2, n2, c2, off2))
            if with_pi:
                ok2 = st2 == 0
                abp2, ps2 = pi_fit(ctx, k2, n2, c2, off2, torch.nan_to_num(p_e, nan=0.0), torch.where(ok2, pc2, torch.full_like(pc2, float('nan'))),
                                   max_cells=int(lens2.max().item()) if lens2.numel() else 1)
                _, _, nw2 = group_sums(ctx, counts, strand, barcode, n_barcodes, sel=sel, conversions=g)
                pic2 = torch.where(ok2 & (ps2 == 0), abp2[:, 2], torch.full_like(pc2, float('nan')))
                entry.update(pi=abp2, pi_status=ps2, alpha=alpha(ctx, n_reads, nw2, pic2))
            res['groups'][name(g)] = entry
        mark('other_groups')
    if method == 'pi_g':
        rows_g = aggregate_kn(ctx, counts, strand, barcode, gene, n_barcodes, n_genes, conversions, sel=sel, first_appearance=False)
        gk, gn, gc, goff, pair_key, pair_total = kn_csr_pairs(ctx, rows_g['keys'], rows_g['count32'], rows_g['n_rows'])
        bc_of_pair = pair_key >> gene_bits
        fit = (em_status[bc_of_pair] == 0) & (pair_total >= cell_gene_threshold)
        pc_g = torch.where(fit, p_c[bc_of_pair], torch.full_like(p_c[bc_of_pair], float('nan')))
        pe_g = torch.nan_to_num(p_e[bc_of_pair], nan=0.0)
        glens = goff[1:] - goff[:-1]
        g_max = int(glens.max().item()) if glens.numel() else 1
        abp_g, st_g = pi_fit(ctx, gk, gn, gc, goff, pe_g, pc_g, max_cells=g_max)
        st_g = torch.where(fit, st_g, torch.full_like(st_g, -1))          # -1: below the threshold / no p_c for the barcode
        res.update(pi_g=abp_g, pi_g_status=st_g, pi_g_pair=pair_key, pi_g_reads=pair_total, pi_g_hist=(gk, gn, gc, goff))
        mark('pi_g')
    if timings is not None:
        stream.synchronize()
        for (_, a), (name, b) in zip(marks[:-1], marks[1:]):
            timings[name] = timings.get(name, 0.0) + a.elapsed_time(b)
    return res


# --------------------------------------------------------------------------------------------------
# From a BAM file: streamed ingest (device or host) -> tally per chunk -> the path above
# --------------------------------------------------------------------------------------------------
def tally_bam(ctx: Context, path: str, gene_ids, barcode_tag: Optional[str] = 'CB', umi_tag: Optional[str] = 'UB',
              gene_tag: str = 'GX', require_gene: bool = True, quality: int = 27, nasc: bool = False, part: int = 0,
              n_parts: int = 1, chunk_bytes: int = 0, ingest: str = 'device', n_threads: int = 8, snps=None,
              stats: Optional[dict] = None):
    """Streams block range `part` of `n_parts` of a coordinate-sorted BAM through the tally: every chunk of the reader
    (bamio.DeviceBamStream: inflate + pack on the GPU; bamio.BamStream: on the host cores, then cudaMemcpyAsync from
    page-locked memory) is tallied as it arrives while the next one is being read.  ingest = 'device' | 'host' | 'auto'
    (device, falling back to the host reader for files the device path refuses: paired records, records straddling
    BGZF blocks).  Returns the per-read columns of the whole range on the device: counts [n, 16] uint8, conv_n int16,
    barcode_key / umi_key int64 (2-bit packed), gene_idx int32 (index into gene_ids, -1 no tag, -2 unknown), score int32,
    transcriptome uint8 (gene tag present), status (int, DYN_ST_* bits)."""
    from . import bamio
    dev = torch.device('cuda', ctx.device)
    kw = dict(barcode_tag=barcode_tag, umi_tag=umi_tag, gene_tag=gene_tag, require_gene=require_gene, part=part, n_parts=n_parts,
              chunk_bytes=chunk_bytes)
    cols = {k: [] for k in ('counts', 'conv_n', 'barcode_key', 'umi_key', 'gene_idx', 'score', 'transcriptome')}
    status = torch.zeros(1, dtype=torch.int64, device=dev)
    seen = units = compressed = inflated = 0

    def tally_chunk(records, rec_off, n):
        out = TallyBuffers(n, 2 * n + 4096, device=ctx.device)
        tally_reads(ctx, records, rec_off, out, quality=quality, nasc=nasc, emit_conv=True, snps=snps)
        status.bitwise_or_(out.scalars[1:2])
        cols['counts'].append(out.counts)
        cols['conv_n'].append(out.conv_n)

    def run_device():
        nonlocal seen, units, compressed, inflated
        with bamio.DeviceBamStream(ctx, path, lean=True, gene_ids=list(gene_ids), n_threads=n_threads, **kw) as st:
            for ch in st:
                seen += ch['n_records_seen']; compressed += ch['compressed_bytes']; inflated += ch['inflated_bytes']
                n = ch['n_units']
                if not n:
                    continue
                units += n
                tally_chunk(ch['records'], ch['rec_off'], n)
                # the chunk's buffers are reused two chunks later: keep copies of the (small) key columns
                cols['barcode_key'].append(ch['barcode_key'].clone()); cols['umi_key'].append(ch['umi_key'].clone())
                cols['gene_idx'].append(ch['gene_idx'].clone()); cols['score'].append(ch['score'].clone())
                cols['transcriptome'].append(ch['gx_assigned'].clone())

    def run_host():
        nonlocal seen, units
        pending = []        # (chunk, event): a chunk's page-locked buffers are released once its copies have run
        with bamio.BamStream(path, keys=True, lean=True, gene_ids=list(gene_ids), n_threads=n_threads, **kw) as st:
            for ch in st:
                seen += ch.n_records_seen
                n = ch.n_units
                if n:
                    units += n
                    h2d = lambda a: torch.from_numpy(a).to(dev, non_blocking=True)
                    records = h2d(ch.blob)
                    rec_off = torch.from_numpy(ch.rec_off.view(np.int32)).to(dev)
                    cols['barcode_key'].append(torch.from_numpy(ch.barcode_key.view(np.int64)).to(dev))
                    cols['umi_key'].append(torch.from_numpy(ch.umi_key.view(np.int64)).to(dev))
                    cols['gene_idx'].append(torch.from_numpy(ch.gene_idx).to(dev)); cols['score'].append(torch.from_numpy(ch.score).to(dev))
                    cols['transcriptome'].append(torch.from_numpy(ch.gx_assigned).to(dev))
                    tally_chunk(records, rec_off, n)
                ev = torch.cuda.Event()
                ev.record()
                pending.append((ch, ev))
                while len(pending) > 2:
                    old, oev = pending.pop(0)
                    oev.synchronize()
                    old.close()
            for old, oev in pending:
                oev.synchronize()
                old.close()

    if ingest in ('device', 'auto'):
        try:
            run_device()
        except _lib.DynastB200Error as exc:
            if ingest == 'device' or not any(w in str(exc) for w in ('straddle', 'paired')):
                raise
            for v in cols.values():
                v.clear()
            seen = units = compressed = inflated = 0
            run_host()
            ingest = 'host (the device path refused the file: %s)' % str(exc).split(': ', 2)[-1][:60]
    else:
        run_host()
    empty = dict(counts=torch.empty((0, 16), dtype=torch.uint8, device=dev), conv_n=torch.empty(0, dtype=torch.int16, device=dev),
                 barcode_key=torch.empty(0, dtype=torch.int64, device=dev), umi_key=torch.empty(0, dtype=torch.int64, device=dev),
                 gene_idx=torch.empty(0, dtype=torch.int32, device=dev), score=torch.empty(0, dtype=torch.int32, device=dev),
                 transcriptome=torch.empty(0, dtype=torch.uint8, device=dev))
    res = {k: (torch.cat(v) if len(v) > 1 else (v[0] if v else empty[k])) for k, v in cols.items()}
    res['status'] = int(status.item()) & 0xffffffff
    if stats is not None:
        stats.update(records_seen=seen, units=units, compressed_bytes=compressed, inflated_bytes=inflated, ingest=ingest)
    return res


def count_from_bam(ctx: Context, path: str, gene_ids, gene_strand, conversions=('TC',), group=None, timings: Optional[dict] = None,
                   stats: Optional[dict] = None, cell_threshold: int = 1000, with_pi: bool = True, **bam_kw):
    """`dynast count` + `dynast estimate` on the device from a BAM file: tally_bam over this rank's block range
    (part = rank, n_parts = world when `group` is given), barcode / UMI keys -> dense ids, then count_to_estimate on the
    tallied reads (with the per-read exchange to the barcode owner when several ranks share the file).
    gene_strand: uint8 per gene of gene_ids (0 '+', 1 '-').  Returns count_to_estimate's dict plus `barcode_keys`
    (the 2-bit key of every dense barcode id; bamio.decode_key turns it back into the string)."""
    world, rank = 1, 0
    if group is not None:
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        bam_kw.update(part=rank, n_parts=world)
    t = tally_bam(ctx, path, gene_ids, stats=stats, **bam_kw)
    dev = t['counts'].device
    if t['status'] & 0x2:
        raise ValueError('malformed alignment record (CIGAR / MD / sequence disagree)')
    if t['gene_idx'].numel() and int(t['gene_idx'].min().item()) < 0:
        raise KeyError('a read carries no gene tag or one that is not in gene_ids (KeyError in complement_counts, conversion.py:116)')
    def global_unique(keys):
        """Sorted distinct values of `keys` over all ranks: dense ids (searchsorted) then agree on every rank."""
        u = torch.unique(keys)
        if world == 1:
            return u
        import torch.distributed as dist
        n_max = torch.tensor([u.numel()], dtype=torch.int64, device=dev)
        dist.all_reduce(n_max, op=dist.ReduceOp.MAX, group=group)
        sentinel = torch.iinfo(torch.int64).max
        pad = torch.full((int(n_max.item()),), sentinel, dtype=torch.int64, device=dev)
        pad[:u.numel()] = u
        every = torch.empty(world * pad.numel(), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(every, pad, group=group)
        u = torch.unique(every)
        return u[u != sentinel]

    bc_keys = global_unique(t['barcode_key'])       # owner of a barcode = its dense id % world
    barcode = torch.searchsorted(bc_keys, t['barcode_key']).to(torch.int32)
    n_barcodes = int(bc_keys.numel())
    umi = t['umi_key']
    umi_max = umi.max().reshape(1) if umi.numel() else torch.zeros(1, dtype=torch.int64, device=dev)
    if world > 1:       # the group key is packed on the OWNER's side of the exchange: every rank must use the same widths
        import torch.distributed as dist
        dist.all_reduce(umi_max, op=dist.ReduceOp.MAX, group=group)
    umi_bits = max(1, int(umi_max.item()).bit_length())
    gene_bits, bc_bits = _bits(max(len(gene_ids) - 1, 0)), _bits(max(n_barcodes - 1, 0))
    if umi_bits + gene_bits + bc_bits > 64:      # long UMIs: dense ids instead of the 2-bit value
        u = global_unique(umi)
        umi, umi_bits = torch.searchsorted(u, umi).to(torch.int64), _bits(max(int(u.numel()) - 1, 0))
    strand_table = torch.as_tensor(np.ascontiguousarray(gene_strand, dtype=np.uint8), device=dev)
    strand = strand_table[t['gene_idx'].long()]
    res = count_to_estimate(ctx, None, None, barcode, umi, t['gene_idx'], t['score'], strand, n_barcodes, len(gene_ids), None,
                            conversions=conversions, umi_bits=umi_bits, cell_threshold=cell_threshold, with_pi=with_pi, timings=timings,
                            group=group, exchange=world > 1, tallied=(t['counts'], t['conv_n']), transcriptome=t['transcriptome'])
    res['barcode_keys'] = bc_keys
    res['n_reads_tallied'] = int(barcode.numel())
    return res
