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
                 barcode_base: int = 0, with_keys: bool = True):
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
        sp.pos_span = 1 << 27
        sp.p_c, sp.p_e, sp.frac_n = p_c, p_e, frac_n
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
        """Unpadded record bytes (16 + 4 n_cigar + ceil(l/2) + l + md_len per read) + 4-byte offsets."""
        off = (self.rec_off[:-1].to(torch.int64) & 0xffffffff) * 16
        total = 0
        step = 1 << 24
        for i in range(0, self.n_reads, step):
            o = off[i:i + step]
            lo = self.records[o + 8].to(torch.int64) | (self.records[o + 9].to(torch.int64) << 8)
            nc = self.records[o + 10].to(torch.int64) | (self.records[o + 11].to(torch.int64) << 8)
            ml = self.records[o + 12].to(torch.int64) | (self.records[o + 13].to(torch.int64) << 8)
            total += int((16 + 4 * nc + (lo + 1) // 2 + lo + ml).sum().item())
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
