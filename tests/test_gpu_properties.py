"""GPU: size-independent properties of the device path at BASELINE scale (C2-shaped batches generated on the
device, tens of millions of reads): partition invariance of the tally, complement involution, dedup idempotence
and key uniqueness, checksum-of-checksums between the per-read counters, the per-barcode sums and the (k, n)
histograms, order invariance of the EM, and empty inputs through every entry point."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

N_READS = 20_000_000
N_BARCODES = 4000


@pytest.fixture(scope='module')
def batch(gpu_ctx):
    from dynast_b200 import pipeline
    b = pipeline.SynthBatch(N_READS, seed=424242, n_barcodes=N_BARCODES, device=0)
    out = pipeline.TallyBuffers(N_READS, int(N_READS * 1.2), device=0)
    pipeline.tally_reads(gpu_ctx, b.records, b.rec_off, out, quality=27, emit_conv=True)
    torch.cuda.synchronize()
    assert int(out.scalars[1].item()) == 0
    return b, out


def test_tally_partition_invariance(gpu_ctx, batch):
    """Tallying two halves of the read range separately gives the same counters and the same mismatch records."""
    from dynast_b200 import pipeline
    b, out = batch
    half = N_READS // 2
    off = b.rec_off.to(torch.int64) & 0xffffffff
    for lo, hi in ((0, half), (half, N_READS)):
        part = pipeline.TallyBuffers(hi - lo, int(N_READS * 0.7), device=0)
        pipeline.tally_reads(gpu_ctx, b.records, b.rec_off[lo:hi + 1].contiguous(), part, quality=27, emit_conv=True)
        torch.cuda.synchronize()
        assert torch.equal(part.counts, out.counts[lo:hi])
        assert torch.equal(part.conv_n, out.conv_n[lo:hi])
        assert int(part.scalars[0].item()) == int(out.conv_n[lo:hi].to(torch.int64).sum().item())
    # every mismatch record is owned by the read whose range holds it, and record counts add up
    total = int(out.scalars[0].item())
    assert total == int(out.conv_n.to(torch.int64).sum().item())
    reads = out.conv_read[:total].to(torch.int64)
    assert torch.equal(torch.bincount(reads, minlength=N_READS), out.conv_n.to(torch.int64))
    # content never exceeds the read length + deletions, conversions never exceed the content of their base
    c = out.counts.to(torch.int32)
    assert int(c[:, 12:].sum(dim=1).max().item()) <= 91 + 3
    for j, base in enumerate((0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 3)):
        assert bool((c[:, j] <= c[:, 12 + base]).all())


def test_complement_involution_at_scale(gpu_ctx, batch):
    from dynast_b200 import pipeline
    b, out = batch
    counts = out.counts.clone()
    pipeline.complement_counts(gpu_ctx, counts, b.strand)
    rev = b.strand != 0
    assert torch.equal(counts[~rev], out.counts[~rev])
    assert torch.equal(counts[rev][:, :12], out.counts[rev][:, :12].flip(1))
    assert torch.equal(counts[rev][:, 12:], out.counts[rev][:, 12:].flip(1))
    pipeline.complement_counts(gpu_ctx, counts, b.strand)
    assert torch.equal(counts, out.counts)


def _key(b):
    return (b.barcode.to(torch.int64) << 39) | ((b.umi & 0xffffff) << 15) | b.gene.to(torch.int64)


def test_dedup_properties_and_checksums(gpu_ctx, batch):
    from dynast_b200 import pipeline
    b, out = batch
    key = _key(b)
    ones = torch.ones(N_READS, dtype=torch.uint8, device='cuda')
    sel = pipeline.dedup(gpu_ctx, key, out.counts, b.strand, ones, b.score.to(torch.int16), out.conv_n, b.barcode, N_BARCODES,
                         conversions=('TC',), use_conversions=True, key_lo_bits=53).to(torch.int64)
    # one survivor per (barcode, umi, gene) group, every group represented
    assert sel.numel() == torch.unique(key).numel()
    assert torch.unique(key[sel]).numel() == sel.numel()
    # the survivor maximises the priority inside its group (has TC conversion, then score, then TC count)
    comp = out.counts.clone()
    pipeline.complement_counts(gpu_ctx, comp, b.strand)
    tc = comp[:, 10].to(torch.int64)
    prio = ((tc > 0).to(torch.int64) << 40) | (b.score.to(torch.int64) << 16) | tc
    uk, inv = torch.unique(key, return_inverse=True)
    best = torch.zeros(uk.numel(), dtype=torch.int64, device='cuda').scatter_reduce_(0, inv, prio, reduce='amax', include_self=False)
    assert torch.equal(prio[sel], best[inv[sel]])
    # idempotence: deduplicating the survivors changes nothing
    sub = sel.sort().values
    sel2 = pipeline.dedup(gpu_ctx, key[sub].contiguous(), out.counts[sub].contiguous(), b.strand[sub].contiguous(), ones[sub].contiguous(),
                          b.score[sub].to(torch.int16).contiguous(), out.conv_n[sub].contiguous(), b.barcode[sub].contiguous(),
                          N_BARCODES, conversions=('TC',), use_conversions=True, key_lo_bits=53)
    assert sel2.numel() == sub.numel()
    # checksum of checksums: per-barcode sums == column sums of the survivors' complemented counters ...
    sel32 = sel.to(torch.int32)
    sums, n_reads, n_with = pipeline.group_sums(gpu_ctx, out.counts, b.strand, b.barcode, N_BARCODES, sel=sel32, conversions=('TC',))
    assert torch.equal(sums.sum(dim=0), comp[sel].to(torch.int64).sum(dim=0))
    assert int(n_reads.sum().item()) == sel.numel()
    assert int(n_with.sum().item()) == int((tc[sel] > 0).sum().item())
    # ... and the (k, n) histogram holds every survivor once, with the same k and n totals
    rows = pipeline.aggregate_kn(gpu_ctx, out.counts, b.strand, b.barcode, None, N_BARCODES, 0, ('TC',), sel=sel32, first_appearance=False)
    assert int(rows['count'].sum().item()) == sel.numel()
    assert int((rows['k'] * rows['count']).sum().item()) == int(sums[:, 10].sum().item())
    assert int((rows['n'] * rows['count']).sum().item()) == int(sums[:, 15].sum().item())
    k, n, c, off, total = pipeline.kn_csr(gpu_ctx, rows['keys'], rows['count32'], rows['n_rows'], N_BARCODES)
    assert torch.equal(total, n_reads)
    assert int(off[-1].item()) == rows['keys'].numel() and bool((off[1:] >= off[:-1]).all())
    # rates / p_e from the sums agree with a float64 recomputation
    rates, p_e = pipeline.rates_pe(gpu_ctx, sums, pe_conversions=['AC', 'AG', 'AT', 'CA', 'CG', 'CT', 'GA', 'GC', 'GT'])
    s = sums.to(torch.float64)
    want = torch.stack([s[:, j] / s[:, 12 + j // 3] for j in range(9)], dim=1)
    assert torch.allclose(p_e, torch.nanmean(want, dim=1), rtol=1e-12, equal_nan=True)


def test_em_order_invariance(gpu_ctx):
    """Shuffling a barcode's triplets (and splitting cells into duplicates) does not change p_c bit for bit."""
    import sys, os
    from helpers import synth_em_batch
    from dynast_b200 import device
    k, n, c, off, pe = synth_em_batch(300, seed=99, adversarial_every=50)
    p0, s0, _ = device.em_pc_host(k, n, c, off, pe)
    rng = np.random.default_rng(5)
    k2, n2, c2 = k.copy(), n.copy(), c.copy()
    for b in range(len(off) - 1):
        perm = rng.permutation(off[b + 1] - off[b]) + off[b]
        k2[off[b]:off[b + 1]], n2[off[b]:off[b + 1]], c2[off[b]:off[b + 1]] = k[perm], n[perm], c[perm]
    p1, s1, _ = device.em_pc_host(k2, n2, c2, off, pe)
    assert (s0 == s1).all() and (p0[s0 == 0] == p1[s1 == 0]).all()


def test_empty_inputs(gpu_ctx):
    from dynast_b200 import device, pipeline
    dev = torch.device('cuda', 0)
    e8 = torch.zeros(0, dtype=torch.uint8, device=dev)
    e16 = torch.zeros(0, dtype=torch.int16, device=dev)
    e32 = torch.zeros(0, dtype=torch.int32, device=dev)
    e64 = torch.zeros(0, dtype=torch.int64, device=dev)
    counts = torch.zeros((0, 16), dtype=torch.uint8, device=dev)
    assert pipeline.dedup(gpu_ctx, e64, counts, e8, e8, e16, e16, e32, 0).numel() == 0
    sums, n_reads, n_with = pipeline.group_sums(gpu_ctx, counts, e8, e32, 3)
    assert int(sums.sum().item()) == 0 and n_reads.numel() == 3
    rows = pipeline.aggregate_kn(gpu_ctx, counts, e8, e32, e32, 1, 1, ('TC',))
    assert rows['count'].numel() == 0
    p_c, st, it = device.em_pc_host(np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.int64), np.zeros(1, np.int64),
                                    np.zeros(0))
    assert len(p_c) == 0
    # barcodes without any read: EM reports 'p_c <= p_e', pi reports an empty group
    off = torch.zeros(4, dtype=torch.int64, device=dev)
    p_c, st, it = pipeline.em_pc(gpu_ctx, e32, e32, e64, off, torch.full((3,), 1e-3, dtype=torch.float64, device=dev))
    assert st.tolist() == [1, 1, 1]
    abp, pst = pipeline.pi_fit(gpu_ctx, e32, e32, e64, off, torch.full((3,), 1e-3, dtype=torch.float64, device=dev),
                               torch.full((3,), 5e-2, dtype=torch.float64, device=dev), max_cells=1)
    assert pst.tolist() == [1, 1, 1] and bool(torch.isnan(abp).all())
    out = pipeline.consensus(gpu_ctx, torch.zeros(16, dtype=torch.uint8, device=dev), e32, torch.zeros(1, dtype=torch.int64, device=dev))
    assert out['off'].tolist() == [0]


def test_group_key(gpu_ctx):
    """dyn_group_key == the bit expression it replaces."""
    from dynast_b200.pipeline import check, ptr, _stream
    dev = torch.device('cuda', 0)
    g = torch.Generator().manual_seed(5)
    n = 100003
    bc = torch.randint(0, 9000, (n,), generator=g).to(torch.int32).to(dev)
    umi = torch.randint(0, 1 << 40, (n,), generator=g).to(dev)
    gene = torch.randint(0, 30000, (n,), generator=g).to(torch.int32).to(dev)
    for umi_bits, gene_bits in ((24, 15), (32, 16), (0, 15), (40, 0)):
        key = torch.empty(n, dtype=torch.int64, device=dev)
        check(gpu_ctx.lib.dyn_group_key(gpu_ctx.handle, ptr(bc), ptr(umi), ptr(gene), n, umi_bits, gene_bits, ptr(key), _stream()))
        want = (bc.to(torch.int64) << (umi_bits + gene_bits)) | ((umi & ((1 << umi_bits) - 1)) << gene_bits) | gene.to(torch.int64)
        assert torch.equal(key, want)


@pytest.mark.parametrize('n,n_barcodes,n_umis', [(1, 1, 1), (300_000, 37, 900), (200_000, 3, 20)])
def test_dedup_hash_path_equals_the_stable_sort_definition(gpu_ctx, n, n_barcodes, n_umis):
    """dyn_dedup_umi's atomicMax winner selection against the numpy restatement of the reference's stable sorts
    (tests/fake_device.py::dedup: conversion.py:225-243, 541-606) on tie-heavy groups: identical survivor list, in order."""
    import fake_device
    from dynast_b200 import pipeline
    g = torch.Generator().manual_seed(n + n_barcodes)
    dev = torch.device('cuda', 0)
    barcode = torch.randint(0, n_barcodes, (n,), generator=g, dtype=torch.int32)
    umi = torch.randint(0, n_umis, (n,), generator=g, dtype=torch.int64)
    gene = torch.randint(0, 50, (n,), generator=g, dtype=torch.int32)
    strand = (gene % 2).to(torch.uint8)
    counts = torch.randint(0, 3, (n, 16), generator=g, dtype=torch.uint8)
    counts[:, :12] = counts[:, :12] * (torch.rand((n, 12), generator=g) < 0.15)
    score = torch.randint(50, 54, (n,), generator=g, dtype=torch.int16)          # few distinct scores: many priority ties
    transcriptome = (torch.rand(n, generator=g) < 0.7).to(torch.uint8)
    conv_n = counts[:, :12].sum(dim=1).to(torch.int16)
    key = (barcode.to(torch.int64) << 40) | (umi << 8) | gene.to(torch.int64)
    for use_conv in (True, False):
        want = fake_device.dedup(None, key, counts, strand, transcriptome, score, conv_n, barcode, n_barcodes, conversions=('TC',),
                                 use_conversions=use_conv, key_lo_bits=48, split_threshold=20000)
        got = pipeline.dedup(gpu_ctx, key.to(dev), counts.to(dev), strand.to(dev), transcriptome.to(dev), score.to(dev), conv_n.to(dev),
                             barcode.to(dev), n_barcodes, conversions=('TC',), use_conversions=use_conv, key_lo_bits=48, split_threshold=20000)
        torch.cuda.synchronize()
        assert torch.equal(got.cpu().to(torch.int64), want.to(torch.int64))
