"""GPU: the per-read exchange before UMI dedup.  Kernels against numpy on one GPU; the whole
multi-GPU path (read-range sharding -> tally -> all-to-all to the barcode owner -> dedup -> ... -> EM) against the
single-GPU path on the same reads when two GPUs are present."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('n,world', [(0, 2), (1, 1), (1000, 3), (300000, 8), (70001, 64)])
def test_partition_pack_unpack(gpu_ctx, n, world):
    import fake_device
    from dynast_b200 import pipeline
    from dynast_b200.pipeline import Context
    ctx = Context.get(0)
    g = torch.Generator().manual_seed(n + world)
    cols = dict(barcode=torch.randint(0, 5000, (n,), generator=g).to(torch.int32),
                umi=torch.randint(0, 1 << 40, (n,), generator=g),
                gene=torch.randint(0, 30000, (n,), generator=g).to(torch.int32),
                score=torch.randint(-300, 300, (n,), generator=g).to(torch.int16),
                conv_n=torch.randint(0, 40000, (n,), generator=g).to(torch.int16),
                strand=torch.randint(0, 2, (n,), generator=g).to(torch.uint8),
                transcriptome=torch.randint(0, 2, (n,), generator=g).to(torch.uint8),
                counts=torch.randint(0, 256, (n, 16), generator=g).to(torch.uint8))
    d = {k: v.cuda() for k, v in cols.items()}
    perm, cnt = pipeline.owner_partition(ctx, d['barcode'], world)
    want_perm, want_cnt = fake_device.owner_partition(None, cols['barcode'], world)
    assert torch.equal(cnt.cpu(), want_cnt)
    assert torch.equal(perm.cpu(), want_perm)
    rows = pipeline.pack_reads(ctx, perm, d['counts'], d['barcode'], d['umi'], d['gene'], d['score'], d['conv_n'], d['strand'],
                               d['transcriptome'])
    want_rows = fake_device.pack_reads(None, want_perm, cols['counts'], cols['barcode'], cols['umi'], cols['gene'], cols['score'],
                                       cols['conv_n'], cols['strand'], cols['transcriptome'])
    assert torch.equal(rows.cpu(), want_rows)
    back = pipeline.unpack_reads(ctx, rows)
    for k, v in back.items():
        assert torch.equal(v.cpu(), cols[k][want_perm.long()]), k
    # perm None = identity, transcriptome None = all ones
    rows = pipeline.pack_reads(ctx, None, d['counts'], d['barcode'], d['umi'], d['gene'], d['score'], d['conv_n'], d['strand'])
    back = pipeline.unpack_reads(ctx, rows)
    assert torch.equal(back['umi'].cpu(), cols['umi']) and bool((back['transcriptome'] == 1).all())


@pytest.mark.parametrize('n,world', [(1, 2), (255, 2), (257, 3), (300000, 8), (70001, 16)])
def test_push_equals_partition_pack(gpu_ctx, n, world):
    """dyn_push_plan + dyn_push_reads (the fused exchange) put every row where dyn_owner_partition + dyn_pack_reads + an
    all-to-all would: here all `world` receive buffers live on this GPU, and this rank plays source `me` of `world` with
    made-up row counts from the lower ranks (dst_base)."""
    import ctypes as C
    from dynast_b200 import pipeline
    from dynast_b200._lib import check, ptr
    ctx = gpu_ctx
    g = torch.Generator().manual_seed(3 * n + world)
    d = dict(barcode=torch.randint(0, 5000, (n,), generator=g).to(torch.int32).cuda(),
             umi=torch.randint(0, 1 << 40, (n,), generator=g).cuda(),
             gene=torch.randint(0, 30000, (n,), generator=g).to(torch.int32).cuda(),
             score=torch.randint(-300, 300, (n,), generator=g).to(torch.int16).cuda(),
             conv_n=torch.randint(0, 40000, (n,), generator=g).to(torch.int16).cuda(),
             strand=torch.randint(0, 2, (n,), generator=g).to(torch.uint8).cuda(),
             transcriptome=torch.randint(0, 2, (n,), generator=g).to(torch.uint8).cuda(),
             counts=torch.randint(0, 256, (n, 16), generator=g).to(torch.uint8).cuda())
    perm, cnt = pipeline.owner_partition(ctx, d['barcode'], world)
    want = pipeline.pack_reads(ctx, perm, d['counts'], d['barcode'], d['umi'], d['gene'], d['score'], d['conv_n'], d['strand'],
                               d['transcriptome']).cpu()
    n_tiles = (n + 255) // 256
    tile_base = torch.empty(max(n_tiles * world, 1), dtype=torch.int32, device='cuda')
    cnt2 = torch.empty(world, dtype=torch.int64, device='cuda')
    check(ctx.lib.dyn_push_plan(ctx.handle, ptr(d['barcode']), n, world, ptr(tile_base), ptr(cnt2), pipeline._stream()))
    assert torch.equal(cnt2, cnt)
    cnt = cnt.cpu().numpy()
    start = np.concatenate([[0], np.cumsum(cnt)[:-1]])
    before = [(7 * o + 3) % 11 for o in range(world)]                 # rows "lower ranks" already sent to owner o
    bufs = [torch.full((before[o] + int(cnt[o]) + 5, 40), 0xee, dtype=torch.uint8, device='cuda') for o in range(world)]
    peers = (C.c_void_p * world)(*[b.data_ptr() for b in bufs])
    dst_base = (C.c_int64 * world)(*[before[o] - int(start[o]) for o in range(world)])
    check(ctx.lib.dyn_push_reads(ctx.handle, ptr(tile_base), n, world, ptr(d['counts']), ptr(d['barcode']), ptr(d['umi']), ptr(d['gene']),
                                 ptr(d['score']), ptr(d['conv_n']), ptr(d['strand']), ptr(d['transcriptome']), peers, dst_base,
                                 pipeline._stream()))
    torch.cuda.synchronize()
    for o in range(world):
        got = bufs[o].cpu()
        assert bool((got[:before[o]] == 0xee).all()) and bool((got[before[o] + int(cnt[o]):] == 0xee).all())
        assert torch.equal(got[before[o]:before[o] + int(cnt[o])], want[int(start[o]):int(start[o]) + int(cnt[o])]), o


def _run(rank, world, port, n_reads, n_barcodes, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import conftest  # noqa: F401  (package alias)
    import torch.distributed as dist
    from dynast_b200 import distributed, pipeline
    torch.cuda.set_device(rank)
    if world > 1:
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    ctx = pipeline.Context.get(rank)
    b = pipeline.SynthBatch(n_reads, seed=77, n_barcodes=n_barcodes, n_genes=300, device=rank, reads_per_umi=3.0)
    lo, hi = distributed.read_range(n_reads, rank, world)
    off = b.rec_off[lo:hi + 1].to(torch.int64) & 0xffffffff
    rec = b.records[int(off[0]) * 16:int(off[-1]) * 16]
    rec_off = (off - off[0]).to(torch.int32)
    out = pipeline.TallyBuffers(hi - lo, 4 * (hi - lo) + 1024, device=rank)
    if world > 1:       # a small exchange first: the peer buffers of the real one then have to GROW (collectively)
        z = lambda n, dt: torch.zeros(n, dtype=dt, device=rank)
        small = distributed.exchange_reads(ctx, b.barcode[lo:lo + 500], b.umi[lo:lo + 500], b.gene[lo:lo + 500], z(500, torch.int16), z(500, torch.int16),
                                           b.strand[lo:lo + 500], torch.zeros((500, 16), dtype=torch.uint8, device=rank), world, mode='push')
        assert int((small['barcode'] % world == rank).all()) == 1
        cap_small = distributed.PeerExchange.get(ctx, None).capacity
    res = pipeline.count_to_estimate(ctx, rec, rec_off, b.barcode[lo:hi], b.umi[lo:hi], b.gene[lo:hi], b.score[lo:hi],
                                     b.strand[lo:hi], n_barcodes, b.n_genes, out, cell_threshold=50, with_pi=False,
                                     exchange=world > 1)
    k, n, c, off_h = res['hist']
    if world > 1:       # the fused peer-store exchange ran (not the all-to-all fall-back), and both transports agree row for row
        px = distributed.PeerExchange.get(ctx, None)
        assert px.usable and px.capacity > cap_small > 0, 'peer memory exchange not in use, or its buffers did not grow'
        sc = b.score[lo:hi].to(torch.int16)
        args = (ctx, b.barcode[lo:hi], b.umi[lo:hi], b.gene[lo:hi], sc, out.conv_n, b.strand[lo:hi], out.counts, world)
        via_push = distributed.exchange_reads(*args, mode='push')
        via_nccl = distributed.exchange_reads(*args, mode='nccl')
        for key_ in via_nccl:
            assert torch.equal(via_push[key_], via_nccl[key_]), key_
    np.savez(os.path.join(out_dir, f'r{world}_{rank}.npz'), p_e=res['p_e'].cpu().numpy(), p_c=res['p_c'].cpu().numpy(),
             st=res['em_status'].cpu().numpy(), n_reads=res['n_reads'].cpu().numpy(), k=k.cpu().numpy(), n=n.cpu().numpy(),
             c=c.cpu().numpy(), off=off_h.cpu().numpy(), n_sel=int(res['selected'].numel()), n_local=int(res['n_local_reads']))
    if world > 1:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_two_gpu_path_equals_one_gpu(tmp_path):
    import torch.multiprocessing as mp
    n_reads, n_barcodes = 400000, 101
    port = 29600 + (os.getpid() % 2000)
    mp.spawn(_run, args=(1, port, n_reads, n_barcodes, str(tmp_path)), nprocs=1, join=True)
    mp.spawn(_run, args=(2, port + 1, n_reads, n_barcodes, str(tmp_path)), nprocs=2, join=True)
    one = np.load(tmp_path / 'r1_0.npz')
    two = [np.load(tmp_path / f'r2_{r}.npz') for r in range(2)]
    assert sum(int(t['n_local']) for t in two) == n_reads
    assert sum(int(t['n_sel']) for t in two) == int(one['n_sel'])
    checked = 0
    for b in range(n_barcodes):
        t, j = two[b % 2], b // 2
        assert one['n_reads'][b] == t['n_reads'][j]
        assert one['p_e'][b] == t['p_e'][j] or (np.isnan(one['p_e'][b]) and np.isnan(t['p_e'][j]))
        a0, a1 = one['off'][b], one['off'][b + 1]
        b0, b1 = t['off'][j], t['off'][j + 1]
        assert np.array_equal(one['k'][a0:a1], t['k'][b0:b1]) and np.array_equal(one['n'][a0:a1], t['n'][b0:b1])
        assert np.array_equal(one['c'][a0:a1], t['c'][b0:b1])
        assert one['st'][b] == t['st'][j]
        if one['st'][b] == 0:
            assert one['p_c'][b] == t['p_c'][j]
            checked += 1
    assert checked > 20


@pytest.mark.parametrize('n,world', [(0, 1), (1, 1), (5000, 1), (200000, 4)])
def test_merge_kn(gpu_ctx, n, world):
    """dyn_merge_kn + dyn_kn_csr against numpy: repeated keys are summed, group ids become local, rows sorted."""
    import fake_device
    from dynast_b200 import pipeline
    ctx = pipeline.Context.get(0)
    g = torch.Generator().manual_seed(n + world)
    n_groups = 300
    grp = torch.randint(0, n_groups, (n,), generator=g) * world + (world - 1)     # this rank owns ids = world - 1 mod world
    keys = (grp << 22) | (torch.randint(0, 6, (n,), generator=g) << 10) | torch.randint(0, 40, (n,), generator=g)
    cnt = torch.randint(1, 1000, (n,), generator=g).to(torch.int32)
    mk, mc, n_rows = pipeline.merge_kn(ctx, keys.cuda(), cnt.cuda(), world)
    wk, wc, wn = fake_device.merge_kn(None, keys, cnt, world)
    m = int(wn[0])
    assert int(n_rows.item()) == m
    assert torch.equal(mk.cpu()[:m], wk[:m]) and torch.equal(mc.cpu()[:m], wc[:m])
    got = pipeline.kn_csr(ctx, mk, mc, n_rows, n_groups)
    want = fake_device.kn_csr(None, wk, wc, wn, n_groups)
    tot = int(want[3][-1])
    for a, b in zip(got[:3], want[:3]):
        assert torch.equal(a.cpu()[:tot], b[:tot])
    assert torch.equal(got[3].cpu(), want[3]) and torch.equal(got[4].cpu(), want[4])
