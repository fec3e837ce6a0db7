"""CPU, world_size 2 over gloo: the histogram exchange before the EM (all-gather of rank-local (barcode, k, n,
count) rows, ownership by barcode, CSR assembly, result gather) gives every rank the same p_c vector as a
single-process run over the union of the rows.  The EM itself is the C oracle here (test infrastructure)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rows(seed, n_barcodes, n_rows):
    rng = np.random.default_rng(seed)
    b = rng.integers(0, n_barcodes, n_rows)
    n = rng.integers(5, 40, n_rows)
    k = np.minimum(rng.binomial(n, np.where(rng.random(n_rows) < 0.3, 0.06, 0.002)), n)
    c = rng.integers(1, 60, n_rows)
    return np.stack([b, k, n, c], axis=1).astype(np.int64)


def _oracle_em(k, n, c, off, p_e):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import oracle
    from helpers import oracle_em_batch
    p_c, st = oracle_em_batch(oracle.lib(), k.numpy(), n.numpy(), c.numpy(), off.numpy(), p_e.numpy())
    return torch.from_numpy(p_c), torch.from_numpy(st)


def _worker(rank, world, port, n_barcodes, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from dynast_b200 import distributed
    rows = torch.from_numpy(_rows(100 + rank, n_barcodes, 400 + 150 * rank))
    p_e = torch.full((n_barcodes,), 1e-3, dtype=torch.float64)
    gathered = distributed.all_gather_rows(rows)
    assert gathered.shape[0] == sum(400 + 150 * r for r in range(world))
    p_c, status = distributed.distributed_p_c(rows, p_e, n_barcodes, _oracle_em)
    np.save(os.path.join(out_dir, f'pc{rank}.npy'), p_c.numpy())
    np.save(os.path.join(out_dir, f'st{rank}.npy'), status.numpy())
    # per-read exchange to the barcode owner (all-to-all): every row must land on rank barcode % world, rows of
    # one source rank in their original order, sources in rank order
    g = torch.Generator().manual_seed(7 + rank)
    bc = torch.randint(0, n_barcodes, (300 + 50 * rank,), generator=g)
    payload = torch.stack([bc, torch.arange(len(bc)) + 1000 * rank], dim=1)
    counts = torch.randint(0, 255, (len(bc), 16), generator=g).to(torch.uint8)
    got_payload, got_counts = distributed.exchange_by_owner(bc % world, payload, counts)
    assert (got_payload[:, 0] % world == rank).all()
    assert got_counts.shape == (got_payload.shape[0], 16)
    src = got_payload[:, 1] // 1000
    assert (src[1:] >= src[:-1]).all()
    for r in range(world):
        idx = got_payload[src == r, 1]
        assert (idx[1:] > idx[:-1]).all()
    np.save(os.path.join(out_dir, f'xchg{rank}.npy'), got_payload.numpy())
    # the packed form of the same exchange (distributed.exchange_reads; numpy stand-ins for the three kernels)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import fake_device
    from dynast_b200 import pipeline
    for name in ('owner_partition', 'pack_reads', 'unpack_reads'):
        setattr(pipeline, name, getattr(fake_device, name))
    m = len(bc)
    cols = dict(barcode=bc.to(torch.int32), umi=torch.arange(m, dtype=torch.int64) + 10**9 * rank,
                gene=torch.randint(0, 500, (m,), generator=g).to(torch.int32),
                score=torch.randint(-50, 90, (m,), generator=g).to(torch.int16),
                conv_n=torch.randint(0, 9, (m,), generator=g).to(torch.int16),
                strand=torch.randint(0, 2, (m,), generator=g).to(torch.uint8))
    got = distributed.exchange_reads(None, cols['barcode'], cols['umi'], cols['gene'], cols['score'], cols['conv_n'],
                                     cols['strand'], counts, world)
    assert (got['barcode'] % world == rank).all() and (got['transcriptome'] == 1).all()
    assert torch.equal(got['counts'], got_counts) and torch.equal(got['barcode'].to(torch.int64), got_payload[:, 0])
    mine = got['umi'] // 10**9 == rank          # rows that stayed: columns still belong together
    keep = cols['barcode'] % world == rank
    for name in ('gene', 'score', 'conv_n', 'strand'):
        assert torch.equal(got[name][mine], cols[name][keep])
    # the key-encoded histogram exchange (distributed.distributed_p_c_keys) gives the same p_c as the row form
    for name in ('merge_kn', 'kn_csr'):
        setattr(pipeline, name, getattr(fake_device, name))
    keys = (rows[:, 0] << 22) | (rows[:, 1] << 10) | rows[:, 2]
    p_c2, status2 = distributed.distributed_p_c_keys(None, keys, rows[:, 3].to(torch.int32), p_e, n_barcodes, _oracle_em)
    assert torch.equal(status2, status)
    assert np.array_equal(p_c2.numpy(), p_c.numpy(), equal_nan=True)
    lo, hi = distributed.read_range(1001, rank, world)
    assert (lo, hi) == ((0, 501) if rank == 0 else (501, 1001))
    dist.destroy_process_group()


def test_histogram_exchange_world2(tmp_path, oracle_lib):
    world, n_barcodes = 2, 37
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n_barcodes, str(tmp_path)), nprocs=world, join=True)
    pcs = [np.load(tmp_path / f'pc{r}.npy') for r in range(world)]
    sts = [np.load(tmp_path / f'st{r}.npy') for r in range(world)]
    assert np.array_equal(sts[0], sts[1])
    assert np.array_equal(pcs[0], pcs[1], equal_nan=True)
    got = np.concatenate([np.load(tmp_path / f'xchg{r}.npy') for r in range(world)])
    assert len(got) == sum(300 + 50 * r for r in range(world)) and len(np.unique(got[:, 1])) == len(got)
    # single-process answer over the union of the rows
    sys.path.insert(0, ROOT)
    from dynast_b200 import distributed
    rows = torch.from_numpy(np.concatenate([_rows(100 + r, n_barcodes, 400 + 150 * r) for r in range(world)]))
    owned, k, n, c, off = distributed.owned_csr(rows, n_barcodes, 0, 1)
    want_pc, want_st = _oracle_em(k, n, c, off, torch.full((n_barcodes,), 1e-3, dtype=torch.float64))
    assert np.array_equal(sts[0], want_st.numpy())
    ok = want_st.numpy() == 0
    assert ok.sum() > 5
    assert np.array_equal(pcs[0][ok], want_pc.numpy()[ok])
