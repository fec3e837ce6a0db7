"""GPU: the whole path from a BAM FILE (pipeline.count_from_bam: streamed ingest -> tally per chunk -> UMI dedup -> p_e ->
histograms -> EM) equals the same path on the packed records the file was written from -- with the device ingest
(inflate + pack on the GPU) and with the host reader."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from helpers import ref_path

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def synth_bam(tmp_path_factory):
    from dynast_b200 import pipeline, synth
    ctx = pipeline.Context.get(0)
    batch = pipeline.SynthBatch(300_000, seed=91, n_barcodes=60, n_genes=400, device=0, lean=False, reads_per_umi=2.5)
    path = str(tmp_path_factory.mktemp('bam') / 'synth.bam')
    info = synth.write_synth_bam(path, batch, level=6, n_threads=4)
    # the reference result: the same reads, in the file's order, through the device path
    out = pipeline.TallyBuffers(batch.n_reads, 3 * batch.n_reads, device=0)
    pipeline.tally_reads(ctx, batch.records, batch.rec_off, out, quality=27)
    order = torch.from_numpy(info['order']).cuda()
    want = pipeline.count_to_estimate(ctx, None, None, batch.barcode[order], batch.umi[order], batch.gene[order], batch.score[order],
                                      batch.strand[order], 60, 400, None, cell_threshold=200, with_pi=False,
                                      tallied=(out.counts[order], out.conv_n[order]))
    torch.cuda.synchronize()
    strand = batch.tables['gene_strand'].cpu().numpy()
    return dict(path=path, info=info, want=want, strand=strand, n=batch.n_reads)


@pytest.mark.parametrize('ingest,chunk_bytes', [('device', 8 << 20), ('host', 8 << 20), ('device', 256 << 20)])
def test_count_from_bam_equals_packed_path(gpu_ctx, synth_bam, ingest, chunk_bytes):
    from dynast_b200 import bamio, pipeline, synth
    stats = {}
    got = pipeline.count_from_bam(gpu_ctx, synth_bam['path'], synth_bam['info']['gene_ids'], synth_bam['strand'], cell_threshold=200,
                                  with_pi=False, ingest=ingest, chunk_bytes=chunk_bytes, n_threads=4, stats=stats)
    torch.cuda.synchronize()
    want = synth_bam['want']
    assert stats['records_seen'] == stats['units'] == synth_bam['n'] == got['n_reads_tallied']
    assert int(got['selected'].numel()) == int(want['selected'].numel())
    # dense barcode ids here are ranks of the 2-bit keys: map the generator's ids onto them
    keys = [bamio.decode_key(k) for k in got['barcode_keys'].cpu().numpy()]
    names = [bytes(r).decode() for r in synth.barcode_strings(np.arange(60))]
    pos = [keys.index(nm) for nm in names]
    for name in ('n_reads', 'p_e', 'p_c', 'em_status'):
        a, b = want[name].cpu().numpy(), got[name].cpu().numpy()[pos]
        assert np.array_equal(a, b, equal_nan=(a.dtype.kind == 'f')), name
    assert (want['em_status'] == 0).sum() > 20
    k, n, c, off = (x.cpu().numpy() for x in want['hist'])
    k2, n2, c2, off2 = (x.cpu().numpy() for x in got['hist'])
    for b in range(60):
        j = pos[b]
        assert np.array_equal(k[off[b]:off[b + 1]], k2[off2[j]:off2[j + 1]]) and np.array_equal(c[off[b]:off[b + 1]], c2[off2[j]:off2[j + 1]])


def test_count_from_bam_auto_falls_back_to_the_host_reader(gpu_ctx):
    """Paired records: the device path refuses, `auto` reads the file on the host."""
    from dynast_b200 import bamio, pipeline
    path = ref_path('smartseq_align', 'Aligned.sortedByCoord.out.bam')
    with bamio.PackedBam(path, barcode_tag='RG', require_gene=False, n_threads=2) as pb:
        n = pb.n_units
    stats = {}
    t = pipeline.tally_bam(gpu_ctx, path, [], barcode_tag='RG', umi_tag=None, require_gene=False, ingest='auto', stats=stats)
    assert stats['ingest'].startswith('host') and t['counts'].shape[0] == n == stats['units']
    with pytest.raises(Exception, match='paired'):
        pipeline.tally_bam(gpu_ctx, path, [], barcode_tag='RG', umi_tag=None, require_gene=False, ingest='device')


def test_tally_bam_on_straddling_blocks_equals_the_host_reader(gpu_ctx, tmp_path):
    """Records straddling BGZF blocks (what STAR writes): device ingest and host reader give the same tally."""
    from dynast_b200 import bamio, pipeline
    out = str(tmp_path / 'split.bam')
    subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'make_bam.py'), ref_path('SRR11683995_align', 'Aligned.sortedByCoord.out.bam'),
                    out, '6'], check=True, capture_output=True)
    with bamio.PackedBam(out, barcode_tag='CB', umi_tag='UB', require_gene=True, n_threads=2) as pb:
        genes = sorted(set(pb.gene))
        n = pb.n_units
    stats = {}
    t = pipeline.tally_bam(gpu_ctx, out, genes, ingest='device', chunk_bytes=1 << 20, stats=stats)
    assert t['counts'].shape[0] == n == stats['units']
    h = pipeline.tally_bam(gpu_ctx, out, genes, ingest='host', chunk_bytes=0)
    assert torch.equal(t['counts'], h['counts']) and torch.equal(t['barcode_key'], h['barcode_key'])


def _bam_rank(rank, world, port, path, gene_ids, strand, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import conftest  # noqa: F401  (package alias)
    import torch.distributed as dist
    from dynast_b200 import pipeline
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    ctx = pipeline.Context.get(rank)
    stats = {}
    got = pipeline.count_from_bam(ctx, path, gene_ids, strand, cell_threshold=200, with_pi=False, group=dist.group.WORLD, chunk_bytes=8 << 20,
                                  n_threads=2, stats=stats)
    torch.cuda.synchronize()
    k, n, c, off = (x.cpu().numpy() for x in got['hist'])
    np.savez(os.path.join(out_dir, f'bam_r{rank}.npz'), keys=got['barcode_keys'].cpu().numpy(), n_reads=got['n_reads'].cpu().numpy(),
             p_e=got['p_e'].cpu().numpy(), p_c=got['p_c'].cpu().numpy(), st=got['em_status'].cpu().numpy(), n_sel=int(got['selected'].numel()),
             units=stats['units'], k=k, n=n, c=c, off=off)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_one_bam_on_two_gpus_equals_one_gpu(synth_bam, tmp_path):
    """Two ranks take the two block ranges of ONE file (device ingest), intern barcodes globally, exchange every read's row to
    the barcode's owner and run the rest of the path on the barcodes they own: per barcode the same reads, p_e, histograms
    and p_c as one GPU over the whole file, bit for bit."""
    import torch.multiprocessing as mp
    from dynast_b200 import bamio, synth
    port = 29700 + (os.getpid() % 2000)
    mp.spawn(_bam_rank, args=(2, port, synth_bam['path'], synth_bam['info']['gene_ids'], synth_bam['strand'], str(tmp_path)), nprocs=2, join=True)
    parts = [np.load(tmp_path / f'bam_r{r}.npz') for r in range(2)]
    want = synth_bam['want']
    assert sum(int(p['units']) for p in parts) == synth_bam['n']
    assert sum(int(p['n_sel']) for p in parts) == int(want['selected'].numel())
    assert np.array_equal(parts[0]['keys'], parts[1]['keys'])                      # one global barcode table
    keys = [bamio.decode_key(k) for k in parts[0]['keys']]
    names = [bytes(r).decode() for r in synth.barcode_strings(np.arange(60))]
    w = {name: want[name].cpu().numpy() for name in ('n_reads', 'p_e', 'p_c', 'em_status')}
    k, n, c, off = (x.cpu().numpy() for x in want['hist'])
    checked = 0
    for b, nm in enumerate(names):
        g = keys.index(nm)                     # global dense id: owner g % 2, local id g // 2
        p, j = parts[g % 2], g // 2
        assert w['n_reads'][b] == p['n_reads'][j] and w['em_status'][b] == p['st'][j]
        assert w['p_e'][b] == p['p_e'][j] or (np.isnan(w['p_e'][b]) and np.isnan(p['p_e'][j]))
        assert np.array_equal(k[off[b]:off[b + 1]], p['k'][p['off'][j]:p['off'][j + 1]])
        assert np.array_equal(c[off[b]:off[b + 1]], p['c'][p['off'][j]:p['off'][j + 1]])
        if w['em_status'][b] == 0:
            assert w['p_c'][b] == p['p_c'][j]
            checked += 1
    assert checked > 20
