"""Config C3 from FILE to FILE: a coordinate-sorted UMI BAM -> preprocessing.call_consensus -> consensus.bam (sorted, indexed)
-> count_from_bam on it.   python tools/consensus_file_bench.py [reads]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import conftest  # noqa
import numpy as np, torch
from dynast_b200 import pipeline, synth, bamio
from dynast_b200.preprocessing import call_consensus
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
ctx = pipeline.Context.get(0)
batch = pipeline.SynthBatch(n, seed=2003, n_barcodes=2000, n_genes=2000, reads_per_umi=3.5, device=0, clustered=True, lean=False)
path, out = '/dev/shm/dyn_c3_in.bam', '/dev/shm/dyn_c3_out.bam'
info = synth.write_synth_bam(path, batch, level=1, n_threads=os.cpu_count())
strand = batch.tables['gene_strand'].cpu().numpy(); contig = batch.tables['gene_contig'].cpu().numpy()
gene_infos = {g: dict(chromosome=f'chr{int(contig[i]) + 1}', strand='-' if strand[i] else '+', segment=(0, 1 << 28), gene_name=g) for i, g in enumerate(info['gene_ids'])}
del batch; torch.cuda.empty_cache()
print(f'input: {n} records, {os.path.getsize(path) / 1e6:.0f} MB')
for rep in range(2):
    t0 = time.perf_counter()
    call_consensus(path, out, gene_infos, strand='forward', umi_tag='UB', barcode_tag='CB', gene_tag='GX', quality=27, n_threads=os.cpu_count())
    t1 = time.perf_counter()
    print(f'call_consensus: {t1 - t0:.2f} s = {n / (t1 - t0):.3e} input reads/s -> {os.path.getsize(out) / 1e6:.0f} MB')
    try:
        t2 = time.perf_counter()
        r = pipeline.count_from_bam(ctx, out, info['gene_ids'], strand, cell_threshold=100, with_pi=False)
        torch.cuda.synchronize(); t3 = time.perf_counter()
        print(f'count_from_bam on the consensus BAM: {t3 - t2:.3f} s, {r["n_reads_tallied"]} consensus / pass-through reads, {int((r["em_status"] == 0).sum())} barcodes with p_c')
    except Exception as e:
        print('count on the consensus BAM failed:', str(e)[:300])
for p in (path, out, out + '.bai'):
    if os.path.exists(p): os.remove(p)
