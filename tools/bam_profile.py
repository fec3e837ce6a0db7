"""Where the from-BAM path spends its time: python tools/bam_profile.py [reads]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import conftest  # noqa
import numpy as np, torch
from dynast_b200 import pipeline, synth, bamio
from dynast_b200._lib import Context
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16_000_000
ctx = Context.get(0)
path = '/dev/shm/dyn_prof.bam'
gen = pipeline.SynthBatch(n, seed=2002, device=0, lean=False)
info = synth.write_synth_bam(path, gen, level=1, n_threads=os.cpu_count())
strand = gen.tables['gene_strand'].cpu().numpy(); del gen; torch.cuda.empty_cache()
kw = dict(barcode_tag='CB', umi_tag='UB', gene_tag='GX', require_gene=True, n_threads=16)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with bamio.DeviceBamStream(ctx, path, lean=True, gene_ids=info['gene_ids'], **kw) as st:
        t_open = time.perf_counter()
        nn = 0
        for ch in st: nn += ch['n_units']
        torch.cuda.synchronize(); t_iter = time.perf_counter()
    t_close = time.perf_counter()
    print(f'ingest only: open {1e3*(t_open-t0):.1f} ms, chunks {1e3*(t_iter-t_open):.1f} ms, close {1e3*(t_close-t_iter):.1f} ms')
    torch.cuda.synchronize(); t0 = time.perf_counter()
    t = pipeline.tally_bam(ctx, path, info['gene_ids'], **kw)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    tm = {}
    r = pipeline.count_from_bam(ctx, path, info['gene_ids'], strand, timings=tm, **kw)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f'tally_bam {1e3*(t1-t0):.1f} ms; count_from_bam {1e3*(t2-t1):.1f} ms, device stages {tm}')
