"""From-BAM ingest rates on this box: the host reader (dyn_bam_stream_*, all cores) and the device ingest
(dyn_bam_dev_*: inflate + parse + pack on the GPU) over one synthetic coordinate-sorted BAM of C2-like reads.
    python tools/ingest_bench2.py [reads] [zlib level] [chunk MB]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import conftest  # noqa: F401
import numpy as np
import torch
from dynast_b200 import bamio, pipeline, synth
from dynast_b200._lib import Context

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
level = int(sys.argv[2]) if len(sys.argv) > 2 else 6
chunk = (int(sys.argv[3]) if len(sys.argv) > 3 else 256) << 20
cores = os.cpu_count() or 8
ctx = Context.get(0)
path = '/dev/shm/dyn_synth.bam' if os.path.isdir('/dev/shm') else '/tmp/dyn_synth.bam'
t0 = time.perf_counter()
batch = pipeline.SynthBatch(n, seed=2001, device=0, lean=False)
info = synth.write_synth_bam(path, batch, level=level, n_threads=cores)
del batch
torch.cuda.empty_cache()
size = os.path.getsize(path)
print(f'wrote {path}: {n} records, {size / 1e6:.1f} MB ({size / n:.1f} B/record), zlib level {level}, {time.perf_counter() - t0:.1f} s, {cores} cores')
kw = dict(barcode_tag='CB', umi_tag='UB', gene_tag='GX', require_gene=True)
for threads in sorted({cores, max(1, cores // 2)}):
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        units = seen = packed = 0
        with bamio.BamStream(path, keys=True, lean=True, gene_ids=info['gene_ids'], n_threads=threads, chunk_bytes=chunk, **kw) as st:
            for ch in st:
                units += ch.n_units; seen += ch.n_records_seen; packed += ch.blob.size
                ch.close()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    print(f'host reader, {threads:3d} threads: {best:7.3f} s  {seen / best / 1e6:7.2f} M records/s  {size / best / 1e6:8.1f} MB/s of BAM  ({units} units, {packed / max(units, 1):.1f} packed B/unit)')
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    units = seen = comp = infl = chunks = 0
    with bamio.DeviceBamStream(ctx, path, lean=True, gene_ids=info['gene_ids'], chunk_bytes=chunk, **kw) as st:
        for ch in st:
            units += ch['n_units']; seen += ch['n_records_seen']; comp += ch['compressed_bytes']; infl += ch['inflated_bytes']; chunks += 1
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f'device ingest, rep {rep}: {dt:7.3f} s  {seen / dt / 1e6:7.2f} M records/s  {comp / dt / 1e9:6.2f} GB/s compressed in  {infl / dt / 1e9:6.2f} GB/s inflated  ({units} units, {chunks} chunks)')
