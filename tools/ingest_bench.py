"""Host BAM ingest rate (dyn_bam_pack: BGZF inflate + filter + pack) on this machine's cores:
    python tools/ingest_bench.py <bam> [threads ...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import conftest  # noqa: F401  (package alias)
from dynast_b200 import bamio

path = sys.argv[1]
threads = [int(x) for x in sys.argv[2:]] or [1, 2, 4, os.cpu_count() or 8]
size = os.path.getsize(path)
for t in threads:
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        with bamio.PackedBam(path, barcode_tag='CB', umi_tag='UB', gene_tag='GX', require_gene=False, n_threads=t) as pb:
            n_units, n_seen, packed = pb.n_units, pb.n_records_seen, len(pb.blob)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    print(f'threads {t:3d}: {best:7.3f} s  {n_seen / best / 1e6:6.2f} M records/s  {n_units / best / 1e6:6.2f} M kept reads/s  '
          f'{size / best / 1e6:7.1f} MB/s of BAM  {packed / best / 1e6:7.1f} MB/s of packed records')
