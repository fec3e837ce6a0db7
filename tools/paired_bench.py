"""Throughput of the tally on paired units (the sequential kernel): the Smart-seq fixture's units repeated.
    python tools/paired_bench.py [copies]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import conftest  # noqa: F401
import numpy as np
import torch
from dynast_b200 import bamio, pipeline

copies = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
path = os.path.join(ROOT, 'tests', 'golden', 'ref', 'smartseq_align', 'Aligned.sortedByCoord.out.bam')
with bamio.PackedBam(path, barcode_tag='RG', umi_tag=None, gene_tag='GX', require_gene=True, n_threads=4) as pb:
    blob, off = pb.blob.copy(), pb.rec_off.copy().astype(np.int64)
n1 = len(off) - 1
big = np.tile(blob, copies)
offs = (off[None, :-1] + (np.arange(copies) * off[-1])[:, None]).reshape(-1)
offs = np.concatenate([offs, [copies * off[-1]]]).astype(np.uint32)
n = n1 * copies
ctx = pipeline.Context.get(0)
rec = torch.from_numpy(big).cuda()
ro = torch.from_numpy(offs.view(np.int32)).cuda()
out = pipeline.TallyBuffers(n, 8 * n, device=0)
for _ in range(3):
    pipeline.tally_reads(ctx, rec, ro, out, quality=27)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    pipeline.tally_reads(ctx, rec, ro, out, quality=27)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f'{n} paired units ({big.nbytes / 1e6:.0f} MB): {ms:.3f} ms  {n / ms / 1e3:.1f} M units/s  {big.nbytes / ms / 1e6:.1f} GB/s')
