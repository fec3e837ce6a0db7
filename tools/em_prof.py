"""Small EM run for profiling: python tools/em_prof.py [n_barcodes]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from dynast_b200 import pipeline
from dynast_b200._lib import Context
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
dev = torch.device('cuda', 0)
ctx = Context.get(0)
k, n, c, off, pe = bench.build_em_batch_gpu(nb, 2004, dev)
for _ in range(3):
    p_c, st, it = pipeline.em_pc(ctx, k, n, c, off, pe)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
p_c, st, it = pipeline.em_pc(ctx, k, n, c, off, pe)
e1.record()
torch.cuda.synchronize()
print('ms', e0.elapsed_time(e1), 'barcodes/s', nb / e0.elapsed_time(e1) * 1e3, 'mean iters', float(it.float().mean()))
