"""Small consensus run for profiling: python tools/cons_prof.py [n_reads]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dynast_b200 import pipeline
from dynast_b200._lib import Context
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
dev = torch.device('cuda', 0)
ctx = Context.get(0)
cb = pipeline.SynthBatch(n, seed=2003, n_barcodes=20000, reads_per_umi=3.5, device=0, clustered=True)
key = (cb.barcode.to(torch.int64) << 39) | ((cb.umi & 0xffffff) << 15) | cb.gene.to(torch.int64)
skey, order = torch.sort(key, stable=True)
uk, counts = torch.unique_consecutive(skey, return_counts=True)
multi = counts >= 2
keep = torch.repeat_interleave(multi, counts)
goff = torch.zeros(int(multi.sum().item()) + 1, dtype=torch.int64, device=dev)
goff[1:] = torch.cumsum(counts[multi], 0)
items = cb.rec_off[order[keep]].contiguous()
for _ in range(2):
    r = pipeline.consensus(ctx, cb.records, items, goff)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
r = pipeline.consensus(ctx, cb.records, items, goff)
e1.record()
torch.cuda.synchronize()
print('ms', e0.elapsed_time(e1), 'member reads/s', items.numel() / e0.elapsed_time(e1) * 1e3, 'groups', goff.numel() - 1)
# the count stage on the consensus records (second half of config C3): sizes, CIGAR shapes, tally time, parked units
rec_off = r['off']
n_cons = rec_off.numel() - 1
sizes = (rec_off[1:] - rec_off[:-1]).to(torch.int64) * 16
hdr = r['records'].view(torch.int32)
w = (rec_off[:-1].to(torch.int64) & 0xffffffff) * 4
l_seq = hdr[w + 2] & 0xffff
n_cig = (hdr[w + 2] >> 16) & 0xffff
n_tok = hdr[w + 3] & 0xffff
print('consensus records', n_cons, 'bytes mean/max', float(sizes.float().mean()), int(sizes.max()), 'l_seq mean/max', float(l_seq.float().mean()),
      int(l_seq.max()), 'n_cigar mean/max', float(n_cig.float().mean()), int(n_cig.max()), 'n_tok mean/max', float(n_tok.float().mean()), int(n_tok.max()))
out = pipeline.TallyBuffers(n_cons, 2 * n_cons + 1024, device=0)
for _ in range(2):
    pipeline.tally_reads(ctx, r['records'], rec_off, out, quality=27)
torch.cuda.synchronize()
e0.record()
pipeline.tally_reads(ctx, r['records'], rec_off, out, quality=27)
e1.record()
torch.cuda.synchronize()
print('tally of the consensus records: ms', e0.elapsed_time(e1), 'records/s', n_cons / e0.elapsed_time(e1) * 1e3, 'status', int(out.scalars[1].item()))
