#!/usr/bin/env python
"""Benchmark of the count -> estimate hot path (contract: see the task statement / DESIGN.md section 6).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the device hot path over one batch of synthetic reads of BASELINE config C2
(50 M single-end 91-nt reads, 10 k barcodes, TC conversions) per GPU.  `value` is reads/s with the
packed records resident in HBM; `e2e` is the same work through the host-buffer C-ABI call
(dyn_tally_reads_host: pinned host records -> H2D -> kernels -> D2H).  The p_c EM (config C4:
1e5 barcodes x (k<=30, n<=150) histograms) is timed separately and reported under "em".
For N > 1 launch with torchrun; every rank runs its own read range (weak scaling), rank 0 prints.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import numpy as np  # noqa: E402

METRIC = 'aligned reads/s conversion counting; barcodes/s p_c EM'
UNIT = 'reads/s'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--reads', type=int, default=50_000_000, help='reads per GPU (C2: 50 M)')
    ap.add_argument('--barcodes', type=int, default=10_000)
    ap.add_argument('--em-barcodes', type=int, default=100_000, help='C4: 1e5 barcodes (whole job)')
    ap.add_argument('--cpu-sample-reads', type=int, default=8_000_000)
    ap.add_argument('--cpu-sample-barcodes', type=int, default=4000)
    ap.add_argument('--skip-e2e', action='store_true')
    ap.add_argument('--skip-cpu', action='store_true')
    ap.add_argument('--skip-em', action='store_true')
    ap.add_argument('--skip-pipeline', action='store_true')
    ap.add_argument('--skip-consensus', action='store_true')
    ap.add_argument('--consensus-reads', type=int, default=20_000_000, help='C3-like reads per GPU for the consensus stage')
    ap.add_argument('--no-emit', action='store_true', help='diagnostic: counters only, no mismatch records')
    ap.add_argument('--c5', action='store_true', help='also run config C5 per GPU: 62.5 M reads, 125 k barcodes, TC + AG labelling groups')
    ap.add_argument('--bam-max-total', type=int, default=48_000_000, help='cap of the BAM leg\'s file over all GPUs (records)')
    ap.add_argument('--skip-bam', action='store_true')
    ap.add_argument('--skip-other-layout', action='store_true')
    ap.add_argument('--skip-pi-g', action='store_true')
    ap.add_argument('--bam-reads', type=int, default=16_000_000, help='records per GPU of the synthetic BAM of the from-file leg')
    ap.add_argument('--bam-level', type=int, default=1, help='zlib level of that BAM (STAR writes level 1, samtools sort level 6)')
    ap.add_argument('--full-records', action='store_true',
                    help='records with their whole Phred stream (round-1 layout) instead of the lean single-end layout')
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '100', '-i',
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def mark(self):
        return time.time()

    def stop(self):
        if self.proc:
            self.proc.terminate()

    def summary(self, t0, t1):
        sm, mx, reasons = [], [], set()
        for t, line in self.rows:
            if t < t0 or t > t1 + 0.2:
                continue
            f = [x.strip() for x in line.split(',')]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons),
                'samples': len(sm)}


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def build_em_batch_gpu(n_barcodes, seed, device):
    """C4 histograms generated on the device with torch (setup, untimed). Returns CSR triplets."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    rng = np.random.default_rng(seed)
    ks, ns, cs, offs = [], [], [], [0]
    pes = rng.uniform(2e-4, 2e-3, n_barcodes)
    chunk = 2000
    for b0 in range(0, n_barcodes, chunk):
        b1 = min(n_barcodes, b0 + chunk)
        nb = b1 - b0
        R = np.maximum(1000, rng.lognormal(np.log(3000), 0.5, nb).astype(np.int64))
        pi = torch.from_numpy(rng.beta(2, 5, nb)).to(device)
        pcb = torch.from_numpy(rng.uniform(0.02, 0.08, nb)).to(device)
        peb = torch.from_numpy(pes[b0:b1]).to(device)
        bid = torch.repeat_interleave(torch.arange(nb, device=device), torch.from_numpy(R).to(device))
        n = torch.poisson(torch.full((bid.numel(),), 23.0, device=device), generator=g).clamp_(0, 150)
        lab = torch.rand(bid.numel(), device=device, generator=g) < pi[bid]
        p = torch.where(lab, pcb[bid], peb[bid])
        k = torch.binomial(n.double(), p, generator=g).clamp_(0, 30).long()
        n = n.long()
        # adversarial 1 %: force one read to touch k = 30, n = 150 (int64-wrap region of the coefficient)
        first = torch.cumsum(torch.from_numpy(R).to(device), 0) - torch.from_numpy(R).to(device)
        adv = first[torch.arange(0, nb, 100, device=device)]
        k[adv] = 30
        n[adv] = 150
        key = (bid * 31 + k) * 151 + n
        uk, cnt = torch.unique(key, return_counts=True)
        ub = uk // (31 * 151)
        ks.append(((uk // 151) % 31).int())
        ns.append((uk % 151).int())
        cs.append(cnt.long())
        per = torch.bincount(ub, minlength=nb).cpu().numpy()
        for x in per:
            offs.append(offs[-1] + int(x))
    return (torch.cat(ks), torch.cat(ns), torch.cat(cs), torch.tensor(offs, dtype=torch.int64, device=device),
            torch.from_numpy(pes).to(device))


def cpu_tally_baseline(blob, rec_off, n_threads):
    """C oracle (oracle/tally_oracle.c) over the host sample, split across threads (ctypes drops the GIL)."""
    import ctypes as C
    from concurrent.futures import ThreadPoolExecutor
    import oracle
    lib = oracle.lib()
    n = len(rec_off) - 1
    counts = np.zeros((n, 16), dtype=np.uint8)
    bounds = np.linspace(0, n, n_threads + 1).astype(np.int64)

    def work(i):
        st = C.c_uint32(0)
        lib.dyn_oracle_tally(blob.ctypes.data, rec_off.ctypes.data, int(bounds[i]), int(bounds[i + 1]), None, 0, 27, 0,
                             counts.ctypes.data, None, None, None, None, 0, None, C.byref(st))
        return st.value

    t0 = time.perf_counter()
    with ThreadPoolExecutor(n_threads) as ex:
        list(ex.map(work, range(n_threads)))
    dt = time.perf_counter() - t0
    return n / dt, counts


def cpu_em_baseline(k, n, c, off, pe, n_threads):
    from concurrent.futures import ThreadPoolExecutor
    import oracle
    from helpers import oracle_em_batch
    lib = oracle.lib()
    B = len(off) - 1
    bounds = np.linspace(0, B, n_threads + 1).astype(np.int64)

    def work(i):
        a, b = int(bounds[i]), int(bounds[i + 1])
        o = off[a:b + 1] - off[a]
        return oracle_em_batch(lib, k[off[a]:off[b]], n[off[a]:off[b]], c[off[a]:off[b]], o, pe[a:b])

    t0 = time.perf_counter()
    with ThreadPoolExecutor(n_threads) as ex:
        res = list(ex.map(work, range(n_threads)))
    dt = time.perf_counter() - t0
    return B / dt, np.concatenate([r[0] for r in res]), np.concatenate([r[1] for r in res])


def bind_near_gpu(index):
    """Several ranks on one box: keep this process (and the page-locked buffers it allocates, first-touch placement)
    on the CPUs NVML reports as local to its GPU, so that the host-buffer legs do not cross the socket interconnect.
    Returns the number of CPUs bound to, or None when NVML gives no answer."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_cpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [i for i in range(n_cpu) if (mask[i // 64] >> (i % 64)) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return len(allowed)
    except Exception:
        pass
    return None


def workload_config(args, world, numa=None):
    return {
        'workload': 'C2 synthetic scSLAM-seq-like 10x BAM records: 50M reads x 91 nt, 10k barcodes, TC, per GPU',
        'reads_per_gpu': args.reads, 'barcodes_per_gpu': args.barcodes, 'read_len': 91, 'quality': 27,
        'record_layout': 'full (Phred stream kept)' if args.full_records else
                         'lean (single-end: Phred only at mismatches, inside the MD tokens; include/dynast_b200.h DYN_REC_LEAN)',
        'l2_policy': 'inputs (~%.1f GB packed records per step) are far larger than the 126 MB L2' % (8.8 if args.full_records else 4.2),
        'parallelism': f'read-range x{world}', 'cpus_bound_near_gpu': numa,
    }


def reference_arm(args):
    """`--impl reference`: the reference's per-read tally (C restatement, oracle/tally_oracle.c) on all host threads.
    Nothing of the product is imported here: the input comes from the oracle's own numpy generator."""
    from oracle import synth_np
    n_cores = os.cpu_count() or 1
    sample = min(args.reads, args.cpu_sample_reads)
    unit = min(sample, 1_000_000)          # generated once, tiled up to the sample size (the numpy generator does ~0.2 M reads/s)
    blob1, off1 = synth_np.make_records(unit, seed=2001, lean=not args.full_records)
    reps = max(1, sample // unit)
    sample = reps * unit
    blob = np.tile(blob1, reps)
    rec_off = np.concatenate([off1[:-1].astype(np.int64) + r * int(off1[-1]) for r in range(reps)] + [[reps * int(off1[-1])]]).astype(np.uint32)
    vals = []
    for i in range(args.warmup + args.steps):
        v, _ = cpu_tally_baseline(blob, rec_off, n_cores)
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * sample / value,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
        'config': workload_config(args, int(os.environ.get('WORLD_SIZE', 1))),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': n_cores, 'kind': 'port',
                         'sample': f'{sample} reads of the C2 workload per step (oracle/synth_np.py, same generative model); C '
                                   'restatement of the reference tally (oracle/tally_oracle.c) on all host threads; the reference '
                                   'itself is CPython + pysam and cannot run on this box'},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))
    return 0


def multi_gpu_bit_equal(ctx, dist, rank, world, device, local_rank):
    """ONE global batch, sharded by read range over the ranks and pushed through the whole path with the per-read
    exchange, against the same batch on one GPU: per-barcode read counts, p_e, EM status and p_c must be identical bit
    for bit.  Returns True / False on every rank."""
    import torch
    from dynast_b200 import distributed, pipeline
    n_reads, n_bc = 2_000_000, 503
    b = pipeline.SynthBatch(n_reads, seed=77, n_barcodes=n_bc, n_genes=300, device=local_rank, reads_per_umi=3.0)

    def run(lo, hi, exchange):
        off = b.rec_off[lo:hi + 1].to(torch.int64) & 0xffffffff
        rec = b.records[int(off[0]) * 16:int(off[-1]) * 16]
        out = pipeline.TallyBuffers(hi - lo, 4 * (hi - lo) + 1024, device=local_rank)
        return pipeline.count_to_estimate(ctx, rec, (off - off[0]).to(torch.int32), b.barcode[lo:hi], b.umi[lo:hi], b.gene[lo:hi],
                                          b.score[lo:hi], b.strand[lo:hi], n_bc, b.n_genes, out, cell_threshold=50, with_pi=False,
                                          exchange=exchange)

    def columns(res):       # [4, n] int64: read count, p_e bits, EM status, p_c bits where the fit succeeded
        ok = res['em_status'] == 0
        return torch.stack([res['n_reads'].to(torch.int64), res['p_e'].view(torch.int64), res['em_status'].to(torch.int64),
                            torch.where(ok, res['p_c'].view(torch.int64), torch.zeros_like(res['p_c'].view(torch.int64)))])

    lo, hi = distributed.read_range(n_reads, rank, world)
    mine = columns(run(lo, hi, True))                 # barcodes rank, rank + world, ...
    per = (n_bc + world - 1) // world
    pad = torch.zeros((4, per), dtype=torch.int64, device=device)
    pad[:, :mine.shape[1]] = mine
    every = torch.empty((world, 4, per), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(every, pad)
    same = torch.zeros(1, dtype=torch.int64, device=device)
    if rank == 0:
        whole = columns(run(0, n_reads, False))
        sharded = every.permute(1, 2, 0).reshape(4, per * world)[:, :n_bc]      # barcode g sits at [g // world] of rank g % world
        same[0] = int(torch.equal(whole, sharded))
    dist.broadcast(same, src=0)
    return bool(same.item())


def exchange_transport(ctx):
    """Which transport the per-read exchange of the pipeline leg used (distributed.exchange_reads)."""
    from dynast_b200 import distributed
    px = distributed.PeerExchange._instances.get((id(ctx), id(None)))
    if px is not None and px.usable and px.capacity > 0:
        return ('40-byte per-read rows stored straight into the memory of rank barcode % world over NVLink by the gather kernel '
                '(dyn_push_reads, cudaIpc peer buffers); NCCL only for the counts and the barrier')
    return 'all-to-all of 40-byte per-read rows to rank barcode % world before dedup (dyn_pack_reads + NCCL all_to_all_single)'


def main():
    args = parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        return reference_arm(args) if rank == 0 else 0
    import torch
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device; dynast_b200 has no CPU fallback'
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    numa = bind_near_gpu(local_rank) if world > 1 else None
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=device)

    from dynast_b200 import pipeline
    from dynast_b200._lib import Context
    ctx = Context.get(local_rank)
    n_cores = os.cpu_count() or 1
    n_reads = args.reads
    config = workload_config(args, world, numa)

    # ---------------------------------------------------------------- workload (setup, untimed)
    batch = pipeline.SynthBatch(n_reads, seed=2001 + rank, n_barcodes=args.barcodes, device=local_rank,
                                barcode_base=rank * args.barcodes, lean=not args.full_records)
    conv_cap = int(n_reads * 1.5) + 1024
    out = pipeline.TallyBuffers(n_reads, conv_cap, device=local_rank)
    alg_in = batch.algorithmic_input_bytes()
    stream = torch.cuda.current_stream()

    def step():
        pipeline.tally_reads(ctx, batch.records, batch.rec_off, out, quality=27, emit_conv=not args.no_emit)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    n_conv = int(out.scalars[0].item())
    status = int(out.scalars[1].item()) & 0xffffffff
    assert status == 0 or os.environ.get('DYN_BENCH_IGNORE_STATUS'), f'tally status {status}'
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    t_mark0 = sampler.mark()
    ev[0].record(stream)
    for i in range(args.steps):
        step()
        ev[i + 1].record(stream)
    barrier()
    t_mark1 = sampler.mark()
    total_ms = ev[0].elapsed_time(ev[-1])
    kernel_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    if dist is not None:
        t = torch.tensor([total_ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * n_reads / (ms_per_step * 1e-3)

    # roofline of the dominant kernel (tally): algorithmic bytes / mean launch duration (CUDA events on
    # the launching stream; the step is that one kernel plus an 32-byte memset)
    alg_out = n_reads * (16 + 4 + 2) + n_conv * 12
    alg_bytes = alg_in + alg_out
    peak, peak_src = peaks()
    achieved = alg_bytes / (np.mean(kernel_ms) * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, 'profiles', 'r1_tally_traffic.json' if args.full_records else 'r2_tally_traffic.json')
    if os.path.exists(tp):      # DRAM bytes per read from the committed ncu --set full capture, scaled to this launch
        with open(tp) as f:
            tj = json.load(f)
        traffic, traffic_src = tj['bytes_per_read'] * n_reads, tj['source']
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': traffic, 'traffic_unit': 'bytes per launch (ncu DRAM bytes per read x reads in the launch)',
                'traffic_source': traffic_src, 'algorithmic_bytes_per_launch': alg_bytes, 'kernel': 'tally_kernel',
                'peak_source': peak_src, 'algorithmic_bytes_per_read': alg_bytes / n_reads}

    # the same kernel on the other record layout of the same reads (same seed): the lean layout halves the bytes a read needs,
    # the kernel's time does not change (it is bound by instruction issue), so its HBM fraction halves
    if not args.skip_other_layout:
        other = pipeline.SynthBatch(n_reads, seed=2001 + rank, n_barcodes=args.barcodes, device=local_rank,
                                    barcode_base=rank * args.barcodes, lean=bool(args.full_records))
        for _ in range(3):
            pipeline.tally_reads(ctx, other.records, other.rec_off, out, quality=27, emit_conv=not args.no_emit)
        barrier()
        oe = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        oe[0].record(stream)
        for i in range(args.steps):
            pipeline.tally_reads(ctx, other.records, other.rec_off, out, quality=27, emit_conv=not args.no_emit)
            oe[i + 1].record(stream)
        barrier()
        o_ms = float(np.mean([oe[i].elapsed_time(oe[i + 1]) for i in range(args.steps)]))
        o_bytes = other.algorithmic_input_bytes() + n_reads * (16 + 4 + 2) + int(out.scalars[0].item()) * 12
        roofline['other_layout'] = {'record_layout': 'lean' if args.full_records else 'full (Phred stream kept)', 'ms_per_launch': o_ms,
                                    'reads_per_s': n_reads / (o_ms * 1e-3), 'algorithmic_bytes_per_read': o_bytes / n_reads,
                                    'achieved': o_bytes / (o_ms * 1e-3) / 1e9, 'frac': o_bytes / (o_ms * 1e-3) / 1e9 / peak}
        del other
        torch.cuda.empty_cache()
        for _ in range(2):      # leave `out` holding the main batch's results for the legs below
            step()
        barrier()

    # ---------------------------------------------------------------- e2e through the host-buffer C-ABI
    e2e = None
    if not args.skip_e2e:
        import ctypes as C
        from dynast_b200._lib import check, ptr
        # ~10.5 GB of page-locked host memory per rank; if any rank cannot get it (many ranks on a small host), every
        # rank skips the leg together instead of one of them dying inside a collective
        pinned_ok, pin_error = True, None
        try:
            h_rec = batch.records.cpu().pin_memory()
            h_off = batch.rec_off.cpu().pin_memory()
            h_counts = torch.empty((n_reads, 16), dtype=torch.uint8).pin_memory()
            h_conv_off = torch.empty(n_reads, dtype=torch.int32).pin_memory()
            h_conv_n = torch.empty(n_reads, dtype=torch.int16).pin_memory()
            h_conv_rec = torch.empty(conv_cap, dtype=torch.int64).pin_memory()
            h_conv_read = torch.empty(conv_cap, dtype=torch.int32).pin_memory()
        except RuntimeError as exc:
            pinned_ok, pin_error = False, str(exc).splitlines()[0][:200]
        if dist is not None:
            flag = torch.tensor([1 if pinned_ok else 0], device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            pinned_ok = bool(flag.item())
        tot, st = C.c_uint64(0), C.c_uint32(0)
    if not args.skip_e2e and not pinned_ok:
        e2e = {'value': None, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0,
               'skipped': 'page-locked host buffers could not be allocated on every rank: %s' % (pin_error or 'another rank failed')}
    elif not args.skip_e2e:

        def e2e_step():
            check(ctx.lib.dyn_tally_reads_host(
                ctx.handle, ptr(h_rec), ptr(h_off), n_reads, None, 0, 27, pipeline.TALLY_EMIT_CONV, ptr(h_counts),
                ptr(h_conv_off), ptr(h_conv_n), ptr(h_conv_rec), ptr(h_conv_read), conv_cap,
                C.cast(C.byref(tot), C.c_void_p), C.cast(C.byref(st), C.c_void_p)))

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 3))
        for _ in range(n_e2e):
            e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / n_e2e
        if dist is not None:
            t = torch.tensor([dt], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        assert st.value == 0 and tot.value == n_conv
        assert torch.equal(h_counts[:100000], out.counts[:100000].cpu())
        e2e = {'value': world * n_reads / dt, 'unit': UNIT, 'h2d_bytes_per_step': int(batch.n_bytes + 4 * (n_reads + 1)),
               'd2h_bytes_per_step': int(n_reads * 22 + n_conv * 12 + 12), 'ms_per_step': dt * 1e3,
               'timing': 'host perf_counter around the blocking C-ABI call, device synchronised on both sides'}
        del h_rec, h_counts, h_conv_rec, h_conv_read

    # ---------------------------------------------------------------- whole device path on the same batch (untimed stages)
    pipe = None
    if not args.skip_pipeline:
        timings = {}
        reps = 2
        # N > 1: read-range sharding means a barcode's reads sit on every rank.  Give every (barcode, UMI) group a
        # global barcode id whose owner (id % world) is spread over the ranks, and let the path do its per-read
        # all-to-all to the owner before dedup; every rank ends up with ~n_reads reads of args.barcodes barcodes.
        xchg = world > 1
        p_bc = batch.barcode - rank * args.barcodes
        if xchg:
            p_bc = (p_bc.to(torch.int64) * world + batch.umi % world).to(torch.int32)
        p_nb = args.barcodes * world

        def run_path(t=None):
            return pipeline.count_to_estimate(ctx, batch.records, batch.rec_off, p_bc, batch.umi, batch.gene, batch.score,
                                              batch.strand, p_nb, batch.n_genes, out, timings=t, exchange=xchg)

        run_path()        # warm-up
        barrier()
        reps = 5
        per_rep, rep_ms = [], []
        for _ in range(reps):
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t_rep = {}
            p0.record(stream)
            res = run_path(t_rep)
            p1.record(stream)
            barrier()
            per_rep.append(t_rep)
            rep_ms.append(p0.elapsed_time(p1))
            if os.environ.get('DYN_BENCH_DEBUG'):
                st = torch.cuda.memory_stats()
                sys.stderr.write('alloc stats: device_alloc %d device_free %d retries %d reserved %.1f GB active %.1f GB\n' % (
                    st.get('num_device_alloc', -1), st.get('num_device_free', -1), st.get('num_alloc_retries', -1),
                    st.get('reserved_bytes.all.current', 0) / 1e9, st.get('active_bytes.all.current', 0) / 1e9))
        if os.environ.get('DYN_BENCH_DEBUG'):
            sys.stderr.write('pipeline per rep: %r %r\n' % (rep_ms, per_rep))
        # the MEDIAN repetition, all of them listed (several stages are host-driven -- split plan, shape discovery,
        # allocations -- and a descheduled host thread shows up as a spike in some repetition)
        best = int(np.argsort(rep_ms)[len(rep_ms) // 2])
        pipe_ms = float(rep_ms[best])
        timings = dict(per_rep[best])
        if dist is not None:
            t = torch.tensor([pipe_ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            pipe_ms = float(t.item())
        pipe = {'metric': 'reads/s, tally -> UMI dedup -> p_e -> (k,n) histograms -> EM -> pi/alpha, device resident',
                'value': world * n_reads / (pipe_ms * 1e-3), 'unit': UNIT, 'ms': pipe_ms,
                'stage_ms': timings, 'repetitions': reps, 'statistic': 'median repetition', 'rep_ms': [float(x) for x in rep_ms],
                'reads_after_dedup': int(res['selected'].numel()),
                'barcodes_with_p_c': int((res['em_status'] == 0).sum().item()),
                'exchange': (exchange_transport(ctx) if xchg else None),
                'note': 'includes the host round trips of the dedup split plan and the EM / pi shape discovery'}
        del res
        # estimate(method='pi_g'): one fit per (cell, gene) pair with >= 16 reads, batched on the device (pi.py:253-296 runs one
        # Stan optimisation per pair in a process pool)
        if world == 1 and not args.skip_pi_g:
            def run_pi_g(t=None):
                return pipeline.count_to_estimate(ctx, batch.records, batch.rec_off, p_bc, batch.umi, batch.gene, batch.score, batch.strand,
                                                  p_nb, batch.n_genes, out, timings=t, method='pi_g', cell_gene_threshold=16)
            run_pi_g()
            barrier()
            tg = {}
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record(stream)
            rg = run_pi_g(tg)
            g1.record(stream)
            barrier()
            fitted = int((rg['pi_g_status'] >= 0).sum().item())
            pipe['pi_g'] = {'metric': '(cell, gene) pairs/s, labeled-fraction fit per pair (estimate method pi_g)', 'pairs': int(rg['pi_g_pair'].numel()),
                            'pairs_fitted': fitted, 'pairs_ok': int((rg['pi_g_status'] == 0).sum().item()), 'ms_pi_g_stage': tg.get('pi_g'),
                            'value': fitted / (tg['pi_g'] * 1e-3) if tg.get('pi_g') else None, 'unit': 'pairs/s', 'ms_whole_path': g0.elapsed_time(g1),
                            'threshold': 'pairs with at least 16 reads of barcodes whose p_c was estimated'}
            del rg

    # ---------------------------------------------------------------- C5 (behind --c5): dual labelling, 1 M barcodes over 8 GPUs
    c5 = None
    if args.c5:
        n5, b5 = 62_500_000, 125_000
        g5 = pipeline.SynthBatch(n5, seed=2005 + rank, n_barcodes=b5, device=local_rank, lean=not args.full_records)
        o5 = pipeline.TallyBuffers(n5, int(n5 * 1.5) + 1024, device=local_rank)

        def run5(t=None):
            return pipeline.count_to_estimate(ctx, g5.records, g5.rec_off, g5.barcode, g5.umi, g5.gene, g5.score, g5.strand, b5, g5.n_genes, o5,
                                              conversions=('AG', 'TC'), conversion_groups=[('TC',), ('AG',)], cell_threshold=300, timings=t)

        run5()
        barrier()
        ms5, t5 = [], []
        for _ in range(3):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            tt = {}
            a0.record(stream); r5 = run5(tt); a1.record(stream)
            barrier()
            ms5.append(a0.elapsed_time(a1)); t5.append(tt)
        mid = int(np.argsort(ms5)[1])
        c5 = {'workload': 'C5 per GPU: 62.5 M reads, 125 k barcodes, --conversion TC --conversion AG (two aggregate / EM / pi passes)',
              'value': world * n5 / (ms5[mid] * 1e-3), 'unit': UNIT, 'ms': ms5[mid], 'rep_ms': ms5, 'stage_ms': t5[mid],
              'reads_after_dedup': int(r5['selected'].numel()),
              'barcodes_with_p_c': {k_: int((v['em_status'] == 0).sum().item()) for k_, v in r5['groups'].items()}}
        del g5, o5, r5
        torch.cuda.empty_cache()

    # ---------------------------------------------------------------- from a BAM FILE: ingest -> tally -> dedup -> estimate
    e2e_bam = None
    if not args.skip_bam:
        from dynast_b200 import synth
        # one file for the whole job, capped so that writing it (rank 0's host cores, untimed) and holding it in memory stay
        # bounded at N = 8: beyond the cap the per-GPU share shrinks and the JSON says so (records_per_gpu)
        n_bam = min(args.bam_reads * world, max(args.bam_max_total, args.bam_reads)) // world * world
        need = int(n_bam * 150 * 1.15)
        shm = None
        for cand in ('/dev/shm', '/tmp'):
            try:
                import shutil
                if os.path.isdir(cand) and shutil.disk_usage(cand).free > need:
                    shm = cand
                    break
            except OSError:
                pass
        if dist is not None:        # rank 0 decides where the file goes (or that there is no room for it)
            where = [shm]
            dist.broadcast_object_list(where, src=0)
            shm = where[0]
    if not args.skip_bam and shm is None:
        e2e_bam = {'value': None, 'unit': UNIT, 'skipped': f'no room for a {need >> 20} MiB BAM in /dev/shm or /tmp'}
    elif not args.skip_bam:
        bam_path = os.path.join(shm, f'dyn_bench_{os.getuid()}_{n_bam}_{args.bam_level}.bam')
        t_gen = time.perf_counter()
        if rank == 0:       # one coordinate-sorted file for the whole job; every rank then takes its block range of it
            gen = pipeline.SynthBatch(n_bam, seed=2002, n_barcodes=args.barcodes * world, device=local_rank, lean=False)
            synth.write_synth_bam(bam_path, gen, level=args.bam_level, n_threads=n_cores)
            np.save(bam_path + '.strand.npy', gen.tables['gene_strand'].cpu().numpy())
            n_genes_bam = gen.n_genes
            del gen
            torch.cuda.empty_cache()
        barrier()
        t_gen = time.perf_counter() - t_gen
        gene_strand = np.load(bam_path + '.strand.npy')
        gene_ids = synth.gene_names(len(gene_strand))
        grp = dist.group.WORLD if dist is not None else None
        bam_kw = dict(barcode_tag='CB', umi_tag='UB', gene_tag='GX', require_gene=True, n_threads=max(1, n_cores // world))

        def bam_run(ingest, stats):
            torch.cuda.synchronize()
            barrier()
            t0 = time.perf_counter()
            r = pipeline.count_from_bam(ctx, bam_path, gene_ids, gene_strand, group=grp, ingest=ingest, stats=stats, **bam_kw)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if dist is not None:
                t = torch.tensor([dt], device=device)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            return r, dt

        st = {}
        bam_run('auto', st)          # warm-up (buffers, page cache) ...
        bam_run('auto', {})          # ... twice: torch's caching allocator settles on its block sizes in the second pass
        runs = []
        for _ in range(5):
            st = {}
            r, dt = bam_run('auto', st)
            runs.append(dt)
            if os.environ.get('DYN_BENCH_DEBUG'):
                ms_ = torch.cuda.memory_stats()
                sys.stderr.write('bam rep %.3f s: device_alloc %d device_free %d reserved %.1f GB\n' % (
                    dt, ms_.get('num_device_alloc', -1), ms_.get('num_device_free', -1), ms_.get('reserved_bytes.all.current', 0) / 1e9))
        dt = float(np.median(runs))
        tallied = torch.tensor([r['n_reads_tallied'], int(r['selected'].numel()), int((r['em_status'] == 0).sum().item())], device=device)
        if dist is not None:
            dist.all_reduce(tallied)
        assert int(tallied[0].item()) == n_bam, (int(tallied[0].item()), n_bam)
        host_dt = None
        if world == 1:
            _, host_dt = bam_run('host', {})
        size = os.path.getsize(bam_path)
        e2e_bam = {'metric': 'aligned reads/s from a coordinate-sorted BAM FILE: BGZF inflate + record packing + tally + UMI dedup + p_e + '
                             '(k,n) histograms + EM + pi/alpha',
                   'value': n_bam / dt, 'unit': UNIT, 'records': n_bam, 'records_per_gpu': n_bam // world, 's': dt, 'rep_s': runs, 'statistic': 'median of 5 after two warm-ups', 'best_s': float(min(runs)),
                   'ingest': st.get('ingest'), 'file_bytes': size, 'bytes_per_record': size / n_bam, 'zlib_level': args.bam_level,
                   'file_on': shm, 'h2d_bytes_per_step': int(st.get('compressed_bytes', 0)), 'reads_after_dedup': int(tallied[1].item()),
                   'barcodes_with_p_c': int(tallied[2].item()),
                   'host_reader_value': (n_bam / host_dt) if host_dt else None,
                   'host_reader_note': f'same path with BGZF inflate + packing on the {n_cores} host cores (dyn_bam_stream_*)' if host_dt else None,
                   'sharding': f'block range {world} ways, per-read exchange to the barcode owner' if world > 1 else 'one range',
                   'timing': 'host perf_counter around pipeline.count_from_bam, device synchronised on both sides, max over ranks',
                   'setup_s_untimed': t_gen}
        if rank == 0:
            for p in (bam_path, bam_path + '.strand.npy'):
                try:
                    os.remove(p)
                except OSError:
                    pass

    # ---------------------------------------------------------------- UMI consensus (C3-like: 3.5 reads per UMI group)
    cons = None
    if not args.skip_consensus:
        nc = args.consensus_reads
        cb = pipeline.SynthBatch(nc, seed=2003 + rank, n_barcodes=2 * args.barcodes, reads_per_umi=3.5, device=local_rank, clustered=True)
        key = (cb.barcode.to(torch.int64) << 39) | ((cb.umi & 0xffffff) << 15) | cb.gene.to(torch.int64)

        def group_items():   # consensus.py:338-430: group by (gene, barcode, UMI); here a device sort of the interned key
            skey, order = torch.sort(key, stable=True)
            uk, counts = torch.unique_consecutive(skey, return_counts=True)
            multi = counts >= 2
            keep = torch.repeat_interleave(multi, counts)
            goff = torch.zeros(int(multi.sum().item()) + 1, dtype=torch.int64, device=device)
            goff[1:] = torch.cumsum(counts[multi], 0)
            group_items.keys = uk[multi]
            return (cb.rec_off[order[keep]]).contiguous(), goff

        items, goff = group_items()
        pipeline.consensus(ctx, cb.records, items, goff, quality=27)
        barrier()
        c0, c1, c2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        c0.record(stream)
        items, goff = group_items()
        c1.record(stream)
        cres = pipeline.consensus(ctx, cb.records, items, goff, quality=27)
        c2.record(stream)
        barrier()
        ok = bool((cres['status'] == 0).all().item())
        # C3 continues with `dynast count` on the consensus reads (dedup mode `exon`, count.py:295-304): tally + UMI dedup
        n_cons = int(goff.numel() - 1)
        gk = group_items.keys
        c_bc, c_umi, c_gene = (gk >> 39).to(torch.int32), (gk >> 15) & 0xffffff, (gk & 0x7fff).to(torch.int32)
        c_out = pipeline.TallyBuffers(n_cons, 2 * n_cons + 1024, device=local_rank)
        c_strand = cb.tables['gene_strand'][c_gene.long()]
        ckey = torch.empty(n_cons, dtype=torch.int64, device=device)
        c_score = torch.full((n_cons,), 91, dtype=torch.int16, device=device)
        c_tr = torch.ones(n_cons, dtype=torch.uint8, device=device)
        c3, c3b, c4 = (torch.cuda.Event(enable_timing=True) for _ in range(3))

        def count_on_consensus():
            c3.record(stream)
            pipeline.tally_reads(ctx, cres['records'], cres['off'], c_out, quality=27)
            c3b.record(stream)
            pipeline.check(ctx.lib.dyn_group_key(ctx.handle, pipeline.ptr(c_bc), pipeline.ptr(c_umi), pipeline.ptr(c_gene), n_cons, 24, 15,
                                                 pipeline.ptr(ckey), pipeline._stream()))
            sel_ = pipeline.dedup(ctx, ckey, c_out.counts, c_strand, c_tr, c_score, c_out.conv_n, c_bc, 2 * args.barcodes,
                                  conversions=('TC',), use_conversions=False, key_lo_bits=15 + 24 + 15)
            c4.record(stream)
            return sel_

        c_umi = c_umi.contiguous()
        count_on_consensus()            # warm-up: scratch of this size class is allocated on first use
        barrier()
        c_sel = count_on_consensus()
        barrier()
        count_ms = c3.elapsed_time(c4)
        cons = {'metric': 'member reads/s, UMI consensus (per-group shared-memory accumulation, tuple sort for scattered groups, assemble)',
                'value': world * int(items.numel()) / (c1.elapsed_time(c2) * 1e-3), 'unit': UNIT, 'ms_consensus': c1.elapsed_time(c2),
                'ms_grouping': c0.elapsed_time(c1), 'reads': nc, 'member_reads': int(items.numel()), 'groups': int(goff.numel() - 1),
                'all_status_ok': ok, 'workload': 'C3-like: 91-nt reads, 3.5 reads per (barcode, UMI, gene) group, starts within 300 nt',
                'count_on_consensus_ms': count_ms, 'count_on_consensus_tally_ms': c3.elapsed_time(c3b), 'count_on_consensus_rows': int(c_sel.numel()),
                'count_on_consensus_note': 'tally of the consensus records + UMI dedup in mode `exon` (the RN tag of a consensus BAM '
                                           'selects it, count.py:295-304): the second half of config C3',
                'c3_value': world * int(items.numel()) / ((c1.elapsed_time(c2) + c0.elapsed_time(c1) + count_ms) * 1e-3)}
        del cb, cres, items, goff, key
        torch.cuda.empty_cache()

    # ---------------------------------------------------------------- p_c EM (C4), barcodes split over ranks
    em = None
    if not args.skip_em:
        from dynast_b200 import distributed
        nb = args.em_barcodes // world
        n_global = nb * world
        k, n, c, off, pe = build_em_batch_gpu(nb, 2004 + rank, device)
        # rank-local histogram rows with global barcode ids; the EM of barcode b runs on rank b % world after
        # an all-gather of the rows (NCCL over NVLink) -- the one exchange of the path
        # global barcode id = local * world + (local + rank) % world: a rank's barcodes are owned by all ranks in turn, so
        # (world - 1) / world of its histogram rows really cross NVLink, as they do when every rank tallies a read range
        local_b = torch.repeat_interleave(torch.arange(nb, device=device), off[1:] - off[:-1])
        em_keys = (((local_b * world + (local_b + rank) % world) << 22) | (k.long() << 10) | n.long()).contiguous()
        em_cnt32 = c.to(torch.int32).contiguous()
        if dist is not None:
            pe_all = torch.empty(n_global, dtype=torch.float64, device=device)
            dist.all_gather_into_tensor(pe_all, pe.contiguous())
            gid = torch.arange(n_global, device=device)
            loc = gid // world
            pe_global = pe_all.view(world, nb)[(gid % world - loc) % world, loc].contiguous()
        else:
            pe_global = pe
        iters_box = {}

        def em_fn(k_, n_, c_, off_, pe_):
            p_c_, st_, it_ = pipeline.em_pc(ctx, k_, n_, c_, off_, pe_.contiguous())
            iters_box['iters'] = it_
            return p_c_, st_

        def em_step():
            if dist is None:      # one GPU: the CSR histograms go straight to the kernel, there is nothing to exchange
                return em_fn(k, n, c, off, pe)
            return distributed.distributed_p_c_keys(ctx, em_keys, em_cnt32, pe_global, n_global, em_fn)

        for _ in range(2):
            p_c, st_em = em_step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record(stream)
        for _ in range(reps):
            p_c, st_em = em_step()
        e1.record(stream)
        barrier()
        em_ms = e0.elapsed_time(e1) / reps
        if dist is not None:
            t = torch.tensor([em_ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            em_ms = float(t.item())
        iters = iters_box['iters']
        em = {'metric': 'barcodes/s p_c EM', 'value': n_global / (em_ms * 1e-3), 'unit': 'barcodes/s',
              'barcodes': n_global, 'ms': em_ms, 'ok_fraction': float((st_em == 0).float().mean().item()),
              'mean_iters': float(iters.float().mean().item()), 'workload': 'C4: k<=30, n<=150, 1% adversarial full shape',
              'dtype': 'f64 E step / f32 M step (as the reference)',
              'exchange': 'none (1 GPU)' if dist is None else 'all-to-all of (barcode,k,n) keys + counts to the barcode owner, device merge + CSR, all-gather of p_c: inside the timed region'}

    # ---------------------------------------------------------------- N > 1: the sharded path against one GPU, bit for bit
    bit_equal = None
    if dist is not None:
        bit_equal = multi_gpu_bit_equal(ctx, dist, rank, world, device, local_rank)

    # ---------------------------------------------------------------- CPU baseline beside it (rank 0, N = 1)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        sample = min(n_reads, args.cpu_sample_reads)
        hi = int(batch.rec_off[sample].item()) & 0xffffffff
        blob = batch.records[:hi * 16].cpu().numpy()
        rec_off = batch.rec_off[:sample + 1].cpu().numpy().view(np.uint32)
        v, cpu_counts = cpu_tally_baseline(blob, rec_off, n_cores)
        same = bool((cpu_counts == out.counts[:sample].cpu().numpy()).all())
        cpu_baseline = {'value': v, 'unit': UNIT, 'cores': n_cores, 'kind': 'port',
                        'sample': f'first {sample} reads of the same batch, C restatement of the reference tally '
                                  f'(oracle/tally_oracle.c), {n_cores} threads; counters identical to the GPU: {same}'}
        if em is not None:
            sb = min(nb, args.cpu_sample_barcodes)
            o = off[:sb + 1].cpu().numpy()
            ve, cp, cs = cpu_em_baseline(k[:o[-1]].cpu().numpy(), n[:o[-1]].cpu().numpy(), c[:o[-1]].cpu().numpy(), o,
                                         pe[:sb].cpu().numpy(), n_cores)
            g_pc, g_st = p_c[:sb].cpu().numpy(), st_em[:sb].cpu().numpy()
            okm = cs == 0
            em['cpu_baseline'] = {'value': ve, 'unit': 'barcodes/s', 'cores': n_cores, 'kind': 'port',
                                  'sample': f'{sb} barcodes, oracle/em_oracle.c',
                                  'status_equal': bool((cs == g_st).all()),
                                  'bit_exact': bool((cp[okm] == g_pc[okm]).all())}
    sampler.stop()
    clocks = sampler.summary(t_mark0, t_mark1)

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic', 'config': config, 'gpu_launches': 2 * args.steps,
            'roofline': roofline, 'cpu_baseline': cpu_baseline, 'e2e': e2e, 'e2e_bam': e2e_bam, 'c5': c5, 'em': em, 'pipeline': pipe, 'consensus': cons, 'clocks': clocks,
            'conversions_per_read': n_conv / n_reads, 'multi_gpu_bit_equal': bit_equal,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
