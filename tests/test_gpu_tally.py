"""GPU: the CUDA tally (through the C-ABI, host-buffer form) is bit-exact against the C oracle and the
reference fixtures."""
import numpy as np
import pytest

from helpers import ref_path, ref_rows

pytestmark = pytest.mark.gpu


def _compare(gpu, orc):
    assert gpu['status'] == orc['status']
    assert (gpu['counts'] == orc['counts']).all()
    assert gpu['total'] == orc['total']
    assert (gpu['conv_n'] == orc['conv_n']).all()
    # placement of a read's records is not part of the contract (one atomic per tile); content is
    for name in ('conv_rec', 'conv_read'):
        g = np.concatenate([gpu[name][o:o + n] for o, n in zip(gpu['conv_off'], gpu['conv_n'])]) if gpu['total'] else gpu[name]
        r = np.concatenate([orc[name][o:o + n] for o, n in zip(orc['conv_off'], orc['conv_n'])]) if orc['total'] else orc[name]
        assert (g == r).all()


@pytest.mark.parametrize('align,kw', [
    ('SRR11683995_align', dict(umi_tag='UB', barcode_tag='CB', velocity=True)),
])
def test_fixture_bam_single_end(gpu_ctx, oracle_lib, align, kw):
    from oracle import pack
    from dynast_b200 import device, records as rec
    blob, off, meta, refs = pack.units_from_bam(ref_path(align, 'Aligned.sortedByCoord.out.bam'), rec.pack_one,
                                                rec.pack_units, **kw)
    orc = pack.run_tally(blob, off, quality=27)
    gpu = device.tally_reads_host(blob, off, quality=27)
    _compare(gpu, orc)
    # and against the fixture CSV directly
    ali = {(r['read_id'], int(r['index'])): r for r in ref_rows('SRR11683995_count', 'alignments.csv')}
    hit = 0
    for i, m in enumerate(meta):
        row = ali.get((m['read_id'], m['index']))
        if row:
            hit += 1
            assert [int(row[b]) for b in 'ACGT'] == [int(x) for x in gpu['counts'][i, 12:16]]
    assert hit == 250


@pytest.mark.parametrize('align,kw,nasc', [
    ('smartseq_align', dict(umi_tag=None, barcode_tag='RG', velocity=False), False),
    ('nasc_align', dict(umi_tag=None, barcode_tag='RG', velocity=False), True),
    ('nasc_align', dict(umi_tag=None, barcode_tag='RG', velocity=False), False),
])
def test_fixture_bam_paired(gpu_ctx, oracle_lib, align, kw, nasc):
    """Paired units (first-seen mate packed behind the read): overlap / mate-call rules of bam.py:307-352, 393-419."""
    from oracle import pack
    from dynast_b200 import device, records as rec
    blob, off, meta, refs = pack.units_from_bam(ref_path(align, 'Aligned.sortedByCoord.out.bam'), rec.pack_one,
                                                rec.pack_units, **kw)
    assert len(meta) > 100
    orc = pack.run_tally(blob, off, quality=27, nasc=nasc)
    gpu = device.tally_reads_host(blob, off, quality=27, nasc=nasc)
    assert orc['total'] > 0
    _compare(gpu, orc)


def test_event_overflow_and_long_reads(gpu_ctx, oracle_lib):
    """Reads with more mismatches than event slots, reads whose counters pass 255, and one record larger
    than the staging buffer all take the sequential walker and must still equal the oracle."""
    from oracle import pack
    from dynast_b200 import device, records as rec, synth
    d = synth.make_reads(600, seed=21, p_e=0.04)           # ~11 mismatches per read
    orc = pack.run_tally(d['blob'], d['rec_off'], quality=27)
    gpu = device.tally_reads_host(d['blob'], d['rec_off'], quality=27)
    assert (orc['conv_n'] > 8).sum() > 100
    _compare(gpu, orc)
    d = synth.make_reads(300, seed=22, read_len=300, p_e=0.002)   # A/C/G/T content can exceed 255 -> status bit
    orc = pack.run_tally(d['blob'], d['rec_off'], quality=27)
    gpu = device.tally_reads_host(d['blob'], d['rec_off'], quality=27)
    _compare(gpu, orc)
    # 400 short reads and one 40 kb record in the middle of a tile
    small = synth.make_reads(400, seed=23)
    big = synth.make_reads(1, seed=24, read_len=40000, p_e=0.001)
    units = []
    for src, lo, hi in ((small, 0, 200), (big, 0, 1), (small, 200, 400)):
        for i in range(lo, hi):
            a, b = int(src['rec_off'][i]) * 16, int(src['rec_off'][i + 1]) * 16
            units.append([bytes(src['blob'][a:b])])
    blob, off = rec.pack_units(units)
    orc = pack.run_tally(blob, off, quality=27)
    gpu = device.tally_reads_host(blob, off, quality=27)
    _compare(gpu, orc)


@pytest.mark.parametrize('n_reads,seed', [(1, 3), (31, 4), (257, 5), (20000, 6)])
def test_synthetic_vs_oracle(gpu_ctx, oracle_lib, n_reads, seed):
    from oracle import pack
    from dynast_b200 import device, synth
    d = synth.make_reads(n_reads, seed=seed)
    for quality in (27, 0, 41):
        orc = pack.run_tally(d['blob'], d['rec_off'], quality=quality)
        gpu = device.tally_reads_host(d['blob'], d['rec_off'], quality=quality)
        _compare(gpu, orc)


@pytest.mark.parametrize('read_len,p_e,seed', [(91, 5e-4, 31), (150, 0.01, 32), (40, 0.03, 33), (17, 0.05, 34)])
def test_complex_cigars(gpu_ctx, oracle_lib, read_len, p_e, seed):
    """Hard / soft clips at both ends, =/X pieces, several insertions, deletions and junctions per read, insertions
    next to the clips; N calls inside clipped and inserted stretches."""
    from oracle import pack
    from dynast_b200 import device, synth
    d = synth.make_reads(3000, seed=seed, read_len=read_len, p_e=p_e, frac_n=0.01, complex_cigars=True)
    orc = pack.run_tally(d['blob'], d['rec_off'], quality=20)
    gpu = device.tally_reads_host(d['blob'], d['rec_off'], quality=20)
    assert orc['status'] == 0 and orc['total'] > 0
    _compare(gpu, orc)


def test_snp_filter(gpu_ctx, oracle_lib):
    from oracle import pack
    from dynast_b200 import device, records as rec, synth
    d = synth.make_reads(5000, seed=11)
    base = pack.run_tally(d['blob'], d['rec_off'], quality=27)
    # declare every 3rd observed TC / AG mismatch position a SNP
    pos, refl, readl, q = rec.decode_conv(base['conv_rec'])
    hdr = d['blob'].view(np.uint8)
    ref_ids = np.array([int(np.frombuffer(d['blob'][int(o) * 16 + 4:int(o) * 16 + 8].tobytes(), '<i4')[0])
                        for o in d['rec_off'][:-1]])
    keys = []
    for j in range(0, len(pos), 3):
        conv = refl[j] + readl[j]
        if conv in rec.CONVERSION_COLUMNS:
            keys.append((rec.CONVERSION_COLUMNS.index(conv) << 56) | (int(ref_ids[base['conv_read'][j]]) << 32) | int(pos[j]))
    snps = np.unique(np.asarray(keys, dtype=np.uint64))
    orc = pack.run_tally(d['blob'], d['rec_off'], quality=27, snps=snps)
    gpu = device.tally_reads_host(d['blob'], d['rec_off'], quality=27, snps=snps)
    assert orc['counts'][:, :12].sum() < base['counts'][:, :12].sum()
    _compare(gpu, orc)


def test_no_emit_and_empty(gpu_ctx, oracle_lib):
    from oracle import pack
    from dynast_b200 import device, synth
    d = synth.make_reads(3000, seed=12)
    orc = pack.run_tally(d['blob'], d['rec_off'])
    gpu = device.tally_reads_host(d['blob'], d['rec_off'], emit_conv=False)
    assert (gpu['counts'] == orc['counts']).all()
    empty = device.tally_reads_host(np.zeros(0, np.uint8), np.zeros(1, np.uint32))
    assert empty['counts'].shape == (0, 16) and empty['total'] == 0


@pytest.mark.parametrize('paired', [False, True])
def test_conversion_capacity(gpu_ctx, oracle_lib, paired):
    """A mismatch-record buffer that is too small: DYN_ST_CONV_CAPACITY is raised, counters are unaffected, a read
    either has all of its records (inside the buffer) or reports none -- on the tile kernel and on the sequential
    kernel (paired units)."""
    import torch
    from oracle import pack
    from dynast_b200 import pipeline, records as rec, synth
    if paired:
        blob, off, meta, refs = pack.units_from_bam(ref_path('smartseq_align', 'Aligned.sortedByCoord.out.bam'), rec.pack_one,
                                                    rec.pack_units, umi_tag=None, barcode_tag='RG', velocity=False)
    else:
        d = synth.make_reads(5000, seed=41, p_e=0.01)
        blob, off = d['blob'], d['rec_off']
    orc = pack.run_tally(blob, off, quality=27)
    n = len(off) - 1
    cap = orc['total'] // 2
    assert cap > 10
    ctx = pipeline.Context.get(0)
    out = pipeline.TallyBuffers(n, cap, device=0)
    pipeline.tally_reads(ctx, torch.from_numpy(np.ascontiguousarray(blob)).cuda(), torch.from_numpy(off.view(np.int32)).cuda(), out,
                         quality=27)
    torch.cuda.synchronize()
    status = int(out.scalars[1].item()) & 0xffffffff
    assert status & 0x8 and int(out.scalars[0].item()) == orc['total']
    assert (out.counts.cpu().numpy() == orc['counts']).all()
    conv_n = out.conv_n.cpu().numpy().view(np.uint16)
    conv_off = out.conv_off.cpu().numpy().view(np.uint32)
    conv_rec = out.conv_rec.cpu().numpy().view(np.uint64)
    conv_read = out.conv_read.cpu().numpy().view(np.uint32)
    kept = dropped = 0
    for i in range(n):
        want = orc['conv_rec'][orc['conv_off'][i]:orc['conv_off'][i] + orc['conv_n'][i]]
        if conv_n[i]:
            assert conv_off[i] + conv_n[i] <= cap
            assert (conv_rec[conv_off[i]:conv_off[i] + conv_n[i]] == want).all()
            assert (conv_read[conv_off[i]:conv_off[i] + conv_n[i]] == i).all()
            kept += 1
        elif len(want):
            dropped += 1
    assert kept > 0 and dropped > 0


@pytest.mark.parametrize('kw', [
    dict(n_reads=20000, seed=51),
    dict(n_reads=3000, seed=52, read_len=150, p_e=0.01, frac_n=0.01, complex_cigars=True),
    dict(n_reads=600, seed=53, p_e=0.04),                      # event overflow -> sequential kernel on lean records
    dict(n_reads=300, seed=54, read_len=300, p_e=0.002),
])
def test_lean_records_give_the_results_of_full_records(gpu_ctx, oracle_lib, kw):
    """DYN_REC_LEAN (no Phred stream, the mismatching bases' qualities inside the MD tokens): the kernel and the
    oracle on lean records both equal the oracle on the full records -- counters, mismatch records, status."""
    from oracle import pack
    from dynast_b200 import device, synth
    full = synth.make_reads(**kw)
    lean = synth.make_reads(lean=True, **kw)
    assert lean['blob'].size < 0.7 * full['blob'].size
    for quality in (27, 0):
        want = pack.run_tally(full['blob'], full['rec_off'], quality=quality)
        _compare(pack.run_tally(lean['blob'], lean['rec_off'], quality=quality), want)
        _compare(device.tally_reads_host(lean['blob'], lean['rec_off'], quality=quality), want)


def test_lean_device_generator_matches_full(gpu_ctx, oracle_lib):
    """csrc/synth.cu with lean = 1 writes the same reads in the lean layout: same tally, about half the bytes."""
    import torch
    from oracle import pack
    from dynast_b200 import pipeline
    ctx = pipeline.Context.get(0)
    res = {}
    for lean in (False, True):
        b = pipeline.SynthBatch(50_000, seed=61, n_barcodes=50, n_genes=300, device=0, lean=lean)
        out = pipeline.TallyBuffers(b.n_reads, b.n_reads * 2 + 1024, device=0)
        pipeline.tally_reads(ctx, b.records, b.rec_off, out, quality=27)
        torch.cuda.synchronize()
        res[lean] = (out.counts.cpu().numpy(), int(out.scalars[0].item()), b.n_bytes,
                     pack.run_tally(b.records.cpu().numpy(), b.rec_off.cpu().numpy().view(np.uint32), quality=27))
        assert int(out.scalars[1].item()) == 0
    assert (res[False][0] == res[True][0]).all() and res[False][1] == res[True][1]
    assert (res[True][0] == res[True][3]['counts']).all() and (res[False][0] == res[False][3]['counts']).all()
    assert res[True][2] < 0.55 * res[False][2]


def test_lean_pair_is_rejected(gpu_ctx):
    """A unit with a mate cannot be lean (the mate rule compares qualities the lean layout does not carry)."""
    from dynast_b200 import device, records as rec
    a = rec.pack_one(100, 0, 0x1 | 0x40, [(0, 20)], 'A' * 20, [30] * 20, '20', lean=True)
    b = rec.pack_one(110, 0, 0x1 | 0x80, [(0, 20)], 'A' * 20, [30] * 20, '20')
    blob, off = rec.pack_units([[a, b]])
    out = device.tally_reads_host(blob, off)
    assert out['status'] & 0x2
