"""GPU: dyn_consensus against the consensus oracle (pinned by the reference's known-answer test) -- the
known-answer group itself and randomised groups with spliced reads, deletions, insertions, soft clips, N bases,
low qualities and ties."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LETTERS = 'ACGT'


def simulate_read(rng, genome, pos, cigar, p_mis=0.05, p_n=0.02, low_q=0.15):
    """cigar: [(op letter, n)]. Returns (seq, qual, md) consistent with `genome` (list of letters, may hold 'N')."""
    seq, qual, md = [], [], []
    run = 0
    g = pos
    for op, n in cigar:
        if op == 'M':
            for _ in range(n):
                ref = genome[g]
                g += 1
                b = ref if ref != 'N' else LETTERS[int(rng.integers(0, 4))]
                if rng.random() < p_mis:
                    b = LETTERS[(LETTERS.index(b) + int(rng.integers(1, 4))) % 4]
                if rng.random() < p_n:
                    b = 'N'
                seq.append(b)
                qual.append(int(rng.integers(2, 27)) if rng.random() < low_q else int(rng.integers(27, 41)))
                if b == ref:
                    run += 1
                else:
                    md.append(str(run)); md.append(ref); run = 0
        elif op == 'D':
            md.append(str(run)); run = 0
            md.append('^' + ''.join(genome[g:g + n]))
            g += n
        elif op == 'N':
            g += n
        elif op in 'IS':
            for _ in range(n):
                seq.append(LETTERS[int(rng.integers(0, 4))])
                qual.append(int(rng.integers(2, 41)))
    md.append(str(run))
    return ''.join(seq), qual, ''.join(md)


def random_groups(seed, n_groups, max_reads=7):
    from oracle.consensus_oracle import make_read
    rng = np.random.default_rng(seed)
    genome = [LETTERS[i] for i in rng.integers(0, 4, 6000)]
    for i in rng.integers(0, len(genome), 40):
        genome[i] = 'N'
    groups = []
    for g in range(n_groups):
        base = int(rng.integers(0, 3000))
        reads = []
        for r in range(int(rng.integers(2, max_reads + 1))):
            pos = base + int(rng.integers(0, 60))
            kind = rng.random()
            L = int(rng.integers(20, 70))
            if kind < 0.4:
                cigar = [('M', L)]
            elif kind < 0.6:
                a = int(rng.integers(5, L - 5))
                cigar = [('M', a), ('N', int(rng.integers(20, 900))), ('M', L - a)]
            elif kind < 0.75:
                a = int(rng.integers(5, L - 5))
                cigar = [('M', a), ('D', int(rng.integers(1, 4))), ('M', L - a)]
            elif kind < 0.85:
                a = int(rng.integers(5, L - 5))
                cigar = [('M', a), ('I', int(rng.integers(1, 4))), ('M', L - a)]
            elif kind < 0.93:
                cigar = [('S', int(rng.integers(1, 8))), ('M', L), ('S', int(rng.integers(1, 5)))]
            else:
                a = int(rng.integers(5, L - 10))
                cigar = [('M', a), ('I', 2), ('D', 2), ('M', 5), ('N', 50), ('D', 1), ('M', L - a - 5)]
            seq, qual, md = simulate_read(rng, genome, pos, cigar)
            reads.append(make_read(f'g{g}r{r}', pos, cigar, seq, qual, md, ref_id=g % 3))
        groups.append(reads)
    return groups


def run_gpu(groups, quality=27):
    import torch
    from oracle import bamio
    from dynast_b200 import pipeline, records as rec
    from dynast_b200._lib import Context
    ctx = Context.get(0)
    units, item_rec, group_off = [], [], [0]
    total16 = 0
    for reads in groups:
        for r in reads:
            b = rec.pack_one(r.reference_start, r.reference_id, r.flag, r.cigar, r.query_sequence, r.query_qualities, r.tags['MD'])
            item_rec.append(total16)
            total16 += len(b) // 16
            units.append([b])
        group_off.append(len(item_rec))
    blob, _ = rec.pack_units(units)
    dev = torch.device('cuda', 0)
    out = pipeline.consensus(ctx, torch.from_numpy(np.ascontiguousarray(blob)).to(dev),
                             torch.from_numpy(np.asarray(item_rec, dtype=np.int32)).to(dev),
                             torch.from_numpy(np.asarray(group_off, dtype=np.int64)).to(dev), quality=quality)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}


def decode(out, g):
    from dynast_b200 import records as rec
    o = int(out['off'][g]) * 16
    b = out['records'][o:int(out['off'][g + 1]) * 16].tobytes()
    pos, ref_id = np.frombuffer(b[:8], '<i4')
    l_seq, n_cigar, n_tok, flag = (int(x) for x in np.frombuffer(b[8:16], '<u2'))
    cig = np.frombuffer(b[16:16 + 4 * n_cigar], '<u4')
    tok = np.frombuffer(b[16 + 4 * n_cigar:16 + 4 * n_cigar + 4 * n_tok], '<u4')
    p = 16 + 4 * n_cigar + 4 * n_tok
    nb = ((l_seq + 1) // 2 + 3) & ~3
    seqb = b[p:p + nb]
    seq = ''.join(rec.SEQ_CODES[(seqb[i >> 1] >> (4 if i % 2 == 0 else 0)) & 15] for i in range(l_seq))
    qual = list(b[p + nb:p + nb + l_seq])
    md = out['md'][int(out['md_off'][g]):int(out['md_off'][g + 1])].tobytes().decode()
    cigar = [('MIDNSHP=X'[int(c) & 15], int(c) >> 4) for c in cig]
    # the record's MD tokens say the same as its MD text (same letters at the same aligned offsets)
    letters = [rec.SEQ_CODES[(int(t) >> 16) & 15] for t in tok]
    assert letters == [c for c in md if c.isalpha()]
    assert all(0 <= (int(t) & 0xffff) <= l_seq for t in tok)
    return dict(start=int(pos), ref_id=int(ref_id), seq=seq, qual=qual, md=md, flag=int(flag), cigar=cigar, nm=int(out['nm'][g]))


def check_groups(groups, quality=27):
    from oracle.consensus_oracle import call_consensus
    out = run_gpu(groups, quality=quality)
    assert (out['status'] == 0).all()
    for g, reads in enumerate(groups):
        want = call_consensus(reads, quality=quality)
        got = decode(out, g)
        assert got['start'] == want['start'], g
        assert got['cigar'] == want['cigar'], g
        assert got['seq'] == want['seq'], g
        assert got['qual'] == want['qual'], g
        assert got['md'] == want['md'], g
        assert got['nm'] == want['nm'], g
        assert got['ref_id'] == reads[0].reference_id


def test_known_answer(gpu_ctx):
    from test_oracle_consensus import known_answer_reads
    out = run_gpu([known_answer_reads()])
    got = decode(out, 0)
    assert got['seq'] == 'ACTGCCAGCTGC' and got['start'] == 1
    assert ''.join(f'{n}{op}' for op, n in got['cigar']) == '10M1D2M'
    assert got['md'] == '5T4^A2' and got['nm'] == 1


@pytest.mark.parametrize('seed,n_groups,quality', [(1, 50, 27), (2, 400, 27), (3, 200, 35), (4, 200, 0)])
def test_random_groups_vs_oracle(gpu_ctx, seed, n_groups, quality):
    check_groups(random_groups(seed, n_groups), quality=quality)


def test_large_group_and_singleton(gpu_ctx):
    groups = random_groups(9, 3, max_reads=3)
    groups.append(random_groups(10, 1, max_reads=2)[0][:1])          # a single read is its own consensus
    big = []
    for g in random_groups(11, 40, max_reads=7):
        big.extend(g)                                                # 150+ reads spread over 3 kb
    for r in big:
        r.reference_id = 0
    groups.append(big)
    check_groups(groups)


def test_deep_groups_both_paths(gpu_ctx):
    """256 reads within one window stay on the direct path (16-bit quality sums: 256 * 93 fits), 257 take the sorted
    path; both must equal the oracle."""
    from oracle.consensus_oracle import make_read
    rng = np.random.default_rng(12)
    genome = [LETTERS[i] for i in rng.integers(0, 4, 400)]
    groups = []
    for depth in (256, 257, 300):
        reads = []
        for r in range(depth):
            pos = int(rng.integers(0, 40))
            L = int(rng.integers(30, 90))
            cigar = [('M', L)] if r % 5 else [('M', 10), ('D', 2), ('M', L - 10)]
            seq, qual, md = simulate_read(rng, genome, pos, cigar, p_mis=0.2, low_q=0.05)
            qual = [93 if q > 30 else q for q in qual]
            reads.append(make_read(f'd{depth}r{r}', pos, cigar, seq, qual, md, ref_id=1))
        groups.append(reads)
    check_groups(groups)


def test_long_introns(gpu_ctx):
    """Gaps between covered positions: up to 65 535 skipped positions ride in the direct path's cells, longer ones send
    the group to the sorted path."""
    from oracle.consensus_oracle import make_read
    rng = np.random.default_rng(13)
    genome = [LETTERS[i] for i in rng.integers(0, 4, 200000)]
    groups = []
    for gap in (1000, 65535, 65536, 150000):
        reads = []
        for r in range(3):
            pos = int(rng.integers(0, 30))
            cigar = [('M', 40), ('N', gap - r), ('M', 35)]
            seq, qual, md = simulate_read(rng, genome, pos, cigar)
            reads.append(make_read(f'i{gap}r{r}', pos, cigar, seq, qual, md, ref_id=0))
        groups.append(reads)
    check_groups(groups)
