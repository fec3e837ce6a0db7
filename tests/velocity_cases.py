"""Synthetic annotations and alignments for the velocity tests (shared by the CPU and the GPU test files)."""
import os

import numpy as np

from helpers import GOLDEN

FIXTURE_GTF = os.path.join(GOLDEN, 'velocity_fixture.gtf')


def random_gtf(path, seed=0, n_contigs=3, genes_per_contig=40, span=60_000):
    """Overlapping genes on both strands, 1-5 transcripts each, exons that sometimes touch or overlap (collapse /
    invert edge cases), transcripts shorter than their gene.  Returns the contig names."""
    rng = np.random.default_rng(seed)
    contigs = [f'c{i}' for i in range(n_contigs)]
    lines = ['# synthetic']
    gid = 0
    for c in contigs:
        for _ in range(genes_per_contig):
            g0 = int(rng.integers(0, span - 4000))
            g1 = g0 + int(rng.integers(300, 4000))
            strand = '+-'[int(rng.integers(0, 2))]
            gene = f'G{gid:04d}'
            gid += 1
            attr = f'gene_id "{gene}"; gene_name "N{gene}";'
            rows = [(c, 'gene', g0 + 1, g1, strand, attr)]
            for t in range(int(rng.integers(1, 6))):
                tid = f'{gene}.t{t}'
                t0 = int(rng.integers(g0, g0 + (g1 - g0) // 3))
                t1 = int(rng.integers(t0 + 100, g1 + 1)) if t0 + 100 < g1 else g1
                tattr = f'{attr} transcript_id "{tid}"; transcript_name "{tid}";'
                rows.append((c, 'transcript', t0 + 1, t1, strand, tattr))
                p = t0
                while p < t1:
                    e = min(t1, p + int(rng.integers(20, 300)))
                    rows.append((c, 'exon', p + 1, e, strand, tattr))
                    r = rng.random()
                    if r < 0.1:
                        p = e                 # the next exon touches this one
                    elif r < 0.2:
                        p = max(p + 1, e - 10)      # ... or overlaps it
                    else:
                        p = e + int(rng.integers(30, 600))
            order = rng.permutation(len(rows)) if rng.random() < 0.3 else range(len(rows))   # GTFs need not be sorted
            for i in order:
                cc, feat, a, b, s, at = rows[i]
                lines.append(f'{cc}\tsynth\t{feat}\t{a}\t{b}\t.\t{s}\t.\t{at}')
    with open(path, 'w') as f:
        f.write('\n'.join(lines) + '\n')
    return contigs


def random_alignments(gene_infos, contigs, n, seed=1, paired_frac=0.25, span=60_000):
    """n alignment units around the genes: list of dicts(ref_id, units=[(pos, cigar, flag)], gx).  CIGARs mix M, N
    (splices), D, I and soft clips; a quarter are pairs (read + first-seen mate); a third carry a gene tag."""
    rng = np.random.default_rng(seed)
    genes = list(gene_infos)
    out = []

    def one(pos):
        cigar = []
        if rng.random() < 0.2:
            cigar.append((4, int(rng.integers(1, 8))))
        for k in range(int(rng.integers(1, 5))):
            cigar.append((0, int(rng.integers(5, 60))))
            r = rng.random()
            if r < 0.35:
                cigar.append((3, int(rng.integers(20, 900))))
            elif r < 0.5:
                cigar.append((2, int(rng.integers(1, 4))))
            elif r < 0.65:
                cigar.append((1, int(rng.integers(1, 4))))
        while cigar[-1][0] in (1, 2, 3):
            cigar.pop()
        if rng.random() < 0.2:
            cigar.append((4, int(rng.integers(1, 8))))
        return pos, cigar

    for _ in range(n):
        g = genes[int(rng.integers(0, len(genes)))]
        info = gene_infos[g]
        a, b = info['segment']
        ref_id = contigs.index(info['chromosome'])
        pos = max(0, int(rng.integers(a - 150, b)))
        flag_rev = 0x10 if rng.random() < 0.5 else 0
        units = []
        if rng.random() < paired_frac:
            first_is_r1 = rng.random() < 0.5
            p1, c1 = one(pos)
            p2, c2 = one(max(0, pos + int(rng.integers(-40, 200))))
            units = [(p2, c2, 0x1 | (0x80 if first_is_r1 else 0x40) | (0 if flag_rev else 0x10)),      # the read (second seen)
                     (p1, c1, 0x1 | (0x40 if first_is_r1 else 0x80) | flag_rev)]                       # its first-seen mate
        else:
            p1, c1 = one(pos)
            units = [(p1, c1, flag_rev)]
        gx = ''
        r = rng.random()
        if r < 0.3:
            gx = g
        elif r < 0.36:       # a gene tag that disagrees with the position (still legal: same contig)
            same = [x for x in genes if gene_infos[x]['chromosome'] == info['chromosome']]
            gx = same[int(rng.integers(0, len(same)))]
        out.append(dict(ref_id=ref_id, units=units, gx=gx))
    return out


def reference_positions(pos, cigar):
    out = []
    r = pos
    for op, n in cigar:
        if op in (0, 7, 8):
            out.extend(range(r, r + n))
            r += n
        elif op in (2, 3):
            r += n
    return out


def md_for(cigar):
    """An MD tag consistent with the CIGAR (all matches, deletions of A's)."""
    out = []
    run = 0
    for op, n in cigar:
        if op in (0, 7, 8):
            run += n
        elif op == 2:
            out.append(f'{run}^' + 'A' * n)
            run = 0
    out.append(str(run))
    return ''.join(out)
