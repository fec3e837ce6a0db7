"""The reference's OWN count_conversions (dynast/preprocessing/conversion.py:441-608, unmodified, loaded through
oracle/refstub.py) timed on this machine's cores: the CPU baseline SURVEY 8d asks for next to the C port.  Needs
/root/reference (the build container); the GPU box does not have it.
    python tools/cpu_reference.py [reads] [threads]  ->  profiles/r2_cpu_reference.txt"""
import gzip
import os
import pickle
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import refstub


def main():
    n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
    threads = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
    if not refstub.available():
        print('reference tree not present: nothing to time')
        return
    refstub.install()
    import dynast.preprocessing.conversion as conv
    rng = np.random.default_rng(1)
    n_genes, n_bc = 2000, 2000
    genes = {f'G{i:05d}': {'strand': '+' if i % 2 else '-', 'chromosome': f'chr{i % 25}'} for i in range(n_genes)}
    names = list(genes)
    bcs = [''.join(rng.choice(list('ACGT'), 16)) for _ in range(n_bc)]
    umis = [''.join(rng.choice(list('ACGT'), 12)) for _ in range(n_reads // 2)]
    t0 = time.perf_counter()
    ali, cv, index = [], [], []
    ali_pos = len('read_id,index,barcode,umi,GX,A,C,G,T,velocity,transcriptome,score\n')
    conv_pos = len('read_id,index,contig,genome_i,conversion,quality\n')
    gi = rng.integers(0, n_genes, n_reads); bi = rng.integers(0, n_bc, n_reads); ui = rng.integers(0, len(umis), n_reads)
    content = rng.integers(15, 30, (n_reads, 4)); m = rng.poisson(0.55, n_reads)
    for i in range(n_reads):
        rid = f'SYN:{i:09d}'
        gx = names[gi[i]]
        if m[i]:
            start = conv_pos
            for _ in range(m[i]):
                a = int(rng.integers(0, 4)); b = (a + int(rng.integers(1, 4))) % 4
                line = f'{rid},1,{genes[gx]["chromosome"]},{int(rng.integers(0, 1 << 27))},{"ACGT"[a]}{"ACGT"[b]},{int(rng.integers(2, 41))}\n'
                cv.append(line); conv_pos += len(line)
            index.append((start, int(m[i]), ali_pos))
        line = f'{rid},1,{bcs[bi[i]]},{umis[ui[i]]},{gx},{content[i, 0]},{content[i, 1]},{content[i, 2]},{content[i, 3]},unassigned,True,{91 - 2 * int(m[i])}\n'
        ali.append(line); ali_pos += len(line)
    tmp = tempfile.mkdtemp()
    ap, cp, ip, op = (os.path.join(tmp, x) for x in ('alignments.csv', 'conversions.csv', 'conversions.idx', 'counts.csv'))
    open(ap, 'w').write('read_id,index,barcode,umi,GX,A,C,G,T,velocity,transcriptome,score\n' + ''.join(ali))
    open(cp, 'w').write('read_id,index,contig,genome_i,conversion,quality\n' + ''.join(cv))
    with gzip.open(ip, 'wb') as f:
        pickle.dump(index, f)
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    conv.count_conversions(cp, ap, ip, op, genes, quality=27, conversions=['TC'], dedup_use_conversions=True, n_threads=threads, temp_dir=tmp)
    dt = time.perf_counter() - t0
    rows = sum(1 for _ in open(op)) - 1
    text = (f'reference dynast.preprocessing.conversion.count_conversions (unmodified, via oracle/refstub.py), {threads} threads, '
            f'{os.cpu_count()} cores in this container\n'
            f'input: {n_reads} synthetic reads (C2-like: 2 000 barcodes, 2 000 genes, ~0.55 mismatches per read) as conversions.csv / '
            f'alignments.csv / conversions.idx (generated in {t_gen:.1f} s, untimed)\n'
            f'count_conversions: {dt:.2f} s = {n_reads / dt:.3e} reads/s -> {rows} rows of counts_TC.csv\n'
            f'(this stage starts from the CSVs: the BAM parse of bam.parse_all_reads, which needs pysam, is not included; the C '
            f'port of the per-read loop that bench.py times runs ~{4e7:.0e} reads/s on 16 threads)\n')
    print(text)
    with open(os.path.join(ROOT, 'profiles', 'r2_cpu_reference.txt'), 'w') as f:
        f.write(text)


if __name__ == '__main__':
    main()
