"""TEST INFRASTRUCTURE ONLY -- (re)generate the committed golden vectors under tests/golden/.

Run in the build container, where /root/reference exists:   python -m oracle.make_golden

  ref/...            verbatim copies of the reference's own small test fixtures for the hot path
                     (tests/fixtures/*: three tiny STAR BAMs and the per-stage output CSVs; text files
                     are gzip-compressed).  They are the reference's golden vectors.
  gene_strands.json  gene -> strand for the genes the fixtures mention (from genes.pkl.gz, which is a
                     2 MB pickle of ngs_tools objects and is not copied)
  em_golden.npz      outputs of the reference's numba functions p_c.expectation_maximization /
                     expectation_maximization_nasc / binomial_pmf on seeded synthetic histograms
  dedup_golden.npz, counts_golden/... (added by later sections of this script)

Nothing here runs on the GPU box; tests read only the committed files.
"""
import gzip
import json
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, 'tests', 'golden')
sys.path.insert(0, ROOT)

from oracle import refstub  # noqa: E402

FIXTURE_FILES = {
    'SRR11683995_align': ['Aligned.sortedByCoord.out.bam'],
    'smartseq_align': ['Aligned.sortedByCoord.out.bam'],
    'nasc_align': ['Aligned.sortedByCoord.out.bam'],
    'SRR11683995_count': ['alignments.csv', 'conversions.csv', 'conversions.idx', 'counts_TC.csv', 'rates.csv'],
    'SRR11683995_count_control': ['coverage.csv', 'snps.csv', 'counts_TC.csv', 'rates.csv'],
    'smartseq_count': ['alignments.csv', 'conversions.csv', 'conversions.idx', 'counts_TC.csv', 'rates.csv'],
    'nasc_count': ['alignments.csv', 'conversions.csv', 'conversions.idx', 'counts_TC.csv', 'rates.csv'],
    'SRR11683995_estimate': ['A_total_TC.csv', 'p_c_TC.csv', 'p_e.csv', 'pi_total_TC.csv', 'pi_c_total_TC.csv'],
    'SRR11683995_estimate_alpha': None,
    'SRR11683995_estimate_control': None,
    'SRR11683995_estimate_with_control': None,
    'smartseq_estimate': None,
    'nasc_estimate': ['A_transcriptome_TC.csv', 'p_c_TC.csv', 'p_e.csv', 'pi_transcriptome_TC.csv',
                      'pi_c_transcriptome_TC.csv'],
}


def copy_fixtures():
    for d, files in FIXTURE_FILES.items():
        src_dir = refstub.fixture(d)
        if files is None:
            files = [f for f in os.listdir(src_dir) if f.endswith('.csv')]
        for f in files:
            src = os.path.join(src_dir, f)
            if not os.path.exists(src):
                print('missing', src)
                continue
            dst_dir = os.path.join(GOLD, 'ref', d)
            os.makedirs(dst_dir, exist_ok=True)
            if f.endswith('.csv'):
                with open(src, 'rb') as fi, open(os.path.join(dst_dir, f + '.gz'), 'wb') as raw:
                    with gzip.GzipFile(fileobj=raw, mode='wb', mtime=0) as fo:
                        shutil.copyfileobj(fi, fo)
            else:
                shutil.copyfile(src, os.path.join(dst_dir, f))
    for f in ('p_e_int.csv', 'p_c_int.csv', 'pi_int.csv'):
        shutil.copyfile(refstub.fixture(f), os.path.join(GOLD, 'ref', f))


def gene_strands():
    import csv
    genes = refstub._read_pickle(refstub.fixture('SRR11683995_count', 'genes.pkl.gz'))
    wanted = set()
    for d in ('SRR11683995_count', 'smartseq_count', 'nasc_count', 'SRR11683995_count_control'):
        for name in ('alignments.csv', 'counts_TC.csv'):
            p = refstub.fixture(d, name)
            if os.path.exists(p):
                with open(p) as f:
                    for row in csv.DictReader(f):
                        wanted.add(row['GX'])
    out = {g: dict(strand=genes[g]['strand'], chromosome=genes[g]['chromosome'], start=genes[g]['segment'].start,
                   end=genes[g]['segment'].end, gene_name=genes[g].get('gene_name', ''))
           for g in sorted(wanted) if g in genes}
    with open(os.path.join(GOLD, 'gene_infos.json'), 'w') as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print('gene_infos', len(out))


def synth_hist(rng, full=False, small=False, unlabelled=False):
    R = int(rng.integers(30, 300)) if small else int(rng.integers(200, 6000))
    pi = 0.0 if unlabelled else rng.beta(2, 5)
    pcb = rng.uniform(0.02, 0.08)
    peb = rng.uniform(2e-4, 2e-3)
    n = np.clip(rng.poisson(23, R), 0, 150)
    lab = rng.random(R) < pi
    k = np.minimum(rng.binomial(n, np.where(lab, pcb, peb)), 30)
    if full:
        k[0], n[0] = 30, 150
    return k.astype(np.int32), n.astype(np.int32), peb


def em_golden():
    import dynast.estimation.p_c as pc
    rng = np.random.default_rng(20240)
    ks, ns, cs, off, pes, pcs, sts, nasc_pcs, nasc_sts = [], [], [], [0], [], [], [], [], []
    for i in range(160):
        k, n, pe = synth_hist(rng, full=(i % 16 == 0), small=(i % 5 == 1), unlabelled=(i % 10 == 3))
        if i == 77:
            k[1], n[1] = 70, 149   # k! wraps to 0 for k >= 66 -> ZeroDivisionError in the reference
        K, N = k.max() + 1, n.max() + 1
        v = np.zeros((K, N), dtype=np.int64)
        np.add.at(v, (k, n), 1)
        kk, nn = np.nonzero(v)
        # shuffle triplets and split some cells into duplicates (csr semantics: duplicates are summed)
        cnt = v[kk, nn]
        dup = cnt > 3
        kk2 = np.concatenate([kk, kk[dup]])
        nn2 = np.concatenate([nn, nn[dup]])
        cc2 = np.concatenate([np.where(dup, cnt - 2, cnt), np.full(dup.sum(), 2)])
        perm = rng.permutation(len(kk2))
        ks.append(kk2[perm]); ns.append(nn2[perm]); cs.append(cc2[perm])
        off.append(off[-1] + len(perm))
        pes.append(pe)
        try:
            pcs.append(pc.expectation_maximization(v, pe)); sts.append(0)
        except ValueError:
            pcs.append(np.nan); sts.append(1)
        except ZeroDivisionError:
            pcs.append(np.nan); sts.append(2)
        if i < 40:
            try:
                nasc_pcs.append(pc.expectation_maximization_nasc(v, pe)); nasc_sts.append(0)
            except ZeroDivisionError:
                nasc_pcs.append(np.nan); nasc_sts.append(2)
    pmf_args = [(k, n, p) for k in (0, 1, 2, 5, 9, 12, 20, 25, 30) for n in (0, 1, 5, 10, 50, 100, 150)
                for p in (0.001, 0.05, 0.5)]
    pmf = [pc.binomial_pmf(*a) for a in pmf_args]
    np.savez_compressed(
        os.path.join(GOLD, 'em_golden.npz'),
        k=np.concatenate(ks).astype(np.int32), n=np.concatenate(ns).astype(np.int32),
        count=np.concatenate(cs).astype(np.int64), off=np.asarray(off, dtype=np.int64),
        p_e=np.asarray(pes), p_c=np.asarray(pcs), status=np.asarray(sts, dtype=np.int32),
        nasc_p_c=np.asarray(nasc_pcs), nasc_status=np.asarray(nasc_sts, dtype=np.int32),
        pmf_args=np.asarray(pmf_args, dtype=np.float64), pmf=np.asarray(pmf),
    )
    print('em_golden: ok', sts.count(0), 'p_c<=p_e', sts.count(1), 'zerodiv', sts.count(2))


def _write_gz(path, text):
    with open(path, 'wb') as raw:
        with gzip.GzipFile(fileobj=raw, mode='wb', mtime=0) as fo:
            fo.write(text.encode())


def synth_counts_case(name, seed, n_reads, umi=True, dup_read_ids=False, with_snps=False, conversions=frozenset({'TC'}),
                      use_conversions=True, threshold=60):
    """Synthetic conversions.csv / alignments.csv / conversions.idx -> the REFERENCE's count_conversions, rates,
    aggregates, p_e and alpha on them (tie-heavy UMI groups, +/- genes, several splits)."""
    import pickle
    import tempfile
    import pandas as pd
    import dynast.config as config
    import dynast.preprocessing.conversion as conv
    import dynast.preprocessing.aggregation as agg
    import dynast.estimation.p_e as p_e
    import dynast.estimation.alpha as alpha
    rng = np.random.default_rng(seed)
    out_dir = os.path.join(GOLD, 'synth_counts', name)
    os.makedirs(out_dir, exist_ok=True)
    genes = {f'G{i:02d}': {'strand': '+' if i % 3 else '-', 'chromosome': f'chr{i % 4}'} for i in range(9)}
    gene_names = list(genes)
    bc_w = np.array([40, 25, 10, 8, 5, 4, 3, 2, 1.5, 1, 0.3, 0.2])
    bcs = [''.join(rng.choice(list('ACGT'), 8)) for _ in bc_w]
    ali_lines, conv_lines, index = [], [], []
    ali_pos = len('read_id,index,barcode,umi,GX,A,C,G,T,velocity,transcriptome,score\n')
    conv_pos = len('read_id,index,contig,genome_i,conversion,quality\n')
    letters = 'ACGT'
    prev_id = None
    for i in range(n_reads):
        rid = f'read{i:06d}'
        bc = bcs[rng.choice(len(bcs), p=bc_w / bc_w.sum())]
        if dup_read_ids and prev_id is not None and rng.random() < 0.03:
            rid, bc = prev_id, prev_bc   # a multimapping read keeps its barcode
        prev_id, prev_bc = rid, bc
        um = ''.join(rng.choice(list('ACGT'), 2)) if umi else 'NA'   # 16 UMIs -> many duplicates
        gx = gene_names[int(rng.integers(0, len(gene_names)))]
        content = rng.integers(8, 30, 4)
        vel = ['spliced', 'unspliced', 'ambiguous'][int(rng.integers(0, 3))]
        tr = bool(rng.random() < 0.6)
        score = int(rng.integers(55, 60))
        m = int(rng.poisson(0.9))
        if m:
            start = conv_pos
            for _ in range(m):
                a = int(rng.integers(0, 4))
                b = (a + int(rng.integers(1, 4))) % 4
                if rng.random() < 0.5:
                    a, b = (3, 1) if genes[gx]['strand'] == '+' else (0, 2)
                line = f'{rid},1,{genes[gx]["chromosome"]},{int(rng.integers(0, 400))},{letters[a]}{letters[b]},{int(rng.integers(20, 41))}\n'
                conv_lines.append(line)
                conv_pos += len(line)
            index.append((start, m, ali_pos))
        line = f'{rid},1,{bc},{um},{gx},{content[0]},{content[1]},{content[2]},{content[3]},{vel},{tr},{score}\n'
        ali_lines.append(line)
        ali_pos += len(line)
    ali_text = 'read_id,index,barcode,umi,GX,A,C,G,T,velocity,transcriptome,score\n' + ''.join(ali_lines)
    conv_text = 'read_id,index,contig,genome_i,conversion,quality\n' + ''.join(conv_lines)
    snps = None
    if with_snps:
        snps = {}
        for line in conv_lines[::7]:
            f = line.strip().split(',')
            snps.setdefault(f[4], {}).setdefault(f[2], set()).add(int(f[3]))
    tmp = tempfile.mkdtemp()
    ap, cp, ip, op = (os.path.join(tmp, x) for x in ('alignments.csv', 'conversions.csv', 'conversions.idx', 'counts.csv'))
    open(ap, 'w').write(ali_text)
    open(cp, 'w').write(conv_text)
    with gzip.open(ip, 'wb') as f:
        pickle.dump(index, f)
    old = config.COUNTS_SPLIT_THRESHOLD
    config.COUNTS_SPLIT_THRESHOLD = threshold
    try:
        conv.count_conversions(cp, ap, ip, op, genes, snps=snps, quality=27, conversions=conversions,
                               dedup_use_conversions=use_conversions, n_threads=2, temp_dir=tmp)
    finally:
        config.COUNTS_SPLIT_THRESHOLD = old
    counts_text = open(op).read()
    _write_gz(os.path.join(out_dir, 'alignments.csv.gz'), ali_text)
    _write_gz(os.path.join(out_dir, 'conversions.csv.gz'), conv_text)
    _write_gz(os.path.join(out_dir, 'counts.csv.gz'), counts_text)
    meta = dict(index=index, genes=genes, snps={c: {k: sorted(v) for k, v in d.items()} for c, d in (snps or {}).items()},
                conversions=sorted(conversions) if conversions is not None else None, use_conversions=use_conversions,
                threshold=threshold)
    with open(os.path.join(out_dir, 'meta.json'), 'w') as f:
        json.dump(meta, f)
    # downstream stages of the reference on its own counts
    df = conv.complement_counts(conv.read_counts(op), genes)
    convs = conversions if conversions is not None else frozenset({'TC'})
    agg.calculate_mutation_rates(df, os.path.join(tmp, 'rates.csv'), group_by='barcode')
    agg.aggregate_counts(df, os.path.join(tmp, 'A.csv'), conversions=convs)
    p_e.estimate_p_e(df, os.path.join(tmp, 'p_e.csv'), conversions=frozenset({convs}), group_by=['barcode'])
    pi_c = {bc: 0.1 + 0.05 * i for i, bc in enumerate(sorted(set(df['barcode'])))}
    pi_c[sorted(pi_c)[0]] = 0.0
    alpha.estimate_alpha(df, pi_c, os.path.join(tmp, 'alpha.csv'), convs, group_by=['barcode'], pi_c_group_by=['barcode'])
    for f in ('rates.csv', 'A.csv', 'p_e.csv', 'alpha.csv'):
        _write_gz(os.path.join(out_dir, f + '.gz'), open(os.path.join(tmp, f)).read())
    with open(os.path.join(out_dir, 'pi_c.json'), 'w') as f:
        json.dump(pi_c, f)
    print('synth_counts', name, len(counts_text.splitlines()) - 1, 'rows')


def synth_counts():
    synth_counts_case('umi_tc', 31, 2500, with_snps=True)
    synth_counts_case('umi_exon', 32, 2500, use_conversions=False)
    synth_counts_case('umi_dual', 33, 2000, conversions=frozenset({'TC', 'AG'}))
    synth_counts_case('umi_noconv', 34, 1500, conversions=None)
    synth_counts_case('noumi', 35, 1500, umi=False, conversions=None)
    synth_counts_case('noumi_dups', 36, 1500, umi=False, dup_read_ids=True)


def pi_golden():
    """Posterior means of pi.stan by QUADPACK (oracle/pi_oracle.py) on seeded histograms of several shapes."""
    from oracle import pi_oracle
    rng = np.random.default_rng(515)
    cases = []
    for reads, pi, pc, pe in [(16, 0.3, 0.05, 1e-3), (40, 0.0, 0.04, 5e-4), (60, 1.0, 0.06, 2e-3), (300, 0.2, 0.03, 1e-3),
                              (300, 0.85, 0.08, 2e-4), (2500, 0.5, 0.05, 1e-3), (2500, 0.02, 0.05, 1.5e-3), (20000, 0.3, 0.02, 2e-3)]:
        n = np.clip(rng.poisson(23, reads), 1, 150)
        lab = rng.random(reads) < pi
        k = rng.binomial(n, np.where(lab, pc, pe))
        key, cnt = np.unique(k.astype(np.int64) * 1024 + n, return_counts=True)
        cases.append((key // 1024, key % 1024, cnt, pe * rng.uniform(0.8, 1.2), pc * rng.uniform(0.8, 1.2)))
    ks, ns, cs, off, pes, pcs, out = [], [], [], [0], [], [], []
    for k, n, c, pe, pc in cases:
        out.append(pi_oracle.posterior_means(k, n, c, pe, pc))
        ks.append(k); ns.append(n); cs.append(c); off.append(off[-1] + len(k)); pes.append(pe); pcs.append(pc)
        print('pi_golden', len(out), out[-1])
    np.savez_compressed(os.path.join(GOLD, 'pi_golden.npz'), k=np.concatenate(ks).astype(np.int32),
                        n=np.concatenate(ns).astype(np.int32), count=np.concatenate(cs).astype(np.int64),
                        off=np.asarray(off, dtype=np.int64), p_e=np.asarray(pes), p_c=np.asarray(pcs),
                        alpha_beta_pi=np.asarray(out))


def adata_golden():
    """Layers of the reference's utils.results_to_adata (anndata replaced by a capturing stand-in) for the
    synthetic counts of synth_counts/umi_dual, with pi and with alpha estimates."""
    import io
    import dynast.utils as du
    import dynast.preprocessing.conversion as conv

    class Capture:
        def __init__(self, X=None, obs=None, var=None, layers=None):
            self.X, self.obs, self.var, self.layers = X, obs, var, layers

    du.anndata.AnnData = Capture
    name = 'umi_dual'
    with open(os.path.join(GOLD, 'synth_counts', name, 'meta.json')) as f:
        meta = json.load(f)
    with gzip.open(os.path.join(GOLD, 'synth_counts', name, 'counts.csv.gz'), 'rt') as f:
        df = conv.complement_counts(conv.read_counts(io.StringIO(f.read())), meta['genes'])
    conversions = frozenset({frozenset({'TC'}), frozenset({'AG'})})
    rng = np.random.default_rng(77)
    pairs = sorted(set(zip(df['barcode'], df['GX'])))
    pis = {key: {('TC',): {p: float(rng.random()) for p in pairs[::2]}, ('AG',): {p: float(rng.random()) for p in pairs[1::3]}}
           for key in ('transcriptome', 'total', 'spliced', 'unspliced')}
    barcodes = sorted(set(df['barcode']))
    alphas = {key: {('TC',): {b: float(rng.uniform(0.2, 1.3)) for b in barcodes[::2]}} for key in ('transcriptome', 'total', 'spliced')}
    out = {}
    for tag, kw in (('plain', {}), ('pi', dict(pis=pis)), ('alpha', dict(alphas=alphas))):
        ad = du.results_to_adata(df, conversions, gene_infos=None, **kw)
        out[f'{tag}/X'] = ad.X.toarray()
        for k, v in ad.layers.items():
            out[f'{tag}/layer/{k}'] = v.toarray()
        for c in ad.obs.columns:
            out[f'{tag}/obs/{c}'] = ad.obs[c].to_numpy(dtype=np.float64)
    np.savez_compressed(os.path.join(GOLD, 'adata_golden.npz'), **out)
    with open(os.path.join(GOLD, 'adata_golden_inputs.json'), 'w') as f:
        json.dump(dict(pis={k: {'_'.join(c): {'|'.join(p): v for p, v in d.items()} for c, d in byc.items()} for k, byc in pis.items()},
                       alphas={k: {'_'.join(c): d for c, d in byc.items()} for k, byc in alphas.items()}), f)
    print('adata_golden', len(out), 'arrays')


def smartseq_counts_tc():
    """counts_TC.csv of the Smart-seq fixture as `dynast count --conversion TC` writes it: count() passes
    conversions=['TC'] to count_conversions (count.py:306-318), which changes the tie order of the no-UMI path; the
    committed fixture was written with conversions=None (what the reference's own test_conversions_paired reproduces).
    The REFERENCE's count_conversions runs here on the fixture's conversions / alignments / index files."""
    import tempfile
    from dynast.preprocessing import conversion as conv
    import shutil
    genes = refstub.read_genes() if hasattr(refstub, 'read_genes') else None
    if genes is None:
        import pickle
        with gzip.open(refstub.fixture('genes.pkl.gz'), 'rb') as f:
            genes = pickle.load(f)
    tmp = tempfile.mkdtemp()
    src = refstub.fixture('smartseq_count')
    op = os.path.join(tmp, 'counts_TC.csv')
    conv.count_conversions(os.path.join(src, 'conversions.csv'), os.path.join(src, 'alignments.csv'), os.path.join(src, 'conversions.idx'),
                           op, genes, quality=27, conversions=['TC'], dedup_use_conversions=True, n_threads=2, temp_dir=tmp)
    out_dir = os.path.join(GOLD, 'smartseq_count_tc')
    os.makedirs(out_dir, exist_ok=True)
    _write_gz(os.path.join(out_dir, 'counts_TC.csv.gz'), open(op).read())
    print('smartseq_count_tc', len(open(op).read().splitlines()) - 1, 'rows')
    shutil.rmtree(tmp, ignore_errors=True)


def main():
    refstub.install()
    os.makedirs(GOLD, exist_ok=True)
    copy_fixtures()
    gene_strands()
    em_golden()
    synth_counts()
    pi_golden()
    adata_golden()
    smartseq_counts_tc()


if __name__ == '__main__':
    main()
