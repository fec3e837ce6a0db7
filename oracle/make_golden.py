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


def main():
    refstub.install()
    os.makedirs(GOLD, exist_ok=True)
    copy_fixtures()
    gene_strands()
    em_golden()


if __name__ == '__main__':
    main()
