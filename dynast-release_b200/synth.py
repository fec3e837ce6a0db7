"""Synthetic scSLAM-seq-like reads (SURVEY.md section 8d, config C2 generative model; follows the
simulator in dynast/benchmarking/simulation.py:22-79 for the labelling model).

This is the slow numpy/Python generator used for tests and small inputs; the bench generates the
full-size configs on the GPU with the same per-read decisions (csrc/synth.cu)."""
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import records

BASES = np.array(list('ACGT'))


def _complex_cigar(rng, L):
    """Hard clips, soft clips at either end, several M/=/X pieces separated by I, D and N -- including an insertion
    right after the leading clip and right before the trailing one."""
    while True:
        lead = int(rng.integers(1, 15)) if rng.random() < 0.4 else 0
        trail = int(rng.integers(1, 15)) if rng.random() < 0.4 else 0
        n_seg = int(rng.integers(1, 5))
        gaps = [int(rng.choice([1, 1, 2, 3])) for _ in range(n_seg - 1)]
        gap_len = [int(rng.integers(1, 5)) if g != 3 else int(rng.integers(20, 3000)) for g in gaps]
        edge_ins = [int(rng.integers(1, 4)) if rng.random() < 0.15 else 0 for _ in range(2)]
        body = L - lead - trail - sum(n for g, n in zip(gaps, gap_len) if g == 1) - sum(edge_ins)
        if body >= n_seg:
            break
    cuts = np.sort(rng.choice(np.arange(1, body), n_seg - 1, replace=False)) if n_seg > 1 else np.array([], dtype=int)
    pieces = np.diff(np.concatenate([[0], cuts, [body]])).astype(int)
    cigar = []
    if rng.random() < 0.2:
        cigar.append((5, int(rng.integers(1, 10))))
    if lead:
        cigar.append((4, lead))
    if edge_ins[0]:
        cigar.append((1, edge_ins[0]))
    for j, m in enumerate(pieces):
        cigar.append((int(rng.choice([0, 0, 0, 7, 8])), int(m)))
        if j < n_seg - 1:
            cigar.append((gaps[j], gap_len[j]))
    if edge_ins[1]:
        cigar.append((1, edge_ins[1]))
    if trail:
        cigar.append((4, trail))
    if rng.random() < 0.2:
        cigar.append((5, int(rng.integers(1, 10))))
    return cigar


def make_reads(
    n_reads: int,
    seed: int = 2001,
    read_len: int = 91,
    n_barcodes: int = 100,
    n_genes: int = 200,
    n_contigs: int = 25,
    p_c: float = 0.05,
    p_e: float = 5e-4,
    frac_n: float = 0.001,
    umi_len: int = 12,
    reads_per_umi: float = 2.0,
    complex_cigars: bool = False,
    lean: bool = False,
):
    """Returns dict(blob, rec_off, barcode_id, umi, gene_id, score, strand_gene, truth...) with one
    single-end read per counting unit.  CIGAR mix: 70 % LM, 18 % aMbNcM, 8 % sS(L-s)M, 2 % one 1-3 nt D,
    2 % one 1-3 nt I.  Reference bases iid uniform; labelled reads convert T->C per T w.p. p_c; every
    base is substituted to each specific other base w.p. p_e; Phred 90 % in 36..40, 10 % in 2..27."""
    rng = np.random.default_rng(seed)
    units = []
    raw = []
    L = read_len
    gene_strand = rng.integers(0, 2, n_genes)
    gene_contig = rng.integers(0, n_contigs, n_genes)
    gene_pi = rng.beta(2, 5, n_genes)
    bc_w = rng.lognormal(np.log(5000), 0.6, n_barcodes)
    bc_w /= bc_w.sum()
    gene_w = 1.0 / np.arange(1, n_genes + 1)**1.1
    gene_w /= gene_w.sum()
    n_groups = max(1, int(n_reads / reads_per_umi))
    grp_bc = rng.choice(n_barcodes, n_groups, p=bc_w)
    grp_gene = rng.choice(n_genes, n_groups, p=gene_w)
    grp_umi = rng.integers(0, 4**umi_len, n_groups)
    grp_of_read = rng.integers(0, n_groups, n_reads)
    barcode_id = grp_bc[grp_of_read].astype(np.uint32)
    gene_id = grp_gene[grp_of_read].astype(np.uint32)
    umi = grp_umi[grp_of_read].astype(np.uint64)
    pos = np.sort(rng.integers(0, 1 << 26, n_reads)).astype(np.int64)
    score = np.zeros(n_reads, dtype=np.int32)
    labelled = rng.random(n_reads) < gene_pi[gene_id]
    kind = rng.choice(5, n_reads, p=[0.70, 0.18, 0.08, 0.02, 0.02])
    for i in range(n_reads):
        k = kind[i]
        if complex_cigars:
            cigar = _complex_cigar(rng, L)
        elif k == 0:
            cigar = [(0, L)]
        elif k == 1:
            a = int(rng.integers(10, L - 10))
            cigar = [(0, a), (3, int(rng.integers(50, 5000))), (0, L - a)]
        elif k == 2:
            s = int(rng.integers(1, 20))
            cigar = [(4, s), (0, L - s)]
        elif k == 3:
            a = int(rng.integers(10, L - 10))
            cigar = [(0, a), (2, int(rng.integers(1, 4))), (0, L - a)]
        else:
            a = int(rng.integers(10, L - 10))
            ins = int(rng.integers(1, 4))
            cigar = [(0, a), (1, ins), (0, L - a - ins)]
        n_ref = sum(n for op, n in cigar if op in (0, 2, 7, 8))
        ref = rng.integers(0, 4, n_ref)
        qual = np.where(rng.random(L) < 0.9, rng.integers(36, 41, L), rng.integers(2, 28, L))
        seq = []
        md = []
        run = 0
        ri = 0
        nm = 0
        strand_rev = gene_strand[gene_id[i]] == 1
        conv_from, conv_to = (0, 2) if strand_rev else (3, 1)   # A>G on reverse-strand genes, T>C on forward
        for op, n in cigar:
            if op in (0, 7, 8):
                for _ in range(n):
                    g = int(ref[ri])
                    ri += 1
                    b = g
                    if labelled[i] and g == conv_from and rng.random() < p_c:
                        b = conv_to
                    u = rng.random()
                    if u < 3 * p_e:
                        b = (g + 1 + int(u / p_e)) % 4
                    if rng.random() < frac_n:
                        seq.append('N')
                        md.append(str(run))
                        md.append('ACGT'[g])
                        run = 0
                        nm += 1
                        continue
                    seq.append('ACGT'[b])
                    if b == g:
                        run += 1
                    else:
                        md.append(str(run))
                        md.append('ACGT'[g])
                        run = 0
                        nm += 1
            elif op == 2:
                md.append(str(run))
                run = 0
                md.append('^' + ''.join('ACGT'[int(x)] for x in ref[ri:ri + n]))
                ri += n
            elif op in (1, 4):
                seq.extend('ACGT'[int(x)] for x in rng.integers(0, 4, n))
        md.append(str(run))
        flag = 16 if rng.random() < 0.5 else 0
        score[i] = L - 2 * nm
        units.append([records.pack_one(int(pos[i]), int(gene_contig[gene_id[i]]), flag, cigar, ''.join(seq), qual, ''.join(md), lean=lean)])
        raw.append((int(pos[i]), cigar, ''.join(seq), [int(x) for x in qual], ''.join(md)))
    blob, rec_off = records.pack_units(units)
    return dict(
        blob=blob,
        rec_off=rec_off,
        barcode_id=barcode_id,
        umi=umi,
        gene_id=gene_id,
        score=score,
        gene_strand=gene_strand.astype(np.uint8),
        labelled=labelled,
        raw=raw,
    )


# ------------------------------------------------------------------------------------------------
# Synthetic BAM files (bench / test utility): the device generator's reads written through dyn_bam_write
# ------------------------------------------------------------------------------------------------
def _digits(values: np.ndarray, width: int, base: int, alphabet: bytes) -> np.ndarray:
    """[n, width] uint8 matrix of the base-`base` digits of `values`, most significant first, drawn from `alphabet`."""
    v = values.astype(np.uint64)
    out = np.empty((len(v), width), dtype=np.uint8)
    table = np.frombuffer(alphabet, dtype=np.uint8)
    for j in range(width - 1, -1, -1):
        out[:, j] = table[(v % np.uint64(base)).astype(np.int64)]
        v //= np.uint64(base)
    return out


def barcode_strings(ids: np.ndarray, width: int = 16) -> np.ndarray:
    """[n, width] ACGT letters of barcode ids (a multiplicative hash spreads consecutive ids over the letters)."""
    return _digits((ids.astype(np.uint64) * np.uint64(2654435761)) & np.uint64(4**width - 1), width, 4, b'ACGT')


def gene_names(n_genes: int):
    return [f'ENSG{g:011d}' for g in range(n_genes)]


def write_synth_bam(path: str, batch, n_contigs: int = 25, level: int = 6, n_threads: int = 8, index: bool = False) -> Dict:
    """A coordinate-sorted BAM of a pipeline.SynthBatch (full record layout, with keys): Illumina-style read names,
    tags NH HI AS nM CB UB GX (and the MD tag dyn_bam_write rebuilds).  Returns dict(n_records, gene_ids, refs)."""
    from . import bamio
    assert not batch.lean and batch.barcode is not None, 'write_synth_bam needs full records with keys'
    n = batch.n_reads
    blob = batch.records.cpu().numpy()
    rec_off = batch.rec_off.cpu().numpy().view(np.uint32)
    barcode = batch.barcode.cpu().numpy().astype(np.int64)
    umi = batch.umi.cpu().numpy().view(np.uint64)
    gene = batch.gene.cpu().numpy().astype(np.int64)
    score = np.clip(batch.score.cpu().numpy(), 0, 255).astype(np.uint8)
    ref_id = blob[rec_off[:-1].astype(np.int64) * 16 + 4].astype(np.int64)       # low byte of the header's ref_id (< 256 contigs)
    pos = blob.view(np.int32)[rec_off[:-1].astype(np.int64) * 4]
    if bool((np.diff(pos) >= 0).all()):
        order = np.argsort(ref_id, kind='stable').astype(np.int64)               # positions already ascend with the index
    else:                                                                        # clustered batches (UMI groups around random anchors)
        order = np.lexsort((pos, ref_id)).astype(np.int64)
    # read names: SYN:<index>:<tile>:<x>:<y>
    idx = np.arange(n, dtype=np.uint64)
    name = np.concatenate([np.broadcast_to(np.frombuffer(b'SYN0001:7:HXXXXXXXX:1:', np.uint8), (n, 22)), _digits(idx % 2000 + 1101, 4, 10, b'0123456789'),
                           np.broadcast_to(np.frombuffer(b':', np.uint8), (n, 1)), _digits((idx * 7919) % 30000, 5, 10, b'0123456789'),
                           np.broadcast_to(np.frombuffer(b':', np.uint8), (n, 1)), _digits(idx, 9, 10, b'0123456789')], axis=1)
    name_off = (np.arange(n + 1, dtype=np.uint64) * name.shape[1]).astype(np.uint32)
    const = lambda b: np.broadcast_to(np.frombuffer(b, np.uint8), (n, len(b)))
    aux = np.concatenate([const(b'NHC\x01HIC\x01ASC'), score[:, None], const(b'nMC\x00CBZ'), barcode_strings(barcode), const(b'\x00UBZ'),
                          _digits(umi & np.uint64(4**12 - 1), 12, 4, b'ACGT'), const(b'\x00GXZENSG'), _digits(gene.astype(np.uint64), 11, 10, b'0123456789'),
                          const(b'\x00')], axis=1)
    aux_off = (np.arange(n + 1, dtype=np.uint64) * aux.shape[1]).astype(np.uint32)
    refs = [(f'chr{c + 1}', 1 << 28) for c in range(n_contigs)]
    bamio.write_bam(path, refs, blob, rec_off, order=order, names=(np.ascontiguousarray(name).tobytes(), name_off),
                    aux=(np.ascontiguousarray(aux).tobytes(), aux_off), mapq=np.full(n, 255, np.uint8), level=level, n_threads=n_threads,
                    index=index)
    return dict(n_records=n, gene_ids=gene_names(batch.n_genes), refs=refs, order=order)
