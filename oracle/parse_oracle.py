"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's BAM -> conversions/alignments parse.

Follows dynast/preprocessing/bam.py:121-444 (`parse_read_contig`) and :648-780 (`parse_all_reads`,
single-process arm: contigs in header order, files concatenated, index rebased), on top of the
pysam-semantics reader in ``oracle/bamio.py``.  Velocity assignment (bam.py:180-228) needs
``ngs_tools`` segment collections and the missing ``transcripts.pkl.gz`` blob, so it is injected:
``velocity_lookup`` maps (read_id, HI) -> (gene, assignment) and reads absent from it are dropped,
exactly what ``if not assignment or not gx: continue`` (bam.py:382-383) does.

Pinned by byte equality with the reference fixtures (tests/test_oracle_parse.py):
  SRR11683995_count/{alignments.csv,conversions.csv,conversions.idx}  (velocity injected)
  smartseq_count/...  and  nasc_count/...                            (velocity=False, fully computed)
The product path never imports this file.
"""
from collections import Counter, OrderedDict
from typing import Dict, Iterable, List, Optional, Tuple

from . import bamio

CONVERSIONS_HEADER = 'read_id,index,contig,genome_i,conversion,quality\n'
ALIGNMENTS_HEADER = 'read_id,index,barcode,umi,GX,A,C,G,T,velocity,transcriptome,score\n'


def _skip(read: bamio.Read, required_tags: List[str]) -> bool:
    # bam.py:173-174 -- supplementary alignments are NOT filtered
    if read.is_secondary or read.is_unmapped or read.is_duplicate:
        return True
    return any(not read.has_tag(t) for t in required_tags)


def parse_contig(
    reads: Iterable[bamio.Read],
    contig: str,
    umi_tag: Optional[str] = None,
    barcode_tag: Optional[str] = None,
    gene_tag: str = 'GX',
    barcodes: Optional[set] = None,
    nasc: bool = False,
    velocity: bool = True,
    velocity_lookup: Optional[Dict[Tuple[str, int], Tuple[str, str]]] = None,
    annotation=None,
    strand: str = 'forward',
    strict_exon_overlap: bool = False,
):
    """Returns (conversion_lines, alignment_lines, index) for one contig; index offsets are local.
    `annotation` = (gene_infos of this contig, transcript_infos): velocity is computed by
    oracle/velocity_oracle.py (bam.py:177-228, 363-383) instead of being looked up."""
    if velocity and annotation is not None:
        from . import velocity_oracle
        v_genes, v_transcripts = annotation
        v_order = velocity_oracle.gene_order(v_genes)
    required = []
    if not velocity:
        required.append(gene_tag)
    if umi_tag:
        required.append(umi_tag)
    if barcode_tag:
        required.append(barcode_tag)

    conv_lines: List[str] = []
    align_lines: List[str] = []
    index = []
    conv_bytes = 0
    align_bytes = 0
    pending: Dict[Tuple[str, int], bamio.Read] = {}

    for read in reads:
        if _skip(read, required):
            continue
        barcode = read.get_tag(barcode_tag) if barcode_tag else 'NA'
        if barcode == '-' or (barcodes and barcode not in barcodes):
            continue
        read_id = read.query_name
        hi = read.get_tag('HI')
        score = read.get_tag('AS')
        ref_positions = read.get_reference_positions()
        content = Counter()
        mate_calls: 'OrderedDict[int, Tuple[int, str, str]]' = OrderedDict()

        if read.is_paired:
            key = (read_id, hi)
            if key not in pending:
                pending[key] = read
                continue
            mate = pending.pop(key)
            mate_seq = mate.query_sequence.upper()
            mate_q = mate.query_qualities
            mate_positions = mate.get_reference_positions()
            content.update(mate.get_reference_sequence().upper())
            shared = set(mate_positions) & set(ref_positions)
            overlap = Counter()
            for qi, gi, gb in mate.get_aligned_pairs(matches_only=True, with_seq=True):
                rb = mate_seq[qi]
                gb = gb.upper()
                if gb == 'N' or rb == 'N':
                    continue
                if gi in shared:
                    overlap[gb] += 1
                if gb != rb and (not nasc or gi not in shared):
                    mate_calls[gi] = (mate_q[qi], gb, rb)
            content = content - overlap  # Counter subtraction floors at zero (bam.py:346)
            if velocity:
                ref_positions = ref_positions + mate_positions

        umi = read.get_tag(umi_tag) if umi_tag else 'NA'
        gx_assigned = read.has_tag(gene_tag)
        gx = read.get_tag(gene_tag) if gx_assigned else ''
        seq = read.query_sequence.upper()
        quals = read.query_qualities
        content.update(read.get_reference_sequence().upper())

        assignment = 'unassigned'
        if velocity and annotation is not None:
            gx, assignment = velocity_oracle.assign_velocity_type(
                ref_positions, v_genes, v_transcripts, v_order, gx=gx,
                strand=velocity_oracle.read_strand(read.flag, strand), strict_exon_overlap=strict_exon_overlap)
            if not assignment or not gx:
                continue
        elif velocity:
            hit = (velocity_lookup or {}).get((read_id, hi))
            if hit is None:
                continue
            gx, assignment = hit

        n_lines = 0
        start_bytes = conv_bytes
        for qi, gi, gb in read.get_aligned_pairs(matches_only=True, with_seq=True):
            rb = seq[qi]
            gb = gb.upper()
            if gb == 'N' or rb == 'N':
                continue
            q = quals[qi]
            if gi in mate_calls:
                if not nasc:
                    mq, _, mrb = mate_calls[gi]
                    if mq > q:
                        rb, q = mrb, mq
                del mate_calls[gi]
            if gb != rb:
                line = f'{read_id},{hi},{contig},{gi},{gb}{rb},{q}\n'
                conv_lines.append(line)
                conv_bytes += len(line)
                n_lines += 1
        for gi, (q, gb, rb) in mate_calls.items():
            line = f'{read_id},{hi},{contig},{gi},{gb}{rb},{q}\n'
            conv_lines.append(line)
            conv_bytes += len(line)
            n_lines += 1
        if n_lines > 0:
            index.append((start_bytes, n_lines, align_bytes))
        line = (
            f'{read_id},{hi},{barcode},{umi},{gx},{content["A"]},{content["C"]},{content["G"]},{content["T"]},'
            f'{assignment},{gx_assigned},{score}\n'
        )
        align_lines.append(line)
        align_bytes += len(line)
    return conv_lines, align_lines, index


def parse_all_reads(
    bam_path: str,
    conversions_path: str,
    alignments_path: str,
    umi_tag=None,
    barcode_tag=None,
    gene_tag='GX',
    barcodes=None,
    nasc=False,
    velocity=True,
    velocity_lookup=None,
    annotation=None,
    strand='forward',
    strict_exon_overlap=False,
):
    """Whole-BAM driver (bam.py:648-780 with n_threads=1). Writes both CSVs, returns the index list."""
    genes_by_contig = None
    if annotation is not None:
        from . import velocity_oracle
        genes_by_contig = velocity_oracle.by_contig(annotation[0])
    _, refs, reads = bamio.read_bam(bam_path)
    per_contig: Dict[int, List[bamio.Read]] = {}
    for r in reads:
        if r.reference_id >= 0:
            per_contig.setdefault(r.reference_id, []).append(r)
    index = []
    with open(conversions_path, 'w') as fc, open(alignments_path, 'w') as fa:
        fc.write(CONVERSIONS_HEADER)
        fa.write(ALIGNMENTS_HEADER)
        cpos, apos = len(CONVERSIONS_HEADER), len(ALIGNMENTS_HEADER)
        for ref_id in sorted(per_contig):
            conv, ali, idx = parse_contig(
                per_contig[ref_id],
                refs[ref_id][0],
                umi_tag=umi_tag,
                barcode_tag=barcode_tag,
                gene_tag=gene_tag,
                barcodes=set(barcodes) if barcodes else None,
                nasc=nasc,
                velocity=velocity,
                velocity_lookup=velocity_lookup,
                annotation=None if annotation is None else (genes_by_contig.get(refs[ref_id][0], {}), annotation[1]),
                strand=strand,
                strict_exon_overlap=strict_exon_overlap,
            )
            fc.write(''.join(conv))
            fa.write(''.join(ali))
            index.extend((cpos + p, n, apos + a) for p, n, a in idx)
            cpos += sum(len(s) for s in conv)
            apos += sum(len(s) for s in ali)
    return index
