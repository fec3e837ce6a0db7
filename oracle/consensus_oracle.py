"""TEST INFRASTRUCTURE ONLY -- CPU restatement of consensus calling for one UMI group.

Follows dynast/preprocessing/consensus.py:57-200 (`call_consensus_from_reads`): L x 4 quality-sum matrix over
the group's span, first-seen reference base (deletions included), per position N / D / M with the tie rules
of :137-146, quality clipped to 42, run-length CIGAR, incrementally built MD and NM.
Pinned by the reference's known-answer test (tests/preprocessing/test_consensus.py:13-80: 6 reads ->
ACTGCCAGCTGC, start 1, end 14, 10M1D2M, MD 5T4^A2) in tests/test_oracle_consensus.py.
Reads are `oracle.bamio.Read` objects (pysam-like accessors)."""
from typing import List

import numpy as np

from . import bamio

BASES = ('A', 'C', 'G', 'T')
BASE_IDX = {b: i for i, b in enumerate(BASES)}


def make_read(name: str, pos: int, cigar, seq: str, qual, md: str, ref_id: int = 0, flag: int = 0, tags=None) -> bamio.Read:
    """cigar: list of (op, len) with op as BAM code or as a letter of 'MIDNSHP=X'."""
    r = bamio.Read()
    r.query_name = name
    r.flag = flag
    r.reference_id = ref_id
    r.reference_name = str(ref_id)
    r.reference_start = pos
    r.mapping_quality = 255
    r.cigar = [(bamio.CIGAR_OPS.index(op) if isinstance(op, str) else op, n) for op, n in cigar]
    r.next_reference_id = -1
    r.next_reference_start = -1
    r.template_length = 0
    r.query_sequence = seq
    r.query_qualities = list(qual)
    r.tags = dict(tags or {})
    r.tags['MD'] = md
    return r


def call_consensus(reads: List[bamio.Read], quality: int = 27):
    """Returns dict(seq, qual, start, end, cigar [(op letter, n)], md, nm)."""
    left_pos = min(r.reference_start for r in reads)
    right_pos = max(r.reference_end for r in reads)
    length = right_pos - left_pos
    sequence = np.zeros((length, 4), dtype=np.uint32)
    reference = np.full(length, -1, dtype=np.int8)
    for read in reads:
        read_sequence = read.query_sequence.upper()
        read_qualities = read.query_qualities
        for read_i, genome_i, genome_base in read.get_aligned_pairs(matches_only=False, with_seq=True):
            if genome_i is None or genome_base is None:
                continue
            i = genome_i - left_pos
            genome_base = genome_base.upper()
            if genome_base == 'N':
                continue
            if read_i is None:
                if reference[i] < 0:
                    reference[i] = BASE_IDX[genome_base]
                continue
            read_base = read_sequence[read_i]
            if read_base == 'N':
                continue
            if reference[i] < 0:
                reference[i] = BASE_IDX[genome_base]
            sequence[i, BASE_IDX[read_base]] += read_qualities[read_i]
    consensus, qualities, cigar, md = [], [], [], []
    last_op, cigar_n = None, 0
    md_n, md_zero, md_del, nm = 0, True, False, 0
    for i in range(length):
        ref = reference[i]
        op = 'N'
        if ref >= 0:
            seq = sequence[i]
            if (seq == 0).all():
                op = 'D'
                if md_n > 0 or md_zero:
                    md.append(str(md_n))
                    md_n = 0
                if not md_del:
                    md.append('^')
                md.append(BASES[ref])
                md_del = True
            else:
                md_del = False
                base_q = seq.max()
                if base_q < quality:
                    base = ref
                else:
                    bases = (seq == base_q).nonzero()[0]
                    base = ref if ref in bases else bases[0]
                op = 'M'
                if ref == base:
                    md_n += 1
                    md_zero = False
                else:
                    if md_n > 0 or md_zero:
                        md.append(str(md_n))
                        md_n = 0
                    md.append(BASES[ref])
                    md_zero = True
                    nm += 1
                consensus.append(BASES[base])
                qualities.append(int(min(base_q, 42)))
        if op == last_op:
            cigar_n += 1
        else:
            if last_op:
                cigar.append((last_op, cigar_n))
            last_op, cigar_n = op, 1
    md.append(str(md_n))
    cigar.append((last_op, cigar_n))
    return dict(seq=''.join(consensus), qual=qualities, start=left_pos, end=right_pos, cigar=cigar, md=''.join(md), nm=nm)


# ------------------------------------------------------------------------------------------------
# The whole consensus caller over a BAM (dynast/preprocessing/consensus.py:227-494): read filter, mate pairing,
# gene lookup, (gene, barcode, UMI) groups, the periodic flush with its singleton quirk, tags, names, strands.
# Returns the records in the order the reference WRITES them (before `samtools sort`).  Each record is a dict:
#   kind 'read' (a read of the input written through; `tags` after swap_gene_tags when `swapped`) or 'consensus'.
# Grouping parity is "unpinned" in SURVEY.md (no fixture holds a consensus.bam); this restatement follows the source
# line by line and is what tests/test_consensus_operator.py compares the product's BAM with.
# ------------------------------------------------------------------------------------------------
def _read_strand(read, strand):
    if read.is_paired:
        if read.is_read1:
            if strand == 'forward':
                return '+' if read.is_reverse else '-'
            if strand == 'reverse':
                return '-' if read.is_reverse else '+'
        else:
            if strand == 'forward':
                return '-' if read.is_reverse else '+'
            if strand == 'reverse':
                return '+' if read.is_reverse else '-'
        return None
    if strand == 'forward':
        return '-' if read.is_reverse else '+'
    if strand == 'reverse':
        return '+' if read.is_reverse else '-'
    return None


def call_consensus_bam(reads, gene_infos, strand='forward', umi_tag=None, barcode_tag=None, gene_tag='GX', barcodes=None,
                       quality=27, add_RS_RI=False):
    from hashlib import sha256

    def seg(g):
        s = gene_infos[g]['segment'] if 'segment' in gene_infos[g] else (gene_infos[g]['start'], gene_infos[g]['end'])
        return tuple(s)

    contig_gene_order = {}
    for gene_id, info in gene_infos.items():
        contig_gene_order.setdefault(info['chromosome'], []).append(gene_id)
    for contig in contig_gene_order:
        contig_gene_order[contig] = sorted(contig_gene_order[contig], key=seg)

    def find_genes(contig, start, end, read_strand):
        genes = []
        for gene in contig_gene_order.get(contig, []):
            if read_strand and read_strand != gene_infos[gene]['strand']:
                continue
            a, b = seg(gene)
            if end <= a:
                break
            if start >= a and end <= b:
                genes.append(gene)
        return genes

    def swap_gene_tags(read, gene):
        tags = dict(read.tags)
        if gene_tag and gene_tag in tags:
            del tags[gene_tag]
        tags['GX'] = gene
        gn = gene_infos.get(gene, {}).get('gene_name')
        if gn:
            tags['GN'] = gn
        return dict(kind='read', read=read, tags=tags, swapped=True)

    def consensus_record(barcode, umi, group, gene):
        info = gene_infos[gene]
        tags = {'AS': sum(r.get_tag('AS') for r in group), 'NH': 1, 'HI': 1, 'RN': len(group)}
        if barcode_tag:
            tags[barcode_tag] = barcode
        if umi_tag:
            tags[umi_tag] = umi
        if gene_tag:
            gene_id = next((r.get_tag(gene_tag) for r in group if r.has_tag(gene_tag)), None)
            if gene_id:
                tags['GX'] = gene_id
                if info.get('gene_name'):
                    tags['GN'] = info['gene_name']
        if add_RS_RI:
            tags['RS'] = ';'.join(r.query_name for r in group)
            tags['RI'] = ';'.join(str(r.get_tag('HI')) for r in group)
        gene_strand = info['strand']
        cons_strand = gene_strand
        if strand == 'reverse':
            cons_strand = '-' if gene_strand == '+' else '+'
        if len({r.reference_id for r in group}) > 1:
            raise Exception('Can not call consensus from reads mapping to multiple contigs.')
        c = call_consensus(group, quality=quality)
        tags.update({'MD': c['md'], 'NM': c['nm']})
        return dict(kind='consensus', name=sha256(''.join(r.query_name for r in group).encode('utf-8')).hexdigest(), seq=c['seq'],
                    qual=c['qual'], reference_id=group[0].reference_id, reference_name=group[0].reference_name, start=c['start'],
                    cigar=c['cigar'], tags=tags, reverse=cons_strand == '-', mapq=255)

    required = [t for t in (umi_tag, barcode_tag) if t]
    out = []
    groups = {}
    paired = {}
    for i, read in enumerate(r for r in reads if r.reference_id >= 0):          # bam.fetch(): the placed records
        if read.is_secondary or read.is_unmapped or any(not read.has_tag(t) for t in required):
            continue
        barcode = read.get_tag(barcode_tag) if barcode_tag else None
        if barcode == '-' or (barcodes and barcode not in barcodes):
            continue
        contig = read.reference_name
        umi = read.get_tag(umi_tag) if umi_tag else None
        start, end = read.reference_start, read.reference_end
        key = (read.query_name, read.get_tag('HI'))
        mate = None
        if read.is_paired:
            if key not in paired:
                paired[key] = read
                continue
            mate = paired.pop(key)
            if not umi:
                start = mate.reference_start
                umi = (start, end)
        read_strand = _read_strand(read, strand)
        gx_assigned = read.has_tag(gene_tag) if gene_tag else False
        genes = [read.get_tag(gene_tag)] if gx_assigned else find_genes(contig, start, end, read_strand)
        if len(genes) != 1:
            out.append(dict(kind='read', read=read, tags=dict(read.tags), swapped=False))
            if read.is_paired:
                out.append(dict(kind='read', read=mate, tags=dict(mate.tags), swapped=False))
            continue
        groups.setdefault(genes[0], {}).setdefault(barcode, {}).setdefault(umi, []).append(read)
        if read.is_paired:
            groups[genes[0]][barcode][umi].append(mate)
        if i % 10000 == 0:
            leftmost = start if not paired else next(iter(paired.values())).reference_start
            for gene in list(groups.keys()):
                info = gene_infos[gene]
                if info['chromosome'] < contig or (info['chromosome'] == contig and seg(gene)[1] <= leftmost):
                    for bc, umi_groups in groups.pop(gene).items():
                        for u, rs in umi_groups.items():
                            if len(rs) == 1:
                                out.append(swap_gene_tags(rs[0], gene))          # no `continue`: a consensus of one follows
                            out.append(consensus_record(bc, u, rs, gene))
                else:
                    break
    for gene, bc_groups in groups.items():
        for bc, umi_groups in bc_groups.items():
            for u, rs in umi_groups.items():
                if len(rs) == 1:
                    out.append(swap_gene_tags(rs[0], gene))
                    continue
                out.append(consensus_record(bc, u, rs, gene))
    return out
