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
