"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's velocity (spliced / unspliced / ambiguous)
assignment and of the third-party pieces it stands on.

What is restated
  * dynast/preprocessing/bam.py:177-228  ``assign_velocity_type`` (gene search order, strand / range tests, the
    any_exon_only / any_exon_overlap / any_intron_overlap flags over a gene's transcripts, multi-gene -> dropped,
    ``strict_exon_overlap``) and bam.py:363-383 (read strand from the flags and the protocol strandedness).
  * ``ngs_tools.gtf`` (requirements.txt:2 ``ngs-tools>=1.8.0``; NOT vendored in /root/reference and not installable
    here): ``Segment``, ``SegmentCollection`` (sorted + collapsed segments, ``from_positions``, ``invert``,
    ``is_overlapping``, ``is_subset``) and ``genes_and_transcripts_from_gtf`` (0-based half-open coordinates:
    ``Segment(start - 1, end)``; gene / transcript / exon features; introns = exons inverted inside the transcript
    segment).  Call sites in the reference: count.py:146-147, bam.py:180-228.

Parity status: **unpinned beyond the fixture**.  ``tests/fixtures/transcripts.pkl.gz`` is a missing blob and the
reference's own velocity test (tests/preprocessing/test_bam.py:40-68) cannot run here, so the published
semantics of ngs_tools are restated from its documentation.  The anchor is the ``GX`` / ``velocity`` columns of
``SRR11683995_count/alignments.csv``: ``oracle/make_velocity_gtf.py`` synthesises an annotation whose gene
segments / strands / contigs are the ones of ``genes.pkl.gz`` and whose exon structure is consistent with every
row of that file, and this module has to reproduce the file byte for byte from the BAM plus that annotation
(tests/test_oracle_velocity.py).  Details of ngs_tools that the fixture cannot distinguish (touching segments are
NOT merged by ``collapse``; an empty collection has start = end = 0) are documented where they are decided.
The product never imports this file.
"""
import re
from typing import Dict, Iterable, List, Optional, Tuple


class Segment:
    """ngs_tools.gtf.Segment: a half-open interval [start, end)."""
    __slots__ = ('start', 'end')

    def __init__(self, start: int, end: int):
        assert end >= start
        self.start = start
        self.end = end

    def is_overlapping(self, other: 'Segment') -> bool:
        return self.start < other.end and self.end > other.start

    def is_subset(self, other: 'Segment') -> bool:
        return self.start >= other.start and self.end <= other.end

    def __iter__(self):
        return iter((self.start, self.end))

    def __lt__(self, other):
        return (self.start, self.end) < (other.start, other.end)

    def __eq__(self, other):
        return (self.start, self.end) == (other.start, other.end)

    def __repr__(self):
        return f'Segment({self.start}, {self.end})'


class SegmentCollection:
    """ngs_tools.gtf.SegmentCollection: sorted, non-overlapping segments."""

    def __init__(self, segments: Optional[Iterable[Segment]] = None):
        self.segments: List[Segment] = sorted(segments) if segments else []
        self.collapse()

    def collapse(self):
        """Overlapping segments are merged; segments that merely touch ([a, b) and [b, c)) stay apart."""
        out: List[Segment] = []
        for s in self.segments:
            if out and out[-1].is_overlapping(s):
                out[-1] = Segment(out[-1].start, max(out[-1].end, s.end))
            else:
                out.append(Segment(s.start, s.end))
        self.segments = out

    @property
    def start(self) -> int:
        return self.segments[0].start if self.segments else 0

    @property
    def end(self) -> int:
        return self.segments[-1].end if self.segments else 0

    def __iter__(self):
        return iter(self.segments)

    def __len__(self):
        return len(self.segments)

    def invert(self, bounds: Segment) -> 'SegmentCollection':
        """The parts of `bounds` not covered by the collection (the introns of a transcript)."""
        gaps = []
        prev = bounds.start
        for s in self.segments:
            if prev < s.start:
                gaps.append(Segment(prev, s.start))
            prev = max(prev, s.end)
        if prev < bounds.end:
            gaps.append(Segment(prev, bounds.end))
        return SegmentCollection(gaps)

    def is_overlapping(self, other: 'SegmentCollection') -> bool:
        return any(a.is_overlapping(b) for a in self.segments for b in other.segments)

    def is_subset(self, other: 'SegmentCollection') -> bool:
        """Every segment lies inside ONE segment of `other`."""
        return all(any(a.is_subset(b) for b in other.segments) for a in self.segments)

    @classmethod
    def from_positions(cls, positions) -> 'SegmentCollection':
        """Runs of consecutive positions (duplicates and order do not matter: a read pair contributes the
        concatenation of both mates' positions, bam.py:352)."""
        pos = sorted(set(positions))
        segs = []
        for p in pos:
            if segs and segs[-1][1] == p:
                segs[-1][1] = p + 1
            else:
                segs.append([p, p + 1])
        return cls(Segment(a, b) for a, b in segs)


_ATTR = re.compile(r'(\S+?)\s*"(.*?)"')


def parse_gtf(gtf_path: str, use_version: bool = False):
    """ngs_tools.gtf.genes_and_transcripts_from_gtf: (gene_infos, transcript_infos).

    gene_infos[gene_id] = {'segment', 'chromosome', 'strand', 'gene_name', 'transcripts': [ids in file order]}
    transcript_infos[transcript_id] = {'gene_id', 'transcript_name', 'segment', 'chromosome', 'strand', 'exons',
    'introns'}.  A transcript without a `transcript` line gets the span of its exons."""
    gene_infos: Dict[str, dict] = {}
    transcript_infos: Dict[str, dict] = {}
    exons: Dict[str, List[Segment]] = {}
    with open(gtf_path) as f:
        for line in f:
            if line.startswith('#') or not line.strip():
                continue
            cols = line.rstrip('\n').split('\t')
            if len(cols) < 9 or cols[2] not in ('gene', 'transcript', 'exon'):
                continue
            attrs = dict(_ATTR.findall(cols[8]))
            start, end = int(cols[3]) - 1, int(cols[4])
            gene_id = attrs.get('gene_id')
            if gene_id is None:
                continue
            if use_version and attrs.get('gene_version'):
                gene_id = f'{gene_id}.{attrs["gene_version"]}'
            tid = attrs.get('transcript_id')
            if tid is not None and use_version and attrs.get('transcript_version'):
                tid = f'{tid}.{attrs["transcript_version"]}'
            if cols[2] == 'gene':
                gene_infos.setdefault(gene_id, {'transcripts': []}).update(
                    segment=Segment(start, end), chromosome=cols[0], strand=cols[6], gene_name=attrs.get('gene_name', ''))
            elif cols[2] == 'transcript':
                g = gene_infos.setdefault(gene_id, {'transcripts': []})
                if tid not in g['transcripts']:
                    g['transcripts'].append(tid)
                transcript_infos.setdefault(tid, {}).update(
                    gene_id=gene_id, transcript_name=attrs.get('transcript_name', ''), segment=Segment(start, end),
                    chromosome=cols[0], strand=cols[6])
            else:
                exons.setdefault(tid, []).append(Segment(start, end))
                g = gene_infos.setdefault(gene_id, {'transcripts': []})
                if tid not in g['transcripts']:
                    g['transcripts'].append(tid)
                transcript_infos.setdefault(tid, {}).setdefault('gene_id', gene_id)
                transcript_infos[tid].setdefault('chromosome', cols[0])
                transcript_infos[tid].setdefault('strand', cols[6])
    for tid, info in transcript_infos.items():
        ex = SegmentCollection(exons.get(tid, []))
        if 'segment' not in info:
            info['segment'] = Segment(ex.start, ex.end)
        info['exons'] = ex
        info['introns'] = ex.invert(info['segment'])
    for gid, info in gene_infos.items():
        if 'segment' not in info:      # gene line missing: span of its transcripts
            ts = [transcript_infos[t] for t in info['transcripts']]
            info.update(segment=Segment(min(t['segment'].start for t in ts), max(t['segment'].end for t in ts)),
                        chromosome=ts[0]['chromosome'], strand=ts[0]['strand'], gene_name='')
    return gene_infos, transcript_infos


def read_strand(flag: int, strand: str) -> Optional[str]:
    """bam.py:363-380."""
    paired, reverse, read1 = bool(flag & 0x1), bool(flag & 0x10), bool(flag & 0x40)
    if paired:
        if read1:
            if strand == 'forward':
                return '+' if reverse else '-'
            if strand == 'reverse':
                return '-' if reverse else '+'
        else:
            if strand == 'forward':
                return '-' if reverse else '+'
            if strand == 'reverse':
                return '+' if reverse else '-'
        return None
    if strand == 'forward':
        return '-' if reverse else '+'
    if strand == 'reverse':
        return '+' if reverse else '-'
    return None


def gene_order(gene_infos: dict) -> List[str]:
    """bam.py:177-178: genes sorted by (start, end) of their segment (stable in dictionary order)."""
    return sorted(gene_infos.keys(), key=lambda g: tuple(gene_infos[g]['segment']))


def assign_velocity_type(positions, gene_infos: dict, transcript_infos: dict, order: List[str], gx: str = '',
                         strand: Optional[str] = None, strict_exon_overlap: bool = False) -> Tuple[Optional[str], Optional[str]]:
    """bam.py:180-228 (``gene_infos`` / ``order`` hold the genes of the read's contig only, bam.py:714-726)."""
    alignment = SegmentCollection.from_positions(positions)
    assigned_gene = gx
    assigned_type = None
    for gene in (order if not gx else [gx]):
        if not gx:
            if strand and strand != gene_infos[gene]['strand']:
                continue
            seg = gene_infos[gene]['segment']
            if not (alignment.start >= seg.start and alignment.end <= seg.end):
                continue
        any_exon_only = any_exon_overlap = any_intron_overlap = False
        for tid in gene_infos[gene]['transcripts']:
            t = transcript_infos[tid]
            if not any_exon_overlap:
                any_exon_overlap |= alignment.is_overlapping(t['exons'])
            if any_exon_overlap and not any_exon_only:
                any_exon_only |= alignment.is_subset(t['exons'])
            if not any_intron_overlap:
                any_intron_overlap |= alignment.is_overlapping(t['introns'])
        if not gx:
            if assigned_gene and (any_exon_overlap or any_intron_overlap):
                return None, None
            if not any_exon_overlap and not any_intron_overlap:
                continue
        if any_exon_only and (not strict_exon_overlap or not any_intron_overlap):
            assigned_gene, assigned_type = gene, 'spliced'
        elif any_intron_overlap and (not strict_exon_overlap or not any_exon_only):
            assigned_gene, assigned_type = gene, 'unspliced'
        else:
            assigned_gene, assigned_type = gene, 'ambiguous'
    return assigned_gene, assigned_type


def by_contig(gene_infos: dict) -> Dict[str, dict]:
    """bam.py:702-706, 719-721: the per-contig gene dictionaries parse_read_contig receives."""
    out: Dict[str, dict] = {}
    for gid, info in gene_infos.items():
        out.setdefault(info['chromosome'], {})[gid] = info
    return out
