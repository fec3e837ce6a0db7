"""Gene / transcript annotation for velocity assignment: GTF -> the dictionaries `dynast count` builds with
``ngs_tools.gtf.genes_and_transcripts_from_gtf`` (count.py:146-147) -> flat interval tables in HBM for
``dyn_velocity`` (csrc/velocity.cu, bam.py:177-228).

Coordinates are 0-based half-open (GTF start - 1, end), as in the reference.  ``gene_infos`` / ``transcript_infos``
produced elsewhere (e.g. unpickled from the reference's ``genes.pkl.gz``) work as well: anything with ``.start`` /
``.end`` or that unpacks to ``(start, end)`` is accepted as a segment, and any iterable of those as a collection.
"""
import re
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

_ATTRIBUTE = re.compile(r'(\S+?)\s*"(.*?)"')


class Segment(tuple):
    """(start, end), half open."""
    __slots__ = ()

    def __new__(cls, start: int, end: int):
        return super().__new__(cls, (int(start), int(end)))

    start = property(lambda self: self[0])
    end = property(lambda self: self[1])


def _span(seg) -> Tuple[int, int]:
    if hasattr(seg, 'start') and hasattr(seg, 'end'):
        return int(seg.start), int(seg.end)
    a, b = seg
    return int(a), int(b)


def collapse(segments: Iterable) -> List[Segment]:
    """Sorted segments with overlapping ones merged (touching ones stay apart)."""
    out: List[List[int]] = []
    for a, b in sorted(_span(s) for s in segments):
        if out and a < out[-1][1] and b > out[-1][0]:
            out[-1][1] = max(out[-1][1], b)
        else:
            out.append([a, b])
    return [Segment(a, b) for a, b in out]


def invert(segments: Sequence[Segment], bounds) -> List[Segment]:
    """The parts of `bounds` that the (collapsed) segments leave uncovered: a transcript's introns."""
    lo, hi = _span(bounds)
    out = []
    prev = lo
    for a, b in segments:
        if prev < a:
            out.append(Segment(prev, a))
        prev = max(prev, b)
    if prev < hi:
        out.append(Segment(prev, hi))
    return out


def genes_and_transcripts_from_gtf(gtf_path: str, use_version: bool = False):
    """(gene_infos, transcript_infos) of a GTF: `gene`, `transcript` and `exon` features.

    gene_infos[gene_id] = dict(segment, chromosome, strand, gene_name, transcripts=[ids in file order]);
    transcript_infos[transcript_id] = dict(gene_id, transcript_name, segment, chromosome, strand, exons, introns)."""
    gene_infos: Dict[str, dict] = {}
    transcript_infos: Dict[str, dict] = {}
    exons: Dict[str, list] = {}
    opener = open
    if gtf_path.endswith('.gz'):
        import gzip
        opener = lambda p: gzip.open(p, 'rt')
    with opener(gtf_path) as f:
        for line in f:
            if not line.strip() or line[0] == '#':
                continue
            cols = line.rstrip('\n').split('\t')
            if len(cols) < 9:
                continue
            feature = cols[2]
            if feature not in ('gene', 'transcript', 'exon'):
                continue
            attributes = dict(_ATTRIBUTE.findall(cols[8]))
            gene_id = attributes.get('gene_id')
            if gene_id is None:
                continue
            if use_version and attributes.get('gene_version'):
                gene_id = f'{gene_id}.{attributes["gene_version"]}'
            transcript_id = attributes.get('transcript_id')
            if transcript_id is not None and use_version and attributes.get('transcript_version'):
                transcript_id = f'{transcript_id}.{attributes["transcript_version"]}'
            segment = Segment(int(cols[3]) - 1, int(cols[4]))
            gene = gene_infos.setdefault(gene_id, {'transcripts': []})
            if feature == 'gene':
                gene.update(segment=segment, chromosome=cols[0], strand=cols[6], gene_name=attributes.get('gene_name', ''))
                continue
            if transcript_id is None:
                continue
            if transcript_id not in transcript_infos:
                transcript_infos[transcript_id] = {'gene_id': gene_id, 'chromosome': cols[0], 'strand': cols[6]}
                gene['transcripts'].append(transcript_id)
            if feature == 'transcript':
                transcript_infos[transcript_id].update(segment=segment, transcript_name=attributes.get('transcript_name', ''))
            else:
                exons.setdefault(transcript_id, []).append(segment)
    for transcript_id, info in transcript_infos.items():
        ex = collapse(exons.get(transcript_id, []))
        if 'segment' not in info:
            info['segment'] = Segment(ex[0].start, ex[-1].end) if ex else Segment(0, 0)
        info.setdefault('transcript_name', '')
        info['exons'] = ex
        info['introns'] = invert(ex, info['segment'])
    for gene_id, info in gene_infos.items():
        if 'segment' not in info:
            spans = [transcript_infos[t] for t in info['transcripts']]
            info.update(segment=Segment(min(t['segment'].start for t in spans), max(t['segment'].end for t in spans)),
                        chromosome=spans[0]['chromosome'], strand=spans[0]['strand'], gene_name='')
    return gene_infos, transcript_infos


class AnnotationTables:
    """Flat tables of an annotation for the device (all int32 unless noted):

      gene_start / gene_end / gene_pmax [G]   genes grouped by contig (BAM header order) and sorted by (start, end)
                                              inside a contig -- bam.py:177-178; gene_pmax = running maximum of
                                              gene_end within the contig (bounds the backward scan for containing genes)
      gene_strand u8 [G]                      0 '+', 1 '-', 2 anything else
      contig_off [n_ref + 1]                  genes of contig c are [contig_off[c], contig_off[c + 1])
      gene_tr_off [G + 1]                     transcripts of gene g (file order)
      tr_exon_off / tr_intron_off [T + 1]     collapsed exons / introns of transcript t in `exon_*` / `intron_*`
    gene_index maps gene id -> row; genes whose contig is absent from the BAM header get no row."""

    def __init__(self, gene_infos: dict, transcript_infos: dict, ref_names: Sequence[str]):
        ref_id = {name: i for i, name in enumerate(ref_names)}
        rows = []
        for position, (gene_id, info) in enumerate(gene_infos.items()):
            c = ref_id.get(str(info['chromosome']))
            if c is None or ('segment' not in info and 'start' not in info):
                continue
            a, b = _span(info['segment']) if 'segment' in info else (int(info['start']), int(info['end']))
            rows.append((c, a, b, position, gene_id))
        rows.sort(key=lambda r: r[:4])       # stable in dictionary order, as sorted(..., key=segment) is
        G = len(rows)
        self.gene_ids = [r[4] for r in rows]
        self.gene_index = {g: i for i, g in enumerate(self.gene_ids)}
        self.gene_start = np.array([r[1] for r in rows], dtype=np.int32)
        self.gene_end = np.array([r[2] for r in rows], dtype=np.int32)
        contig = np.array([r[0] for r in rows], dtype=np.int64)
        self.contig_off = np.searchsorted(contig, np.arange(len(ref_names) + 1)).astype(np.int32)
        self.gene_pmax = self.gene_end.copy()
        for c in range(len(ref_names)):
            lo, hi = self.contig_off[c], self.contig_off[c + 1]
            if hi > lo:
                self.gene_pmax[lo:hi] = np.maximum.accumulate(self.gene_end[lo:hi])
        strand_code = {'+': 0, '-': 1}
        self.gene_strand = np.array([strand_code.get(gene_infos[g]['strand'], 2) for g in self.gene_ids], dtype=np.uint8)
        tr_off = [0]
        ex_off, in_off = [0], [0]
        ex_s, ex_e, in_s, in_e = [], [], [], []
        for g in self.gene_ids:
            for t in gene_infos[g].get('transcripts', []):
                info = transcript_infos[t]
                for a, b in collapse(info['exons']):
                    ex_s.append(a); ex_e.append(b)
                for a, b in collapse(info['introns']):
                    in_s.append(a); in_e.append(b)
                ex_off.append(len(ex_s)); in_off.append(len(in_s))
            tr_off.append(len(ex_off) - 1)
        i32 = lambda v: np.asarray(v, dtype=np.int32)
        self.gene_tr_off = i32(tr_off)
        self.tr_exon_off, self.tr_intron_off = i32(ex_off), i32(in_off)
        self.exon_start, self.exon_end = i32(ex_s), i32(ex_e)
        self.intron_start, self.intron_end = i32(in_s), i32(in_e)
        self.n_genes = G
        self._device = None

    ARRAYS = ('gene_start', 'gene_end', 'gene_pmax', 'gene_strand', 'contig_off', 'gene_tr_off', 'tr_exon_off',
              'tr_intron_off', 'exon_start', 'exon_end', 'intron_start', 'intron_end')

    def to_device(self, device):
        import torch
        if self._device is None or self._device[0] != device:
            self._device = (device, {k: torch.from_numpy(np.ascontiguousarray(getattr(self, k))).to(device) for k in self.ARRAYS})
        return self._device[1]
