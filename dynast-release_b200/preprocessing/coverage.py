"""calculate_coverage -- same name, arguments and coverage.csv as dynast/preprocessing/coverage.py (reference)."""
import re
from typing import Dict, List, Optional, Set, Tuple

import numpy as np
import torch

from .. import _lib, bamio
from .._lib import check, ptr

COVERAGE_PARSER = re.compile(r'^(?P<contig>[^,]*),(?P<genome_i>[^,]*),(?P<coverage>[^,]*)\n$')


def read_coverage(coverage_path: str) -> Dict[str, Dict[int, int]]:
    """coverage.read_coverage (coverage.py:21-41)."""
    coverage = {}
    with open(coverage_path, 'r') as f:
        for i, line in enumerate(f):
            groups = COVERAGE_PARSER.match(line).groupdict()
            try:
                contig, genome_i, count = groups['contig'], int(groups['genome_i']), int(groups['coverage'])
                coverage.setdefault(contig, {})[genome_i] = count
            except ValueError as e:
                if i == 0:
                    continue
                raise e
    return coverage


def coverage_of_units(blob, rec_off, poi: np.ndarray, selected: Optional[np.ndarray] = None):
    """dyn_coverage on host arrays: returns (cov uint32 [n_poi], first uint64 [n_poi])."""
    import ctypes as C
    ctx = _lib.current_context()
    dev = _lib.current_torch_device()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    n_poi = len(poi)
    cov = torch.zeros(max(n_poi, 1), dtype=torch.int32, device=dev)
    first = torch.zeros(max(n_poi, 1), dtype=torch.int64, device=dev)
    blob_d = t(np.asarray(blob).view(np.uint8)) if len(blob) else torch.zeros(16, dtype=torch.uint8, device=dev)
    off_d = t(np.asarray(rec_off).view(np.int32))
    sel_d = None if selected is None else t(selected.astype(np.uint8))
    check(ctx.lib.dyn_coverage(ctx.handle, ptr(blob_d), ptr(off_d), len(rec_off) - 1, ptr(sel_d), ptr(t(poi.view(np.int64))),
                               n_poi, ptr(cov), ptr(first), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return cov[:n_poi].cpu().numpy().view(np.uint32), first[:n_poi].cpu().numpy().view(np.uint64)


def calculate_coverage(
    bam_path: str,
    conversions: Dict[str, Set[int]],
    coverage_path: str,
    alignments: Optional[Set[Tuple[str, int]]] = None,
    umi_tag: Optional[str] = None,
    barcode_tag: Optional[str] = None,
    gene_tag: str = 'GX',
    barcodes: Optional[List[str]] = None,
    temp_dir: Optional[str] = None,
    velocity: bool = True,
) -> str:
    """coverage.calculate_coverage (coverage.py:163-243): number of selected alignments covering each position
    where a conversion was observed; rows in first-touch order per contig, contigs in BAM header order."""
    with bamio.PackedBam(bam_path, barcode_tag=barcode_tag, umi_tag=umi_tag, gene_tag=gene_tag, require_gene=not velocity) as pb:
        n = pb.n_units
        ref_names = list(pb.ref_name)
        selected = np.ones(n, dtype=bool)
        if barcodes:
            selected &= np.isin(pb.barcode, list(barcodes))
        if alignments:
            rid, hi = pb.read_id, pb.hi
            selected &= np.fromiter(((rid[i], int(hi[i])) in alignments for i in range(n)), dtype=bool, count=n)
        ref_id = {name: i for i, name in enumerate(ref_names)}
        keys = [np.asarray([(ref_id[c] << 32) | int(p) for p in pos], dtype=np.uint64)
                for c, pos in conversions.items() if c in ref_id and len(pos)]
        poi = np.unique(np.concatenate(keys)) if keys else np.zeros(0, dtype=np.uint64)
        cov, first = coverage_of_units(pb.blob, pb.rec_off, poi, selected)
    touched = np.flatnonzero(cov > 0)
    order = touched[np.lexsort((first[touched], poi[touched] >> np.uint64(32)))]
    with open(coverage_path, 'w') as out:
        out.write('contig,genome_i,coverage\n')
        for i in order:
            out.write(f'{ref_names[int(poi[i] >> np.uint64(32))]},{int(poi[i] & np.uint64(0xffffffff))},{int(cov[i])}\n')
    return coverage_path
