"""`dynast consensus` -- thin orchestrator with the checks, options and output file of dynast/consensus.py:14-123
(reference), calling preprocessing.call_consensus (device consensus + BAM writer)."""
import datetime as dt
import json
import os
import sys
import time
import warnings
from typing import List, Optional

from . import constants, gtf, preprocessing
from .preprocessing import bam as bam_ops


def consensus(
    bam_path: str,
    gtf_path,
    out_dir: str,
    strand: str = 'forward',
    umi_tag: Optional[str] = None,
    barcode_tag: Optional[str] = None,
    gene_tag: str = 'GX',
    barcodes: Optional[List[str]] = None,
    quality: int = 27,
    add_RS_RI: bool = False,
    n_threads: int = 8,
    temp_dir: Optional[str] = None,
):
    """consensus.consensus (consensus.py:14-123): writes <out_dir>/consensus.bam (+ .bai) and the run_info JSON.
    `gtf_path` may also be a ready gene_infos dictionary.  The BAM must be coordinate sorted (the reference sorts it with
    samtools when it is not)."""
    t0 = time.time()
    start = dt.datetime.now()
    os.makedirs(out_dir, exist_ok=True)
    stats_path = os.path.join(out_dir, f'{constants.STATS_PREFIX}_{dt.datetime.strftime(start, "%Y%m%d_%H%M%S_%f")}.json')
    bam_ops.check_bam_tags(bam_path, umi_tag=umi_tag, barcode_tag=barcode_tag, gene_tag=gene_tag, n_threads=n_threads)
    gene_infos = gtf_path if isinstance(gtf_path, dict) else gtf.genes_and_transcripts_from_gtf(gtf_path, use_version=False)[0]
    consensus_path = os.path.join(out_dir, constants.CONSENSUS_BAM_FILENAME)
    preprocessing.call_consensus(bam_path, consensus_path, gene_infos, strand=strand, umi_tag=umi_tag, barcode_tag=barcode_tag,
                                 gene_tag=gene_tag, barcodes=barcodes, quality=quality, add_RS_RI=add_RS_RI, temp_dir=temp_dir,
                                 n_threads=n_threads)
    with open(stats_path, 'w') as f:
        json.dump({'start_time': start.isoformat(), 'end_time': dt.datetime.now().isoformat(), 'elapsed': time.time() - t0,
                   'call': ' '.join(sys.argv), 'version': 'dynast_b200'}, f, indent=4)
    return consensus_path
