"""`dynast count` -- thin orchestrator with the stage order, options and output files of dynast/count.py:15-393
(reference), calling the B200 operators of dynast_b200.preprocessing.

The GTF is parsed by dynast_b200.gtf (ngs_tools is not needed); the BAM must already be coordinate sorted (the
reference sorts it with samtools when it is not).  The AnnData object is written only when `anndata` is
importable; the layers are returned either way."""
import datetime as dt
import gzip
import json
import os
import pickle
import sys
import time
from typing import Dict, FrozenSet, List, Optional, Tuple

import pandas as pd

from . import constants, gtf, preprocessing, utils


def _write_pickle(obj, path):
    with gzip.open(path, 'wb') as f:
        pickle.dump(obj, f)


def _all_exists(*paths):
    return all(os.path.exists(p) for p in paths)


def count(
    bam_path: str,
    gtf_path,
    out_dir: str,
    strand: str = 'forward',
    umi_tag: Optional[str] = None,
    barcode_tag: Optional[str] = None,
    gene_tag: str = 'GX',
    barcodes: Optional[List[str]] = None,
    control: bool = False,
    quality: int = 27,
    conversions: FrozenSet[FrozenSet[str]] = frozenset({frozenset({'TC'})}),
    snp_threshold: Optional[float] = None,
    snp_min_coverage: int = 1,
    snp_csv: Optional[str] = None,
    n_threads: int = 8,
    temp_dir: Optional[str] = None,
    velocity: bool = True,
    strict_exon_overlap: bool = False,
    dedup_mode: str = 'auto',
    by_name: bool = False,
    nasc: bool = False,
    overwrite: bool = False,
    velocity_assignments: Optional[Dict[Tuple[str, int], Tuple[str, str]]] = None,
    consensus_bam: Optional[bool] = None,
    transcript_infos: Optional[dict] = None,
):
    """count.count (count.py:15-393). Returns the results object (AnnData or utils.Results; None for --control).

    `gtf_path` is the annotation GTF, parsed by gtf.genes_and_transcripts_from_gtf as count.py:146-147 does with
    ngs_tools; a ready-made `gene_infos` dictionary is accepted in its place (then `transcript_infos` carries the
    exon / intron tables velocity assignment needs).  `consensus_bam` overrides the RN-tag detection of
    count.py:295-304 (None: look at the BAM)."""
    t0 = time.time()
    start = dt.datetime.now()
    os.makedirs(out_dir, exist_ok=True)
    stats_path = os.path.join(out_dir, f'{constants.STATS_PREFIX}_{dt.datetime.strftime(start, "%Y%m%d_%H%M%S_%f")}.json')
    all_conversions = sorted(utils.flatten_iter(conversions))

    # ---- BAM tags and alignment checks (count.py:91-135)
    from .preprocessing import bam as bam_ops
    tags = bam_ops.check_bam_tags(bam_path, umi_tag=umi_tag, barcode_tag=barcode_tag, gene_tag=gene_tag, n_threads=n_threads)

    # ---- parse BAM (count.py:138-171): skipped when the outputs exist and not --overwrite
    conversions_path = os.path.join(out_dir, constants.CONVERSIONS_FILENAME)
    index_path = os.path.join(out_dir, constants.CONVERSIONS_INDEX_FILENAME)
    alignments_path = os.path.join(out_dir, constants.ALIGNMENTS_FILENAME)
    genes_path = os.path.join(out_dir, constants.GENES_FILENAME)
    bam_parsed = False
    gene_infos = gtf_path if isinstance(gtf_path, dict) else None
    if not _all_exists(conversions_path, index_path, alignments_path, genes_path) or overwrite:
        if gene_infos is None:
            gene_infos, transcript_infos = gtf.genes_and_transcripts_from_gtf(gtf_path, use_version=False)
        # plain tuples for the segments: the pickle must load without this package (and without ngs_tools)
        _write_pickle({g: {**info, 'segment': tuple(info['segment'])} if 'segment' in info else info
                       for g, info in gene_infos.items()}, genes_path)
        preprocessing.parse_all_reads(
            bam_path, conversions_path, alignments_path, index_path, gene_infos, transcript_infos, strand=strand, umi_tag=umi_tag,
            barcode_tag=barcode_tag, gene_tag=gene_tag, barcodes=barcodes, n_threads=n_threads, temp_dir=temp_dir, nasc=nasc,
            control=control, velocity=velocity, strict_exon_overlap=strict_exon_overlap,
            velocity_assignments=velocity_assignments)
        bam_parsed = True
    elif gene_infos is None:
        with gzip.open(genes_path, 'rb') as f:      # count.py:170-171
            gene_infos = pickle.load(f)

    # ---- SNP detection (count.py:199-282)
    convs_path = os.path.join(out_dir, constants.CONVS_FILENAME)
    redo_snp = True
    if _all_exists(convs_path):
        with gzip.open(convs_path, 'rb') as f:
            redo_snp = pickle.load(f) != conversions
    coverage_path = os.path.join(out_dir, constants.COVERAGE_FILENAME)
    snps_path = os.path.join(out_dir, constants.SNPS_FILENAME)
    if snp_threshold is not None:
        if not _all_exists(convs_path, coverage_path, snps_path) or redo_snp or bam_parsed:
            alignments = preprocessing.select_alignments(preprocessing.read_alignments(alignments_path))
            snp_conversions = set(all_conversions + [preprocessing.CONVERSION_COMPLEMENT[c] for c in all_conversions])
            dfc = preprocessing.read_conversions(conversions_path)
            dfc = dfc[[key in alignments for key in zip(dfc['read_id'].astype(str), dfc['index'].astype(int))]]
            dfc = dfc.loc[dfc['conversion'].isin(snp_conversions), ['contig', 'genome_i']]
            coverage_path = preprocessing.calculate_coverage(
                bam_path, {str(contig): set(part['genome_i']) for contig, part in dfc.groupby('contig', sort=False, observed=True)},
                coverage_path, alignments=alignments, umi_tag=umi_tag, barcode_tag=barcode_tag, gene_tag=gene_tag,
                barcodes=barcodes, temp_dir=temp_dir, velocity=velocity)
            coverage = preprocessing.read_coverage(coverage_path)
            snps_path = preprocessing.detect_snps(
                conversions_path, index_path, coverage, snps_path, alignments=alignments, conversions=snp_conversions,
                quality=quality, threshold=snp_threshold, min_coverage=snp_min_coverage, n_threads=n_threads)
            _write_pickle(conversions, convs_path)
    else:
        _write_pickle(conversions, convs_path)

    # ---- count conversions (count.py:284-321)
    counts_path = os.path.join(out_dir, f'{constants.COUNTS_PREFIX}_{"_".join(all_conversions)}.csv')
    snps = {}
    for part in ([preprocessing.read_snps(snps_path)] if snp_threshold is not None else []) + (
            [preprocessing.read_snp_csv(snp_csv)] if snp_csv else []):
        for conv, by_contig in part.items():
            for contig, pos in by_contig.items():
                snps.setdefault(conv, {}).setdefault(contig, set()).update(pos)
    if umi_tag and dedup_mode == 'auto':      # count.py:295-304: a BAM of consensus reads (RN tag) prioritises exonic reads
        is_consensus = (constants.BAM_CONSENSUS_READ_COUNT_TAG in tags) if consensus_bam is None else consensus_bam
        dedup_mode = 'exon' if is_consensus else 'conversion'
    counts_path = preprocessing.count_conversions(
        conversions_path, alignments_path, index_path, counts_path, gene_infos, barcodes=barcodes, snps=snps, quality=quality,
        conversions=all_conversions, dedup_use_conversions=dedup_mode == 'conversion', n_threads=n_threads, temp_dir=temp_dir)
    df_uncomplemented = preprocessing.read_counts(counts_path)
    df_complemented = preprocessing.complement_counts(df_uncomplemented, gene_infos)

    # ---- mutation rates (count.py:333-358)
    df_counts = df_uncomplemented if nasc else df_complemented
    rates_paths = {'all': preprocessing.calculate_mutation_rates(
        df_counts, os.path.join(out_dir, f'{constants.RATES_PREFIX}.csv'), group_by=['barcode'])}
    if df_complemented['transcriptome'].any():
        rates_paths['X'] = preprocessing.calculate_mutation_rates(
            df_counts[df_counts['transcriptome']], os.path.join(out_dir, f'{constants.RATES_PREFIX}_X.csv'), group_by=['barcode'])
    for key in df_complemented['velocity'].unique():
        if key in constants.VELOCITY_BLACKLIST:
            continue
        rates_paths[key] = preprocessing.calculate_mutation_rates(
            df_counts[df_counts['velocity'] == key], os.path.join(out_dir, f'{constants.RATES_PREFIX}_{key}.csv'),
            group_by=['barcode'])

    # ---- AnnData (count.py:360-392)
    result = None
    if not control:
        if by_name:
            df_complemented['GX'] = df_complemented['GX'].apply(lambda gx: gene_infos[gx].get('gene_name') or gx)
        result = utils.results_to_adata(df_complemented, conversions, gene_infos=gene_infos if not by_name else None)
        obsm = {}
        for key, path in rates_paths.items():
            rates = preprocessing.read_rates(path)
            name = 'rates' if key == 'all' else f'rates_{key}'
            if isinstance(rates, pd.Series):
                obsm[name] = pd.DataFrame([rates] * len(result.obs), index=result.obs.index)
            else:
                obsm[name] = rates.set_index('barcode').reindex(result.obs.index, fill_value=0.0)
        if hasattr(result, 'write'):
            for name, value in obsm.items():
                result.obsm[name] = value
            result.write(os.path.join(out_dir, constants.ADATA_FILENAME), compression='gzip')
        else:
            result.obsm = obsm
    with open(stats_path, 'w') as f:      # stats.py:45-109
        json.dump({'start_time': start.isoformat(), 'end_time': dt.datetime.now().isoformat(), 'elapsed': time.time() - t0,
                   'call': ' '.join(sys.argv), 'version': 'dynast_b200'}, f, indent=4)
    return result
