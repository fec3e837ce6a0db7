"""Operator layer mirroring dynast/preprocessing/__init__.py:2-31 for the count -> estimate hot path."""
from .aggregation import (aggregate_counts, calculate_mutation_rates, merge_aggregates, read_aggregates, read_rates)
from .bam import (check_bam_contains_duplicate, check_bam_contains_secondary, check_bam_contains_unmapped, check_bam_is_paired,
                  check_bam_tags_exist, get_tags_from_bam, parse_all_reads, read_alignments, read_conversions, select_alignments)
from .consensus import call_consensus, call_consensus_groups
from .conversion import (CONVERSION_COMPLEMENT, complement_counts, count_conversions, deduplicate_counts,
                         drop_multimappers, read_counts, split_counts_by_velocity, subset_counts)
from .coverage import calculate_coverage, read_coverage
from .snp import detect_snps, extract_conversions, read_snp_csv, read_snps
