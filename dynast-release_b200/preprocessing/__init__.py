"""Operator layer mirroring dynast/preprocessing/__init__.py:2-31 for the count -> estimate hot path."""
from .aggregation import (aggregate_counts, calculate_mutation_rates, merge_aggregates, read_aggregates, read_rates)
from .conversion import (CONVERSION_COMPLEMENT, complement_counts, count_conversions, deduplicate_counts,
                         drop_multimappers, read_counts, split_counts_by_velocity, subset_counts)
