"""Synthetic spatial-transcriptomics data with the semantics of ``STMiner/Simulate``."""
from .patterns import CONFIGS, make_templates, simulate_replicas, simulate_spot_matrix, simulate_adata, simulate_long_genes, simulate_long_genes_device  # noqa: F401
