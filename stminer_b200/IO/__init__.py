"""Host-side IO helpers the hot path's callers use (mirror of ``STMiner/IO``; file formats need anndata)."""
from .IOUtil import (STMinerIOError, bin_spatial_adata, merge_bin_coordinate, require_positive_int,  # noqa: F401
                     validate_spatial_array)
