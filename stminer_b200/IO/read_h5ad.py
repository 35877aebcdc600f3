"""``read_h5ad`` with the reference's semantics (``STMiner/IO/read_h5ad.py:16-111``).

Reading the HDF5 container itself needs ``anndata`` (not installable in the build image); everything after
the read — validation, amplification, ``merge_bin`` binning or coordinate re-labelling — is done here and also
works on an in-memory AnnData / AnnDataLite through ``process_spatial_adata``.
"""
from __future__ import annotations

import os

import numpy as np

from .IOUtil import (STMinerIOError, bin_spatial_adata, format_io_error, merge_bin_coordinate, require_positive_int,
                     validate_spatial_array)


def process_spatial_adata(adata, bin_size=1, merge_bin=False, amplification=1, path=None):
    """The part of ``read_h5ad`` after the file is loaded (read_h5ad.py:63-111)."""
    bin_size = require_positive_int(bin_size, "bin_size", path=path)
    amplification = require_positive_int(amplification, "amplification", path=path)
    if "spatial" not in adata.obsm:
        raise STMinerIOError(format_io_error("Missing required spatial coordinates", operation="read_h5ad", path=path,
                                             expected="adata.obsm['spatial']",
                                             actual="obsm_keys=%s" % list(adata.obsm.keys()),
                                             hint="Provide h5ad with spatial coordinates in adata.obsm['spatial']"))
    position = validate_spatial_array(adata.obsm["spatial"], path=path)
    if merge_bin:
        if int(amplification) != 1:
            adata = adata.copy()
            adata.obsm["spatial"] = position * np.int32(amplification)
        try:
            return bin_spatial_adata(adata, bin_size=bin_size)
        except STMinerIOError:
            raise
        except Exception as e:
            raise STMinerIOError(format_io_error("Failed while binning spatial AnnData", operation="bin_spatial_adata",
                                                 path=path, actual="%s: %s" % (type(e).__name__, e),
                                                 hint="Check that adata.obsm['spatial'] exists and contains valid "
                                                      "coordinates")) from e
    amp = np.int32(amplification)
    adata.obs["x"] = merge_bin_coordinate(position[:, 0] * amp, position[:, 0].min() * amp, bin_size=bin_size)
    adata.obs["y"] = merge_bin_coordinate(position[:, 1] * amp, position[:, 1].min() * amp, bin_size=bin_size)
    adata.misc = {"up": position[:, 1].min(), "down": position[:, 1].max(),
                  "left": position[:, 0].min(), "right": position[:, 0].max()}
    return adata


def read_h5ad(file, bin_size=1, merge_bin=False, amplification=1):
    path = str(file)
    require_positive_int(bin_size, "bin_size", path=path)
    require_positive_int(amplification, "amplification", path=path)
    if not os.path.isfile(path):
        raise FileNotFoundError(format_io_error("H5AD file does not exist", operation="read_h5ad", path=path,
                                                expected="existing .h5ad file",
                                                hint="Check file path and access permission"))
    try:
        import anndata
    except ImportError as e:
        raise STMinerIOError(format_io_error("Cannot read h5ad files", operation="read_h5ad", path=path,
                                             expected="the 'anndata' package",
                                             actual="ImportError: %s" % e,
                                             hint="Install anndata, or pass an in-memory AnnData / AnnDataLite to "
                                                  "SPFinder(adata) / process_spatial_adata")) from e
    try:
        adata = anndata.read_h5ad(path)
    except Exception as e:
        raise STMinerIOError(format_io_error("Failed to read h5ad file", operation="anndata.read_h5ad", path=path,
                                             expected="valid AnnData .h5ad file",
                                             actual="%s: %s" % (type(e).__name__, e),
                                             hint="File may be corrupted or not in AnnData format")) from e
    return process_spatial_adata(adata, bin_size=bin_size, merge_bin=merge_bin, amplification=amplification, path=path)
