"""Coordinate binning and validated IO errors — semantics of ``STMiner/IO/IOUtil.py``.

Kept because the drop-in API names them (``read_h5ad(bin_size, merge_bin)``, ``SPFinder.merge_bin``):
  * ``merge_bin_coordinate``  floor((c - c_min) / bin)                      (IOUtil.py:136-138)
  * ``bin_spatial_adata``     rows grouped by ``coords // bin`` (no min shift), X summed or averaged per
                              bin into a dense matrix, obs.x/obs.y = bin indices, names ``bin_{gx}_{gy}``
                              (IOUtil.py:146-224) — done here with one sparse group-sum instead of a Python
                              loop over bins
  * parameter validation      bool rejected, float coerced with a warning, <= 0 rejected (IOUtil.py:37-96)
  * errors                    ``STMinerIOError(ValueError)`` with ``message | key=value | ...`` text (IOUtil.py:10-34)
"""
from __future__ import annotations

import warnings

import numpy as np
import pandas as pd
from scipy import sparse


class STMinerIOError(ValueError):
    """IO layer error with actionable context for users."""


def format_io_error(message, *, path=None, operation=None, expected=None, actual=None, hint=None) -> str:
    parts = [message]
    for key, val in (("operation", operation), ("path", path), ("expected", expected), ("actual", actual),
                     ("hint", hint)):
        if val:
            parts.append("%s=%s" % (key, val))
    return " | ".join(parts)


def require_positive_int(value, name: str, *, path=None) -> int:
    def fail(expected, actual):
        raise STMinerIOError(format_io_error("Invalid parameter '%s'" % name, operation="validate parameter",
                                             path=path, expected=expected, actual=actual))
    if isinstance(value, bool):
        fail("positive integer", "%s (%s)" % (value, type(value).__name__))
    if isinstance(value, (float, np.floating)):
        if not np.isfinite(value):
            fail("finite positive integer", str(value))
        as_int = int(value)
        warnings.warn(format_io_error("Float parameter '%s' converted to int" % name, operation="validate parameter",
                                      path=path, expected="integer input", actual="%s -> %s" % (value, as_int),
                                      hint="Pass int explicitly to avoid implicit conversion"),
                      UserWarning, stacklevel=2)
        value = as_int
    elif not isinstance(value, (int, np.integer)):
        fail("> 0", "%s (%s)" % (value, type(value).__name__))
    value = int(value)
    if value <= 0:
        fail("> 0", str(value))
    return value


def validate_spatial_array(spatial, *, path=None, source="adata.obsm['spatial']") -> np.ndarray:
    arr = np.asarray(spatial)
    if arr.ndim != 2 or arr.shape[1] < 2:
        raise STMinerIOError(format_io_error("Invalid spatial coordinates", operation="validate spatial", path=path,
                                             expected="2D array with at least 2 columns", actual="shape=%s" % (arr.shape,),
                                             hint="Ensure %s exists and stores x/y coordinates" % source))
    if not np.issubdtype(arr.dtype, np.number):
        raise STMinerIOError(format_io_error("Spatial coordinates must be numeric", operation="validate spatial",
                                             path=path, expected="numeric array", actual=str(arr.dtype),
                                             hint="Convert %s to numeric values before loading" % source))
    xy = arr[:, :2].astype(np.float64)
    if not np.isfinite(xy).all():
        raise STMinerIOError(format_io_error("Spatial coordinates contain NaN/Inf", operation="validate spatial",
                                             path=path, expected="finite numeric x/y values",
                                             actual="contains NaN or Inf"))
    return arr


def merge_bin_coordinate(coordinate, coordinate_min, bin_size):
    bin_size = require_positive_int(bin_size, "bin_size")
    return np.floor((np.asarray(coordinate) - coordinate_min) / bin_size).astype(int)


def get_bin_center(bin_coordinate, coordinate_min, bin_size):
    bin_size = require_positive_int(bin_size, "bin_size")
    return bin_coordinate * bin_size + coordinate_min + int(bin_size / 2)


def bin_spatial_adata(adata, bin_size=50, method="sum"):
    from ..adata_lite import AnnDataLite
    bin_size = require_positive_int(bin_size, "bin_size")
    if method not in {"sum", "mean"}:
        raise STMinerIOError(format_io_error("Invalid aggregation method", operation="bin_spatial_adata",
                                             expected="'sum' or 'mean'", actual=repr(method)))
    if "spatial" not in adata.obsm:
        raise STMinerIOError(format_io_error("Missing spatial coordinates", operation="bin_spatial_adata",
                                             expected="adata.obsm['spatial']",
                                             actual="obsm_keys=%s" % list(adata.obsm.keys()),
                                             hint="Load data with spatial coordinates before merge_bin=True"))
    coords = validate_spatial_array(adata.obsm["spatial"], source="adata.obsm['spatial']")
    if adata.n_obs == 0:
        raise STMinerIOError(format_io_error("Cannot bin empty AnnData", operation="bin_spatial_adata",
                                             expected="adata.n_obs > 0", actual=str(adata.n_obs)))
    gx = (coords[:, 0] // bin_size).astype(int)
    gy = (coords[:, 1] // bin_size).astype(int)
    # bins in (gx, gy) lexicographic order — the order pandas' groupby iterates them in the reference
    keys = pd.MultiIndex.from_arrays([gx, gy])
    codes, uniques = pd.factorize(keys, sort=True)
    n_bins = len(uniques)
    P = sparse.csr_matrix((np.ones(len(codes)), (codes, np.arange(len(codes)))), shape=(n_bins, len(codes)))
    X = adata.X
    agg = P @ (X.tocsr() if sparse.issparse(X) else np.asarray(X))
    agg = np.asarray(agg.todense()) if sparse.issparse(agg) else np.asarray(agg)
    if method == "mean":
        agg = agg / np.asarray(P.sum(axis=1))
    bx = np.array([u[0] for u in uniques], dtype=int)
    by = np.array([u[1] for u in uniques], dtype=int)
    obs = pd.DataFrame({"x": bx, "y": by}, index=["bin_%d_%d" % (a, b) for a, b in zip(bx, by)])
    new_coords = np.column_stack([(bx + 0.5) * bin_size, (by + 0.5) * bin_size]).astype(int)
    if isinstance(adata, AnnDataLite):
        return AnnDataLite(agg, obs=obs, var=adata.var.copy(), obsm={"spatial": new_coords}, uns=dict(adata.uns))
    import anndata as ad
    out = ad.AnnData(X=agg, var=adata.var.copy(), uns=adata.uns.copy())
    out.obsm["spatial"] = new_coords
    out.obs_names = list(obs.index)
    out.obs["x"] = bx
    out.obs["y"] = by
    return out
