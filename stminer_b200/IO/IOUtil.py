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
    """Raised by the IO helpers; subclass of ValueError like the reference's (IOUtil.py:10)."""


_FIELDS = ("operation", "path", "expected", "actual", "hint")


def format_io_error(message, **context) -> str:
    """``message | operation=... | path=... | expected=... | actual=... | hint=...`` — empty fields are skipped and
    the field order is fixed, which is the text format of the reference's IO errors (IOUtil.py:14-34)."""
    return " | ".join([message] + ["%s=%s" % (f, context[f]) for f in _FIELDS if context.get(f)])


def _bad_parameter(name, path, expected, actual):
    return STMinerIOError(format_io_error("Invalid parameter '%s'" % name, operation="validate parameter", path=path,
                                          expected=expected, actual=actual))


def require_positive_int(value, name: str, *, path=None) -> int:
    """Positive-integer parameter check with the reference's policy (IOUtil.py:37-96): bools are rejected, floats
    are truncated with a UserWarning (non-finite ones rejected), anything else non-integral is rejected, and the
    result must be > 0."""
    kind = type(value).__name__
    if isinstance(value, bool):
        raise _bad_parameter(name, path, "positive integer", "%s (%s)" % (value, kind))
    if isinstance(value, (float, np.floating)):
        if not np.isfinite(value):
            raise _bad_parameter(name, path, "finite positive integer", str(value))
        truncated = int(value)
        warnings.warn(format_io_error("Float parameter '%s' converted to int" % name, operation="validate parameter",
                                      path=path, expected="integer input", actual="%s -> %s" % (value, truncated),
                                      hint="Pass int explicitly to avoid implicit conversion"),
                      UserWarning, stacklevel=2)
        value = truncated
    elif not isinstance(value, (int, np.integer)):
        raise _bad_parameter(name, path, "> 0", "%s (%s)" % (value, kind))
    if int(value) <= 0:
        raise _bad_parameter(name, path, "> 0", str(int(value)))
    return int(value)


def validate_spatial_array(spatial, *, path=None, source="adata.obsm['spatial']") -> np.ndarray:
    """Spatial coordinates must be a numeric (n, >=2) array with finite x / y (IOUtil.py:99-133)."""
    arr = np.asarray(spatial)
    problem = None
    if arr.ndim != 2 or arr.shape[1] < 2:
        problem = ("Invalid spatial coordinates", "2D array with at least 2 columns", "shape=%s" % (arr.shape,),
                   "Ensure %s exists and stores x/y coordinates" % source)
    elif not np.issubdtype(arr.dtype, np.number):
        problem = ("Spatial coordinates must be numeric", "numeric array", str(arr.dtype),
                   "Convert %s to numeric values before loading" % source)
    elif not np.isfinite(arr[:, :2].astype(np.float64)).all():
        problem = ("Spatial coordinates contain NaN/Inf", "finite numeric x/y values", "contains NaN or Inf", None)
    if problem is not None:
        raise STMinerIOError(format_io_error(problem[0], operation="validate spatial", path=path, expected=problem[1],
                                             actual=problem[2], hint=problem[3]))
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
