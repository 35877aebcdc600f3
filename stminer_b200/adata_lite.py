"""Minimal AnnData duck type for the hot path.

``anndata`` / ``scanpy`` / ``h5py`` are not installable in this image, and the hot
path only touches ``adata.X``, ``adata.obs['x'|'y']``, ``adata.var.index`` / ``var_names``,
``adata.obsm['spatial']`` and ``adata.shape``.  ``AnnDataLite`` provides exactly
that surface; every function here also accepts a real ``anndata.AnnData`` when one is available.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import pandas as pd
from scipy import sparse

from .staging import SpotMatrix


class AnnDataLite:
    def __init__(self, X, obs: Optional[pd.DataFrame] = None, var: Optional[pd.DataFrame] = None,
                 obsm: Optional[dict] = None, uns: Optional[dict] = None):
        self.X = X if sparse.issparse(X) else np.asarray(X)
        n_obs, n_vars = self.X.shape
        self.obs = obs if obs is not None else pd.DataFrame(index=[str(i) for i in range(n_obs)])
        self.var = var if var is not None else pd.DataFrame(index=[str(i) for i in range(n_vars)])
        if len(self.obs) != n_obs or len(self.var) != n_vars:
            raise ValueError("obs/var length does not match X shape %s" % (self.X.shape,))
        self.obsm = obsm if obsm is not None else {}
        self.uns = uns if uns is not None else {}

    # --- the AnnData surface the hot path touches
    @property
    def shape(self):
        return self.X.shape

    @property
    def n_obs(self):
        return self.X.shape[0]

    @property
    def n_vars(self):
        return self.X.shape[1]

    @property
    def var_names(self):
        return self.var.index

    @var_names.setter
    def var_names(self, names):
        self.var.index = pd.Index(list(names))

    @property
    def obs_names(self):
        return self.obs.index

    def __len__(self):
        return self.n_obs

    def copy(self):
        return AnnDataLite(self.X.copy(), self.obs.copy(), self.var.copy(),
                           {k: np.array(v) for k, v in self.obsm.items()}, dict(self.uns))

    def _subset(self, obs_mask=None, var_mask=None):
        """In-place subsetting (what scanpy's filter_genes / filter_cells do to the reference's adata)."""
        X = self.X
        if obs_mask is not None:
            obs_mask = np.asarray(obs_mask)
            X = X[obs_mask]
            self.obs = self.obs.loc[obs_mask].copy() if obs_mask.dtype == bool else self.obs.iloc[obs_mask].copy()
            self.obsm = {k: np.asarray(v)[obs_mask] for k, v in self.obsm.items()}
        if var_mask is not None:
            var_mask = np.asarray(var_mask)
            X = X[:, var_mask]
            self.var = self.var.loc[var_mask].copy() if var_mask.dtype == bool else self.var.iloc[var_mask].copy()
        self.X = X

    def __getitem__(self, key):
        """``adata[:, gene]`` / ``adata[:, [genes]]`` / ``adata[:, bool_mask]`` views (copies here)."""
        if not (isinstance(key, tuple) and len(key) == 2):
            raise IndexError("AnnDataLite supports adata[obs_sel, var_sel] indexing only")
        rsel, csel = key
        if isinstance(csel, str):
            csel = [csel]
        if isinstance(csel, (list, tuple, pd.Index)) and len(csel) and isinstance(csel[0], str):
            names = np.asarray(self.var_names)
            cidx = np.concatenate([np.flatnonzero(names == c) for c in csel])
        elif isinstance(csel, slice):
            cidx = np.arange(self.n_vars)[csel]
        else:
            c = np.asarray(csel)
            cidx = np.flatnonzero(c) if c.dtype == bool else c
        X = self.X[:, cidx]
        out = AnnDataLite(X, self.obs.copy(), self.var.iloc[cidx].copy(), dict(self.obsm), dict(self.uns))
        if not (isinstance(rsel, slice) and rsel == slice(None)):
            r = np.asarray(rsel)
            out._subset(obs_mask=r)
        return out


def column_indices(adata, gene_name_list: Sequence[str]) -> np.ndarray:
    """First column of every requested gene; an unknown gene raises the reference's
    ``ValueError('Gene id: ...')`` (AlgUtils.py:27-28 returns None -> distribution.py:200-202)."""
    names = np.asarray(adata.var_names)
    first = {}
    for i, n in enumerate(names):
        if n not in first:
            first[n] = i
    cols = []
    for g in gene_name_list:
        if g not in first:
            raise ValueError("Gene id: " + str(g) + "\nError: " + str(g) + " is not in adata gene list.")
        cols.append(first[g])
    return np.asarray(cols, dtype=np.int64)


def spot_matrix_from_adata(adata, gene_name_list: Optional[Sequence[str]] = None,
                           truncate_int16: bool = True) -> SpotMatrix:
    """One-pass staging of ``adata`` (real AnnData or AnnDataLite) into the hot path's CSC layout."""
    genes = list(adata.var_names)
    cols = None
    if gene_name_list is not None:
        cols = column_indices(adata, gene_name_list)
    x = np.asarray(adata.obs["x"])
    y = np.asarray(adata.obs["y"])
    return SpotMatrix.from_matrix(adata.X, x, y, genes, columns=cols, truncate_int16=truncate_int16)
