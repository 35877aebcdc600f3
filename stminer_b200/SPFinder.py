"""``SPFinder`` — drop-in façade of the reference's ``STMiner/SPFinder.py`` for the hot path.

Method names, arguments, defaults, attributes and error behaviour follow the reference; the bodies of the
three hot methods are re-plumbed onto the CUDA library:

* ``fit_pattern``                  -> one batched EM launch        (reference: SPFinder.py:337-398)
* ``build_distance_array('gmm')``  -> one batched pair launch      (reference: SPFinder.py:424-439)
* ``spatial_high_variable_genes``  -> one batched spot-EMD launch  (reference: SPFinder.py:251-326)

Everything else here is host glue (preprocessing restatements of the scanpy calls, DataFrame assembly).
Plotting, GUI, BMK/GEM readers, image patterns are outside the hot path ("out of scope").
"""
from __future__ import annotations

import logging
from collections import Counter
from collections.abc import Mapping
from typing import Dict, List, Optional

import numpy as np
import pandas as pd
from scipy import sparse
from scipy.sparse import csr_matrix
from scipy.stats import zscore

from .adata_lite import AnnDataLite, spot_matrix_from_adata
from .Algorithm.algorithm import cluster
from .Algorithm.distance import (build_gmm_distance_array, build_ot_distance_array, compare_gmm_distance,
                                 ot_distances_to_source)
from .Algorithm.distribution import fit_gmms
from .IO.IOUtil import merge_bin_coordinate

logger = logging.getLogger(__name__)


def scale_array(exp_matrix: np.ndarray, total_count: np.ndarray) -> np.ndarray:
    """SPFinder.py:32-38."""
    total_sum = np.sum(exp_matrix)
    if total_sum == 0:
        return np.zeros_like(total_count)
    return exp_matrix * (100 / total_sum)


class GeneImageDict(Mapping):
    """``csr_dict`` of the reference — gene -> sparse (x, y) image (SPFinder.py:178-249) — built lazily.

    The staging keeps ONE CSC matrix of the selected genes (values already clipped at the ``vmax`` percentile) and
    the spot coordinates; ``d[gene]`` materialises the same ``csr_matrix((column, (x, y)))`` the reference stores
    (cached), while the SVG ranking reads every gene's support straight from the CSC columns (``supports``)
    instead of creating and re-parsing thousands of scipy objects."""

    def __init__(self, genes, columns_csc, x, y):
        self._genes = list(genes)
        self._index = {g: i for i, g in enumerate(self._genes)}
        self._M = columns_csc            # (n_obs, n_genes) CSC, float64, clipped
        self._x = np.asarray(x).astype(np.int64)
        self._y = np.asarray(y).astype(np.int64)
        self._cache = {}

    def __len__(self):
        return len(self._genes)

    def __iter__(self):
        return iter(self._genes)

    def __contains__(self, gene):
        return gene in self._index

    def __getitem__(self, gene):
        if gene not in self._cache:
            j = self._index[gene]
            data = np.asarray(self._M[:, j].todense()).ravel()
            self._cache[gene] = csr_matrix((data, (self._x, self._y)))      # SPFinder.py:244 (explicit zeros kept)
        return self._cache[gene]

    def supports(self, genes=None):
        """``[(x, y, mass)]`` of the positive cells of every gene image, as ``csr_to_support(d[gene])`` returns them
        (duplicate (x, y) rows summed, row-major order)."""
        genes = self._genes if genes is None else list(genes)
        ymul = int(self._y.max()) + 1
        cell = self._x * ymul + self._y
        uniq, inv = np.unique(cell, return_inverse=True)
        merge = len(uniq) != len(cell)
        out = []
        M = self._M
        for g in genes:
            j = self._index[g]
            a, b = M.indptr[j], M.indptr[j + 1]
            r, v = M.indices[a:b], M.data[a:b]
            if merge:                                         # rows sharing a cell are summed by the csr constructor
                c = inv[r]
                s = np.bincount(c, weights=v, minlength=len(uniq))
                keep = np.flatnonzero(s > 0)
                cc, vv = uniq[keep], s[keep]
            else:
                keep = v > 0
                cc, vv = cell[r[keep]], v[keep]
                o = np.argsort(cc, kind="stable")
                cc, vv = cc[o], vv[o]
            out.append(((cc // ymul).astype(np.int32), (cc % ymul).astype(np.int32), vv.astype(np.float64)))
        return out


def _loess_quadratic(x: np.ndarray, y: np.ndarray, span: float = 0.3) -> np.ndarray:
    """Local quadratic regression (tricube weights on the ``span`` fraction of nearest points), evaluated at every x —
    the estimator of Cleveland's loess(degree=2, family="gaussian") without the kd-tree surface interpolation."""
    n = len(x)
    if n == 0:
        return np.zeros(0)
    q = int(min(n, max(3, np.ceil(span * n))))
    order = np.argsort(x, kind="stable")
    xs, ys = x[order], y[order]
    fit = np.empty(n)
    lo = 0
    for i in range(n):                      # the q nearest neighbours of a sorted array form a sliding window
        while lo + q < n and xs[lo + q] - xs[i] < xs[i] - xs[lo]:
            lo += 1
        xw, yw = xs[lo:lo + q], ys[lo:lo + q]
        h = max(xs[i] - xw[0], xw[-1] - xs[i])
        if h <= 0:
            fit[i] = yw.mean()
            continue
        w = np.clip(1.0 - (np.abs(xw - xs[i]) / h) ** 3, 0.0, None) ** 3
        dx = xw - xs[i]
        A = np.stack([np.ones_like(dx), dx, dx * dx], axis=1) * np.sqrt(w)[:, None]
        coef, *_ = np.linalg.lstsq(A, yw * np.sqrt(w), rcond=None)
        fit[i] = coef[0]
    out = np.empty(n)
    out[order] = fit
    return out


def _n_nonzero(X, axis):
    if sparse.issparse(X):
        return np.asarray((X > 0).sum(axis=axis)).ravel()
    return np.asarray((np.asarray(X) > 0).sum(axis=axis)).ravel()


class SPFinder:
    def __init__(self, adata=None):
        self.adata = None
        self.patterns = None
        self.genes_distance_array = None
        self.filtered_distance_array = None
        self.genes_labels = None
        self.kmeans_fit_result = None
        self.mds_features = None
        self.image_gmm = None
        self.global_distance = None
        self.all_labels = None
        self.custom_pattern = None
        self.csr_dict = {}
        self.patterns_matrix_dict = {}
        self.patterns_binary_matrix_dict = {}
        self._highly_variable_genes = []
        self._gene_expression_edge = {}
        self._scope = None
        self.device = None           # CUDA device of the hot path (None = current device)
        self.seed = 0                # seed of the on-device mixture initialisation
        if adata is not None:
            self.set_adata(adata)

    # ------------------------------------------------------------------ data
    def set_adata(self, adata) -> None:
        """SPFinder.py:77-99."""
        self.adata = adata
        has_obs_coordinates = "x" in self.adata.obs and "y" in self.adata.obs
        if has_obs_coordinates:
            if "spatial" not in self.adata.obsm:
                self.adata.obsm["spatial"] = self.adata.obs[["x", "y"]].to_numpy()
        elif "spatial" in self.adata.obsm:
            position = np.asarray(self.adata.obsm["spatial"])
            if len(position.shape) != 2 or position.shape[1] < 2:
                raise ValueError('adata.obsm["spatial"] must be a 2D array with at least two columns.')
            self.adata.obs["x"] = merge_bin_coordinate(position[:, 0], position[:, 0].min(), bin_size=1)
            self.adata.obs["y"] = merge_bin_coordinate(position[:, 1], position[:, 1].min(), bin_size=1)
        else:
            raise ValueError('AnnData must contain either obs["x"] and obs["y"] or obsm["spatial"].')
        self._scope = (0, max(self.adata.obs["y"].max(), self.adata.obs["x"].max()))

    def read_h5ad(self, file: str, amplification: int = 1, bin_size: int = 1, merge_bin: bool = False) -> None:
        """SPFinder.py:101-122 -> IO/read_h5ad.py:16-111."""
        from .IO.read_h5ad import read_h5ad
        self.set_adata(read_h5ad(file, amplification=amplification, bin_size=bin_size, merge_bin=merge_bin))

    def merge_bin(self, bin_width: int) -> None:
        """SPFinder.py:128-151: re-label obs.x / obs.y in place, min-shifted."""
        self.adata.obs["x"] = merge_bin_coordinate(self.adata.obs["x"], self.adata.obs["x"].min(), bin_size=bin_width)
        self.adata.obs["y"] = merge_bin_coordinate(self.adata.obs["y"], self.adata.obs["y"].min(), bin_size=bin_width)
        self._scope = (0, max(self.adata.obs["y"].max(), self.adata.obs["x"].max()))

    # ------------------------------------------------------------------ preprocessing (scanpy restatements)
    def preprocess(self, normalize: bool, exclude_highly_expressed: bool, log1p: bool, min_cells: int,
                   min_genes: int, n_top_genes: int = 2000) -> None:
        """SPFinder.py:400-422.  In-place, like the scanpy calls of the reference:
        ``filter_genes(min_cells)`` -> ``filter_cells(min_genes)`` -> [``highly_variable_genes``] ->
        [``normalize_total``] -> [``log1p``].  A real AnnData with scanpy installed goes through scanpy."""
        a = self.adata
        if not isinstance(a, AnnDataLite):
            import scanpy as sc
            sc.pp.filter_genes(a, min_cells=min_cells)
            sc.pp.filter_cells(a, min_genes=min_genes)
            sc.pp.highly_variable_genes(a, flavor="seurat_v3", n_top_genes=n_top_genes)
            self._highly_variable_genes = list(a.var[a.var["highly_variable"]].index)
            if normalize:
                sc.pp.normalize_total(a, exclude_highly_expressed=exclude_highly_expressed)
            if log1p:
                sc.pp.log1p(a)
            return
        n_cells = _n_nonzero(a.X, 0)
        a.var["n_cells"] = n_cells
        a._subset(var_mask=n_cells >= min_cells)                      # sc.pp.filter_genes
        n_genes = _n_nonzero(a.X, 1)
        a.obs["n_genes"] = n_genes
        a._subset(obs_mask=n_genes >= min_genes)                      # sc.pp.filter_cells
        if n_top_genes is not None and n_top_genes > 0 and n_top_genes < a.n_vars:
            self._highly_variable_genes = self._seurat_v3_hvg(n_top_genes)
        else:
            self._highly_variable_genes = list(a.var.index)
        if normalize:                                                 # sc.pp.normalize_total(target_sum=None)
            X = a.X.tocsr().astype(np.float64) if sparse.issparse(a.X) else np.asarray(a.X, dtype=np.float64)
            counts = np.asarray(X.sum(axis=1)).ravel()
            if exclude_highly_expressed:
                # scanpy 1.10 normalize_total(exclude_highly_expressed=True, max_fraction=0.05): a gene that holds more
                # than 5 % of the counts of at least one cell does not enter the size factors
                if sparse.issparse(X):
                    big = X.data > np.repeat(counts, np.diff(X.indptr)) * 0.05
                    hits = np.bincount(X.indices[big], minlength=X.shape[1])
                else:
                    hits = (X > counts[:, None] * 0.05).sum(axis=0)
                keep = np.asarray(hits).ravel() == 0
                counts = np.asarray(X[:, keep].sum(axis=1)).ravel()
            target = np.median(counts[counts > 0])
            scale = np.where(counts > 0, target / np.where(counts > 0, counts, 1.0), 1.0)
            a.X = sparse.diags(scale) @ X if sparse.issparse(X) else X * scale[:, None]
        if log1p:                                                     # sc.pp.log1p
            if sparse.issparse(a.X):
                a.X = a.X.copy(); a.X.data = np.log1p(a.X.data)
            else:
                a.X = np.log1p(a.X)

    def _seurat_v3_hvg(self, n_top_genes: int) -> List[str]:
        """Restatement of scanpy's ``highly_variable_genes(flavor="seurat_v3")`` (SPFinder.py:409-413; scanpy 1.10
        ``_highly_variable_genes_seurat_v3``): log10(variance) regressed on log10(mean) by a local quadratic fit (span 0.3),
        counts clipped at ``mean + sqrt(N) * regularised std``, genes ranked by the variance of the standardised counts.
        scanpy takes the fit from scikit-misc's loess (Cleveland's Fortran code with a kd-tree-interpolated surface, not
        in this image); here the local regression is evaluated exactly at every gene — the same estimator without the
        interpolation error, so the ranking can differ in near-ties only.  Consumed only when ``n_top_genes > 0``."""
        X = self.adata.X
        N = X.shape[0]
        if sparse.issparse(X):
            Xc = X.tocsc().astype(np.float64)
            mean = np.asarray(Xc.mean(axis=0)).ravel()
            sq = np.asarray(Xc.multiply(Xc).sum(axis=0)).ravel()
        else:
            Xc = np.asarray(X, dtype=np.float64)
            mean = Xc.mean(axis=0)
            sq = (Xc * Xc).sum(axis=0)
        var = (sq - N * mean ** 2) / max(N - 1, 1)                     # ddof = 1, as scanpy's _get_mean_var
        not_const = var > 0
        est = np.zeros_like(var)
        est[not_const] = _loess_quadratic(np.log10(mean[not_const]), np.log10(var[not_const]), span=0.3)
        reg_std = np.sqrt(10.0 ** est)
        clip = reg_std * np.sqrt(N) + mean
        if sparse.issparse(Xc):                                        # sums of the clipped counts, column by column
            d = Xc.data.copy()
            col = np.repeat(np.arange(Xc.shape[1]), np.diff(Xc.indptr))
            np.minimum(d, clip[col], out=d)
            csum = np.bincount(col, weights=d, minlength=Xc.shape[1])
            csq = np.bincount(col, weights=d * d, minlength=Xc.shape[1])
        else:
            d = np.minimum(Xc, clip[None, :])
            csum, csq = d.sum(axis=0), (d * d).sum(axis=0)
        with np.errstate(divide="ignore", invalid="ignore"):
            norm_var = (1.0 / ((N - 1) * reg_std ** 2)) * (N * mean ** 2 + csq - 2.0 * csum * mean)
        norm_var[~np.isfinite(norm_var)] = 0.0
        order = np.argsort(-norm_var, kind="stable")[:n_top_genes]
        mask = np.zeros(self.adata.n_vars, dtype=bool); mask[order] = True
        self.adata.var["highly_variable"] = mask
        self.adata.var["variances_norm"] = norm_var
        return list(self.adata.var.index[mask])

    # ------------------------------------------------------------------ spot-level EMD path
    def get_genes_csr_array(self, min_cells: int, min_genes: int = 1, normalize: bool = True,
                            exclude_highly_expressed: bool = False, log1p: bool = False, vmax: float = 100.0,
                            gene_list: Optional[List[str]] = None) -> None:
        """SPFinder.py:178-249: preprocess in place, then one sparse (x, y) image per gene with the values
        clipped at the ``vmax`` percentile (taken over ALL spots, zeros included, and only if it is > 0)."""
        self.csr_dict = {}
        self.preprocess(normalize, exclude_highly_expressed, log1p, min_cells, min_genes)
        arr_list = gene_list if gene_list is not None else list(self.adata.var.index)
        names = np.asarray(self.adata.var_names)
        counts = Counter(names.tolist())
        first = {}
        for i, n in enumerate(names):
            first.setdefault(n, i)
        x = np.asarray(self.adata.obs["x"].values).ravel()
        y = np.asarray(self.adata.obs["y"].values).ravel()
        Xc = (self.adata.X.tocsc() if sparse.issparse(self.adata.X) else sparse.csc_matrix(np.asarray(self.adata.X)))
        Xc = Xc.astype(np.float64)
        warned, seen, keep_genes, keep_cols = set(), set(), [], []
        for gene in arr_list:
            if gene not in counts:                 # ``self.adata[:, gene]`` raises KeyError (SPFinder.py:224)
                raise KeyError("Values [%r], from [%r], are not valid obs/ var names or indices." % (gene, gene))
            if counts[gene] != 1:                  # SPFinder.py:225-234: duplicated var names are skipped with a warning
                if gene not in warned:
                    warned.add(gene)
                    logger.warning("Gene [%s] has more than one index, skip.", gene)
                continue
            if gene not in seen:
                seen.add(gene); keep_genes.append(gene); keep_cols.append(first[gene])
        M = Xc[:, keep_cols].tocsc(copy=True)
        M.sort_indices()
        if vmax < 100:                                # clip at the percentile over ALL spots, zeros included, if it is > 0
            n_obs = M.shape[0]
            for j in range(M.shape[1]):
                a, b = M.indptr[j], M.indptr[j + 1]
                col = np.zeros(n_obs); col[:b - a] = M.data[a:b]
                p = np.percentile(col, vmax)
                if p > 0:
                    np.minimum(M.data[a:b], p, out=M.data[a:b])
        # vmax = 100: the percentile is the column maximum and the clip is the identity
        self.csr_dict = GeneImageDict(keep_genes, M, x, y)

    def spatial_high_variable_genes(self, vmax: float = 100.0, thread: int = 1) -> None:
        """SPFinder.py:251-326: exact EMD between the per-spot total image and every gene image, then
        ``DataFrame[Gene, Distance, z-score]`` sorted by Distance descending.  ``thread`` is accepted for
        compatibility; all genes go into one batched device launch."""
        if len(self.csr_dict) == 0:
            self.get_genes_csr_array(min_cells=50, vmax=vmax, normalize=True)
        data = np.array(self.adata.X.sum(axis=1)).flatten().astype(np.float64)
        data[data > np.percentile(data, vmax)] = np.percentile(data, vmax)
        row_indices = np.array(self.adata.obs["x"].values).flatten()
        column_indices = np.array(self.adata.obs["y"].values).flatten()
        global_matrix = csr_matrix((data, (row_indices, column_indices)))
        keys = list(self.csr_dict.keys())
        if isinstance(self.csr_dict, GeneImageDict):          # supports straight from the staged CSC columns
            from .Algorithm.distance import csr_to_support
            from .sharding import emd_sharded
            dist, status = emd_sharded(csr_to_support(global_matrix), self.csr_dict.supports(keys), device=self.device)
        else:
            dist, status = ot_distances_to_source(global_matrix, [self.csr_dict[k] for k in keys], device=self.device)
        distance_dict = {}
        for k, d, s in zip(keys, dist, status):
            # status 0 = optimal, 1 = stopped at numItermax with the current basis cost (what POT returns too).
            # Anything else (2 = empty image, >= 3 = the solver could not run the problem) is what an exception is
            # in the reference's loop: logged, gene skipped (SPFinder.py:294-299) — never a NaN in the table.
            if s >= 2 or not np.isfinite(d):
                logger.error("Failed to compute OT distance for gene %s (solver status %d).", k, int(s))
                continue
            distance_dict[k] = d
        self.global_distance = pd.DataFrame(list(distance_dict.items()), columns=["Gene", "Distance"]) \
            .sort_values(by="Distance", ascending=False)
        self.global_distance["z-score"] = zscore(np.log1p(self.global_distance["Distance"]))

    # ------------------------------------------------------------------ mixture fit
    def fit_pattern(self, n_top_genes: int = -1, n_comp: int = 20, normalize: bool = False,
                    exclude_highly_expressed: bool = False, log1p: bool = False, min_cells: int = 20,
                    gene_list: Optional[List[str]] = None, remove_low_exp_spots: bool = False) -> None:
        """SPFinder.py:337-398."""
        self.preprocess(normalize=normalize, exclude_highly_expressed=exclude_highly_expressed, log1p=log1p,
                        min_cells=min_cells, min_genes=1, n_top_genes=n_top_genes)
        if gene_list is not None and isinstance(gene_list, list) and len(gene_list) > 0:
            gene_to_fit = gene_list
        elif n_top_genes > 0:
            gene_to_fit = self._highly_variable_genes
        else:
            gene_to_fit = list(self.adata.var.index)
        try:
            self.patterns = fit_gmms(self.adata, gene_to_fit, n_comp=n_comp,
                                     remove_low_exp_spots=remove_low_exp_spots, seed=self.seed, device=self.device)
        except Exception:
            logger.error("Failed to fit patterns.", exc_info=True)

    def compare_gene_to_genes(self, gene_name: str) -> pd.DataFrame:
        """SPFinder.py:165-176."""
        return compare_gmm_distance(self.patterns[gene_name], self.patterns, device=self.device)

    # ------------------------------------------------------------------ distances / clustering
    def build_distance_array(self, method: str = "gmm", gene_list: Optional[List[str]] = None) -> None:
        """SPFinder.py:424-439 -> services/spfinder_analysis.py:18-35."""
        if gene_list is None:
            gene_list = list(self.adata.var.index)
        if method == "gmm":
            self.genes_distance_array = build_gmm_distance_array(self.patterns, device=self.device)
        elif method == "ot":
            self.genes_distance_array = build_ot_distance_array(self.csr_dict, gene_list, device=self.device)
        elif method in ("mse", "cs"):
            raise NotImplementedError("method '%s' is a dense-grid O(G^2 XY) path outside the B200 hot path "
                                      "(SURVEY.md §2 row 2)" % method)
        else:
            raise ValueError("Unknown method, method should be one of gmm, mse, cs, ot")

    def cluster_gene(self, n_clusters: int, mds_components: int = 20, use_highly_variable_gene: bool = False,
                     n_top_genes: int = 500) -> None:
        """SPFinder.py:526-539 -> services/spfinder_analysis.py:38-62."""
        if use_highly_variable_gene:
            df = pd.DataFrame(self.genes_distance_array.mean(axis=1), columns=["mean"])
            df = df.sort_values(by="mean", ascending=False)
            hv_gene_list = list(df[:n_top_genes].index)
            self.filtered_distance_array = self.genes_distance_array.loc[hv_gene_list, hv_gene_list]
            target = self.filtered_distance_array
        else:
            target = self.genes_distance_array
        self.genes_labels, self.kmeans_fit_result, self.mds_features = cluster(
            target, n_clusters=n_clusters, mds_components=mds_components, device=self.device)

    # ------------------------------------------------------------------ pattern voting
    def get_pattern_array(self, vote_rate: float = 0.0, mode: str = "vote") -> None:
        """SPFinder.py:441-457.  All labels are accumulated in one device pass over the CSR matrix when it qualifies
        (``_patterns_on_device``); otherwise label by label on the host."""
        self.patterns_binary_matrix_dict = {}
        if mode == "vote":
            labels = list(set(self.genes_labels["labels"]))
            lists = {label: list(self.genes_labels[self.genes_labels["labels"] == label]["gene_id"]) for label in labels}
            done = self._patterns_on_device(lists, vote_rate)
            for label in labels:
                binary_arr, total_count = done[label] if done is not None else self._genes_to_pattern(lists[label], vote_rate)
                self.patterns_matrix_dict[label] = total_count
                self.patterns_binary_matrix_dict[label] = binary_arr
        elif mode == "test":
            pass                                     # unimplemented in the reference as well (SPFinder.py:452-455)
        else:
            raise ValueError("mode should be vote or test")

    def _patterns_on_device(self, gene_lists: Dict, vote_rate: float):
        """{label: (binary_arr, total_count)} of ``_genes_to_pattern`` for every label of ``gene_lists`` in ONE pass of
        the vote kernel over the spot-by-gene CSR matrix (``stm_pattern_vote``); ``None`` when the matrix does not qualify
        (not canonical CSR float32 / small integers, observations sharing a spot, a gene in two lists, > 64 labels, no
        CUDA device) — the caller then runs the host loop."""
        from .adata_lite import column_indices
        from .staging import device_staging_supported, pattern_vote_on_device
        try:
            import torch
            if not torch.cuda.is_available():
                return None
        except ImportError:      # pragma: no cover
            return None
        labels = list(gene_lists)
        x = np.asarray(self.adata.obs["x"]); y = np.asarray(self.adata.obs["y"])
        if not labels or len(labels) > 64 or not device_staging_supported(self.adata.X, x, y, None):
            return None
        n_var = self.adata.X.shape[1]
        col_label = np.full(n_var, -1, dtype=np.int32)
        for li, label in enumerate(labels):
            cols = column_indices(self.adata, gene_lists[label])
            if len(set(cols.tolist())) != len(cols) or (col_label[cols] >= 0).any():
                return None
            col_label[cols] = li
        total, votes = pattern_vote_on_device(self.adata.X, n_var, col_label, len(labels), device=self.device)
        shape = (int(x.max()) + 1, int(y.max()) + 1)
        xi, yi = x.astype(np.int64), y.astype(np.int64)
        out = {}
        for li, label in enumerate(labels):
            total_count = np.zeros(shape, dtype=np.float64); vote_cnt = np.zeros(shape, dtype=np.float64)
            total_count[xi, yi] = total[li]; vote_cnt[xi, yi] = votes[li]
            vote_array = (vote_cnt > 0) & (vote_cnt / len(gene_lists[label]) >= vote_rate)
            total_count *= vote_array
            out[label] = (np.where(total_count != 0, 1, total_count), total_count)
        return out

    def _genes_to_pattern(self, gene_list: List[str], vote_rate: float):
        """SPFinder.py:459-489 on the CSC staging: per gene the int16 grid scaled to sum 100, accumulated;
        cells where at least ``vote_rate`` of the genes are expressed are kept."""
        sm = spot_matrix_from_adata(self.adata, gene_list)
        shape = (int(np.asarray(self.adata.obs["x"]).max()) + 1, int(np.asarray(self.adata.obs["y"]).max()) + 1)
        # all genes at once on the CSC staging (one bincount per quantity instead of a Python loop over genes)
        w = np.trunc(np.asarray(sm.val, dtype=np.float64))
        # AlgUtils.py:8-36 keeps negative cells in the grid; nonzero cells vote (np.nonzero, SPFinder.py:471)
        nz = w != 0
        gene_of = np.repeat(np.arange(sm.n_genes), np.diff(sm.col_ptr))
        sums = np.bincount(gene_of, weights=w, minlength=sm.n_genes)               # scale_array: 100 / sum, 0 if sum == 0
        scale = np.divide(100.0, sums, out=np.zeros_like(sums), where=sums != 0)
        cell = sm.spot_x[sm.row_idx].astype(np.int64) * shape[1] + sm.spot_y[sm.row_idx]
        n_cells = shape[0] * shape[1]
        total_count = np.bincount(cell[nz], weights=(w * scale[gene_of])[nz], minlength=n_cells).reshape(shape)
        votes = np.bincount(cell[nz], minlength=n_cells).reshape(shape).astype(np.float64)
        vote_array = (votes > 0) & (votes / len(gene_list) >= vote_rate)
        total_count *= vote_array
        binary_arr = np.where(total_count != 0, 1, total_count)
        return binary_arr, total_count

    def get_custom_pattern(self, gene_list: List[str], n_components: int = 20, vote_rate: float = 0.0,
                           mode: str = "vote") -> None:
        """SPFinder.py:491-524: mixture of the voted pattern image, fitted on the device."""
        if mode == "vote":
            done = self._patterns_on_device({"custom": list(gene_list)}, vote_rate)
            _, total_count = done["custom"] if done is not None else self._genes_to_pattern(gene_list, vote_rate)
            # SPFinder.py:516-517: GaussianMixture(n_components).fit -> sklearn's default max_iter = 100
            self.custom_pattern = _fit_image(np.round(total_count).astype(np.int32), n_components, self.seed, self.device,
                                             max_iter=100)
        elif mode == "test":
            pass
        else:
            raise ValueError("mode should be vote or test")

    def get_pattern_of_given_genes(self, gene_list: List[str], n_comp: int = 20) -> None:
        """SPFinder.py:572-598: pattern of a user-provided gene list; genes absent from ``adata.var.index`` are silently
        ignored, the result is stored by ``get_custom_pattern`` (``self.custom_pattern``)."""
        if self.adata is None:
            raise ValueError("Please load ST data first.")
        var_index = set(self.adata.var.index)
        _genes = [gene for gene in gene_list if gene in var_index]
        self.get_custom_pattern(gene_list=_genes, n_components=n_comp, vote_rate=0)

    def get_all_labels(self) -> None:
        """SPFinder.py:556-570: nearest voted pattern (by GMM distance) of every fitted gene."""
        df_list = []
        # the reference hard-codes 20 components here, which only works when the genes were fitted with 20
        # (distance.py:29 takes K from the first mixture); use the fitted genes' K so other K work too
        n_comp = next(iter(self.patterns.values())).weights_.size
        for i in self.patterns_binary_matrix_dict:
            gmm = _fit_image(np.asarray(self.patterns_binary_matrix_dict[i], dtype=np.int32), n_comp, self.seed,
                             self.device)
            df = compare_gmm_distance(gmm, self.patterns, device=self.device)
            df.rename(columns={0: i}, inplace=True)
            df_list.append(df)
        total = pd.concat(df_list, axis=1)
        label = total.idxmin(axis=1)
        all_labels = pd.DataFrame(label, columns=["label"])
        all_labels["value"] = [total.loc[g, lab] for g, lab in zip(all_labels.index, all_labels["label"])]
        self.all_labels = all_labels


def _fit_image(image: np.ndarray, n_components: int, seed: int, device, max_iter: int = 200):
    """``get_gmm`` (distribution.py:37-41, max_iter=200 there) for a dense count image, on the device."""
    from .Algorithm.distribution import fit_spot_matrix, result_to_gmm_dict
    from .staging import DeviceSpotMatrix, SpotMatrix
    xs, ys = np.nonzero(image > 0)
    sm = SpotMatrix(np.array([0, len(xs)], dtype=np.int64), np.arange(len(xs), dtype=np.int32),
                    image[xs, ys].astype(np.float32), xs.astype(np.int32), ys.astype(np.int32), ["pattern"])
    res = fit_spot_matrix(DeviceSpotMatrix(sm, device=device), n_components, init=None, max_iter=max_iter, seed=seed)
    d = result_to_gmm_dict(sm.genes, res.to_host())
    if "pattern" not in d:
        raise ValueError("pattern image has fewer distinct spots than components")
    return d["pattern"]
