"""CPU: host-side logic of the drop-in façade — preprocessing restatements, binning/IO helpers, pattern voting."""
import warnings

import numpy as np
import pandas as pd
import pytest
from scipy import sparse

from stminer_b200.adata_lite import AnnDataLite
from stminer_b200.IO.IOUtil import (STMinerIOError, bin_spatial_adata, merge_bin_coordinate, require_positive_int,
                                    validate_spatial_array)
from stminer_b200.IO.read_h5ad import process_spatial_adata, read_h5ad
from stminer_b200.SPFinder import SPFinder, scale_array


def _adata(seed=0, n_obs=120, n_var=30, spatial_only=False):
    rng = np.random.default_rng(seed)
    X = rng.poisson(0.7, size=(n_obs, n_var)).astype(np.float64)
    xy = np.column_stack([rng.integers(100, 164, n_obs), rng.integers(50, 90, n_obs)])
    obs = pd.DataFrame(index=["c%d" % i for i in range(n_obs)]) if spatial_only else \
        pd.DataFrame({"x": xy[:, 0] - 100, "y": xy[:, 1] - 50}, index=["c%d" % i for i in range(n_obs)])
    var = pd.DataFrame(index=["g%d" % i for i in range(n_var)])
    return AnnDataLite(sparse.csr_matrix(X), obs=obs, var=var, obsm={"spatial": xy}), X, xy


def test_set_adata_and_merge_bin_semantics():
    """tests/test_STMiner.py:11-21 of the reference: obs.x from obsm['spatial'] min-shifted; merge_bin(2) halves it."""
    adata, X, xy = _adata(spatial_only=True)
    sp = SPFinder()
    assert sp.adata is None
    sp.set_adata(adata)
    assert sp.adata.obs["x"].max() == xy[:, 0].max() - xy[:, 0].min()
    mx = sp.adata.obs["x"].max()
    sp.merge_bin(bin_width=2)
    assert sp.adata.obs["x"].max() == mx // 2
    assert sp._scope == (0, max(sp.adata.obs["x"].max(), sp.adata.obs["y"].max()))
    with pytest.raises(ValueError):
        SPFinder(AnnDataLite(np.zeros((3, 2))))


def test_merge_bin_coordinate_and_validation():
    assert merge_bin_coordinate(np.array([10, 14, 15, 29]), 10, 5).tolist() == [0, 0, 1, 3]
    with pytest.raises(STMinerIOError, match="bin_size"):
        merge_bin_coordinate(np.arange(3), 0, 0)
    with pytest.raises(STMinerIOError):
        require_positive_int(True, "bin_size")
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert require_positive_int(3.7, "bin_size") == 3
        assert len(w) == 1 and "converted to int" in str(w[0].message)
    with pytest.raises(STMinerIOError, match="expected=2D array"):
        validate_spatial_array(np.zeros(5))
    with pytest.raises(STMinerIOError, match="NaN"):
        validate_spatial_array(np.array([[0.0, np.nan]]))
    assert isinstance(STMinerIOError("x"), ValueError)
    with pytest.raises(FileNotFoundError):
        read_h5ad("/nonexistent/file.h5ad")


def test_bin_spatial_adata_groups_and_sums():
    """IOUtil.py:146-224: rows grouped by coords // bin (no min shift), X summed per bin, names bin_{gx}_{gy}."""
    adata, X, xy = _adata(n_obs=200)
    out = bin_spatial_adata(adata, bin_size=5)
    gx, gy = xy[:, 0] // 5, xy[:, 1] // 5
    keys = sorted(set(zip(gx.tolist(), gy.tolist())))
    assert out.shape == (len(keys), X.shape[1])
    assert list(out.obs_names) == ["bin_%d_%d" % k for k in keys]
    for r, (a, b) in enumerate(keys[:25]):
        np.testing.assert_allclose(np.asarray(out.X)[r], X[(gx == a) & (gy == b)].sum(axis=0))
        assert out.obs["x"].iloc[r] == a and out.obs["y"].iloc[r] == b
    mean = bin_spatial_adata(adata, bin_size=5, method="mean")
    np.testing.assert_allclose(np.asarray(mean.X)[0], X[(gx == keys[0][0]) & (gy == keys[0][1])].mean(axis=0))
    with pytest.raises(STMinerIOError, match="aggregation"):
        bin_spatial_adata(adata, bin_size=5, method="max")
    # read_h5ad's post-load path, both modes (read_h5ad.py:77-111)
    a2, _, xy2 = _adata(spatial_only=True)
    merged = process_spatial_adata(a2, bin_size=4, merge_bin=True)
    assert merged.n_obs == len(set(zip((xy2[:, 0] // 4).tolist(), (xy2[:, 1] // 4).tolist())))
    a3, _, xy3 = _adata(spatial_only=True)
    rel = process_spatial_adata(a3, bin_size=4, merge_bin=False, amplification=2)
    assert rel.n_obs == a3.n_obs
    assert rel.obs["x"].tolist() == np.floor((xy3[:, 0] * 2 - xy3[:, 0].min() * 2) / 4).astype(int).tolist()


def test_preprocess_filters_and_normalises_in_place():
    """SPFinder.py:400-422 restated without scanpy: filter_genes(min_cells), filter_cells(min_genes),
    normalize_total (target = median of per-cell totals), log1p — all in place."""
    adata, X, _ = _adata(seed=3)
    sp = SPFinder(adata)
    sp.preprocess(normalize=True, exclude_highly_expressed=False, log1p=True, min_cells=45, min_genes=5)
    keep_g = (X > 0).sum(axis=0) >= 45
    Xg = X[:, keep_g]
    keep_c = (Xg > 0).sum(axis=1) >= 5
    Xc = Xg[keep_c]
    assert sp.adata.shape == Xc.shape
    assert list(sp.adata.var_names) == [n for n, k in zip(["g%d" % i for i in range(30)], keep_g) if k]
    tot = Xc.sum(axis=1)
    ref = np.log1p(Xc / tot[:, None] * np.median(tot))
    np.testing.assert_allclose(np.asarray(sp.adata.X.todense()), ref, rtol=1e-12)
    assert sp._highly_variable_genes == list(sp.adata.var_names)


def test_normalize_total_exclude_highly_expressed():
    """scanpy 1.10 ``normalize_total(exclude_highly_expressed=True)`` (SPFinder.py:417-420): genes holding more than 5 % of
    some cell's counts stay out of the size factors; dense restatement of scanpy's rule as the check."""
    from stminer_b200.Simulate import simulate_adata
    a = simulate_adata("T", seed=4)
    X0 = a.X.toarray().astype(float)
    sp = SPFinder(a)
    sp.preprocess(normalize=True, exclude_highly_expressed=True, log1p=False, min_cells=1, min_genes=1, n_top_genes=-1)
    cpc = X0.sum(1)
    sub = (X0 > cpc[:, None] * 0.05).sum(0) == 0
    assert 0 < sub.sum() < X0.shape[1]
    c2 = X0[:, sub].sum(1)
    c2 = c2 / np.median(c2[c2 > 0])
    c2[c2 == 0] = 1
    np.testing.assert_allclose(sp.adata.X.toarray(), X0 / c2[:, None], rtol=1e-12, atol=1e-12)


def test_seurat_v3_hvg_restatement():
    """highly_variable_genes(flavor="seurat_v3") (SPFinder.py:409-413) restated with an exactly evaluated local quadratic
    fit: Poisson genes get a standardised variance of about 1, over-dispersed genes rank first, and the loess reproduces a
    quadratic exactly."""
    from scipy import sparse
    from stminer_b200.adata_lite import AnnDataLite
    from stminer_b200.SPFinder import _loess_quadratic
    x = np.sort(np.random.default_rng(1).uniform(-2, 3, 300))
    np.testing.assert_allclose(_loess_quadratic(x, 0.5 * x * x - x + 2.0), 0.5 * x * x - x + 2.0, rtol=1e-9, atol=1e-9)
    rng = np.random.default_rng(0)
    N, G = 1500, 400
    mu = np.exp(rng.normal(0, 1.2, G))
    X = rng.poisson(mu[None, :], (N, G)).astype(np.float32)
    hv = rng.choice(G, 30, replace=False)
    X[:, hv] = rng.negative_binomial(1.0, 1.0 / (1.0 + mu[hv][None, :]), (N, 30))
    genes = ["g%d" % i for i in range(G)]
    a = AnnDataLite(sparse.csr_matrix(X), obs=pd.DataFrame({"x": np.arange(N) // 50, "y": np.arange(N) % 50}),
                    var=pd.DataFrame(index=genes), obsm={})
    sp = SPFinder(a)
    top = sp._seurat_v3_hvg(30)
    assert len(set(top) & {genes[i] for i in hv}) >= 26
    nv = sp.adata.var["variances_norm"].values
    assert 0.85 < np.median(np.delete(nv, hv)) < 1.15
    assert int(sp.adata.var["highly_variable"].sum()) == 30
    # the dense path gives the same ranking
    b = AnnDataLite(X.copy(), obs=a.obs.copy(), var=pd.DataFrame(index=genes), obsm={})
    assert SPFinder(b)._seurat_v3_hvg(30) == top


def test_get_genes_csr_array_percentile_clip():
    """SPFinder.py:220-249: values clipped at the vmax percentile over ALL spots (only if that percentile > 0)."""
    adata, X, xy = _adata(seed=4)
    sp = SPFinder(adata)
    sp.get_genes_csr_array(min_cells=1, normalize=False, vmax=90)
    assert set(sp.csr_dict) == set(sp.adata.var_names)
    g = "g7"
    col = X[:, 7].copy()
    p = np.percentile(col, 90)
    if p > 0:
        col[col > p] = p
    ref = sparse.csr_matrix((col, (xy[:, 0] - 100, xy[:, 1] - 50)))
    assert (sp.csr_dict[g] != ref).nnz == 0
    # duplicate gene names are skipped (SPFinder.py:225-234)
    adata2, _, _ = _adata(seed=4)
    adata2.var.index = pd.Index(["dup", "dup"] + ["g%d" % i for i in range(2, 30)])
    sp2 = SPFinder(adata2)
    sp2.get_genes_csr_array(min_cells=1, normalize=False)
    assert "dup" not in sp2.csr_dict and len(sp2.csr_dict) == 28


def test_genes_to_pattern_vote():
    """SPFinder.py:459-489: per gene the int16 grid scaled to sum 100; cells kept where >= vote_rate of the genes
    are expressed."""
    adata, X, xy = _adata(seed=6, n_obs=60, n_var=6)
    sp = SPFinder(adata)
    genes = ["g0", "g2", "g5"]
    binary, total = sp._genes_to_pattern(genes, vote_rate=0.5)
    shape = (int(sp.adata.obs["x"].max()) + 1, int(sp.adata.obs["y"].max()) + 1)
    ref_total = np.zeros(shape); votes = np.zeros(shape)
    for g in genes:
        grid = np.zeros(shape, dtype=np.int16)
        np.add.at(grid, (xy[:, 0] - 100, xy[:, 1] - 50), X[:, int(g[1:])].astype(np.int16))
        ref_total += scale_array(grid.astype(float), ref_total)
        votes += grid != 0
    mask = (votes / len(genes) >= 0.5) & (votes > 0)
    np.testing.assert_allclose(total, ref_total * mask)
    assert np.array_equal(binary, (total != 0).astype(float))
    with pytest.raises(ValueError, match="mode"):
        sp.get_pattern_array(mode="nope")
    with pytest.raises(ValueError, match="Unknown method"):
        sp.build_distance_array(method="nope")


@pytest.mark.parametrize("n_obs,vmax", [(120, 100.0), (4000, 100.0), (4000, 85.0)])
def test_gene_image_dict_supports_equal_the_materialised_images(n_obs, vmax):
    """The SVG ranking reads each gene's support straight from the staged CSC columns; it must equal what the
    reference's per-gene ``csr_matrix((column, (x, y)))`` + ``nonzero`` staging yields (distance.py:192-199) — also when
    several spots share a cell (n_obs = 4000 on a 64 x 40 grid: the csr constructor sums them)."""
    from stminer_b200.Algorithm.distance import csr_to_support
    adata, X, xy = _adata(seed=8, n_obs=n_obs, n_var=12)
    sp = SPFinder(adata)
    sp.get_genes_csr_array(min_cells=1, normalize=False, vmax=vmax)
    keys = list(sp.csr_dict.keys())
    assert keys == list(sp.adata.var_names) and len(sp.csr_dict) == len(keys) and keys[3] in sp.csr_dict
    sup = sp.csr_dict.supports()
    for k, s in zip(keys, sup):
        ref = csr_to_support(sp.csr_dict[k])
        assert all(np.array_equal(a, b) for a, b in zip(ref, s)), k
    part = sp.csr_dict.supports(keys[2:5])
    assert all(np.array_equal(a, b) for a, b in zip(part[0], sup[2]))
