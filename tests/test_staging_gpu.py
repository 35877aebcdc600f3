"""GPU: device-side staging (csrc/staging.cu, through the C ABI) against the host staging it replaces
(``SpotMatrix.from_matrix`` = AlgUtils.py:8-36 semantics, tested on its own in tests/test_staging.py)."""
import numpy as np
import pytest
from scipy import sparse

pytestmark = pytest.mark.gpu


def _random_adata(n_x, n_y, n_genes, density, seed, shuffle=True, big_values=False):
    from stminer_b200.adata_lite import AnnDataLite
    import pandas as pd
    rng = np.random.default_rng(seed)
    xs, ys = np.meshgrid(np.arange(n_x), np.arange(n_y), indexing="ij")
    x, y = xs.ravel(), ys.ravel()
    keep = rng.random(len(x)) < 0.9                      # a tissue with holes
    x, y = x[keep], y[keep]
    if shuffle:                                          # observations in arbitrary order, as in a real h5ad
        o = rng.permutation(len(x)); x, y = x[o], y[o]
    n_obs = len(x)
    X = sparse.random(n_obs, n_genes, density=density, format="csr", random_state=seed, dtype=np.float32)
    X.data = (rng.gamma(1.2, 2.0, X.nnz) * (rng.random(X.nnz) < 0.9)).astype(np.float32)      # floats in (0, 1) truncate to 0, exact zeros stored
    if big_values:
        X.data[::17] = 40000.5                           # wraps to int16 (AlgUtils.py:35)
        X.data[::23] = -3.7                              # truncates toward zero, stays negative
    genes = ["g%d" % i for i in range(n_genes)]
    obs = pd.DataFrame({"x": x.astype(int), "y": y.astype(int)}, index=["s%d" % i for i in range(n_obs)])
    var = pd.DataFrame({"gene_ids": genes}, index=genes)
    return AnnDataLite(X, obs=obs, var=var, obsm={})


@pytest.mark.parametrize("n_x,n_y,n_genes,density,big", [(17, 19, 40, 0.3, False), (60, 70, 300, 0.05, True), (300, 250, 64, 0.02, False)])
def test_device_staging_equals_host_staging(cuda_device, n_x, n_y, n_genes, density, big):
    """Same col_ptr / spot indices / values (and spot table) as SpotMatrix.from_matrix(...).compact(); 75,000 spots in
    the last case use the 8 B / value encoding."""
    from stminer_b200.adata_lite import column_indices
    from stminer_b200.staging import SpotMatrix, device_staging_supported, stage_csr_on_device
    a = _random_adata(n_x, n_y, n_genes, density, seed=n_genes, big_values=big)
    rng = np.random.default_rng(1)
    pick = [a.var_names[i] for i in rng.permutation(n_genes)[:max(3, n_genes // 2)]]     # a subset, in another order
    cols = column_indices(a, pick)
    x, y = np.asarray(a.obs["x"]), np.asarray(a.obs["y"])
    assert device_staging_supported(a.X, x, y, cols)
    dsm = stage_csr_on_device(a.X, x, y, list(a.var_names), columns=cols)
    ref = SpotMatrix.from_matrix(a.X, x, y, list(a.var_names), columns=cols).compact()
    assert dsm.host.genes == ref.genes == pick
    np.testing.assert_array_equal(dsm.host.col_ptr, ref.col_ptr)
    np.testing.assert_array_equal(dsm.host.spot_x, ref.spot_x); np.testing.assert_array_equal(dsm.host.spot_y, ref.spot_y)
    nnz = int(ref.col_ptr[-1])
    assert dsm.compact == ref.is_compact
    row = dsm.row_idx[:nnz].cpu().numpy(); val = dsm.val[:nnz].cpu().numpy()
    if dsm.compact:
        row = row.view(np.uint16)
    np.testing.assert_array_equal(row.astype(np.int64), ref.row_idx.astype(np.int64))
    np.testing.assert_array_equal(val.astype(np.float64), ref.val.astype(np.float64))


def test_fit_gmms_same_result_through_device_and_host_staging(cuda_device, monkeypatch):
    """fit_gmms(adata, genes): the device-staged path (default for a CSR AnnData) and the host-staged path give
    bit-identical mixtures — the device init is keyed by the staged data, which are the same arrays."""
    import stminer_b200.staging as staging
    from stminer_b200.Algorithm.distribution import fit_gmms
    a = _random_adata(40, 40, 30, 0.4, seed=3)
    genes = list(a.var_names)[5:25]
    dev = fit_gmms(a, genes, n_comp=6, seed=2)
    monkeypatch.setattr(staging, "device_staging_supported", lambda *args, **kw: False)
    host = fit_gmms(a, genes, n_comp=6, seed=2)
    assert list(dev) == list(host) and len(dev) > 10
    for g in dev:
        np.testing.assert_array_equal(dev[g].weights_, host[g].weights_)
        np.testing.assert_array_equal(dev[g].means_, host[g].means_)
        np.testing.assert_array_equal(dev[g].covariances_, host[g].covariances_)


def test_unsupported_inputs_take_the_host_path(cuda_device):
    from stminer_b200.staging import device_staging_supported
    a = _random_adata(10, 10, 8, 0.5, seed=0)
    x, y = np.asarray(a.obs["x"]), np.asarray(a.obs["y"])
    assert not device_staging_supported(a.X.tocsc(), x, y, None)                  # not CSR
    assert not device_staging_supported(a.X.astype(np.float64), x, y, None)       # values that float32 may not hold
    assert not device_staging_supported(a.X, x, y, [1, 2, 1])                     # a gene asked for twice
    assert not device_staging_supported(a.X, np.zeros_like(x), y * 0, None)       # observations sharing a spot


def test_pattern_vote_on_device_matches_host_loop(cuda_device):
    """get_pattern_array (SPFinder.py:441-489): all labels in one pass of the vote kernel == the host restatement of the
    reference's per-label loop (tests/test_host_logic.py pins that one against a dense re-computation); the AnnData rows
    are shuffled and carry floats that truncate to 0, negatives and int16 wraps."""
    import pandas as pd
    from stminer_b200 import SPFinder
    a = _random_adata(23, 31, 60, 0.25, seed=5, big_values=True)
    sp = SPFinder(a)
    genes = list(a.var_names)
    rng = np.random.default_rng(0)
    lab = rng.integers(0, 4, len(genes))
    sp.genes_labels = pd.DataFrame({"gene_id": genes, "labels": lab})
    for vote_rate in (0.0, 0.3):
        sp.patterns_matrix_dict, sp.patterns_binary_matrix_dict = {}, {}
        sp.get_pattern_array(vote_rate=vote_rate)
        assert set(sp.patterns_matrix_dict) == set(lab.tolist())
        for label in set(lab.tolist()):
            gl = [g for g, l in zip(genes, lab) if l == label]
            binary, total = sp._genes_to_pattern(gl, vote_rate)
            np.testing.assert_allclose(sp.patterns_matrix_dict[label], total, rtol=1e-12, atol=1e-12)
            np.testing.assert_array_equal(sp.patterns_binary_matrix_dict[label], binary)
    done = sp._patterns_on_device({"a": genes[:10], "b": genes[5:15]}, 0.0)
    assert done is None                                     # a gene in two lists: host path
    sp.get_custom_pattern(genes[:20], n_components=4, vote_rate=0.1)
    assert sp.custom_pattern.means_.shape == (4, 2)
