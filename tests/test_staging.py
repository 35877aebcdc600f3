"""CPU: one-pass CSC staging keeps the reference's per-gene semantics (AlgUtils.py:26-36)."""
import numpy as np
import pandas as pd
from scipy import sparse

from oracle import gmm_oracle as go
from stminer_b200.adata_lite import AnnDataLite, column_indices, spot_matrix_from_adata
from stminer_b200.staging import SpotMatrix
import pytest


def _random_adata(seed=0, n_obs=300, n_var=12, dup=True, dense=False):
    rng = np.random.default_rng(seed)
    x = rng.integers(0, 9, n_obs); y = rng.integers(0, 11, n_obs)       # many duplicate (x, y) rows
    if not dup:
        idx = rng.permutation(9 * 11)[:min(n_obs, 99)]
        x, y = idx // 11, idx % 11
        n_obs = len(idx)
    X = rng.gamma(0.6, 2.0, size=(n_obs, n_var)) * (rng.random((n_obs, n_var)) < 0.4)
    X[:, 3] *= -1                                                         # a negative column
    Xs = X if dense else sparse.csr_matrix(X)
    obs = pd.DataFrame({"x": x, "y": y})
    var = pd.DataFrame(index=["g%d" % i for i in range(n_var)])
    return AnnDataLite(Xs, obs=obs, var=var), X, x, y


@pytest.mark.parametrize("dup,dense", [(True, False), (False, False), (True, True)])
def test_spot_matrix_matches_reference_per_gene(dup, dense):
    adata, X, x, y = _random_adata(dup=dup, dense=dense)
    genes = ["g5", "g0", "g3", "g7"]
    sm = spot_matrix_from_adata(adata, genes)
    assert sm.genes == genes
    for gi, g in enumerate(genes):
        col = int(g[1:])
        xy_ref, c_ref = go.gene_column_to_counts(x, y, X[:, col])
        xy, c = sm.gene_spots(gi)
        assert xy.tolist() == xy_ref.tolist()
        assert c.tolist() == c_ref.tolist()


def test_unknown_gene_raises_like_reference():
    adata, *_ = _random_adata()
    with pytest.raises(ValueError, match="Gene id: nope"):
        column_indices(adata, ["g1", "nope"])


def test_select_and_lpt_order():
    adata, *_ = _random_adata()
    sm = spot_matrix_from_adata(adata)
    order = sm.lpt_order()
    assert (np.diff(sm.nnz_per_gene()[order]) <= 0).all()
    sub = sm.select([4, 1])
    a = sm.gene_spots(4); b = sub.gene_spots(0)
    assert a[0].tolist() == b[0].tolist() and a[1].tolist() == b[1].tolist()


def test_compact_wire_format_round_trip():
    """uint16 spot index + int16 value (4 B / value) carries the same per-gene spots; matrices that do not
    qualify (float values, > 65,536 spots) are left alone."""
    adata, *_ = _random_adata()
    sm = spot_matrix_from_adata(adata)
    c = sm.compact()
    assert c.is_compact and c.row_idx.dtype == np.uint16 and c.val.dtype == np.int16
    for g in range(sm.n_genes):
        a, b = sm.gene_spots(g), c.gene_spots(g)
        assert a[0].tolist() == b[0].tolist() and a[1].tolist() == b[1].tolist()
    part = c.slice_genes(2, 7)
    assert part.is_compact and part.n_genes == 5 and part.gene_spots(0)[1].tolist() == sm.gene_spots(2)[1].tolist()
    floats = spot_matrix_from_adata(adata, truncate_int16=False)
    assert floats.compact() is floats
    cuts = sm.nnz_balanced_chunks(4)
    assert cuts[0] == 0 and cuts[-1] == sm.n_genes and cuts == sorted(cuts)


@pytest.mark.parametrize("binary,cut,remove_low,thr", [(True, False, False, 95), (False, True, False, 97.5), (True, True, True, 90),
                                                       (False, True, True, 99), (True, False, False, 150)])
def test_preprocess_array_binary_cut_matches_reference_dense_path(binary, cut, remove_low, thr):
    """SpotMatrix.preprocess_array (host transform of the staged columns) against the oracle's dense-grid restatement of
    get_exp_array + preprocess_array (AlgUtils.py:8-36, distribution.py:133-145)."""
    from oracle import gmm_oracle as go
    from stminer_b200.Simulate import simulate_spot_matrix
    sm = simulate_spot_matrix("T", seed=2, n_templates=3, replicas=3, truncate_int16=False)
    out = sm.preprocess_array(binary=binary, cut=cut and not binary, threshold=thr, remove_low_exp_spots=remove_low)
    for g in range(sm.n_genes):
        a, b = int(sm.col_ptr[g]), int(sm.col_ptr[g + 1])
        r = sm.row_idx[a:b]
        xy_ref, c_ref = go.gene_column_to_counts(sm.spot_x[r], sm.spot_y[r], sm.val[a:b], remove_low_exp_spots=remove_low,
                                                 binary=binary, cut=cut, threshold=thr)
        xy, c = out.gene_spots(g)
        assert np.array_equal(xy, xy_ref) and np.array_equal(c, c_ref), g
