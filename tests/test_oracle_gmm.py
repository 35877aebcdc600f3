"""CPU: the float64 count-weighted EM oracle against golden vectors made with the real sklearn
GaussianMixture on the replicated list (tests/golden/make_golden.py), and against sklearn live."""
import os

import numpy as np
import pytest

from oracle import gmm_oracle as go

GOLD = os.path.join(os.path.dirname(__file__), "golden", "gmm_fixed_init.npz")


def _cases():
    z = np.load(GOLD)
    for i in range(int(z["n"])):
        p = "g%d_" % i
        yield i, {k[len(p):]: z[k] for k in z.files if k.startswith(p)}


@pytest.mark.parametrize("i,c", list(_cases()))
def test_weighted_em_matches_sklearn_golden(i, c):
    r = go.weighted_em(c["xy"], c["counts"], c["w0"], c["mu0"], c["cov0"])
    assert r["n_iter"] == int(c["n_iter"])
    assert r["converged"] == bool(c["converged"])
    np.testing.assert_allclose(r["weights"], c["weights"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(r["means"], c["means"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(r["covariances"], c["covariances"], rtol=1e-8, atol=1e-9)
    assert abs(r["lower_bound"] - float(c["lower_bound"])) < 1e-10


def test_weighted_em_matches_sklearn_live():
    _, c = next(_cases())
    a = go.weighted_em(c["xy"], c["counts"], c["w0"], c["mu0"], c["cov0"])
    b = go.sklearn_fit_from_init(c["xy"], c["counts"], c["w0"], c["mu0"], c["cov0"])
    assert a["n_iter"] == b["n_iter"]
    np.testing.assert_allclose(a["means"], b["means"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(a["covariances"], b["covariances"], rtol=1e-8, atol=1e-9)


def test_gene_column_to_counts_semantics():
    # AlgUtils.py:35 truncation toward zero, duplicates summed, cells <= 0 dropped, row-major order
    x = np.array([2, 0, 2, 1, 1, 0]); y = np.array([1, 3, 1, 0, 0, 0])
    v = np.array([1.9, 2.2, 3.7, -4.0, 0.9, 0.3])
    xy, c = go.gene_column_to_counts(x, y, v)
    assert xy.tolist() == [[0, 3], [2, 1]]          # (1,0) sums to -4+0 -> dropped; (0,0) truncates to 0
    assert c.tolist() == [2, 4]                      # trunc(1.9)+trunc(3.7) = 1+3
    rep = go.array_to_list(xy, c)
    assert rep.shape == (6, 2) and rep[:2].tolist() == [[0, 3], [0, 3]]


def test_remove_low_exp_spots():
    x = np.arange(6); y = np.zeros(6, dtype=int)
    v = np.array([1, 1, 2, 6, 10, 0])
    xy, c = go.gene_column_to_counts(x, y, v, remove_low_exp_spots=True)
    # mean of nonzero = 4 -> max(v-4,0) = [0,0,0,2,6,0]
    assert xy[:, 0].tolist() == [3, 4] and c.tolist() == [2, 6]
