"""The numpy SMACOF restatement against the real scikit-learn of this image (same Guttman update; rule 1 is the
installed version's stopping rule)."""
import numpy as np
import pytest

from oracle import mds_oracle as mo


def _dist(n, d, seed):
    rng = np.random.default_rng(seed)
    P = rng.normal(size=(n, d + 2))
    D = np.sqrt(((P[:, None] - P[None]) ** 2).sum(-1))
    return D


@pytest.mark.parametrize("n,d,eps", [(60, 3, 1e-6), (150, 20, 1e-5), (40, 2, 1e-3)])
def test_rule1_matches_installed_sklearn(n, d, eps):
    from sklearn.manifold import _mds
    D = _dist(n, d, seed=n)
    X0 = np.random.default_rng(1).uniform(size=(n, d))
    Xs, stress_s, it_s = _mds._smacof_single(D, metric=True, n_components=d, init=X0.copy(), max_iter=300, eps=eps,
                                             normalized_stress=False)
    Xo, stress_o, it_o = mo.smacof_single(D, X0, max_iter=300, eps=eps, rule=1)
    assert it_o == it_s
    np.testing.assert_allclose(Xo, Xs, rtol=1e-9, atol=1e-9)
    assert abs(stress_o - stress_s) <= 1e-9 * max(1.0, stress_s)


def test_rule0_descends_and_stops():
    D = _dist(80, 4, seed=3)
    X0 = np.random.default_rng(2).uniform(size=(80, 4))
    X, stress, n_iter = mo.smacof_single(D, X0, max_iter=300, eps=1e-3, rule=0)
    X1, stress1, _ = mo.smacof_single(D, X0, max_iter=1, rule=0)
    assert 1 < n_iter < 300 and stress < stress1          # SMACOF is a descent method
    # same Guttman sequence as rule 1: n iterations of either rule give the same configuration
    Xa, _, _ = mo.smacof_single(D, X0, max_iter=5, eps=-np.inf, rule=0)
    Xb, _, _ = mo.smacof_single(D, X0, max_iter=5, eps=-np.inf, rule=1)
    np.testing.assert_allclose(Xa, Xb, rtol=1e-12, atol=1e-12)


def test_mds_fit_consumes_global_random_state_like_sklearn():
    D = _dist(30, 2, seed=5)
    np.random.seed(7)
    Xa, sa, _ = mo.mds_fit(D, 2)
    np.random.seed(7)
    starts = [np.random.uniform(size=60).reshape(30, 2) for _ in range(4)]
    runs = [mo.smacof_single(D, s) for s in starts]
    best = min(runs, key=lambda r: r[1])
    np.testing.assert_array_equal(Xa, best[0])
