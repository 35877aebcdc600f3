"""GPU parity: device SMACOF (through the C ABI) vs the numpy oracle and vs the installed scikit-learn."""
import numpy as np
import pandas as pd
import pytest

from oracle import mds_oracle as mo

pytestmark = pytest.mark.gpu


def _dist(n, d, seed):
    rng = np.random.default_rng(seed)
    P = rng.normal(size=(n, d + 2))
    return np.sqrt(((P[:, None] - P[None]) ** 2).sum(-1))


@pytest.mark.parametrize("n,d", [(50, 2), (333, 20), (257, 7), (64, 32), (9, 1)])
@pytest.mark.parametrize("rule", [0, 1])
def test_smacof_matches_oracle(cuda_device, n, d, rule):
    from stminer_b200.Algorithm.algorithm import mds_smacof
    D = _dist(n, d, seed=n + d)
    X0 = np.random.default_rng(3).uniform(size=(n, d))
    eps = 1e-3 if rule == 0 else 1e-6
    X, stress, n_iter = mds_smacof(D, X0, max_iter=300, eps=eps, rule=rule)
    Xo, so, ito = mo.smacof_single(D, X0, max_iter=300, eps=eps, rule=rule)
    assert int(n_iter.cpu()[0]) == ito
    np.testing.assert_allclose(X.cpu().numpy(), Xo, rtol=1e-9, atol=1e-9)
    assert abs(float(stress.cpu()[0]) - so) <= 1e-9 * max(1.0, so)


def test_smacof_matches_installed_sklearn_and_iteration_cap(cuda_device):
    from sklearn.manifold import _mds
    from stminer_b200.Algorithm.algorithm import mds_smacof
    n, d = 200, 20
    D = _dist(n, d, seed=1)
    X0 = np.random.default_rng(4).uniform(size=(n, d))
    Xs, ss, its = _mds._smacof_single(D, metric=True, n_components=d, init=X0.copy(), max_iter=300, eps=1e-6,
                                      normalized_stress=False)
    X, stress, n_iter = mds_smacof(D, X0, max_iter=300, eps=1e-6, rule=1)
    assert int(n_iter.cpu()[0]) == its
    np.testing.assert_allclose(X.cpu().numpy(), Xs, rtol=1e-8, atol=1e-8)
    for rule in (0, 1):     # iteration cap: n_iter = max_iter, configuration after exactly 3 Guttman updates
        X3, s3, it3 = mds_smacof(D, X0, max_iter=3, eps=-1e300, rule=rule)
        Xo, so, ito = mo.smacof_single(D, X0, max_iter=3, eps=-np.inf, rule=rule)
        assert int(it3.cpu()[0]) == ito == 3
        np.testing.assert_allclose(X3.cpu().numpy(), Xo, rtol=1e-11, atol=1e-11)
        assert abs(float(s3.cpu()[0]) - so) <= 1e-9 * so
    # coincident points: dis == 0 -> 1e-5 as sklearn
    X0[5] = X0[6]
    Xc, _, _ = mds_smacof(D, X0, max_iter=2, eps=-1e300, rule=0)
    Xco, _, _ = mo.smacof_single(D, X0, max_iter=2, eps=-np.inf, rule=0)
    np.testing.assert_allclose(Xc.cpu().numpy(), Xco, rtol=1e-9, atol=1e-9)


def test_cluster_labels_match_reference_path(cuda_device):
    """cluster() (algorithm.py:29-61) at a fixed numpy seed: device MDS + KMeans gives the labels of the
    reference's CPU path (oracle SMACOF with the pinned sklearn's semantics + the same KMeans call)."""
    from sklearn.cluster import KMeans
    from stminer_b200.Algorithm.algorithm import cluster
    rng = np.random.default_rng(0)
    centres = rng.normal(scale=4.0, size=(5, 6))
    P = np.concatenate([c + rng.normal(scale=0.6, size=(40, 6)) for c in centres])
    D = np.sqrt(((P[:, None] - P[None]) ** 2).sum(-1))
    names = ["g%d" % i for i in range(len(P))]
    df = pd.DataFrame(D, index=names, columns=names)
    np.random.seed(42)
    res, fit, emb = cluster(df, mds_components=8, n_clusters=5)
    np.random.seed(42)
    emb_ref, _, _ = mo.mds_fit(D, 8)
    ref = KMeans(n_clusters=5, random_state=0, max_iter=1000, n_init=20).fit(emb_ref)
    np.testing.assert_allclose(emb, emb_ref, rtol=1e-8, atol=1e-8)
    assert list(res["labels"]) == list(ref.labels_)
    assert list(res["gene_id"]) == names
    with pytest.raises(ValueError):
        cluster(df, method="nope")


@pytest.mark.parametrize("n,d,k,seed", [(500, 20, 6, 0), (2000, 20, 10, 1), (300, 10, 3, 2), (1200, 5, 8, 3), (64, 20, 2, 4)])
def test_device_kmeans_matches_sklearn(cuda_device, n, d, k, seed):
    """cluster()'s KMeans(n_clusters, random_state=0, max_iter=1000, n_init=20) (algorithm.py:57): same labels, centres and
    inertia as scikit-learn from the same k-means++ draws — blobs with overlap, like an MDS embedding of gene distances."""
    from sklearn.cluster import KMeans
    from stminer_b200.Algorithm.algorithm import kmeans_fit
    rng = np.random.default_rng(seed)
    centres = rng.normal(0, 1.0, (k + 2, d))
    X = centres[rng.integers(0, k + 2, n)] + rng.normal(0, 0.6, (n, d))
    ref = KMeans(n_clusters=k, random_state=0, max_iter=1000, n_init=20).fit(X.copy())
    got = kmeans_fit(X, k, n_init=20, max_iter=1000, random_state=0)
    np.testing.assert_array_equal(got.labels_, ref.labels_)
    np.testing.assert_allclose(got.cluster_centers_, ref.cluster_centers_, rtol=1e-10, atol=1e-12)
    assert abs(got.inertia_ - ref.inertia_) <= 1e-10 * ref.inertia_
    assert got.n_iter_ == ref.n_iter_
