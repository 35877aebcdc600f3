"""GPU: size-independent properties of the hot path at sizes the float64 oracle cannot reach in seconds
(Stereo-seq bin50-shaped data, thousands of genes, millions of pairs)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def s50_fit():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from stminer_b200.Algorithm.distribution import fit_spot_matrix
    from stminer_b200.Simulate import simulate_spot_matrix
    from stminer_b200.staging import DeviceSpotMatrix
    sm = simulate_spot_matrix("S50", seed=5, n_templates=12, replicas=100)       # 1,200 genes x 40,000 bins
    res = fit_spot_matrix(DeviceSpotMatrix(sm), 20, init=None, seed=7).to_host()
    return sm, res


def test_fit_invariants_at_scale(s50_fit):
    sm, res = s50_fit
    assert sm.n_genes == 1200 and sm.n_spots == 40000
    assert (res["status"] == 0).all() and (res["n_iter"] >= 2).all() and (res["n_iter"] <= 100).all()
    np.testing.assert_allclose(res["weights"].sum(axis=1), 1.0, rtol=0, atol=1e-9)
    assert (res["weights"] > 0).all()
    cov = res["covariances"]
    assert np.array_equal(cov[..., 0, 1], cov[..., 1, 0])
    det = cov[..., 0, 0] * cov[..., 1, 1] - cov[..., 0, 1] ** 2
    assert (cov[..., 0, 0] >= 1e-3).all() and (cov[..., 1, 1] >= 1e-3).all() and (det > 0).all()   # reg_covar floor, PD
    mu = res["means"]
    assert (mu[..., 0] > -1).all() and (mu[..., 0] < 200).all() and (mu[..., 1] > -1).all() and (mu[..., 1] < 200).all()
    # the mixture mean equals the count-weighted centroid of the gene (exact property of every M-step)
    for g in (0, 377, 1199):
        xy, c = sm.gene_spots(g)
        centroid = (xy * c[:, None]).sum(axis=0) / c.sum()
        np.testing.assert_allclose((res["weights"][g][:, None] * mu[g]).sum(axis=0), centroid, rtol=1e-5)


def test_fit_is_independent_of_gene_order_and_batch(s50_fit):
    """Genes are independent (distribution.py:184) and the device init is keyed by the gene's data: any subset, in any
    order, reproduces the full run bit for bit."""
    from stminer_b200.Algorithm.distribution import fit_spot_matrix
    from stminer_b200.staging import DeviceSpotMatrix
    sm, res = s50_fit
    perm = np.random.default_rng(0).permutation(sm.n_genes)[:300]
    sub = fit_spot_matrix(DeviceSpotMatrix(sm.select(perm)), 20, init=None, seed=7).to_host()
    for k in ("weights", "means", "covariances", "n_iter", "lower_bound", "status"):
        assert np.array_equal(sub[k], res[k][perm]), k


def test_distance_matrix_properties_at_scale(s50_fit):
    from stminer_b200.Algorithm.distance import gmm_one_vs_all, gmm_pair_distances, pairs_to_square
    sm, res = s50_fit
    W, MU, COV = res["weights"], res["means"], res["covariances"]
    G = len(W)
    pairs = gmm_pair_distances(W, MU, COV)
    assert pairs.numel() == G * (G - 1) // 2
    D = pairs_to_square(pairs, G).cpu().numpy()
    assert np.array_equal(D, D.T) and (np.diag(D) == 0).all()
    assert (D >= 0).all() and (D <= 2.0 + 1e-12).all()          # sum_i (w1_i + w2_sigma(i)) * H, H in [0, 1]
    # any slice of the linearised triangle reproduces the full run; one-vs-all is the same arithmetic
    a, b = 123457, 345679
    assert np.array_equal(gmm_pair_distances(W, MU, COV, a, b).cpu().numpy(), pairs[a:b].cpu().numpy())
    for q in (0, 611):
        ova = gmm_one_vs_all(W[q], MU[q], COV[q], W, MU, COV).cpu().numpy()
        mask = np.arange(G) != q
        np.testing.assert_allclose(ova[mask], D[q][mask], rtol=1e-12, atol=1e-14)   # (j, q) pairs see the transposed costs
        assert np.array_equal(ova[q + 1:], D[q][q + 1:]) and ova[q] < 1e-6          # d(g, g): identity assignment, H = 0
    # replicas of one template are closer to each other than to other templates, on average
    same = np.mean([D[i, j] for i in range(0, 100, 7) for j in range(1, 100, 11) if i != j])
    other = np.mean([D[i, 100 + j] for i in range(0, 100, 7) for j in range(1, 1100, 97)])
    assert same < other


def test_emd_metric_properties(cuda_device):
    """Exact LP optimum: EMD(a, a) = 0, symmetric in its arguments, invariant under a common translation, and scaling
    the coordinates by s scales the squared-Euclidean cost by s^2."""
    from stminer_b200.Algorithm.distance import emd_batched
    rng = np.random.default_rng(3)
    R = 40
    def img(n):
        s = rng.permutation(R * R)[:n]
        return (s // R).astype(np.int32), (s % R).astype(np.int32), rng.gamma(1.0, 1.0, n) + 0.01
    a, b, c = img(700), img(450), img(900)
    (ab, aa, ac), st = emd_batched(a, [b, a, c])
    assert (st == 0).all() and aa == 0.0
    (ba,), _ = emd_batched(b, [a])
    assert abs(ab - ba) <= 1e-9 * ab
    sh = lambda t: (t[0] + 17, t[1] + 5, t[2])
    (ab_shift,), _ = emd_batched(sh(a), [sh(b)])
    assert abs(ab - ab_shift) <= 1e-9 * ab
    sc = lambda t: (t[0] * 3, t[1] * 3, t[2])
    (ab_scaled,), _ = emd_batched(sc(a), [sc(b)])
    assert abs(ab_scaled - 9.0 * ab) <= 1e-9 * ab_scaled
    mass = lambda t, f: (t[0], t[1], t[2] * f)
    (ab_mass,), _ = emd_batched(mass(a, 7.5), [mass(b, 0.2)])         # masses are normalised: a / sum(a)
    assert abs(ab - ab_mass) <= 1e-9 * ab
