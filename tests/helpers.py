"""Shared helpers for the parity tests (the oracle is the checker, never the product path)."""
import numpy as np

from oracle import gmm_oracle as go


def oracle_inits(sm, K, seed=0, reg_covar=1e-3):
    """Per-gene initial parameters made the way the reference's GaussianMixture would (KMeans init)."""
    G = sm.n_genes
    W = np.zeros((G, K)); MU = np.zeros((G, K, 2)); COV = np.tile(np.eye(2), (G, K, 1, 1)).astype(np.float64)
    ok = np.zeros(G, dtype=bool)
    for g in range(G):
        xy, c = sm.gene_spots(g)
        if len(c) < K:
            W[g] = 1.0 / K
            continue
        W[g], MU[g], COV[g] = go.kmeans_init(xy, c, K, seed=seed + g, reg_covar=reg_covar)
        ok[g] = True
    return W, MU, COV, ok


def oracle_fit_all(sm, K, W, MU, COV, ok, **kw):
    out = []
    for g in range(sm.n_genes):
        if not ok[g]:
            out.append(None)
            continue
        xy, c = sm.gene_spots(g)
        out.append(go.weighted_em(xy, c, W[g], MU[g], COV[g], **kw))
    return out


def assert_gmm_close(res, g, ref, rtol=1e-4, what=""):
    """rtol 1e-4 on weights / means / covariances (north-star tolerance, FP32 E-step)."""
    w, mu, cov = res["weights"][g], res["means"][g], res["covariances"][g]
    # absolute floors: a weight is compared relative to 1/K-scale mass, a mean relative to the grid scale,
    # a covariance entry relative to that component's own variance scale
    np.testing.assert_allclose(w, ref["weights"], rtol=rtol, atol=rtol * 1e-2, err_msg="weights " + what)
    np.testing.assert_allclose(mu, ref["means"], rtol=rtol, atol=rtol, err_msg="means " + what)
    scale = np.sqrt(ref["covariances"][:, 0, 0] * ref["covariances"][:, 1, 1])[:, None, None]
    np.testing.assert_allclose(cov / scale, ref["covariances"] / scale, rtol=rtol, atol=rtol,
                               err_msg="covariances " + what)
