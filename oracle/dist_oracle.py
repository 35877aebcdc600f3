"""Oracle for the GMM-to-GMM distance (hot path 2b).  TEST INFRASTRUCTURE ONLY.

Restates, with the reference's own third-party solver:

* ``STMiner/Algorithm/distance.py:50-80``   ``get_hellinger_distance`` (numba njit in the reference; the
  same NumPy calls — ``np.linalg.det`` / ``np.linalg.inv`` — are used here, jitted when numba is present)
* ``STMiner/Algorithm/distance.py:13-36``   ``distribution_distance`` (K x K cost, (w1_i + w2_j) * H_ij)
* ``STMiner/Algorithm/algorithm.py:10-26``  ``linear_sum`` -> ``scipy.optimize.linear_sum_assignment``
  (reference pins scipy==1.11.1, requirements.txt:13; this image has 1.18.1 — same C++ LSAP)
* ``STMiner/Algorithm/distance.py:83-99``   ``build_gmm_distance_array`` (i<j loop, mirrored, zero diagonal)
* ``STMiner/Algorithm/distance.py:102-111`` ``compare_gmm_distance`` (one-vs-all, sorted ascending)

Pinning: the reference holds no golden values for these (tests/test_STMiner.py:61-64 checks shape only;
readme.md:205-216 gives the range 0.89 .. 1.9997).  ``STMiner/Algorithm/algorithm.py`` itself imports
stand-alone in this image, so ``tests/golden/make_golden.py`` runs the reference's own ``linear_sum`` on the
cost matrices and commits the results; analytic cases (identical GMMs -> 0, far-apart GMMs -> 2) are
tested as well.
"""
from __future__ import annotations

import numpy as np
from scipy.optimize import linear_sum_assignment

try:  # the reference jit-compiles this function (distance.py:50)
    from numba import njit
except Exception:  # pragma: no cover
    def njit(f):
        return f


@njit
def get_hellinger_distance(gmm1_covs, gmm1_means, gmm2_covs, gmm2_means):
    """distance.py:71-80."""
    mean_cov = (gmm1_covs + gmm2_covs) / 2
    mean_cov_det = np.linalg.det(mean_cov)
    mean_cov_inv = np.linalg.inv(mean_cov)
    d1 = np.linalg.det(gmm1_covs)
    d2 = np.linalg.det(gmm2_covs)
    diff = gmm1_means - gmm2_means
    first = np.exp(-(diff.T @ mean_cov_inv @ diff) / 8)
    second = (np.power(d1, 0.25) * np.power(d2, 0.25)) / np.sqrt(mean_cov_det)
    return np.sqrt(np.abs(1 - first * second))


def cost_matrix(w1, mu1, cov1, w2, mu2, cov2):
    """distance.py:29-34."""
    K = w1.size
    C = np.zeros((K, K))
    for i in range(K):
        for j in range(K):
            C[i, j] = (w1[i] + w2[j]) * get_hellinger_distance(
                np.ascontiguousarray(cov1[i]), np.ascontiguousarray(mu1[i]),
                np.ascontiguousarray(cov2[j]), np.ascontiguousarray(mu2[j]))
    return C


def linear_sum(cost):
    """algorithm.py:24-26."""
    r, c = linear_sum_assignment(cost)
    return cost[r, c].sum()


def distribution_distance(w1, mu1, cov1, w2, mu2, cov2):
    """distance.py:13-36."""
    return linear_sum(cost_matrix(w1, mu1, cov1, w2, mu2, cov2))


def cost_matrix_vectorised(w1, mu1, cov1, w2, mu2, cov2):
    """Closed-form 2x2 version of ``cost_matrix`` — used for the larger parity
    sweeps and the CPU baseline timing where K^2 numba dispatches per pair would dominate."""
    S = 0.5 * (cov1[:, None] + cov2[None])
    det = S[..., 0, 0] * S[..., 1, 1] - S[..., 0, 1] * S[..., 1, 0]
    d = mu1[:, None] - mu2[None]
    q = (S[..., 1, 1] * d[..., 0] ** 2 - (S[..., 0, 1] + S[..., 1, 0]) * d[..., 0] * d[..., 1]
         + S[..., 0, 0] * d[..., 1] ** 2) / det
    d1 = cov1[:, 0, 0] * cov1[:, 1, 1] - cov1[:, 0, 1] * cov1[:, 1, 0]
    d2 = cov2[:, 0, 0] * cov2[:, 1, 1] - cov2[:, 0, 1] * cov2[:, 1, 0]
    t = np.exp(-q / 8) * (d1[:, None] ** 0.25) * (d2[None] ** 0.25) / np.sqrt(det)
    return (w1[:, None] + w2[None]) * np.sqrt(np.abs(1 - t))


def build_gmm_distance_array(W, MU, COV, vectorised=True):
    """distance.py:83-99 on stacked parameters [G,K], [G,K,2], [G,K,2,2] -> dense G x G float64."""
    G = W.shape[0]
    cm = cost_matrix_vectorised if vectorised else cost_matrix
    D = np.zeros((G, G))
    for i in range(G):
        for j in range(i + 1, G):
            D[i, j] = D[j, i] = linear_sum(cm(W[i], MU[i], COV[i], W[j], MU[j], COV[j]))
    return D


def compare_gmm_distance(qw, qmu, qcov, W, MU, COV, vectorised=True):
    """distance.py:102-111 (unsorted vector; the caller sorts)."""
    cm = cost_matrix_vectorised if vectorised else cost_matrix
    return np.array([linear_sum(cm(qw, qmu, qcov, W[g], MU[g], COV[g])) for g in range(W.shape[0])])


def random_gmms(G, K, extent=100.0, seed=0):
    """Synthetic GMM set of SURVEY.md §8d config P: mu ~ U(grid), Sigma = A A^T s + 1e-3 I, w ~ Dirichlet(1)."""
    rng = np.random.default_rng(seed)
    W = rng.dirichlet(np.ones(K), size=G)
    MU = rng.uniform(0, extent, size=(G, K, 2))
    A = rng.normal(size=(G, K, 2, 2))
    s = rng.uniform(0.5, 0.05 * extent, size=(G, K, 1, 1)) ** 2
    COV = A @ np.swapaxes(A, -1, -2) * s + 1e-3 * np.eye(2)
    return W, MU, COV
