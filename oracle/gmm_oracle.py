"""Oracle for the per-gene 2-D Gaussian-mixture fit (hot path 1).  TEST INFRASTRUCTURE ONLY.

Reference path restated here (paths relative to the STMiner tree):

* ``STMiner/Algorithm/AlgUtils.py:8-36``   gene column -> int16 count grid
* ``STMiner/Algorithm/distribution.py:369-373`` ``array_to_list`` (replicate coords by count)
* ``STMiner/Algorithm/distribution.py:87-130``  ``fit_gmm`` (distinct-spot guard + sklearn fit)
* ``STMiner/Algorithm/distribution.py:148-205`` ``fit_gmms`` (serial loop, dropped-gene rule)

The EM arithmetic itself lives in a third-party dependency that is NOT under
/root/reference: scikit-learn ``GaussianMixture`` (reference pins
``scikit_learn==1.3.0``, requirements.txt:12; this image has 1.9.0 whose EM
numerics are the same).  ``weighted_em`` restates its published algorithm
(sklearn ``mixture/_base.py`` fit loop, ``_gaussian_mixture.py``
``_estimate_gaussian_parameters`` / ``_compute_precision_cholesky`` /
``_estimate_log_gaussian_prob``) in the count-weighted form, and
``sklearn_fit_from_init`` runs the real sklearn class on the replicated list
exactly as the reference does.

Pinning: the reference holds NO golden vectors for GMM parameters
(tests/test_STMiner.py:54-58 only asserts ``len(patterns) > 0``), so this oracle
is pinned against the real ``sklearn.mixture.GaussianMixture`` run in this
image on the replicated list from identical injected initial parameters
(``tests/test_oracle_gmm.py``; committed fixtures ``tests/golden/gmm_*.npz``
made by ``tests/golden/make_golden.py``).
"""
from __future__ import annotations

import numpy as np

LOG_2PI = np.log(2.0 * np.pi)


# --------------------------------------------------------------------------- staging
def gene_column_to_counts(x, y, values, remove_low_exp_spots: bool = False, binary: bool = False, cut: bool = False,
                          threshold: float = 95):
    """Gene column -> (xy[nnz,2] int64, counts[nnz] int64) on DISTINCT positive spots.

    Follows ``AlgUtils._preprocess`` (AlgUtils.py:26-36): the column is cast to
    int16 by ``coo_matrix(..., dtype=np.int16)`` (truncation toward zero),
    ``todense()`` sums duplicate (x, y) (AlgUtils.py:15); optional
    ``max(dense - mean(nonzero), 0)`` (AlgUtils.py:16-17) and ``np.round`` ->
    int16 (AlgUtils.py:22-23); optional ``binary`` / ``cut`` of ``preprocess_array``
    (distribution.py:133-145: percentile over the whole dense grid).  ``array_to_list``
    (distribution.py:369-373) then keeps cells ``> 0`` in row-major (x, y) order.
    """
    x = np.asarray(x).astype(np.int64)
    y = np.asarray(y).astype(np.int64)
    v16 = np.asarray(values).astype(np.int16)          # truncation toward zero, as coo_matrix(dtype=int16)
    shape = (int(x.max()) + 1, int(y.max()) + 1)
    dense = np.zeros(shape, dtype=np.int16)
    np.add.at(dense, (x, y), v16)                       # duplicates summed in int16 (todense)
    if remove_low_exp_spots:
        nz = dense[dense != 0]
        d = np.maximum(dense - np.mean(nz), 0)
        dense = np.array(np.round(d), dtype=np.int16)
    if binary:                                          # preprocess_array, distribution.py:135-140
        if threshold > 100 or threshold < 0:
            threshold = 100
        dense = np.where(dense > np.percentile(dense, threshold), 1, 0)
    elif cut:                                           # distribution.py:142-143
        dense = dense.copy()
        dense[dense < np.percentile(dense, threshold)] = 0
    xs, ys = np.where(dense > 0)
    counts = dense[xs, ys].astype(np.int64)
    return np.column_stack([xs, ys]).astype(np.int64), counts


def array_to_list(xy, counts):
    """Replicated coordinate list, distribution.py:369-373 (``np.repeat(coords, counts)``)."""
    return np.repeat(np.asarray(xy), np.asarray(counts), axis=0)


# --------------------------------------------------------------------------- EM restatement
def _chol2x2_params(cov):
    """cov (K,2,2) -> lower Cholesky L and sklearn's precisions_cholesky U = inv(L).T."""
    L = np.linalg.cholesky(cov)                         # raises LinAlgError like sklearn -> ValueError
    K = cov.shape[0]
    U = np.empty_like(cov)
    for k in range(K):
        U[k] = np.linalg.solve(L[k], np.eye(2)).T       # _gaussian_mixture.py:361-370
    return L, U


def weighted_log_prob(xy, weights, means, U):
    """log(w_k N(x_i | mu_k, Sigma_k)) for all i,k; _gaussian_mixture.py:490-553 + _base.py:524."""
    X = np.asarray(xy, dtype=np.float64)
    K = means.shape[0]
    lp = np.empty((X.shape[0], K))
    for k in range(K):
        y = (X - means[k]) @ U[k]
        log_det = np.log(U[k, 0, 0]) + np.log(U[k, 1, 1])
        lp[:, k] = -0.5 * (2 * LOG_2PI + np.sum(y * y, axis=1)) + log_det
    return lp + np.log(weights)


def _logsumexp(a):
    m = np.max(a, axis=1)
    return m + np.log(np.sum(np.exp(a - m[:, None]), axis=1))


def weighted_em(xy, counts, weights_init, means_init, covs_init,
                reg_covar=1e-3, tol=1e-3, max_iter=100):
    """Count-weighted EM on distinct spots, float64; mathematically the reference's
    replicated-list fit (distribution.py:126-127 -> sklearn fit loop _base.py:257-278).

    Returns dict(weights, means, covariances, n_iter, converged, lower_bound).
    """
    X = np.asarray(xy, dtype=np.float64)
    c = np.asarray(counts, dtype=np.float64)
    N = c.sum()
    w = np.asarray(weights_init, dtype=np.float64).copy()
    mu = np.asarray(means_init, dtype=np.float64).copy()
    cov = np.asarray(covs_init, dtype=np.float64).copy()
    K = w.shape[0]
    _, U = _chol2x2_params(cov)
    lower_bound = -np.inf
    converged = False
    n_iter = 0
    eps10 = 10 * np.finfo(np.float64).eps
    for n_iter in range(1, max_iter + 1):
        prev = lower_bound
        # E-step (_base.py:552-582)
        lp = weighted_log_prob(X, w, mu, U)
        lse = _logsumexp(lp)
        resp = np.exp(lp - lse[:, None])
        lower_bound = float(np.sum(c * lse) / N)         # mean over replicated rows (_base.py:332)
        # M-step (_gaussian_mixture.py:282-320, 883-901)
        cr = resp * c[:, None]
        nk = cr.sum(axis=0) + eps10
        mu = (cr.T @ X) / nk[:, None]
        for k in range(K):
            d = X - mu[k]
            cov[k] = (cr[:, k] * d.T) @ d / nk[k]
            cov[k, 0, 0] += reg_covar
            cov[k, 1, 1] += reg_covar
        w = nk / nk.sum()
        _, U = _chol2x2_params(cov)
        if abs(lower_bound - prev) < tol:
            converged = True
            break
    return dict(weights=w, means=mu, covariances=cov, n_iter=n_iter,
                converged=converged, lower_bound=lower_bound)


def sklearn_fit_from_init(xy, counts, weights_init, means_init, covs_init,
                          reg_covar=1e-3, tol=1e-3, max_iter=100):
    """The real sklearn GaussianMixture on the replicated list (distribution.py:126-127)
    with injected initial parameters (KMeans skipped: _gaussian_mixture.py:836-847)."""
    from sklearn.mixture import GaussianMixture
    import warnings
    X = array_to_list(xy, counts).astype(np.float64)
    K = len(weights_init)
    gm = GaussianMixture(n_components=K, max_iter=max_iter, reg_covar=reg_covar, tol=tol,
                         weights_init=np.asarray(weights_init, dtype=np.float64),
                         means_init=np.asarray(means_init, dtype=np.float64),
                         precisions_init=np.linalg.inv(np.asarray(covs_init, dtype=np.float64)))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                   # distribution.py:20
        gm.fit(X)
    return dict(weights=gm.weights_, means=gm.means_, covariances=gm.covariances_,
                n_iter=gm.n_iter_, converged=gm.converged_, lower_bound=gm.lower_bound_)


# --------------------------------------------------------------------------- initial parameters
def params_from_labels(xy, counts, labels, K, reg_covar=1e-3):
    """One-hot responsibilities -> (w, mu, cov) exactly as sklearn's ``_initialize``
    (_gaussian_mixture.py:849-881 via _estimate_gaussian_parameters :282-320)."""
    X = np.asarray(xy, dtype=np.float64)
    c = np.asarray(counts, dtype=np.float64)
    resp = np.zeros((X.shape[0], K))
    resp[np.arange(X.shape[0]), labels] = 1.0
    cr = resp * c[:, None]
    nk = cr.sum(axis=0) + 10 * np.finfo(np.float64).eps
    mu = (cr.T @ X) / nk[:, None]
    cov = np.empty((K, 2, 2))
    for k in range(K):
        d = X - mu[k]
        cov[k] = (cr[:, k] * d.T) @ d / nk[k]
        cov[k, 0, 0] += reg_covar
        cov[k, 1, 1] += reg_covar
    w = nk / c.sum()
    return w, mu, cov


def kmeans_init(xy, counts, K, seed=0, reg_covar=1e-3):
    """Initial (w, mu, cov) the way GaussianMixture's default ``init_params='kmeans'`` makes
    them (_base.py:119-128): KMeans(K, n_init=1) on the replicated list, one-hot resp, M-step."""
    from sklearn.cluster import KMeans
    X = array_to_list(xy, counts).astype(np.float64)
    km = KMeans(n_clusters=K, n_init=1, random_state=seed).fit(X)
    # label of each distinct spot (all replicas of a spot share a label)
    first = np.concatenate([[0], np.cumsum(counts)[:-1]])
    labels = km.labels_[first]
    return params_from_labels(xy, counts, labels, K, reg_covar)


# --------------------------------------------------------------------------- reference fit path (CPU baseline)
def fit_gmm_reference(xy, counts, n_comp, max_iter=100, reg_covar=1e-3, random_state=None):
    """``fit_gmm`` as the reference runs it (distribution.py:122-130): distinct-spot guard,
    then an un-initialised sklearn fit (KMeans init included).  Returns the sklearn object or None."""
    from sklearn.mixture import GaussianMixture
    import warnings
    if len(counts) < n_comp:                              # distribution.py:125,129-130
        return None
    X = array_to_list(xy, counts)
    gm = GaussianMixture(n_components=n_comp, max_iter=max_iter, reg_covar=reg_covar,
                         random_state=random_state)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gm.fit(X)
    return gm
