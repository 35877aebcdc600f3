"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the MDS step of ``cluster`` — STMiner/Algorithm/algorithm.py:54-55.

The arithmetic lives in a third-party dependency: scikit-learn, pinned ``scikit_learn==1.3.0`` by the reference
(requirements.txt:12); this image has 1.9.0, whose SMACOF uses the same Guttman update but a different stopping
rule and different ``MDS`` defaults.  So the published 1.3.0 algorithm is restated here in numpy:

  * ``smacof_single(rule=0)``  — sklearn 1.3.0 ``manifold/_mds.py::_smacof_single`` (metric branch):
        dis = euclidean_distances(X); stress = ((dis - D)**2).sum() / 2; dis[dis == 0] = 1e-5
        X = 1/n * B X  with  B = -D/dis, B_ii += sum_j D_ij/dis_ij
        stop when  old - stress / sum_i ||X_i||  <  eps   (old = previous value of stress / norm)
  * ``smacof_single(rule=1)``  — the installed sklearn's rule (same update; stress of the NEW configuration,
        relative decrease) — this branch is what pins the restatement: tests compare it with the real
        ``sklearn.manifold._mds._smacof_single`` from the same initial configuration.
  * ``mds_fit``  — ``MDS(n_components, dissimilarity='precomputed')`` of 1.3.0: n_init=4 runs, every run drawing its
        start ``random_state.uniform(size=n*d)`` from the SAME RandomState (global numpy state for random_state=None,
        smacof() serial branch), best final stress wins; max_iter=300, eps=1e-3.

Pinning: rule 1 is checked against the real scikit-learn of this image (tests/test_oracle_mds.py) — the same Guttman
update the pinned version runs; the stopping rule and restart scheme of 1.3.0 (rule 0, ``mds_fit``) are restated from the
published source and cannot be run here: PARITY UNPINNED for the iteration at which rule 0 stops.

Only ``tests/`` and the CPU legs of ``bench.py`` may import this module.
"""
from __future__ import annotations

import numpy as np


def _pairwise(X):
    diff = X[:, None, :] - X[None, :, :]
    return np.sqrt((diff * diff).sum(axis=2))


def smacof_single(D, X0, max_iter=300, eps=1e-3, rule=0):
    """One SMACOF run from ``X0``; returns (X, stress, n_iter) with sklearn's meanings."""
    D = np.asarray(D, dtype=np.float64)
    X = np.array(X0, dtype=np.float64)
    n = D.shape[0]
    old = None
    if rule == 0:
        for it in range(max_iter):
            dis = _pairwise(X)
            stress = ((dis.ravel() - D.ravel()) ** 2).sum() / 2
            dis[dis == 0] = 1e-5
            ratio = D / dis
            B = -ratio
            B[np.arange(n), np.arange(n)] += ratio.sum(axis=1)
            X = 1.0 / n * np.dot(B, X)
            nrm = np.sqrt((X ** 2).sum(axis=1)).sum()
            if old is not None and (old - stress / nrm) < eps:
                break
            old = stress / nrm
        return X, stress, it + 1
    dis = _pairwise(X)
    for it in range(max_iter):
        dis[dis == 0] = 1e-5
        ratio = D / dis
        B = -ratio
        B[np.arange(n), np.arange(n)] += ratio.sum(axis=1)
        X = 1.0 / n * np.dot(B, X)
        dis = _pairwise(X)
        stress = ((dis.ravel() - D.ravel()) ** 2).sum() / 2
        if old is not None:
            ssd = (dis.ravel() ** 2).sum()
            if (old - stress) / (ssd / 2) < eps:
                break
        old = stress
    return X, stress, it + 1


def mds_fit(D, n_components, n_init=4, max_iter=300, eps=1e-3, random_state=None, rule=0):
    """Embedding of ``MDS(n_components, dissimilarity='precomputed').fit_transform(D)`` as scikit-learn 1.3.0."""
    rs = np.random.mtrand._rand if random_state is None else random_state
    n = np.asarray(D).shape[0]
    best = None
    for _ in range(n_init):
        X0 = rs.uniform(size=n * n_components).reshape((n, n_components))
        X, stress, n_iter = smacof_single(D, X0, max_iter=max_iter, eps=eps, rule=rule)
        if best is None or stress < best[1]:
            best = (X, stress, n_iter)
    return best
