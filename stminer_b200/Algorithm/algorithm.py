"""Host mirror of ``STMiner/Algorithm/algorithm.py``: ``linear_sum`` on the device LSAP kernel, ``cluster`` =
device SMACOF MDS (scikit-learn 1.3.0 semantics, the version the reference pins) + scikit-learn KMeans / spectral."""
from __future__ import annotations

import numpy as np
import pandas as pd

from .. import _lib


def linear_sum_batched(costs, device=None):
    """Optimal assignment cost of every K x K matrix of ``costs`` [B, K, K] (numpy or CUDA tensor, float64)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    t = costs if isinstance(costs, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(costs, dtype=np.float64))
    if device is None:
        device = t.device if t.is_cuda else "cuda"
    t = t.to(device=device, dtype=torch.float64).contiguous()
    if t.dim() != 3 or t.shape[1] != t.shape[2]:
        raise ValueError("expected cost matrices of shape [B, K, K], got %s" % (tuple(t.shape),))
    out = torch.empty((t.shape[0],), dtype=torch.float64, device=t.device)
    with torch.cuda.device(t.device):
        st = lib.stm_lsap_batched(_lib.ptr(t), int(t.shape[1]), int(t.shape[0]), _lib.ptr(out), _lib.current_stream_ptr())
    _lib.check(st, "stm_lsap_batched")
    return out


def linear_sum(cost, device=None) -> float:
    """Drop-in for ``linear_sum`` (algorithm.py:10-26): minimal total cost of a complete assignment.
    Like scipy, non-finite entries are an error (scipy: 'matrix contains invalid numeric entries')."""
    cost = np.asarray(cost, dtype=np.float64)
    if cost.ndim != 2 or cost.shape[0] != cost.shape[1]:
        raise ValueError("the device LSAP takes square cost matrices (the hot path's K x K costs)")
    v = float(linear_sum_batched(cost[None], device=device).cpu().numpy()[0])
    if not np.isfinite(v):
        raise ValueError("matrix contains invalid numeric entries")
    return v


def mds_smacof(dissimilarities, init, max_iter: int = 300, eps: float = 1e-3, rule: int = 0, device=None):
    """One SMACOF run on the device from the start configuration ``init`` [n, d] — sklearn's ``_smacof_single``
    (metric MDS on a precomputed matrix).  Returns CUDA tensors ``(X[n, d], stress[1], n_iter[1])``; nothing is
    synchronised.  ``rule`` 0 = stopping rule of scikit-learn <= 1.6 (the reference pins 1.3.0), 1 = scikit-learn >= 1.7."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = torch.device(device if device is not None else "cuda")

    def dv(a):
        t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))
        return t.to(device=dev, dtype=torch.float64).contiguous()
    D, X0 = dv(dissimilarities), dv(init)
    n, d = int(X0.shape[0]), int(X0.shape[1])
    if tuple(D.shape) != (n, n):
        raise ValueError("dissimilarities must be [n, n] with n = init.shape[0]; got %s" % (tuple(D.shape),))
    nbytes = int(lib.stm_mds_workspace_bytes(n, d))
    if nbytes < 0:
        raise ValueError("n_components must be in [1, 32]")
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    X = torch.empty((n, d), dtype=torch.float64, device=dev)
    stress = torch.empty(1, dtype=torch.float64, device=dev)
    n_iter = torch.empty(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        st = lib.stm_mds_smacof(_lib.ptr(D), n, d, _lib.ptr(X0), int(max_iter), float(eps), int(rule), _lib.ptr(ws),
                                nbytes, _lib.ptr(X), _lib.ptr(stress), _lib.ptr(n_iter), _lib.current_stream_ptr())
    _lib.check(st, "stm_mds_smacof")
    X._keepalive = (D, X0, ws)
    return X, stress, n_iter


def mds_fit_transform(dissimilarities, n_components: int = 2, n_init: int = 4, max_iter: int = 300, eps: float = 1e-3,
                      random_state=None, rule: int = 0, device=None):
    """``MDS(n_components, dissimilarity='precomputed').fit_transform(D)`` with the defaults of the scikit-learn the
    reference pins (1.3.0: metric, n_init=4, max_iter=300, eps=1e-3; algorithm.py:54-55): every restart draws its
    start configuration from the same RandomState — numpy's global one when ``random_state`` is None, so
    ``np.random.seed`` fixes the result as it does for the reference — the runs execute on the device and the
    lowest final stress wins.  Returns ``(embedding ndarray [n, d], stress, n_iter)``."""
    D = np.asarray(dissimilarities, dtype=np.float64)
    n = D.shape[0]
    if D.ndim != 2 or D.shape[1] != n:
        raise ValueError("expected a square dissimilarity matrix")
    if isinstance(random_state, (int, np.integer)):
        random_state = np.random.RandomState(int(random_state))
    rs = np.random.mtrand._rand if random_state is None else random_state
    torch = _lib.require_cuda()
    Dd = torch.from_numpy(np.ascontiguousarray(D)).to(device if device is not None else "cuda")
    runs = []
    # the restarts are independent: each goes to its own stream so that their passes (n/8 CTAs each) fill the GPU together
    main = torch.cuda.current_stream(Dd.device)
    streams = [torch.cuda.Stream(device=Dd.device) for _ in range(n_init)] if n_init > 1 else [main]
    for r in range(n_init):
        X0 = rs.uniform(size=n * n_components).reshape((n, n_components))     # drawn in restart order, as sklearn does
        streams[r].wait_stream(main)
        with torch.cuda.stream(streams[r]):
            runs.append(mds_smacof(Dd, X0, max_iter=max_iter, eps=eps, rule=rule, device=Dd.device))
    for st in streams:
        main.wait_stream(st)
    best = None
    for X, stress, n_iter in runs:      # first-lowest wins, as `if best_stress is None or stress < best_stress`
        s = float(stress.cpu().numpy()[0])
        if best is None or s < best[1]:
            best = (X, s, int(n_iter.cpu().numpy()[0]))
    return best[0].cpu().numpy(), best[1], best[2]


class KMeansResult:
    """What ``cluster`` hands back as its fit result: the attributes of a fitted ``sklearn.cluster.KMeans`` that callers
    of the reference read (``labels_``, ``cluster_centers_``, ``inertia_``, ``n_iter_``)."""

    def __init__(self, labels, centers, inertia, n_iter, n_clusters):
        self.labels_, self.cluster_centers_, self.inertia_, self.n_iter_ = labels, centers, inertia, n_iter
        self.n_clusters = n_clusters

    def predict(self, X):
        X = np.asarray(X, dtype=np.float64)
        d2 = ((X[:, None, :] - self.cluster_centers_[None]) ** 2).sum(axis=2)
        return np.argmin(d2, axis=1).astype(np.int32)


def kmeans_fit(X, n_clusters: int, n_init: int = 20, max_iter: int = 1000, tol: float = 1e-4, random_state=0, device=None):
    """``KMeans(n_clusters, random_state, max_iter, n_init).fit(X)`` (algorithm.py:57) with the Lloyd iterations of all
    restarts on the device (``stm_kmeans_lloyd``).  The start centres are drawn exactly as ``KMeans.fit`` draws them —
    scikit-learn's own k-means++ on the mean-centred data, one shared RandomState consumed restart after restart — and the
    best restart is chosen by KMeans' rule (lower inertia and not the same clustering), so the labels are those of
    scikit-learn (tests compare them).  A restart in which a cluster becomes empty (scikit-learn relocates it to the
    farthest point) is not restated: the whole fit is then handed to scikit-learn."""
    from sklearn.cluster import KMeans, kmeans_plusplus
    from sklearn.utils import check_random_state
    from sklearn.utils.extmath import row_norms
    torch = _lib.require_cuda()
    lib = _lib.load()
    X = np.ascontiguousarray(X, dtype=np.float64)
    n, d = X.shape
    if n_clusters > 64 or d > 64 or n < n_clusters:
        return KMeans(n_clusters=n_clusters, random_state=random_state, max_iter=max_iter, n_init=n_init, tol=tol).fit(X)
    rs = check_random_state(random_state)
    tol_abs = float(np.mean(np.var(X, axis=0)) * tol)
    mean = X.mean(axis=0)
    Xc = X - mean
    xsn = row_norms(Xc, squared=True)
    inits = np.stack([kmeans_plusplus(Xc, n_clusters, x_squared_norms=xsn, random_state=rs)[0] for _ in range(n_init)])
    dev = torch.device(device if device is not None else "cuda")
    with torch.cuda.device(dev):
        dX = torch.from_numpy(Xc).to(dev)
        dI = torch.from_numpy(np.ascontiguousarray(inits, dtype=np.float64)).to(dev)
        ws_bytes = int(lib.stm_kmeans_workspace_bytes(n, n_init))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        oc = torch.empty((n_init, n_clusters, d), dtype=torch.float64, device=dev)
        ol = torch.empty((n_init, n), dtype=torch.int32, device=dev)
        oi = torch.empty(n_init, dtype=torch.float64, device=dev)
        on = torch.empty(n_init, dtype=torch.int32, device=dev)
        os_ = torch.empty(n_init, dtype=torch.int32, device=dev)
        st = lib.stm_kmeans_lloyd(_lib.ptr(dX), n, d, int(n_clusters), _lib.ptr(dI), int(n_init), int(max_iter), tol_abs,
                                  _lib.ptr(ws), ws_bytes, _lib.ptr(oc), _lib.ptr(ol), _lib.ptr(oi), _lib.ptr(on), _lib.ptr(os_),
                                  _lib.current_stream_ptr())
        _lib.check(st, "stm_kmeans_lloyd")
        centers, labels, inertia = oc.cpu().numpy(), ol.cpu().numpy(), oi.cpu().numpy()
        n_iter, status = on.cpu().numpy(), os_.cpu().numpy()
    if (status != 0).any():
        import logging
        logging.getLogger(__name__).warning("a KMeans restart emptied a cluster; fitting with scikit-learn on the host")
        return KMeans(n_clusters=n_clusters, random_state=random_state, max_iter=max_iter, n_init=n_init, tol=tol).fit(X)
    best = None
    for r in range(n_init):                 # KMeans.fit: lower inertia AND a different clustering replaces the best
        if best is None or (inertia[r] < inertia[best] and not _same_clustering(labels[r], labels[best], n_clusters)):
            best = r
    return KMeansResult(labels[best].astype(np.int32), centers[best] + mean, float(inertia[best]), int(n_iter[best]), n_clusters)


def _same_clustering(a, b, n_clusters) -> bool:
    """sklearn's ``_is_same_clustering``: equal up to a permutation of the labels."""
    mapping = np.full(n_clusters, -1, dtype=np.int64)
    for x, y in zip(a.tolist(), b.tolist()):
        if mapping[x] == -1:
            mapping[x] = y
        elif mapping[x] != y:
            return False
    return True


def cluster(distance_array: pd.DataFrame, mds_components: int = 20, n_clusters: int = 10, method: str = "kmeans",
            device=None):
    """Drop-in for ``cluster`` (algorithm.py:29-61): MDS on the precomputed distances, then KMeans
    (``random_state=0, max_iter=1000, n_init=20``) or spectral clustering; returns
    ``(DataFrame[gene_id, labels], fit_result, embedding)``.  The MDS (SMACOF, O(G^2 d) per iteration) runs on the
    device with the semantics of the scikit-learn version the reference pins; it is unseeded in the reference too
    (``np.random.seed`` fixes it).  KMeans: scikit-learn's k-means++ starts on the host, the Lloyd iterations of the 20
    restarts on the device (``kmeans_fit``)."""
    from sklearn.cluster import KMeans, SpectralClustering
    if method not in ("kmeans", "spectral"):
        raise ValueError("method should be kmeans or spectral")
    index = distance_array.index
    embedding_position, _, _ = mds_fit_transform(np.asarray(distance_array, dtype=np.float64), mds_components,
                                                 device=device)
    if method == "kmeans":
        fit_result = kmeans_fit(embedding_position, n_clusters, n_init=20, max_iter=1000, random_state=0, device=device)
    else:
        fit_result = SpectralClustering(n_clusters=n_clusters, random_state=0, affinity="rbf").fit(embedding_position)
    result = pd.DataFrame(dict(gene_id=list(index), labels=list(fit_result.labels_)))
    return result, fit_result, embedding_position
