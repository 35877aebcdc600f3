"""Host mirror of ``STMiner/Algorithm/algorithm.py``: ``linear_sum`` on the device LSAP kernel, ``cluster``
(MDS + KMeans / spectral) on the host with scikit-learn exactly as the reference calls it."""
from __future__ import annotations

import numpy as np
import pandas as pd

from .. import _lib


def linear_sum_batched(costs):
    """Optimal assignment cost of every K x K matrix of ``costs`` [B, K, K] (numpy or CUDA tensor, float64)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    t = costs if isinstance(costs, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(costs, dtype=np.float64))
    t = t.to(device="cuda", dtype=torch.float64).contiguous()
    if t.dim() != 3 or t.shape[1] != t.shape[2]:
        raise ValueError("expected cost matrices of shape [B, K, K], got %s" % (tuple(t.shape),))
    out = torch.empty((t.shape[0],), dtype=torch.float64, device=t.device)
    st = lib.stm_lsap_batched(_lib.ptr(t), int(t.shape[1]), int(t.shape[0]), _lib.ptr(out), _lib.current_stream_ptr())
    _lib.check(st, "stm_lsap_batched")
    return out


def linear_sum(cost) -> float:
    """Drop-in for ``linear_sum`` (algorithm.py:10-26): minimal total cost of a complete assignment.
    Like scipy, non-finite entries are an error (scipy: 'matrix contains invalid numeric entries')."""
    cost = np.asarray(cost, dtype=np.float64)
    if cost.ndim != 2 or cost.shape[0] != cost.shape[1]:
        raise ValueError("the device LSAP takes square cost matrices (the hot path's K x K costs)")
    v = float(linear_sum_batched(cost[None]).cpu().numpy()[0])
    if not np.isfinite(v):
        raise ValueError("matrix contains invalid numeric entries")
    return v


def cluster(distance_array: pd.DataFrame, mds_components: int = 20, n_clusters: int = 10, method: str = "kmeans"):
    """Drop-in for ``cluster`` (algorithm.py:29-61): MDS on the precomputed distances, then KMeans
    (``random_state=0, max_iter=1000, n_init=20``) or spectral clustering; returns
    ``(DataFrame[gene_id, labels], fit_result, embedding)``.  Runs on the host with scikit-learn — the
    'next' row of SURVEY.md §8f; MDS is unseeded in the reference too (``np.random.seed`` fixes it)."""
    from sklearn.cluster import KMeans, SpectralClustering
    from sklearn.manifold import MDS
    import inspect
    index = distance_array.index
    kw = {"n_components": mds_components}
    params = inspect.signature(MDS.__init__).parameters
    if "metric" in params and "dissimilarity" in params and params["metric"].default is not True:
        kw["metric"] = "precomputed"          # sklearn >= 1.8 spelling
    else:
        kw["dissimilarity"] = "precomputed"   # the reference's sklearn 1.3 spelling (algorithm.py:54)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        embedding_position = MDS(**kw).fit_transform(np.asarray(distance_array))
    if method == "kmeans":
        fit_result = KMeans(n_clusters=n_clusters, random_state=0, max_iter=1000, n_init=20).fit(embedding_position)
    elif method == "spectral":
        fit_result = SpectralClustering(n_clusters=n_clusters, random_state=0, affinity="rbf").fit(embedding_position)
    else:
        raise ValueError("method should be kmeans or spectral")
    result = pd.DataFrame(dict(gene_id=list(index), labels=list(fit_result.labels_)))
    return result, fit_result, embedding_position
