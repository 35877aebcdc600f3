"""GMM-to-GMM and spot-level distances on B200 — host mirror of ``STMiner/Algorithm/distance.py``.

Same function names and return types as the reference (DataFrames keyed by gene, float64); the K x K
Hellinger-cost assignment of every pair runs in one batched launch through the C ABI.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import pandas as pd

from .. import _lib


def stack_gmms(gmm_dict: Dict[str, object]) -> Tuple[list, np.ndarray, np.ndarray, np.ndarray]:
    """{gene: GMM-like} -> (genes, W[G,K], MU[G,K,2], COV[G,K,2,2]) float64, dict order kept (distance.py:91).
    All mixtures must have the component count of the first (distance.py:29 assumes equal K)."""
    genes = list(gmm_dict.keys())
    if not genes:
        return genes, np.zeros((0, 0)), np.zeros((0, 0, 2)), np.zeros((0, 0, 2, 2))
    K = np.asarray(gmm_dict[genes[0]].weights_).size
    W = np.stack([np.asarray(gmm_dict[g].weights_, dtype=np.float64).reshape(K) for g in genes])
    MU = np.stack([np.asarray(gmm_dict[g].means_, dtype=np.float64).reshape(K, 2) for g in genes])
    COV = np.stack([np.asarray(gmm_dict[g].covariances_, dtype=np.float64).reshape(K, 2, 2) for g in genes])
    return genes, W, MU, COV


def _dev(a, device):
    torch = _lib.require_cuda()
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=torch.float64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(device)


def gmm_pair_distances(W, MU, COV, pair_begin: int = 0, pair_end: Optional[int] = None, device=None):
    """Distances of the linearised upper-triangle pairs [pair_begin, pair_end) as a CUDA float64 tensor."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    device = torch.device(device if device is not None else "cuda")
    w, mu, cov = _dev(W, device), _dev(MU, device), _dev(COV, device)
    G, K = int(w.shape[0]), int(w.shape[1])
    total = G * (G - 1) // 2
    pair_end = total if pair_end is None else int(pair_end)
    out = torch.empty((max(pair_end - pair_begin, 0),), dtype=torch.float64, device=device)
    with torch.cuda.device(device):
        st = lib.stm_gmm_dist_pairs(_lib.ptr(w), _lib.ptr(mu), _lib.ptr(cov), G, K, int(pair_begin), pair_end,
                                    _lib.ptr(out), _lib.current_stream_ptr())
    _lib.check(st, "stm_gmm_dist_pairs")
    return out


def pairs_to_square(pairs, G: int):
    """Upper-triangle pair vector (CUDA tensor) -> dense symmetric G x G CUDA tensor, zero diagonal."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    out = torch.empty((G, G), dtype=torch.float64, device=pairs.device)
    with torch.cuda.device(pairs.device):
        st = lib.stm_gmm_dist_fill_square(_lib.ptr(pairs), int(G), _lib.ptr(out), _lib.current_stream_ptr())
    _lib.check(st, "stm_gmm_dist_fill_square")
    return out


def build_gmm_distance_array(gmm_dict, device=None):
    """Drop-in for ``build_gmm_distance_array`` (distance.py:83-99): symmetric float64 DataFrame with a zero
    diagonal, index = columns = ``list(gmm_dict.keys())``."""
    genes, W, MU, COV = stack_gmms(gmm_dict)
    G = len(genes)
    if G == 0:
        return pd.DataFrame(np.zeros((0, 0)), index=genes, columns=genes)
    pairs = gmm_pair_distances(W, MU, COV, device=device)
    square = pairs_to_square(pairs, G).cpu().numpy()
    return pd.DataFrame(square, index=genes, columns=genes)


def gmm_one_vs_all(qw, qmu, qcov, W, MU, COV, device=None):
    torch = _lib.require_cuda()
    lib = _lib.load()
    device = torch.device(device if device is not None else "cuda")
    w, mu, cov = _dev(W, device), _dev(MU, device), _dev(COV, device)
    q = (_dev(qw, device), _dev(qmu, device), _dev(qcov, device))
    G, K = int(w.shape[0]), int(w.shape[1])
    out = torch.empty((G,), dtype=torch.float64, device=device)
    with torch.cuda.device(device):
        st = lib.stm_gmm_dist_one_vs_all(_lib.ptr(q[0]), _lib.ptr(q[1]), _lib.ptr(q[2]), _lib.ptr(w), _lib.ptr(mu),
                                         _lib.ptr(cov), G, K, _lib.ptr(out), _lib.current_stream_ptr())
    _lib.check(st, "stm_gmm_dist_one_vs_all")
    return out


def compare_gmm_distance(gmm, gmm_dict, device=None):
    """Drop-in for ``compare_gmm_distance`` (distance.py:102-111): one-column DataFrame (column ``0``) indexed
    by gene, sorted ascending."""
    genes, W, MU, COV = stack_gmms(gmm_dict)
    K = W.shape[1]
    d = gmm_one_vs_all(np.asarray(gmm.weights_, dtype=np.float64).reshape(K),
                       np.asarray(gmm.means_, dtype=np.float64).reshape(K, 2),
                       np.asarray(gmm.covariances_, dtype=np.float64).reshape(K, 2, 2), W, MU, COV, device=device)
    df = pd.DataFrame.from_dict(dict(zip(genes, d.cpu().numpy())), orient="index")
    return df.sort_values(0)


def distribution_distance(first_gmm, second_gmm, device=None) -> float:
    """Drop-in for ``distribution_distance`` (distance.py:13-36) — a batch of one pair."""
    _, W, MU, COV = stack_gmms({"a": first_gmm, "b": second_gmm})
    return float(gmm_pair_distances(W, MU, COV, device=device).cpu().numpy()[0])
