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


def gmm_pair_distances(W, MU, COV, pair_begin: int = 0, pair_end: Optional[int] = None, device=None, out=None):
    """Distances of the linearised upper-triangle pairs [pair_begin, pair_end) as a CUDA float64 tensor (``out``: a
    contiguous float64 CUDA tensor of that length to write into, e.g. this rank's slot of a gather buffer)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    device = torch.device(device if device is not None else "cuda")
    w, mu, cov = _dev(W, device), _dev(MU, device), _dev(COV, device)
    G, K = int(w.shape[0]), int(w.shape[1])
    total = G * (G - 1) // 2
    pair_end = total if pair_end is None else int(pair_end)
    n_out = max(pair_end - pair_begin, 0)
    if out is None:
        out = torch.empty((n_out,), dtype=torch.float64, device=device)
    elif out.numel() != n_out or out.dtype != torch.float64 or not out.is_contiguous() or not out.is_cuda:
        raise ValueError("out must be a contiguous float64 tensor of %d entries on %s" % (n_out, device))
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
    diagonal, index = columns = ``list(gmm_dict.keys())``.  The warp LSAP keeps one or two cost columns per
    lane, so mixtures are limited to K <= 64 components (the reference default is 20)."""
    genes, W, MU, COV = stack_gmms(gmm_dict)
    G = len(genes)
    if G == 0:
        return pd.DataFrame(np.zeros((0, 0)), index=genes, columns=genes)
    from ..sharding import gmm_pairs_sharded      # pair slices per rank + all-gather under a process group
    torch = _lib.require_cuda()
    pairs = gmm_pairs_sharded(W, MU, COV, device=device)
    pairs = pairs.to(torch.device(device if device is not None else "cuda"))
    square = pairs_to_square(pairs, G).cpu().numpy()
    _raise_on_invalid(square, genes)
    return pd.DataFrame(square, index=genes, columns=genes)


def _raise_on_invalid(square, genes, query=None):
    """scipy's ``linear_sum_assignment`` raises on a cost matrix with NaN/inf entries (degenerate covariances);
    the device LSAP returns NaN for such a pair — surface it the same way instead of storing it (algorithm.py:24)."""
    bad = ~np.isfinite(square)
    if bad.any():
        idx = np.argwhere(bad)[0]
        a = genes[idx[0]] if query is None else query
        b = genes[idx[-1]]
        raise ValueError("matrix contains invalid numeric entries (GMM distance of genes %s and %s)" % (a, b))


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
    d = d.cpu().numpy()
    _raise_on_invalid(d, genes, query="<query>")
    df = pd.DataFrame.from_dict(dict(zip(genes, d)), orient="index")
    return df.sort_values(0)


def distribution_distance(first_gmm, second_gmm, device=None) -> float:
    """Drop-in for ``distribution_distance`` (distance.py:13-36) — a batch of one pair."""
    _, W, MU, COV = stack_gmms({"a": first_gmm, "b": second_gmm})
    return float(gmm_pair_distances(W, MU, COV, device=device).cpu().numpy()[0])


# ---------------------------------------------------------------------------------------------------
# Spot-level exact EMD (distance.py:179-202 of the reference)
# ---------------------------------------------------------------------------------------------------
def csr_to_support(img):
    """The staging of ``calculate_ot_distance`` (distance.py:192-199) for one sparse (x, y) image:
    supports = coordinates of the stored positive entries (CSR order), masses = ``data[data > 0]`` (the
    kernel divides by their sum, which equals ``data.sum()`` for the non-negative images the reference builds)."""
    coo = img.tocsr().tocoo()
    keep = coo.data > 0
    return (coo.row[keep].astype(np.int32), coo.col[keep].astype(np.int32), coo.data[keep].astype(np.float64))


def _emd_size_classes(lib, sx, sy, sizes, ptr, tx, ty):
    """Masks over the targets of one source: (int32 kernel, 64-bit kernel, global-tree kernel); the rest keep status 3.
    The default kernel keeps int32 potentials: it needs 2 * (span^2 + 1) * nodes < 2^31 for the bounding box of source and
    target together; problems beyond that (scattered coordinates, long thin tissues) go to the 64-bit variant.  Beyond one
    CTA's shared memory — e.g. a Stereo-seq bin50 slide with 40,000 source bins — the tree lives in the CTA's global
    scratch."""
    G = len(sizes)
    nz = sizes > 0
    first = ptr[:-1][nz]
    ext = {}
    for name, a, s0 in (("x", tx, sx), ("y", ty, sy)):
        lo = np.full(G, int(s0.min()), dtype=np.int64); hi = np.full(G, int(s0.max()), dtype=np.int64)
        if nz.any():
            lo[nz] = np.minimum(lo[nz], np.minimum.reduceat(a, first)); hi[nz] = np.maximum(hi[nz], np.maximum.reduceat(a, first))
        ext[name] = hi - lo
    maxc = ext["x"] ** 2 + ext["y"] ** 2
    narrow = 2 * (maxc + 1) * (len(sx) + sizes + 1) + maxc < 2 ** 31 - 1
    nodes = len(sx) + sizes + 1
    cap_n, cap_w = int(lib.stm_emd_smem_node_capacity(0)), int(lib.stm_emd_smem_node_capacity(1))
    cap_g = int(lib.stm_emd_global_node_capacity())
    arcs_ok = len(sx) * sizes < 2 ** 31 - 1
    small_n = narrow & (nodes <= cap_n)
    small_w = ~narrow & (nodes <= cap_w)
    large = ~small_n & ~small_w & (nodes <= cap_g) & arcs_ok
    return small_n, small_w, large


def _emd_pack(src, targets):
    sx, sy, sm = (np.ascontiguousarray(src[0], dtype=np.int32), np.ascontiguousarray(src[1], dtype=np.int32),
                  np.ascontiguousarray(src[2], dtype=np.float64))
    G = len(targets)
    sizes = np.array([len(t[0]) for t in targets], dtype=np.int64)
    ptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    cat = lambda k, dt: (np.concatenate([np.asarray(t[k], dtype=dt) for t in targets]) if G else np.zeros(0, dtype=dt))
    return sx, sy, sm, sizes, ptr, cat(0, np.int32), cat(1, np.int32), cat(2, np.float64)


def emd_pricing_threads(src, targets) -> np.ndarray:
    """Per target: threads per problem of the candidate-list pricing ``emd_batched(src, targets)`` solves it with
    (512 or 256; 0 = plain block search: the 64-bit kernel; -1 = the global-tree kernel, whose blocks grow over fruitless
    scans) — what the CPU oracle takes as ``cand_threads`` to follow the same pivots."""
    _lib.require_cuda()
    lib = _lib.load()
    sx, sy, sm, sizes, ptr, tx, ty, tm = _emd_pack(src, targets)
    out = np.zeros(len(targets), dtype=np.int64)
    if len(targets) == 0 or len(sx) == 0:
        return out
    small_n, _, large = _emd_size_classes(lib, sx, sy, sizes, ptr, tx, ty)
    if small_n.any():
        out[small_n] = int(lib.stm_emd_pricing_threads(len(sx), int(sizes[small_n].max()), 0))
    out[large] = -1
    return out


def emd_batched(src, targets, max_iter: int = 200000, block_size: int = 0, device=None, return_iters: bool = False,
                star_start: bool = False, candidates: bool = True):
    """Exact EMD between one source support ``(x, y, mass)`` and every target support of ``targets``.

    Returns ``(cost[G] float64, status[G] int32)`` as numpy arrays (+ pivots with ``return_iters``).
    ``max_iter`` mirrors ``numItermax=200000`` of the reference (distance.py:201): a problem that hits it
    reports status 1 and the cost of its current basis, as POT does.  ``star_start`` selects the artificial star
    POT starts from instead of the multiscale start basis (same optimum, about ten times the pivots);
    ``candidates=False`` prices by plain block search, one pricing pass per pivot (same optimum, more passes).
    """
    torch = _lib.require_cuda()
    lib = _lib.load()
    device = torch.device(device if device is not None else "cuda")
    G = len(targets)
    sx, sy, sm, sizes, ptr, tx, ty, tm = _emd_pack(src, targets)
    if G == 0:
        out = (np.zeros(0), np.zeros(0, dtype=np.int32))
        return out + (np.zeros(0, dtype=np.int64),) if return_iters else out
    if len(sx) == 0:
        out = (np.zeros(G), np.full(G, 2, dtype=np.int32))
        return out + (np.zeros(G, dtype=np.int64),) if return_iters else out
    small_n, small_w, large = _emd_size_classes(lib, sx, sy, sizes, ptr, tx, ty)
    up = lambda a: torch.from_numpy(a).to(device)
    d = [up(a) for a in (sx, sy, sm, ptr, tx, ty, tm)]
    with torch.cuda.device(device):
        props = torch.cuda.get_device_properties(device)
        cost = torch.full((G,), float("nan"), dtype=torch.float64, device=device)
        status = torch.full((G,), 3, dtype=torch.int32, device=device)
        iters = torch.zeros(G, dtype=torch.int64, device=device)
        keep = []
        for wide, sel, big in ((False, np.flatnonzero(small_n), False), (True, np.flatnonzero(small_w), False),
                               (True, np.flatnonzero(large), True)):
            if len(sel) == 0:
                continue
            order = sel[np.argsort(-sizes[sel], kind="stable")].astype(np.int32)      # largest problems first
            d_order = up(order)
            call_max_tgt = int(sizes[sel].max())
            if big:
                n_ctas = int(min(len(sel), props.multi_processor_count))
            else:
                smem = (len(sx) + call_max_tgt + 10) * (22 if wide else 18) + 4 * 2048 * 2 + 64 + 1024
                per_sm = max(1, min(8, (227 * 1024) // smem))
                n_ctas = int(min(len(sel), props.multi_processor_count * per_sm))
            ws_bytes = int(lib.stm_emd_workspace_bytes(len(sx), call_max_tgt, n_ctas))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
            fn = lib.stm_emd_spots_batched_wide if wide else lib.stm_emd_spots_batched
            st = fn(_lib.ptr(d[0]), _lib.ptr(d[1]), _lib.ptr(d[2]), len(sx), _lib.ptr(d[3]),
                    _lib.ptr(d[4]), _lib.ptr(d[5]), _lib.ptr(d[6]), len(sel), call_max_tgt, _lib.ptr(d_order),
                    int(max_iter), int(block_size),
                    (_lib.STM_EMD_STAR_START if star_start else 0) | (0 if candidates else _lib.STM_EMD_NO_CANDIDATES),
                    _lib.ptr(ws), ws_bytes, n_ctas,
                    _lib.ptr(cost), _lib.ptr(status), _lib.ptr(iters), _lib.current_stream_ptr())
            _lib.check(st, "stm_emd_spots_batched_wide" if wide else "stm_emd_spots_batched")
            keep.append((d_order, ws))
    out = (cost.cpu().numpy(), status.cpu().numpy())
    return out + (iters.cpu().numpy(),) if return_iters else out


def ot_distances_to_source(source, targets, max_iter: int = 200000, device=None):
    """EMD from one sparse image to each image of ``targets`` (the loop of SPFinder.py:285-299)."""
    from ..sharding import emd_sharded            # targets dealt over the ranks under a process group
    return emd_sharded(csr_to_support(source), [csr_to_support(t) for t in targets], max_iter=max_iter, device=device)


def calculate_ot_distance(source, target, device=None) -> float:
    """Drop-in for ``calculate_ot_distance`` (distance.py:191-202)."""
    cost, status = ot_distances_to_source(source, [target], device=device)
    if status[0] >= 2:
        raise ValueError("EMD failed with status %d (2 = empty image, 3 = problem too large for the on-chip solver)"
                         % int(status[0]))
    return float(cost[0])


def build_ot_distance_array(csr_dict, gene_list=None, device=None):
    """Drop-in for ``build_ot_distance_array`` (distance.py:179-188): symmetric DataFrame of pairwise spot EMDs;
    row i is one batched launch (source = gene i, targets = genes j > i)."""
    genes = list(gene_list) if gene_list is not None else list(csr_dict.keys())
    n = len(genes)
    sup = [csr_to_support(csr_dict[g]) for g in genes]
    D = np.zeros((n, n), dtype=np.float64)
    for i in range(n - 1):
        cost, status = emd_batched(sup[i], sup[i + 1:], device=device)
        if (status >= 2).any():
            raise ValueError("EMD failed for a pair involving gene %s" % genes[i])
        D[i, i + 1:] = cost
        D[i + 1:, i] = cost
    return pd.DataFrame(D, index=genes, columns=genes)
