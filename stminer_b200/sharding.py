"""Multi-GPU sharding of the hot path (one process per GPU, ``torch.distributed``).

The path has no exchange step inside the computation: genes are independent
(distribution.py:184) and so are GMM pairs (distance.py:94-98).  Ranks therefore work on disjoint
shards with no data-path collective; one all-gather per stage makes the (tiny) results replicated —
fitted parameters so every rank can form pairs, pair distances so every rank holds the matrix.
The collectives run on whatever backend the process group has (NCCL over NVLink on the GPU box,
gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def shard_genes_lpt(nnz: np.ndarray, world_size: int) -> List[np.ndarray]:
    """Deal genes to ranks longest-first onto the least-loaded rank (work ~ nnz).  Returns, per rank,
    the sorted gene indices it owns.  Deterministic, so every rank computes the same partition."""
    nnz = np.asarray(nnz, dtype=np.int64)
    order = np.argsort(-nnz, kind="stable")
    load = np.zeros(world_size, dtype=np.int64)
    owner = np.empty(len(nnz), dtype=np.int64)
    # snake deal in blocks of world_size keeps it O(G log G) and within one gene of optimal LPT
    for start in range(0, len(order), world_size):
        blk = order[start:start + world_size]
        ranks = np.argsort(load, kind="stable")[:len(blk)]
        owner[blk] = ranks
        load[ranks] += nnz[blk]
    return [np.sort(np.flatnonzero(owner == r)) for r in range(world_size)]


def pair_slice(total_pairs: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous slice [begin, end) of the linearised upper triangle owned by ``rank``."""
    base, rem = divmod(int(total_pairs), world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def allgather_rows(local, owned_idx: List[np.ndarray], group=None):
    """All-gather per-gene rows.  ``local`` is a tensor [n_local, ...] holding the rows of
    ``owned_idx[rank]``; returns the full tensor [G, ...] in original gene order on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n_max = max(len(ix) for ix in owned_idx)
    G = sum(len(ix) for ix in owned_idx)
    pad = torch.zeros((n_max,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:len(owned_idx[rank])] = local
    gathered = torch.empty((world * n_max,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, pad, group=group)
    out = torch.empty((G,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r in range(world):
        ix = torch.as_tensor(owned_idx[r], dtype=torch.long, device=local.device)
        out[ix] = gathered[r * n_max:r * n_max + len(owned_idx[r])]
    return out


def allgather_pairs(local, total_pairs: int, group=None):
    """All-gather the contiguous pair slices into the full pair vector (length ``total_pairs``)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n_max = -(-int(total_pairs) // world)
    pad = torch.zeros((n_max,), dtype=local.dtype, device=local.device)
    pad[:local.numel()] = local
    gathered = torch.empty((world * n_max,), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, pad, group=group)
    parts = []
    for r in range(world):
        b, e = pair_slice(total_pairs, r, world)
        parts.append(gathered[r * n_max:r * n_max + (e - b)])
    return torch.cat(parts)


# ------------------------------------------------------------------------------------------------------------
# Stage drivers: what the host API calls when a process group exists (one process per GPU).  `compute(...)`
# arguments let the world-size-2 gloo tests on CPU exercise the sharding / gathering with a stand-in for the kernels.
# ------------------------------------------------------------------------------------------------------------
def group_rank_world(group=None) -> Tuple[int, int]:
    """(rank, world_size) of the process group, (0, 1) when torch.distributed is not initialised."""
    try:
        import torch.distributed as dist
    except ImportError:      # pragma: no cover
        return 0, 1
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def _collective_device(group=None):
    """Tensors handed to the collectives live on the current CUDA device under NCCL, on the host under gloo."""
    import torch
    import torch.distributed as dist
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")


def fit_genes_sharded(sm, n_comp: int, init=None, group=None, compute=None, **kw) -> dict:
    """GMM fit of every gene of the SpotMatrix ``sm`` with the genes dealt LPT-wise over the ranks:
    each rank uploads and fits only its own columns, one all-gather per output array replicates the result.
    Returns the same host dict as ``fit_spot_matrix_host``, in ``sm`` gene order, on every rank."""
    import torch
    rank, world = group_rank_world(group)
    if compute is None:
        from .Algorithm.distribution import fit_spot_matrix_host as compute
    if world == 1:
        return compute(sm, n_comp, init=init, **kw)
    owned = shard_genes_lpt(sm.nnz_per_gene(), world)
    mine = owned[rank]
    sub_init = None if init is None else tuple(np.asarray(a)[mine] for a in init)
    local = compute(sm.select(mine), n_comp, init=sub_init, **kw)
    dev = _collective_device(group)
    out = {}
    for k, v in local.items():
        out[k] = allgather_rows(torch.from_numpy(np.ascontiguousarray(v)).to(dev), owned, group=group).cpu().numpy()
    return out


def gmm_pairs_sharded(W, MU, COV, group=None, device=None, compute=None):
    """All pair distances of G mixtures with the linearised upper triangle cut into one contiguous slice per rank;
    returns the full pair vector (tensor on the collective's device) on every rank."""
    import torch
    rank, world = group_rank_world(group)
    if compute is None:
        from .Algorithm.distance import gmm_pair_distances as compute
    G = int(W.shape[0])
    total = G * (G - 1) // 2
    if world == 1:
        return compute(W, MU, COV, 0, total, device=device)
    b, e = pair_slice(total, rank, world)
    local = compute(W, MU, COV, b, e, device=device)
    return allgather_pairs(local.to(_collective_device(group)), total, group=group)


def emd_sharded(src, targets, group=None, compute=None, **kw):
    """Spot-level EMD of one source against every target, targets dealt LPT-wise by support size; (cost, status)
    numpy arrays in target order on every rank."""
    import torch
    rank, world = group_rank_world(group)
    if compute is None:
        from .Algorithm.distance import emd_batched as compute
    if world == 1:
        return compute(src, targets, **kw)
    sizes = np.array([len(t[0]) for t in targets], dtype=np.int64)
    owned = shard_genes_lpt(sizes, world)
    cost, status = compute(src, [targets[i] for i in owned[rank]], **kw)[:2]
    dev = _collective_device(group)
    cost = allgather_rows(torch.from_numpy(np.ascontiguousarray(cost, dtype=np.float64)).to(dev), owned, group=group)
    status = allgather_rows(torch.from_numpy(np.ascontiguousarray(status, dtype=np.int32)).to(dev), owned, group=group)
    return cost.cpu().numpy(), status.cpu().numpy()
