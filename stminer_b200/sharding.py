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
