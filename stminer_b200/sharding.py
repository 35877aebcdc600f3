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
    """Contiguous slice [begin, end) of the linearised upper triangle owned by ``rank``: every rank owns
    ceil(total / world) pairs except the last ones, which own what is left — so the ranks' slices, laid out at a fixed
    stride, ARE the pair vector and the all-gather needs neither padding nor a re-assembly copy."""
    n_max = -(-int(total_pairs) // world_size) if total_pairs > 0 else 0
    begin = min(int(total_pairs), rank * n_max)
    return begin, min(int(total_pairs), begin + n_max)


def shard_genes_contiguous(col_ptr: np.ndarray, world_size: int) -> np.ndarray:
    """Gene boundaries [world_size + 1] of contiguous gene ranges with (nearly) equal numbers of stored values: rank r
    owns genes [b[r], b[r + 1]).  Contiguous ranges are views of the CSC matrix (no host gather before the upload) and
    their results are contiguous rows of the gathered arrays; with thousands of genes per rank the work (stored values x
    EM iterations) balances as well as a longest-first deal."""
    col_ptr = np.asarray(col_ptr, dtype=np.int64)
    G = len(col_ptr) - 1
    targets = np.linspace(0, int(col_ptr[-1]), world_size + 1)
    b = np.searchsorted(col_ptr, targets, side="left").astype(np.int64)
    b[0], b[-1] = 0, G
    return np.maximum.accumulate(np.minimum(b, G))


def shard_gene_chunks(col_ptr: np.ndarray, world_size: int, chunk_genes: int = 16) -> List[List[Tuple[int, int]]]:
    """Strong-scaling partition of the genes: runs of ``chunk_genes`` consecutive genes are dealt longest-first (by stored
    values) onto the least-loaded rank.  Chunks keep the host slices contiguous views (each is one DMA from the page-locked
    matrix, no host gather), the deal balances the stored values to within one chunk, and consecutive chunks land on
    different ranks — neighbouring genes (replicates of one pattern, genes of one family) tend to need similar numbers of
    EM iterations, which a cut into a few long contiguous ranges would pile onto one rank.  Returns, per rank, its
    (g0, g1) ranges in increasing order; deterministic, so every rank computes the same partition."""
    col_ptr = np.asarray(col_ptr, dtype=np.int64)
    G = len(col_ptr) - 1
    starts = np.arange(0, G, max(1, int(chunk_genes)), dtype=np.int64)
    ends = np.minimum(starts + max(1, int(chunk_genes)), G)
    nnz = col_ptr[ends] - col_ptr[starts]
    import heapq
    order = np.argsort(-nnz, kind="stable")
    heap = [(0, 0, r) for r in range(world_size)]          # (stored values, chunks, rank): least loaded first
    owner = np.empty(len(starts), dtype=np.int64)
    for c in order.tolist():
        load, count, r = heapq.heappop(heap)
        owner[c] = r
        heapq.heappush(heap, (load + int(nnz[c]), count + 1, r))
    return [[(int(starts[c]), int(ends[c])) for c in np.flatnonzero(owner == r)] for r in range(world_size)]


def ranges_to_index(ranges: List[Tuple[int, int]]) -> np.ndarray:
    return (np.concatenate([np.arange(a, b, dtype=np.int64) for a, b in ranges]) if ranges else np.zeros(0, dtype=np.int64))


def allgather_rows(local, owned_idx: List[np.ndarray], group=None):
    """All-gather per-gene rows.  ``local`` is a tensor [n_local, ...] holding the rows of
    ``owned_idx[rank]``; returns the full tensor [G, ...] in original gene order on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n_max = max(len(ix) for ix in owned_idx)
    G = sum(len(ix) for ix in owned_idx)
    gathered = torch.empty((world * n_max,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    mine = gathered[rank * n_max:(rank + 1) * n_max]
    mine[:len(owned_idx[rank])] = local
    mine[len(owned_idx[rank]):] = 0
    dist.all_gather_into_tensor(gathered, mine, group=group)           # in place: the input is this rank's slot
    out = torch.empty((G,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r in range(world):
        ix = torch.as_tensor(owned_idx[r], dtype=torch.long, device=local.device)
        out[ix] = gathered[r * n_max:r * n_max + len(owned_idx[r])]
    return out


def pair_gather_buffer(total_pairs: int, world_size: int, dtype, device):
    """Buffer of world x ceil(total / world) entries whose first ``total_pairs`` entries are the pair vector once every
    rank has written its slice at ``pair_slice`` and ``allgather_pairs_inplace`` has run."""
    import torch
    n_max = -(-int(total_pairs) // world_size) if total_pairs > 0 else 0
    return torch.empty((world_size * n_max,), dtype=dtype, device=device)


def allgather_pairs_inplace(buf, total_pairs: int, group=None):
    """In-place all-gather of the slices the ranks wrote into ``buf`` (see ``pair_gather_buffer``); returns the view
    ``buf[:total_pairs]`` — no padding copy, no concatenation."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n_max = buf.numel() // world
    dist.all_gather_into_tensor(buf, buf[rank * n_max:(rank + 1) * n_max], group=group)
    return buf[:int(total_pairs)]


def allgather_pairs(local, total_pairs: int, group=None):
    """All-gather the contiguous pair slices into the full pair vector (length ``total_pairs``); ``local`` is this rank's
    slice.  (One copy of the local slice into the gather buffer; ``gmm_pairs_sharded`` writes there directly.)"""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    buf = pair_gather_buffer(total_pairs, world, local.dtype, local.device)
    b, e = pair_slice(total_pairs, rank, world)
    buf[b:e] = local
    return allgather_pairs_inplace(buf, total_pairs, group=group)


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


FIT_FIELDS = ("weights", "means", "covariances", "n_iter", "converged", "lower_bound", "status")


def pack_fit_result(res, K: int):
    """One [n, 7K + 4] float64 tensor holding every output array of a fit (device ``FitResult`` or host dict): the
    gather of a stage is ONE collective instead of seven.  The int32 fields are exact in float64."""
    import torch
    K = int(K)
    width = dict(weights=K, means=2 * K, covariances=4 * K, n_iter=1, converged=1, lower_bound=1, status=1)
    if isinstance(res, dict):
        n = len(res["status"])
        cols = [torch.from_numpy(np.ascontiguousarray(res[k])).to(torch.float64).reshape(n, width[k]) for k in FIT_FIELDS]
    else:
        n = res.status.shape[0]
        cols = [getattr(res, k).to(torch.float64).reshape(n, width[k]) for k in FIT_FIELDS]
    return torch.cat(cols, dim=1)


def unpack_fit_rows(rows: np.ndarray, K: int) -> dict:
    """Inverse of ``pack_fit_result`` on a host array [G, 7K + 4]; every field is a fresh array (``rows`` may be a
    buffer that the next call reuses)."""
    G = rows.shape[0]
    return dict(weights=rows[:, :K].copy(), means=rows[:, K:3 * K].copy().reshape(G, K, 2),
                covariances=rows[:, 3 * K:7 * K].copy().reshape(G, K, 2, 2),
                n_iter=rows[:, 7 * K].astype(np.int32), converged=rows[:, 7 * K + 1].astype(np.int32),
                lower_bound=rows[:, 7 * K + 2].copy(), status=rows[:, 7 * K + 3].astype(np.int32))


_PINNED = {}


def _pinned_like(t):
    """A cached page-locked host tensor of the shape / dtype of the CUDA tensor ``t`` (the read-back target of a stage)."""
    import torch
    key = (tuple(t.shape), t.dtype)
    if key not in _PINNED:
        _PINNED.clear()                      # one live shape at a time is enough for the stage drivers
        _PINNED[key] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    return _PINNED[key]


def gather_fit_rows(packed, owned_idx: List[np.ndarray], group=None):
    """All-gather of the ranks' packed rows (rank r holds the genes ``owned_idx[r]``, in that order) — one in-place
    collective on a buffer with a fixed stride per rank, one read-back; returns the host array [G, C] in gene order."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lens = [len(ix) for ix in owned_idx]
    n_max = max(lens)
    G = int(sum(lens))
    buf = torch.empty((world * n_max, packed.shape[1]), dtype=packed.dtype, device=packed.device)
    buf[rank * n_max:rank * n_max + packed.shape[0]] = packed
    dist.all_gather_into_tensor(buf, buf[rank * n_max:(rank + 1) * n_max], group=group)
    if buf.is_cuda:
        # gene order restored on the device (one indexed copy), then ONE read-back into page-locked memory
        src = np.concatenate([r * n_max + np.arange(lens[r], dtype=np.int64) for r in range(world)])
        dst = np.concatenate(owned_idx)
        perm = np.empty(G, dtype=np.int64)
        perm[dst] = src
        ordered = buf.index_select(0, torch.from_numpy(perm).to(buf.device, non_blocking=True))
        h = _pinned_like(ordered)
        h.copy_(ordered, non_blocking=True)
        torch.cuda.current_stream(buf.device).synchronize()
        return h.numpy()        # a view of the cached page-locked buffer: unpack_fit_rows copies every field out of it
    h = buf.numpy()
    out = np.empty((G, packed.shape[1]), dtype=h.dtype)
    for r in range(world):
        out[owned_idx[r]] = h[r * n_max:r * n_max + lens[r]]
    return out


def default_chunk_genes(n_genes: int, world_size: int) -> int:
    """Chunk length of ``shard_gene_chunks`` for a sharded fit: 16 genes, longer when that would give a rank more than
    ~64 chunks (every chunk is a pair of host-to-device copies)."""
    return max(16, -(-int(n_genes) // (64 * max(1, world_size))))


def fit_genes_sharded(sm, n_comp: int, init=None, group=None, compute=None, chunk_genes: int = 0, **kw) -> dict:
    """GMM fit of every gene of the SpotMatrix ``sm`` with the genes dealt over the ranks in chunks of consecutive
    genes (``shard_gene_chunks``): each rank uploads only its own columns — every chunk one DMA from the
    host matrix into its place in the rank's device matrix, no host gather, in three groups so that the fit of a group
    runs while the next one is still arriving — fits them, packs the fitted arrays into one device buffer, ONE
    all-gather replicates them and one read-back returns them.
    Returns the same host dict as ``fit_spot_matrix_host``, in ``sm`` gene order, on every rank."""
    rank, world = group_rank_world(group)
    if world == 1:
        if compute is None:
            from .Algorithm.distribution import fit_spot_matrix_host as compute
        return compute(sm, n_comp, init=init, **kw)
    parts = shard_gene_chunks(sm.col_ptr, world, chunk_genes or default_chunk_genes(sm.n_genes, world))
    owned = [ranges_to_index(p) for p in parts]
    mine = owned[rank]
    sub_init = None if init is None else tuple(np.asarray(a)[mine] for a in init)
    if compute is not None:
        local = compute(sm.select(mine), n_comp, init=sub_init, **kw)
        packed = pack_fit_result(local, n_comp).to(_collective_device(group))
    else:
        import torch
        from .Algorithm.distribution import _copy_stream, fit_spot_matrix
        from .staging import DeviceSpotMatrix
        dev = torch.device(kw.pop("device", None) or "cuda")
        kw.pop("n_chunks", None)
        smc = sm.compact()
        ranges = parts[rank]
        # groups of ranges holding ~15 / 35 / 50 % of the rank's stored values: the first fit starts early
        sizes = np.array([smc.col_ptr[b] - smc.col_ptr[a] for a, b in ranges], dtype=np.float64)
        cum = np.cumsum(sizes) / max(1.0, sizes.sum())
        cuts = sorted(set([0] + [int(np.searchsorted(cum, f, side="left")) + 1 for f in (0.15, 0.5)] + [len(ranges)]))
        cuts = [min(c, len(ranges)) for c in cuts]
        with torch.cuda.device(dev):
            compute_stream = torch.cuda.current_stream()
            copy = _copy_stream(dev)
            copy.wait_stream(compute_stream)
            packs, keep, at = [], [], 0
            for c0, c1 in zip(cuts[:-1], cuts[1:]):
                if c1 <= c0:
                    continue
                grp = ranges[c0:c1]
                n_grp = int(sum(b - a for a, b in grp))
                with torch.cuda.stream(copy):
                    dsm = DeviceSpotMatrix.from_ranges(smc, grp, device=dev)
                    ready = torch.cuda.Event()
                    ready.record(copy)
                compute_stream.wait_event(ready)
                g_init = None if sub_init is None else tuple(a[at:at + n_grp] for a in sub_init)
                res = fit_spot_matrix(dsm, n_comp, init=g_init, **kw)
                packs.append(pack_fit_result(res, n_comp))
                keep.append((dsm, res))
                at += n_grp
            packed = torch.cat(packs, dim=0) if packs else torch.zeros((0, 7 * int(n_comp) + 4), dtype=torch.float64, device=dev)
            out = unpack_fit_rows(gather_fit_rows(packed, owned, group=group), int(n_comp))
            copy.wait_stream(compute_stream)
        return out
    return unpack_fit_rows(gather_fit_rows(packed, owned, group=group), int(n_comp))


def gmm_pairs_sharded(W, MU, COV, group=None, device=None, compute=None):
    """All pair distances of G mixtures with the linearised upper triangle cut into one contiguous slice per rank;
    every rank computes its slice straight into its slot of the gather buffer and one in-place all-gather makes the
    pair vector (tensor on the collective's device) complete on every rank."""
    import torch
    rank, world = group_rank_world(group)
    own_kernel = compute is None
    if compute is None:
        from .Algorithm.distance import gmm_pair_distances as compute
    G = int(W.shape[0])
    total = G * (G - 1) // 2
    if world == 1:
        return compute(W, MU, COV, 0, total, device=device)
    b, e = pair_slice(total, rank, world)
    cdev = _collective_device(group)
    if own_kernel and cdev.type == "cuda":
        buf = pair_gather_buffer(total, world, torch.float64, cdev)
        compute(W, MU, COV, b, e, device=cdev, out=buf[b:e])
        return allgather_pairs_inplace(buf, total, group=group)
    local = compute(W, MU, COV, b, e, device=device)
    return allgather_pairs(local.to(cdev), total, group=group)


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
