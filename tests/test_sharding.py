"""CPU: multi-GPU sharding logic and the all-gather plumbing on a world_size-2 gloo group."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stminer_b200.sharding import allgather_pairs, allgather_rows, pair_slice, shard_genes_lpt


def test_shard_genes_lpt_balances_and_partitions():
    rng = np.random.default_rng(0)
    nnz = (rng.pareto(1.5, 1000) * 1000 + 100).astype(np.int64)
    for world in (1, 2, 4, 8):
        parts = shard_genes_lpt(nnz, world)
        allg = np.sort(np.concatenate(parts))
        assert np.array_equal(allg, np.arange(1000))
        loads = np.array([nnz[p].sum() for p in parts])
        assert loads.max() <= loads.mean() * 1.05 + nnz.max()


def test_pair_slice_covers_triangle():
    for total, world in ((7998000, 8), (10, 3), (0, 4), (5, 8)):
        cuts = [pair_slice(total, r, world) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == total
        assert all(cuts[r][1] == cuts[r + 1][0] for r in range(world - 1))
        sizes = [e - b for b, e in cuts]
        assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, G, total_pairs):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nnz = (np.arange(G) * 37 % 101 + 5).astype(np.int64)
        parts = shard_genes_lpt(nnz, world)
        # every rank "fits" its genes: row g holds f(g); the gather must restore gene order on all ranks
        mine = torch.tensor(parts[rank], dtype=torch.float64)
        local = torch.stack([mine * 2.0 + 1.0, mine ** 2], dim=1).reshape(len(mine), 2, 1)
        full = allgather_rows(local, parts)
        g = torch.arange(G, dtype=torch.float64)
        assert torch.equal(full[:, 0, 0], g * 2.0 + 1.0) and torch.equal(full[:, 1, 0], g ** 2)
        b, e = pair_slice(total_pairs, rank, world)
        pairs = allgather_pairs(torch.arange(b, e, dtype=torch.float64) * 0.5, total_pairs)
        assert torch.equal(pairs, torch.arange(total_pairs, dtype=torch.float64) * 0.5)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("G,total_pairs", [(23, 253), (8, 11)])
def test_allgather_world_size_2_gloo(G, total_pairs):
    port = 29500 + (os.getpid() + G) % 2000
    mp.spawn(_worker, args=(2, port, G, total_pairs), nprocs=2, join=True)
