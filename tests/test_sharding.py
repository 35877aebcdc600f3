"""CPU: multi-GPU sharding logic and the all-gather plumbing on a world_size-2 gloo group."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stminer_b200.sharding import (allgather_pairs, allgather_rows, pair_slice, ranges_to_index, shard_gene_chunks,
                                   shard_genes_contiguous, shard_genes_lpt)


def test_shard_genes_lpt_balances_and_partitions():
    rng = np.random.default_rng(0)
    nnz = (rng.pareto(1.5, 1000) * 1000 + 100).astype(np.int64)
    for world in (1, 2, 4, 8):
        parts = shard_genes_lpt(nnz, world)
        allg = np.sort(np.concatenate(parts))
        assert np.array_equal(allg, np.arange(1000))
        loads = np.array([nnz[p].sum() for p in parts])
        assert loads.max() <= loads.mean() * 1.05 + nnz.max()


def test_shard_genes_contiguous_balances_stored_values():
    rng = np.random.default_rng(1)
    nnz = (rng.pareto(1.5, 5000) * 1000 + 100).astype(np.int64)
    col_ptr = np.concatenate([[0], np.cumsum(nnz)])
    for world in (1, 2, 4, 8):
        b = shard_genes_contiguous(col_ptr, world)
        assert b[0] == 0 and b[-1] == 5000 and (np.diff(b) >= 0).all() and len(b) == world + 1
        loads = np.diff(col_ptr[b])
        assert loads.max() <= loads.mean() + nnz.max()
    assert shard_genes_contiguous(np.array([0, 3, 5]), 8)[-1] == 2            # more ranks than genes: empty ranges


def test_shard_gene_chunks_partitions_and_balances():
    rng = np.random.default_rng(2)
    nnz = np.repeat((rng.pareto(1.5, 100) * 1000 + 100).astype(np.int64), 50)        # runs of 50 equal genes ("replicates")
    col_ptr = np.concatenate([[0], np.cumsum(nnz)])
    for world in (1, 2, 4, 8):
        parts = shard_gene_chunks(col_ptr, world, chunk_genes=16)
        idx = [ranges_to_index(p) for p in parts]
        assert np.array_equal(np.sort(np.concatenate(idx)), np.arange(5000))
        loads = np.array([nnz[i].sum() for i in idx])
        assert loads.max() <= loads.mean() + 16 * nnz.max()
        if world == 8:        # a run of 50 replicates is spread over several ranks, not piled onto one
            owner = np.empty(5000, dtype=int)
            for r, i in enumerate(idx):
                owner[i] = r
            assert np.mean([len(set(owner[k:k + 50])) for k in range(0, 5000, 50)]) >= 3


def test_pair_slice_covers_triangle():
    for total, world in ((7998000, 8), (10, 3), (0, 4), (5, 8)):
        cuts = [pair_slice(total, r, world) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == total
        assert all(cuts[r][1] == cuts[r + 1][0] for r in range(world - 1))
        sizes = [e - b for b, e in cuts]
        n_max = -(-total // world) if total else 0          # fixed stride: all slices full except the trailing ones
        assert all(b == min(total, r * n_max) for r, (b, e) in enumerate(cuts)) and max(sizes, default=0) <= n_max


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


def _stage_worker(rank, world, port):
    """The stage drivers with stand-ins for the kernels: sharded result == unsharded result on every rank."""
    from stminer_b200.sharding import emd_sharded, fit_genes_sharded, gmm_pairs_sharded
    from stminer_b200.Simulate import simulate_spot_matrix
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sm = simulate_spot_matrix("T", seed=1, n_templates=3, replicas=3)
        K = 3

        def fake_fit(m, n_comp, init=None, **kw):       # per-gene values that depend only on the gene's own data
            nnz = m.nnz_per_gene().astype(np.float64)
            tot = np.array([m.val[m.col_ptr[g]:m.col_ptr[g + 1]].sum() for g in range(m.n_genes)], dtype=np.float64)
            extra = 0.0 if init is None else np.asarray(init[0]).sum(axis=1)
            G_ = m.n_genes
            return {"weights": np.repeat((nnz + extra)[:, None], n_comp, axis=1), "means": np.tile(tot[:, None, None], (1, n_comp, 2)),
                    "covariances": np.tile((nnz * 0.5)[:, None, None, None], (1, n_comp, 2, 2)), "n_iter": (nnz % 7).astype(np.int32),
                    "converged": (nnz % 2).astype(np.int32), "lower_bound": -tot / 3.0, "status": (nnz > 0).astype(np.int32)}
        init = (np.arange(sm.n_genes * K, dtype=np.float64).reshape(sm.n_genes, K),)
        for ini in (None, init):
            a = fit_genes_sharded(sm, K, init=ini, compute=fake_fit)
            b = fake_fit(sm, K, init=ini)
            for k in b:
                assert np.array_equal(a[k], b[k]), k

        G = 11
        W = torch.arange(G * K, dtype=torch.float64).reshape(G, K)

        def fake_pairs(W, MU, COV, b, e, device=None):
            return torch.arange(b, e, dtype=torch.float64) * 3.0 + W[0, 0]
        full = gmm_pairs_sharded(W, None, None, compute=fake_pairs)
        assert torch.equal(full, torch.arange(G * (G - 1) // 2, dtype=torch.float64) * 3.0)

        targets = [(np.zeros(n), np.zeros(n), np.full(n, float(n))) for n in (5, 1, 9, 3, 7, 2, 8)]

        def fake_emd(src, tg, **kw):
            return (np.array([t[2].sum() for t in tg], dtype=np.float64), np.array([len(t[0]) % 3 for t in tg], dtype=np.int32))
        cost, status = emd_sharded(None, targets, compute=fake_emd)
        c0, s0 = fake_emd(None, targets)
        assert np.array_equal(cost, c0) and np.array_equal(status, s0)
    finally:
        dist.destroy_process_group()


def test_stage_drivers_world_size_2_gloo():
    port = 31500 + os.getpid() % 2000
    mp.spawn(_stage_worker, args=(2, port), nprocs=2, join=True)


def test_stage_drivers_without_process_group():
    from stminer_b200.sharding import group_rank_world
    assert group_rank_world() == (0, 1)
