"""Multi-GPU check (NOT collected by pytest): run under torchrun with one process per GPU, e.g.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py

Every rank drives the SPFinder façade with the process group initialised (genes / pair slices / EMD targets are
sharded over the ranks and all-gathered) and compares the result with the same stages computed locally, unsharded.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    try:
        from stminer_b200.SPFinder import SPFinder
        from stminer_b200.Simulate import simulate_adata
        from stminer_b200.Algorithm.distance import emd_batched, gmm_pair_distances, stack_gmms, csr_to_support
        from stminer_b200.Algorithm.distribution import fit_spot_matrix_host, result_to_gmm_dict
        from stminer_b200.adata_lite import spot_matrix_from_adata
        adata = simulate_adata("T", seed=3, n_templates=6, replicas=5)
        sp = SPFinder(adata)
        sp.get_genes_csr_array(min_cells=20, normalize=False)
        sp.spatial_high_variable_genes()
        genes = list(sp.global_distance["Gene"])[:24]
        sp.fit_pattern(n_comp=6, gene_list=genes, min_cells=1)
        sp.build_distance_array()
        # ---- the same stages without sharding, on this rank ----
        sm = spot_matrix_from_adata(sp.adata, list(sp.patterns.keys()))
        host = fit_spot_matrix_host(sm, 6, init=None, seed=sp.seed)
        ref = result_to_gmm_dict(sm.genes, host)
        for g in ref:
            assert np.array_equal(ref[g].means_, sp.patterns[g].means_), g
            assert np.array_equal(ref[g].covariances_, sp.patterns[g].covariances_), g
        _, W, MU, COV = stack_gmms(sp.patterns)
        pairs = gmm_pair_distances(W, MU, COV).cpu().numpy()
        G = len(sp.patterns)
        iu = np.triu_indices(G, 1)
        assert np.array_equal(np.asarray(sp.genes_distance_array)[iu], pairs)
        names = list(sp.csr_dict.keys())
        from scipy.sparse import csr_matrix
        data = np.array(sp.adata.X.sum(axis=1)).flatten().astype(np.float64)
        data[data > np.percentile(data, 100.0)] = np.percentile(data, 100.0)
        gm = csr_matrix((data, (np.array(sp.adata.obs["x"].values).flatten(), np.array(sp.adata.obs["y"].values).flatten())))
        cost, status = emd_batched(csr_to_support(gm), [csr_to_support(sp.csr_dict[k]) for k in names])
        gd = sp.global_distance.set_index("Gene")["Distance"]
        assert (status == 0).all() and np.array_equal(gd.loc[names].to_numpy(), cost)
        t = torch.tensor([1.0], device="cuda")
        dist.all_reduce(t)
        assert int(t.item()) == world
        print("rank %d/%d: sharded SPFinder stages match the unsharded ones (%d genes, %d pairs, %d EMDs)"
              % (rank, world, G, len(pairs), len(names)), flush=True)
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
