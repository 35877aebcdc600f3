"""The reference's canonical workflow (readme / tests/test_STMiner.py) end to end on synthetic Visium-shaped data,
timed per stage: staging -> SVG ranking (exact EMD per gene) -> mixture fit of the top genes -> distance matrix ->
MDS + KMeans -> pattern voting."""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from stminer_b200.Simulate import simulate_adata
from stminer_b200.SPFinder import SPFinder

n_top = int(sys.argv[1]) if len(sys.argv) > 1 else 500
t0 = time.perf_counter(); adata = simulate_adata("V", seed=0); t_gen = time.perf_counter() - t0
sp = SPFinder(adata)
def stage(name, fn):
    torch.cuda.synchronize(); t = time.perf_counter(); fn(); torch.cuda.synchronize()
    print("%-46s %8.3f s" % (name, time.perf_counter() - t), flush=True)
print("synthetic Visium data: %d spots x %d genes (generated in %.1f s)" % (adata.n_obs, adata.n_vars, t_gen))
stage("get_genes_csr_array(min_cells=50)", lambda: sp.get_genes_csr_array(min_cells=50, normalize=False))
stage("spatial_high_variable_genes()  [%d EMDs]" % len(sp.csr_dict), lambda: sp.spatial_high_variable_genes())
top = list(sp.global_distance["Gene"][:n_top])
stage("fit_pattern(n_comp=20, top %d SVGs)" % n_top, lambda: sp.fit_pattern(n_comp=20, gene_list=top))
stage("build_distance_array()  [%d pairs]" % (len(sp.patterns) * (len(sp.patterns) - 1) // 2), lambda: sp.build_distance_array())
np.random.seed(0)
stage("cluster_gene(n_clusters=6, mds_components=20)", lambda: sp.cluster_gene(n_clusters=6, mds_components=20))
stage("get_pattern_array(vote_rate=0.2)", lambda: sp.get_pattern_array(vote_rate=0.2))
print("labels:", np.bincount(np.asarray(sp.genes_labels["labels"])).tolist(), " top SVG:", top[0],
      " z-score range %.2f..%.2f" % (sp.global_distance["z-score"].min(), sp.global_distance["z-score"].max()))
