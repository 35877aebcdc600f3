import sys, time, cProfile, pstats
sys.path.insert(0, '.')
import numpy as np, torch
from stminer_b200.Simulate import simulate_adata
from stminer_b200.SPFinder import SPFinder
adata = simulate_adata("V", seed=0)
sp = SPFinder(adata)
sp.get_genes_csr_array(min_cells=50, normalize=False)
sp.spatial_high_variable_genes()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
sp.spatial_high_variable_genes(); torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
