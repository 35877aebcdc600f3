import sys, numpy as np
sys.path.insert(0, '.')
import torch
from oracle import dist_oracle as do
from stminer_b200.Algorithm.distance import gmm_pair_distances
G = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
W, MU, COV = do.random_gmms(G, 20, seed=1)
for _ in range(2):
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); p = gmm_pair_distances(W, MU, COV); b.record(); torch.cuda.synchronize()
    print('%d GMMs: %.3f ms, %.3e pairs/s' % (G, a.elapsed_time(b), p.numel() / a.elapsed_time(b) * 1e3))
