import sys, time, numpy as np
sys.path.insert(0, '.')
from oracle import dist_oracle as do
from stminer_b200.Algorithm.distance import gmm_pair_distances, pairs_to_square, gmm_one_vs_all
import torch
for K in (10, 20, 32, 40, 64, 3):
    G = 30
    W, MU, COV = do.random_gmms(G, K, seed=K)
    ref = do.build_gmm_distance_array(W, MU, COV)
    p = gmm_pair_distances(W, MU, COV)
    sq = pairs_to_square(p, G).cpu().numpy()
    ova = gmm_one_vs_all(W[3], MU[3], COV[3], W, MU, COV).cpu().numpy()
    print(K, np.abs(sq - ref).max(), np.abs(ova - ref[3] - 0).max(), ova[3], ref.max())
W, MU, COV = do.random_gmms(4000, 20, seed=1)
for it in range(3):
    torch.cuda.synchronize(); t = time.time()
    p = gmm_pair_distances(W, MU, COV); torch.cuda.synchronize()
    print('4000 GMMs K=20: %.3f s, %.3e pairs/s' % (time.time() - t, p.numel() / (time.time() - t)))
