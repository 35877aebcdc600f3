import sys, time, numpy as np, pandas as pd
sys.path.insert(0, ".")
import torch
from oracle import dist_oracle as do
from stminer_b200.Algorithm.distance import gmm_pair_distances, pairs_to_square
from stminer_b200.Algorithm.algorithm import mds_fit_transform, kmeans_fit, cluster
for G in (500, 2000):
    W, MU, COV = do.random_gmms(G, 20, seed=1)
    D = pairs_to_square(gmm_pair_distances(W, MU, COV), G).cpu().numpy()
    df = pd.DataFrame(D)
    cluster(df.iloc[:64, :64], 20, 4)
    for rep in range(2):
        np.random.seed(0)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        emb, stress, nit = mds_fit_transform(D, 20)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        km = kmeans_fit(emb, 6)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        print("G=%d mds %.3f s (n_iter %d)  kmeans %.3f s (n_iter %d)" % (G, t1 - t0, nit, t2 - t1, km.n_iter_))
    import cProfile, pstats
    np.random.seed(0)
    pr = cProfile.Profile(); pr.enable(); mds_fit_transform(D, 20); torch.cuda.synchronize(); pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(8)
