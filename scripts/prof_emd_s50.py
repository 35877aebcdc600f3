"""Spot EMD at Stereo-seq bin50 size (40,000 source bins): the global-tree kernel.  python scripts/prof_emd_s50.py [n_problems]"""
import sys
import time

import numpy as np
import torch
from scipy import sparse

sys.path.insert(0, ".")
from stminer_b200.Algorithm.distance import emd_batched
from stminer_b200.Simulate import simulate_spot_matrix

G = int(sys.argv[1]) if len(sys.argv) > 1 else 16
sm = simulate_spot_matrix("S50", seed=0, n_templates=max(1, (G + 3) // 4), replicas=4, truncate_int16=False)
X = sparse.csc_matrix((sm.val, sm.row_idx, sm.col_ptr), shape=(sm.n_spots, sm.n_genes))
tot = np.asarray(X.sum(axis=1)).ravel()
keep = tot > 0
src = (sm.spot_x[keep], sm.spot_y[keep], tot[keep])
tg = []
for gi in range(min(G, sm.n_genes)):
    a, b = sm.col_ptr[gi], sm.col_ptr[gi + 1]
    tg.append((sm.spot_x[sm.row_idx[a:b]], sm.spot_y[sm.row_idx[a:b]], sm.val[a:b].astype(np.float64)))
print("source bins %d, target sizes %s" % (len(src[0]), [len(t[0]) for t in tg]))
torch.cuda.synchronize()
t0 = time.perf_counter()
cost, status, iters = emd_batched(src, tg, return_iters=True, max_iter=10 ** 9)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("%d problems in %.2f s = %.2f problems/s (one CTA per problem, %d in flight); pivots %s; status %s"
      % (len(tg), dt, len(tg) / dt, min(len(tg), 148), iters.tolist(), status.tolist()))
print("costs", np.round(cost, 4).tolist())
print("pivots on the problem itself vs numItermax=200000 of the reference: all-level totals above; see STM_EMD_PROFILE for the split")
