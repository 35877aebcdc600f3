import sys, time, numpy as np
sys.path.insert(0, '.')
from stminer_b200.Algorithm.distance import emd_batched
from stminer_b200.Simulate import simulate_spot_matrix
from scipy import sparse
sm = simulate_spot_matrix("V", 0, n_templates=2, replicas=2, truncate_int16=False)
X = sparse.csc_matrix((sm.val, sm.row_idx, sm.col_ptr), shape=(sm.n_spots, sm.n_genes))
tot = np.asarray(X.sum(axis=1)).ravel()
src = (sm.spot_x, sm.spot_y, tot)
tg = []
for g in range(sm.n_genes):
    a, b = sm.col_ptr[g], sm.col_ptr[g + 1]
    tg.append((sm.spot_x[sm.row_idx[a:b]], sm.spot_y[sm.row_idx[a:b]], sm.val[a:b].astype(np.float64)))
import torch
for bs in (0,):
    torch.cuda.synchronize(); t = time.time()
    cost, status, iters = emd_batched(src, tg, max_iter=10**7, block_size=bs, return_iters=True)
    torch.cuda.synchronize(); print('block', bs, 'time %.2f' % (time.time() - t), iters)
