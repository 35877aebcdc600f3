"""Visium-size EMD: pricing block size (whole target rows) vs time and pivots."""
import sys, time, numpy as np
sys.path.insert(0, '.')
import torch
from scipy import sparse
from stminer_b200.Algorithm.distance import emd_batched
from stminer_b200.Simulate import simulate_spot_matrix
vsm = simulate_spot_matrix("V", seed=0, n_templates=37, replicas=4, truncate_int16=False)
X = sparse.csc_matrix((vsm.val, vsm.row_idx, vsm.col_ptr), shape=(vsm.n_spots, vsm.n_genes))
src = (vsm.spot_x, vsm.spot_y, np.asarray(X.sum(axis=1)).ravel())
tg = []
for gi in range(148):
    a, b = vsm.col_ptr[gi], vsm.col_ptr[gi + 1]
    tg.append((vsm.spot_x[vsm.row_idx[a:b]], vsm.spot_y[vsm.row_idx[a:b]], vsm.val[a:b].astype(np.float64)))
emd_batched(src, tg[:4])
n = len(src[0])
for rows in (0, 5, 7, 9, 11, 14, 18):
    torch.cuda.synchronize(); t = time.perf_counter()
    c, st, it = emd_batched(src, tg, block_size=rows * n, return_iters=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print("rows %2d: %.3f s (%.1f problems/s), mean pivots %.0f, optimal %d" % (rows, dt, len(tg) / dt, it.mean(), (st == 0).sum()), flush=True)
