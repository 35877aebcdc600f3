"""Spot-EMD timing on Visium-size problems: multiscale start (default) vs the star start.
   python scripts/prof_emd.py [n_problems] [--star] [--nocand]   (--nocand: also time the plain block search)"""
import sys
import time

import numpy as np
import torch
from scipy import sparse

sys.path.insert(0, ".")
from stminer_b200.Algorithm.distance import emd_batched
from stminer_b200.Simulate import simulate_spot_matrix

G = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 148
vsm = simulate_spot_matrix("V", seed=0, n_templates=max(1, (G + 3) // 4), replicas=4, truncate_int16=False)
X = sparse.csc_matrix((vsm.val, vsm.row_idx, vsm.col_ptr), shape=(vsm.n_spots, vsm.n_genes))
src = (vsm.spot_x, vsm.spot_y, np.asarray(X.sum(axis=1)).ravel())
tg = []
for gi in range(min(G, vsm.n_genes)):
    a, b = vsm.col_ptr[gi], vsm.col_ptr[gi + 1]
    tg.append((vsm.spot_x[vsm.row_idx[a:b]], vsm.spot_y[vsm.row_idx[a:b]], vsm.val[a:b].astype(np.float64)))
emd_batched(src, tg[:4])
torch.cuda.synchronize()
res = {}
for star in ([False, True] if "--star" in sys.argv else [False]):
    t0 = time.perf_counter()
    cost, status, iters = emd_batched(src, tg, return_iters=True, star_start=star)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    res[star] = cost
    print("%s start: %d problems in %.3f s = %.1f problems/s; mean pivots %.0f; status counts %s"
          % ("star" if star else "multiscale", len(tg), dt, len(tg) / dt, iters.mean(), np.bincount(status).tolist()))
if "--nocand" in sys.argv:
    t0 = time.perf_counter()
    cost, status, iters = emd_batched(src, tg, return_iters=True, candidates=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("plain block search (no candidate list): %d problems in %.3f s = %.1f problems/s; mean pivots %.0f; max rel diff %.3e"
          % (len(tg), dt, len(tg) / dt, iters.mean(), np.max(np.abs(cost - res[False]) / res[False])))
if len(res) == 2:
    print("max rel diff between the two starts: %.3e" % np.max(np.abs(res[True] - res[False]) / res[True]))
