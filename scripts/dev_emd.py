import sys, time, numpy as np
sys.path.insert(0, '.')
from oracle import emd_oracle as eo
from stminer_b200.Algorithm.distance import emd_batched
rng = np.random.default_rng(0)
def prob(n, m, R):
    s = rng.permutation(R * R)[:n]; t = rng.permutation(R * R)[:m]
    return (s // R, s % R, rng.gamma(1.0, 1.0, n) + 0.01), (t // R, t % R, rng.gamma(1.0, 1.0, m) + 0.01)
for (n, ms, R) in [(5, [4, 3, 5], 6), (30, [20, 25, 7], 10), (282, [150, 100, 282, 1], 17), (1000, [500, 300, 800], 40)]:
    src, _ = prob(n, 1, R)
    tg = [prob(1, m, R)[1] for m in ms]
    cost, status, iters = emd_batched(src, tg, return_iters=True)
    for g, t in enumerate(tg):
        c, st, it = eo.emd_network_simplex(*src, *t)
        print(n, len(t[0]), 'gpu %.10f st %d it %d | oracle %.10f st %d it %d | rel %.2e' % (cost[g], status[g], iters[g], c, st, it, abs(cost[g] - c) / max(c, 1e-300)))
# V-size timing
from stminer_b200.Simulate import simulate_spot_matrix
from scipy import sparse
sm = simulate_spot_matrix("V", 0, n_templates=8, replicas=4, truncate_int16=False)
X = sparse.csc_matrix((sm.val, sm.row_idx, sm.col_ptr), shape=(sm.n_spots, sm.n_genes))
tot = np.asarray(X.sum(axis=1)).ravel()
src = (sm.spot_x, sm.spot_y, tot)
tg = []
for g in range(sm.n_genes):
    a, b = sm.col_ptr[g], sm.col_ptr[g + 1]
    tg.append((sm.spot_x[sm.row_idx[a:b]], sm.spot_y[sm.row_idx[a:b]], sm.val[a:b].astype(np.float64)))
import torch
for bs in (0, 8192, 16384):
    torch.cuda.synchronize(); t = time.time()
    cost, status, iters = emd_batched(src, tg, max_iter=10**7, block_size=bs, return_iters=True)
    torch.cuda.synchronize(); dt = time.time() - t
    print('V %d genes block %d: %.2f s, status %s, mean pivots %.0f, max %d; cost[:3] %s' % (len(tg), bs, dt, np.bincount(status), iters.mean(), iters.max(), cost[:3]))
c, st, it = eo.emd_network_simplex(*src, *tg[0], max_iter=10**7)
print('oracle gene0 %.8f pivots %d' % (c, it), 'gpu', cost[0])
