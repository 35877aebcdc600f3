import sys, numpy as np
sys.path.insert(0, '.')
import torch
from stminer_b200.Algorithm.distance import emd_batched
rng = np.random.default_rng(0)
R, n, G = 60, 1500, 296
s = rng.permutation(R * R)[:n]
src = (s // R, s % R, rng.gamma(1.0, 1.0, n) + 0.01)
tg = []
for g in range(G):
    m = int(rng.integers(400, 900)); t = rng.permutation(R * R)[:m]
    tg.append((t // R, t % R, rng.gamma(1.0, 1.0, m) + 0.01))
for _ in range(2):
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); c, st, it = emd_batched(src, tg, return_iters=True); b.record(); torch.cuda.synchronize()
    print('%d problems (1500 x ~650): %.1f ms, mean pivots %.0f, status %s' % (G, a.elapsed_time(b), it.mean(), np.bincount(st)))
