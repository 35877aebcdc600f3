"""Profiling driver: SMACOF MDS on a 4,000 x 4,000 distance matrix, 20 components, a few iterations."""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
from stminer_b200.Algorithm.algorithm import mds_smacof
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
rng = np.random.default_rng(0)
P = torch.from_numpy(rng.normal(size=(n, 24))).cuda()
D = torch.cdist(P, P).contiguous(); D.fill_diagonal_(0.0); D = 0.5 * (D + D.T)
X0 = rng.uniform(size=(n, 20))
for _ in range(2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); X, s, it = mds_smacof(D, X0, max_iter=iters, eps=-1e300); b.record(); torch.cuda.synchronize()
    print("n=%d d=20: %d iterations in %.3f ms (%.3f ms each), stress %.4f" % (n, int(it.cpu()[0]), a.elapsed_time(b), a.elapsed_time(b) / iters, float(s.cpu()[0])))
