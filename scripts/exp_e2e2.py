import sys, time, numpy as np
sys.path.insert(0, '.')
import torch
from bench import workload
from stminer_b200.Algorithm.distribution import fit_spot_matrix_host
sm = workload("S50", 0, 0).compact().pin_memory()
for fr in (None, (0.05, 0.2, 0.5, 1.0), (0.1, 0.3, 0.6, 1.0), (0.04, 0.12, 0.3, 0.6, 1.0), (0.08, 0.4, 1.0), (0.03, 0.1, 0.25, 0.5, 0.75, 1.0)):
    for _ in range(2):
        fit_spot_matrix_host(sm, 20, init=None, seed=1, chunk_fractions=fr)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5):
        out = fit_spot_matrix_host(sm, 20, init=None, seed=1, chunk_fractions=fr)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
    print('fractions %s: %.2f ms  (%.0f genes/s)' % (fr, dt * 1e3, sm.n_genes / dt), flush=True)
