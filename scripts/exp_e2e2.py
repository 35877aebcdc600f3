import sys, time, numpy as np
sys.path.insert(0, '.')
import torch
from bench import workload
from stminer_b200.Algorithm.distribution import fit_spot_matrix_host
sm = workload("S50", 0, 0).compact().pin_memory()
for fr, tl in (((0.05, 0.2, 0.5, 1.0), False), ((0.05, 0.2, 0.5, 1.0), True), ((0.03, 0.1, 0.25, 0.5, 0.75, 1.0), True), ((0.02, 0.06, 0.14, 0.3, 0.5, 0.7, 0.85, 1.0), True)):
    for _ in range(2):
        fit_spot_matrix_host(sm, 20, init=None, seed=1, chunk_fractions=fr, two_lanes=tl)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5):
        out = fit_spot_matrix_host(sm, 20, init=None, seed=1, chunk_fractions=fr, two_lanes=tl)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
    print("lanes %s fractions %s: %.2f ms  (%.0f genes/s)" % (tl, fr, dt * 1e3, sm.n_genes / dt), flush=True)
