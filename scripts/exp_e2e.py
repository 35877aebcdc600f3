import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import torch
from bench import workload
from stminer_b200.Algorithm.distribution import fit_spot_matrix_host
sm = workload("S50", 0, 0).compact().pin_memory()
for nc in (1, 2, 4, 6, 8, 12, 16):
    for _ in range(2):
        fit_spot_matrix_host(sm, 20, init=None, seed=1, n_chunks=nc)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(4):
        out = fit_spot_matrix_host(sm, 20, init=None, seed=1, n_chunks=nc)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 4
    print('n_chunks %2d: %.2f ms  (%.0f genes/s)' % (nc, dt * 1e3, sm.n_genes / dt))
# raw H2D rate
import torch
a = torch.empty(512 << 20, dtype=torch.uint8).pin_memory(); b = torch.empty_like(a, device='cuda')
for _ in range(2): b.copy_(a, non_blocking=True)
torch.cuda.synchronize(); t = time.perf_counter(); b.copy_(a, non_blocking=True); torch.cuda.synchronize()
print('pinned H2D: %.1f GB/s' % (a.numel() / (time.perf_counter() - t) / 1e9))
