"""Profiling driver: V-config fit with device init, then EM-only from fixed parameters (for ncu / timing)."""
import sys, time, numpy as np
sys.path.insert(0, '.')
import torch
from bench import workload
from stminer_b200.Algorithm.distribution import fit_spot_matrix
from stminer_b200.staging import DeviceSpotMatrix
name = sys.argv[1] if len(sys.argv) > 1 else "V"
genes = int(sys.argv[2]) if len(sys.argv) > 2 else 0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
sm = workload(name, 0, genes)
dsm = DeviceSpotMatrix(sm)
K = 20
def timed(fn):
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); r = fn(); b.record(); torch.cuda.synchronize(); return r, a.elapsed_time(b)
one = fit_spot_matrix(dsm, K, init=None, seed=1, max_iter=1)
init = (one.weights.clone(), one.means.clone(), one.covariances.clone())
for i in range(reps):
    full, t_full = timed(lambda: fit_spot_matrix(dsm, K, init=None, seed=1))
    em, t_em = timed(lambda: fit_spot_matrix(dsm, K, init=init))
    _, t_one = timed(lambda: fit_spot_matrix(dsm, K, init=None, seed=1, max_iter=1))
nnz = sm.nnz_per_gene().astype(float)
it_full = full.n_iter.cpu().numpy(); it_em = em.n_iter.cpu().numpy()
fl = lambda it: 30 * 20 * float((nnz * it).sum())
print("%s G=%d nnz=%d buckets=%s" % (name, sm.n_genes, sm.nnz, dsm.plan(K)[1].tolist()))
print("full (init+EM): %.3f ms, mean n_iter %.2f, EM-FLOP rate %.2f TFLOP/s" % (t_full, it_full.mean(), fl(it_full) / t_full / 1e9))
print("EM only (fixed init): %.3f ms, mean n_iter %.2f, %.2f TFLOP/s" % (t_em, it_em.mean(), fl(it_em) / t_em / 1e9))
print("init + 1 EM iteration: %.3f ms" % t_one)
