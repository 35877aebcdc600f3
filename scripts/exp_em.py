"""Experiment driver (one process per library variant): EM-only and device-init timings + fit quality."""
import os, sys, time, json, numpy as np
sys.path.insert(0, '.')
import torch
from bench import workload
from stminer_b200.Algorithm.distribution import fit_spot_matrix
from stminer_b200.staging import DeviceSpotMatrix
name = sys.argv[1]; genes = int(sys.argv[2]); mode = sys.argv[3]
sm = workload(name, 0, genes)
dsm = DeviceSpotMatrix(sm)
K = 20
nnz = sm.nnz_per_gene().astype(float)
def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); r = fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return r, best
tag = os.path.basename(os.environ.get("STM_LIB_PATH", "default"))
one = fit_spot_matrix(dsm, K, init=None, seed=1, max_iter=1)
init = (one.weights.clone(), one.means.clone(), one.covariances.clone())
em, t_em = timed(lambda: fit_spot_matrix(dsm, K, init=init))
it = em.n_iter.cpu().numpy()
print("%s %s G=%d EM-only %.3f ms  n_iter %.2f  %.2f TFLOP/s  lb %.4f" % (tag, name, sm.n_genes, t_em, it.mean(), 600 * float((nnz * it).sum()) / t_em / 1e9, em.lower_bound.mean().item()), flush=True)
if mode == "init":
    for ll, cand, cap in [(255, 0, 0), (16, 0, 0), (8, 0, 0), (4, 0, 0), (2, 0, 0), (8, 1, 0), (4, 1, 0), (8, 2, 0), (8, 0, 512), (8, 0, 2048), (16, 0, 512), (1, 1, 0)]:
        try:
            r, t = timed(lambda: fit_spot_matrix(dsm, K, init=None, seed=1, lloyd_iters=ll, kpp_candidates=cand, seed_cap=cap))
            _, t1 = timed(lambda: fit_spot_matrix(dsm, K, init=None, seed=1, lloyd_iters=ll, kpp_candidates=cand, seed_cap=cap, max_iter=1))
            print("  init lloyd=%3d cand=%d cap=%4d: total %.3f ms (init+1 iter %.3f)  n_iter %.2f  lb %.4f  conv %.3f" % (ll, cand, cap, t, t1, r.n_iter.float().mean().item(), r.lower_bound.mean().item(), r.converged.float().mean().item()), flush=True)
        except Exception as e:
            print("  init", ll, cand, cap, "failed", e)
