"""A/B timing of library variants (STM_LIB_PATH): full fit (device init + EM) on S50 and on an S10 subset."""
import os, sys, numpy as np
sys.path.insert(0, '.')
import torch
from bench import workload
from stminer_b200.Algorithm.distribution import fit_spot_matrix
from stminer_b200.staging import DeviceSpotMatrix
from stminer_b200.Simulate import simulate_long_genes
tag = os.path.basename(os.environ.get("STM_LIB_PATH", "default"))
K = 20
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run(name, sm, reps=5):
    dsm = DeviceSpotMatrix(sm.compact() if name == "S50" else sm)
    res = None
    for _ in range(2):
        res = fit_spot_matrix(dsm, K, init=None, seed=1, out=res)
    ts = []
    for r in range(reps):
        flush.fill_(r); torch.cuda.synchronize()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); res = fit_spot_matrix(dsm, K, init=None, seed=1, out=res); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    h = res.to_host()
    fl = 600.0 * float((sm.nnz_per_gene() * h["n_iter"] * (h["status"] == 0)).sum())
    print("%-22s %-4s G=%5d  min %.3f  med %.3f ms  %.2f TFLOP/s  n_iter %.2f lb %.4f" % (
        tag, name, sm.n_genes, min(ts), float(np.median(ts)), fl / np.median(ts) / 1e9, h["n_iter"].mean(), h["lower_bound"].mean()), flush=True)
which = sys.argv[1] if len(sys.argv) > 1 else "S50,S10"
if "S50" in which:
    run("S50", workload("S50", 0))
if "S10" in which:
    p = "/tmp/stm_s10_296.npz"
    if os.path.exists(p):
        from stminer_b200.staging import SpotMatrix
        z = np.load(p); s10 = SpotMatrix(z["col_ptr"], z["row_idx"], z["val"], z["spot_x"], z["spot_y"], ["g%d" % i for i in range(len(z["col_ptr"]) - 1)])
    else:
        s10 = simulate_long_genes(296, seed=0)
        np.savez(p, col_ptr=s10.col_ptr, row_idx=s10.row_idx, val=s10.val, spot_x=s10.spot_x, spot_y=s10.spot_y)
    run("S10", s10)
