"""Long-gene (Stereo-seq bin10-shaped) fit: thread-block-cluster split vs one CTA streaming from L2."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from stminer_b200.Simulate import simulate_long_genes
from stminer_b200.staging import DeviceSpotMatrix
from stminer_b200.Algorithm.distribution import fit_spot_matrix
from stminer_b200 import _lib

G = int(sys.argv[1]) if len(sys.argv) > 1 else 296
sm = simulate_long_genes(G, seed=0)
print("genes %d nnz %d (mean %.0f, max %d)" % (sm.n_genes, sm.nnz, sm.nnz / G, sm.nnz_per_gene().max()), flush=True)
dsm = DeviceSpotMatrix(sm)
K = 20
for no_cl in (False, True):
    fl = _lib.fit_flags(device_init=True, no_clusters=no_cl)
    print("no_clusters=%s buckets=%s" % (no_cl, dsm.plan(K, fl)[1].tolist()))
    res = None
    for _ in range(2):
        res = fit_spot_matrix(dsm, K, init=None, seed=1, out=res, no_clusters=no_cl)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        res = fit_spot_matrix(dsm, K, init=None, seed=1, out=res, no_clusters=no_cl)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    h = res.to_host()
    flops = 30.0 * float((sm.nnz_per_gene() * h["n_iter"] * (h["status"] == 0)).sum()) * K
    print("  %.2f ms per fit of %d genes = %.0f genes/s, %.2f Gspot/s, EM %.1f TFLOP/s, mean n_iter %.1f, ok %d, mean lb %.4f"
          % (ms, G, G / ms * 1e3, sm.nnz / ms / 1e6, flops / ms / 1e9, h["n_iter"].mean(), (h["status"] == 0).sum(),
             h["lower_bound"].mean()), flush=True)
