"""Profiling driver: one fit of a Stereo-seq bin10-shaped long-gene set (cluster buckets + one-CTA buckets)."""
import sys
sys.path.insert(0, '.')
import torch
from stminer_b200.Simulate import simulate_long_genes
from stminer_b200.Algorithm.distribution import fit_spot_matrix
from stminer_b200.staging import DeviceSpotMatrix
sm = simulate_long_genes(int(sys.argv[1]) if len(sys.argv) > 1 else 296, seed=0)
dsm = DeviceSpotMatrix(sm)
print("buckets", dsm.plan(20, 2)[1].tolist())
for _ in range(2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); r = fit_spot_matrix(dsm, 20, init=None, seed=1); b.record(); torch.cuda.synchronize()
    print("fit %.3f ms" % a.elapsed_time(b))
