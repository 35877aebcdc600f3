"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck): multiscale EMD (shared-memory and
global-memory tree), device staging, pattern vote, EM fit, pair distances."""
import sys
import numpy as np
sys.path.insert(0, ".")
import torch
from stminer_b200.Algorithm.distance import emd_batched, gmm_pair_distances
from stminer_b200.Algorithm.distribution import fit_spot_matrix
from stminer_b200.Simulate import simulate_spot_matrix
from stminer_b200.staging import DeviceSpotMatrix, stage_csr_on_device
from scipy import sparse

rng = np.random.default_rng(0)
s = rng.permutation(40 * 40)[:700]; src = (s // 40, s % 40, rng.random(700) + 0.1)
tg = []
for m in (300, 120, 1):
    t = rng.permutation(40 * 40)[:m]; tg.append((t // 40, t % 40, rng.random(m) + 0.1))
c, st = emd_batched(src, tg); print("emd smem", c, st)
c2, st2 = emd_batched(src, tg, star_start=True); print("emd star", np.abs(c - c2).max(), st2)
# global-memory tree: a problem beyond the shared-memory capacity
n = 9500; cells = rng.permutation(120 * 120)[:n]
big = (cells // 120, cells % 120, rng.random(n) + 0.1)
t = rng.permutation(120 * 120)[:2200]
c3, st3 = emd_batched(big, [(t // 120, t % 120, rng.random(2200) + 0.1)], max_iter=3000); print("emd global", c3, st3)
sm = simulate_spot_matrix("T", seed=1, n_templates=3, replicas=3)
res = fit_spot_matrix(DeviceSpotMatrix(sm), 5, init=None, seed=1).to_host(); print("fit", res["status"])
p = gmm_pair_distances(res["weights"], res["means"], res["covariances"]); print("pairs", float(p.sum()))
X = sparse.random(300, 50, density=0.2, format="csr", random_state=0, dtype=np.float32); X.data *= 9
xs = np.arange(300) // 20; ys = np.arange(300) % 20
d = stage_csr_on_device(X, xs, ys, ["g%d" % i for i in range(50)], columns=np.arange(5, 40)); torch.cuda.synchronize(); print("stage", d.host.nnz)
from stminer_b200.staging import pattern_vote_on_device
lab = np.full(50, -1, dtype=np.int32); lab[:30] = np.arange(30) % 3
tot, vot = pattern_vote_on_device(X, 50, lab, 3); print("vote", tot.sum(), vot.sum())
