"""Generates the committed golden fixtures (run in the build container, where /root/reference exists).

    python tests/golden/make_golden.py

* gmm_fixed_init.npz  — real ``sklearn.mixture.GaussianMixture`` run on the count-replicated list exactly as
  STMiner/Algorithm/distribution.py:126-127 does, from injected initial parameters.
* gmm_dist.npz        — K x K Hellinger-cost assignment distances where the LSAP step is the reference's OWN
  ``linear_sum`` (STMiner/Algorithm/algorithm.py:10-26, loaded stand-alone by file path) and the Hellinger
  step is STMiner/Algorithm/distance.py:50-80 restated verbatim (that file cannot import here: it needs POT).

Nothing at test time reads /root/reference; only this script does.
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import dist_oracle as do  # noqa: E402
from oracle import gmm_oracle as go  # noqa: E402
from stminer_b200.Simulate import simulate_spot_matrix  # noqa: E402


def make_gmm():
    out = {}
    idx = 0
    for config, K, nt, rep in [("T", 10, 3, 2), ("V", 20, 3, 1), ("T", 4, 2, 1)]:
        sm = simulate_spot_matrix(config, seed=21, n_templates=nt, replicas=rep)
        for g in range(sm.n_genes):
            xy, c = sm.gene_spots(g)
            if len(c) < K:
                continue
            w0, mu0, cov0 = go.kmeans_init(xy, c, K, seed=100 + idx)
            ref = go.sklearn_fit_from_init(xy, c, w0, mu0, cov0)
            pre = "g%d_" % idx
            out[pre + "xy"] = xy.astype(np.int32); out[pre + "counts"] = c.astype(np.int32)
            out[pre + "w0"] = w0; out[pre + "mu0"] = mu0; out[pre + "cov0"] = cov0
            out[pre + "weights"] = ref["weights"]; out[pre + "means"] = ref["means"]
            out[pre + "covariances"] = ref["covariances"]
            out[pre + "n_iter"] = np.int64(ref["n_iter"]); out[pre + "converged"] = np.bool_(ref["converged"])
            out[pre + "lower_bound"] = np.float64(ref["lower_bound"])
            idx += 1
    out["n"] = np.int64(idx)
    np.savez_compressed(os.path.join(HERE, "gmm_fixed_init.npz"), **out)
    print("gmm_fixed_init.npz:", idx, "genes")


def make_dist():
    ref_alg = "/root/reference/STMiner/Algorithm/algorithm.py"
    spec = importlib.util.spec_from_file_location("ref_algorithm", ref_alg)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)                       # the reference's own linear_sum
    out = {}
    for n, (G, K) in enumerate([(12, 10), (10, 20), (6, 5), (5, 33)]):
        W, MU, COV = do.random_gmms(G, K, extent=60.0, seed=40 + n)
        D = np.zeros((G, G))
        for i in range(G):
            for j in range(i + 1, G):
                C = do.cost_matrix(W[i], MU[i], COV[i], W[j], MU[j], COV[j])   # verbatim distance.py:29-34,71-80
                D[i, j] = D[j, i] = mod.linear_sum(C)
        out["s%d_W" % n] = W; out["s%d_MU" % n] = MU; out["s%d_COV" % n] = COV; out["s%d_D" % n] = D
    out["n"] = np.int64(4)
    np.savez_compressed(os.path.join(HERE, "gmm_dist.npz"), **out)
    print("gmm_dist.npz written")


def make_emd():
    """Exact transport-LP optima from scipy's HiGHS (an independent solver; POT is not installable here)."""
    from oracle import emd_oracle as eo
    rng = np.random.default_rng(77)
    out = {}
    cases = [(6, 5, 5), (40, 25, 9), (100, 140, 14), (282, 150, 17), (200, 200, 17), (64, 1, 8), (1, 30, 8)]
    for c, (n, m, R) in enumerate(cases):
        s = rng.permutation(R * R)[:n]; t = rng.permutation(R * R)[:m]
        sm = rng.gamma(0.8, 2.0, n) + 1e-3; tm = rng.gamma(0.8, 2.0, m) + 1e-3
        if c == 4:   # integer counts with many ties (degenerate pivots)
            sm = rng.integers(1, 4, n).astype(float); tm = rng.integers(1, 4, m).astype(float)
        ref = eo.emd_linprog(s // R, s % R, sm, t // R, t % R, tm)
        pre = "c%d_" % c
        out[pre + "sx"] = (s // R).astype(np.int32); out[pre + "sy"] = (s % R).astype(np.int32); out[pre + "sm"] = sm
        out[pre + "tx"] = (t // R).astype(np.int32); out[pre + "ty"] = (t % R).astype(np.int32); out[pre + "tm"] = tm
        out[pre + "cost"] = np.float64(ref)
    out["n"] = np.int64(len(cases))
    np.savez_compressed(os.path.join(HERE, "emd_small.npz"), **out)
    print("emd_small.npz:", len(cases), "problems")


if __name__ == "__main__":
    make_gmm()
    make_dist()
    make_emd()
