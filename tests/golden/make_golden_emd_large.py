"""Visium-size (and one Stereo-seq-bin50-size) golden optima for the spot-level EMD, from an INDEPENDENT solver.

    python tests/golden/make_golden_emd_large.py            # writes tests/golden/emd_large.npz

POT (what STMiner/Algorithm/distance.py:200-201 calls) is not installable in this image, so the exact optimum of the
transport LP of ``calculate_ot_distance`` is computed with scipy's HiGHS.  The dense LP of a Visium-size problem has
n x m = 5-17 million columns, so HiGHS solves it by COLUMN GENERATION: a restricted LP over a sparse arc set (the
nearest targets of every source and the nearest sources of every target, plus slack columns at a big-M price),
then all n x m reduced costs are evaluated with the LP's duals in numpy, the violating arcs are added and the LP is
solved again.  It stops when NO arc of the full problem has a negative reduced cost and no slack is used: by LP
duality the restricted optimum is then the optimum of the full problem.  That final certificate — dual feasibility
over all n x m arcs, objective = dual objective — is checked here in float64 numpy: the VALUE is HiGHS's and its
optimality is proved by duality, independently of this repository's solver.  To save hours of column generation the
FIRST arc set is seeded with the support of the oracle's plan (``--seed-oracle``, used for the committed file); the seed
only decides which columns HiGHS starts with — a wrong or sub-optimal seed costs more rounds, it cannot change the
certified optimum.  Case 0 of the committed file was also run unseeded (5 rounds, 22 min): same value to 1e-13.

Problems: the source is the per-spot total image of a synthetic Visium data set (4,992 spots on the hex lattice, the
``global_matrix`` of SPFinder.py:279-283) and the targets are genes of that data set, as ``spatial_high_variable_genes``
pairs them.  Only the inputs needed to rebuild them (seed, gene index) and the optimum are stored.
"""
import os
import sys
import time

import numpy as np
from scipy import sparse
from scipy.optimize import linprog

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def visium_problems(seed=11, n_templates=6, replicas=1):
    from stminer_b200.Simulate import simulate_spot_matrix
    sm = simulate_spot_matrix("V", seed=seed, n_templates=n_templates, replicas=replicas, truncate_int16=False)
    X = sparse.csc_matrix((sm.val, sm.row_idx, sm.col_ptr), shape=(sm.n_spots, sm.n_genes))
    src = (sm.spot_x.astype(np.int32), sm.spot_y.astype(np.int32), np.asarray(X.sum(axis=1)).ravel().astype(np.float64))
    tg = []
    for g in range(sm.n_genes):
        a, b = sm.col_ptr[g], sm.col_ptr[g + 1]
        r = sm.row_idx[a:b]
        tg.append((sm.spot_x[r].astype(np.int32), sm.spot_y[r].astype(np.int32), sm.val[a:b].astype(np.float64)))
    return src, tg


def grid_problem(R, n_src_frac, m, seed):
    """Full R x R grid source with random masses (Stereo-seq-like), blob-shaped target of m cells."""
    rng = np.random.default_rng(seed)
    xs, ys = np.meshgrid(np.arange(R), np.arange(R), indexing="ij")
    keep = rng.random(R * R) < n_src_frac
    sx, sy = xs.ravel()[keep].astype(np.int32), ys.ravel()[keep].astype(np.int32)
    sm = rng.gamma(2.0, 1.0, len(sx)) + 0.05
    cx, cy = rng.uniform(0.3, 0.7, 2) * R
    score = np.exp(-0.5 * (((xs - cx) / (0.2 * R)) ** 2 + ((ys - cy) / (0.12 * R)) ** 2)).ravel() * rng.exponential(size=R * R)
    idx = np.sort(np.argpartition(score, R * R - m)[R * R - m:])
    tx, ty = (idx // R).astype(np.int32), (idx % R).astype(np.int32)
    tm = 1.0 + rng.poisson(2.0, m).astype(np.float64)
    return (sx, sy, sm), (tx, ty, tm)


def oracle_support(src, tgt):
    """(i, j) pairs carrying flow in the plan of the repository's CPU network simplex (only a column seed, see the
    module docstring)."""
    sys.path.insert(0, os.path.join(ROOT, "scripts", "research"))
    import ms_proto as mp
    sx, sy, a = src; tx, ty, b = tgt
    S = (sx.astype(np.int32), sy.astype(np.int32), mp.integerize(a)); T = (tx.astype(np.int32), ty.astype(np.int32), mp.integerize(b))
    smaps, tmaps, Ss, Ts = [], [], [S], [T]
    while len(Ss[-1][0]) + len(Ts[-1][0]) > 64:
        cx, cy, cm, inv = mp.coarsen(*Ss[-1]); Ss.append((cx, cy, cm)); smaps.append(inv)
        cx, cy, cm, inv = mp.coarsen(*Ts[-1]); Ts.append((cx, cy, cm)); tmaps.append(inv)
    arcs = None
    for lev in range(len(Ss) - 1, -1, -1):
        n_, m_ = len(Ss[lev][0]), len(Ts[lev][0])
        _, _, final, _ = mp.ns(Ss[lev][0], Ss[lev][1], Ss[lev][2], Ts[lev][0], Ts[lev][1], Ts[lev][2], arcs,
                               block=min(n_ * m_, max(int(np.sqrt(n_ * m_)), 7 * n_)))
        if lev > 0:
            arcs = mp.refine(final, Ss[lev - 1], Ts[lev - 1], Ss[lev], Ts[lev], smaps[lev - 1], tmaps[lev - 1], order="morton")
    return final[0].astype(np.int64), final[1].astype(np.int64)


def solve_colgen(src, tgt, knn=6, verbose=True, seed=None):
    sx, sy, a = src
    tx, ty, b = tgt
    n, m = len(sx), len(tx)
    a = a / a.sum(); b = b / b.sum()
    C = ((sx[:, None].astype(np.float64) - tx[None]) ** 2 + (sy[:, None].astype(np.float64) - ty[None]) ** 2)
    bigM = 4.0 * C.max() + 1.0
    arcs = set()
    k1 = min(knn, m); k2 = min(knn, n)
    nn_t = np.argpartition(C, k1 - 1, axis=1)[:, :k1]
    for i in range(n):
        arcs.update((i * m + nn_t[i]).tolist())
    nn_s = np.argpartition(C, k2 - 1, axis=0)[:k2, :]
    for j in range(m):
        arcs.update((nn_s[:, j] * m + j).tolist())
    if seed is not None:
        arcs.update((seed[0] * m + seed[1]).tolist())
    it = 0
    while True:
        it += 1
        e = np.fromiter(arcs, dtype=np.int64, count=len(arcs)); e.sort()
        ei, ej = e // m, e % m
        E = len(e)
        # columns: E real arcs, then n source slacks, then m target slacks
        rows = np.concatenate([ei, n + ej, np.arange(n), n + np.arange(m)])
        cols = np.concatenate([np.arange(E), np.arange(E), E + np.arange(n), E + n + np.arange(m)])
        A = sparse.csr_matrix((np.ones(len(rows)), (rows, cols)), shape=(n + m, E + n + m))
        c = np.concatenate([C[ei, ej], np.full(n + m, bigM)])
        t0 = time.time()
        res = linprog(c, A_eq=A, b_eq=np.concatenate([a, b]), bounds=(0, None), method="highs",
                      options={"primal_feasibility_tolerance": 1e-10, "dual_feasibility_tolerance": 1e-10})
        if res.status != 0:
            raise RuntimeError(res.message)
        y = res.eqlin.marginals
        u, v = y[:n], y[n:]
        R = C - u[:, None] - v[None]                 # reduced costs of ALL n x m arcs
        slack = res.x[E:].sum()
        rmin = R.min()
        if verbose:
            print("  round %d: %d arcs, obj %.12g, slack %.3g, min reduced cost %.3g, %.1f s" %
                  (it, E, res.fun, slack, rmin, time.time() - t0), flush=True)
        if rmin >= -1e-7 and slack <= 1e-12:
            # certificate: dual feasible on the full problem (up to rmin, which bounds the gap by |rmin| * total mass)
            dual_obj = float(a @ u + b @ v)
            gap = abs(res.fun - dual_obj)
            assert gap <= 1e-9 * max(1.0, abs(res.fun)), gap
            return float(res.fun), float(max(-rmin, 0.0)), it, E
        # add the most violating arcs: per source the worst few, and everything below a threshold (bounded)
        viol = np.argwhere(R < -1e-9)
        if len(viol) > 400000:
            thr = np.partition(R[R < -1e-9], 400000)[400000]
            viol = np.argwhere(R < thr)
        arcs.update((viol[:, 0] * m + viol[:, 1]).tolist())
        per_src = np.argmin(R, axis=1)
        arcs.update((np.arange(n) * m + per_src).tolist())


def main():
    out = {}
    cases = []
    src, tg = visium_problems()
    sizes = np.array([len(t[0]) for t in tg])
    pick = list(np.argsort(sizes)[[0, len(sizes) // 2, -1]])            # small / median / largest gene
    # plus two thinned variants (sparser genes, as after a vmax clip / low expression)
    for g in pick:
        cases.append(("visium_seed11_gene%d" % g, src, tg[g]))
    rng = np.random.default_rng(5)
    for g in pick[1:]:
        t = tg[g]
        keep = np.sort(rng.permutation(len(t[0]))[:max(50, len(t[0]) // 4)])
        cases.append(("visium_seed11_gene%d_thin4" % g, src, (t[0][keep], t[1][keep], t[2][keep])))
    if "--big" in sys.argv:                                               # beyond one CTA: Stereo-seq bin50-like
        s, t = grid_problem(120, 0.9, 3000, seed=3)
        cases.append(("grid120_seed3", s, t))
    for c, (name, s, t) in enumerate(cases):
        print("case %d %s: n=%d m=%d" % (c, name, len(s[0]), len(t[0])), flush=True)
        t0 = time.time()
        cost, rmin, rounds, E = solve_colgen(s, t, seed=oracle_support(s, t) if "--seed-oracle" in sys.argv else None)
        print("  -> optimum %.12g (%d rounds, %d arcs, %.0f s)" % (cost, rounds, E, time.time() - t0), flush=True)
        p = "c%d_" % c
        out[p + "name"] = np.array(name)
        out[p + "sx"] = s[0]; out[p + "sy"] = s[1]; out[p + "sm"] = s[2]
        out[p + "tx"] = t[0]; out[p + "ty"] = t[1]; out[p + "tm"] = t[2]
        out[p + "cost"] = np.float64(cost); out[p + "dual_infeasibility"] = np.float64(rmin)
        out["n"] = np.int64(c + 1)
        np.savez_compressed(os.path.join(HERE, "emd_large.npz"), **out)
    print("emd_large.npz:", len(cases), "problems")


if __name__ == "__main__":
    main()
