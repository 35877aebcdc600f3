"""Research prototype: pivot counts of the network simplex from different start bases (star vs multiscale-refined).
Not product code.   python scripts/research/ms_proto.py"""
import ctypes as C
import os
import subprocess
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
so = os.path.join("/tmp", "ns_proto.so")
subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-o", so, os.path.join(HERE, "ns_proto.c"), "-lm"], check=True)
lib = C.CDLL(so)
lib.ns_solve.restype = C.c_int64
P = C.c_void_p
lib.ns_solve.argtypes = [P, P, P, C.c_int, P, P, P, C.c_int, P, P, P, C.c_int, C.c_int64, C.c_int64, P, P, P, P]


def ns(sx, sy, sup, tx, ty, dem, arcs=None, block=0, max_iter=10 ** 9):
    sx = np.ascontiguousarray(sx, np.int32); sy = np.ascontiguousarray(sy, np.int32)
    tx = np.ascontiguousarray(tx, np.int32); ty = np.ascontiguousarray(ty, np.int32)
    sup = np.ascontiguousarray(sup, np.int64); dem = np.ascontiguousarray(dem, np.int64)
    n, m = len(sx), len(tx)
    if arcs is None:
        ai = aj = np.zeros(0, np.int32); af = np.zeros(0, np.int64)
    else:
        ai, aj, af = (np.ascontiguousarray(arcs[0], np.int32), np.ascontiguousarray(arcs[1], np.int32),
                      np.ascontiguousarray(arcs[2], np.int64))
    cost = C.c_double(0)
    par = np.zeros(n + m + 1, np.int32); fl = np.zeros(n + m + 1, np.int64); st = np.zeros(4, np.int64)
    it = lib.ns_solve(sx.ctypes.data, sy.ctypes.data, sup.ctypes.data, n, tx.ctypes.data, ty.ctypes.data, dem.ctypes.data, m,
                      ai.ctypes.data, aj.ctypes.data, af.ctypes.data, len(ai), max_iter, block, C.byref(cost),
                      par.ctypes.data, fl.ctypes.data, st.ctypes.data)
    assert it >= 0
    # final real arcs with positive flow
    v = np.arange(n + m)
    keep = (par[:n + m] != n + m) & (fl[:n + m] > 0)
    vs = v[keep]; ps = par[:n + m][keep]
    I = np.where(vs < n, vs, ps); J = np.where(vs < n, ps, vs) - n
    return cost.value, it, (I.astype(np.int32), J.astype(np.int32), fl[:n + m][keep]), st


def integerize(mass):
    q = np.floor(mass / mass.sum() * 2.0 ** 50 + 0.5).astype(np.int64)
    q[np.argmax(q)] += (1 << 50) - q.sum()
    return q


def coarsen(x, y, mass):
    """(x>>1, y>>1) cells: returns coarse x, y, mass and the fine->coarse map."""
    key = (x.astype(np.int64) >> 1) * (1 << 20) + (y.astype(np.int64) >> 1)
    uniq, inv = np.unique(key, return_inverse=True)
    cm = np.zeros(len(uniq), np.int64)
    np.add.at(cm, inv, mass)
    return (uniq >> 20).astype(np.int32), (uniq & ((1 << 20) - 1)).astype(np.int32), cm, inv


def nwc(supply_ids, supply_amt, demand_ids, demand_amt):
    """North-west corner between two ordered lists (equal totals); yields (supply id, demand id, flow > 0)."""
    out = []
    a, b = 0, 0
    ra = supply_amt[0]; rb = demand_amt[0]
    while True:
        f = min(ra, rb)
        if f > 0:
            out.append((supply_ids[a], demand_ids[b], f))
        ra -= f; rb -= f
        if ra == 0:
            a += 1
            if a == len(supply_ids):
                break
            ra = supply_amt[a]
        if rb == 0:
            b += 1
            if b == len(demand_ids):
                break
            rb = demand_amt[b]
    return out


def refine(coarse_arcs, fs, ft, cs, ct, smap, tmap, order="lex"):
    """coarse_arcs (I, J, F) on the coarse level -> feasible forest arcs (i, j, f) on the fine level.
    fs/ft = (x, y, mass) of the fine sources/targets, cs/ct of the coarse ones, smap/tmap fine->coarse maps."""
    I, J, F = coarse_arcs
    nS, nT = len(cs[0]), len(ct[0])
    kids_s = [[] for _ in range(nS)]
    for i, c in enumerate(smap):
        kids_s[c].append(i)
    kids_t = [[] for _ in range(nT)]
    for j, c in enumerate(tmap):
        kids_t[c].append(j)
    out_arcs = [[] for _ in range(nS)]
    for a in range(len(I)):
        out_arcs[I[a]].append((J[a], F[a]))

    def mort(x, y):
        k = 0
        for b in range(15):
            k |= ((int(x) >> b) & 1) << (2 * b + 1) | ((int(y) >> b) & 1) << (2 * b)
        return k
    if order == "lex":
        def key_s(i):
            return (fs[0][i], fs[1][i])

        def key_t(j):
            return (ft[0][j], ft[1][j])
        ckey = lambda a: (ct[0][a[0]], ct[1][a[0]])
    else:
        def key_s(i):
            return mort(fs[0][i], fs[1][i])

        def key_t(j):
            return mort(ft[0][j], ft[1][j])
        ckey = lambda a: mort(ct[0][a[0]], ct[1][a[0]])
    pieces = [[] for _ in range(nT)]            # per coarse target: (fine source, flow)
    for c in range(nS):
        if not out_arcs[c]:
            continue
        oa = sorted(out_arcs[c], key=ckey)
        kids = sorted(kids_s[c], key=key_s)
        for (i, Jc, f) in nwc(kids, [fs[2][i] for i in kids], [a[0] for a in oa], [a[1] for a in oa]):
            pieces[Jc].append((i, f))
    ai, aj, af = [], [], []
    for c in range(nT):
        pc = sorted(pieces[c], key=lambda p: key_s(p[0]))
        kids = sorted(kids_t[c], key=key_t)
        # a fine source may appear once per coarse target, so (i, j) pairs are distinct
        for (k, j, f) in nwc(list(range(len(pc))), [p[1] for p in pc], kids, [ft[2][j] for j in kids]):
            ai.append(pc[k][0]); aj.append(j); af.append(f)
    return np.array(ai, np.int32), np.array(aj, np.int32), np.array(af, np.int64)


def multiscale(src, tgt, block_of=lambda n, m: 0, min_nodes=64, verbose=True, order='lex'):
    sx, sy, a = src; tx, ty, b = tgt
    S = [(sx.astype(np.int32), sy.astype(np.int32), integerize(a))]
    T = [(tx.astype(np.int32), ty.astype(np.int32), integerize(b))]
    smaps, tmaps = [], []
    while len(S[-1][0]) + len(T[-1][0]) > min_nodes and max(S[-1][0].max(), S[-1][1].max()) > 1:
        cx, cy, cm, inv = coarsen(*S[-1]); S.append((cx, cy, cm)); smaps.append(inv)
        cx, cy, cm, inv = coarsen(*T[-1]); T.append((cx, cy, cm)); tmaps.append(inv)
    arcs = None
    total_it = []
    for lev in range(len(S) - 1, -1, -1):
        n, m = len(S[lev][0]), len(T[lev][0])
        t0 = time.time()
        cost, it, final, st = ns(S[lev][0], S[lev][1], S[lev][2], T[lev][0], T[lev][1], T[lev][2], arcs, block=block_of(n, m))
        if verbose:
            print("   level %d: n=%d m=%d start arcs %s pivots %d (%.2f per node, %d degenerate, cycle %.1f max %d, range %.0f) %.2fs"
                  % (lev, n, m, "star" if arcs is None else len(arcs[0]), it, it / (n + m), st[0], st[1] / max(it, 1), st[3],
                     st[2] / max(it, 1), time.time() - t0))
        total_it.append((n + m, it))
        if lev > 0:
            arcs = refine(final, S[lev - 1], T[lev - 1], S[lev], T[lev], smaps[lev - 1], tmaps[lev - 1], order=order)
    return cost / 2.0 ** 50, total_it


if __name__ == "__main__":
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_golden_emd_large import visium_problems
    src, tg = visium_problems()
    blk = lambda n, m: max(int(np.sqrt(n * m)), min(n * m, 7 * n))
    for g in [0, 2, 5]:
        t = tg[g]
        print("gene %d: n=%d m=%d" % (g, len(src[0]), len(t[0])))
        t0 = time.time()
        c0, it0, _, st = ns(src[0], src[1], integerize(src[2]), t[0], t[1], integerize(t[2]), None, block=blk(len(src[0]), len(t[0])))
        print("  star: cost %.10g pivots %d (%.2f per node; degenerate %d, cycle %.1f, range %.0f) %.1fs"
              % (c0 / 2.0 ** 50, it0, it0 / (len(src[0]) + len(t[0])), st[0], st[1] / it0, st[2] / it0, time.time() - t0))
        t0 = time.time()
        for order in ("lex", "morton"):
            c1, its = multiscale(src, t, blk, verbose=False, order=order)
            print("  multiscale %s: cost %.10g  fine pivots %d, all levels %s  %.1fs" % (order, c1, its[-1][1], its, time.time() - t0))
            assert abs(c1 - c0 / 2.0 ** 50) < 1e-9 * c1
