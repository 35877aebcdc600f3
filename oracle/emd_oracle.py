"""Python side of the spot-level EMD oracle.  TEST INFRASTRUCTURE ONLY.

* ``emd_network_simplex`` — ctypes binding of ``oracle/emd_oracle.c`` (see its header for the reference
  citations and pinning).
* ``emd_linprog`` — independent exact LP solve of the same transport problem with scipy's HiGHS; used to pin
  the network simplex on small problems (POT==0.9.1, which the reference calls, is not installable here).
* ``csr_to_problem`` — the staging of ``calculate_ot_distance`` (STMiner/Algorithm/distance.py:191-199):
  supports = ``nonzero()`` coordinates, masses = ``data[data > 0] / data.sum()``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "libstm_oracle.so")
        src = os.path.join(HERE, "emd_oracle.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            subprocess.run(["make", "-C", HERE], check=True, capture_output=True)
        lib = C.CDLL(path)
        lib.stm_oracle_emd.restype = C.c_int
        lib.stm_oracle_emd.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
        lib.stm_oracle_emd_ms.restype = C.c_int
        lib.stm_oracle_emd_ms.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_int, C.c_int64, C.c_int64, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _LIB = lib
    return _LIB


def csr_to_problem(img):
    """distance.py:192-199 for one image: (x[int32], y[int32], mass[float64]) of the stored positive entries."""
    img = img.tocsr()
    coo = img.tocoo()
    # the reference pairs nonzero() coordinates with data[data > 0]; both walk the CSR storage in order
    keep = coo.data > 0
    x, y = coo.row[keep].astype(np.int32), coo.col[keep].astype(np.int32)
    mass = coo.data[keep].astype(np.float64) / img.data.sum()
    return x, y, mass


def emd_network_simplex(sx, sy, smass, tx, ty, tmass, max_iter=200000, block_size=0, cand_threads=0):
    """(cost, status, pivots).  max_iter mirrors ``numItermax=200000`` (distance.py:201).  ``cand_threads`` = T
    prices like the CUDA kernel with T threads per problem (candidate list on whole-row blocks, see emd_oracle.c;
    T = ``stminer_b200.Algorithm.distance.emd_pricing_threads``); 0 = plain block search."""
    sx = np.ascontiguousarray(sx, dtype=np.int32); sy = np.ascontiguousarray(sy, dtype=np.int32)
    tx = np.ascontiguousarray(tx, dtype=np.int32); ty = np.ascontiguousarray(ty, dtype=np.int32)
    smass = np.ascontiguousarray(smass, dtype=np.float64); tmass = np.ascontiguousarray(tmass, dtype=np.float64)
    cost = C.c_double(0.0); iters = C.c_int64(0)
    st = _lib().stm_oracle_emd(sx.ctypes.data, sy.ctypes.data, smass.ctypes.data, len(sx),
                               tx.ctypes.data, ty.ctypes.data, tmass.ctypes.data, len(tx),
                               int(max_iter), int(block_size), int(cand_threads), C.byref(cost), C.byref(iters))
    return cost.value, st, iters.value


def emd_multiscale(sx, sy, smass, tx, ty, tmass, max_iter=200000, block_size=0, star=False, cand_threads=0):
    """(cost, status, total pivots, pivots per level [0] = finest) with the multiscale start basis the CUDA kernel uses
    by default (see emd_oracle.c); ``max_iter`` bounds the pivots of the finest level, i.e. of the original problem."""
    sx = np.ascontiguousarray(sx, dtype=np.int32); sy = np.ascontiguousarray(sy, dtype=np.int32)
    tx = np.ascontiguousarray(tx, dtype=np.int32); ty = np.ascontiguousarray(ty, dtype=np.int32)
    smass = np.ascontiguousarray(smass, dtype=np.float64); tmass = np.ascontiguousarray(tmass, dtype=np.float64)
    cost = C.c_double(0.0); iters = C.c_int64(0)
    lev = np.zeros(32, dtype=np.int64)
    st = _lib().stm_oracle_emd_ms(sx.ctypes.data, sy.ctypes.data, smass.ctypes.data, len(sx),
                                  tx.ctypes.data, ty.ctypes.data, tmass.ctypes.data, len(tx),
                                  int(max_iter), int(block_size), 1 if star else 0, int(cand_threads), C.byref(cost), C.byref(iters),
                                  lev.ctypes.data)
    return cost.value, st, iters.value, lev


def emd_linprog(sx, sy, smass, tx, ty, tmass):
    """Exact transport LP with HiGHS: min <G, M> s.t. G 1 = a, G^T 1 = b, G >= 0 (what ot.emd2 returns)."""
    from scipy import sparse
    from scipy.optimize import linprog
    n, m = len(sx), len(tx)
    a = np.asarray(smass, dtype=np.float64) / np.sum(smass)
    b = np.asarray(tmass, dtype=np.float64) / np.sum(tmass)
    M = (np.asarray(sx, dtype=np.float64)[:, None] - np.asarray(tx, dtype=np.float64)[None]) ** 2 + \
        (np.asarray(sy, dtype=np.float64)[:, None] - np.asarray(ty, dtype=np.float64)[None]) ** 2
    rows = sparse.kron(sparse.eye(n), np.ones((1, m)))
    cols = sparse.kron(np.ones((1, n)), sparse.eye(m))
    A = sparse.vstack([rows, cols]).tocsr()
    res = linprog(M.ravel(), A_eq=A, b_eq=np.concatenate([a, b]), bounds=(0, None), method="highs")
    if res.status != 0:
        raise RuntimeError(res.message)
    return float(res.fun)


def calculate_ot_distance(source_csr, target_csr, max_iter=200000):
    """distance.py:191-202 with the network-simplex oracle standing in for ot.emd2."""
    s = csr_to_problem(source_csr); t = csr_to_problem(target_csr)
    return emd_network_simplex(*s, *t, max_iter=max_iter)[0]
