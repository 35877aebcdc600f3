"""CPU: the C-ABI library loads, exports every symbol include/*.h declares, and its host-side entry
points (version, argument validation, launch plan) behave — no compute calls, no GPU."""
import ctypes
import glob
import os
import re

import numpy as np

from stminer_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(stm_[a-z0-9_]+)\s*\(", src))
    return names


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    decl = declared_symbols()
    assert {"stm_gmm_fit_batched", "stm_gmm_dist_pairs", "stm_version"} <= decl
    for name in decl:
        assert hasattr(lib, name), name
    assert decl == set(_lib.SIGNATURES), "ctypes signature table and header disagree"


def test_version_and_error_string():
    lib = _lib.load()
    assert lib.stm_version() >= 100
    st = lib.stm_gmm_dist_pairs(None, None, None, 4, 0, 0, 1, None, None)   # K = 0
    assert st == -1
    assert b"K" in lib.stm_last_error_string()
    st = lib.stm_gmm_dist_pairs(None, None, None, 4, 3, 0, 100, None, None)  # range outside triangle
    assert st == -1
    # G == 0 is a no-op, not an error
    assert lib.stm_gmm_fit_batched(*([None] * 5), 0, 0, 5, None, None, None, None, None, 1e-3, 1e-3, 100, 0, 0,
                                   *([None] * 8)) == 0
    st = lib.stm_gmm_fit_batched(*([None] * 5), 10, 3, 5, None, None, None, None, None, 1e-3, 1e-3, 100, 0, 0,
                                 *([None] * 8))
    assert st == -1 and b"null" in lib.stm_last_error_string()


def test_fit_plan_orders_by_nnz_and_buckets():
    lib = _lib.load()
    nnz = np.array([10, 30000, 500, 6000, 4000, 0, 12000, 100000], dtype=np.int64)
    col = np.concatenate([[0], np.cumsum(nnz)]).astype(np.int64)
    order = np.empty(len(nnz), dtype=np.int32)
    buckets = np.zeros(_lib.STM_FIT_N_BUCKETS, dtype=np.int32)
    assert lib.stm_gmm_fit_plan(col.ctypes.data, len(nnz), 20, 0, order.ctypes.data, buckets.ctypes.data) == 0
    assert (np.diff(nnz[order]) <= 0).all()
    assert sorted(order.tolist()) == list(range(len(nnz)))
    assert buckets.sum() == len(nnz)
    # 8-byte staged records: the small bucket (4 CTAs/SM) holds ~6.5k spots, the medium one (2 CTAs/SM) ~13k
    # 30,000 and 100,000 spots exceed one CTA's staging budget (~27.5k): clusters of 2 and 4 CTAs
    assert buckets.tolist() == [0, 0, 1, 1, 0, 1, 5]
    assert set(nnz[order[:2]].tolist()) == {30000, 100000} and nnz[order[2]] == 12000
    # with the device init a seed area is reserved, which lowers the boundaries (6000 no longer fits 4 CTAs/SM)
    assert lib.stm_gmm_fit_plan(col.ctypes.data, len(nnz), 20, _lib.fit_flags(device_init=True),
                                order.ctypes.data, buckets.ctypes.data) == 0
    assert buckets.tolist() == [0, 1, 0, 1, 1, 1, 4]
    assert lib.stm_gmm_fit_plan(col.ctypes.data, len(nnz), 20, _lib.fit_flags(device_init=True, no_clusters=True),
                                order.ctypes.data, buckets.ctypes.data) == 0
    assert buckets.tolist() == [0, 0, 0, 0, 3, 1, 4]
    # K > 40 has one kernel shape and no cluster buckets
    assert lib.stm_gmm_fit_plan(col.ctypes.data, len(nnz), 64, 0, order.ctypes.data, buckets.ctypes.data) == 0
    assert buckets[:4].sum() == 0 and buckets.sum() == len(nnz)


def test_product_code_never_touches_the_oracle_or_the_reference_tree():
    """The oracle is test infrastructure: nothing under stminer_b200/ (nor the library sources) may import, load or
    read it, and nothing that runs on the GPU box may read /root/reference."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bad = []
    for base, _, files in os.walk(os.path.join(root, "stminer_b200")):
        if "build" in os.path.basename(base):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(base, f), encoding="utf-8").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, re.M) or "libstm_oracle" in src or "/root/reference" in src:
                    bad.append(os.path.join(base, f))
    assert not bad, bad
    for f in ("bench.py", "__graft_entry__.py"):
        assert "/root/reference" not in open(os.path.join(root, f), encoding="utf-8").read(), f
