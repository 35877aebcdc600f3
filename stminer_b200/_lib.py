"""ctypes binding of ``libstminer_b200.so`` (the C ABI declared in ``include/stminer_b200.h``).

There is NO CPU fallback: if the library cannot be loaded, or no CUDA device is present when a
compute entry point is called, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

from . import build as _build

_lock = threading.Lock()
_lib = None

c_i32p = C.c_void_p
_P = C.c_void_p

STM_GENE_OK = 0
STM_GENE_DROPPED = 1
STM_GENE_ILL_CONDITIONED = 2
STM_FIT_REMOVE_LOW_EXP_SPOTS = 1
STM_FIT_DEVICE_INIT = 2
STM_FIT_COMPACT_INPUT = 4
STM_FIT_NO_CLUSTERS = 8
STM_EMD_STAR_START = 1
STM_EMD_NO_CANDIDATES = 2
STM_FIT_N_BUCKETS = 7     # cluster x16 / x8 / x4 / x2, then one CTA per gene: large / medium / small


def fit_flags(remove_low_exp_spots=False, device_init=False, lloyd_iters=0, kpp_candidates=0, seed_cap=0,
              compact=False, no_clusters=False):
    """Compose the ``flags`` word of ``stm_gmm_fit_batched`` / ``stm_gmm_fit_plan`` (0 = library default)."""
    f = (STM_FIT_REMOVE_LOW_EXP_SPOTS if remove_low_exp_spots else 0) | (STM_FIT_DEVICE_INIT if device_init else 0)
    f |= STM_FIT_COMPACT_INPUT if compact else 0
    f |= STM_FIT_NO_CLUSTERS if no_clusters else 0
    f |= (int(lloyd_iters) & 0xff) << 8
    del kpp_candidates   # retired: the seeding is plain D^2 sampling by exponential race
    f |= ((int(seed_cap) + 255) // 256 & 0xff) << 20
    return f

# every symbol include/stminer_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "stm_version": (C.c_int, []),
    "stm_last_error_string": (C.c_char_p, []),
    "stm_gmm_fit_batched": (C.c_int, [
        _P, _P, _P, _P, _P, C.c_int32,                 # col_ptr,row_idx,val,spot_x,spot_y,n_spots
        C.c_int32, C.c_int32, _P, _P,                  # G,K,gene_order,bucket_counts_host
        _P, _P, _P,                                    # init_w,init_mu,init_cov
        C.c_double, C.c_double, C.c_int32, C.c_uint32, C.c_uint64,  # reg_covar,tol,max_iter,flags,seed
        _P, _P, _P, _P, _P, _P, _P, _P,                # out_w,out_mu,out_cov,n_iter,converged,lb,status,stream
    ]),
    "stm_gmm_fit_plan": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_uint32, _P, _P]),
    "stm_gmm_dist_pairs": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int64, C.c_int64, _P, _P]),
    "stm_lsap_batched": (C.c_int, [_P, C.c_int32, C.c_int64, _P, _P]),
    "stm_gmm_dist_fill_square": (C.c_int, [_P, C.c_int32, _P, _P]),
    "stm_gmm_dist_one_vs_all": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, _P, _P]),
    "stm_emd_workspace_bytes": (C.c_int64, [C.c_int32, C.c_int64, C.c_int32]),
    "stm_emd_spots_batched": (C.c_int, [_P, _P, _P, C.c_int32, _P, _P, _P, _P, C.c_int32, C.c_int64, _P,
                                        C.c_int64, C.c_int32, C.c_uint32, _P, C.c_int64, C.c_int32, _P, _P, _P, _P]),
    "stm_emd_spots_batched_wide": (C.c_int, [_P, _P, _P, C.c_int32, _P, _P, _P, _P, C.c_int32, C.c_int64, _P,
                                             C.c_int64, C.c_int32, C.c_uint32, _P, C.c_int64, C.c_int32, _P, _P, _P, _P]),
    "stm_emd_smem_node_capacity": (C.c_int32, [C.c_int32]),
    "stm_emd_global_node_capacity": (C.c_int32, []),
    "stm_emd_pricing_threads": (C.c_int32, [C.c_int32, C.c_int64, C.c_int32]),
    "stm_stage_blocks": (C.c_int32, [C.c_int32]),
    "stm_stage_workspace_bytes": (C.c_int64, [C.c_int32, C.c_int32]),
    "stm_stage_csr_columns": (C.c_int, [_P, _P, _P, C.c_int32, _P, C.c_int32, _P, C.c_uint32, _P, C.c_int64, C.c_int64,
                                        _P, _P, _P, _P]),
    "stm_pattern_vote": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _P, C.c_int32, _P, _P, _P, _P]),
    "stm_kmeans_workspace_bytes": (C.c_int64, [C.c_int32, C.c_int32]),
    "stm_kmeans_lloyd": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, _P, C.c_int32, C.c_int32, C.c_double, _P, C.c_int64,
                                   _P, _P, _P, _P, _P, _P]),
    "stm_mds_workspace_bytes": (C.c_int64, [C.c_int32, C.c_int32]),
    "stm_mds_smacof": (C.c_int, [_P, C.c_int32, C.c_int32, _P, C.c_int32, C.c_double, C.c_int32, _P, C.c_int64,
                                 _P, _P, _P, _P]),
    "stm_bench_fma": (C.c_int, [C.c_int32, C.c_int32, _P, _P]),
}


class StmError(RuntimeError):
    """A C-ABI call returned a negative status."""


def lib_path() -> str:
    """The in-tree library; STM_LIB_PATH selects a tuning variant built by ``build.build(defines=..., out=...)``."""
    return os.environ.get("STM_LIB_PATH") or _build.LIB_PATH


def load():
    """Load (building in-tree first if the sources are newer and nvcc is present) the CUDA library."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = lib_path()
        stale = path == _build.LIB_PATH and _build.needs_build() and _build.have_nvcc()
        if not os.path.exists(path) or stale or os.environ.get("STM_REBUILD") == "1":
            _build.build(force=os.environ.get("STM_REBUILD") == "1")
        if not os.path.exists(path):
            raise RuntimeError(
                "stminer_b200: CUDA library %s is missing and could not be built; there is no CPU fallback" % path)
        lib = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().stm_last_error_string()
        raise StmError("%s failed with status %d: %s" % (what, status, (msg or b"").decode()))


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("stminer_b200 requires a CUDA device (sm_100a); there is no CPU fallback")
    return torch


def ptr(t) -> int:
    """Device pointer of a torch tensor (None -> NULL)."""
    return 0 if t is None else t.data_ptr()


def current_stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
