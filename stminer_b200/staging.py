"""Host staging of the expression matrix into the hot path's HBM layout.

The reference slices ``adata[:, gene]`` and builds a dense int16 grid per gene
(STMiner/Algorithm/AlgUtils.py:26-36, O(nnz_total) per gene).  Here ONE pass converts the whole
matrix into CSC over DISTINCT spots — ``col_ptr / row_idx / val`` plus ``spot_x / spot_y`` — which is
uploaded once and stays resident for every stage (fit, distances, patterns).

Semantics kept from the reference:
  * int16 truncation toward zero of every stored value (``coo_matrix(..., dtype=np.int16)``, AlgUtils.py:35),
  * rows sharing an (x, y) are summed after truncation, with int16 wrap-around (``todense()``, AlgUtils.py:15),
  * only cells > 0 reach the mixture fit (``array_to_list``, distribution.py:369-373).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
from scipy import sparse


def _wrap_i16(a: np.ndarray) -> np.ndarray:
    return ((a.astype(np.int64) + 32768) % 65536 - 32768).astype(np.int32)


def _wrap_i16_float(a: np.ndarray) -> np.ndarray:
    """int16 wrap of integer-valued floats, as floats; one min/max pass when nothing is out of range (the usual case)."""
    if len(a) == 0 or (a.min() >= -32768.0 and a.max() <= 32767.0):
        return a
    return _wrap_i16(a).astype(np.float64)


@dataclass
class SpotMatrix:
    """CSC expression matrix over distinct spots (host copy, numpy)."""

    col_ptr: np.ndarray   # int64 [G+1]
    row_idx: np.ndarray   # int32 [nnz]
    val: np.ndarray       # float32 [nnz]
    spot_x: np.ndarray    # int32 [n_spots]
    spot_y: np.ndarray    # int32 [n_spots]
    genes: List[str]

    @property
    def n_genes(self) -> int:
        return len(self.col_ptr) - 1

    @property
    def n_spots(self) -> int:
        return len(self.spot_x)

    @property
    def nnz(self) -> int:
        return int(self.col_ptr[-1])

    def nnz_per_gene(self) -> np.ndarray:
        return np.diff(self.col_ptr)

    def lpt_order(self) -> np.ndarray:
        """Genes by decreasing nnz (longest processing time first)."""
        return np.argsort(-self.nnz_per_gene(), kind="stable").astype(np.int32)

    def gene_spots(self, g: int):
        """(xy[nnz_g, 2] int64 sorted row-major, counts[nnz_g] int64) of gene g, as the oracle wants them."""
        a, b = int(self.col_ptr[g]), int(self.col_ptr[g + 1])
        r = self.row_idx[a:b].astype(np.int64)
        w = np.trunc(self.val[a:b]).astype(np.int64)
        keep = w > 0
        xy = np.column_stack([self.spot_x[r][keep], self.spot_y[r][keep]]).astype(np.int64)
        w = w[keep]
        order = np.lexsort((xy[:, 1], xy[:, 0]))
        return xy[order], w[order]

    @property
    def is_compact(self) -> bool:
        """True when ``row_idx`` is uint16 and ``val`` int16 (the 4 B / value wire format)."""
        return self.row_idx.dtype == np.uint16 and self.val.dtype == np.int16

    def compact(self) -> "SpotMatrix":
        """4 B / value encoding — uint16 spot index + int16 value — for matrices with <= 65,536 spots whose values
        are already int16 integers (what ``from_matrix`` produces).  Halves the host->device transfer; the
        kernel reads it with ``STM_FIT_COMPACT_INPUT``.  Returns ``self`` when not applicable."""
        if self.is_compact:
            return self
        if self.n_spots > 65536 or self.nnz == 0:
            return self
        v = self.val
        if not (np.all(v == np.trunc(v)) and v.min() >= -32768 and v.max() <= 32767):
            return self
        return SpotMatrix(self.col_ptr, self.row_idx.astype(np.uint16), v.astype(np.int16), self.spot_x, self.spot_y,
                          list(self.genes))

    def pin_memory(self) -> "SpotMatrix":
        """Copy whose arrays live in page-locked host memory, so uploads are true async DMA."""
        import torch
        def pin(a):
            a = np.ascontiguousarray(a)
            if a.dtype == np.uint16:      # torch has no uint16 arithmetic; the bytes are what travels
                return torch.from_numpy(a.view(np.int16)).pin_memory()
            return torch.from_numpy(a).pin_memory()
        pins = [pin(a) for a in (self.col_ptr, self.row_idx, self.val, self.spot_x, self.spot_y)]
        arrs = [t.numpy() for t in pins]
        if self.row_idx.dtype == np.uint16:
            arrs[1] = arrs[1].view(np.uint16)
        out = SpotMatrix(*arrs, list(self.genes))
        out._pins = pins
        return out

    def slice_genes(self, g0: int, g1: int) -> "SpotMatrix":
        """Contiguous gene range [g0, g1) as views (no copy of the values; page-locked storage stays page-locked)."""
        a, b = int(self.col_ptr[g0]), int(self.col_ptr[g1])
        out = SpotMatrix((self.col_ptr[g0:g1 + 1] - self.col_ptr[g0]).astype(np.int64), self.row_idx[a:b],
                         self.val[a:b], self.spot_x, self.spot_y, self.genes[g0:g1])
        pins = getattr(self, "_pins", None)
        if pins is not None:
            import torch
            out._pins = [torch.from_numpy(out.col_ptr), pins[1][a:b], pins[2][a:b], pins[3], pins[4]]
        return out

    def nnz_balanced_chunks(self, n_chunks: int, fractions: Optional[Sequence[float]] = None):
        """Gene boundaries of ``n_chunks`` contiguous ranges with roughly equal numbers of stored values, or — with
        ``fractions`` (increasing, ending at 1) — ranges ending at those cumulative fractions of the stored values."""
        G = self.n_genes
        n_chunks = max(1, min(int(n_chunks), G))
        targets = np.linspace(0, self.nnz, n_chunks + 1)
        if fractions is not None:
            targets = np.concatenate([[0.0], np.asarray(fractions, dtype=np.float64)]) * self.nnz
        cuts = np.searchsorted(self.col_ptr, targets, side="left")
        cuts[0], cuts[-1] = 0, G
        return sorted(set(int(c) for c in cuts))

    def select(self, idx: Sequence[int]) -> "SpotMatrix":
        idx = np.asarray(idx, dtype=np.int64)
        lens = np.diff(self.col_ptr)[idx]
        col_ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        take = np.concatenate([np.arange(self.col_ptr[g], self.col_ptr[g + 1]) for g in idx]) if len(idx) else \
            np.zeros(0, dtype=np.int64)
        return SpotMatrix(col_ptr, self.row_idx[take], self.val[take], self.spot_x, self.spot_y,
                          [self.genes[g] for g in idx])

    def preprocess_array(self, binary: bool = False, cut: bool = False, threshold: float = 95,
                         remove_low_exp_spots: bool = False) -> "SpotMatrix":
        """``preprocess_array`` (distribution.py:133-145) on the staged columns: the optional ``binary`` (cells above the
        ``threshold`` percentile become 1, the rest 0) and ``cut`` (cells below the percentile become 0) transforms.
        The percentile is taken over the WHOLE dense count grid of ``get_exp_array`` — shape (max x + 1, max y + 1),
        empty cells included (AlgUtils.py:35,15) — after the optional ``remove_low_exp_spots`` step (AlgUtils.py:16-17),
        which is therefore applied here too: fit the returned matrix WITHOUT the remove-low flag."""
        if not (binary or cut):
            return self
        if binary and (threshold > 100 or threshold < 0):
            import logging
            logging.getLogger(__name__).warning("the threshold is illegal, the value in [0, 100] is accepted.")
            threshold = 100
        n_cells = (int(self.spot_x.max()) + 1) * (int(self.spot_y.max()) + 1)
        val = np.empty(len(self.val), dtype=np.float32)
        for g in range(self.n_genes):
            a, b = int(self.col_ptr[g]), int(self.col_ptr[g + 1])
            v = _wrap_i16(np.trunc(np.asarray(self.val[a:b], dtype=np.float64)).astype(np.int64)).astype(np.int64)
            if remove_low_exp_spots:
                nz = v[v != 0]
                mean = nz.mean() if len(nz) else 0.0
                if mean < 0:
                    raise NotImplementedError("negative mean expression: empty cells would become positive")
                v = np.round(np.maximum(v - mean, 0)).astype(np.int64)
            dense = np.zeros(max(n_cells, b - a), dtype=np.int64)
            dense[:b - a] = v
            thr = np.percentile(dense, threshold)
            if thr < 0:
                raise NotImplementedError("negative percentile: empty cells would pass the threshold")
            val[a:b] = (v > thr).astype(np.float32) if binary else np.where(v < thr, 0, v).astype(np.float32)
        return SpotMatrix(self.col_ptr, self.row_idx, val, self.spot_x, self.spot_y, self.genes)

    # ------------------------------------------------------------------ constructors
    @classmethod
    def from_matrix(cls, X, x, y, genes: Sequence[str], columns: Optional[Sequence[int]] = None,
                    truncate_int16: bool = True) -> "SpotMatrix":
        """Build from an (n_obs, n_var) matrix (scipy sparse or dense) and per-row integer coordinates."""
        x = np.asarray(x).astype(np.int64).ravel()
        y = np.asarray(y).astype(np.int64).ravel()
        if sparse.issparse(X):
            M = X.tocsc()
        else:
            M = sparse.csc_matrix(np.asarray(X))
        if columns is not None:
            M = M[:, np.asarray(columns, dtype=np.int64)]
            genes = [genes[c] for c in columns]
        if truncate_int16:
            # AlgUtils.py:35; trunc copies the data, the structure arrays are copied because eliminate_zeros below works in
            # place and tocsc() hands back the caller's own arrays when X already is CSC
            M = sparse.csc_matrix((_wrap_i16_float(np.trunc(M.data.astype(np.float64, copy=False))),
                                   M.indices.copy(), M.indptr.copy()), shape=M.shape)
        else:
            M = M.astype(np.float64, copy=True)
        # distinct spots, row-major (x, y) order
        ymul = int(y.max()) + 1 if len(y) else 1
        key = x * ymul + y
        uniq, inverse = np.unique(key, return_inverse=True)
        n_obs, n_spots = len(key), len(uniq)
        if n_spots != n_obs or not np.array_equal(inverse, np.arange(n_obs)):
            P = sparse.csr_matrix((np.ones(n_obs), (inverse, np.arange(n_obs))), shape=(n_spots, n_obs))
            M = (P @ M).tocsc()                                              # duplicates summed (todense, AlgUtils.py:15)
            if truncate_int16:
                M.data = _wrap_i16_float(M.data)
        M.eliminate_zeros()
        M.sort_indices()
        return cls(M.indptr.astype(np.int64), M.indices.astype(np.int32), M.data.astype(np.float32),
                   (uniq // ymul).astype(np.int32), (uniq % ymul).astype(np.int32), list(genes))


class DeviceSpotMatrix:
    """SpotMatrix resident in HBM (torch tensors are only the buffer owners)."""

    def __init__(self, sm: SpotMatrix, device=None, non_blocking: bool = True, spots=None):
        import torch
        from . import _lib
        _lib.require_cuda()
        self.host = sm
        self.device = torch.device(device if device is not None else "cuda")
        self._plans = {}

        pins = getattr(sm, "_pins", None)   # page-locked torch tensors behind the arrays (SpotMatrix.pin_memory)

        def up(a, k=None):
            # async DMA straight from the page-locked tensor when there is one, staged copy otherwise
            if pins is not None and k is not None:
                t = pins[k]
            else:
                a = np.ascontiguousarray(a)
                t = torch.from_numpy(a.view(np.int16) if a.dtype == np.uint16 else a)
            return t.to(self.device, non_blocking=non_blocking)

        self.compact = sm.is_compact

        self.col_ptr = up(sm.col_ptr, 0)
        self.row_idx = up(sm.row_idx, 1)
        self.val = up(sm.val, 2)
        # the spot table is shared by the chunks of a pipelined upload
        self.spot_x, self.spot_y = spots if spots is not None else (up(sm.spot_x, 3), up(sm.spot_y, 4))
        self._up = up
        self.h2d_bytes = sum(a.nbytes for a in (sm.col_ptr, sm.row_idx, sm.val))
        self.h2d_bytes += 0 if spots is not None else sm.spot_x.nbytes + sm.spot_y.nbytes

    @classmethod
    def from_ranges(cls, sm: "SpotMatrix", ranges, device=None) -> "DeviceSpotMatrix":
        """The genes of the (g0, g1) ``ranges`` of ``sm``, in that order, as one device matrix: every range is copied
        straight from the host arrays (async DMA when ``sm`` is page-locked) to its place in the device buffers — no
        host-side gather.  What a rank of a sharded fit uploads."""
        import torch
        from . import _lib
        _lib.require_cuda()
        dev = torch.device(device if device is not None else "cuda")
        lens = [np.diff(sm.col_ptr[a:b + 1]) for a, b in ranges]
        col_ptr = np.concatenate([[0], np.cumsum(np.concatenate(lens) if lens else np.zeros(0, dtype=np.int64))]).astype(np.int64)
        nnz = int(col_ptr[-1])
        pins = getattr(sm, "_pins", None)
        as_t = lambda a: torch.from_numpy(a.view(np.int16) if a.dtype == np.uint16 else a)
        src_row = pins[1] if pins is not None else as_t(np.ascontiguousarray(sm.row_idx))
        src_val = pins[2] if pins is not None else as_t(np.ascontiguousarray(sm.val))
        with torch.cuda.device(dev):
            row_idx = torch.empty(max(nnz, 1), dtype=src_row.dtype, device=dev)
            val = torch.empty(max(nnz, 1), dtype=src_val.dtype, device=dev)
            at = 0
            for a, b in ranges:
                e0, e1 = int(sm.col_ptr[a]), int(sm.col_ptr[b])
                row_idx[at:at + e1 - e0].copy_(src_row[e0:e1], non_blocking=True)
                val[at:at + e1 - e0].copy_(src_val[e0:e1], non_blocking=True)
                at += e1 - e0
        genes = [g for a, b in ranges for g in sm.genes[a:b]]
        out = cls.__new__(cls)
        out.host = _StagedHost(col_ptr, genes, sm.spot_x, sm.spot_y, sm.is_compact)
        out.device, out.compact, out._plans = dev, sm.is_compact, {}
        up = lambda a, k=None: torch.from_numpy(np.ascontiguousarray(a)).to(dev, non_blocking=True)
        out.col_ptr, out.row_idx, out.val = up(col_ptr), row_idx, val
        out.spot_x, out.spot_y = up(sm.spot_x), up(sm.spot_y)
        out._up = up
        out.h2d_bytes = int(col_ptr.nbytes + nnz * (src_row.element_size() + src_val.element_size()) + 2 * sm.spot_x.nbytes)
        return out

    def plan(self, K: int, flags: int = 0):
        """(gene_order CUDA tensor, bucket_counts host int32[STM_FIT_N_BUCKETS]) from ``stm_gmm_fit_plan`` for K components."""
        from . import _lib
        key = (K, flags & ~5)   # (remove_low / compact do not move boundaries)   # the device-init fields move the bucket boundaries
        if key not in self._plans:
            lib = _lib.load()
            G = self.host.n_genes
            col = np.ascontiguousarray(self.host.col_ptr, dtype=np.int64)
            order = np.empty(max(G, 1), dtype=np.int32)
            buckets = np.zeros(_lib.STM_FIT_N_BUCKETS, dtype=np.int32)
            _lib.check(lib.stm_gmm_fit_plan(col.ctypes.data, G, int(K), int(flags), order.ctypes.data,
                                            buckets.ctypes.data), "stm_gmm_fit_plan")
            self._plans[key] = (self._up(order[:G]), buckets)
            self.h2d_bytes += order[:G].nbytes
        return self._plans[key]

    @property
    def n_genes(self) -> int:
        return self.host.n_genes

    @property
    def n_spots(self) -> int:
        return self.host.n_spots


class StreamedDeviceSpotMatrix:
    """Device buffers for the WHOLE matrix allocated once (on the caller's stream), filled gene range by gene range
    on a copy stream.  ``gene_view(g0, g1)`` is what ``fit_spot_matrix`` needs to fit that range: the column
    pointers of the range keep their absolute offsets into the full ``row_idx`` / ``val`` buffers, so no re-basing
    and no per-chunk allocation is involved."""

    def __init__(self, sm: SpotMatrix, device):
        import torch
        from . import _lib
        _lib.require_cuda()
        self.host, self.device, self.compact = sm, torch.device(device), sm.is_compact
        pins = getattr(sm, "_pins", None)
        as_t = lambda a: torch.from_numpy(a.view(np.int16) if a.dtype == np.uint16 else a)
        self._src = pins if pins is not None else [as_t(np.ascontiguousarray(a)) for a in
                                                   (sm.col_ptr, sm.row_idx, sm.val, sm.spot_x, sm.spot_y)]
        empty = lambda t: torch.empty(t.shape, dtype=t.dtype, device=self.device)
        self.col_ptr, self.row_idx, self.val = empty(self._src[0]), empty(self._src[1]), empty(self._src[2])
        self.spot_x = self._src[3].to(self.device, non_blocking=True)
        self.spot_y = self._src[4].to(self.device, non_blocking=True)
        self.col_ptr.copy_(self._src[0], non_blocking=True)
        self.h2d_bytes = sum(int(t.numel() * t.element_size()) for t in self._src)

    def upload_genes(self, g0: int, g1: int):
        """Enqueue (on the current stream) the copy of the stored values of genes [g0, g1)."""
        a, b = int(self.host.col_ptr[g0]), int(self.host.col_ptr[g1])
        self.row_idx[a:b].copy_(self._src[1][a:b], non_blocking=True)
        self.val[a:b].copy_(self._src[2][a:b], non_blocking=True)

    def gene_view(self, g0: int, g1: int) -> "DeviceSpotMatrix":
        view = DeviceSpotMatrix.__new__(DeviceSpotMatrix)
        view.host = _GeneRange(self.host, g0, g1)
        view.device, view.compact, view._plans = self.device, self.compact, {}
        view.col_ptr, view.row_idx, view.val = self.col_ptr[g0:g1 + 1], self.row_idx, self.val
        view.spot_x, view.spot_y = self.spot_x, self.spot_y
        import torch
        view._up = lambda a, k=None: torch.from_numpy(np.ascontiguousarray(a)).to(self.device, non_blocking=True)
        view.h2d_bytes = 0
        return view


class _GeneRange:
    """The part of a SpotMatrix the launch plan needs for genes [g0, g1): counts, not values."""

    def __init__(self, sm: SpotMatrix, g0: int, g1: int):
        self.col_ptr = sm.col_ptr[g0:g1 + 1] - sm.col_ptr[g0]
        self.genes = sm.genes[g0:g1]
        self.spot_x = sm.spot_x

    @property
    def n_genes(self) -> int:
        return len(self.col_ptr) - 1

    @property
    def n_spots(self) -> int:
        return len(self.spot_x)


class _StagedHost:
    """Host-side description of a matrix staged ON the device: what the launch plan and the result assembly need
    (column pointers, gene names, the spot table) — the stored values themselves never existed on the host."""

    def __init__(self, col_ptr, genes, spot_x, spot_y, compact):
        self.col_ptr = np.ascontiguousarray(col_ptr, dtype=np.int64)
        self.genes = list(genes)
        self.spot_x, self.spot_y = spot_x, spot_y
        self.is_compact = bool(compact)

    @property
    def n_genes(self) -> int:
        return len(self.col_ptr) - 1

    @property
    def n_spots(self) -> int:
        return len(self.spot_x)

    @property
    def nnz(self) -> int:
        return int(self.col_ptr[-1])

    def nnz_per_gene(self) -> np.ndarray:
        return np.diff(self.col_ptr)


def device_staging_supported(X, x, y, columns) -> bool:
    """The device staging takes a canonical CSR matrix of float32 / small-integer values whose rows sit on DISTINCT (x, y)
    spots, and a gene selection without repeats; everything else goes through ``SpotMatrix.from_matrix`` on the host."""
    if not sparse.issparse(X) or X.format != "csr":
        return False
    if X.dtype not in (np.float32, np.int16, np.int32, np.uint8, np.uint16, np.int8):
        return False
    if X.dtype == np.int32 and X.nnz and (np.abs(X.data).max() >= 2 ** 24):
        return False
    if columns is not None and len(set(int(c) for c in columns)) != len(columns):
        return False
    key = np.asarray(x).astype(np.int64).ravel() * (int(np.max(y)) + 1 if len(y) else 1) + np.asarray(y).astype(np.int64).ravel()
    return len(np.unique(key)) == len(key)


def stage_csr_on_device(X, x, y, genes, columns=None, device=None) -> DeviceSpotMatrix:
    """``SpotMatrix.from_matrix`` + upload, done on the device (csrc/staging.cu): the spot-by-gene CSR arrays are
    uploaded as they are and ONE stable transposition writes the selected genes' columns in the hot path's layout —
    int16 truncation / wrap (AlgUtils.py:35), zeros dropped, entries by increasing spot index, the 4 B / value encoding
    when there are at most 65,536 spots.  Same arrays as the host staging produces (tests compare them).  The caller
    checks ``device_staging_supported`` first."""
    import torch
    from . import _lib
    _lib.require_cuda()
    lib = _lib.load()
    dev = torch.device(device if device is not None else "cuda")
    if not X.has_canonical_format:
        X = X.copy()
        X.sum_duplicates()
    x = np.asarray(x).astype(np.int64).ravel()
    y = np.asarray(y).astype(np.int64).ravel()
    n_obs, n_var = X.shape
    ymul = int(y.max()) + 1 if len(y) else 1
    key = x * ymul + y
    row_order = np.argsort(key, kind="stable").astype(np.int32)          # spot s (row-major (x, y) order) <- input row
    skey = key[row_order]
    if columns is None:
        columns = np.arange(n_var, dtype=np.int64)
    columns = np.asarray(columns, dtype=np.int64)
    n_out = len(columns)
    col_map = np.full(n_var, -1, dtype=np.int32)
    col_map[columns] = np.arange(n_out, dtype=np.int32)
    compact = n_obs <= 65536
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev, non_blocking=True)
    with torch.cuda.device(dev):
        d_indptr = up(X.indptr.astype(np.int64, copy=False))
        d_indices = up(X.indices.astype(np.int32, copy=False))
        d_data = up(X.data.astype(np.float32, copy=False))
        d_map, d_order = up(col_map), up(row_order)
        cap = int(X.nnz)
        row_idx = torch.empty(max(cap, 1), dtype=torch.int16 if compact else torch.int32, device=dev)
        val = torch.empty(max(cap, 1), dtype=torch.int16 if compact else torch.float32, device=dev)
        col_ptr = torch.empty(n_out + 1, dtype=torch.int64, device=dev)
        ws_bytes = int(lib.stm_stage_workspace_bytes(n_obs, n_out))
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
        st = lib.stm_stage_csr_columns(_lib.ptr(d_indptr), _lib.ptr(d_indices), _lib.ptr(d_data), n_obs, _lib.ptr(d_map), n_out,
                                       _lib.ptr(d_order), 1 if compact else 0, _lib.ptr(ws), ws_bytes, cap,
                                       _lib.ptr(col_ptr), _lib.ptr(row_idx), _lib.ptr(val), _lib.current_stream_ptr())
        _lib.check(st, "stm_stage_csr_columns")
        host_col_ptr = col_ptr.cpu().numpy()                             # the launch plan needs the column lengths
    spot_x = (skey // ymul).astype(np.int32); spot_y = (skey % ymul).astype(np.int32)
    dsm = DeviceSpotMatrix.__new__(DeviceSpotMatrix)
    dsm.host = _StagedHost(host_col_ptr, [genes[c] for c in columns], spot_x, spot_y, compact)
    dsm.device, dsm.compact, dsm._plans = dev, compact, {}
    dsm.col_ptr, dsm.row_idx, dsm.val = col_ptr, row_idx, val
    dsm.spot_x, dsm.spot_y = up(spot_x), up(spot_y)
    dsm._up = lambda a, k=None: torch.from_numpy(np.ascontiguousarray(a)).to(dev, non_blocking=True)
    dsm.h2d_bytes = int(X.indptr.nbytes + X.nnz * 8 + col_map.nbytes + row_order.nbytes + 2 * spot_x.nbytes)
    dsm._keepalive = (d_indptr, d_indices, d_data, d_map, d_order, ws)
    return dsm


def pattern_vote_on_device(X, n_var: int, col_label: np.ndarray, n_labels: int, device=None):
    """``stm_pattern_vote`` on a canonical CSR matrix: (total[n_labels, n_obs] float64, votes[n_labels, n_obs] int32) — the
    accumulation of ``_genes_to_pattern`` (SPFinder.py:459-489) for every label in one pass (csrc/staging.cu)."""
    import torch
    from . import _lib
    _lib.require_cuda()
    lib = _lib.load()
    dev = torch.device(device if device is not None else "cuda")
    if not X.has_canonical_format:
        X = X.copy()
        X.sum_duplicates()
    n_obs = X.shape[0]
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev, non_blocking=True)
    with torch.cuda.device(dev):
        d_indptr, d_indices = up(X.indptr.astype(np.int64, copy=False)), up(X.indices.astype(np.int32, copy=False))
        d_data, d_label = up(X.data.astype(np.float32, copy=False)), up(np.asarray(col_label, dtype=np.int32))
        gsum = torch.empty(max(n_var, 1), dtype=torch.int64, device=dev)
        total = torch.empty((n_labels, n_obs), dtype=torch.float64, device=dev)
        votes = torch.empty((n_labels, n_obs), dtype=torch.int32, device=dev)
        st = lib.stm_pattern_vote(_lib.ptr(d_indptr), _lib.ptr(d_indices), _lib.ptr(d_data), n_obs, n_var, _lib.ptr(d_label),
                                  int(n_labels), _lib.ptr(gsum), _lib.ptr(total), _lib.ptr(votes), _lib.current_stream_ptr())
        _lib.check(st, "stm_pattern_vote")
        return total.cpu().numpy(), votes.cpu().numpy()
