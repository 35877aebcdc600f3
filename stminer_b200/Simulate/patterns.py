"""Vectorised restatement of the reference's synthetic-data generator.

``Simulator.generate`` (STMiner/Simulate/Simulate.py:30-86) is a Python triple loop over cells; at the
benchmark sizes (40 k bins x 10 k genes) it is unusable, so the same per-replica semantics are applied
with array operations:

  * intensity scale ~ U(0.5, 2)                                   (Simulate.py:41-42)
  * every non-zero cell moves with probability ``offset_probability`` to a uniformly drawn cell of the
    clipped +-``offset_radius`` box, later writes overwriting earlier ones in row-major order (:44-55)
  * noise: ``gauss``  N(mean*arg, 1) clipped at 0 and added to every cell (simUtils.py:13-16),
           ``periodicity`` x0.1 on every ``arg``-th row/column (:19-29),
           ``undersampling`` Bernoulli mask (:6-10), ``uniform`` U(0,1)*arg added (:32-35).

The generator draws from ``numpy.random.Generator`` (the reference mixes numpy's global RandomState and
``random.choice``), so streams differ but distributions match.  Output goes straight to the hot path's
CSC layout (values truncated to int16 as AlgUtils.py:35 does before the fit).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import pandas as pd
from scipy import sparse

from ..staging import SpotMatrix

# name -> (rows, cols, hex_lattice, n_templates, replicas)
CONFIGS = {
    "T": (17, 17, False, 15, 20),          # stand-in for the reference test path (282 bins, ~300 genes)
    "V": (78, 128, True, 20, 100),         # Visium-shaped: 4,992 spots x 2,000 genes
    "S50": (200, 200, False, 100, 100),    # Stereo-seq bin50: 40,000 bins x 10,000 genes
    "R32": (32, 32, False, 27, 20),        # 1,024 bins x 540 genes: the smallest set with a top-500 SVG ranking
}


def make_templates(shape: Tuple[int, int], n_templates: int, rng: np.random.Generator) -> List[np.ndarray]:
    """Template expression maps: mixtures of 1-4 Gaussian blobs / stripes with Poisson counts."""
    R, C = shape
    rr, cc = np.meshgrid(np.arange(R, dtype=np.float64), np.arange(C, dtype=np.float64), indexing="ij")
    out = []
    for _ in range(n_templates):
        lam = np.zeros(shape)
        for _b in range(int(rng.integers(1, 5))):
            amp = rng.uniform(2.0, 6.0)
            if rng.random() < 0.25:   # stripe
                theta = rng.uniform(0, np.pi)
                off = rng.uniform(0.2, 0.8) * (R * np.cos(theta) ** 2 + C * np.sin(theta) ** 2)
                width = rng.uniform(0.02, 0.06) * max(R, C)
                d = (rr * np.cos(theta) + cc * np.sin(theta) - off) / width
                lam += amp * np.exp(-0.5 * d * d)
            else:                      # anisotropic blob
                cx, cy = rng.uniform(0.1, 0.9) * R, rng.uniform(0.1, 0.9) * C
                sx, sy = rng.uniform(0.04, 0.18) * R, rng.uniform(0.04, 0.18) * C
                rho = rng.uniform(-0.7, 0.7)
                dx, dy = (rr - cx) / sx, (cc - cy) / sy
                lam += amp * np.exp(-0.5 * (dx * dx - 2 * rho * dx * dy + dy * dy) / (1 - rho * rho))
        out.append(rng.poisson(lam).astype(np.float64))
    return out


def simulate_replicas(template: np.ndarray, count: int, rng: np.random.Generator, offset_radius: int = 2,
                      offset_probability: float = 0.5, noise_type: Optional[str] = "gauss",
                      noise_argument: float = 0.1) -> np.ndarray:
    """``count`` noisy replicas of one template, shape (count, R, C) float64 (Simulate.py:36-68)."""
    R, C = template.shape
    nz_r, nz_c = np.nonzero(template)            # row-major visiting order of np.ndenumerate
    vals = template[nz_r, nz_c]
    n = len(vals)
    out = np.zeros((count, R, C))
    for i in range(count):
        scale = rng.uniform(0.5, 2.0)
        move = rng.random(n) > (1.0 - offset_probability)
        lo_r = np.maximum(nz_r - offset_radius, 0); hi_r = np.minimum(nz_r + offset_radius + 1, R)
        lo_c = np.maximum(nz_c - offset_radius, 0); hi_c = np.minimum(nz_c + offset_radius + 1, C)
        new_r = np.where(move, rng.integers(lo_r, hi_r), nz_r)
        new_c = np.where(move, rng.integers(lo_c, hi_c), nz_c)
        sim = out[i]
        sim[new_r, new_c] = vals * scale           # repeated index: last write wins, as the sequential loop
        if noise_type == "gauss":
            noise = rng.normal(loc=sim.mean() * noise_argument, scale=1.0, size=sim.shape)
            noise[noise < 0] = 0
            sim += noise
        elif noise_type == "periodicity":
            k = int(noise_argument)
            sim[::k, :] *= 0.1
            sim[:, ::k] *= 0.1      # rows first, then columns overwrite to 0.1 exactly as simUtils.py:25-28
            sim[::k, ::k] *= 10.0   # crossings were scaled twice above; the reference scales them once
        elif noise_type == "undersampling":
            sim *= (rng.random(sim.shape) > noise_argument)
        elif noise_type == "uniform":
            sim += rng.random(sim.shape) * noise_argument
    return out


def _grid_spots(R: int, C: int, hex_lattice: bool):
    rr, cc = np.meshgrid(np.arange(R), np.arange(C), indexing="ij")
    mask = ((rr + cc) % 2 == 0) if hex_lattice else np.ones((R, C), dtype=bool)
    return mask, rr[mask].astype(np.int32), cc[mask].astype(np.int32)


def simulate_spot_matrix(config: str = "V", seed: int = 0, n_templates: Optional[int] = None,
                         replicas: Optional[int] = None, noise_type: Optional[str] = "gauss",
                         noise_argument: float = 0.1, truncate_int16: bool = True,
                         return_templates: bool = False):
    """Synthetic data set of a named BASELINE config directly in the hot path's CSC layout.

    Spots are the grid cells (the even-parity cells for the Visium hex lattice), x = row, y = column
    (Utils/utils.py:117-118).  Gene names follow the reference: ``gene_P{template}_N{replica}``.
    """
    R, C, hexl, nt, rep = CONFIGS[config]
    nt = nt if n_templates is None else n_templates
    rep = rep if replicas is None else replicas
    rng = np.random.default_rng(seed)
    mask, sx, sy = _grid_spots(R, C, hexl)
    templates = [t * mask for t in make_templates((R, C), nt, rng)]
    n_spots = int(mask.sum())
    blocks, genes = [], []
    for ti, t in enumerate(templates):
        sims = simulate_replicas(t, rep, rng, noise_type=noise_type, noise_argument=noise_argument)
        flat = sims[:, mask]                                        # (rep, n_spots)
        if truncate_int16:
            flat = np.trunc(flat)
        blocks.append(sparse.csc_matrix(flat.T.astype(np.float32)))  # (n_spots, rep)
        genes += ["gene_P%d_N%d" % (ti, i) for i in range(rep)]
    M = sparse.hstack(blocks, format="csc")
    M.eliminate_zeros(); M.sort_indices()
    sm = SpotMatrix(M.indptr.astype(np.int64), M.indices.astype(np.int32), M.data.astype(np.float32),
                    sx, sy, genes)
    assert sm.n_spots == n_spots
    return (sm, templates) if return_templates else sm


def simulate_long_genes(n_genes: int, seed: int = 0, shape: Tuple[int, int] = (1000, 1000), nnz_lo: float = 1e3,
                        nnz_hi: float = 5e5, nnz: Optional[Sequence[int]] = None) -> SpotMatrix:
    """Stereo-seq bin10-shaped genes (BASELINE configs[3], SURVEY.md §8d "S10"): a 1000 x 1000 full grid, per-gene
    distinct-spot counts log-uniform in [nnz_lo, nnz_hi] (or given by ``nnz``), counts 1 + Poisson(intensity).

    A gene's support is the ``nnz_g`` cells with the largest ``intensity x Exp(1)`` — a thinned sample of a blob
    mixture — so long genes cover most of the tissue and short ones only the blob cores.
    """
    R, C = shape
    rng = np.random.default_rng(seed)
    rr, cc = np.meshgrid(np.arange(R, dtype=np.float32), np.arange(C, dtype=np.float32), indexing="ij")
    targets = (np.asarray(nnz, dtype=np.int64) if nnz is not None else
               np.exp(rng.uniform(np.log(nnz_lo), np.log(nnz_hi), size=n_genes)).astype(np.int64))
    targets = np.minimum(targets, R * C)
    cols, rows, vals = [0], [], []
    for n in targets:
        lam = np.full((R, C), 0.05, dtype=np.float32)
        for _b in range(int(rng.integers(2, 6))):
            cx, cy = rng.uniform(0.1, 0.9) * R, rng.uniform(0.1, 0.9) * C
            sx, sy = rng.uniform(0.05, 0.25) * R, rng.uniform(0.05, 0.25) * C
            rho = rng.uniform(-0.7, 0.7)
            dx, dy = (rr - cx) / sx, (cc - cy) / sy
            lam += rng.uniform(1.0, 4.0) * np.exp(-0.5 * (dx * dx - 2 * rho * dx * dy + dy * dy) / (1 - rho * rho))
        score = (lam * rng.exponential(size=lam.shape).astype(np.float32)).ravel()
        idx = np.sort(np.argpartition(score, score.size - int(n))[score.size - int(n):])
        rows.append(idx.astype(np.int32))
        vals.append((1 + rng.poisson(lam.ravel()[idx])).astype(np.float32))
        cols.append(cols[-1] + len(idx))
    sx_all = (np.arange(R * C) // C).astype(np.int32); sy_all = (np.arange(R * C) % C).astype(np.int32)
    return SpotMatrix(np.asarray(cols, dtype=np.int64), np.concatenate(rows), np.concatenate(vals), sx_all, sy_all,
                      ["long_%d" % i for i in range(len(targets))])


def simulate_long_genes_device(n_genes: int, seed: int = 0, shape: Tuple[int, int] = (1000, 1000), nnz_lo: float = 1e3,
                               nnz_hi: float = 5e5, device=None):
    """``simulate_long_genes`` generated ON the device with torch (the full BASELINE configs[3] gene count — thousands of
    genes of up to 5e5 spots — takes minutes in numpy): same model (blob-mixture intensity, support = the ``nnz_g`` cells
    with the largest ``intensity x Exp(1)``, counts 1 + Poisson(intensity)), torch's generator instead of numpy's.
    Returns a ``DeviceSpotMatrix`` (8 B / value: a million spots do not fit the compact encoding)."""
    import torch
    from ..staging import DeviceSpotMatrix, _StagedHost
    dev = torch.device(device if device is not None else "cuda")
    R, C = shape
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed))
    rng = np.random.default_rng(seed)
    targets = np.minimum(np.exp(rng.uniform(np.log(nnz_lo), np.log(nnz_hi), size=n_genes)).astype(np.int64), R * C)
    rr = torch.arange(R, dtype=torch.float32, device=dev)[:, None]
    cc = torch.arange(C, dtype=torch.float32, device=dev)[None, :]
    col_ptr = np.concatenate([[0], np.cumsum(targets)]).astype(np.int64)
    row_idx = torch.empty(int(col_ptr[-1]), dtype=torch.int32, device=dev)
    val = torch.empty(int(col_ptr[-1]), dtype=torch.float32, device=dev)
    for g, n in enumerate(targets.tolist()):
        lam = torch.full((R, C), 0.05, dtype=torch.float32, device=dev)
        for _b in range(int(rng.integers(2, 6))):
            cx, cy = rng.uniform(0.1, 0.9) * R, rng.uniform(0.1, 0.9) * C
            sx, sy = rng.uniform(0.05, 0.25) * R, rng.uniform(0.05, 0.25) * C
            rho = rng.uniform(-0.7, 0.7)
            dx, dy = (rr - cx) / sx, (cc - cy) / sy
            lam += float(rng.uniform(1.0, 4.0)) * torch.exp(-0.5 * (dx * dx - 2 * rho * dx * dy + dy * dy) / (1 - rho * rho))
        score = (lam * torch.empty((R, C), dtype=torch.float32, device=dev).exponential_(generator=gen)).ravel()
        idx = torch.sort(torch.topk(score, n, sorted=False).indices).values
        a, b = int(col_ptr[g]), int(col_ptr[g + 1])
        row_idx[a:b] = idx.to(torch.int32)
        val[a:b] = 1.0 + torch.poisson(lam.ravel()[idx], generator=gen)
    sx_all = (np.arange(R * C) // C).astype(np.int32); sy_all = (np.arange(R * C) % C).astype(np.int32)
    dsm = DeviceSpotMatrix.__new__(DeviceSpotMatrix)
    dsm.host = _StagedHost(col_ptr, ["long_%d" % i for i in range(n_genes)], sx_all, sy_all, False)
    dsm.device, dsm.compact, dsm._plans = dev, False, {}
    up = lambda a, k=None: torch.from_numpy(np.ascontiguousarray(a)).to(dev, non_blocking=True)
    dsm.col_ptr, dsm.row_idx, dsm.val = up(col_ptr), row_idx, val
    dsm.spot_x, dsm.spot_y = up(sx_all), up(sy_all)
    dsm._up = up
    dsm.h2d_bytes = 0
    return dsm


def simulate_adata(config: str = "T", seed: int = 0, **kw):
    """Same data as an AnnData-like object (float values, not truncated) for the SPFinder façade."""
    from ..adata_lite import AnnDataLite
    sm = simulate_spot_matrix(config, seed, truncate_int16=False, **kw)
    X = sparse.csc_matrix((sm.val, sm.row_idx, sm.col_ptr), shape=(sm.n_spots, sm.n_genes)).tocsr()
    obs = pd.DataFrame({"x": sm.spot_x.astype(int), "y": sm.spot_y.astype(int)},
                       index=["spot_%d" % i for i in range(sm.n_spots)])
    var = pd.DataFrame({"gene_ids": sm.genes}, index=sm.genes)
    return AnnDataLite(X, obs=obs, var=var, obsm={"spatial": np.column_stack([sm.spot_x, sm.spot_y])})
