#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native STMiner hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload S50|V|T]

A "step" is one pass of the hot path over one batch: the count-weighted GMM fit (on-device
k-means++/Lloyd initialisation + EM, K=20, reg_covar=tol=1e-3, max_iter=100 — the reference's
``fit_pattern`` defaults) of every gene of the synthetic Stereo-seq bin50 data set
(40,000 bins x 10,000 genes, BASELINE.json configs[2], the configuration the genes/s target is quoted on).
With N ranks the genes of that ONE data set are dealt over the ranks in chunks of consecutive genes (BASELINE
configs[2]: "genes sharded over 1/2/4/8 B200" — strong scaling); a step then ends with the one packed NCCL
all-gather that replicates the fitted parameters on every rank.

One JSON line is printed by rank 0.  ``value`` is genes/s with the shards resident in HBM (fit + all-gather);
``e2e`` is the same metric through the host-buffer API ``sharding.fit_genes_sharded`` (pinned H2D of the rank's
columns + fit + all-gather + D2H of the parameters inside the timed region); ``pairs`` reports the second half of
the metric (SVG-pair GMM distances/s on 4,000 fitted genes, 7,998,000 pairs, pair slices sharded over the ranks and
all-gathered in place); ``emd`` the spot-level EMD of ``spatial_high_variable_genes`` on Visium-size problems,
sharded over the ranks as well.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K_COMP = 20
FLOP_PER_SPOT_COMP_ITER = 30          # algorithmic FLOPs of one (spot, component, iteration)
NOMINAL_FP32_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12   # 74.4, SURVEY.md §8d


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="S50", choices=["S50", "V", "T"])
    ap.add_argument("--genes", type=int, default=0, help="override gene count (debug)")
    ap.add_argument("--pairs-genes", type=int, default=4000)
    ap.add_argument("--cpu-sample", type=int, default=0, help="genes in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--emd-genes", type=int, default=2000, help="Visium-size spot-EMD problems in the emd leg (0 = skip)")
    ap.add_argument("--s10-genes", type=int, default=2072, help="long (Stereo-seq bin10-shaped) genes in the s10 leg (0 = skip)")
    ap.add_argument("--no-mds", action="store_true", help="skip the MDS (cluster_gene) leg")
    ap.add_argument("--no-adata", action="store_true", help="skip the fit_gmms(adata) end-to-end leg")
    ap.add_argument("--emd50-genes", type=int, default=8, help="Stereo-seq bin50-size spot-EMD problems (0 = skip)")
    return ap.parse_args()


def workload(name, seed, genes=0):
    """Synthetic data set of a BASELINE config, cached under /tmp (generation is setup, never timed)."""
    from stminer_b200.Simulate import CONFIGS, simulate_spot_matrix
    from stminer_b200.staging import SpotMatrix
    R, C, hexl, nt, rep = CONFIGS[name]
    if genes:
        nt = max(1, min(nt, genes // max(1, min(rep, genes))))
        rep = max(1, genes // nt)
    path = "/tmp/stm_bench_%s_%d_%d_%d.npz" % (name, seed, nt, rep)
    if os.path.exists(path):
        z = np.load(path)
        return SpotMatrix(z["col_ptr"], z["row_idx"], z["val"], z["spot_x"], z["spot_y"],
                          ["gene_%d" % i for i in range(len(z["col_ptr"]) - 1)])
    sm = simulate_spot_matrix(name, seed=seed, n_templates=nt, replicas=rep)
    try:
        tmp = "%s.%d.tmp.npz" % (path[:-4], os.getpid())
        np.savez(tmp, col_ptr=sm.col_ptr, row_idx=sm.row_idx, val=sm.val, spot_x=sm.spot_x, spot_y=sm.spot_y)
        os.replace(tmp, path)              # atomic: a concurrent reader never sees a partial file
    except OSError:
        pass
    return sm


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = max(mx, float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arm
def _cpu_fit_one(args):
    os.environ["OMP_NUM_THREADS"] = "1"
    xy, c, K = args
    from oracle import gmm_oracle as go
    t = time.perf_counter()
    gm = go.fit_gmm_reference(xy, c, K)
    return time.perf_counter() - t, (0 if gm is None else int(gm.n_iter_)), (float("nan") if gm is None else float(gm.lower_bound_))


class CpuReference:
    """The reference CPU path (array_to_list + sklearn GaussianMixture.fit, KMeans init included;
    STMiner/Algorithm/distribution.py:122-130) on a bounded sample: one process per core, 1 BLAS thread each."""

    def __init__(self, cores):
        import multiprocessing as mp
        self.cores = cores
        self.pool = mp.get_context("spawn").Pool(cores, initializer=_pool_init)
        self.pool.map(_noop, range(cores * 2))             # spin the workers up (imports) before timing

    def rate(self, sm, n_sample, seed=123):
        rng = np.random.default_rng(seed)
        idx = rng.choice(sm.n_genes, size=min(n_sample, sm.n_genes), replace=False)
        jobs = [(*sm.gene_spots(int(g)), K_COMP) for g in idx]
        t = time.perf_counter()
        res = self.pool.map(_cpu_fit_one, jobs, chunksize=1)
        wall = time.perf_counter() - t
        self.last_idx = idx
        self.last_lower_bound = np.array([r[2] for r in res])
        return len(jobs) / wall, wall, float(np.mean([r[0] for r in res])), len(jobs)

    def close(self):
        self.pool.terminate()


def _pool_init():
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["MKL_NUM_THREADS"] = "1"
    sys.path.insert(0, ROOT)
    import sklearn.mixture  # noqa: F401


def _noop(_):
    return 0


def _cpu_pairs(job):
    """distribution_distance (distance.py:13-36) of the pairs (ia[k], ib[k]) on one core: (seconds, pairs, distances)."""
    W, MU, COV, ia, ib = job
    from oracle import dist_oracle as do
    t = time.perf_counter()
    d = np.array([do.distribution_distance(W[a], MU[a], COV[a], W[b], MU[b], COV[b]) for a, b in zip(ia, ib)])
    return time.perf_counter() - t, len(ia), d


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the host cores."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sm = workload(args.workload, seed=0, genes=args.genes)
    per_step = args.cpu_sample or 6 * max(cores, 8)      # several fits per core and step: the wall is not one slow gene
    rates, walls = [], []
    cpu = CpuReference(cores)
    for s in range(args.warmup + args.steps):
        rate, wall, mean_fit, n = cpu.rate(sm, per_step, seed=s)
        if s >= args.warmup:
            rates.append(rate); walls.append(wall)
    cpu.close()
    value = float(len(rates) * per_step / sum(walls))
    sample = "%d genes of the %s set per step, process-per-gene pool over %d cores, 1 BLAS thread each" % (
        per_step, args.workload, cores)
    line = {"impl": "reference", "metric": "genes_per_sec_gmm_fit", "value": value, "unit": "genes/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * float(np.mean(walls)), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, sm, 1),
            "cpu_baseline": {"value": value, "unit": "genes/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "genes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args, sm, world):
    """The workload description — the same dict in both arms (nothing here depends on the arm)."""
    names = {"S50": "synthetic Stereo-seq bin50: 40,000 bins x 10,000 genes, K=20 (BASELINE configs[2])",
             "V": "synthetic Visium-shaped: 4,992 spots x 2,000 genes, K=20 (BASELINE configs[1])",
             "T": "synthetic 17x17 test-path stand-in, K=20"}
    return {"workload": names[args.workload], "genes": sm.n_genes, "spots": sm.n_spots,
            "nnz": sm.nnz, "n_comp": K_COMP, "reg_covar": 1e-3, "tol": 1e-3, "max_iter": 100,
            "init": "each arm's own initialisation: sklearn KMeans (reference arm) / on-device weighted k-means++ + Lloyd (GPU arm)",
            "sharding": "GPU arm: the genes of the one data set dealt over the ranks in chunks of 16 consecutive genes, longest "
                        "first (strong scaling), one packed all-gather per step; pairs and EMD problems sharded likewise",
            "l2": "GPU arm: L2 flushed (256 MiB write) between timed steps; the inputs (4 B per stored value) also exceed "
                  "the 126 MB L2 at one GPU"}


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from stminer_b200 import _lib
    from stminer_b200.Algorithm.distribution import fit_spot_matrix, fit_spot_matrix_host
    from stminer_b200.Algorithm.distance import gmm_pair_distances
    from stminer_b200.sharding import gmm_pairs_sharded
    from stminer_b200.staging import DeviceSpotMatrix

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = _lib.load()
    from stminer_b200.sharding import (default_chunk_genes, fit_genes_sharded, pack_fit_result, ranges_to_index,
                                       shard_gene_chunks)
    # ONE data set for all ranks (strong scaling); host copy in the 4 B / value wire format (uint16 spot, int16 value),
    # page-locked.  The genes are dealt over the ranks in chunks of 16 consecutive genes, longest-first.
    if world > 1:                       # rank 0 generates (or finds) the cached data set, the others then load it
        if rank == 0:
            workload(args.workload, seed=0, genes=args.genes)
        dist.barrier()
    sm = workload(args.workload, seed=0, genes=args.genes).compact().pin_memory()
    G = sm.n_genes
    parts = shard_gene_chunks(sm.col_ptr, world, default_chunk_genes(sm.n_genes, world))
    owned = [ranges_to_index(p) for p in parts]
    n_mine = len(owned[rank])
    dsm = DeviceSpotMatrix.from_ranges(sm, parts[rank], device=dev)
    shard = dsm.host
    order, buckets = dsm.plan(K_COMP, _lib.fit_flags(device_init=True, compact=dsm.compact))
    n_launch = int((buckets > 0).sum())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    n_max = max(len(ix) for ix in owned)
    C = 7 * K_COMP + 4
    gbuf = torch.empty((world * n_max, C), dtype=torch.float64, device=dev)      # packed results of all ranks

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def step(res):
        """One step of the hot path: fit this rank's genes; with several ranks pack the seven output arrays into the
        rank's slot of the gather buffer and replicate them with one in-place all-gather."""
        res = fit_spot_matrix(dsm, K_COMP, init=None, seed=1, out=res)
        if world > 1:
            gbuf[rank * n_max:rank * n_max + n_mine] = pack_fit_result(res, K_COMP)
            dist.all_gather_into_tensor(gbuf, gbuf[rank * n_max:(rank + 1) * n_max])
        return res

    res = None
    for _ in range(args.warmup):
        res = step(res)
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # ---- timed: exactly `steps` steps, shards resident in HBM ----------------------------------------
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall = time.perf_counter()
    for s in range(args.steps):
        flush.fill_(s)                       # evict L2 between timed steps (not timed)
        if world > 1:
            dist.barrier()                   # every step starts together on all ranks (the all-gather ends it together)
        ev[s][0].record()
        res = step(res)
        ev[s][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall
    step_ms = [a.elapsed_time(b) for a, b in ev]
    dev_ms = max_over_ranks(float(np.sum(step_ms)))
    ms_per_step = dev_ms / args.steps
    value = G / (ms_per_step / 1e3)

    host_local = res.to_host()
    n_iter_l = host_local["n_iter"].astype(np.float64)
    nnz_l = shard.nnz_per_gene().astype(np.float64)
    ok_l = host_local["status"] == 0
    flops_local = FLOP_PER_SPOT_COMP_ITER * float((nnz_l * n_iter_l * ok_l).sum()) * K_COMP
    flops = flops_local
    if world > 1:
        t = torch.tensor([flops_local], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        flops = float(t.item())
    em_tflops = flops / (ms_per_step / 1e3) / 1e12          # whole job (all ranks) over the step time

    # ---- e2e: host buffers in / out through the public sharded API -----------------------------------
    for _ in range(2):
        fit_genes_sharded(sm, K_COMP, init=None, device=dev, seed=1)
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        host = fit_genes_sharded(sm, K_COMP, init=None, device=dev, seed=1)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0) / args.steps
    # bytes copied per step, summed over the ranks: every rank uploads its own columns (+ the spot table and its chunk
    # orders) and reads the gathered parameters back
    h2d = int(sm.col_ptr.nbytes + sm.row_idx.nbytes + sm.val.nbytes + world * (sm.spot_x.nbytes + sm.spot_y.nbytes) + 4 * G)
    d2h = int(world * G * C * 8)
    n_iter = host["n_iter"].astype(np.float64)
    ok = host["status"] == 0

    # ---- second half of the metric: SVG-pair distances -------------------------------------------
    Gp = min(args.pairs_genes, int(ok.sum()))
    sel = np.flatnonzero(ok)[:Gp]
    W = torch.from_numpy(host["weights"][sel]).to(dev)
    MU = torch.from_numpy(host["means"][sel]).to(dev)
    COV = torch.from_numpy(host["covariances"][sel]).to(dev)
    total_pairs = Gp * (Gp - 1) // 2

    def pairs_step():
        return gmm_pairs_sharded(W, MU, COV, device=dev)      # slice -> kernel writes its slot -> in-place all-gather
    for _ in range(2):
        pairs_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        full = pairs_step()
    e1.record()
    barrier()
    pairs_ms = max_over_ranks(e0.elapsed_time(e1) / 3)
    clocks = sampler.stop() if rank == 0 else None

    # ---- third kernel of the path: exact spot-level EMD (global image vs gene image), Visium-size problems, ----
    # ---- sharded over the ranks like the loop of spatial_high_variable_genes (strong scaling) --------------
    emd = None
    if args.emd_genes > 0:
        from scipy import sparse as _sp
        from stminer_b200.sharding import emd_sharded
        from stminer_b200.Simulate import simulate_spot_matrix
        nt = max(1, (args.emd_genes + 3) // 4)
        vsm = simulate_spot_matrix("V", seed=0, n_templates=nt, replicas=4, truncate_int16=False)
        Xv = _sp.csc_matrix((vsm.val, vsm.row_idx, vsm.col_ptr), shape=(vsm.n_spots, vsm.n_genes))
        src = (vsm.spot_x, vsm.spot_y, np.asarray(Xv.sum(axis=1)).ravel())
        tg = []
        for gi in range(min(args.emd_genes, vsm.n_genes)):
            a_, b_ = vsm.col_ptr[gi], vsm.col_ptr[gi + 1]
            tg.append((vsm.spot_x[vsm.row_idx[a_:b_]], vsm.spot_y[vsm.row_idx[a_:b_]], vsm.val[a_:b_].astype(np.float64)))
        emd_sharded(src, tg[:4 * world], device=dev)                          # warm-up (module load, allocator)
        barrier()
        t_e = time.perf_counter()
        ecost, estatus = emd_sharded(src, tg, device=dev)
        barrier()
        t_e = max_over_ranks(time.perf_counter() - t_e)
        emd = {"metric": "spot_emd_problems_per_sec", "value": len(tg) / t_e, "unit": "problems/s",
               "problems": len(tg), "source_spots": int(len(src[0])), "mean_target_spots": float(np.mean([len(t[0]) for t in tg])),
               "seconds": t_e, "optimal": int((estatus == 0).sum()),
               "sharding": "problems dealt longest-first over %d rank(s), one all-gather of (cost, status)" % world,
               "note": "exact network simplex with the multiscale start basis, numItermax=200000 as the reference; wall "
                       "time of sharding.emd_sharded incl. H2D/D2H of the problem data and the gather",
               "issue_slot_util": ncu_metric("emd_issue_active_pct"),
               "ncu_file": "profiles/r2_ncu_emd_visium.txt"}

    # ---- the user's call: fit_gmms(adata, genes) from an AnnData-like object (CSR float32 counts), staging included ----
    from_adata = None
    if rank == 0 and world == 1 and not args.no_adata:
        import pandas as pd
        from scipy import sparse as _sp2
        from stminer_b200.adata_lite import AnnDataLite
        from stminer_b200.Algorithm.distribution import fit_gmms
        Xc = _sp2.csc_matrix((sm.val.astype(np.float32), sm.row_idx.astype(np.int32), sm.col_ptr), shape=(sm.n_spots, G))
        Xr = Xc.tocsr()                                   # what an h5ad holds: observations x genes, CSR
        del Xc
        gnames = ["gene_%d" % i for i in range(G)]
        adata = AnnDataLite(Xr, obs=pd.DataFrame({"x": sm.spot_x.astype(int), "y": sm.spot_y.astype(int)}),
                            var=pd.DataFrame(index=gnames), obsm={})
        fit_gmms(adata, gnames[:64], n_comp=K_COMP, seed=1)                    # warm-up
        torch.cuda.synchronize()
        t_a = []
        for _ in range(3):
            t0a = time.perf_counter()
            gm = fit_gmms(adata, gnames, n_comp=K_COMP, seed=1)
            t_a.append(time.perf_counter() - t0a)
        from stminer_b200.staging import stage_csr_on_device
        t0a = time.perf_counter()
        dsm_a = stage_csr_on_device(adata.X, sm.spot_x, sm.spot_y, gnames, device=dev)
        torch.cuda.synchronize()
        t_stage = time.perf_counter() - t0a
        from_adata = {"metric": "genes_per_sec_fit_gmms_from_adata", "value": G / float(np.median(t_a)), "unit": "genes/s",
                      "seconds": float(np.median(t_a)), "genes_fitted": len(gm), "staging_seconds": t_stage,
                      "h2d_bytes": int(dsm_a.h2d_bytes),
                      "note": "wall time of fit_gmms(adata, all genes, n_comp=20) — the reference's call "
                              "(distribution.py:148) — from a CSR float32 matrix in pageable host memory: upload of the "
                              "CSR arrays, device staging (csrc/staging.cu: stable transposition + int16 truncation), "
                              "fit, read-back and the construction of the 10,000 GMM objects; staging_seconds = upload + "
                              "staging alone (the host staging it replaces took 4.5 s)"}
        del dsm_a, adata, Xr

    # ---- spot EMD at Stereo-seq bin50 size: the per-spot total image of the S50 set (40,000 bins) against sparse genes ----
    emd50 = None
    if rank == 0 and world == 1 and args.emd50_genes > 0:
        from stminer_b200.Algorithm.distance import emd_batched as _emd
        tot50 = np.zeros(sm.n_spots)
        np.add.at(tot50, sm.row_idx.astype(np.int64), sm.val.astype(np.float64))
        k50 = tot50 > 0
        src50 = (sm.spot_x[k50], sm.spot_y[k50], tot50[k50])
        rng50 = np.random.default_rng(0)
        tg50 = []
        for gi in range(args.emd50_genes):               # every 100th gene (one per pattern), thinned to ~2,000 bins
            g_ = (gi * 100) % G
            a_, b_ = int(sm.col_ptr[g_]), int(sm.col_ptr[g_ + 1])
            pick = np.sort(rng50.permutation(b_ - a_)[:min(b_ - a_, 2000)]) + a_
            r_ = sm.row_idx[pick].astype(np.int64)
            tg50.append((sm.spot_x[r_], sm.spot_y[r_], sm.val[pick].astype(np.float64)))
        torch.cuda.synchronize()
        t_5 = time.perf_counter()
        c50, s50_, i50 = _emd(src50, tg50, device=dev, return_iters=True)
        torch.cuda.synchronize()
        t_5 = time.perf_counter() - t_5
        emd50 = {"metric": "spot_emd_bin50_seconds", "problems": len(tg50), "source_bins": int(k50.sum()),
                 "mean_target_bins": float(np.mean([len(t[0]) for t in tg50])), "seconds": t_5,
                 "optimal": int((s50_ == 0).sum()), "mean_pivots": float(i50.mean()),
                 "note": "global-memory-tree kernel (beyond one CTA's shared memory), one CTA per problem, all in flight "
                         "together; 8 dense bin50 genes (26,000 bins each) in flight take 14.5 s (profiles/r2_tuning_log.txt)"}

    # ---- the step after the distance matrix: cluster() = SMACOF MDS (4 restarts, 20 components) on the G x G matrix ----
    mds = None
    if rank == 0 and world == 1 and not args.no_mds and Gp >= 64:
        from stminer_b200.Algorithm.algorithm import mds_fit_transform
        from stminer_b200.Algorithm.distance import pairs_to_square
        local_all = gmm_pair_distances(W, MU, COV, 0, total_pairs, device=dev)
        square = pairs_to_square(local_all, Gp)
        torch.cuda.synchronize()
        Dh = square.cpu().numpy()
        mds_fit_transform(Dh[:256, :256], 20, random_state=0, device=dev)       # warm-up
        torch.cuda.synchronize()
        t_m = time.perf_counter()
        emb, stress, n_it = mds_fit_transform(Dh, 20, random_state=0, device=dev)
        t_m = time.perf_counter() - t_m
        mds = {"metric": "mds_smacof_seconds", "genes": Gp, "n_components": 20, "n_init": 4, "seconds": t_m,
               "best_stress": stress, "n_iter_best": n_it,
               "note": "cluster() MDS half (algorithm.py:54-55), sklearn 1.3.0 semantics, host matrix in / embedding out"}
        if not args.no_cpu_baseline:
            from oracle import mds_oracle as mo
            n_s = min(1000, Gp)
            Ds = Dh[:n_s, :n_s]
            t_c = time.perf_counter()
            mo.smacof_single(Ds, np.random.RandomState(0).uniform(size=(n_s, 20)), max_iter=3, eps=-np.inf)
            t_c = (time.perf_counter() - t_c) / 3
            mds["cpu_seconds_per_iteration_%d_genes" % n_s] = t_c

    # ---- long genes (BASELINE configs[3], Stereo-seq bin10-shaped): one gene per thread-block cluster ----------------
    s10 = None
    if rank == 0 and world == 1 and args.s10_genes > 0:
        from stminer_b200.Simulate import simulate_long_genes_device
        ldsm = simulate_long_genes_device(args.s10_genes, seed=0, device=dev)       # generated on the device (setup, untimed)
        lsm = ldsm.host
        _, lb = ldsm.plan(K_COMP, _lib.fit_flags(device_init=True))
        lres = None
        for _ in range(2):
            lres = fit_spot_matrix(ldsm, K_COMP, init=None, seed=1, out=lres)
        torch.cuda.synchronize()
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0.record()
        for _ in range(3):
            lres = fit_spot_matrix(ldsm, K_COMP, init=None, seed=1, out=lres)
        l1.record()
        torch.cuda.synchronize()
        lms = l0.elapsed_time(l1) / 3
        lh = lres.to_host()
        lflops = FLOP_PER_SPOT_COMP_ITER * K_COMP * float((lsm.nnz_per_gene() * lh["n_iter"] * (lh["status"] == 0)).sum())
        s10 = {"metric": "genes_per_sec_gmm_fit_long_genes", "value": lsm.n_genes / (lms / 1e3), "unit": "genes/s",
               "genes": lsm.n_genes, "grid": "1000 x 1000", "nnz": int(lsm.nnz), "max_nnz": int(lsm.nnz_per_gene().max()),
               "ms": lms, "spots_per_sec": lsm.nnz / (lms / 1e3), "em_tflops": lflops / (lms / 1e3) / 1e12,
               "bucket_counts_cluster16_8_4_2_large_medium_small": lb.tolist(), "genes_ok": int((lh["status"] == 0).sum()),
               "note": "BASELINE configs[3]-shaped genes on the 1000 x 1000 grid (distinct-spot counts log-uniform in [1e3, 5e5], "
                       "generated on the device); genes beyond one CTA's shared memory are split over thread-block clusters"}
        del ldsm, lres

    # ---- FP32 FMA peak measured on this GPU (roofline denominator) --------------------------------
    scratch = torch.zeros(4, dtype=torch.float32, device=dev)
    blocks, iters = 148 * 8, 20000
    for _ in range(2):
        lib.stm_bench_fma(blocks, iters, scratch.data_ptr(), _lib.current_stream_ptr())
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    lib.stm_bench_fma(blocks, iters, scratch.data_ptr(), _lib.current_stream_ptr())
    f1.record()
    torch.cuda.synchronize()
    fma_peak = blocks * 256 * iters * 32 / (f0.elapsed_time(f1) / 1e3) / 1e12

    if rank != 0:
        return
    line = {
        "metric": "genes_per_sec_gmm_fit", "value": value, "unit": "genes/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, sm, world),
        "e2e": {"value": G / e2e_s, "unit": "genes/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_s,
                "api": "stminer_b200.sharding.fit_genes_sharded (what fit_gmms calls): host CSC matrix in, host parameter "
                       "arrays out, on every rank"},
        "gpu_launches": args.steps * n_launch * world,
        "input_encoding": "uint16 spot + int16 value (4 B/value)" if sm.is_compact else "int32 spot + float32 value",
        "roofline": {"bound": "fp32_fma", "achieved": em_tflops, "peak": world * fma_peak, "unit": "TFLOP/s",
                     "frac": em_tflops / (world * fma_peak),
                     # ncu --set full of this command's step launches (profiles/r1h_ncu_bench_step.txt):
                     # dram__bytes_read+write = 374.3 MB (large bucket) + 150.4 MB (medium) vs 506 MB of input
                     "traffic": ncu_metric("em_dram_bytes_per_step") if (args.workload == "S50" and not args.genes and world == 1) else None,
                     "traffic_unit": "bytes per step (sum over the step's bucket launches), dram__bytes_read + write of the "
                                     "committed ncu --set full capture (profiles/r2_ncu_metrics.json)",
                     "kernel": "gmm_em_kernel (%d nnz-bucket launches per step, run concurrently)" % n_launch,
                     "peak_source": "measured here with stm_bench_fma (FFMA chains) on one GPU x n_gpus; nominal %.1f TFLOP/s per "
                                    "GPU; MEASURED_PEAKS.json has no FP32 figure" % NOMINAL_FP32_TFLOPS,
                     "algorithmic_flops_per_step": flops,
                     "note": "EM sweeps only: 30 FLOP x nnz_g x K x n_iter_g; the k-means++/Lloyd init inside the "
                             "same launch is not credited",
                     "hbm_frac": (sm.nnz * (4 if sm.is_compact else 8) + G * K_COMP * 7 * 8) / (ms_per_step / 1e3) / 1e9 / (world * peak_hbm())},
        "pairs": {"metric": "svg_pair_gmm_distances_per_sec", "value": total_pairs / (pairs_ms / 1e3),
                  "unit": "pairs/s", "genes": Gp, "pairs": total_pairs, "ms": pairs_ms,
                  "sharding": "contiguous pair slices per rank written into the gather buffer + one in-place NCCL all-gather" if world > 1 else "single GPU",
                  "issue_slot_util": ncu_metric("dist_issue_active_pct"), "ncu_file": "profiles/r2_ncu_dist.txt"},
        "fit": {"mean_n_iter": float(n_iter[ok].mean()), "genes_ok": int(ok.sum()), "genes_dropped": int((~ok).sum()),
                "genes_per_rank": [len(ix) for ix in owned],
                "bucket_counts_cluster16_8_4_2_large_medium_small_rank0": buckets.tolist(), "wall_ms_per_step": 1e3 * t_wall / args.steps},
        "e2e_from_adata": from_adata,
        "emd": emd,
        "emd_bin50": emd50,
        "mds": mds,
        "s10": s10,
        "clocks": clocks,
    }
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        n_sample = args.cpu_sample or min(512, 6 * max(cores, 8))
        cpu = CpuReference(cores)
        rate, wall, mean_fit, n = cpu.rate(sm, n_sample)
        cpu.close()
        # quality of the on-device initialisation against the reference's (sklearn KMeans) on the SAME sample genes: final
        # lower bound (mean log-likelihood per count) of the GPU fit minus that of the CPU fit
        lb_gpu = host["lower_bound"][cpu.last_idx]
        lb_cpu = cpu.last_lower_bound
        both = np.isfinite(lb_cpu) & (host["status"][cpu.last_idx] == 0)
        line["fit"]["lower_bound_vs_cpu"] = {
            "genes": int(both.sum()), "gpu_mean": float(lb_gpu[both].mean()), "cpu_mean": float(lb_cpu[both].mean()),
            "mean_diff": float((lb_gpu[both] - lb_cpu[both]).mean()),
            "frac_gpu_within_0.01_or_better": float(np.mean(lb_gpu[both] >= lb_cpu[both] - 0.01)),
            "note": "same genes as cpu_baseline.sample; CPU = sklearn GaussianMixture.lower_bound_ (KMeans init, unseeded), "
                    "GPU = on-device k-means++ (exponential race) + capped Lloyd init; nats per count"}
        # the other two stages of the path, on bounded samples (the reference runs them serially on one core)
        try:
            n_p = 10240
            Wh, MUh, COVh = host["weights"][sel], host["means"][sel], host["covariances"][sel]
            rngp = np.random.default_rng(1)
            ia = rngp.integers(0, len(sel), n_p); ib = (ia + 1 + rngp.integers(0, len(sel) - 1, n_p)) % len(sel)
            chunks = np.array_split(np.arange(n_p), cores)
            import multiprocessing as mp
            with mp.get_context("spawn").Pool(cores, initializer=_pool_init) as pool:
                pool.map(_cpu_pairs, [(Wh[:2], MUh[:2], COVh[:2], np.array([0]), np.array([1]))] * cores)   # numba compile
                t_c = time.perf_counter()
                outp = pool.map(_cpu_pairs, [(Wh, MUh, COVh, ia[c], ib[c]) for c in chunks])
                t_c = time.perf_counter() - t_c
            one_core = float(np.mean([o[1] / max(o[0], 1e-9) for o in outp]))
            line["pairs"]["cpu_pairs_per_sec"] = n_p / t_c
            line["pairs"]["cpu_pairs_per_sec_1_core"] = one_core
            line["pairs"]["cpu_cores"] = cores
            line["pairs"]["cpu_note"] = ("reference arithmetic (K^2 njit Hellinger calls + scipy LSAP per pair, "
                                         "distance.py:13-36) on %d random pairs of the fitted genes over a %d-process pool "
                                         "(the reference itself has no parallel path for this stage)" % (n_p, cores))
            # agreement of the device result on the sampled pairs
            fullh = full.cpu().numpy()
            lin = lambda i, j: (i * (2 * Gp - i - 1)) // 2 + (j - i - 1)
            ref_d = np.concatenate([o[2] for o in outp])
            ii = np.minimum(ia, ib); jj = np.maximum(ia, ib)
            got_d = fullh[[lin(int(i), int(j)) for i, j in zip(ii, jj)]]
            line["pairs"]["max_rel_diff_vs_cpu_sample"] = float(np.max(np.abs(got_d - ref_d) / np.maximum(np.abs(ref_d), 1e-12)))
        except Exception as e:        # pragma: no cover
            line["pairs"]["cpu_note"] = "cpu sample failed: %r" % (e,)
        if emd is not None:
            try:
                from oracle import emd_oracle as eo
                small = min(range(len(tg)), key=lambda i: len(tg[i][0]))
                t_c = time.perf_counter()
                c_cpu, st_cpu, it_cpu = eo.emd_network_simplex(*src, *tg[small])
                t_c = time.perf_counter() - t_c
                emd["cpu_seconds_per_problem_1_core"] = t_c
                emd["cpu_note"] = ("C restatement of POT's network simplex from the artificial star (POT is not installable "
                                   "here) on the smallest problem of the batch (%d target spots, %d pivots), one core; GPU cost "
                                   "agrees to %.1e" % (len(tg[small][0]), it_cpu, abs(c_cpu - float(ecost[small])) / max(1.0, abs(c_cpu))))
            except Exception as e:    # pragma: no cover
                emd["cpu_note"] = "cpu sample failed: %r" % (e,)
        line["cpu_baseline"] = {"value": rate, "unit": "genes/s", "cores": cores, "kind": "port",
                                "sample": "%d random genes of the same data set, reference path (array_to_list + "
                                          "sklearn GaussianMixture.fit, KMeans init included), process-per-gene "
                                          "pool, 1 BLAS thread each; %.1f s wall, %.2f s mean per fit" % (n, wall, mean_fit)}
    print(json.dumps(line), flush=True)


def ncu_metric(name):
    """A number taken from the committed ncu captures of this command (profiles/r2_ncu_metrics.json, written by
    scripts/ncu_summary.py from the .ncu-rep files); None when the file or the entry is missing."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_metrics.json"))).get(name)
    except Exception:
        return None


def peak_hbm():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0   # B200_PROFILING.md fallback


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        import torch
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
