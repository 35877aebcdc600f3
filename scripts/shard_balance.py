"""Predicts the strong-scaling step time at N ranks on ONE GPU: fits every rank's shard of the S50 set on its own and
reports the slowest (what the all-gather waits for), for the contiguous and the chunk-dealt partition."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from bench import K_COMP, workload
from stminer_b200.Algorithm.distribution import fit_spot_matrix
from stminer_b200.sharding import shard_gene_chunks, shard_genes_contiguous
from stminer_b200.staging import DeviceSpotMatrix

sm = workload("S50", 0).compact().pin_memory()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def time_shard(ranges):
    dsm = DeviceSpotMatrix.from_ranges(sm, ranges)
    res = None
    ts = []
    for i in range(6):
        flush.fill_(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); res = fit_spot_matrix(dsm, K_COMP, init=None, seed=1, out=res); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    it = res.n_iter.cpu().numpy()
    return float(np.median(ts[2:])), float(it.mean()), dsm.host.nnz


for world in (int(a) for a in (sys.argv[1:] or ["8"])):
    b = shard_genes_contiguous(sm.col_ptr, world)
    cont = [[(int(b[r]), int(b[r + 1]))] for r in range(world)]
    for name, parts in (("contiguous", cont), ("chunks of 16", shard_gene_chunks(sm.col_ptr, world, 16)),
                        ("chunks of 64", shard_gene_chunks(sm.col_ptr, world, 64))):
        res = [time_shard(p) for p in parts]
        t = [r[0] for r in res]
        print("N=%d %-13s per-rank ms %s  max %.2f mean %.2f  (mean n_iter per rank %s)"
              % (world, name, " ".join("%.2f" % x for x in t), max(t), np.mean(t), " ".join("%.1f" % r[1] for r in res)))
