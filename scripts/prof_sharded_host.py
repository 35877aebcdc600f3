"""cProfile of sharding.fit_genes_sharded on rank 0 under torchrun (host overheads of the strong-scaling e2e path)."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, ".")
import numpy as np, torch, torch.distributed as dist
from bench import K_COMP, workload
from stminer_b200.sharding import fit_genes_sharded
rank = int(os.environ.get("RANK", 0)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
if rank == 0:
    workload("S50", 0)
dist.barrier()
sm = workload("S50", 0).compact().pin_memory()
dev = torch.device("cuda", lr)
for _ in range(3):
    fit_genes_sharded(sm, K_COMP, init=None, device=dev, seed=1)
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    fit_genes_sharded(sm, K_COMP, init=None, device=dev, seed=1)
dist.barrier(); torch.cuda.synchronize()
if rank == 0:
    print("world %d: %.2f ms per call" % (dist.get_world_size(), (time.perf_counter() - t0) / 5 * 1e3))
pr = cProfile.Profile(); pr.enable()
fit_genes_sharded(sm, K_COMP, init=None, device=dev, seed=1)
pr.disable()
if rank == 0:
    pstats.Stats(pr).sort_stats("tottime").print_stats(14)
dist.destroy_process_group()
