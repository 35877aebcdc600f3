import sys
sys.path.insert(0, '.')
import numpy as np, torch
from stminer_b200.Simulate import simulate_long_genes
from stminer_b200.Algorithm.distribution import fit_spot_matrix
from stminer_b200.staging import DeviceSpotMatrix
sm = simulate_long_genes(int(sys.argv[1]) if len(sys.argv) > 1 else 296, seed=0)
dsm = DeviceSpotMatrix(sm)
r = fit_spot_matrix(dsm, 20, init=None, seed=1)
torch.cuda.synchronize()
