import sys
sys.path.insert(0, '.')
import torch
from bench import workload
from stminer_b200.Algorithm.distribution import fit_spot_matrix
from stminer_b200.staging import DeviceSpotMatrix
name = sys.argv[1]; genes = int(sys.argv[2])
sm = workload(name, 0, genes).compact()
dsm = DeviceSpotMatrix(sm)
r = fit_spot_matrix(dsm, 20, init=None, seed=1)
torch.cuda.synchronize()
