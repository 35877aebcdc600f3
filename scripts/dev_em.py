import numpy as np, time, sys
sys.path.insert(0, '/root/repo')
from stminer_b200.Simulate import simulate_spot_matrix
from tests.helpers import *
from stminer_b200.Algorithm.distribution import fit_spot_matrix
from stminer_b200.staging import DeviceSpotMatrix
import torch
for config,K,nt,rep in [("T",10,4,3),("V",20,4,2)]:
    sm = simulate_spot_matrix(config, seed=7, n_templates=nt, replicas=rep)
    W, MU, COV, ok = oracle_inits(sm, K, seed=11)
    dsm = DeviceSpotMatrix(sm)
    res = fit_spot_matrix(dsm, K, init=(W,MU,COV)).to_host()
    ref = oracle_fit_all(sm, K, W, MU, COV, ok)
    for g in range(sm.n_genes):
        if not ok[g]: print(g, 'dropped', res['status'][g]); continue
        r=ref[g]
        print(g, res['status'][g], res['n_iter'][g], r['n_iter'], res['lower_bound'][g]-r['lower_bound'],
              np.abs(res['weights'][g]-r['weights']).max(), np.abs(res['means'][g]-r['means']).max(),
              np.abs(res['covariances'][g]-r['covariances']).max()/np.abs(r['covariances']).max())
