import numpy as np, time, sys
sys.path.insert(0, '.')
from stminer_b200.Simulate import simulate_spot_matrix
from tests.helpers import *
from oracle import gmm_oracle as go
from stminer_b200.Algorithm.distribution import fit_spot_matrix
from stminer_b200.staging import DeviceSpotMatrix
import torch
for config,K,nt,rep in [("T",10,4,3),("V",20,4,2)]:
    sm = simulate_spot_matrix(config, seed=7, n_templates=nt, replicas=rep)
    W, MU, COV, ok = oracle_inits(sm, K, seed=11)
    dsm = DeviceSpotMatrix(sm)
    print('plan', dsm.plan(K)[1])
    res = fit_spot_matrix(dsm, K, init=(W,MU,COV)).to_host()
    ref = oracle_fit_all(sm, K, W, MU, COV, ok)
    for g in range(sm.n_genes):
        if not ok[g]: print(g, 'dropped', res['status'][g]); continue
        r=ref[g]
        print(g, res['status'][g], res['n_iter'][g], r['n_iter'], res['lower_bound'][g]-r['lower_bound'],
              np.abs(res['weights'][g]-r['weights']).max(), np.abs(res['means'][g]-r['means']).max(),
              np.abs(res['covariances'][g]-r['covariances']).max()/np.abs(r['covariances']).max())
    # device init: max_iter=... compare EM-from-device-init against oracle EM from the same init
    init = fit_spot_matrix(dsm, K, init=None, max_iter=1, tol=1e30, seed=5)  # one iteration (params after 1 M-step)
    full = fit_spot_matrix(dsm, K, init=None, seed=5).to_host()
    print('device init fit: status', np.bincount(full['status']), 'n_iter', full['n_iter'][:12], 'lb', full['lower_bound'][:4])
    for g in range(min(4, sm.n_genes)):
        xy, c = sm.gene_spots(g)
        gm = go.fit_gmm_reference(xy, c, K, random_state=0)
        print('  gene', g, 'device lb %.4f n_iter %d | sklearn lb %.4f n_iter %d' % (full['lower_bound'][g], full['n_iter'][g], gm.lower_bound_, gm.n_iter_))
# throughput V full
sm = simulate_spot_matrix("V", seed=0)
dsm = DeviceSpotMatrix(sm)
print('plan', dsm.plan(20)[1])
for it in range(3):
    torch.cuda.synchronize(); t=time.time()
    r = fit_spot_matrix(dsm, 20, init=None, seed=1); torch.cuda.synchronize(); dt=time.time()-t
    print('V 2000 genes device-init fit: %.4f s  %.1f genes/s  mean n_iter %.1f' % (dt, 2000/dt, r.n_iter.float().mean().item()))
