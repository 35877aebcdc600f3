import sys, time, numpy as np
sys.path.insert(0, ".")
import torch
from scipy import sparse
from stminer_b200 import SPFinder
from stminer_b200.Simulate import simulate_spot_matrix
from stminer_b200.adata_lite import AnnDataLite
import pandas as pd
sm = simulate_spot_matrix("S50", seed=0, n_templates=3, replicas=2, truncate_int16=False)
X = sparse.csc_matrix((sm.val, sm.row_idx, sm.col_ptr), shape=(sm.n_spots, sm.n_genes)).tocsr()
# thin the genes (sparser genes, as most real genes are): keep ~8 % of each gene's spots
rng = np.random.default_rng(0)
Xc = X.tocsc()
keep = rng.random(Xc.nnz) < 0.08
keep[Xc.indptr[0]:Xc.indptr[2]] = True           # two genes stay dense: the per-spot total then covers the whole slide
Xc.data = Xc.data * keep
Xc.eliminate_zeros()
X = Xc.tocsr()
obs = pd.DataFrame({"x": sm.spot_x.astype(int), "y": sm.spot_y.astype(int)})
a = AnnDataLite(X, obs=obs, var=pd.DataFrame(index=sm.genes), obsm={})
sp = SPFinder(a)
t0 = time.perf_counter(); sp.get_genes_csr_array(min_cells=50, normalize=False, gene_list=list(sm.genes[2:])); t1 = time.perf_counter()
sp.spatial_high_variable_genes(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("spots with counts:", int((np.asarray(X.sum(axis=1)).ravel() > 0).sum()), "gene nnz:", np.diff(X.tocsc().indptr).tolist())
print("staging %.2f s, SVG %.2f s" % (t1 - t0, t2 - t1)); print(sp.global_distance)
