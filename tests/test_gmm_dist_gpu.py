"""GPU parity: K x K Hellinger-cost assignment distances (through the C ABI) vs the oracle / golden values."""
import os

import numpy as np
import pytest

from oracle import dist_oracle as do

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "gmm_dist.npz")


@pytest.mark.parametrize("n", range(4))
def test_pairs_match_reference_golden(cuda_device, n):
    from stminer_b200.Algorithm.distance import gmm_pair_distances, pairs_to_square
    z = np.load(GOLD)
    W, MU, COV, D = z["s%d_W" % n], z["s%d_MU" % n], z["s%d_COV" % n], z["s%d_D" % n]
    G = W.shape[0]
    sq = pairs_to_square(gmm_pair_distances(W, MU, COV), G).cpu().numpy()
    # float64 kernel vs the reference's own linear_sum: far inside the north-star rtol of 1e-4
    np.testing.assert_allclose(sq, D, rtol=1e-9, atol=1e-12)
    assert (np.diag(sq) == 0).all() and np.array_equal(sq, sq.T)


@pytest.mark.parametrize("G,K", [(40, 20), (25, 10), (9, 1), (7, 2), (12, 32), (8, 45), (6, 64)])
def test_pairs_match_oracle_random(cuda_device, G, K):
    from stminer_b200.Algorithm.distance import gmm_pair_distances, pairs_to_square
    W, MU, COV = do.random_gmms(G, K, extent=80.0, seed=G * 100 + K)
    ref = do.build_gmm_distance_array(W, MU, COV)
    sq = pairs_to_square(gmm_pair_distances(W, MU, COV), G).cpu().numpy()
    np.testing.assert_allclose(sq, ref, rtol=1e-9, atol=1e-12)


def test_pair_slices_concatenate(cuda_device):
    """A rank computes an arbitrary [begin, end) slice of the linearised triangle (multi-GPU sharding)."""
    from stminer_b200.Algorithm.distance import gmm_pair_distances
    from stminer_b200.sharding import pair_slice
    W, MU, COV = do.random_gmms(33, 20, seed=5)
    full = gmm_pair_distances(W, MU, COV).cpu().numpy()
    total = 33 * 32 // 2
    parts = []
    for r in range(5):
        b, e = pair_slice(total, r, 5)
        parts.append(gmm_pair_distances(W, MU, COV, b, e).cpu().numpy())
    assert np.array_equal(np.concatenate(parts), full)
    assert gmm_pair_distances(W, MU, COV, 7, 7).numel() == 0


def test_one_vs_all_and_analytic(cuda_device):
    from stminer_b200.Algorithm.distance import gmm_one_vs_all
    W, MU, COV = do.random_gmms(20, 20, seed=9)
    ref = do.compare_gmm_distance(W[4], MU[4], COV[4], W, MU, COV)
    got = gmm_one_vs_all(W[4], MU[4], COV[4], W, MU, COV).cpu().numpy()
    np.testing.assert_allclose(got, ref, rtol=1e-9, atol=1e-7)   # self-distance ~1e-8 on both sides
    assert got[4] < 1e-6
    far = gmm_one_vs_all(W[0], MU[0] + 1e6, COV[0], W, MU, COV).cpu().numpy()
    np.testing.assert_allclose(far, 2.0, rtol=0, atol=1e-12)     # all Hellinger = 1 -> distance 2


def test_nan_mixture_gives_nan_not_hang(cuda_device):
    """scipy's LSAP raises on NaN costs (reference: ValueError); the kernel returns NaN for that pair only."""
    from stminer_b200.Algorithm.distance import gmm_pair_distances
    W, MU, COV = do.random_gmms(4, 6, seed=2)
    MU[1, 2, 0] = np.nan
    out = gmm_pair_distances(W, MU, COV).cpu().numpy()
    idx = {(0, 1): 0, (0, 2): 1, (0, 3): 2, (1, 2): 3, (1, 3): 4, (2, 3): 5}
    for (i, j), p in idx.items():
        assert np.isnan(out[p]) == (1 in (i, j))


def test_host_api_dataframes(cuda_device):
    """build_gmm_distance_array / compare_gmm_distance keep the reference's return types (distance.py:83-111)."""
    import pandas as pd
    from stminer_b200.Algorithm.distance import build_gmm_distance_array, compare_gmm_distance, distribution_distance
    from stminer_b200.Algorithm.distribution import GMM
    W, MU, COV = do.random_gmms(6, 10, seed=1)
    d = {}
    for g in range(6):
        m = GMM(10); m.set_weights(W[g]); m.set_mean(MU[g]); m.set_covariances(COV[g])
        d["gene%d" % g] = m
    df = build_gmm_distance_array(d)
    assert isinstance(df, pd.DataFrame) and list(df.index) == list(d) and list(df.columns) == list(d)
    ref = do.build_gmm_distance_array(W, MU, COV)
    np.testing.assert_allclose(df.values, ref, rtol=1e-9, atol=1e-12)
    one = compare_gmm_distance(d["gene2"], d)
    assert one.index[0] == "gene2" and list(one.columns) == [0]
    assert (np.diff(one[0].values) >= 0).all()
    assert abs(distribution_distance(d["gene0"], d["gene1"]) - ref[0, 1]) < 1e-12
