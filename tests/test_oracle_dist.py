"""CPU: the distance oracle against golden values made with the reference's own linear_sum, plus analytic cases."""
import os

import numpy as np
import pytest

from oracle import dist_oracle as do

GOLD = os.path.join(os.path.dirname(__file__), "golden", "gmm_dist.npz")


@pytest.mark.parametrize("n", range(4))
def test_distance_matrix_matches_reference_golden(n):
    z = np.load(GOLD)
    W, MU, COV, D = z["s%d_W" % n], z["s%d_MU" % n], z["s%d_COV" % n], z["s%d_D" % n]
    got = do.build_gmm_distance_array(W, MU, COV, vectorised=True)
    np.testing.assert_allclose(got, D, rtol=1e-10, atol=1e-12)
    if n == 2:  # the verbatim njit restatement too (slow: K^2 dispatches per pair)
        got2 = do.build_gmm_distance_array(W, MU, COV, vectorised=False)
        np.testing.assert_allclose(got2, D, rtol=1e-12, atol=1e-14)


def test_analytic_cases():
    W, MU, COV = do.random_gmms(2, 8, seed=3)
    # identical mixtures: every diagonal Hellinger is 0 -> the identity assignment costs 0
    d = do.distribution_distance(W[0], MU[0], COV[0], W[0], MU[0], COV[0])
    assert d < 1e-6
    # far apart: all Hellinger = 1 -> sum_i (w1_i + w2_sigma(i)) = 2   (readme.md:205-216 range [0, 2])
    d = do.distribution_distance(W[0], MU[0], COV[0], W[1], MU[1] + 1e6, COV[1])
    assert abs(d - 2.0) < 1e-12
    # symmetric
    a = do.distribution_distance(W[0], MU[0], COV[0], W[1], MU[1], COV[1])
    b = do.distribution_distance(W[1], MU[1], COV[1], W[0], MU[0], COV[0])
    assert abs(a - b) < 1e-12
