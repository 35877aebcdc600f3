"""GPU parity: the warp LSAP on arbitrary cost matrices vs scipy.optimize.linear_sum_assignment
(the solver behind the reference's linear_sum, STMiner/Algorithm/algorithm.py:10-26)."""
import numpy as np
import pytest
from scipy.optimize import linear_sum_assignment

pytestmark = pytest.mark.gpu


def _ref(C):
    r, c = linear_sum_assignment(C)
    return C[r, c].sum()


@pytest.mark.parametrize("K", [1, 2, 3, 7, 10, 20, 31, 32, 33, 48, 64])
def test_lsap_random_and_adversarial(cuda_device, K):
    from stminer_b200.Algorithm.algorithm import linear_sum_batched
    rng = np.random.default_rng(K)
    mats = [rng.random((K, K)) for _ in range(40)]
    mats += [rng.integers(0, 4, (K, K)).astype(float) for _ in range(40)]          # heavy ties
    mats += [np.outer(rng.random(K), rng.random(K)) for _ in range(10)]            # rank-1: many equal optima
    mats += [np.ones((K, K)), np.zeros((K, K)), np.eye(K), 1 - np.eye(K)]
    mats += [rng.random((K, K)) * 1e6 - 5e5 for _ in range(10)]                     # negative costs
    a = rng.random((K, K)); mats.append(a + a.T)
    m = np.tile(rng.random(K), (K, 1)); mats.append(m); mats.append(m.T.copy())      # identical rows / columns
    got = linear_sum_batched(np.stack(mats)).cpu().numpy()
    ref = np.array([_ref(C) for C in mats])
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-9)


@pytest.mark.parametrize("K", [2, 5, 20, 40])
def test_lsap_forbidden_entries(cuda_device, K):
    """+inf entries are forbidden assignments (scipy accepts them); a matrix without a finite perfect matching is
    infeasible: scipy raises ValueError, the kernel reports NaN."""
    from stminer_b200.Algorithm.algorithm import linear_sum_batched
    rng = np.random.default_rng(100 + K)
    mats = []
    for _ in range(30):
        C = rng.random((K, K))
        C[rng.random((K, K)) < 0.4] = np.inf
        perm = rng.permutation(K)
        C[np.arange(K), perm] = rng.random(K)          # one finite perfect matching is guaranteed
        mats.append(C)
    got = linear_sum_batched(np.stack(mats)).cpu().numpy()
    ref = np.array([_ref(C) for C in mats])
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-12)
    bad = rng.random((K, K)); bad[:, 0] = np.inf       # column 0 cannot be used by anyone
    bad2 = rng.random((K, K)); bad2[0, :] = np.inf
    out = linear_sum_batched(np.stack([bad, bad2])).cpu().numpy()
    assert np.isnan(out).all()
    with pytest.raises(ValueError):
        linear_sum_assignment(bad)


def test_linear_sum_drop_in(cuda_device):
    from stminer_b200.Algorithm.algorithm import linear_sum
    C = np.array([[4.0, 1, 3], [2, 0, 5], [3, 2, 2]])
    assert linear_sum(C) == 5.0
    with pytest.raises(ValueError):
        linear_sum(np.array([[1.0, np.nan], [0.0, 1.0]]))
