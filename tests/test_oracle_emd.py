"""CPU: the network-simplex EMD oracle against exact LP optima (HiGHS) — committed golden values and live."""
import os

import numpy as np
import pytest
from scipy.sparse import csr_matrix

from oracle import emd_oracle as eo

GOLD = os.path.join(os.path.dirname(__file__), "golden", "emd_small.npz")


def cases():
    z = np.load(GOLD)
    for c in range(int(z["n"])):
        p = "c%d_" % c
        yield c, {k[len(p):]: z[k] for k in z.files if k.startswith(p)}


@pytest.mark.parametrize("c,d", list(cases()))
def test_network_simplex_matches_lp_golden(c, d):
    cost, status, pivots = eo.emd_network_simplex(d["sx"], d["sy"], d["sm"], d["tx"], d["ty"], d["tm"])
    assert status == 0
    assert abs(cost - float(d["cost"])) <= 1e-9 * max(1.0, float(d["cost"]))
    # the pivot rule's block size does not change the optimum
    cost2, status2, _ = eo.emd_network_simplex(d["sx"], d["sy"], d["sm"], d["tx"], d["ty"], d["tm"], block_size=17)
    assert status2 == 0 and abs(cost2 - cost) <= 1e-11 * max(1.0, cost)
    # neither does the candidate-list variant of the pivot rule
    cost3, status3, _ = eo.emd_network_simplex(d["sx"], d["sy"], d["sm"], d["tx"], d["ty"], d["tm"], cand_threads=256)
    assert status3 == 0 and abs(cost3 - cost) <= 1e-11 * max(1.0, cost)


def test_network_simplex_matches_lp_live():
    rng = np.random.default_rng(5)
    s = rng.permutation(144)[:60]; t = rng.permutation(144)[:45]
    sm = rng.random(60) + 0.1; tm = rng.random(45) + 0.1
    a = eo.emd_network_simplex(s // 12, s % 12, sm, t // 12, t % 12, tm)[0]
    b = eo.emd_linprog(s // 12, s % 12, sm, t // 12, t % 12, tm)
    assert abs(a - b) < 1e-9 * b


def test_iteration_limit_reports_status_1():
    """POT stops at numItermax and returns the current (sub-optimal) value with a warning (distance.py:201)."""
    _, d = list(cases())[3]
    full, st0, pivots = eo.emd_network_simplex(d["sx"], d["sy"], d["sm"], d["tx"], d["ty"], d["tm"])
    cut, st1, it1 = eo.emd_network_simplex(d["sx"], d["sy"], d["sm"], d["tx"], d["ty"], d["tm"], max_iter=pivots // 2)
    assert st0 == 0 and st1 == 1 and it1 == pivots // 2


def test_calculate_ot_distance_staging_and_properties():
    """distance.py:191-202 staging from sparse images; EMD of an image with itself is 0; a translation by
    (dx, dy) costs dx^2 + dy^2; scaling the masses changes nothing (both sides are normalised)."""
    rng = np.random.default_rng(1)
    img = np.zeros((12, 15)); idx = rng.permutation(180)[:40]
    img[idx // 15, idx % 15] = rng.gamma(1.0, 2.0, 40)
    a = csr_matrix(img)
    assert eo.calculate_ot_distance(a, a) < 1e-12
    shifted = np.zeros((16, 18)); shifted[3:15, 2:17] = img
    assert abs(eo.calculate_ot_distance(a, csr_matrix(shifted)) - (3 ** 2 + 2 ** 2)) < 1e-9
    b_img = np.zeros((12, 15)); j = rng.permutation(180)[:25]; b_img[j // 15, j % 15] = rng.random(25) + 0.05
    d1 = eo.calculate_ot_distance(a, csr_matrix(b_img))
    d2 = eo.calculate_ot_distance(csr_matrix(7.5 * img), csr_matrix(0.2 * b_img))
    assert abs(d1 - d2) < 1e-9 * d1


GOLD_LARGE = os.path.join(os.path.dirname(__file__), "golden", "emd_large.npz")


@pytest.mark.parametrize("c,d", list(cases()))
def test_multiscale_start_reaches_the_same_optimum(c, d):
    cost, status, pivots, lev = eo.emd_multiscale(d["sx"], d["sy"], d["sm"], d["tx"], d["ty"], d["tm"])
    assert status == 0 and pivots == lev.sum()
    assert abs(cost - float(d["cost"])) <= 1e-9 * max(1.0, float(d["cost"]))
    cost2, status2, _, _ = eo.emd_multiscale(d["sx"], d["sy"], d["sm"], d["tx"], d["ty"], d["tm"], block_size=23)
    assert status2 == 0 and abs(cost2 - cost) <= 1e-11 * max(1.0, cost)
    # the pivot rules of the CUDA kernels (candidate list + growing blocks for 256 / 512 threads per problem, the
    # global-tree kernel's variant) change the pivots, never the optimum
    for T in (256, 512, -1):
        cost3, status3, _, _ = eo.emd_multiscale(d["sx"], d["sy"], d["sm"], d["tx"], d["ty"], d["tm"], cand_threads=T)
        assert status3 == 0 and abs(cost3 - cost) <= 1e-11 * max(1.0, cost), T


def test_oracle_matches_visium_size_lp_golden():
    """The oracle (multiscale start; the star start takes ~10 s per problem) against the independent Visium-size LP
    optima of tests/golden/emd_large.npz (HiGHS + column generation + full dual-feasibility certificate)."""
    z = np.load(GOLD_LARGE)
    for c in range(int(z["n"])):
        p = "c%d_" % c
        cost, status, pivots, lev = eo.emd_multiscale(z[p + "sx"], z[p + "sy"], z[p + "sm"], z[p + "tx"], z[p + "ty"], z[p + "tm"])
        ref = float(z[p + "cost"])
        assert status == 0
        assert abs(cost - ref) <= 1e-7 * ref, (c, cost, ref)
        # with the pricing of the kernel that solves a problem of this size (fewer pivots, same optimum)
        T = 512 if len(z[p + "sx"]) + len(z[p + "tx"]) < 11000 else -1
        cost_c, status_c, pivots_c, _ = eo.emd_multiscale(z[p + "sx"], z[p + "sy"], z[p + "sm"], z[p + "tx"], z[p + "ty"], z[p + "tm"],
                                                          cand_threads=T)
        assert status_c == 0 and abs(cost_c - ref) <= 1e-7 * ref and pivots_c < pivots, (c, cost_c, ref, pivots_c, pivots)


def _edge_problems():
    rng = np.random.default_rng(11)
    out = {}
    # duplicate coordinates on both sides (a general support, not a CSR image): cells of the hierarchy with > 4 children
    s = rng.integers(0, 36, 90); t = rng.integers(0, 36, 70)
    out["duplicates"] = ((s // 6, s % 6, rng.random(90) + 0.05), (t // 6, t % 6, rng.random(70) + 0.05))
    # targets outside the source's bounding box, left / below its corner (negative coordinates relative to it)
    s = rng.permutation(100)[:60]; t = rng.permutation(100)[:40]
    out["outside"] = ((s // 10 + 20, s % 10 + 30, rng.random(60) + 0.1), (t // 10, t % 10 + 25, rng.random(40) + 0.1))
    # masses so small that their integer image is 0 (a node without any arc in the refined forest)
    s = rng.permutation(144)[:80]; t = rng.permutation(144)[:50]
    sm = rng.random(80) + 0.1; tm = rng.random(50) + 0.1
    sm[3] = 1e-22; tm[7] = 1e-22; tm[8] = 1e-25
    out["tiny_mass"] = ((s // 12, s % 12, sm), (t // 12, t % 12, tm))
    # one source spot, one target spot, and a single row of cells
    out["one_to_many"] = ((np.array([4]), np.array([5]), np.array([2.0])), (t // 12, t % 12, tm.copy()))
    out["line"] = ((np.arange(70), np.zeros(70, dtype=int), rng.random(70) + 0.1), (np.arange(0, 70, 2), np.zeros(35, dtype=int), rng.random(35) + 0.1))
    # a target too far left of the source's corner for the Morton keys (< -16384): the solver starts from the star
    out["far"] = ((s // 12, s % 12, sm.copy()), (t // 12 - 20000, t % 12, tm.copy()))
    return out


@pytest.mark.parametrize("name", list(_edge_problems()))
def test_multiscale_edge_cases_match_lp(name):
    src, tgt = _edge_problems()[name]
    cost, status, pivots, lev = eo.emd_multiscale(*src, *tgt)
    assert status == 0
    ref = eo.emd_linprog(*src, *tgt)
    assert abs(cost - ref) <= 1e-8 * max(1.0, ref), (name, cost, ref)
    star = eo.emd_network_simplex(*src, *tgt)[0]
    assert abs(cost - star) <= 1e-11 * max(1.0, star)
    # the global-tree kernel's rule searches on the low 32 bits of the reduced costs; nodes hanging on artificial arcs
    # (potentials +-BIG: "tiny_mass") make those wrap, and the 64-bit re-pricing / certificate sweep must catch it
    for T in (256, -1):
        cost_t, status_t, _, _ = eo.emd_multiscale(*src, *tgt, cand_threads=T)
        assert status_t == 0 and abs(cost_t - star) <= 1e-11 * max(1.0, star), (name, T, cost_t, star)
