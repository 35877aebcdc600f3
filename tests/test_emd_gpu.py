"""GPU parity: batched exact spot-level EMD (through the C ABI) vs golden LP optima and the network-simplex oracle."""
import os

import numpy as np
import pytest
from scipy.sparse import csr_matrix

from oracle import emd_oracle as eo

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "emd_small.npz")


def test_emd_matches_lp_golden(cuda_device):
    from stminer_b200.Algorithm.distance import emd_batched
    z = np.load(GOLD)
    for c in range(int(z["n"])):
        p = "c%d_" % c
        cost, status = emd_batched((z[p + "sx"], z[p + "sy"], z[p + "sm"]), [(z[p + "tx"], z[p + "ty"], z[p + "tm"])])
        assert status[0] == 0
        # north-star tolerance for global_distance is rtol 1e-4; the integer simplex is exact to ~1e-12
        assert abs(cost[0] - float(z[p + "cost"])) <= 1e-9 * max(1.0, float(z[p + "cost"]))


def _random_problem(rng, n, ms, R):
    s = rng.permutation(R * R)[:n]
    src = (s // R, s % R, rng.gamma(1.0, 1.0, n) + 0.01)
    tg = []
    for m in ms:
        t = rng.permutation(R * R)[:m]
        tg.append((t // R, t % R, rng.gamma(1.0, 1.0, m) + 0.01))
    return src, tg


@pytest.mark.parametrize("n,ms,R", [(282, [150, 282, 31, 1, 90], 17), (1200, [700, 400, 1100], 40), (50, list(range(1, 40, 3)), 8)])
def test_emd_batch_matches_oracle_and_pivot_sequence(cuda_device, n, ms, R):
    """One source, ragged targets in one launch; same block size -> identical pivot count and cost as the oracle."""
    from stminer_b200.Algorithm.distance import emd_batched, emd_pricing_threads
    src, tg = _random_problem(np.random.default_rng(n), n, ms, R)
    # default block size: optimum must agree
    cost0, status0 = emd_batched(src, tg)
    cost, status, iters = emd_batched(src, tg, block_size=777, return_iters=True, star_start=True, candidates=False)
    cost_c, status_c, iters_c = emd_batched(src, tg, block_size=777, return_iters=True, star_start=True)
    T = emd_pricing_threads(src, tg)
    for g, t in enumerate(tg):
        c, st, it = eo.emd_network_simplex(*src, *t, block_size=777)
        assert status0[g] == 0 and abs(cost0[g] - c) <= 1e-12 * max(1.0, c)
        assert status[g] == st == 0
        assert iters[g] == it
        assert abs(cost[g] - c) <= 1e-12 * max(1.0, c)
        # candidate-list pricing (the default; active where the block is whole target rows): the oracle's restatement
        c2, st2, it2 = eo.emd_network_simplex(*src, *t, block_size=777, cand_threads=int(T[g]))
        assert status_c[g] == st2 == 0 and iters_c[g] == it2, (g, iters_c[g], it2)
        assert abs(cost_c[g] - c) <= 1e-12 * max(1.0, c)


def test_emd_iteration_limit_and_empty(cuda_device):
    from stminer_b200.Algorithm.distance import emd_batched
    src, tg = _random_problem(np.random.default_rng(3), 282, [150], 17)
    full, st, it = emd_batched(src, tg, block_size=500, return_iters=True, star_start=True)
    cut, st1, it1 = emd_batched(src, tg, max_iter=int(it[0]) // 2, block_size=500, return_iters=True, star_start=True)
    ref, rst, rit = eo.emd_network_simplex(*src, *tg[0], max_iter=int(it[0]) // 2, block_size=500)
    assert st[0] == 0 and st1[0] == 1 and rst == 1 and it1[0] == rit
    assert abs(cut[0] - ref) <= 1e-12 * ref          # same sub-optimal basis as the oracle (POT semantics)
    empty = (np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int32), np.zeros(0))
    c2, s2 = emd_batched(src, [tg[0], empty])
    assert s2.tolist() == [0, 2]


def test_emd_beyond_one_cta_runs_with_the_tree_in_global_memory(cuda_device):
    """Problems whose tree does not fit a CTA's shared memory (> ~11.9k nodes: Visium HD, Stereo-seq bins) run the same
    solver with the tree in the CTA's global scratch: same optimum and — at an explicit block size — the same pivots per
    level as the oracle.  A batch that mixes sizes is split by the host; results stay in target order."""
    from stminer_b200.Algorithm.distance import emd_batched, emd_pricing_threads
    rng = np.random.default_rng(0)
    n = 9000
    cells = rng.permutation(150 * 150)[:n]
    src = (cells // 150, cells % 150, rng.random(n) + 0.1)
    tgs = []
    for m in (4000, 700, 5200):
        t = rng.permutation(150 * 150)[:m]
        tgs.append((t // 150, t % 150, rng.random(m) + 0.1))
    cost, status, iters = emd_batched(src, tgs, return_iters=True)
    costB, statusB, itersB = emd_batched(src, tgs, block_size=40000, return_iters=True)
    T = emd_pricing_threads(src, tgs)          # -1 for the global-tree kernel, 512 for the target that fits shared memory
    assert sorted(T.tolist()) == [-1, -1, 512]
    costR, statusR, itersR = emd_batched(src, tgs, block_size=8 * n, return_iters=True)      # whole target rows
    for g, t in enumerate(tgs):
        c, st, it, lev = eo.emd_multiscale(*src, *t, block_size=40000, cand_threads=int(T[g]))
        assert status[g] == statusB[g] == st == 0, (g, status, statusB)
        assert abs(cost[g] - c) <= 1e-11 * c and abs(costB[g] - c) <= 1e-11 * c, (g, cost[g], costB[g], c)
        assert itersB[g] == it, (g, itersB[g], it, lev[:8])
        # blocks of whole target rows: the global-tree kernel lengthens them over fruitless scans, the oracle likewise
        c2, st2, it2, lev2 = eo.emd_multiscale(*src, *t, block_size=8 * n, cand_threads=int(T[g]))
        assert statusR[g] == st2 == 0 and itersR[g] == it2, (g, itersR[g], it2, lev2[:8])
        assert abs(costR[g] - c) <= 1e-11 * c


def test_emd_global_tree_32bit_search_survives_wrapping_reduced_costs(cuda_device):
    """The global-tree kernel searches on the low 32 bits of the reduced costs.  Targets whose integer mass is 0 hang on an
    artificial arc with potential +BIG (~1e13 here), so the reduced costs of their arcs do not fit 32 bits: whatever the
    wrapped search makes of them, the 64-bit re-pricing of the selected arc and the 64-bit certificate sweep keep the result
    exact — same optimum and same pivots as the oracle's restatement (cand_threads = -1)."""
    from stminer_b200.Algorithm.distance import emd_batched, emd_pricing_threads
    rng = np.random.default_rng(5)
    n, m = 9000, 4200
    cells = rng.permutation(150 * 150)[:n]
    src = (cells // 150, cells % 150, rng.random(n) + 0.1)
    t = rng.permutation(150 * 150)[:m]
    tm = rng.random(m) + 0.1
    tm[[17, 2900, 4100]] = 1e-26                       # integer image 0: no arc in any refined forest
    tgt = (t // 150, t % 150, tm)
    assert emd_pricing_threads(src, [tgt]).tolist() == [-1]
    cost, status, iters = emd_batched(src, [tgt], block_size=8 * n, return_iters=True)
    c, st, it, lev = eo.emd_multiscale(*src, *tgt, block_size=8 * n, cand_threads=-1)
    c0, st0, _, _ = eo.emd_multiscale(*src, *tgt)                                   # plain 64-bit block search
    assert status[0] == st == st0 == 0
    assert abs(cost[0] - c0) <= 1e-11 * c0 and abs(c - c0) <= 1e-11 * c0
    assert iters[0] == it, (iters[0], it, lev[:8])
    cost_d, status_d = emd_batched(src, [tgt])                                      # the kernel's default blocks
    assert status_d[0] == 0 and abs(cost_d[0] - c0) <= 1e-11 * c0


def test_emd_node_capacity_limit_reports_status_3(cuda_device):
    """Beyond the global-tree capacity (or 2^31 arcs) a problem is refused with status 3 and NaN, never a wrong value."""
    from stminer_b200 import _lib
    from stminer_b200.Algorithm.distance import emd_batched
    cap = int(_lib.load().stm_emd_global_node_capacity())
    rng = np.random.default_rng(1)
    n = cap - 100
    src = (np.arange(n) // 400, np.arange(n) % 400, rng.random(n) + 0.1)
    tgt = (np.arange(300) // 400, np.arange(300) % 400, rng.random(300) + 0.1)
    cost, status = emd_batched(src, [tgt])
    assert status[0] == 3 and np.isnan(cost[0])


def test_host_api_csr_images(cuda_device):
    """calculate_ot_distance / build_ot_distance_array on sparse (x, y) images (distance.py:179-202)."""
    from stminer_b200.Algorithm.distance import build_ot_distance_array, calculate_ot_distance
    rng = np.random.default_rng(2)
    imgs = {}
    for g in range(5):
        a = np.zeros((14, 11)); idx = rng.permutation(154)[:30 + 5 * g]
        a[idx // 11, idx % 11] = rng.gamma(1.0, 2.0, len(idx))
        imgs["g%d" % g] = csr_matrix(a)
    d01 = calculate_ot_distance(imgs["g0"], imgs["g1"])
    assert abs(d01 - eo.calculate_ot_distance(imgs["g0"], imgs["g1"])) < 1e-10 * d01
    assert calculate_ot_distance(imgs["g2"], imgs["g2"]) < 1e-12
    df = build_ot_distance_array(imgs)
    assert list(df.index) == list(imgs) and np.allclose(df.values, df.values.T) and (np.diag(df.values) == 0).all()
    assert abs(df.loc["g0", "g1"] - d01) < 1e-12
    ref34 = eo.calculate_ot_distance(imgs["g3"], imgs["g4"])
    assert abs(df.loc["g3", "g4"] - ref34) < 1e-10 * ref34


@pytest.mark.parametrize("n,m", [(250, 200), (600, 450), (640, 640), (2500, 1900), (5000, 3000)])
def test_emd_one_dimensional_images_deep_trees(cuda_device, n, m):
    """Spots on a single grid line: the optimal plan is monotone and the basis trees are long chains, i.e. long cycle
    paths — the stress case for the parallel path extraction (ordered compaction over many threads).  Same pivot count
    and cost as the oracle; the 1-D optimum also has a closed form (monotone coupling).  The two longest lines exceed
    the int32 potentials of the default kernel ((span^2 + 1) * nodes >= 2^30) and run on the 64-bit variant."""
    from stminer_b200.Algorithm.distance import emd_batched, emd_pricing_threads
    rng = np.random.default_rng(n)
    span = n + n // 3
    sx = np.sort(rng.choice(span, size=n, replace=False)).astype(np.int32)
    tx = np.sort(rng.choice(span, size=m, replace=False)).astype(np.int32)
    src = (sx, np.zeros(n, dtype=np.int32), rng.gamma(2.0, 1.0, n) + 0.05)
    tgt = (tx, np.full(m, 3, dtype=np.int32), rng.gamma(2.0, 1.0, m) + 0.05)
    cost, status, iters = emd_batched(src, [tgt], return_iters=True, block_size=4 * n, star_start=True, candidates=False)
    c, st, it = eo.emd_network_simplex(*src, *tgt, block_size=4 * n)
    assert status[0] == st == 0 and iters[0] == it
    assert abs(cost[0] - c) <= 1e-12 * max(1.0, c)
    T = int(emd_pricing_threads(src, [tgt])[0])                   # 256 for the short lines, 0 on the 64-bit kernel
    cost_c, status_c, iters_c = emd_batched(src, [tgt], return_iters=True, block_size=4 * n, star_start=True)
    c2, st2, it2 = eo.emd_network_simplex(*src, *tgt, block_size=4 * n, cand_threads=T)
    assert status_c[0] == st2 == 0 and iters_c[0] == it2 and abs(cost_c[0] - c) <= 1e-12 * max(1.0, c)
    if True:           # monotone (north-west corner) coupling of the sorted supports is optimal for a convex cost on a line
        a, b = src[2] / src[2].sum(), tgt[2] / tgt[2].sum()
        i = j = 0
        ra, rb, tot = a[0], b[0], 0.0
        while i < n and j < m:
            f = min(ra, rb)
            tot += f * ((int(sx[i]) - int(tx[j])) ** 2 + 9)
            ra -= f; rb -= f
            if ra <= 1e-18 and i < n - 1: i += 1; ra = a[i]
            elif rb <= 1e-18 and j < m - 1: j += 1; rb = b[j]
            else: break
        assert abs(cost[0] - tot) <= 1e-6 * tot


def test_emd_batch_mixes_int32_and_64bit_problems(cuda_device):
    """One launch request whose targets need different potential widths: a target far away from the source makes
    (span^2 + 1) * nodes overflow int32, so the host routes it to stm_emd_spots_batched_wide; results stay in target
    order and both match the oracle (cost and pivot count)."""
    from stminer_b200.Algorithm.distance import emd_batched, emd_pricing_threads
    rng = np.random.default_rng(9)
    src, tg = _random_problem(rng, 300, [200, 150, 260], 40)
    far = (tg[1][0] + 3000, tg[1][1] + 1700, tg[1][2])
    targets = [tg[0], far, tg[2]]
    cost, status, iters = emd_batched(src, targets, block_size=500, return_iters=True, star_start=True)
    T = emd_pricing_threads(src, targets)
    assert T.tolist() == [256, 0, 256]
    for g, t in enumerate(targets):
        c, st, it = eo.emd_network_simplex(*src, *t, block_size=500, cand_threads=int(T[g]))
        assert status[g] == st == 0 and iters[g] == it, g
        assert abs(cost[g] - c) <= 1e-12 * max(1.0, c)
    assert cost[1] > 1e6 > cost[0]


def test_emd_one_cta_per_sm_kernel_matches_oracle(cuda_device):
    """A problem large enough for the one-CTA-per-SM kernel with int32 potentials and the register-
    resident whole-row pricing path: cost AND pivot count of the oracle at the same block size (9 target rows)."""
    from stminer_b200.Algorithm.distance import emd_batched, emd_pricing_threads
    rng = np.random.default_rng(11)
    n, m, R = 3600, 2200, 60
    src, tg = _random_problem(rng, n, [m], R)
    B = 9 * n
    cost, status, iters = emd_batched(src, tg, block_size=B, return_iters=True, star_start=True, candidates=False)
    c, st, it = eo.emd_network_simplex(*src, *tg[0], block_size=B)
    assert status[0] == st == 0 and iters[0] == it
    assert abs(cost[0] - c) <= 1e-12 * max(1.0, c)
    # with the candidate list (default): the pivots of the oracle pricing like 512 threads, fewer than the plain search
    assert emd_pricing_threads(src, tg).tolist() == [512]
    cost_c, status_c, iters_c = emd_batched(src, tg, block_size=B, return_iters=True, star_start=True)
    c2, st2, it2 = eo.emd_network_simplex(*src, *tg[0], block_size=B, cand_threads=512)
    assert status_c[0] == st2 == 0 and iters_c[0] == it2, (iters_c[0], it2)
    assert abs(cost_c[0] - c) <= 1e-12 * max(1.0, c)
    cost_d, status_d = emd_batched(src, tg)                       # the kernel's own default block size: same optimum
    assert status_d[0] == 0 and abs(cost_d[0] - c) <= 1e-12 * max(1.0, c)


# ---------------------------------------------------------------------------------------------------------------------
# multiscale start basis (the default): same optimum, and the same pivots as the oracle's restatement of the scheme
# ---------------------------------------------------------------------------------------------------------------------
GOLD_LARGE = os.path.join(os.path.dirname(__file__), "golden", "emd_large.npz")


@pytest.mark.parametrize("n,ms,R,B", [(282, [150, 282, 31, 1, 90], 17, 777), (1200, [700, 400, 1100], 40, 5000),
                                      (50, list(range(1, 40, 3)), 8, 100), (3600, [2200, 900], 60, 9 * 3600)])
def test_emd_multiscale_matches_oracle_levels(cuda_device, n, ms, R, B):
    """Default start (Morton hierarchy, coarse-to-fine refinement): cost, status and the pivot count over all levels
    equal the oracle's stm_oracle_emd_ms for the same block size (with the candidate list of the default pricing, and
    with the plain block search); the optimum equals the star start's."""
    from stminer_b200.Algorithm.distance import emd_batched, emd_pricing_threads
    src, tg = _random_problem(np.random.default_rng(n + 1), n, ms, R)
    cost, status, iters = emd_batched(src, tg, block_size=B, return_iters=True)
    cost_p, status_p, iters_p = emd_batched(src, tg, block_size=B, return_iters=True, candidates=False)
    cost_d, status_d = emd_batched(src, tg)                                # the kernel's default block sizes
    T = emd_pricing_threads(src, tg)
    for g, t in enumerate(tg):
        c, st, it, lev = eo.emd_multiscale(*src, *t, block_size=B, cand_threads=int(T[g]))
        c_p, st_p, it_p, lev_p = eo.emd_multiscale(*src, *t, block_size=B)      # plain block search
        assert status_p[g] == st_p == 0 and iters_p[g] == it_p, (g, iters_p[g], it_p, lev_p[:8])
        assert abs(cost_p[g] - c_p) <= 1e-12 * max(1.0, c_p)
        c_star, st_star, _ = eo.emd_network_simplex(*src, *t)
        assert status[g] == st == 0 and status_d[g] == 0 and st_star == 0
        assert iters[g] == it, (g, iters[g], it, lev[:8])
        assert abs(cost[g] - c) <= 1e-12 * max(1.0, c)
        assert abs(cost[g] - c_star) <= 1e-11 * max(1.0, c_star)
        assert abs(cost_d[g] - c_star) <= 1e-11 * max(1.0, c_star)


def test_emd_multiscale_iteration_limit_applies_to_the_problem_itself(cuda_device):
    """numItermax bounds the pivots on the original problem (finest level); the coarse levels always run to their
    optimum.  Same sub-optimal basis cost as the oracle, status 1."""
    from stminer_b200.Algorithm.distance import emd_batched, emd_pricing_threads
    src, tg = _random_problem(np.random.default_rng(3), 600, [400], 30)
    T = int(emd_pricing_threads(src, tg)[0])
    c_full, st_full, it_full, lev = eo.emd_multiscale(*src, *tg[0], block_size=900, cand_threads=T)
    assert st_full == 0 and lev[0] > 20
    cap = int(lev[0]) // 2
    cut, st1, it1 = emd_batched(src, tg, max_iter=cap, block_size=900, return_iters=True)
    c_cut, st_cut, it_cut, lev_cut = eo.emd_multiscale(*src, *tg[0], max_iter=cap, block_size=900, cand_threads=T)
    assert st1[0] == st_cut == 1 and lev_cut[0] == cap and it1[0] == it_cut
    assert abs(cut[0] - c_cut) <= 1e-12 * c_cut and cut[0] > c_full


def test_emd_visium_size_matches_independent_lp_golden(cuda_device):
    """Visium-size problems (4,992 source spots x 800-3,300 target spots) against optima computed by HiGHS through
    column generation with a full n x m dual-feasibility certificate (tests/golden/make_golden_emd_large.py) — an
    independent solver, neither the kernel nor oracle/emd_oracle.c.  Both start bases; rtol 1e-7 = the LP solver's
    feasibility tolerance (the north-star bar for global_distance is rtol 1e-4)."""
    from stminer_b200.Algorithm.distance import emd_batched
    z = np.load(GOLD_LARGE)
    for c in range(int(z["n"])):
        p = "c%d_" % c
        src = (z[p + "sx"], z[p + "sy"], z[p + "sm"]); tgt = (z[p + "tx"], z[p + "ty"], z[p + "tm"])
        ref = float(z[p + "cost"])
        cost, status = emd_batched(src, [tgt])      # the last case (12,979 x 3,000) runs with the tree in global memory
        assert status[0] == 0, (c, status)
        assert abs(cost[0] - ref) <= 1e-7 * ref, (c, cost[0], ref)
        cost_s, status_s = emd_batched(src, [tgt], star_start=True)
        if len(src[0]) + len(tgt[0]) < 11000:
            assert status_s[0] == 0 and abs(cost_s[0] - cost[0]) <= 1e-11 * ref
        else:
            # from the artificial star (the start of POT, which the reference calls with numItermax=200000) the 16k-node
            # problem is still sub-optimal after 200,000 pivots: status 1 and a larger cost, as POT would return
            assert status_s[0] == 1 and cost_s[0] > cost[0]


def test_emd_multiscale_edge_cases(cuda_device):
    """Supports the reference's CSR images never produce but the API accepts: duplicate coordinates, targets outside the
    source's bounding box, masses whose integer image is 0, single spots, a single row of cells, coordinates beyond the
    range of the Morton keys (star fallback) — optimum of the independent LP, pivots of the oracle."""
    from stminer_b200.Algorithm.distance import emd_batched, emd_pricing_threads
    from tests.test_oracle_emd import _edge_problems
    for name, (src, tgt) in _edge_problems().items():
        cost, status, iters = emd_batched(src, [tgt], block_size=300, return_iters=True)
        c, st, it, lev = eo.emd_multiscale(*src, *tgt, block_size=300, cand_threads=int(emd_pricing_threads(src, [tgt])[0]))
        assert status[0] == st == 0, (name, status)
        assert iters[0] == it, (name, iters[0], it, lev[:6])
        assert abs(cost[0] - c) <= 1e-12 * max(1.0, c), (name, cost[0], c)
        ref = eo.emd_linprog(*src, *tgt)
        assert abs(cost[0] - ref) <= 1e-8 * max(1.0, ref), (name, cost[0], ref)
