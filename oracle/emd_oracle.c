/*
 * emd_oracle.c — CPU oracle for the spot-level exact EMD (hot path 2a).  TEST INFRASTRUCTURE ONLY.
 *
 * Restates  calculate_ot_distance  STMiner/Algorithm/distance.py:191-202:
 *     xs = source.nonzero().T ; xt = target.nonzero().T           (integer grid coordinates)
 *     a  = source.data[>0] / source.data.sum() ; b likewise
 *     M  = ot.dist(xs, xt)                                        (squared Euclidean, integer valued)
 *     ot.emd2(a, b, M, numItermax=200000)                         (exact LP optimum  sum_ij G_ij M_ij)
 *
 * The LP solver lives in a third-party dependency that is NOT under /root/reference and is not
 * installable in this image: POT (Python Optimal Transport), pinned POT==0.9.1 (requirements.txt:10).
 * POT's emd2 is a primal network simplex (LEMON-derived) on the bipartite transport graph with a block-search
 * pivot rule.  This file restates that published algorithm — spanning-tree basis rooted at an artificial
 * node, block search for the entering arc, strongly-feasible leaving-arc rule, potential update of the
 * re-hung subtree — in a form chosen so the CUDA kernel (stminer_b200/csrc/emd_spots.cu) can run the
 * IDENTICAL pivot sequence:
 *   - masses are scaled to integers (total 2^50) so every flow comparison is exact,
 *   - costs are integers computed on the fly from the coordinates (the n x m matrix is never built),
 *   - the tree is kept as parent / preorder-position / subtree-size arrays (a subtree is a contiguous
 *     segment of the preorder array), which is what makes the tree update data-parallel on the GPU.
 *
 * Two start bases (the LP optimum — the value the reference returns — does not depend on the start):
 *   - STAR  (stm_oracle_emd, or flags & 1): the artificial star POT/LEMON start from;
 *   - MULTISCALE (stm_oracle_emd_ms, default of the CUDA kernel): the nodes are sorted along the Morton curve, the
 *     grid is coarsened one Morton bit per level (cells halved along y, then x; levels that merge < 20 % of the nodes
 *     are skipped), the coarsest problem is solved from the star and every
 *     finer level starts from the refinement of the coarser optimum (two north-west-corner passes: the flow of a
 *     coarse arc is dealt to the children of its source cell, then to the children of its target cell).  The
 *     finest level is the original problem, solved to optimality by the same simplex — about ten times fewer pivots.
 *
 * Pinning — PARITY UNPINNED against the reference itself: POT cannot be imported here and the reference holds no golden
 * value for this path that is usable here (its only pin is the top-2
 * gene ranking on the missing tests/data/test.h5ad, tests/test_STMiner.py:48-49).  The oracle is pinned
 * against an independent exact LP solver (scipy.optimize.linprog, HiGHS): small problems live and as committed golden
 * values tests/golden/emd_small.npz, Visium-size problems (4,992 x 800..3,300) as tests/golden/emd_large.npz (HiGHS by
 * column generation with a full n x m dual-feasibility certificate, tests/golden/make_golden_emd_large.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define STM_EMD_SCALE 1125899906842624.0 /* 2^50 */
#define STM_EMD_MAXLEV 23  /* bit levels 0 .. 23: cells up to 2^11 x 2^12 */
#define STM_EMD_MAXUSE 16  /* levels in use per problem at most */
#define STM_EMD_SKIP_NUM 4   /* a bit level that keeps more than 4/5 of the nodes of the level below is skipped */
#define STM_EMD_SKIP_DEN 5
#define STM_EMD_COARSEST_NODES 64
#define STM_EMD_GROW_G 4 /* the same for the global-tree kernel (cand_threads < 0): blocks of up to 64 target rows */
#define STM_EMD_GROW 3   /* a pricing block after k empty ones is 2^min(k, 3) times as long (candidate-list pricing) */

typedef struct {
  int n, m, N, root;
  const int32_t *sx, *sy, *tx, *ty;
  int32_t* parent;
  int32_t* pos;
  int32_t* size;
  int32_t* order;
  int32_t* tmp;
  int64_t* flow;
  int64_t* pi;
} Net;

static inline int64_t arc_cost(const Net* g, int i, int j) {
  const int64_t dx = (int64_t)g->sx[i] - g->tx[j], dy = (int64_t)g->sy[i] - g->ty[j];
  return dx * dx + dy * dy;
}
static inline int is_anc(const Net* g, int u, int w) { /* u is an ancestor of (or equal to) w */
  return g->pos[u] <= g->pos[w] && g->pos[w] < g->pos[u] + g->size[u];
}

/* Integer masses with total exactly 2^50 on both sides (the largest entry absorbs the rounding). */
static void integerize(const double* mass, int n, int64_t* out) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += mass[i];
  int64_t tot = 0;
  int big = 0;
  for (int i = 0; i < n; ++i) {
    out[i] = (int64_t)floor(mass[i] / s * STM_EMD_SCALE + 0.5);
    tot += out[i];
    if (out[i] > out[big]) big = i;
  }
  out[big] += (int64_t)STM_EMD_SCALE - tot;
}

static void net_alloc(Net* g, int n, int m, const int32_t* sx, const int32_t* sy, const int32_t* tx, const int32_t* ty) {
  g->n = n; g->m = m; g->N = n + m + 1; g->root = n + m;
  g->sx = sx; g->sy = sy; g->tx = tx; g->ty = ty;
  const int N = g->N;
  g->parent = (int32_t*)malloc(sizeof(int32_t) * N);
  g->pos = (int32_t*)malloc(sizeof(int32_t) * N);
  g->size = (int32_t*)malloc(sizeof(int32_t) * N);
  g->order = (int32_t*)malloc(sizeof(int32_t) * N);
  g->tmp = (int32_t*)malloc(sizeof(int32_t) * N);
  g->flow = (int64_t*)malloc(sizeof(int64_t) * N);
  g->pi = (int64_t*)malloc(sizeof(int64_t) * N);
}
static void net_free(Net* g) {
  free(g->parent); free(g->pos); free(g->size); free(g->order); free(g->tmp); free(g->flow); free(g->pi);
}

static int64_t net_big(const Net* g) {
  const int n = g->n, m = g->m;
  int32_t x0 = g->sx[0], x1 = g->sx[0], y0 = g->sy[0], y1 = g->sy[0];
  for (int i = 0; i < n; ++i) { if (g->sx[i] < x0) x0 = g->sx[i]; if (g->sx[i] > x1) x1 = g->sx[i]; if (g->sy[i] < y0) y0 = g->sy[i]; if (g->sy[i] > y1) y1 = g->sy[i]; }
  for (int j = 0; j < m; ++j) { if (g->tx[j] < x0) x0 = g->tx[j]; if (g->tx[j] > x1) x1 = g->tx[j]; if (g->ty[j] < y0) y0 = g->ty[j]; if (g->ty[j] > y1) y1 = g->ty[j]; }
  const int64_t maxc = (int64_t)(x1 - x0) * (x1 - x0) + (int64_t)(y1 - y0) * (y1 - y0);
  return (maxc + 1) * (int64_t)g->N;
}

/* initial basis: star of artificial arcs; sources send to the root, sinks receive from it.
 * g->flow[0..n+m) must hold the integer supplies / demands. */
static void net_star(Net* g) {
  const int n = g->n, m = g->m, N = g->N;
  const int64_t BIG = net_big(g);
  g->order[0] = g->root; g->pos[g->root] = 0; g->size[g->root] = N; g->parent[g->root] = -1; g->pi[g->root] = 0; g->flow[g->root] = 0;
  for (int v = 0; v < n + m; ++v) {
    g->parent[v] = g->root; g->order[v + 1] = v; g->pos[v] = v + 1; g->size[v] = 1;
    g->pi[v] = v < n ? -BIG : BIG;   /* pi_head = pi_tail + cost along the flow direction */
  }
}

/* initial basis from a forest of real arcs with positive flows that meets every supply and demand: every component
 * is rooted at its lowest-index source, which hangs under the root by a zero-flow artificial arc (strongly feasible:
 * the only zero-flow arcs point to the root).  Returns 0, or -1 if the arcs do not form a spanning forest. */
static int net_from_forest(Net* g, const int32_t* ai, const int32_t* aj, const int64_t* af, int E) {
  const int n = g->n, N = g->N, root = g->root;
  const int64_t BIG = net_big(g);
  int32_t* deg = (int32_t*)calloc(N + 1, sizeof(int32_t));
  for (int e = 0; e < E; ++e) { deg[ai[e] + 1]++; deg[n + aj[e] + 1]++; }
  for (int v = 0; v < N; ++v) deg[v + 1] += deg[v];
  int32_t* adj = (int32_t*)malloc(sizeof(int32_t) * (2 * (size_t)E + 1));
  int32_t* cur = (int32_t*)malloc(sizeof(int32_t) * N);
  int32_t* stack = (int32_t*)malloc(sizeof(int32_t) * N);
  memcpy(cur, deg, sizeof(int32_t) * N);
  for (int e = 0; e < E; ++e) { adj[cur[ai[e]]++] = e; adj[cur[n + aj[e]]++] = e; }   /* lists in arc order */
  for (int v = 0; v < N; ++v) g->parent[v] = -2;
  g->parent[root] = -1; g->pi[root] = 0; g->flow[root] = 0;
  int cnt = 0, bad = 0;
  g->order[cnt] = root; g->pos[root] = cnt++;
  for (int s0 = 0; s0 < n && !bad; ++s0) {
    if (g->parent[s0] != -2) continue;
    g->parent[s0] = root; g->flow[s0] = 0; g->pi[s0] = -BIG;
    int sp = 0;
    stack[sp++] = s0;
    while (sp > 0 && !bad) {
      const int v = stack[--sp];
      g->order[cnt] = v; g->pos[v] = cnt++;
      for (int k = deg[v + 1] - 1; k >= deg[v]; --k) {      /* pushed in reverse: children are visited in list order */
        const int e = adj[k];
        const int w = v < n ? n + aj[e] : ai[e];
        if (w == g->parent[v]) continue;
        if (g->parent[w] != -2) { bad = 1; break; }
        g->parent[w] = v; g->flow[w] = af[e];
        const int64_t c = arc_cost(g, ai[e], aj[e]);
        g->pi[w] = w >= n ? g->pi[v] + c : g->pi[v] - c;
        stack[sp++] = w;
      }
    }
  }
  /* targets without any arc (zero integer demand): hung under the root like in the star */
  for (int v = n; v < root && !bad; ++v)
    if (g->parent[v] == -2) { g->parent[v] = root; g->flow[v] = 0; g->pi[v] = BIG; g->order[cnt] = v; g->pos[v] = cnt++; }
  if (cnt != N) bad = 1;
  if (!bad) {
    for (int x = 0; x < N; ++x) g->size[x] = 1;
    for (int x = N - 1; x >= 1; --x) { const int v = g->order[x]; g->size[g->parent[v]] += g->size[v]; }
  }
  free(deg); free(adj); free(cur); free(stack);
  return bad ? -1 : 0;
}

/*
 * The pivot loop on an initialised basis.  Returns 0 (optimal), 1 (iteration limit reached: the current basis is
 * kept, like POT), 3 (internal failure).  *iters is incremented by the number of pivots.
 *
 * cand_threads = 0: plain block search (one pivot per pricing pass).
 * cand_threads = T > 0: CANDIDATE LIST, the way the register-resident pricing path of the CUDA kernel with T threads
 * works (emd_spots.cu; it applies when the block is a whole number of target rows, B % n == 0, and the level has at most
 * R * T sources, R = 5 for T = 256 and ceil(5120 / T) otherwise — emd_rmax there).  The kernel's threads form
 * groups = T / teff row groups of teff = max(32, ceil(n / R)) rounded up to a power of two threads; the arc (i, j) with j the r-th row of the block
 * belongs to thread (class)  (r mod groups) * teff + (i mod teff).  A pricing pass (a "major" iteration) that finds a
 * negative block keeps every class's first most-negative arc of that block; the following "minor" iterations re-price
 * only these <= T candidates (those no longer negative are dropped), enter the most negative one (ties: lowest class)
 * and fall back to the block search where the last pass stopped once none is negative (LEMON's candidate-list idea).
 * Within a pass, the block that follows k blocks without a negative arc is 2^min(k, STM_EMD_GROW) times B long.
 * cand_threads < 0: the pricing of the global-tree kernel (problems beyond one CTA's shared memory) on blocks of at most
 * 64 whole target rows: its 512 threads own the sources i = thread (mod 512); ties inside a block go to the lowest source
 * index, then the first row; candidate list per thread as above; blocks grow up to 2^STM_EMD_GROW_G times B within 64 rows.
 * That kernel searches on the LOW 32 BITS of the reduced costs (its potentials are 64-bit, the reduced costs fit 32 bits in
 * practice): the selected arc is re-priced exactly — a non-negative value is a wrap artefact and switches the rest of the
 * level to 64-bit pricing — and when the 32-bit search finds nothing, one 64-bit sweep over all arcs certifies the optimum.
 */
static int net_run(Net* g, int64_t max_iter, int64_t block_size, int cand_threads, int64_t* iters) {
  const int n = g->n, m = g->m, N = g->N;
  int32_t* pathA = (int32_t*)malloc(sizeof(int32_t) * N);
  int32_t* pathB = (int32_t*)malloc(sizeof(int32_t) * N);
  int32_t* P = (int32_t*)malloc(sizeof(int32_t) * N);
  int32_t* S = (int32_t*)malloc(sizeof(int32_t) * N);
  int64_t* F = (int64_t*)malloc(sizeof(int64_t) * N);
  const int64_t n_arcs = (int64_t)n * m;
  int64_t B = block_size > 0 ? block_size : (int64_t)sqrt((double)n_arcs);
  if (B < 10) B = 10;
  if (B > n_arcs) B = n_arcs;
  int64_t next_arc = 0, it = *iters;
  int status = 0;
  const int grow_g = cand_threads < 0 && B % n == 0 && B / n <= 64;   /* whole-row pricing of the global-tree kernel */
  const int T = cand_threads > 0 ? cand_threads : grow_g ? 512 : 0;      /* candidate classes */
  const int R = T == 256 ? 5 : T > 0 ? (5120 + T - 1) / T : 0;
  const int use_cand = cand_threads > 0 && B % n == 0 && (int64_t)n <= (int64_t)R * T;
  int teff = 32;
  while (use_cand && teff < (n + R - 1) / R) teff <<= 1;
  const int groups = use_cand ? T / teff : 1;
  int64_t* cand_e = (int64_t*)malloc(sizeof(int64_t) * (T + 1));
  int64_t* cls_rc = (int64_t*)malloc(sizeof(int64_t) * (T + 1));
  int have_cand = 0;
  int wide_pr = 0;   /* global-tree kernel: 64-bit pricing for the rest of the level (wrap artefact seen, or the certificate) */
  for (;;) {
    /* ---- entering arc ----------------------------------------------------------------------- */
    int64_t scanned = 0, best_e = -1, best_rc = 0;
    if (have_cand) {   /* minor iteration: re-price the candidates kept by the last pass */
      for (int c = 0; c < T; ++c) {
        const int64_t e = cand_e[c];
        if (e < 0) continue;
        const int j = (int)(e / n), i = (int)(e - (int64_t)j * n);
        const int64_t rc = arc_cost(g, i, j) + g->pi[i] - g->pi[n + j];
        if (rc >= 0) { cand_e[c] = -1; continue; }
        if (rc < best_rc) { best_rc = rc; best_e = e; }   /* lowest class wins ties */
      }
      if (best_e < 0) have_cand = 0;
    }
    int empty = 0;                             /* consecutive blocks of this pass without a negative arc */
  search:
    while (best_e < 0 && scanned < n_arcs) {   /* major pass: block search over arcs e = j*n + i */
      /* with the candidate list, a block that follows k empty ones is 2^min(k, STM_EMD_GROW) times as long (the kernel
         amortises its per-block work over the long fruitless scans near the optimum) */
      int64_t Bk = use_cand ? B << (empty < STM_EMD_GROW ? empty : STM_EMD_GROW) : B;
      if (grow_g) {   /* global-tree kernel: up to STM_EMD_GROW_G doublings while the block stays within 64 target rows */
        int gsh = empty < STM_EMD_GROW_G ? empty : STM_EMD_GROW_G;
        while (gsh > 0 && ((B / n) << gsh) > 64) --gsh;
        Bk = B << gsh;
      }
      int64_t cnt = Bk < n_arcs - scanned ? Bk : n_arcs - scanned;
      ++empty;
      if (use_cand || grow_g) for (int c = 0; c < T; ++c) { cand_e[c] = -1; cls_rc[c] = 0; }
      best_e = -1; best_rc = 0;
      int j = (int)(next_arc / n), i = (int)(next_arc - (int64_t)j * n);
      int r = 0;                                           /* row of the block (use_cand: blocks start on a row) */
      for (int64_t k = 0; k < cnt; ++k) {
        int64_t rc = arc_cost(g, i, j) + g->pi[i] - g->pi[n + j];
        if (grow_g && !wide_pr) rc = (int64_t)(int32_t)(uint32_t)rc;   /* the global-tree kernel prices on the low 32 bits */
        /* first minimum in block order; the global-tree kernel's whole-row blocks: lowest source index, then first row */
        if (rc < best_rc || (grow_g && rc == best_rc && rc < 0 && i < (int)(best_e % n))) { best_rc = rc; best_e = (int64_t)j * n + i; }
        if (use_cand) {
          const int c = (r % groups) * teff + i % teff;
          if (rc < cls_rc[c]) { cls_rc[c] = rc; cand_e[c] = (int64_t)j * n + i; }   /* the class's first minimum */
        } else if (grow_g) {   /* global-tree kernel: thread = source index mod 512; ties: lowest source, then first row */
          const int c = i % T;
          if (rc < cls_rc[c] || (rc == cls_rc[c] && rc < 0 && i < (int)(cand_e[c] % n))) { cls_rc[c] = rc; cand_e[c] = (int64_t)j * n + i; }
        }
        if (++i == n) { i = 0; ++r; if (++j == m) j = 0; }
      }
      if (best_e >= 0 && grow_g && !wide_pr) {   /* the selected arc re-priced in 64 bits */
        const int bj = (int)(best_e / n), bi = (int)(best_e - (int64_t)bj * n);
        best_rc = arc_cost(g, bi, bj) + g->pi[bi] - g->pi[n + bj];
        if (best_rc >= 0) { wide_pr = 1; --empty; best_e = -1; best_rc = 0; continue; }   /* wrap artefact: the block again, in 64 bits */
      }
      scanned += cnt;
      next_arc += cnt;
      if (next_arc >= n_arcs) next_arc -= n_arcs;
      if (best_e >= 0) { have_cand = use_cand || grow_g; break; }
    }
    if (best_e < 0 && grow_g && !wide_pr) {      /* nothing negative in the low words: the certificate is a 64-bit sweep */
      wide_pr = 1; scanned = 0; empty = 0;
      goto search;
    }
    if (best_e < 0) break; /* optimal */
    if (it >= max_iter) { status = 1; break; }
    ++it;
    const int ej = (int)(best_e / n), ei = (int)(best_e - (int64_t)ej * n);
    const int u_i = ei, u_j = n + ej;
    /* ---- cycle: walk both endpoints up to the join node -------------------------------------- */
    int lenA = 0, lenB = 0;
    for (int v = u_i; !is_anc(g, v, u_j); v = g->parent[v]) pathA[lenA++] = v;
    for (int v = u_j; !is_anc(g, v, u_i); v = g->parent[v]) pathB[lenB++] = v;
    /* ---- leaving arc: min flow over the arcs traversed against their direction --------------- */
    /* cycle orientation = entering arc i -> j.  i-side (join ... -> i, downward): arc (v, parent) is
       backward iff v is a source.  j-side (j -> ... join, upward): backward iff v is a sink. */
    int64_t delta = INT64_MAX;
    int side = -1, idx = -1;
    for (int a = 0; a < lenA; ++a) {
      const int v = pathA[a];
      if (v < n && g->flow[v] < delta) { delta = g->flow[v]; side = 0; idx = a; }
    }
    for (int b = 0; b < lenB; ++b) {
      const int v = pathB[b];
      if (v >= n && g->flow[v] <= delta) { delta = g->flow[v]; side = 1; idx = b; }
    }
    if (side < 0) { status = 3; break; } /* unbounded: cannot happen for a transport LP */
    /* ---- flow update ----------------------------------------------------------------------- */
    for (int a = 0; a < lenA; ++a) { const int v = pathA[a]; g->flow[v] += (v < n) ? -delta : delta; }
    for (int b = 0; b < lenB; ++b) { const int v = pathB[b]; g->flow[v] += (v >= n) ? -delta : delta; }
    /* ---- tree update: subtree(q) is re-rooted at the entering endpoint inside it and hung under the
            other endpoint ---------------------------------------------------------------------- */
    int32_t* pathX = side == 0 ? pathA : pathB;
    int32_t* pathY = side == 0 ? pathB : pathA;
    const int lenX = side == 0 ? lenA : lenB, lenY = side == 0 ? lenB : lenA;
    const int k = idx;                      /* v_0 .. v_k = q on pathX */
    const int q = pathX[k];
    const int t_in = side == 0 ? u_j : u_i;
    const int s = g->size[q], a0 = g->pos[q];
    const int64_t shift = side == 0 ? -best_rc : best_rc;   /* makes the entering arc's reduced cost 0 */
    for (int x = a0; x < a0 + s; ++x) g->pi[g->order[x]] += shift;
    /* old positions / sizes of the path nodes (needed by the re-ordering below) */
    for (int t = 0; t <= k; ++t) { P[t] = g->pos[pathX[t]]; S[t] = g->size[pathX[t]]; F[t] = g->flow[pathX[t]]; }
    const int pt = g->pos[t_in];
    const int ptr = pt < a0 ? pt : pt - s;   /* position of t_in once the segment is removed */
    /* new preorder: segment removed, re-rooted segment inserted right after t_in (only [lo, hi) moves) */
    const int lo = pt < a0 ? pt + 1 : a0, hi = pt < a0 ? a0 + s : pt + 1;
    memcpy(g->tmp + lo, g->order + lo, sizeof(int32_t) * (size_t)(hi - lo));
    for (int x = lo; x < hi; ++x) {
      const int v = g->tmp[x];
      int f;
      if (x < a0 || x >= a0 + s) {
        const int xr = x < a0 ? x : x - s;
        f = xr <= ptr ? xr : xr + s;
      } else {
        int l = 0, h = k;                  /* smallest t with P[t] <= x < P[t]+S[t] (nested intervals) */
        while (l < h) { const int mid = (l + h) / 2; if (P[mid] <= x && x < P[mid] + S[mid]) h = mid; else l = mid + 1; }
        const int t = l;
        int r;
        if (t == 0) r = x - P[0];
        else if (x == P[t]) r = S[t - 1];
        else if (x < P[t - 1]) r = S[t - 1] + 1 + (x - (P[t] + 1));
        else r = x - P[t];
        f = ptr + 1 + r;
      }
      g->order[f] = v;
      g->pos[v] = f;
    }
    /* parents, flows and sizes along the reversed path */
    for (int t = k; t >= 1; --t) { g->parent[pathX[t]] = pathX[t - 1]; g->flow[pathX[t]] = F[t - 1]; g->size[pathX[t]] = s - S[t - 1]; }
    g->parent[pathX[0]] = t_in; g->flow[pathX[0]] = delta; g->size[pathX[0]] = s;
    for (int t = k + 1; t < lenX; ++t) g->size[pathX[t]] -= s;
    for (int t = 0; t < lenY; ++t) g->size[pathY[t]] += s;
  }
  *iters = it;
  free(pathA); free(pathB); free(P); free(S); free(F); free(cand_e); free(cls_rc);
  return status;
}

/* objective of the current basis: real tree arcs only (artificial arcs carry no flow at the optimum) */
static double net_cost(const Net* g) {
  double total = 0.0;
  for (int v = 0; v < g->n + g->m; ++v) {
    const int p = g->parent[v];
    if (p == g->root || g->flow[v] == 0) continue;
    const int i = v < g->n ? v : p, j = (v < g->n ? p : v) - g->n;
    total += (double)g->flow[v] * (double)arc_cost(g, i, j);
  }
  return total / STM_EMD_SCALE;
}

/*
 * Exact EMD from the artificial star.  Returns 0 (optimal), 1 (iteration limit reached: cost of the current basis,
 * like POT), 2 (empty input).  *iters receives the number of pivots.  block_size <= 0 selects max(10, sqrt(n*m))
 * like LEMON.
 */
int stm_oracle_emd(const int32_t* sx, const int32_t* sy, const double* smass, int n,
                   const int32_t* tx, const int32_t* ty, const double* tmass, int m,
                   int64_t max_iter, int64_t block_size, int cand_threads, double* cost_out, int64_t* iters) {
  *cost_out = 0.0;
  if (iters) *iters = 0;
  if (n <= 0 || m <= 0) return 2;
  Net g;
  net_alloc(&g, n, m, sx, sy, tx, ty);
  integerize(smass, n, g.flow);
  integerize(tmass, m, g.flow + n);
  net_star(&g);
  int64_t it = 0;
  const int status = net_run(&g, max_iter, block_size, cand_threads, &it);
  *cost_out = net_cost(&g);
  if (iters) *iters = it;
  net_free(&g);
  return status;
}

/* ------------------------------------------------------------------------------------------------------------------
 * Multiscale start (what the CUDA kernel does by default).
 * ---------------------------------------------------------------------------------------------------------------- */
static inline uint32_t morton16(uint32_t x, uint32_t y) {   /* x, y < 2^16: x in the odd bits, y in the even bits */
  uint32_t a = x, b = y;
  a = (a | (a << 8)) & 0x00ff00ffu; a = (a | (a << 4)) & 0x0f0f0f0fu; a = (a | (a << 2)) & 0x33333333u; a = (a | (a << 1)) & 0x55555555u;
  b = (b | (b << 8)) & 0x00ff00ffu; b = (b | (b << 4)) & 0x0f0f0f0fu; b = (b | (b << 2)) & 0x33333333u; b = (b | (b << 1)) & 0x55555555u;
  return (a << 1) | b;
}

typedef struct { uint64_t key; } SortKey;
static int cmp_key(const void* a, const void* b) {
  const uint64_t x = ((const SortKey*)a)->key, y = ((const SortKey*)b)->key;
  return x < y ? -1 : x > y;
}

/* One side's hierarchy over BIT levels: level 0 = the points in Morton order (ties by input index); level b + 1 = runs of
 * level-b nodes whose Morton codes agree after dropping one more bit — cells are halved along y first (odd b: 2^(b/2) wide,
 * 2^(b/2 + 1) high), then along x.  x/y of level b are the cell coordinates (u_x >> b/2, u_y >> (b+1)/2) of the biased
 * coordinates u; child[b][c] .. child[b][c+1] are the children (at level b-1) of node c of level b >= 1. */
typedef struct {
  int L;
  int cnt[STM_EMD_MAXLEV + 1];
  uint32_t* code[STM_EMD_MAXLEV + 1];   /* morton code of the level-0 cell >> b */
  int32_t* x[STM_EMD_MAXLEV + 1];
  int32_t* y[STM_EMD_MAXLEV + 1];
  int64_t* mass[STM_EMD_MAXLEV + 1];
  int32_t* child[STM_EMD_MAXLEV + 1];
} Hier;

static void hier_build(Hier* h, const int32_t* px, const int32_t* py, const int64_t* mass, int n, int32_t x0, int32_t y0, int levels) {
  SortKey* k = (SortKey*)malloc(sizeof(SortKey) * n);
  for (int i = 0; i < n; ++i)
    k[i].key = ((uint64_t)morton16((uint32_t)(px[i] - x0 + 16384), (uint32_t)(py[i] - y0 + 16384)) << 32) | (uint32_t)i;
  qsort(k, n, sizeof(SortKey), cmp_key);
  h->cnt[0] = n;
  h->code[0] = (uint32_t*)malloc(sizeof(uint32_t) * n); h->x[0] = (int32_t*)malloc(sizeof(int32_t) * n);
  h->y[0] = (int32_t*)malloc(sizeof(int32_t) * n); h->mass[0] = (int64_t*)malloc(sizeof(int64_t) * n); h->child[0] = 0;
  for (int i = 0; i < n; ++i) {
    const int o = (int)(k[i].key & 0xffffffffu);
    h->code[0][i] = (uint32_t)(k[i].key >> 32); h->x[0][i] = px[o] - x0; h->y[0][i] = py[o] - y0; h->mass[0][i] = mass[o];
  }
  free(k);
  for (int l = 1; l < levels; ++l) {
    const int pn = h->cnt[l - 1];
    h->code[l] = (uint32_t*)malloc(sizeof(uint32_t) * pn); h->x[l] = (int32_t*)malloc(sizeof(int32_t) * pn);
    h->y[l] = (int32_t*)malloc(sizeof(int32_t) * pn); h->mass[l] = (int64_t*)malloc(sizeof(int64_t) * pn);
    h->child[l] = (int32_t*)malloc(sizeof(int32_t) * (pn + 1));
    int c = -1;
    for (int i = 0; i < pn; ++i) {
      const uint32_t pc = h->code[l - 1][i] >> 1;
      if (c < 0 || pc != h->code[l][c]) {
        ++c;
        h->code[l][c] = pc; h->child[l][c] = i; h->mass[l][c] = 0;
        /* odd level: the y bit goes (the Morton code holds y in the even bits); arithmetic shift = floor: cell coordinates */
        h->x[l][c] = (l & 1) ? h->x[l - 1][i] : h->x[l - 1][i] >> 1;
        h->y[l][c] = (l & 1) ? h->y[l - 1][i] >> 1 : h->y[l - 1][i];
      }
      h->mass[l][c] += h->mass[l - 1][i];
    }
    h->cnt[l] = c + 1;
    h->child[l][c + 1] = pn;
  }
  h->L = levels;
}
static void hier_free(Hier* h) {
  for (int l = 0; l < h->L; ++l) { free(h->code[l]); free(h->x[l]); free(h->y[l]); free(h->mass[l]); if (h->child[l]) free(h->child[l]); }
}

static int cmp_arc_ij(const void* a, const void* b) {   /* arcs packed as (I << 32 | J) */
  const uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
  return x < y ? -1 : x > y;
}

/* Coarse optimum (real tree arcs with positive flow) -> feasible forest of the next finer level.
 * Stage 1, per coarse source I: north-west corner between its children (in order) and its arcs (by increasing J)
 * gives pieces (fine source, J, flow).  Stage 2, per coarse target J: north-west corner between its pieces (by
 * increasing I, then child order = increasing fine source) and its children gives the fine arcs. */
static int refine_level(const Net* gc, const Hier* hs, const Hier* ht, int lev /* coarse level */, int flev /* fine: lev - 1 or lev - 2 */,
                        int32_t** out_i, int32_t** out_j, int64_t** out_f) {
  const int nS = gc->n, nT = gc->m;
  /* coarse arcs sorted by (I, J) */
  int Ec = 0;
  uint64_t* key = (uint64_t*)malloc(sizeof(uint64_t) * (nS + nT));
  for (int v = 0; v < nS + nT; ++v) {
    const int p = gc->parent[v];
    if (p == gc->root || gc->flow[v] == 0) continue;
    const int I = v < nS ? v : p, J = (v < nS ? p : v) - nS;
    key[Ec++] = ((uint64_t)I << 32) | (uint32_t)J;
  }
  qsort(key, Ec, sizeof(uint64_t), cmp_arc_ij);
  int32_t* aI = (int32_t*)malloc(sizeof(int32_t) * (Ec + 1)); int32_t* aJ = (int32_t*)malloc(sizeof(int32_t) * (Ec + 1));
  int64_t* aF = (int64_t*)malloc(sizeof(int64_t) * (Ec + 1));
  for (int e = 0; e < Ec; ++e) {
    aI[e] = (int32_t)(key[e] >> 32); aJ[e] = (int32_t)(key[e] & 0xffffffffu);
    const int I = aI[e], J = nS + aJ[e];
    aF[e] = gc->parent[I] == J ? gc->flow[I] : gc->flow[J];   /* the arc is stored at its child end */
  }
  free(key);
  /* children at the fine level: directly, or through the skipped level in between (descendants stay contiguous) */
  int32_t* cs = (int32_t*)malloc(sizeof(int32_t) * (nS + 1)); int32_t* ct = (int32_t*)malloc(sizeof(int32_t) * (nT + 1));
  for (int I = 0; I <= nS; ++I) cs[I] = flev == lev - 1 ? hs->child[lev][I] : hs->child[lev - 1][hs->child[lev][I]];
  for (int J = 0; J <= nT; ++J) ct[J] = flev == lev - 1 ? ht->child[lev][J] : ht->child[lev - 1][ht->child[lev][J]];
  const int64_t* fs = hs->mass[flev]; const int64_t* ft = ht->mass[flev];
  const int nf = hs->cnt[flev], mf = ht->cnt[flev];
  /* stage 1: pieces per arc (in arc order (I, J)); at most children(I) pieces per arc */
  const int maxp = nf + Ec + 1;
  int32_t* pc_i = (int32_t*)malloc(sizeof(int32_t) * maxp); int64_t* pc_f = (int64_t*)malloc(sizeof(int64_t) * maxp);
  int32_t* pc_first = (int32_t*)malloc(sizeof(int32_t) * (Ec + 1));   /* first piece of arc e */
  int np = 0, e = 0;
  for (int I = 0; I < nS; ++I) {
    int a = cs[I];
    const int a_end = cs[I + 1];
    int64_t ra = a < a_end ? fs[a] : 0;
    for (; e < Ec && aI[e] == I; ++e) {
      pc_first[e] = np;
      int64_t rb = aF[e];
      while (rb > 0 && a < a_end) {
        const int64_t f = ra < rb ? ra : rb;
        if (f > 0) { pc_i[np] = a; pc_f[np] = f; ++np; }
        ra -= f; rb -= f;
        if (ra == 0) { ++a; ra = a < a_end ? fs[a] : 0; }
      }
    }
  }
  pc_first[Ec] = np;
  /* arcs grouped by J, increasing I inside a group (counting sort by J keeps the (I, J) order) */
  int32_t* toff = (int32_t*)calloc(nT + 1, sizeof(int32_t));
  for (int k = 0; k < Ec; ++k) toff[aJ[k] + 1]++;
  for (int J = 0; J < nT; ++J) toff[J + 1] += toff[J];
  int32_t* tlist = (int32_t*)malloc(sizeof(int32_t) * (Ec + 1)); int32_t* tcur = (int32_t*)malloc(sizeof(int32_t) * (nT + 1));
  memcpy(tcur, toff, sizeof(int32_t) * (nT + 1));
  for (int k = 0; k < Ec; ++k) tlist[tcur[aJ[k]]++] = k;
  /* stage 2 */
  const int maxa = np + mf + 1;
  int32_t* oi = (int32_t*)malloc(sizeof(int32_t) * maxa); int32_t* oj = (int32_t*)malloc(sizeof(int32_t) * maxa);
  int64_t* of = (int64_t*)malloc(sizeof(int64_t) * maxa);
  int E = 0;
  for (int J = 0; J < nT; ++J) {
    int b = ct[J];
    const int b_end = ct[J + 1];
    int64_t rb = b < b_end ? ft[b] : 0;
    for (int q = toff[J]; q < toff[J + 1]; ++q) {
      const int k = tlist[q];
      for (int pz = pc_first[k]; pz < pc_first[k + 1]; ++pz) {
        int64_t ra = pc_f[pz];
        while (ra > 0 && b < b_end) {
          const int64_t f = ra < rb ? ra : rb;
          if (f > 0) { oi[E] = pc_i[pz]; oj[E] = b; of[E] = f; ++E; }
          ra -= f; rb -= f;
          if (rb == 0) { ++b; rb = b < b_end ? ft[b] : 0; }
        }
      }
    }
  }
  free(aI); free(aJ); free(aF); free(pc_i); free(pc_f); free(pc_first); free(toff); free(tlist); free(tcur); free(cs); free(ct);
  (void)nf;
  *out_i = oi; *out_j = oj; *out_f = of;
  return E;
}

/*
 * Exact EMD with the multiscale start.  flags bit 0: star start (then identical to stm_oracle_emd).
 * block_size > 0: that block at every level (clamped to the level's arc count); <= 0: max(sqrt(n_l m_l),
 * min(n_l m_l, rows * n_l)) with rows = 7 — whole target rows, as the CUDA kernel prices.
 * level_iters (optional, STM_EMD_MAXLEV + 1 entries): pivots per level in use, [0] = finest.
 */
int stm_oracle_emd_ms(const int32_t* sx, const int32_t* sy, const double* smass, int n,
                      const int32_t* tx, const int32_t* ty, const double* tmass, int m,
                      int64_t max_iter, int64_t block_size, uint32_t flags, int cand_threads, double* cost_out,
                      int64_t* iters, int64_t* level_iters) {
  if (flags & 1u) return stm_oracle_emd(sx, sy, smass, n, tx, ty, tmass, m, max_iter, block_size, cand_threads, cost_out, iters);
  *cost_out = 0.0;
  if (iters) *iters = 0;
  if (level_iters) memset(level_iters, 0, sizeof(int64_t) * (STM_EMD_MAXLEV + 1));
  if (n <= 0 || m <= 0) return 2;
  int64_t* si = (int64_t*)malloc(sizeof(int64_t) * n); int64_t* ti = (int64_t*)malloc(sizeof(int64_t) * m);
  integerize(smass, n, si);
  integerize(tmass, m, ti);
  int32_t x0 = sx[0], y0 = sy[0];
  for (int i = 0; i < n; ++i) { if (sx[i] < x0) x0 = sx[i]; if (sy[i] < y0) y0 = sy[i]; }
  {   /* the Morton keys take coordinates relative to the source's corner in [-16384, 49152); otherwise: star start */
    int ok = 1;
    for (int i = 0; i < n && ok; ++i) ok = sx[i] - x0 < 49152 && sy[i] - y0 < 49152;
    for (int j = 0; j < m && ok; ++j) ok = tx[j] - x0 >= -16384 && tx[j] - x0 < 49152 && ty[j] - y0 >= -16384 && ty[j] - y0 < 49152;
    if (!ok) { free(si); free(ti); return stm_oracle_emd(sx, sy, smass, n, tx, ty, tmass, m, max_iter, block_size, cand_threads, cost_out, iters); }
  }
  /* number of levels: coarsen until the coarsest problem has at most STM_EMD_COARSEST_NODES nodes */
  Hier hs, ht;
  hier_build(&hs, sx, sy, si, n, x0, y0, STM_EMD_MAXLEV + 1);
  hier_build(&ht, tx, ty, ti, m, x0, y0, STM_EMD_MAXLEV + 1);
  /* the bit levels in use, finest first: a level that keeps more than 4/5 of the nodes of the one below it is skipped (on
     the Visium lattice, where (row + col) is even, halving along y merges nothing); never two in a row */
  int seq[STM_EMD_MAXLEV + 1];
  int L = 1;
  seq[0] = 0;
  for (int b = 0; hs.cnt[b] + ht.cnt[b] > STM_EMD_COARSEST_NODES && b + 1 <= STM_EMD_MAXLEV && L < STM_EMD_MAXUSE;) {
    int nb = b + 1;
    if (nb + 1 <= STM_EMD_MAXLEV &&
        (int64_t)STM_EMD_SKIP_DEN * (hs.cnt[nb] + ht.cnt[nb]) > (int64_t)STM_EMD_SKIP_NUM * (hs.cnt[b] + ht.cnt[b])) nb = b + 2;
    seq[L++] = nb;
    b = nb;
  }
  int status = 0;
  int64_t it_total = 0;
  int32_t *ai = 0, *aj = 0; int64_t* af = 0;
  int E = 0;
  double cost = 0.0;
  for (int li = L - 1; li >= 0 && status == 0; --li) {
    const int lev = seq[li];
    Net g;
    const int nl = hs.cnt[lev], ml = ht.cnt[lev];
    /* isotropic cost in units of the cell width: an odd level's cells are twice as high as wide */
    int32_t* ys = (int32_t*)malloc(sizeof(int32_t) * (nl + 1)); int32_t* yt = (int32_t*)malloc(sizeof(int32_t) * (ml + 1));
    for (int i = 0; i < nl; ++i) ys[i] = hs.y[lev][i] * (1 + (lev & 1));
    for (int j = 0; j < ml; ++j) yt[j] = ht.y[lev][j] * (1 + (lev & 1));
    net_alloc(&g, nl, ml, hs.x[lev], ys, ht.x[lev], yt);
    int ok = -1;
    if (ai) {
      ok = net_from_forest(&g, ai, aj, af, E);
      free(ai); free(aj); free(af); ai = aj = 0; af = 0;
    }
    if (ok != 0) {   /* coarsest level (or a refinement that does not span: cannot happen) -> star */
      memcpy(g.flow, hs.mass[lev], sizeof(int64_t) * nl);
      memcpy(g.flow + nl, ht.mass[lev], sizeof(int64_t) * ml);
      net_star(&g);
    }
    const int64_t na = (int64_t)nl * ml;
    int64_t B = block_size;
    if (B <= 0) { B = (int64_t)sqrt((double)na); const int64_t rows = 7 * (int64_t)nl < na ? 7 * (int64_t)nl : na; if (rows > B) B = rows; }
    /* max_iter (numItermax of the reference) bounds the pivots on the ORIGINAL problem, i.e. the finest level; the
       coarse levels only construct its start basis and always run to their optimum */
    int64_t it = 0;
    status = net_run(&g, li == 0 ? max_iter : INT64_MAX, B, cand_threads, &it);
    if (level_iters) level_iters[li] = it;
    it_total += it;
    if (li == 0 || status != 0) cost = li == 0 ? net_cost(&g) : 0.0;
    else E = refine_level(&g, &hs, &ht, lev, seq[li - 1], &ai, &aj, &af);
    net_free(&g);
    free(ys); free(yt);
  }
  if (ai) { free(ai); free(aj); free(af); }
  hier_free(&hs); hier_free(&ht);
  free(si); free(ti);
  *cost_out = cost;
  if (iters) *iters = it_total;
  return status;
}

/* Batched driver with the layout of the C ABI (one source image, G target images). */
int stm_oracle_emd_batched(const int32_t* sx, const int32_t* sy, const double* smass, int n,
                           const int64_t* tgt_ptr, const int32_t* tx, const int32_t* ty, const double* tmass, int G,
                           int64_t max_iter, int64_t block_size, double* cost, int32_t* status, int64_t* iters) {
  for (int g = 0; g < G; ++g) {
    const int64_t a = tgt_ptr[g], b = tgt_ptr[g + 1];
    status[g] = stm_oracle_emd(sx, sy, smass, n, tx + a, ty + a, tmass + a, (int)(b - a), max_iter, block_size, 0,
                               &cost[g], iters ? &iters[g] : 0);
  }
  return 0;
}
