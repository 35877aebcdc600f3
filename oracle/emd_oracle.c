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
 * Pinning — PARITY UNPINNED against the reference itself: POT cannot be imported here and the reference holds no golden
 * value for this path that is usable here (its only pin is the top-2
 * gene ranking on the missing tests/data/test.h5ad, tests/test_STMiner.py:48-49).  The oracle is pinned
 * against an independent exact LP solve (scipy.optimize.linprog, HiGHS) on small problems in
 * tests/test_oracle_emd.py, and against committed golden values tests/golden/emd_small.npz made the same way.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define STM_EMD_SCALE 1125899906842624.0 /* 2^50 */

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

/*
 * Exact EMD.  Returns 0 (optimal), 1 (iteration limit reached: cost of the current basis, like POT), 2 (empty
 * input).  *iters receives the number of pivots.  block_size <= 0 selects max(10, sqrt(n*m)) like LEMON.
 *
 * cand_threads = 0: plain block search (one pivot per pricing pass).
 * cand_threads = T > 0: candidate list, the way the CUDA kernel with T threads prices — arc offset k of a block
 * belongs to class (k mod T) / 32 (the warp that evaluates it); the pass keeps the best arc of every class, and
 * "minor" pivots then re-price only those T/32 candidates until none is negative (LEMON's candidate-list idea).
 */
int stm_oracle_emd(const int32_t* sx, const int32_t* sy, const double* smass, int n,
                   const int32_t* tx, const int32_t* ty, const double* tmass, int m,
                   int64_t max_iter, int64_t block_size, int cand_threads, double* cost_out, int64_t* iters) {
  *cost_out = 0.0;
  if (iters) *iters = 0;
  if (n <= 0 || m <= 0) return 2;
  Net g;
  g.n = n; g.m = m; g.N = n + m + 1; g.root = n + m;
  g.sx = sx; g.sy = sy; g.tx = tx; g.ty = ty;
  const int N = g.N;
  g.parent = (int32_t*)malloc(sizeof(int32_t) * N);
  g.pos = (int32_t*)malloc(sizeof(int32_t) * N);
  g.size = (int32_t*)malloc(sizeof(int32_t) * N);
  g.order = (int32_t*)malloc(sizeof(int32_t) * N);
  g.tmp = (int32_t*)malloc(sizeof(int32_t) * N);
  g.flow = (int64_t*)malloc(sizeof(int64_t) * N);
  g.pi = (int64_t*)malloc(sizeof(int64_t) * N);
  int32_t* pathA = (int32_t*)malloc(sizeof(int32_t) * N);
  int32_t* pathB = (int32_t*)malloc(sizeof(int32_t) * N);

  integerize(smass, n, g.flow);
  integerize(tmass, m, g.flow + n);
  int64_t maxc = 0;
  {
    int32_t x0 = sx[0], x1 = sx[0], y0 = sy[0], y1 = sy[0];
    for (int i = 0; i < n; ++i) { if (sx[i] < x0) x0 = sx[i]; if (sx[i] > x1) x1 = sx[i]; if (sy[i] < y0) y0 = sy[i]; if (sy[i] > y1) y1 = sy[i]; }
    for (int j = 0; j < m; ++j) { if (tx[j] < x0) x0 = tx[j]; if (tx[j] > x1) x1 = tx[j]; if (ty[j] < y0) y0 = ty[j]; if (ty[j] > y1) y1 = ty[j]; }
    maxc = (int64_t)(x1 - x0) * (x1 - x0) + (int64_t)(y1 - y0) * (y1 - y0);
  }
  const int64_t BIG = (maxc + 1) * (int64_t)N;
  /* initial basis: star of artificial arcs; sources send to the root, sinks receive from it */
  g.order[0] = g.root; g.pos[g.root] = 0; g.size[g.root] = N; g.parent[g.root] = -1; g.pi[g.root] = 0; g.flow[g.root] = 0;
  for (int v = 0; v < n + m; ++v) {
    g.parent[v] = g.root; g.order[v + 1] = v; g.pos[v] = v + 1; g.size[v] = 1;
    g.pi[v] = v < n ? -BIG : BIG;   /* pi_head = pi_tail + cost along the flow direction */
  }
  const int64_t n_arcs = (int64_t)n * m;
  int64_t B = block_size > 0 ? block_size : (int64_t)sqrt((double)n_arcs);
  if (B < 10) B = 10;
  if (B > n_arcs) B = n_arcs;
  int64_t next_arc = 0, it = 0;
  int status = 0;
  const int n_cls = cand_threads > 0 ? cand_threads / 32 : 0;
  int64_t cand_e[64];
  int n_cand = 0;
  for (;;) {
    /* ---- entering arc ----------------------------------------------------------------------- */
    int64_t scanned = 0, best_e = -1, best_rc = 0;
    if (n_cand > 0) {   /* minor iteration: re-price the candidates kept by the last pass */
      for (int c = 0; c < n_cls; ++c) {
        const int64_t e = cand_e[c];
        if (e < 0) continue;
        const int j = (int)(e / n), i = (int)(e - (int64_t)j * n);
        const int64_t rc = arc_cost(&g, i, j) + g.pi[i] - g.pi[n + j];
        if (rc < best_rc) { best_rc = rc; best_e = e; }   /* lowest class wins ties */
      }
      if (best_e < 0) n_cand = 0;
    }
    while (best_e < 0 && scanned < n_arcs) {   /* major pass: block search over arcs e = j*n + i */
      int64_t cnt = B < n_arcs - scanned ? B : n_arcs - scanned;
      int64_t cls_rc[64];
      for (int c = 0; c < n_cls; ++c) { cand_e[c] = -1; cls_rc[c] = 0; }
      best_e = -1; best_rc = 0;
      int64_t best_k = -1;
      for (int64_t k = 0; k < cnt; ++k) {
        int64_t e = next_arc + k;
        if (e >= n_arcs) e -= n_arcs;
        const int j = (int)(e / n), i = (int)(e - (int64_t)j * n);
        const int64_t rc = arc_cost(&g, i, j) + g.pi[i] - g.pi[n + j];
        if (rc < best_rc) { best_rc = rc; best_e = e; best_k = k; }   /* first minimum in block order */
        if (n_cls > 0) {
          const int c = (int)((k % cand_threads) / 32);
          if (rc < cls_rc[c]) { cls_rc[c] = rc; cand_e[c] = e; }
        }
      }
      (void)best_k;
      scanned += cnt;
      next_arc += cnt;
      if (next_arc >= n_arcs) next_arc -= n_arcs;
      if (best_e >= 0) { n_cand = n_cls; break; }
    }
    if (best_e < 0) break; /* optimal */
    if (it >= max_iter) { status = 1; break; }
    ++it;
    const int ej = (int)(best_e / n), ei = (int)(best_e - (int64_t)ej * n);
    const int u_i = ei, u_j = n + ej;
    /* ---- cycle: walk both endpoints up to the join node -------------------------------------- */
    int lenA = 0, lenB = 0;
    for (int v = u_i; !is_anc(&g, v, u_j); v = g.parent[v]) pathA[lenA++] = v;
    for (int v = u_j; !is_anc(&g, v, u_i); v = g.parent[v]) pathB[lenB++] = v;
    /* ---- leaving arc: min flow over the arcs traversed against their direction --------------- */
    /* cycle orientation = entering arc i -> j.  i-side (join ... -> i, downward): arc (v, parent) is
       backward iff v is a source.  j-side (j -> ... join, upward): backward iff v is a sink. */
    int64_t delta = INT64_MAX;
    int side = -1, idx = -1;
    for (int a = 0; a < lenA; ++a) {
      const int v = pathA[a];
      if (v < n && g.flow[v] < delta) { delta = g.flow[v]; side = 0; idx = a; }
    }
    for (int b = 0; b < lenB; ++b) {
      const int v = pathB[b];
      if (v >= n && g.flow[v] <= delta) { delta = g.flow[v]; side = 1; idx = b; }
    }
    if (side < 0) { status = 3; break; } /* unbounded: cannot happen for a transport LP */
    /* ---- flow update ----------------------------------------------------------------------- */
    for (int a = 0; a < lenA; ++a) { const int v = pathA[a]; g.flow[v] += (v < n) ? -delta : delta; }
    for (int b = 0; b < lenB; ++b) { const int v = pathB[b]; g.flow[v] += (v >= n) ? -delta : delta; }
    /* ---- tree update: subtree(q) is re-rooted at the entering endpoint inside it and hung under the
            other endpoint ---------------------------------------------------------------------- */
    int32_t* pathX = side == 0 ? pathA : pathB;
    int32_t* pathY = side == 0 ? pathB : pathA;
    const int lenX = side == 0 ? lenA : lenB, lenY = side == 0 ? lenB : lenA;
    const int k = idx;                      /* v_0 .. v_k = q on pathX */
    const int q = pathX[k];
    const int t_in = side == 0 ? u_j : u_i;
    const int s = g.size[q], a0 = g.pos[q];
    const int64_t shift = side == 0 ? -best_rc : best_rc;   /* makes the entering arc's reduced cost 0 */
    for (int x = a0; x < a0 + s; ++x) g.pi[g.order[x]] += shift;
    /* old positions / sizes of the path nodes (needed by the re-ordering below) */
    int32_t* P = (int32_t*)malloc(sizeof(int32_t) * (k + 1));
    int32_t* S = (int32_t*)malloc(sizeof(int32_t) * (k + 1));
    int64_t* F = (int64_t*)malloc(sizeof(int64_t) * (k + 1));
    for (int t = 0; t <= k; ++t) { P[t] = g.pos[pathX[t]]; S[t] = g.size[pathX[t]]; F[t] = g.flow[pathX[t]]; }
    const int pt = g.pos[t_in];
    const int ptr = pt < a0 ? pt : pt - s;   /* position of t_in once the segment is removed */
    /* new preorder: segment removed, re-rooted segment inserted right after t_in */
    memcpy(g.tmp, g.order, sizeof(int32_t) * N);
    for (int x = 0; x < N; ++x) {
      const int v = g.tmp[x];
      int f;
      if (x < a0 || x >= a0 + s) {
        const int xr = x < a0 ? x : x - s;
        f = xr <= ptr ? xr : xr + s;
      } else {
        int lo = 0, hi = k;                  /* smallest t with P[t] <= x < P[t]+S[t] (nested intervals) */
        while (lo < hi) { const int mid = (lo + hi) / 2; if (P[mid] <= x && x < P[mid] + S[mid]) hi = mid; else lo = mid + 1; }
        const int t = lo;
        int r;
        if (t == 0) r = x - P[0];
        else if (x == P[t]) r = S[t - 1];
        else if (x < P[t - 1]) r = S[t - 1] + 1 + (x - (P[t] + 1));
        else r = x - P[t];
        f = ptr + 1 + r;
      }
      g.order[f] = v;
      g.pos[v] = f;
    }
    /* parents, flows and sizes along the reversed path */
    for (int t = k; t >= 1; --t) { g.parent[pathX[t]] = pathX[t - 1]; g.flow[pathX[t]] = F[t - 1]; g.size[pathX[t]] = s - S[t - 1]; }
    g.parent[pathX[0]] = t_in; g.flow[pathX[0]] = delta; g.size[pathX[0]] = s;
    for (int t = k + 1; t < lenX; ++t) g.size[pathX[t]] -= s;
    for (int t = 0; t < lenY; ++t) g.size[pathY[t]] += s;
    free(P); free(S); free(F);
  }
  /* ---- objective: real tree arcs only (artificial arcs carry no flow at the optimum) ------------ */
  double total = 0.0;
  for (int v = 0; v < n + m; ++v) {
    const int p = g.parent[v];
    if (p == g.root || g.flow[v] == 0) continue;
    const int i = v < n ? v : p, j = (v < n ? p : v) - n;
    total += (double)g.flow[v] * (double)arc_cost(&g, i, j);
  }
  *cost_out = total / STM_EMD_SCALE;
  if (iters) *iters = it;
  free(g.parent); free(g.pos); free(g.size); free(g.order); free(g.tmp); free(g.flow); free(g.pi);
  free(pathA); free(pathB);
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
