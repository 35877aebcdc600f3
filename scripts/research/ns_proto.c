/* Research prototype: network simplex on the bipartite transport graph from an arbitrary initial FOREST
 * (list of arcs with positive integer flows), to count pivots for different start bases.  Not product code. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int n, m, N, root;
  const int32_t *sx, *sy, *tx, *ty;
  int32_t *parent, *pos, *size, *order, *tmp;
  int64_t *flow, *pi;
} Net;

static inline int64_t arc_cost(const Net* g, int i, int j) {
  const int64_t dx = (int64_t)g->sx[i] - g->tx[j], dy = (int64_t)g->sy[i] - g->ty[j];
  return dx * dx + dy * dy;
}
static inline int is_anc(const Net* g, int u, int w) { return g->pos[u] <= g->pos[w] && g->pos[w] < g->pos[u] + g->size[u]; }

/* supply[n], demand[m] integer with equal totals.  init arcs: (ai[e], aj[e], af[e]) a forest with feasible flows
 * (every node's balance met), E may be 0 -> star start.  Returns pivots; cost via *cost_out (sum f*c as double);
 * out_parent/out_flow (N entries) receive the final basis if non-NULL. */
int64_t ns_solve(const int32_t* sx, const int32_t* sy, const int64_t* supply, int n,
                 const int32_t* tx, const int32_t* ty, const int64_t* demand, int m,
                 const int32_t* ai, const int32_t* aj, const int64_t* af, int E,
                 int64_t max_iter, int64_t block_size, double* cost_out, int32_t* out_parent, int64_t* out_flow,
                 int64_t* stats /* [4]: degenerate pivots, sum cycle len, sum range, max cycle */) {
  Net g;
  g.n = n; g.m = m; g.N = n + m + 1; g.root = n + m;
  g.sx = sx; g.sy = sy; g.tx = tx; g.ty = ty;
  const int N = g.N, root = g.root;
  g.parent = malloc(sizeof(int32_t) * N); g.pos = malloc(sizeof(int32_t) * N); g.size = malloc(sizeof(int32_t) * N);
  g.order = malloc(sizeof(int32_t) * N); g.tmp = malloc(sizeof(int32_t) * N);
  g.flow = malloc(sizeof(int64_t) * N); g.pi = malloc(sizeof(int64_t) * N);
  int32_t* pathA = malloc(sizeof(int32_t) * N); int32_t* pathB = malloc(sizeof(int32_t) * N);
  int32_t* P = malloc(sizeof(int32_t) * N); int32_t* S = malloc(sizeof(int32_t) * N); int64_t* F = malloc(sizeof(int64_t) * N);
  int64_t maxc = 0;
  {
    int32_t x0 = sx[0], x1 = sx[0], y0 = sy[0], y1 = sy[0];
    for (int i = 0; i < n; ++i) { if (sx[i] < x0) x0 = sx[i]; if (sx[i] > x1) x1 = sx[i]; if (sy[i] < y0) y0 = sy[i]; if (sy[i] > y1) y1 = sy[i]; }
    for (int j = 0; j < m; ++j) { if (tx[j] < x0) x0 = tx[j]; if (tx[j] > x1) x1 = tx[j]; if (ty[j] < y0) y0 = ty[j]; if (ty[j] > y1) y1 = ty[j]; }
    maxc = (int64_t)(x1 - x0) * (x1 - x0) + (int64_t)(y1 - y0) * (y1 - y0);
  }
  const int64_t BIG = (maxc + 1) * (int64_t)N;
  if (E == 0) {
    g.order[0] = root; g.pos[root] = 0; g.size[root] = N; g.parent[root] = -1; g.pi[root] = 0; g.flow[root] = 0;
    for (int v = 0; v < n + m; ++v) {
      g.parent[v] = root; g.order[v + 1] = v; g.pos[v] = v + 1; g.size[v] = 1;
      g.pi[v] = v < n ? -BIG : BIG; g.flow[v] = v < n ? supply[v] : demand[v - n];
    }
  } else {
    /* adjacency (CSR) of the forest */
    int32_t* deg = calloc(N + 1, sizeof(int32_t));
    for (int e = 0; e < E; ++e) { deg[ai[e] + 1]++; deg[n + aj[e] + 1]++; }
    for (int v = 0; v < N; ++v) deg[v + 1] += deg[v];
    int32_t* adj = malloc(sizeof(int32_t) * 2 * E); int32_t* cur = malloc(sizeof(int32_t) * N);
    memcpy(cur, deg, sizeof(int32_t) * N);
    for (int e = 0; e < E; ++e) { adj[cur[ai[e]]++] = e; adj[cur[n + aj[e]]++] = e; }
    for (int v = 0; v < N; ++v) { g.parent[v] = -2; }
    g.parent[root] = -1; g.pi[root] = 0; g.flow[root] = 0;
    int32_t* stack = malloc(sizeof(int32_t) * N);
    int cnt = 0;
    g.order[cnt] = root; g.pos[root] = cnt++;
    /* components: rooted at their lowest-index source, attached to the root by a zero-flow artificial arc */
    for (int s0 = 0; s0 < n; ++s0) {
      if (g.parent[s0] != -2) continue;
      g.parent[s0] = root; g.flow[s0] = 0; g.pi[s0] = -BIG;
      int sp = 0; stack[sp++] = s0;
      while (sp > 0) {
        const int v = stack[--sp];
        g.order[cnt] = v; g.pos[v] = cnt++;
        for (int k = deg[v + 1] - 1; k >= deg[v]; --k) {
          const int e = adj[k];
          const int w = v < n ? n + aj[e] : ai[e];
          if (w == g.parent[v]) continue;
          if (g.parent[w] != -2) { fprintf(stderr, "init arcs contain a cycle\n"); return -1; }
          g.parent[w] = v; g.flow[w] = af[e];
          const int64_t c = arc_cost(&g, ai[e], aj[e]);
          g.pi[w] = w >= n ? g.pi[v] + c : g.pi[v] - c;   /* pi_j = pi_i + c */
          stack[sp++] = w;
        }
      }
    }
    if (cnt != N) { fprintf(stderr, "init forest does not span: %d of %d (a target without arcs?)\n", cnt, N); return -1; }
    for (int x = N - 1; x >= 0; --x) g.size[g.order[x]] = 1;
    for (int x = N - 1; x >= 1; --x) { const int v = g.order[x]; g.size[g.parent[v]] += g.size[v]; }
    /* feasibility check */
    int64_t* bal = calloc(N, sizeof(int64_t));
    for (int v = 0; v < n + m; ++v) { const int p = g.parent[v]; if (p == root) continue; bal[v] += g.flow[v]; bal[p] += g.flow[v]; }
    for (int v = 0; v < n + m; ++v) { const int64_t want = v < n ? supply[v] : demand[v - n]; if (bal[v] != want) { fprintf(stderr, "infeasible init at node %d: %lld vs %lld\n", v, (long long)bal[v], (long long)want); return -1; } }
    free(bal); free(deg); free(adj); free(cur); free(stack);
  }
  const int64_t n_arcs = (int64_t)n * m;
  int64_t B = block_size > 0 ? block_size : (int64_t)sqrt((double)n_arcs);
  if (B < 10) B = 10;
  if (B > n_arcs) B = n_arcs;
  int64_t next_arc = 0, it = 0, n_degen = 0, sum_len = 0, sum_range = 0, max_len = 0;
  for (;;) {
    int64_t scanned = 0, best_e = -1, best_rc = 0;
    while (best_e < 0 && scanned < n_arcs) {
      int64_t cnt = B < n_arcs - scanned ? B : n_arcs - scanned;
      int64_t e = next_arc;
      int j = (int)(e / n), i = (int)(e - (int64_t)j * n);
      for (int64_t k = 0; k < cnt; ++k) {
        const int64_t rc = arc_cost(&g, i, j) + g.pi[i] - g.pi[n + j];
        if (rc < best_rc) { best_rc = rc; best_e = (int64_t)j * n + i; }
        if (++i == n) { i = 0; if (++j == m) j = 0; }
      }
      scanned += cnt; next_arc += cnt;
      if (next_arc >= n_arcs) next_arc -= n_arcs;
    }
    if (best_e < 0) break;
    if (it >= max_iter) break;
    ++it;
    const int ej = (int)(best_e / n), ei = (int)(best_e - (int64_t)ej * n);
    const int u_i = ei, u_j = n + ej;
    int lenA = 0, lenB = 0;
    for (int v = u_i; !is_anc(&g, v, u_j); v = g.parent[v]) pathA[lenA++] = v;
    for (int v = u_j; !is_anc(&g, v, u_i); v = g.parent[v]) pathB[lenB++] = v;
    int64_t delta = INT64_MAX;
    int side = -1, idx = -1;
    for (int a = 0; a < lenA; ++a) { const int v = pathA[a]; if (v < n && g.flow[v] < delta) { delta = g.flow[v]; side = 0; idx = a; } }
    for (int b = 0; b < lenB; ++b) { const int v = pathB[b]; if (v >= n && g.flow[v] <= delta) { delta = g.flow[v]; side = 1; idx = b; } }
    if (side < 0) { fprintf(stderr, "unbounded\n"); break; }
    if (delta == 0) ++n_degen;
    sum_len += lenA + lenB; if (lenA + lenB > max_len) max_len = lenA + lenB;
    for (int a = 0; a < lenA; ++a) { const int v = pathA[a]; g.flow[v] += (v < n) ? -delta : delta; }
    for (int b = 0; b < lenB; ++b) { const int v = pathB[b]; g.flow[v] += (v >= n) ? -delta : delta; }
    int32_t* pathX = side == 0 ? pathA : pathB; int32_t* pathY = side == 0 ? pathB : pathA;
    const int lenX = side == 0 ? lenA : lenB, lenY = side == 0 ? lenB : lenA;
    const int k = idx, q = pathX[k], t_in = side == 0 ? u_j : u_i;
    const int s = g.size[q], a0 = g.pos[q];
    const int64_t shift = side == 0 ? -best_rc : best_rc;
    for (int x = a0; x < a0 + s; ++x) g.pi[g.order[x]] += shift;
    for (int t = 0; t <= k; ++t) { P[t] = g.pos[pathX[t]]; S[t] = g.size[pathX[t]]; F[t] = g.flow[pathX[t]]; }
    const int pt = g.pos[t_in];
    const int ptr = pt < a0 ? pt : pt - s;
    const int lo = pt < a0 ? pt + 1 : a0, hi = pt < a0 ? a0 + s : pt + 1;
    sum_range += hi - lo;
    memcpy(g.tmp + lo, g.order + lo, sizeof(int32_t) * (hi - lo));
    for (int x = lo; x < hi; ++x) {
      const int v = g.tmp[x];
      int f;
      if (x < a0 || x >= a0 + s) { const int xr = x < a0 ? x : x - s; f = xr <= ptr ? xr : xr + s; }
      else {
        int l = 0, h = k;
        while (l < h) { const int mid = (l + h) / 2; if (P[mid] <= x && x < P[mid] + S[mid]) h = mid; else l = mid + 1; }
        const int t = l; int r;
        if (t == 0) r = x - P[0]; else if (x == P[t]) r = S[t - 1]; else if (x < P[t - 1]) r = S[t - 1] + 1 + (x - (P[t] + 1)); else r = x - P[t];
        f = ptr + 1 + r;
      }
      g.order[f] = v; g.pos[v] = f;
    }
    for (int t = k; t >= 1; --t) { g.parent[pathX[t]] = pathX[t - 1]; g.flow[pathX[t]] = F[t - 1]; g.size[pathX[t]] = s - S[t - 1]; }
    g.parent[pathX[0]] = t_in; g.flow[pathX[0]] = delta; g.size[pathX[0]] = s;
    for (int t = k + 1; t < lenX; ++t) g.size[pathX[t]] -= s;
    for (int t = 0; t < lenY; ++t) g.size[pathY[t]] += s;
  }
  double total = 0.0;
  for (int v = 0; v < n + m; ++v) {
    const int p = g.parent[v];
    if (p == root || g.flow[v] == 0) continue;
    const int i = v < n ? v : p, j = (v < n ? p : v) - n;
    total += (double)g.flow[v] * (double)arc_cost(&g, i, j);
  }
  *cost_out = total;
  if (out_parent) memcpy(out_parent, g.parent, sizeof(int32_t) * N);
  if (out_flow) memcpy(out_flow, g.flow, sizeof(int64_t) * N);
  if (stats) { stats[0] = n_degen; stats[1] = sum_len; stats[2] = sum_range; stats[3] = max_len; }
  free(g.parent); free(g.pos); free(g.size); free(g.order); free(g.tmp); free(g.flow); free(g.pi);
  free(pathA); free(pathB); free(P); free(S); free(F);
  return it;
}
