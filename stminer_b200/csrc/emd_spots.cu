// Batched exact spot-level EMD for sm_100a — one CTA per (source image, target image) problem.
//
// Restates calculate_ot_distance (STMiner/Algorithm/distance.py:191-202): supports are integer grid
// coordinates, masses are normalised to 1, the cost is the squared Euclidean distance and the result is
// the exact transport-LP optimum sum_ij G_ij M_ij that POT's emd2 (network simplex, POT==0.9.1) returns.
//
// The solver is a primal network simplex on the bipartite transport graph, arranged for one CTA:
//   * the n x m cost matrix is never materialised: arc costs come from the coordinates in shared memory;
//   * masses are scaled to integers (total 2^50), potentials are integers, so every comparison is exact
//     and the pivot sequence is the one of the CPU oracle (oracle/emd_oracle.c) for the same block size;
//   * pricing is a block search: the CTA evaluates a block of arcs in parallel and enters the most negative
//     reduced cost of the first block that has one (LEMON's rule);
//   * the spanning tree is stored as parent / preorder position / subtree size (uint16) in shared memory, a
//     subtree being a contiguous segment of the preorder array — the two cycle paths are chased by two
//     lanes, everything else (leaving-arc search, flow update, potential shift of the re-hung subtree,
//     preorder re-numbering) is data-parallel over the CTA;
//   * arc flows (int64) live in a per-CTA slice of a global workspace (touched only along the cycle).
#include <limits.h>
#include <math.h>

#include <algorithm>
#include <type_traits>

#include "stm_common.cuh"

namespace stm {
namespace {

constexpr int EMD_PREP_THREADS = 256;
constexpr int EMD_MAXPATH = 2048;
#ifndef STM_EMD_T_BIG
#define STM_EMD_T_BIG 512    // CTA size of the one-CTA-per-SM kernel: the pricing scan is issue-bound (same time at 512 or 1,024
                             // threads) while every barrier-delimited serial phase is cheaper with 16 warps: 116 vs 111 problems/s
#endif
// sources per thread kept in registers by the whole-row pricing path: 5,120 sources per CTA for the big kernel
#ifndef STM_EMD_MINB_SMALL
#define STM_EMD_MINB_SMALL 1   // CTAs per SM the 256-thread kernel is compiled for (register bound)
#endif
constexpr int emd_rmax(int threads, bool big) { return big ? (5120 + threads - 1) / threads : 5; }
constexpr double EMD_SCALE = 1125899906842624.0;  // 2^50
constexpr unsigned short NONE16 = 0xffffu;

struct EmdArgs {
  const int32_t* src_x; const int32_t* src_y; const double* src_mass; int32_t n_src;
  const int64_t* tgt_ptr; const int32_t* tgt_x; const int32_t* tgt_y; const double* tgt_mass;
  int32_t G; int32_t cap_nodes;   // shared-memory capacity in nodes
  int32_t flow_smem_nodes;        // arc flows of nodes below this index live in shared memory, the rest in flow_ws
  int64_t max_iter; int32_t block_size;
  int64_t* src_int;               // [n_src] integer source masses (prep kernel)
  int64_t* flow_ws;               // [gridDim.x][cap_nodes]
  unsigned int* counter;          // dynamic problem queue
  const int32_t* order;           // optional processing order
  double* out_cost; int32_t* out_status; int64_t* out_iters;
};

__device__ __forceinline__ long long shfl_xor_ll(long long v, int o) {
  return __shfl_xor_sync(FULL_MASK, v, o);
}

// Deterministic block sum of doubles (fixed strided order + fixed tree).
template <int EMD_THREADS>
__device__ double block_sum(double v, double* s_scr) {
  constexpr int EMD_WARPS = EMD_THREADS / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) s_scr[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < EMD_WARPS; ++w) t += s_scr[w];
  return t;
}

// mass[i] / sum * 2^50 rounded to integers; the largest entry absorbs the rounding so the total is exactly 2^50.
// Single CTA; `out` may be global or shared.
template <int EMD_THREADS>
__device__ void integerize(const double* mass, int n, int64_t* out, double* s_scr, long long* s_ll) {
  constexpr int EMD_WARPS = EMD_THREADS / 32;
  double part = 0.0;
  for (int i = threadIdx.x; i < n; i += EMD_THREADS) part += mass[i];
  const double total = block_sum<EMD_THREADS>(part, s_scr);
  long long isum = 0, best = -1;
  int best_i = 0;
  for (int i = threadIdx.x; i < n; i += EMD_THREADS) {
    const long long q = (long long)floor(mass[i] / total * EMD_SCALE + 0.5);
    out[i] = q;
    isum += q;
    if (q > best) { best = q; best_i = i; }
  }
  // reduce (sum, argmax with lowest index on ties)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 16; o > 0; o >>= 1) {
    isum += shfl_xor_ll(isum, o);
    const long long ob = shfl_xor_ll(best, o);
    const int oi = __shfl_xor_sync(FULL_MASK, best_i, o);
    if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
  }
  __syncthreads();
  if (lane == 0) { s_ll[3 * warp] = isum; s_ll[3 * warp + 1] = best; s_ll[3 * warp + 2] = best_i; }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long tsum = 0, tb = -1, ti = 0;
    for (int w = 0; w < EMD_WARPS; ++w) {
      tsum += s_ll[3 * w];
      if (s_ll[3 * w + 1] > tb || (s_ll[3 * w + 1] == tb && s_ll[3 * w + 2] < ti)) { tb = s_ll[3 * w + 1]; ti = s_ll[3 * w + 2]; }
    }
    out[ti] += (long long)EMD_SCALE - tsum;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(EMD_PREP_THREADS) emd_prep_kernel(const double* src_mass, int n, int64_t* src_int) {
  __shared__ double s_scr[EMD_PREP_THREADS / 32];
  __shared__ long long s_ll[3 * EMD_PREP_THREADS / 32];
  integerize<EMD_PREP_THREADS>(src_mass, n, src_int, s_scr, s_ll);
}

// WIDE = 64-bit potentials and reduced costs for problems whose coordinate span makes (span^2 + 1) * nodes exceed the
// int32 range of the default kernel (scattered integer coordinates, long thin tissues): 22 instead of 18 bytes of shared
// memory per node and the general (non register-resident) pricing path, same pivot sequence.
template <int EMD_THREADS, bool WIDE, bool ONE_PER_SM>
__global__ void __launch_bounds__(EMD_THREADS, ONE_PER_SM ? 1 : STM_EMD_MINB_SMALL) emd_kernel(const EmdArgs p) {
  constexpr int EMD_WARPS = EMD_THREADS / 32;
  constexpr int EMD_RMAX_T = emd_rmax(EMD_THREADS, ONE_PER_SM);
  using pi_t = typename std::conditional<WIDE, long long, int>::type;
  constexpr pi_t PI_NONE = WIDE ? (pi_t)LLONG_MAX : (pi_t)INT_MAX;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int cap = p.cap_nodes;
  // s_pi holds ADJUSTED potentials: pi_i + |x_i|^2 for sources, pi_j - |x_j|^2 for sinks, so that the reduced
  // cost of arc (i, j) is  s_pi[i] - s_pi[j] - 2 x_i.x_j  (one subtraction + two IMAD per arc); a subtree shift
  // adds the same constant to either form.
  // arc flows (int64): nodes [0, fcap) in shared memory — whatever is left of the CTA's budget once the tree is
  // placed — the others in the per-CTA global slice; the leaving-arc search and the flow update then run at
  // shared-memory instead of L2 latency
  const int fcap = p.flow_smem_nodes;
  long long* s_flow = reinterpret_cast<long long*>(smem_raw);           // [fcap]
  pi_t* s_pi = reinterpret_cast<pi_t*>(s_flow + fcap);                   // [cap]
  int* s_xy = reinterpret_cast<int*>(s_pi + cap);                        // [cap] {int16 x | int16 y}
  unsigned short* s_parent = reinterpret_cast<unsigned short*>(s_xy + cap);
  unsigned short* s_pos = s_parent + cap;
  unsigned short* s_size = s_pos + cap;
  unsigned short* s_order = s_size + cap;
  unsigned short* s_tmp = s_order + cap;
  unsigned short* s_pathA = s_tmp + cap;                                 // [EMD_MAXPATH]
  unsigned short* s_pathB = s_pathA + EMD_MAXPATH;
  unsigned short* s_P = s_pathB + EMD_MAXPATH;
  unsigned short* s_S = s_P + EMD_MAXPATH;
  __shared__ double s_scr[EMD_WARPS];
  __shared__ long long s_ll[3 * EMD_WARPS];
  __shared__ int2 s_wkey[2 * EMD_WARPS];   // per-warp (reduced cost, position) of the pricing arg-min, double-buffered
  __shared__ int s_i[16];
  __shared__ int s_wsum[EMD_WARPS];
  __shared__ long long s_wrc64[WIDE ? 2 * EMD_WARPS : 1];   // per-warp reduced costs of the WIDE pricing arg-min
  __shared__ long long s_delta;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = p.n_src;
  int64_t* flow = p.flow_ws + (size_t)blockIdx.x * cap;
  auto FL = [&](int v) -> long long& { return v < fcap ? s_flow[v] : *reinterpret_cast<long long*>(flow + v); };

  for (;;) {
    __syncthreads();
    if (tid == 0) s_i[0] = (int)atomicAdd(p.counter, 1u);
    __syncthreads();
    const int slot = s_i[0];
    if (slot >= p.G) break;
    const int g = p.order ? p.order[slot] : slot;
    const int64_t t0 = p.tgt_ptr[g];
    const int m = (int)(p.tgt_ptr[g + 1] - t0);
    const int N = n + m + 1, root = n + m;
    if (m <= 0 || n <= 0) {
      if (tid == 0) { p.out_cost[g] = 0.0; p.out_status[g] = 2; if (p.out_iters) p.out_iters[g] = 0; }
      continue;
    }
    // ---- stage coordinates (relative to the source's first spot), bounding box, capacity checks ----
    const int ox = p.src_x[0], oy = p.src_y[0];
    int xmin = INT_MAX, xmax = INT_MIN, ymin = INT_MAX, ymax = INT_MIN;
    // N <= 30 * threads: the cycle-path scan keeps one bit per preorder position of a thread's run in a 32-bit mask
    const bool fits = N <= cap && N < 65535 && N <= 30 * EMD_THREADS && (int64_t)n * m < 2147483647LL;
    for (int v = tid; v < n + m; v += EMD_THREADS) {
      const int x = v < n ? p.src_x[v] : p.tgt_x[t0 + v - n], y = v < n ? p.src_y[v] : p.tgt_y[t0 + v - n];
      xmin = min(xmin, x); xmax = max(xmax, x); ymin = min(ymin, y); ymax = max(ymax, y);
      if (fits) s_xy[v] = ((x - ox) & 0xffff) | ((y - oy) << 16);
    }
    xmin = __reduce_min_sync(FULL_MASK, xmin); xmax = __reduce_max_sync(FULL_MASK, xmax);
    ymin = __reduce_min_sync(FULL_MASK, ymin); ymax = __reduce_max_sync(FULL_MASK, ymax);
    if (lane == 0) { s_ll[3 * warp] = ((long long)xmin << 32) | (unsigned)xmax; s_ll[3 * warp + 1] = ((long long)ymin << 32) | (unsigned)ymax; }
    __syncthreads();
    for (int w = 0; w < EMD_WARPS; ++w) {
      xmin = min(xmin, (int)(s_ll[3 * w] >> 32)); xmax = max(xmax, (int)(unsigned)s_ll[3 * w]);
      ymin = min(ymin, (int)(s_ll[3 * w + 1] >> 32)); ymax = max(ymax, (int)(unsigned)s_ll[3 * w + 1]);
    }
    __syncthreads();
    const long long maxc = (long long)(xmax - xmin) * (xmax - xmin) + (long long)(ymax - ymin) * (ymax - ymin);
    const long long BIG = (maxc + 1) * (long long)N;
    const bool coords_ok = (xmax - ox) < 32768 && (ox - xmin) < 32768 && (ymax - oy) < 32768 && (oy - ymin) < 32768;
    const long long pi_limit = WIDE ? (1LL << 61) : 2147483647LL;
    if (!fits || !coords_ok || 2 * BIG + maxc >= pi_limit) {   // range of the potentials / int16 coordinates / capacity
      if (tid == 0) { p.out_cost[g] = nan(""); p.out_status[g] = 3; if (p.out_iters) p.out_iters[g] = 0; }
      continue;
    }
    // ---- integer masses and the initial star basis ----------------------------------------------
    for (int i = tid; i < n; i += EMD_THREADS) flow[i] = p.src_int[i];
    integerize<EMD_THREADS>(p.tgt_mass + t0, m, flow + n, s_scr, s_ll);
    for (int v = tid; v < n + m; v += EMD_THREADS) {
      s_parent[v] = (unsigned short)root; s_order[v + 1] = (unsigned short)v; s_pos[v] = (unsigned short)(v + 1);
      const int xy = s_xy[v], xr = (xy << 16) >> 16, yr = xy >> 16, r2 = xr * xr + yr * yr;
      s_size[v] = 1; s_pi[v] = v < n ? (pi_t)(-BIG + r2) : (pi_t)(BIG - r2);
    }
    if (tid == 0) {
      s_order[0] = (unsigned short)root; s_pos[root] = 0; s_size[root] = (unsigned short)N; s_parent[root] = NONE16;
      s_pi[root] = 0; flow[root] = 0; s_i[4] = 0;
    }
    __syncthreads();
    for (int v = tid; v < min(fcap, N); v += EMD_THREADS) s_flow[v] = flow[v];
    __syncthreads();

    const int n_arcs = n * m;
    // pricing block: LEMON uses sqrt(#arcs); a CTA prices in parallel, so larger blocks (fewer, better pivots and
    // fewer barriers per arc) win — 32 arcs per thread measured best on Visium-size problems
    int B = p.block_size > 0 ? p.block_size : max((int)sqrt((double)n_arcs), ONE_PER_SM ? 32768 : 32 * EMD_THREADS);
    if (p.block_size <= 0 && n <= EMD_RMAX_T * EMD_THREADS) B = max(1, (B + n / 2) / n) * n;   // whole target rows
    if (B < 10) B = 10;
    if (B > n_arcs) B = n_arcs;
    const bool whole_rows = !WIDE && B % n == 0 && n <= EMD_RMAX_T * EMD_THREADS;   // n_arcs and every block start are multiples of n
    int next_arc = 0;
    long long it = 0;
    int status = 0;
    unsigned round = 0;
    __syncthreads();

#ifdef STM_EMD_PROFILE
    long long c_price = 0, c_walk = 0, c_leave = 0, c_update = 0, n_scanned = 0, c_t0 = 0, sum_len = 0, sum_s = 0, sum_range = 0;
    long long c_scan = 0, c_red = 0, n_blocks = 0, c_t1 = 0;
#define PROF_T(x) { const long long _c = clock64(); x += _c - c_t0; c_t0 = _c; }
#else
#define PROF_T(x)
#endif
    for (;;) {
#ifdef STM_EMD_PROFILE
      c_t0 = clock64();
#endif
      // ---- entering arc: block search over arcs e = j*n + i ----------------------------------------
      int scanned = 0, best_e = -1;
      pi_t best_rc = 0;
      while (scanned < n_arcs) {
        const int cnt = min(B, n_arcs - scanned);
        // The block is arcs [next_arc, next_arc + cnt) of the j-major arc order (wrapping): a few target rows,
        // each a contiguous range of sources.  Per row the target's terms stay in registers and the inner loop is
        // two shared loads, two IMADs and a compare per arc.
#ifdef STM_EMD_PROFILE
        c_t1 = clock64(); ++n_blocks;
#endif
        pi_t g_rc = PI_NONE;         // most negative reduced cost of the block (PI_NONE: none below zero)
        unsigned g_k = 0xffffffffu;  // its position in block order (ties: first in the serial scan order)
        {
          pi_t brc = 0;
          int bi = -1, bj_best = 0;
          bool fast_done = false;
          if constexpr (!WIDE) {
          if (whole_rows) {
            fast_done = true;
            // Fast path (the default block size is a whole number of target rows): every thread keeps ITS sources'
            // terms in registers for the whole block and meets each target row with no shared-memory traffic
            // beyond the row's three broadcast words — six integer instructions per arc.
            int sx[EMD_RMAX_T], sy[EMD_RMAX_T], sp[EMD_RMAX_T];
#pragma unroll
            for (int r = 0; r < EMD_RMAX_T; ++r) {
              const int i = tid + r * EMD_THREADS;
              const int xy = i < n ? s_xy[i] : 0;
              sx[r] = (xy << 16) >> 16; sy[r] = xy >> 16;
              sp[r] = i < n ? s_pi[i] : INT_MAX / 2;     // never the minimum
            }
            // The scan only tracks the thread's minimum (an add, two IMADs and a share of a 3-input min per arc); the
            // arc that attains the block minimum is located afterwards by the few threads whose minimum equals it.
            const int j0 = next_arc / n;
            int j = j0;
            for (int rr = cnt / n; rr > 0; --rr) {
              const int txy = s_xy[n + j], nx2 = -2 * ((txy << 16) >> 16), ny2 = -2 * (txy >> 16), bj = s_pi[n + j];
#pragma unroll
              for (int r = 0; r < EMD_RMAX_T; ++r) brc = min(brc, sx[r] * nx2 + (sy[r] * ny2 + (sp[r] - bj)));
              if (++j >= m) j = 0;
            }
            // ---- level 1: the block's most negative reduced cost --------------------------------------------
            {
              const int w_rc = __reduce_min_sync(FULL_MASK, brc);
              int* wrc = reinterpret_cast<int*>(s_wkey + (round & 1) * EMD_WARPS);
              if (lane == 0) wrc[warp] = w_rc;
              __syncthreads();
              g_rc = __reduce_min_sync(FULL_MASK, lane < EMD_WARPS ? wrc[lane] : 0);
              ++round;
            }
            if (g_rc < 0) {
              // ---- level 2: its first position in block order (rows from the block start, sources ascending) ----
              unsigned myk = 0xffffffffu;
              if (brc == g_rc) {
                j = j0;
                for (int rr = cnt / n; rr > 0 && myk == 0xffffffffu; --rr) {
                  const int txy = s_xy[n + j], nx2 = -2 * ((txy << 16) >> 16), ny2 = -2 * (txy >> 16), bj = s_pi[n + j];
#pragma unroll
                  for (int r = 0; r < EMD_RMAX_T; ++r) {
                    const int rc = sx[r] * nx2 + (sy[r] * ny2 + (sp[r] - bj));
                    if (rc == g_rc && myk == 0xffffffffu) {
                      int k = j * n + tid + r * EMD_THREADS - next_arc;
                      if (k < 0) k += n_arcs;
                      myk = (unsigned)k;
                    }
                  }
                  if (++j >= m) j = 0;
                }
              }
              const unsigned w_k = __reduce_min_sync(FULL_MASK, myk);
              unsigned* wk = reinterpret_cast<unsigned*>(s_wkey + (round & 1) * EMD_WARPS);
              if (lane == 0) wk[warp] = w_k;
              __syncthreads();
              g_k = __reduce_min_sync(FULL_MASK, lane < EMD_WARPS ? wk[lane] : 0xffffffffu);
              ++round;
            } else {
              g_rc = PI_NONE;
            }
          }
          }
          if (!fast_done) {
            // general block: arcs [next_arc, next_arc + cnt) of the j-major order (wrapping) = a few target rows,
            // each a contiguous range of sources
            int left = cnt;
            int j = next_arc / n, lo = next_arc - j * n;
            while (left > 0) {
              const int hi = min(n, lo + left);
              const int txy = s_xy[n + j], nx2 = -2 * ((txy << 16) >> 16), ny2 = -2 * (txy >> 16);
              const pi_t bj = s_pi[n + j];
              for (int i = lo + tid; i < hi; i += EMD_THREADS) {
                const int xy = s_xy[i];
                const pi_t rc = (pi_t)((xy << 16) >> 16) * nx2 + ((pi_t)(xy >> 16) * ny2 + (s_pi[i] - bj));
                if (rc < brc) { brc = rc; bi = i; bj_best = j; }
              }
              left -= hi - lo;
              lo = 0;
              if (++j >= m) j = 0;
            }
          }
#ifdef STM_EMD_PROFILE
          { const long long _c = clock64(); c_scan += _c - c_t1; c_t1 = _c; }
#endif
          if (!fast_done) {
          int k = 0;
          if (bi >= 0) {
            k = bj_best * n + bi - next_arc;   // position in block order decides ties, as in the serial scan
            if (k < 0) k += n_arcs;
          }
          // lexicographic (reduced cost, position) minimum in two 32-bit REDUX steps per level: over the warp, then —
          // through one shared entry per warp, double-buffered, one barrier — over the CTA.  (A 64-bit shared
          // atomicMin is a CAS loop: 32 warps contending on one slot cost several thousand cycles per block.)
          if constexpr (!WIDE) {
            const int w_rc = __reduce_min_sync(FULL_MASK, bi >= 0 ? brc : INT_MAX);
            const unsigned w_k = __reduce_min_sync(FULL_MASK, (bi >= 0 && brc == w_rc) ? (unsigned)k : 0xffffffffu);
            int2* wkey = s_wkey + (round & 1) * EMD_WARPS;
            if (lane == 0) wkey[warp] = make_int2(w_rc, (int)w_k);
            __syncthreads();
            const int2 e = lane < EMD_WARPS ? wkey[lane] : make_int2(INT_MAX, -1);
            g_rc = __reduce_min_sync(FULL_MASK, e.x);
            g_k = __reduce_min_sync(FULL_MASK, e.x == g_rc ? (unsigned)e.y : 0xffffffffu);
          } else {
            // 64-bit reduced costs: lexicographic (rc, position) minimum by shuffles, over the warp and then over the
            // per-warp entries
            long long r64 = bi >= 0 ? (long long)brc : LLONG_MAX;
            unsigned k32 = bi >= 0 ? (unsigned)k : 0xffffffffu;
            for (int o = 16; o > 0; o >>= 1) {
              const long long orc = shfl_xor_ll(r64, o);
              const unsigned ok = __shfl_xor_sync(FULL_MASK, k32, o);
              if (orc < r64 || (orc == r64 && ok < k32)) { r64 = orc; k32 = ok; }
            }
            long long* wr = s_wrc64 + (round & 1) * EMD_WARPS;
            unsigned* wk = reinterpret_cast<unsigned*>(s_wkey + (round & 1) * EMD_WARPS);
            if (lane == 0) { wr[warp] = r64; wk[warp] = k32; }
            __syncthreads();
            r64 = lane < EMD_WARPS ? wr[lane] : LLONG_MAX;
            k32 = lane < EMD_WARPS ? wk[lane] : 0xffffffffu;
            for (int o = 16; o > 0; o >>= 1) {
              const long long orc = shfl_xor_ll(r64, o);
              const unsigned ok = __shfl_xor_sync(FULL_MASK, k32, o);
              if (orc < r64 || (orc == r64 && ok < k32)) { r64 = orc; k32 = ok; }
            }
            g_rc = (pi_t)r64; g_k = k32;
          }
          ++round;
          }
        }
#ifdef STM_EMD_PROFILE
        { const long long _c = clock64(); c_red += _c - c_t1; c_t1 = _c; }
#endif
        const int start = next_arc;
        scanned += cnt;
        next_arc += cnt;
        if (next_arc >= n_arcs) next_arc -= n_arcs;
        if (g_rc < 0) {
          best_rc = g_rc;
          best_e = start + (int)g_k;
          if (best_e >= n_arcs) best_e -= n_arcs;
          break;
        }
      }
      PROF_T(c_price)
#ifdef STM_EMD_PROFILE
      n_scanned += scanned;
#endif
      if (best_e < 0) break;                      // optimal
      if (it >= p.max_iter) { status = 1; break; }
      ++it;
      const int ej = best_e / n, ei = best_e - ej * n;
      const int u_i = ei, u_j = n + ej;
      // ---- cycle: the two tree paths from the entering arc's endpoints up to their join node -----------------
      // A node v is an ancestor of u iff pos[v] <= pos[u] < pos[v] + size[v].  Instead of chasing parent pointers from
      // both endpoints (two lanes, ~180 cycles per step, everyone else waiting), every thread tests a run of consecutive
      // PREORDER POSITIONS: path A = ancestors-or-self of u_i that are not ancestors of u_j, path B the converse; along
      // a path the position decreases, so an ordered compaction (one block scan) yields both paths in walking order.
      {
        const int pu = s_pos[u_i], pw = s_pos[u_j];
        int per = (N + EMD_THREADS - 1) / EMD_THREADS;             // <= 30 for every problem that fits a CTA
        per += (2 - per) & 3;   // per = 2 (mod 4): a lane stride of per uint16 is an odd number of 4-byte words, so the
                                // loads of s_order[x0 + q] across a warp hit 32 different banks
        const int x0 = tid * per, x1 = min(N, x0 + per);
        unsigned mA = 0u, mB = 0u;
        constexpr int PER_UNROLL = 4;                              // independent load pairs in flight
        const int xlim = min(x1, max(pu, pw) + 1);                 // ancestors precede their descendants
        for (int xb = x0; xb < xlim; xb += PER_UNROLL) {
          int vv[PER_UNROLL], sz[PER_UNROLL];
#pragma unroll
          for (int q = 0; q < PER_UNROLL; ++q) vv[q] = xb + q < xlim ? s_order[xb + q] : root;
#pragma unroll
          for (int q = 0; q < PER_UNROLL; ++q) sz[q] = xb + q < xlim ? (int)s_size[vv[q]] : 0;   // 0: never an ancestor
#pragma unroll
          for (int q = 0; q < PER_UNROLL; ++q) {
            const int x = xb + q;
            const unsigned inU = (unsigned)(pu - x) < (unsigned)sz[q], inW = (unsigned)(pw - x) < (unsigned)sz[q];
            mA |= (inU & ~inW & 1u) << ((x - x0) & 31);
            mB |= (inW & ~inU & 1u) << ((x - x0) & 31);
          }
        }
        // exclusive block scan of the two counts, packed in one word
        const int cnt = __popc(mA) | (__popc(mB) << 16);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL_MASK, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        int wtot = lane < EMD_WARPS ? s_wsum[lane] : 0, wincl = wtot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL_MASK, wincl, o); if (lane >= o) wincl += t; }
        const int total = __shfl_sync(FULL_MASK, wincl, 31);
        const int before = __shfl_sync(FULL_MASK, wincl - wtot, warp) + incl - cnt;
        const int nA = total & 0xffff, nB = total >> 16;
        if (tid == 0) { s_i[1] = nA; s_i[2] = nB; }
        if (nA <= EMD_MAXPATH && nB <= EMD_MAXPATH) {
          int ra = before & 0xffff, rb = before >> 16;             // rank in ascending position = distance from the join
          for (unsigned m = mA; m; m &= m - 1) s_pathA[nA - 1 - ra++] = s_order[x0 + __ffs(m) - 1];
          for (unsigned m = mB; m; m &= m - 1) s_pathB[nB - 1 - rb++] = s_order[x0 + __ffs(m) - 1];
        }
      }
      __syncthreads();
      const int lenA = s_i[1], lenB = s_i[2];
      PROF_T(c_walk)
      if (lenA > EMD_MAXPATH || lenB > EMD_MAXPATH) { status = 4; break; }
      // ---- leaving arc (warp 0): min flow over arcs traversed against their direction; ties resolved by
      //      the strongly-feasible rule (i side: closest to i; j side wins and takes the one closest to the join)
      if (warp == 0) {
        long long dmin = LLONG_MAX;
        int seq = -1;
        for (int a = lane; a < lenA; a += 32) {
          const int v = s_pathA[a];
          if (v < n) { const long long f = FL(v); const int sq = lenA - 1 - a; if (f < dmin || (f == dmin && sq > seq)) { dmin = f; seq = sq; } }
        }
        for (int b = lane; b < lenB; b += 32) {
          const int v = s_pathB[b];
          if (v >= n) { const long long f = FL(v); const int sq = lenA + b; if (f < dmin || (f == dmin && sq > seq)) { dmin = f; seq = sq; } }
        }
        for (int o = 16; o > 0; o >>= 1) {
          const long long od = shfl_xor_ll(dmin, o);
          const int os = __shfl_xor_sync(FULL_MASK, seq, o);
          if (od < dmin || (od == dmin && os > seq)) { dmin = od; seq = os; }
        }
        if (lane == 0) { s_delta = dmin; s_i[3] = seq; }
      }
      __syncthreads();
      const long long delta = s_delta;
      const int seq = s_i[3];
      PROF_T(c_leave)
      if (seq < 0) { status = 5; break; }        // unbounded: cannot happen for a transport LP
      const int side = seq >= lenA ? 1 : 0;
      const int k = side == 0 ? lenA - 1 - seq : seq - lenA;   // leaving arc at path index k of its side
      const unsigned short* pathX = side == 0 ? s_pathA : s_pathB;
      const unsigned short* pathY = side == 0 ? s_pathB : s_pathA;
      const int lenX = side == 0 ? lenA : lenB, lenY = side == 0 ? lenB : lenA;
      const int q = pathX[k];
      const int t_in = side == 0 ? u_j : u_i;
      const int s = s_size[q], a0 = s_pos[q];
      const int pt = s_pos[t_in];
      const int ptr = pt < a0 ? pt : pt - s;
      const pi_t shift = side == 0 ? -best_rc : best_rc;
      // ---- flow update along the cycle + old flows / positions / sizes of the reversed path --------
      for (int a = tid; a < lenA; a += EMD_THREADS) { const int v = s_pathA[a]; FL(v) += (v < n) ? -delta : delta; }
      for (int b = tid; b < lenB; b += EMD_THREADS) { const int v = s_pathB[b]; FL(v) += (v >= n) ? -delta : delta; }
      for (int t = tid; t <= k; t += EMD_THREADS) { const int v = pathX[t]; s_P[t] = s_pos[v]; s_S[t] = s_size[v]; }
      // potential shift of the re-hung subtree (a contiguous preorder segment)
      for (int x = a0 + tid; x < a0 + s; x += EMD_THREADS) s_pi[s_order[x]] += shift;
      // affected preorder range
      const int lo = pt < a0 ? pt + 1 : a0, hi = pt < a0 ? a0 + s : pt + 1;
      for (int x = lo + tid; x < hi; x += EMD_THREADS) s_tmp[x] = s_order[x];
      __syncthreads();
      // ---- reversed path: parents, flows, sizes (flows read before anyone overwrites them) ----------
      {
        // every thread handles t = tid, tid + T, ...; two passes: read old flows, then write
        constexpr int MAXT = (EMD_MAXPATH + EMD_THREADS - 1) / EMD_THREADS;
        long long fo[MAXT];
#pragma unroll
        for (int r = 0; r < MAXT; ++r) {
          const int t = tid + r * EMD_THREADS;
          fo[r] = (t >= 1 && t <= k) ? FL(pathX[t - 1]) : 0;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < MAXT; ++r) {
          const int t = tid + r * EMD_THREADS;
          if (t <= k) {
            const int v = pathX[t];
            if (t == 0) { s_parent[v] = (unsigned short)t_in; FL(v) = delta; s_size[v] = (unsigned short)s; }
            else { s_parent[v] = pathX[t - 1]; FL(v) = fo[r]; s_size[v] = (unsigned short)(s - s_S[t - 1]); }
          }
        }
      }
      for (int t = k + 1 + tid; t < lenX; t += EMD_THREADS) s_size[pathX[t]] -= (unsigned short)s;
      for (int t = tid; t < lenY; t += EMD_THREADS) s_size[pathY[t]] += (unsigned short)s;
      // ---- preorder re-numbering of the affected range ------------------------------------------------
      for (int x = lo + tid; x < hi; x += EMD_THREADS) {
        const int v = s_tmp[x];
        int f;
        if (x < a0 || x >= a0 + s) {
          const int xr = x < a0 ? x : x - s;
          f = xr <= ptr ? xr : xr + s;
        } else {
          int l = 0, h = k;   // smallest t with P[t] <= x < P[t] + S[t] (nested intervals)
          while (l < h) { const int mid = (l + h) >> 1; const int Pm = s_P[mid]; if (Pm <= x && x < Pm + s_S[mid]) h = mid; else l = mid + 1; }
          const int t = l, Pt = s_P[t];
          int r;
          if (t == 0) r = x - Pt;
          else if (x == Pt) r = s_S[t - 1];
          else if (x < s_P[t - 1]) r = s_S[t - 1] + 1 + (x - (Pt + 1));
          else r = x - Pt;
          f = ptr + 1 + r;
        }
        s_order[f] = (unsigned short)v;
        s_pos[v] = (unsigned short)f;
      }
      __syncthreads();
      PROF_T(c_update)
#ifdef STM_EMD_PROFILE
      sum_len += lenA + lenB; sum_s += s; sum_range += hi - lo;
#endif
    }
#ifdef STM_EMD_PROFILE
    if (tid == 0 && g < 2)
      printf("[emd prof] g=%d n=%d m=%d pivots=%lld cycles/pivot: price %.0f walk %.0f leave %.0f update %.0f | arcs scanned/pivot %.0f cycle len %.1f subtree %.1f range %.1f | blocks/pivot %.2f, per block: scan %.0f reduce %.0f cycles\n",
             g, n, m, it, (double)c_price / it, (double)c_walk / it, (double)c_leave / it, (double)c_update / it,
             (double)n_scanned / it, (double)sum_len / it, (double)sum_s / it, (double)sum_range / it,
             (double)n_blocks / it, (double)c_scan / n_blocks, (double)c_red / n_blocks);
#endif
    // ---- objective: real tree arcs only ------------------------------------------------------------
    double part = 0.0;
    for (int v = tid; v < n + m; v += EMD_THREADS) {
      const int pr = s_parent[v];
      const long long f = FL(v);
      if (pr != root && f != 0) {
        const int i = v < n ? v : pr, j = v < n ? pr : v;
        const int a = s_xy[i], b = s_xy[j];
        const long long dx = ((a << 16) >> 16) - ((b << 16) >> 16), dy = (a >> 16) - (b >> 16);
        part += (double)f * (double)(dx * dx + dy * dy);
      }
    }
    const double total = block_sum<EMD_THREADS>(part, s_scr);
    if (tid == 0) {
      p.out_cost[g] = total / EMD_SCALE;
      p.out_status[g] = status;
      if (p.out_iters) p.out_iters[g] = it;
    }
  }
}

inline size_t emd_node_bytes(bool wide) { return (wide ? 8 : 4) + 4 + 5 * 2; }   // potential, packed xy, five uint16 tree arrays
inline size_t emd_smem_bytes(int64_t cap_nodes, int64_t flow_nodes = 0, bool wide = false) {
  return (size_t)flow_nodes * 8 + (size_t)cap_nodes * emd_node_bytes(wide) + 4 * (size_t)EMD_MAXPATH * 2 + 64;
}
inline int64_t emd_ws_cap(int64_t n_src, int64_t max_tgt) { return (n_src + max_tgt + 1 + 7) & ~7LL; }

}  // namespace
}  // namespace stm

extern "C" int64_t stm_emd_workspace_bytes(int32_t n_src, int64_t max_tgt, int32_t n_ctas) {
  if (n_src < 0 || max_tgt < 0 || n_ctas <= 0) return -1;
  return 256 + (int64_t)n_src * 8 + (int64_t)n_ctas * stm::emd_ws_cap(n_src, max_tgt) * 8 + 256;
}

namespace stm {
namespace {
int emd_launch(bool wide, const int32_t* src_x, const int32_t* src_y, const double* src_mass, int32_t n_src,
               const int64_t* tgt_ptr, const int32_t* tgt_x, const int32_t* tgt_y,
               const double* tgt_mass, int32_t G, int64_t max_tgt, const int32_t* order,
               int64_t max_iter, int32_t block_size,
               void* workspace, int64_t workspace_bytes, int32_t n_ctas,
               double* out_cost, int32_t* out_status, int64_t* out_iters, void* stream) {
  STM_CHECK_ARG(G >= 0, "G < 0");
  if (G == 0) return STM_OK;
  STM_CHECK_ARG(n_src > 0 && src_x && src_y && src_mass, "empty or null source image");
  STM_CHECK_ARG(tgt_ptr && tgt_x && tgt_y && tgt_mass && out_cost && out_status, "null pointer");
  STM_CHECK_ARG(max_tgt >= 0 && n_ctas > 0 && max_iter >= 0 && block_size >= 0, "bad size argument");
  STM_CHECK_ARG(workspace && workspace_bytes >= stm_emd_workspace_bytes(n_src, max_tgt, n_ctas), "workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int dev = 0, max_optin = 0;
  STM_CHECK_CUDA(cudaGetDevice(&dev));
  STM_CHECK_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  // node capacity of one CTA: the whole tree lives in shared memory; larger problems get status 3
  int64_t cap = emd_ws_cap(n_src, max_tgt);
  const int64_t max_cap = (((int64_t)max_optin - 4096 - (int64_t)emd_smem_bytes(0)) / (int64_t)emd_node_bytes(wide)) & ~7LL;   // static smem + reserve
  if (cap > max_cap) cap = max_cap;
  char* ws = static_cast<char*>(workspace);
  EmdArgs a{};
  a.src_x = src_x; a.src_y = src_y; a.src_mass = src_mass; a.n_src = n_src;
  a.tgt_ptr = tgt_ptr; a.tgt_x = tgt_x; a.tgt_y = tgt_y; a.tgt_mass = tgt_mass;
  a.G = G; a.cap_nodes = (int32_t)cap; a.max_iter = max_iter; a.block_size = block_size;
  a.counter = reinterpret_cast<unsigned int*>(ws);
  a.src_int = reinterpret_cast<int64_t*>(ws + 256);
  a.flow_ws = a.src_int + n_src;
  a.order = order;
  a.out_cost = out_cost; a.out_status = out_status; a.out_iters = out_iters;
  STM_CHECK_CUDA(cudaMemsetAsync(a.counter, 0, 256, st));
  emd_prep_kernel<<<1, EMD_PREP_THREADS, 0, st>>>(src_mass, n_src, a.src_int);
  STM_CHECK_CUDA(cudaGetLastError());
  const size_t tree_smem = emd_smem_bytes((int)cap, 0, wide);
  const int grid = n_ctas < G ? n_ctas : G;
  // Problems that leave room for only one CTA per SM get 1,024 threads (pricing is the dominant phase and
  // scales with the thread count); small problems run several 256-thread CTAs per SM.
  const bool one_per_sm = tree_smem * 2 + 2048 > (size_t)max_optin;
  // what is left of a CTA's share of the SM's shared memory (without lowering the CTAs per SM) holds arc flows
  const int64_t per_sm = one_per_sm ? 1 : std::min<int64_t>(8, (int64_t)max_optin / (int64_t)(tree_smem + 1024));
  const int64_t share = one_per_sm ? (int64_t)max_optin - 4096 : (int64_t)max_optin / per_sm - 1024 - 256;
  int64_t fcap = (share - (int64_t)tree_smem) / 8;
  fcap = std::max<int64_t>(0, std::min<int64_t>(fcap, cap)) & ~1LL;
  a.flow_smem_nodes = (int32_t)fcap;
  const size_t smem = emd_smem_bytes((int)cap, fcap, wide);
  auto launch = [&](auto kern, int threads) -> int {
    STM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, threads, smem, st>>>(a);
    STM_CHECK_CUDA(cudaGetLastError());
    return STM_OK;
  };
  if (one_per_sm)
    return wide ? launch(emd_kernel<STM_EMD_T_BIG, true, true>, STM_EMD_T_BIG) : launch(emd_kernel<STM_EMD_T_BIG, false, true>, STM_EMD_T_BIG);
  return wide ? launch(emd_kernel<256, true, false>, 256) : launch(emd_kernel<256, false, false>, 256);
}
}  // namespace
}  // namespace stm

extern "C" int stm_emd_spots_batched(const int32_t* src_x, const int32_t* src_y, const double* src_mass, int32_t n_src,
                                     const int64_t* tgt_ptr, const int32_t* tgt_x, const int32_t* tgt_y,
                                     const double* tgt_mass, int32_t G, int64_t max_tgt, const int32_t* order,
                                     int64_t max_iter, int32_t block_size,
                                     void* workspace, int64_t workspace_bytes, int32_t n_ctas,
                                     double* out_cost, int32_t* out_status, int64_t* out_iters, void* stream) {
  return stm::emd_launch(false, src_x, src_y, src_mass, n_src, tgt_ptr, tgt_x, tgt_y, tgt_mass, G, max_tgt, order, max_iter,
                         block_size, workspace, workspace_bytes, n_ctas, out_cost, out_status, out_iters, stream);
}

extern "C" int stm_emd_spots_batched_wide(const int32_t* src_x, const int32_t* src_y, const double* src_mass, int32_t n_src,
                                          const int64_t* tgt_ptr, const int32_t* tgt_x, const int32_t* tgt_y,
                                          const double* tgt_mass, int32_t G, int64_t max_tgt, const int32_t* order,
                                          int64_t max_iter, int32_t block_size,
                                          void* workspace, int64_t workspace_bytes, int32_t n_ctas,
                                          double* out_cost, int32_t* out_status, int64_t* out_iters, void* stream) {
  return stm::emd_launch(true, src_x, src_y, src_mass, n_src, tgt_ptr, tgt_x, tgt_y, tgt_mass, G, max_tgt, order, max_iter,
                         block_size, workspace, workspace_bytes, n_ctas, out_cost, out_status, out_iters, stream);
}
