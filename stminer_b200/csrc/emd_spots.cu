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
//     reduced cost of the first block that has one (LEMON's rule).  On top of it (default, STM_EMD_NO_CANDIDATES turns
//     it off): a CANDIDATE LIST — every thread keeps the best arc it saw in a successful block scan and the following
//     "minor" iterations re-price only those <= 512 arcs until none is negative — and GROWING BLOCKS — the block after
//     k fruitless ones is 2^min(k, 3) times as long.  Both belong to the pivot rule the oracle restates (cand_threads);
//   * the spanning tree is stored as parent / preorder position / subtree size (uint16) in shared memory, a
//     subtree being a contiguous segment of the preorder array — cycle paths, leaving-arc search, flow update,
//     potential shift of the re-hung subtree and preorder re-numbering are all data-parallel over the CTA;
//   * arc flows (int64) live in shared memory as far as it reaches, the rest in a per-CTA global slice;
//   * MULTISCALE START (default): the LP optimum does not depend on the start basis, the number of pivots does — by
//     a factor of ten.  Sources and targets are sorted along the Morton curve and the grid is coarsened one Morton bit
//     at a time — cells are halved along y, then along x, ... (a level's nodes are runs of the finer level's nodes; a
//     level that merges < 20 % of the nodes is skipped) —, the coarsest problem (<= 64 nodes) is
//     solved from the artificial star, and every finer level starts from the REFINEMENT of the coarser optimum:
//     two north-west-corner passes deal the flow of each coarse arc first to the children of its source cell, then
//     to the children of its target cell — a feasible forest with positive flows, completed to a strongly
//     feasible spanning tree by hanging each component's first source under the root with a zero-flow
//     artificial arc.  The finest level is the original problem, solved to optimality by the same simplex.
//     Visium-size problems (4,992 x 3,000): ~6k pivots on the original problem + ~8k cheaper ones on the coarse
//     levels instead of ~90k from the star.  STM_EMD_STAR_START selects the star (POT's start; same pivot sequence
//     as oracle/emd_oracle.c::stm_oracle_emd for the same block size).
#include <limits.h>
#include <math.h>

#include <algorithm>
#include <type_traits>

#include "stm_common.cuh"

namespace stm {
namespace {

constexpr int EMD_PREP_THREADS = 1024;
constexpr int EMD_GROW_G = 4;           // the same for the global-tree kernel (blocks of 4 .. 64 target rows)
constexpr int EMD_GROW = 3;             // a pricing block after k empty ones is 2^min(k, 3) times as long (candidate-list pricing)
constexpr int EMD_MAXPATH_S = 2048;     // longest cycle path kept, shared-memory tree
static_assert(2 * EMD_MAXPATH_S <= 0x1fff, "the leaving-arc key keeps 13 bits for the position on the cycle");
constexpr int EMD_MAXPATH_G = 8192;     // global-memory tree
constexpr int EMD_SCAN_CH_G = 8;        // global tree: node capacity 30 * 8 * threads (the workspace and the tests are sized for it)
constexpr int EMD_ROWS_G = 64;          // most target rows of a pricing block of the global-tree kernel
#ifndef STM_EMD_ROWS_MIN_G
#define STM_EMD_ROWS_MIN_G 4
#endif
constexpr int EMD_ROWS_MIN_G = STM_EMD_ROWS_MIN_G;       // and the fewest (when the target has that many)
constexpr int EMD_RG_G = 8;             // sources per thread and tile in its pricing
constexpr int EMD_GLOBAL_THREADS = 512;
constexpr int64_t EMD_GLOBAL_ALLOC_NODES = 9000;
constexpr int64_t EMD_GLOBAL_MAX_NODES = 30LL * EMD_SCAN_CH_G * EMD_GLOBAL_THREADS;   // 122,880
constexpr int EMD_LV = 24;            // BIT levels of the hierarchy kept per side (level 0 = the problem itself; level b = Morton code >> b)
constexpr int EMD_LU = 16;            // levels in use per problem at most (a level that merges < 20 % of the nodes is skipped)
constexpr int EMD_HDR_X0 = 32, EMD_HDR_Y0 = 33, EMD_HDR_OK = 34;   // src_hdr entries behind the EMD_LV node counts
constexpr int EMD_COARSEST = 64;      // the coarsest level solved has at most this many nodes (or is the last level kept / in use)
constexpr int EMD_BIAS = 16384;       // Morton keys take coordinates relative to the source's corner in [-16384, 32768)
#ifndef STM_EMD_T_BIG
#define STM_EMD_T_BIG 512    // CTA size of the one-CTA-per-SM kernel: the pricing scan is issue-bound (same time at 512 or 1,024
                             // threads) while every barrier-delimited serial phase is cheaper with 16 warps: 116 vs 111 problems/s
#endif
// sources per thread kept in registers by the whole-row pricing path: 5,120 sources per CTA for the big kernel
#ifndef STM_EMD_BLOCK_ARCS
#define STM_EMD_BLOCK_ARCS 32768   // arcs per pricing block of the one-CTA-per-SM kernel (whole target rows)
#endif
#ifndef STM_EMD_MINB_SMALL
#define STM_EMD_MINB_SMALL 1   // CTAs per SM the 256-thread kernel is compiled for (register bound)
#endif
constexpr int emd_rmax(int threads, bool big) { return big ? (5120 + threads - 1) / threads : 5; }
constexpr double EMD_SCALE = 1125899906842624.0;  // 2^50
constexpr unsigned short NONE16 = 0xffffu;
constexpr unsigned short UNVIS16 = 0xfffeu;

// One side's hierarchy in global memory: level l occupies [l * stride, l * stride + cnt[l]) of xy / mass and
// [l * (stride + 1), ...] of child; child[l][c] .. child[l][c + 1] are the children (nodes of level l - 1) of node c.
struct EmdHier {
  int32_t* xy;      // {int16 x | int16 y} cell coordinates of the level, relative to the source's corner >> l
  int64_t* mass;    // integer masses (total 2^50 on every level)
  int32_t* child;
  int32_t stride;
};

// per-CTA global scratch (element counts in nodes: capN = node capacity, capT = target capacity)
struct EmdScratch {
  int64_t* flow;      // [capN]   arc flows beyond the shared-memory part
  int64_t* tint;      // [capT]   integer target masses in input order
  EmdHier ht;         // target hierarchy of the problem being solved
  uint32_t* codeA;    // [capT] Morton codes (ping)
  uint32_t* codeB;    // [capT] (pong)
  int32_t* tcnt;      // [EMD_LV]
  // refinement
  int32_t* aI; int32_t* aJ; int64_t* aF;        // [capN] coarse optimum: real arcs with positive flow
  int32_t* offS; int32_t* curS; int32_t* lstS;  // [capN + 1], [capN + 1], [capN] arcs of every coarse source, by target
  int32_t* offT; int32_t* curT; int32_t* lstT;  // same per coarse target (arc ranks, ascending)
  int32_t* pcFirst;                             // [capN + 1] first piece of the arc of rank r
  int32_t* pI; int64_t* pF;                     // [2 capN] pieces (fine source, flow)
  int32_t* aoff;                                // [capN + 1] first fine arc of every coarse target
  int32_t* fI; int32_t* fJ; int64_t* fF;        // [capN] fine forest
  int32_t* tfirst;                              // [capT + 1] first fine arc of every fine target
  uint32_t* gadj;                               // [4 capN + 16] fallback home of the tree-build adjacency
  unsigned long long* sortbuf;                  // global tree only: [pow2(capT)] Morton sort keys of the targets
  char* gtree;                                  // global tree only: potentials, coordinates, five index arrays (32 B per node)
};

__host__ __device__ inline size_t emd_align16(size_t x) { return (x + 15) & ~(size_t)15; }

// Carves a CTA's scratch out of `base` (nullptr: only the size is computed).
__host__ __device__ inline size_t emd_scratch_layout(char* base, int64_t capN, int64_t capT, EmdScratch* s, bool global_tree) {
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += emd_align16(bytes); return p; };
  EmdScratch t;
  t.flow = (int64_t*)take(8 * capN);
  t.tint = (int64_t*)take(8 * capT);
  t.ht.mass = (int64_t*)take(8 * (size_t)EMD_LV * capT);
  t.ht.xy = (int32_t*)take(4 * (size_t)EMD_LV * capT);
  t.ht.child = (int32_t*)take(4 * (size_t)EMD_LV * (capT + 1));
  t.ht.stride = (int32_t)capT;
  t.codeA = (uint32_t*)take(4 * capT);
  t.codeB = (uint32_t*)take(4 * capT);
  t.tcnt = (int32_t*)take(4 * EMD_LV);
  t.aI = (int32_t*)take(4 * capN); t.aJ = (int32_t*)take(4 * capN); t.aF = (int64_t*)take(8 * capN);
  t.offS = (int32_t*)take(4 * (capN + 1)); t.curS = (int32_t*)take(4 * (capN + 1)); t.lstS = (int32_t*)take(4 * capN);
  t.offT = (int32_t*)take(4 * (capN + 1)); t.curT = (int32_t*)take(4 * (capN + 1)); t.lstT = (int32_t*)take(4 * capN);
  t.pcFirst = (int32_t*)take(4 * (capN + 1));
  t.pI = (int32_t*)take(4 * 2 * capN); t.pF = (int64_t*)take(8 * 2 * capN);
  t.aoff = (int32_t*)take(4 * (capN + 1));
  t.fI = (int32_t*)take(4 * capN); t.fJ = (int32_t*)take(4 * capN); t.fF = (int64_t*)take(8 * capN);
  t.tfirst = (int32_t*)take(4 * (capT + 1));
  t.gadj = (uint32_t*)take(4 * (4 * capN + 16));
  t.sortbuf = nullptr; t.gtree = nullptr;
  if (global_tree) {
    int64_t P = 1;
    while (P < capT) P <<= 1;
    t.sortbuf = (unsigned long long*)take(8 * P);
    t.gtree = take((size_t)capN * 32);
  }
  if (s) *s = t;
  return off;
}

struct EmdArgs {
  const int32_t* src_x; const int32_t* src_y; const double* src_mass; int32_t n_src;
  const int64_t* tgt_ptr; const int32_t* tgt_x; const int32_t* tgt_y; const double* tgt_mass;
  int32_t G; int32_t cap_nodes;   // shared-memory capacity in nodes
  int32_t cap_tgt;                // largest target support of the batch
  int32_t flow_smem_nodes;        // arc flows of nodes below this index live in shared memory, the rest in the scratch
  int64_t max_iter; int32_t block_size; uint32_t flags;
  // source side, built once per batch by emd_prep_kernel
  int64_t* src_int;               // [n_src] integer source masses in input order (star start)
  int32_t* src_hdr;               // [0, EMD_LV) node counts per bit level, [32] x0, [33] y0, [34] 1 = coordinates in range
  EmdHier hs;
  char* cta_ws; int64_t cta_ws_bytes;   // per-CTA scratch slices
  unsigned int* counter;          // dynamic problem queue
  const int32_t* order;           // optional processing order
  double* out_cost; int32_t* out_status; int64_t* out_iters;
};

// -DSTM_EMD_CHECK: bounds checks on every index the refinement / tree build computes (compute-sanitizer is not available on
// the pool): a violation prints its site and traps.
#ifdef STM_EMD_CHECK
#define EMD_CHECK(cond, what)                                                                              \
  do {                                                                                                     \
    if (!(cond)) { printf("[emd check] %s failed at line %d (block %d thread %d)\n", what, __LINE__, blockIdx.x, threadIdx.x); __trap(); } \
  } while (0)
#else
#define EMD_CHECK(cond, what)
#endif

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

__device__ __forceinline__ uint32_t morton16(uint32_t x, uint32_t y) {   // x, y < 2^16: x in the odd bits, y in the even bits
  uint32_t a = x, b = y;
  a = (a | (a << 8)) & 0x00ff00ffu; a = (a | (a << 4)) & 0x0f0f0f0fu; a = (a | (a << 2)) & 0x33333333u; a = (a | (a << 1)) & 0x55555555u;
  b = (b | (b << 8)) & 0x00ff00ffu; b = (b | (b << 4)) & 0x0f0f0f0fu; b = (b | (b << 2)) & 0x33333333u; b = (b | (b << 1)) & 0x55555555u;
  return (a << 1) | b;
}

// Bitonic sort of P (a power of two) 64-bit keys by the whole CTA; `a` may be shared or global.
template <int T>
__device__ void cta_bitonic_sort(unsigned long long* a, int P) {
  for (int k = 2; k <= P; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (P >> 1); t += T) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int l = i | j;
        const bool up = (i & k) == 0;
        const unsigned long long x = a[i], y = a[l];
        if ((x > y) == up) { a[i] = y; a[l] = x; }
      }
      __syncthreads();
    }
}

// Exclusive block scan of one int per thread; every warp recomputes the cross-warp level.  s_ws: [T / 32].
template <int T>
__device__ __forceinline__ int cta_excl_scan(int v, int* s_ws, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL_MASK, incl, o); if (lane >= o) incl += t; }
  __syncthreads();                       // s_ws may still be read by the previous call
  if (lane == 31) s_ws[warp] = incl;
  __syncthreads();
  const int w = lane < T / 32 ? s_ws[lane] : 0;
  int wi = w;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL_MASK, wi, o); if (lane >= o) wi += t; }
  if (total) *total = __shfl_sync(FULL_MASK, wi, 31);
  return __shfl_sync(FULL_MASK, wi - w, warp) + incl - v;
}

// a[0 .. len) -> inclusive prefix sums, in place (a[0] = 0 and the count of item k in a[k + 1] gives offsets).
template <int T>
__device__ void cta_prefix_inplace(int32_t* a, int len, int* s_ws) {
  const int per = (len + T - 1) / T;
  const int b = min(len, (int)threadIdx.x * per), e = min(len, b + per);
  int sum = 0;
  for (int i = b; i < e; ++i) sum += a[i];
  int run = cta_excl_scan<T>(sum, s_ws, nullptr);
  for (int i = b; i < e; ++i) { run += a[i]; a[i] = run; }
  __syncthreads();
}

// Bit levels 1 .. EMD_LV - 1 of one side from its level 0 (cnt0 nodes in Morton order, codes in codeA): level b = runs of
// level b - 1 nodes whose codes agree after one more bit is dropped — cells are halved along y first (the code holds y in its
// even bits), then along x; xy of level b = cell coordinates (u_x >> b/2, u_y >> (b+1)/2).  cnt[] receives the node counts.
// Whole CTA.
template <int T>
__device__ void build_levels(const EmdHier h, int cnt0, uint32_t* codeA, uint32_t* codeB, int32_t* cnt, int* s_ws) {
  if (threadIdx.x == 0) cnt[0] = cnt0;
  int pn = cnt0;
  uint32_t* cin = codeA;
  uint32_t* cout = codeB;
  for (int l = 1; l < EMD_LV; ++l) {
    const int32_t* pxy = h.xy + (size_t)(l - 1) * h.stride;
    const int64_t* pm = h.mass + (size_t)(l - 1) * h.stride;
    int32_t* cxy = h.xy + (size_t)l * h.stride;
    int64_t* cm = h.mass + (size_t)l * h.stride;
    int32_t* cch = h.child + (size_t)l * (h.stride + 1);
    const int per = (pn + T - 1) / T;
    const int b = min(pn, (int)threadIdx.x * per), e = min(pn, b + per);
    int heads = 0;
    for (int i = b; i < e; ++i) heads += (i == 0 || (cin[i] >> 1) != (cin[i - 1] >> 1)) ? 1 : 0;
    int total = 0;
    int c = cta_excl_scan<T>(heads, s_ws, &total);
    for (int i = b; i < e; ++i) {
      if (i == 0 || (cin[i] >> 1) != (cin[i - 1] >> 1)) {
        const int xy = pxy[i];
        const int x = ((xy << 16) >> 16) >> ((l & 1) ^ 1), y = (xy >> 16) >> (l & 1);   // arithmetic shifts: cell coordinates
        cxy[c] = (x & 0xffff) | (y << 16);
        cout[c] = cin[i] >> 1;
        cch[c] = i;
        ++c;
      }
    }
    if (threadIdx.x == 0) { cch[total] = pn; cnt[l] = total; }
    __syncthreads();
    for (int k = threadIdx.x; k < total; k += T) {
      long long s = 0;
      for (int i = cch[k]; i < cch[k + 1]; ++i) s += pm[i];
      cm[k] = s;
    }
    __syncthreads();
    pn = total;
    uint32_t* t = cin; cin = cout; cout = t;
  }
}

// Source side of a batch: integer masses, Morton order, hierarchy.  One CTA.
__global__ void __launch_bounds__(EMD_PREP_THREADS) emd_prep_kernel(const int32_t* src_x, const int32_t* src_y, const double* src_mass,
                                                                     int n, int64_t* src_int, unsigned long long* sortbuf, int P,
                                                                     uint32_t* codeA, uint32_t* codeB, EmdHier hs, int32_t* hdr) {
  constexpr int T = EMD_PREP_THREADS;
  __shared__ double s_scr[T / 32];
  __shared__ long long s_ll[3 * T / 32];
  __shared__ int s_ws[T / 32];
  __shared__ int s_box[4];
  const int tid = threadIdx.x;
  integerize<T>(src_mass, n, src_int, s_scr, s_ll);
  if (tid == 0) { s_box[0] = INT_MAX; s_box[1] = INT_MIN; s_box[2] = INT_MAX; s_box[3] = INT_MIN; }
  __syncthreads();
  int x0 = INT_MAX, x1 = INT_MIN, y0 = INT_MAX, y1 = INT_MIN;
  for (int i = tid; i < n; i += T) { const int x = src_x[i], y = src_y[i]; x0 = min(x0, x); x1 = max(x1, x); y0 = min(y0, y); y1 = max(y1, y); }
  x0 = __reduce_min_sync(FULL_MASK, x0); x1 = __reduce_max_sync(FULL_MASK, x1);
  y0 = __reduce_min_sync(FULL_MASK, y0); y1 = __reduce_max_sync(FULL_MASK, y1);
  if ((tid & 31) == 0) { atomicMin(&s_box[0], x0); atomicMax(&s_box[1], x1); atomicMin(&s_box[2], y0); atomicMax(&s_box[3], y1); }
  __syncthreads();
  x0 = s_box[0]; x1 = s_box[1]; y0 = s_box[2]; y1 = s_box[3];
  const bool ok = (long long)x1 - x0 < 32768 && (long long)y1 - y0 < 32768;
  if (tid == 0) { hdr[EMD_HDR_X0] = x0; hdr[EMD_HDR_Y0] = y0; hdr[EMD_HDR_OK] = ok ? 1 : 0; }
  if (!ok) { if (tid < EMD_LV) hdr[tid] = 0; return; }
  for (int i = tid; i < P; i += T)
    sortbuf[i] = i < n ? ((unsigned long long)morton16((uint32_t)(src_x[i] - x0 + EMD_BIAS), (uint32_t)(src_y[i] - y0 + EMD_BIAS)) << 32) | (unsigned)i
                       : ~0ull;
  __syncthreads();
  cta_bitonic_sort<T>(sortbuf, P);
  for (int i = tid; i < n; i += T) {
    const unsigned long long k = sortbuf[i];
    const int o = (int)(k & 0xffffffffu);
    codeA[i] = (uint32_t)(k >> 32);
    hs.xy[i] = ((src_x[o] - x0) & 0xffff) | ((src_y[o] - y0) << 16);
    hs.mass[i] = src_int[o];
  }
  __syncthreads();
  build_levels<T>(hs, n, codeA, codeB, hdr, s_ws);
}

// WIDE = 64-bit potentials and reduced costs for problems whose coordinate span makes (span^2 + 1) * nodes exceed the
// int32 range of the default kernel (scattered integer coordinates, long thin tissues): 22 instead of 18 bytes of shared
// memory per node and the general (non register-resident) pricing path, same pivot sequence.
//
// GLOBAL = the tree (potentials, coordinates, parent / position / size / preorder arrays, 32-bit node indices) lives in the
// CTA's global scratch (L2) instead of shared memory: problems beyond one CTA's shared memory — Stereo-seq bin50 slides,
// 40,000 source bins — up to EMD_GLOBAL_MAX_NODES nodes.  Shared memory then holds arc flows and the cycle paths; pricing
// keeps a tile of sources in registers and meets the block's target rows staged in shared memory.  Same algorithm, same
// pivot rules; implies WIDE.
template <int EMD_THREADS, bool WIDE, bool ONE_PER_SM, bool GLOBAL = false>
__global__ void __launch_bounds__(EMD_THREADS, ONE_PER_SM ? 1 : STM_EMD_MINB_SMALL) emd_kernel(const EmdArgs p) {
  static_assert(!GLOBAL || (WIDE && ONE_PER_SM), "the global-tree variant uses 64-bit potentials and one CTA per SM");
  constexpr int EMD_WARPS = EMD_THREADS / 32;
  constexpr int EMD_RMAX_T = emd_rmax(EMD_THREADS, ONE_PER_SM);
  constexpr int EMD_MAXPATH = GLOBAL ? EMD_MAXPATH_G : EMD_MAXPATH_S;
  constexpr int SCAN_CH = 1;        // 32-position chunks of a thread's run in the cycle-path scan (shared-memory kernels)
  using pi_t = typename std::conditional<WIDE, long long, int>::type;
  using idx_t = typename std::conditional<GLOBAL, uint32_t, unsigned short>::type;
  constexpr idx_t NONE16 = (idx_t)~(idx_t)0, UNVIS16 = (idx_t)(NONE16 - 1);
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
  char* const cta_ws = p.cta_ws + (size_t)blockIdx.x * p.cta_ws_bytes;
  pi_t* s_pi; int* s_xy; idx_t *s_parent, *s_pos, *s_size, *s_order, *s_tmp, *s_pathA;
  if constexpr (GLOBAL) {
    EmdScratch sct; emd_scratch_layout(cta_ws, cap, p.cap_tgt, &sct, true);
    s_pi = reinterpret_cast<pi_t*>(sct.gtree);                           // [cap] int64
    s_xy = reinterpret_cast<int*>(s_pi + cap);                           // [cap]
    s_parent = reinterpret_cast<idx_t*>(s_xy + cap);
    s_pathA = reinterpret_cast<idx_t*>(s_flow + fcap);                   // paths stay in shared memory
  } else {
    s_pi = reinterpret_cast<pi_t*>(s_flow + fcap);                       // [cap]
    s_xy = reinterpret_cast<int*>(s_pi + cap);                           // [cap] {int16 x | int16 y}
    s_parent = reinterpret_cast<idx_t*>(s_xy + cap);
  }
  s_pos = s_parent + cap;
  s_size = s_pos + cap;
  s_order = s_size + cap;
  s_tmp = s_order + cap;
  if constexpr (!GLOBAL) s_pathA = s_tmp + cap;                          // [EMD_MAXPATH]
  idx_t* s_pathB = s_pathA + EMD_MAXPATH;
  idx_t* s_P = s_pathB + EMD_MAXPATH;
  idx_t* s_S = s_P + EMD_MAXPATH;
  __shared__ double s_scr[EMD_WARPS];
  __shared__ long long s_ll[3 * EMD_WARPS];
  __shared__ int2 s_wkey[2 * EMD_WARPS];   // per-warp (reduced cost, position) of the pricing arg-min, double-buffered
  __shared__ int s_i[16];
  __shared__ int s_wsum[EMD_WARPS];
  __shared__ int s_ws[EMD_WARPS];          // block scans of the hierarchy / refinement code
  __shared__ int s_lv[EMD_LU];             // levels in use, finest first: bit level << 20 | node count of the target side
  __shared__ long long s_wrc64[WIDE ? 2 * EMD_WARPS : 1];   // per-warp reduced costs of the WIDE pricing arg-min
  __shared__ int s_row[GLOBAL ? 2 * EMD_ROWS_G : 1];          // the block's target rows: -2x, -2y
  __shared__ long long s_rowp[GLOBAL ? EMD_ROWS_G : 1];       // and potentials (global-tree pricing)
  __shared__ int s_cand[(WIDE || GLOBAL) ? 1 : EMD_THREADS];  // candidate list: every thread's best arc (j << 16 | i) of the last pass
  __shared__ int2 s_candg[GLOBAL ? EMD_THREADS : 1];          // the same for the global-tree kernel: (source, target), source < 0: none

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // the scratch pointers are recomputed where they are needed (hierarchy, tree build, refinement) so that they do not
  // occupy registers across the pivot loop
#define EMD_SCRATCH(sc) EmdScratch sc; emd_scratch_layout(cta_ws, cap, p.cap_tgt, &sc, GLOBAL)
  int64_t* flow;
  { EMD_SCRATCH(sc0); flow = sc0.flow; }
  auto FL = [&](int v) -> long long& { return v < fcap ? s_flow[v] : *reinterpret_cast<long long*>(flow + v); };
  const bool hier_ok = p.src_hdr[EMD_HDR_OK] != 0;
  // CTA-wide lexicographic minimum of (64-bit value, 32-bit key) by shuffles: over the warp, then — through one shared entry
  // per warp, double-buffered by `round`, one barrier — over the per-warp results; every thread gets the result
  auto cta_lexmin64 = [&](long long& r64, unsigned& k32, unsigned& round) {
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
    ++round;
  };

  for (;;) {
    __syncthreads();
    if (tid == 0) s_i[0] = (int)atomicAdd(p.counter, 1u);
    __syncthreads();
    const int slot = s_i[0];
    if (slot >= p.G) break;
    const int g = p.order ? p.order[slot] : slot;
    const int64_t t0 = p.tgt_ptr[g];
    const int m0 = (int)(p.tgt_ptr[g + 1] - t0);
    const int n0 = p.n_src;
    if (m0 <= 0 || n0 <= 0) {
      if (tid == 0) { p.out_cost[g] = 0.0; p.out_status[g] = 2; if (p.out_iters) p.out_iters[g] = 0; }
      continue;
    }
    // ---- bounding box (relative to the source's corner), capacity checks ------------------------------------
    const int ox = hier_ok ? p.src_hdr[EMD_HDR_X0] : p.src_x[0], oy = hier_ok ? p.src_hdr[EMD_HDR_Y0] : p.src_y[0];
    int xmin = INT_MAX, xmax = INT_MIN, ymin = INT_MAX, ymax = INT_MIN;
    int txmin = INT_MAX, tymin = INT_MAX, txmax = INT_MIN, tymax = INT_MIN;
    for (int v = tid; v < n0 + m0; v += EMD_THREADS) {
      const int x = v < n0 ? p.src_x[v] : p.tgt_x[t0 + v - n0], y = v < n0 ? p.src_y[v] : p.tgt_y[t0 + v - n0];
      xmin = min(xmin, x); xmax = max(xmax, x); ymin = min(ymin, y); ymax = max(ymax, y);
      if (v >= n0) { txmin = min(txmin, x); tymin = min(tymin, y); txmax = max(txmax, x); tymax = max(tymax, y); }
    }
    xmin = __reduce_min_sync(FULL_MASK, xmin); xmax = __reduce_max_sync(FULL_MASK, xmax);
    ymin = __reduce_min_sync(FULL_MASK, ymin); ymax = __reduce_max_sync(FULL_MASK, ymax);
    txmin = __reduce_min_sync(FULL_MASK, txmin); tymin = __reduce_min_sync(FULL_MASK, tymin);
    txmax = __reduce_max_sync(FULL_MASK, txmax); tymax = __reduce_max_sync(FULL_MASK, tymax);
    __syncthreads();
    if (lane == 0) {
      s_ll[3 * warp] = ((long long)xmin << 32) | (unsigned)xmax; s_ll[3 * warp + 1] = ((long long)ymin << 32) | (unsigned)ymax;
      s_wkey[warp] = make_int2(txmin, tymin); s_wkey[EMD_WARPS + warp] = make_int2(txmax, tymax);
    }
    __syncthreads();
    for (int w = 0; w < EMD_WARPS; ++w) {
      xmin = min(xmin, (int)(s_ll[3 * w] >> 32)); xmax = max(xmax, (int)(unsigned)s_ll[3 * w]);
      ymin = min(ymin, (int)(s_ll[3 * w + 1] >> 32)); ymax = max(ymax, (int)(unsigned)s_ll[3 * w + 1]);
      txmin = min(txmin, s_wkey[w].x); tymin = min(tymin, s_wkey[w].y);
      txmax = max(txmax, s_wkey[EMD_WARPS + w].x); tymax = max(tymax, s_wkey[EMD_WARPS + w].y);
    }
    __syncthreads();
    const int N0 = n0 + m0 + 1;
    // N + 1 <= cap: the tree build keeps N + 1 offsets in one node array; N <= 30 * threads: the cycle-path scan keeps one
    // bit per preorder position of a thread's run in a 32-bit mask
    const bool fits = N0 + 1 <= cap && (GLOBAL || N0 < 65534) && (GLOBAL ? (int64_t)N0 <= EMD_GLOBAL_MAX_NODES : N0 <= 30 * SCAN_CH * EMD_THREADS) && (int64_t)n0 * m0 < 2147483647LL;
    const long long maxc0 = (long long)(xmax - xmin) * (xmax - xmin) + (long long)(ymax - ymin) * (ymax - ymin);
    const bool coords_ok = (xmax - ox) < 32768 && (ox - xmin) < 32768 && (ymax - oy) < 32768 && (oy - ymin) < 32768;
    const long long pi_limit = WIDE ? (1LL << 61) : 2147483647LL;
    if (!fits || !coords_ok || 2 * (maxc0 + 1) * (long long)N0 + maxc0 >= pi_limit) {   // potentials / int16 coordinates / capacity
      if (tid == 0) { p.out_cost[g] = nan(""); p.out_status[g] = 3; if (p.out_iters) p.out_iters[g] = 0; }
      continue;
    }
    // multiscale start unless asked otherwise (or the Morton keys cannot hold the coordinates)
    bool star = (p.flags & STM_EMD_STAR_START) != 0 || !hier_ok || (ox - txmin) > EMD_BIAS || (oy - tymin) > EMD_BIAS;

    long long it_total = 0;
    int status = 0;
    int n = n0, m = m0, N = N0, root = n0 + m0;
    for (int attempt = 0; attempt < 2; ++attempt) {
    it_total = 0; status = 0;
    int L = 1;
#ifdef STM_EMD_PROFILE
    const long long c_h0 = clock64();
#endif
    if (!star) {
      EMD_SCRATCH(sc);
      // ---- target hierarchy: integer masses, Morton order (sorted in the still unused tree memory), levels ----
      integerize<EMD_THREADS>(p.tgt_mass + t0, m0, sc.tint, s_scr, s_ll);
      unsigned long long* keys = GLOBAL ? sc.sortbuf : reinterpret_cast<unsigned long long*>(smem_raw);
      int P = 1;
      while (P < m0) P <<= 1;
      for (int i = tid; i < P; i += EMD_THREADS)
        keys[i] = i < m0 ? ((unsigned long long)morton16((uint32_t)(p.tgt_x[t0 + i] - ox + EMD_BIAS), (uint32_t)(p.tgt_y[t0 + i] - oy + EMD_BIAS)) << 32) | (unsigned)i
                         : ~0ull;
      __syncthreads();
      cta_bitonic_sort<EMD_THREADS>(keys, P);
      for (int i = tid; i < m0; i += EMD_THREADS) {
        const unsigned long long k = keys[i];
        const int o = (int)(k & 0xffffffffu);
        sc.codeA[i] = (uint32_t)(k >> 32);
        sc.ht.xy[i] = ((p.tgt_x[t0 + o] - ox) & 0xffff) | ((p.tgt_y[t0 + o] - oy) << 16);
        sc.ht.mass[i] = sc.tint[o];
      }
      __syncthreads();
      build_levels<EMD_THREADS>(sc.ht, m0, sc.codeA, sc.codeB, sc.tcnt, s_ws);
      // the bit levels in use: coarsen until at most EMD_COARSEST nodes are left; a level that keeps more than 4/5 of the
      // nodes of the one below it is skipped (on the Visium lattice halving along y merges nothing), never two in a row
      if (tid == 0) {
        int b = 0, Lc = 1;
        s_lv[0] = sc.tcnt[0];
        while (p.src_hdr[b] + sc.tcnt[b] > EMD_COARSEST && b + 1 < EMD_LV && Lc < EMD_LU) {
          int nb = b + 1;
          if (nb + 1 < EMD_LV && 5LL * (p.src_hdr[nb] + sc.tcnt[nb]) > 4LL * (p.src_hdr[b] + sc.tcnt[b])) nb = b + 2;
          s_lv[Lc++] = (nb << 20) | sc.tcnt[nb];
          b = nb;
        }
        s_i[10] = Lc;
      }
      __syncthreads();
      L = s_i[10];
    }

#ifdef STM_EMD_PROFILE
    if (tid == 0 && g < 2) printf("[emd prof] g=%d hierarchy %lld cycles, %d levels\n", g, (long long)(clock64() - c_h0), L);
#endif
    for (int li = L - 1; li >= 0; --li) {
#ifdef STM_EMD_PROFILE
    const long long c_l0 = clock64();
#endif
    const int lev = star ? 0 : s_lv[li] >> 20;          // bit level
    n = star ? n0 : p.src_hdr[lev];
    m = star ? m0 : s_lv[li] & 0xfffff;
    N = n + m + 1; root = n + m;
    const bool from_forest = !star && li < L - 1;
    const int Ef = from_forest ? s_i[8] : 0;
    // ---- stage the level's coordinates, its bounding box -------------------------------------------------------
    {
      EMD_SCRATCH(sc);
      int bx0 = INT_MAX, bx1 = INT_MIN, by0 = INT_MAX, by1 = INT_MIN;
      const int32_t* hsx = p.hs.xy + (size_t)lev * p.hs.stride;
      const int32_t* htx = sc.ht.xy + (size_t)lev * sc.ht.stride;
      for (int v = tid; v < n + m; v += EMD_THREADS) {
        int xy;
        if (star) {
          const int x = v < n ? p.src_x[v] : p.tgt_x[t0 + v - n], y = v < n ? p.src_y[v] : p.tgt_y[t0 + v - n];
          xy = ((x - ox) & 0xffff) | ((y - oy) << 16);
        } else {
          xy = v < n ? hsx[v] : htx[v - n];
          // an odd level's cells are twice as high as wide: y in units of the cell width keeps the cost isotropic
          if (lev & 1) xy = (xy & 0xffff) | (((xy >> 16) * 2) << 16);
        }
        s_xy[v] = xy;
        const int x = (xy << 16) >> 16, y = xy >> 16;
        bx0 = min(bx0, x); bx1 = max(bx1, x); by0 = min(by0, y); by1 = max(by1, y);
      }
      bx0 = __reduce_min_sync(FULL_MASK, bx0); bx1 = __reduce_max_sync(FULL_MASK, bx1);
      by0 = __reduce_min_sync(FULL_MASK, by0); by1 = __reduce_max_sync(FULL_MASK, by1);
      __syncthreads();
      if (lane == 0) { s_wkey[warp] = make_int2(bx0, by0); s_wkey[EMD_WARPS + warp] = make_int2(bx1, by1); }
      __syncthreads();
      for (int w = 0; w < EMD_WARPS; ++w) {
        bx0 = min(bx0, s_wkey[w].x); by0 = min(by0, s_wkey[w].y);
        bx1 = max(bx1, s_wkey[EMD_WARPS + w].x); by1 = max(by1, s_wkey[EMD_WARPS + w].y);
      }
      __syncthreads();
      xmin = bx0; xmax = bx1; ymin = by0; ymax = by1;
    }
    const long long maxc = (long long)(xmax - xmin) * (xmax - xmin) + (long long)(ymax - ymin) * (ymax - ymin);
    const long long BIG = (maxc + 1) * (long long)N;
    bool forest_done = false;
    if (from_forest) {
      EMD_SCRATCH(sc);
      // ---- start basis from the refined forest (sc.fI / fJ / fF, Ef arcs, grouped by target in sc.tfirst) ----------
      // adjacency for the traversal: sources CSR (soff16 / sadj16), targets = their slice of the arc list (ai16)
      idx_t* soff16 = s_tmp;              // [n + 1]
      idx_t* tf16 = s_tmp + n + 1;        // [m + 1]
      const bool adj_smem = (size_t)fcap * 8 >= (size_t)4 * (n + 1) + 2 * sizeof(idx_t) * (size_t)Ef + 16;
      int* deg32 = adj_smem ? reinterpret_cast<int*>(s_flow) : sc.curS;
      idx_t* sadj16 = adj_smem ? reinterpret_cast<idx_t*>(deg32 + n + 1) : reinterpret_cast<idx_t*>(sc.gadj);
      idx_t* ai16 = sadj16 + Ef;
      for (int v = tid; v <= n; v += EMD_THREADS) deg32[v] = 0;
      __syncthreads();
      for (int e = tid; e < Ef; e += EMD_THREADS) { const int i = sc.fI[e]; atomicAdd(&deg32[i + 1], 1); ai16[e] = (idx_t)i; }
      __syncthreads();
      cta_prefix_inplace<EMD_THREADS>(deg32, n + 1, s_ws);
      for (int v = tid; v <= n; v += EMD_THREADS) soff16[v] = (idx_t)deg32[v];
      for (int j = tid; j <= m; j += EMD_THREADS) tf16[j] = (idx_t)sc.tfirst[j];
      __syncthreads();
      for (int e = tid; e < Ef; e += EMD_THREADS) {
        const int at = atomicAdd(&deg32[sc.fI[e]], 1);
        EMD_CHECK(at >= 0 && at < Ef && sc.fI[e] < n && sc.fJ[e] < m, "source adjacency slot");
        sadj16[at] = (idx_t)sc.fJ[e];
      }
      for (int v = tid; v < N; v += EMD_THREADS) s_parent[v] = UNVIS16;
      __syncthreads();
      for (int v = tid; v < n; v += EMD_THREADS) {      // lists in increasing target order (deterministic layout)
        const int a = soff16[v], b = soff16[v + 1];
        for (int x = a + 1; x < b; ++x) {
          const idx_t key = sadj16[x];
          int y = x - 1;
          while (y >= a && sadj16[y] > key) { sadj16[y + 1] = sadj16[y]; --y; }
          sadj16[y + 1] = key;
        }
      }
      __syncthreads();
      if (tid == 0) {
        // depth-first traversal from every component's lowest source: parents, preorder, potentials
        idx_t* stk = s_size;
        int cnt = 0, bad = 0;
        s_order[0] = (idx_t)root; s_pos[root] = 0; s_parent[root] = NONE16; s_pi[root] = 0; cnt = 1;
        for (int s0 = 0; s0 < n; ++s0) {
          if (s_parent[s0] != UNVIS16) continue;
          s_parent[s0] = (idx_t)root;
          { const int xy = s_xy[s0], xr = (xy << 16) >> 16, yr = xy >> 16; s_pi[s0] = (pi_t)(-BIG + (xr * xr + yr * yr)); }
          int sp = 0;
          stk[sp++] = (idx_t)s0;
          while (sp > 0) {
            const int v = stk[--sp];
            EMD_CHECK(cnt < N && v < N, "preorder position");
            s_order[cnt] = (idx_t)v; s_pos[v] = (idx_t)cnt; ++cnt;
            const int pv = s_parent[v];
            const int xy = s_xy[v], xv = (xy << 16) >> 16, yv = xy >> 16;
            if (v < n) {
              const pi_t piv = s_pi[v] - (pi_t)(xv * xv + yv * yv);
              for (int k = (int)soff16[v + 1] - 1; k >= (int)soff16[v]; --k) {
                const int w = n + sadj16[k];
                if (w == pv) continue;
                if (s_parent[w] != UNVIS16) { bad = 1; continue; }
                s_parent[w] = (idx_t)v;
                const int wxy = s_xy[w], xw = (wxy << 16) >> 16, yw = wxy >> 16;
                const int dx = xv - xw, dy = yv - yw;
                s_pi[w] = piv + (pi_t)(dx * dx + dy * dy) - (pi_t)(xw * xw + yw * yw);
                stk[sp++] = (idx_t)w;
              }
            } else {
              const pi_t piv = s_pi[v] + (pi_t)(xv * xv + yv * yv);
              const int j = v - n;
              for (int k = (int)tf16[j + 1] - 1; k >= (int)tf16[j]; --k) {
                const int w = ai16[k];
                if (w == pv) continue;
                if (s_parent[w] != UNVIS16) { bad = 1; continue; }
                s_parent[w] = (idx_t)v;
                const int wxy = s_xy[w], xw = (wxy << 16) >> 16, yw = wxy >> 16;
                const int dx = xv - xw, dy = yv - yw;
                s_pi[w] = piv - (pi_t)(dx * dx + dy * dy) + (pi_t)(xw * xw + yw * yw);
                stk[sp++] = (idx_t)w;
              }
            }
          }
        }
        for (int v = n; v < root; ++v)       // targets without any arc (zero integer demand): under the root, as in the star
          if (s_parent[v] == UNVIS16) {
            s_parent[v] = (idx_t)root;
            const int xy = s_xy[v], xr = (xy << 16) >> 16, yr = xy >> 16;
            s_pi[v] = (pi_t)(BIG - (xr * xr + yr * yr));
            s_order[cnt] = (idx_t)v; s_pos[v] = (idx_t)cnt; ++cnt;
          }
        s_i[9] = (bad == 0 && cnt == N) ? 1 : 0;
      }
      __syncthreads();
      forest_done = s_i[9] != 0;
      if (forest_done) {
        // flows of the tree arcs (the arc of (i, j) sits in j's slice of the arc list), subtree sizes
        for (int v = tid; v < n + m; v += EMD_THREADS) {
          const int pr = s_parent[v];
          long long f = 0;
          if (pr != root) {
            const int i = v < n ? v : pr, j = (v < n ? pr : v) - n;
            for (int e = tf16[j]; e < (int)tf16[j + 1]; ++e)
              if (sc.fI[e] == i) { f = sc.fF[e]; break; }
          }
          // the adjacency scratch shares the flow array's memory: all reads of it are done (barrier above)
          FL(v) = f;
          s_size[v] = 1;
        }
        if (tid == 0) { FL(root) = 0; s_size[root] = 1; s_i[4] = 0; }
        __syncthreads();
        if (tid == 0) {
          for (int x = N - 1; x >= 1; --x) { const int v = s_order[x]; s_size[s_parent[v]] += s_size[v]; }
        }
        __syncthreads();
      }
    }
    if (!forest_done) {
      // ---- integer masses and the initial star basis ----------------------------------------------
      if (star) {
        for (int i = tid; i < n; i += EMD_THREADS) flow[i] = p.src_int[i];
        integerize<EMD_THREADS>(p.tgt_mass + t0, m, flow + n, s_scr, s_ll);
      } else {
        EMD_SCRATCH(sc);
        const int64_t* sm = p.hs.mass + (size_t)lev * p.hs.stride;
        const int64_t* tm = sc.ht.mass + (size_t)lev * sc.ht.stride;
        for (int v = tid; v < n + m; v += EMD_THREADS) flow[v] = v < n ? sm[v] : tm[v - n];
      }
      for (int v = tid; v < n + m; v += EMD_THREADS) {
        s_parent[v] = (idx_t)root; s_order[v + 1] = (idx_t)v; s_pos[v] = (idx_t)(v + 1);
        const int xy = s_xy[v], xr = (xy << 16) >> 16, yr = xy >> 16, r2 = xr * xr + yr * yr;
        s_size[v] = 1; s_pi[v] = v < n ? (pi_t)(-BIG + r2) : (pi_t)(BIG - r2);
      }
      if (tid == 0) {
        s_order[0] = (idx_t)root; s_pos[root] = 0; s_size[root] = (idx_t)N; s_parent[root] = NONE16;
        s_pi[root] = 0; flow[root] = 0; s_i[4] = 0;
      }
      __syncthreads();
      for (int v = tid; v < min(fcap, N); v += EMD_THREADS) s_flow[v] = flow[v];
      __syncthreads();
    }

    const int n_arcs = n * m;
    // pricing block: LEMON uses sqrt(#arcs); a CTA prices in parallel, so larger blocks (fewer, better pivots and
    // fewer barriers per arc) win — 32 arcs per thread measured best on Visium-size problems
    // on the problem itself; coarse levels (few arcs in all) price a sixteenth of their arcs, at least 4,096
    int B = p.block_size > 0 ? p.block_size
                             : max((int)sqrt((double)n_arcs), min(ONE_PER_SM ? STM_EMD_BLOCK_ARCS : 32 * EMD_THREADS, max(4096, n_arcs / 16)));
    if (!GLOBAL && p.block_size <= 0 && n <= EMD_RMAX_T * EMD_THREADS) B = max(1, (B + n / 2) / n) * n;   // whole target rows
    if (B < 10) B = 10;
    if (B > n_arcs) B = n_arcs;
    if (GLOBAL && p.block_size <= 0) {         // whole rows: at least EMD_ROWS_MIN_G (the sources are read once per block)
      int rows = (B + n / 2) / n;
      rows = max(EMD_ROWS_MIN_G, min(EMD_ROWS_G, rows));
      B = min(rows, m) * n;
    }
    const bool whole_rows = !WIDE && B % n == 0 && n <= EMD_RMAX_T * EMD_THREADS;   // n_arcs and every block start are multiples of n
    const bool whole_rows_g = GLOBAL && B % n == 0 && B / n <= EMD_ROWS_G;
    int next_arc = 0;
    int next_row = 0;                               // next_arc / n while the blocks are whole rows
    const int rows_B = whole_rows ? B / n : 0;
    long long it = 0;
    unsigned round = 0;
    // Candidate list (register-resident whole-row pricing only): a pricing pass that finds a negative block leaves every
    // thread's first most-negative arc of that block in s_cand; the following "minor" iterations re-price only these
    // <= EMD_THREADS arcs (four shared loads and one barrier instead of a block scan), enter the most negative one
    // (ties: lowest thread) and return to the block search where it stopped once none of them is negative.  Same
    // optimum, about 10 % fewer pivots and a third of the block scans on the coarse levels (oracle: cand_threads).
    const bool cand_on = !WIDE && !GLOBAL && whole_rows && !(p.flags & STM_EMD_NO_CANDIDATES);
    const bool grow_g = GLOBAL && whole_rows_g && !(p.flags & STM_EMD_NO_CANDIDATES);
    bool wide_pr = false;     // global-tree kernel: 64-bit pricing for the rest of the level (32-bit wrap artefact, or the final certificate)
    const int rows_Bg = whole_rows_g ? B / n : 0;
    bool have_cand = false;
    // numItermax of the reference bounds the pivots on the ORIGINAL problem (level 0); the coarse levels only construct
    // its start basis and run to their optimum
    const long long level_max_iter = li == 0 ? p.max_iter : LLONG_MAX;
    __syncthreads();
#ifdef STM_EMD_PROFILE
    const long long c_l1 = clock64();
#endif

#ifdef STM_EMD_PROFILE
    long long c_price = 0, c_walk = 0, c_leave = 0, c_update = 0, n_scanned = 0, c_t0 = 0, sum_len = 0, sum_s = 0, sum_range = 0;
    long long c_scan = 0, c_red = 0, n_blocks = 0, c_t1 = 0, n_minor = 0;
#define PROF_T(x) { const long long _c = clock64(); x += _c - c_t0; c_t0 = _c; }
#else
#define PROF_T(x)
#endif
    for (;;) {
#ifdef STM_EMD_PROFILE
      c_t0 = clock64();
#endif
      // ---- entering arc: block search over arcs e = j*n + i ----------------------------------------
      int scanned = 0, best_e = -1, cand_j = -1;
      pi_t best_rc = 0;
      if constexpr (!WIDE && !GLOBAL) {
        if (have_cand) {
          // ---- minor iteration: re-price the candidates of the last pricing pass --------------------------------
          const int c = s_cand[tid];
          int rc = 0;
          if (c >= 0) {
            const int ci = c & 0xffff, cj = c >> 16;
            const int sxy = s_xy[ci], txy = s_xy[n + cj];
            const int nx2 = -2 * ((txy << 16) >> 16), ny2 = -2 * (txy >> 16);
            rc = ((sxy << 16) >> 16) * nx2 + ((sxy >> 16) * ny2 + (s_pi[ci] - s_pi[n + cj]));
            if (rc >= 0) { s_cand[tid] = -1; rc = 0; }            // no longer a candidate
          }
          const int w_rc = __reduce_min_sync(FULL_MASK, rc);
          const unsigned w_t = __reduce_min_sync(FULL_MASK, (rc < 0 && rc == w_rc) ? (unsigned)tid : 0xffffffffu);
          int2* wkey = s_wkey + (round & 1) * EMD_WARPS;
          if (lane == 0) wkey[warp] = make_int2(w_rc, (int)w_t);
          __syncthreads();
          const int2 e = lane < EMD_WARPS ? wkey[lane] : make_int2(INT_MAX, -1);
          const int m_rc = __reduce_min_sync(FULL_MASK, e.x);
          const unsigned m_t = __reduce_min_sync(FULL_MASK, e.x == m_rc ? (unsigned)e.y : 0xffffffffu);
          ++round;
          if (m_rc < 0) {
            const int cw = s_cand[m_t];                            // its owner kept it (only dropped entries are rewritten)
            best_rc = m_rc;
            cand_j = cw >> 16;
            best_e = cand_j * n + (cw & 0xffff);
#ifdef STM_EMD_PROFILE
            ++n_minor;
#endif
          } else {
            have_cand = false;
          }
        }
      }
      if constexpr (GLOBAL) {
        if (have_cand) {
          // ---- minor iteration of the global-tree kernel: the same rule with 64-bit reduced costs ----------------
          const int2 c = s_candg[tid];
          long long rc = 0;
          if (c.x >= 0) {
            const int sxy = s_xy[c.x], txy = s_xy[n + c.y];
            const long long nx2 = -2 * ((txy << 16) >> 16), ny2 = -2 * (txy >> 16);
            rc = ((sxy << 16) >> 16) * nx2 + ((sxy >> 16) * ny2 + ((long long)s_pi[c.x] - (long long)s_pi[n + c.y]));
            if (rc >= 0) { s_candg[tid].x = -1; rc = 0; }          // no longer a candidate
          }
          long long r64 = rc;
          unsigned t32 = rc < 0 ? (unsigned)tid : 0xffffffffu;      // ties: lowest thread
          cta_lexmin64(r64, t32, round);
          if (r64 < 0) {
            const int2 cw = s_candg[t32];                          // its owner kept it
            best_rc = (pi_t)r64;
            cand_j = cw.y;
            best_e = cw.y * n + cw.x;
#ifdef STM_EMD_PROFILE
            ++n_minor;
#endif
          } else {
            have_cand = false;
          }
        }
      }
      bool certify;
      do {
      certify = false;
      int empty = 0;                       // consecutive blocks of this pass without a negative arc
      while (best_e < 0 && scanned < n_arcs) {
        // with the candidate list, the block after k empty ones is 2^min(k, EMD_GROW) times as long: near the optimum most
        // blocks hold no negative arc, and the per-block work (sources into registers, reduction, barrier) is paid less often
        int grow = cand_on ? min(empty, EMD_GROW) : 0;
        if constexpr (GLOBAL) {
          // global-tree kernel: the sources stream from L2 once per block, so long blocks pay even more — up to
          // EMD_ROWS_G target rows (what the block's row table in shared memory holds)
          if (grow_g) { grow = min(empty, EMD_GROW_G); while (grow > 0 && (rows_Bg << grow) > EMD_ROWS_G) --grow; }
        }
        const int Bk = B << grow;
        const int cnt = min(Bk, n_arcs - scanned);
        ++empty;
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
            // Layout by level size: the CTA splits into `groups` row groups of `teff` threads; a group's thread ti owns
            // sources ti, ti + teff, ... (all EMD_RMAX_T register slots: the per-row overhead — three shared words, the
            // unpacking, the loop — is amortised over as many arcs as possible) and the groups share the block's rows
            // (rows gidx, gidx + groups, ...), so the small coarse levels of the multiscale start use all lanes too.
            int sx[EMD_RMAX_T], sy[EMD_RMAX_T], sp[EMD_RMAX_T];
            // teff = the power of two >= max(32, ceil(n / slots)): group and lane indices by shifts (the block loop runs
            // thousands of times per problem: no integer division on its path)
            const int tq = (n + EMD_RMAX_T - 1) / EMD_RMAX_T;
            const int tsh = tq <= 32 ? 5 : 32 - __clz(tq - 1);
            const int teff = 1 << tsh;
            const int groups = EMD_THREADS >> tsh;
            const int gidx = tid >> tsh;
            const int i0 = tid & (teff - 1);
#pragma unroll
            for (int r = 0; r < EMD_RMAX_T; ++r) {
              const int i = i0 + r * teff;
              const bool valid = i < n && gidx < groups;
              const int xy = valid ? s_xy[i] : 0;
              sx[r] = (xy << 16) >> 16; sy[r] = xy >> 16;
              sp[r] = valid ? s_pi[i] : INT_MAX / 2;     // never the minimum
            }
            // The scan only tracks the thread's minimum (an add, two IMADs and a share of a 3-input min per arc); the
            // arc that attains the block minimum is located afterwards by the few threads whose minimum equals it.
            const int j0 = next_row;
            const int rows = cnt == Bk ? rows_B << grow : cnt / n;   // (the division: last block of a full sweep only)
            const int rr0 = gidx < groups ? gidx : rows;       // lanes beyond the last row group sit the block out
            int j = j0 + gidx;
            while (j >= m) j -= m;
            int bjrow = 0;                                     // the target row of the thread's (first) minimum
            for (int rr = rr0; rr < rows; rr += groups) {
              const int txy = s_xy[n + j], nx2 = -2 * ((txy << 16) >> 16), ny2 = -2 * (txy >> 16), bj = s_pi[n + j];
              int rmin = INT_MAX;
#pragma unroll
              for (int r = 0; r < EMD_RMAX_T; ++r) rmin = min(rmin, sx[r] * nx2 + (sy[r] * ny2 + (sp[r] - bj)));
              if (rmin < brc) { brc = rmin; bjrow = j; }
              j += groups;
              while (j >= m) j -= m;
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
              int mycand = -1;
              if (brc < 0) {          // the thread's first row holding its minimum: the arc is in that row (slots ascending);
                j = bjrow;            // it is the thread's candidate, and the entering arc if the minimum is the block's
                const int txy = s_xy[n + j], nx2 = -2 * ((txy << 16) >> 16), ny2 = -2 * (txy >> 16), bj = s_pi[n + j];
                int ci = 0;
#pragma unroll
                for (int r = EMD_RMAX_T - 1; r >= 0; --r)
                  if (sx[r] * nx2 + (sy[r] * ny2 + (sp[r] - bj)) == brc) ci = i0 + r * teff;   // the lowest slot is written last
                mycand = (j << 16) | ci;
                if (brc == g_rc) {
                  int k = j * n + ci - next_arc;
                  if (k < 0) k += n_arcs;
                  myk = (unsigned)k;
                }
              }
              if (cand_on) { s_cand[tid] = mycand; have_cand = true; }
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
          if constexpr (GLOBAL) {
          if (whole_rows_g) {
            fast_done = true;
            // Global-tree pricing: the block is `rows` whole target rows, staged in shared memory (-2x, -2y, potential);
            // the sources stream through registers in tiles of EMD_RG_G x threads (coalesced loads of their coordinates
            // and potentials, read once per block) and every tile meets all rows at register speed.
            constexpr int RG = EMD_RG_G;
            const int rows = cnt / n, j0 = next_arc / n;
            __syncthreads();                                   // the previous block's rows may still be read
            for (int r = tid; r < rows; r += EMD_THREADS) {
              int j = j0 + r;
              while (j >= m) j -= m;
              const int txy = s_xy[n + j];
              s_row[2 * r] = -2 * ((txy << 16) >> 16); s_row[2 * r + 1] = -2 * (txy >> 16);
              s_rowp[r] = s_pi[n + j];
            }
            __syncthreads();
            bool redo = false;
            if (grow_g && !wide_pr) {
              // 32-BIT PRICING.  The potentials are 64-bit because (span^2 + 1) * nodes exceeds int32, yet a reduced cost —
              // a difference of two potentials plus a cost — fits 32 bits on every arc anyone has seen: the scan works on the
              // low words with wrapping arithmetic (exact whenever |rc| < 2^31; a third of the instructions of the 64-bit
              // scan), the arc it selects is re-priced in 64 bits (a non-negative value there = a wrap artefact: the level
              // goes on in 64 bits), and so is the whole arc set once the 32-bit search finds nothing (the certificate).
              int trc = 0;
              int tbase = -1;                                  // the tile in which this thread's minimum was first reached
              for (int base = 0; base < n; base += RG * EMD_THREADS) {
                if (base + tid >= n) break;                    // no source of this tile is this thread's
                unsigned sx[RG], sy[RG], sp[RG];
#pragma unroll
                for (int r = 0; r < RG; ++r) {
                  int i = base + tid + r * EMD_THREADS;
                  if (i >= n) i = base + tid;                  // slots beyond n repeat the thread's own first source of the tile
                  const int xy = s_xy[i];
                  sx[r] = (unsigned)((xy << 16) >> 16); sy[r] = (unsigned)(xy >> 16);
                  sp[r] = (unsigned)s_pi[i];
                }
                const int before = trc;
                for (int r2 = 0; r2 < rows; ++r2) {
                  const unsigned nx2 = (unsigned)s_row[2 * r2], ny2 = (unsigned)s_row[2 * r2 + 1], bj = (unsigned)s_rowp[r2];
#pragma unroll
                  for (int r = 0; r < RG; ++r) trc = min(trc, (int)(sx[r] * nx2 + (sy[r] * ny2 + (sp[r] - bj))));
                }
                if (trc < before) tbase = base;
              }
              int g32;
              {
                const int w_rc = __reduce_min_sync(FULL_MASK, trc);
                int* wrc = reinterpret_cast<int*>(s_wkey + (round & 1) * EMD_WARPS);
                if (lane == 0) wrc[warp] = w_rc;
                __syncthreads();
                g32 = __reduce_min_sync(FULL_MASK, lane < EMD_WARPS ? wrc[lane] : 0);
                ++round;
              }
              if (g32 < 0) {
                unsigned mykey = 0xffffffffu;                  // source index * EMD_ROWS_G + row of the block (ties: see below)
                int2 mycand = make_int2(-1, 0);
                if (trc < 0) {
                  bool found = false;
#pragma unroll 1
                  for (int r = 0; r < RG && !found; ++r) {
                    const int i = tbase + tid + r * EMD_THREADS;
                    if (i >= n) break;
                    const int xy = s_xy[i];
                    const unsigned sxi = (unsigned)((xy << 16) >> 16), syi = (unsigned)(xy >> 16), spi = (unsigned)s_pi[i];
                    for (int r2 = 0; r2 < rows; ++r2) {
                      const int rc = (int)(sxi * (unsigned)s_row[2 * r2] + (syi * (unsigned)s_row[2 * r2 + 1] + (spi - (unsigned)s_rowp[r2])));
                      if (rc == trc) {
                        int j = j0 + r2;
                        while (j >= m) j -= m;
                        mycand = make_int2(i, j);
                        if (trc == g32) mykey = (unsigned)i * EMD_ROWS_G + (unsigned)r2;
                        found = true;
                        break;
                      }
                    }
                  }
                }
                s_candg[tid] = mycand; have_cand = true;
                const unsigned w_k = __reduce_min_sync(FULL_MASK, mykey);
                unsigned* wk = reinterpret_cast<unsigned*>(s_wkey + (round & 1) * EMD_WARPS);
                if (lane == 0) wk[warp] = w_k;
                __syncthreads();
                const unsigned g_key = __reduce_min_sync(FULL_MASK, lane < EMD_WARPS ? wk[lane] : 0xffffffffu);
                ++round;
                int j = j0 + (int)(g_key % EMD_ROWS_G);
                while (j >= m) j -= m;
                const int ie = (int)(g_key / EMD_ROWS_G);
                const int sxy = s_xy[ie], txy = s_xy[n + j];
                const long long rc64 = (long long)((sxy << 16) >> 16) * (-2 * ((txy << 16) >> 16)) +
                                       ((long long)(sxy >> 16) * (-2 * (txy >> 16)) + ((long long)s_pi[ie] - (long long)s_pi[n + j]));
                if (rc64 < 0) {
                  int k = j * n + ie - next_arc;
                  if (k < 0) k += n_arcs;
                  g_k = (unsigned)k;
                  g_rc = (pi_t)rc64;
                } else {                                        // wrap artefact: this block again, in 64 bits
                  wide_pr = true; redo = true; have_cand = false;
                }
              } else {
                g_rc = PI_NONE;
              }
            } else {
            long long trc = 0;
            int tbase = -1;                                    // the tile in which this thread's minimum was first reached
            for (int base = 0; base < n; base += RG * EMD_THREADS) {
              int sx[RG], sy[RG];
              long long sp[RG];
#pragma unroll
              for (int r = 0; r < RG; ++r) {
                const int i = base + tid + r * EMD_THREADS;
                const int xy = i < n ? s_xy[i] : 0;
                sx[r] = (xy << 16) >> 16; sy[r] = xy >> 16;
                sp[r] = i < n ? (long long)s_pi[i] : LLONG_MAX / 4;      // never the minimum
              }
              const long long before = trc;
              for (int r2 = 0; r2 < rows; ++r2) {
                const long long nx2 = s_row[2 * r2], ny2 = s_row[2 * r2 + 1], bj = s_rowp[r2];
#pragma unroll
                for (int r = 0; r < RG; ++r) trc = min(trc, sx[r] * nx2 + (sy[r] * ny2 + (sp[r] - bj)));
              }
              if (trc < before) tbase = base;
            }
            {   // the block's most negative reduced cost
              long long r64 = trc;
              for (int o = 16; o > 0; o >>= 1) r64 = min(r64, shfl_xor_ll(r64, o));
              long long* wr = s_wrc64 + (round & 1) * EMD_WARPS;
              if (lane == 0) wr[warp] = r64;
              __syncthreads();
              r64 = lane < EMD_WARPS ? wr[lane] : 0;
              for (int o = 16; o > 0; o >>= 1) r64 = min(r64, shfl_xor_ll(r64, o));
              g_rc = (pi_t)r64;
              ++round;
            }
            if (g_rc < 0) {
              // The entering arc among the arcs that attain the minimum: lowest source index first, then the first row of
              // the block (a thread meets its sources tile by tile, so "first in block order" would need a second scan; this
              // order costs nothing — first tile, first slot, first row — and the oracle resolves ties the same way).
              // Every thread with a negative minimum locates it the same way: its candidate for the minor iterations.
              unsigned mykey = 0xffffffffu;                    // source index * EMD_ROWS_G + row of the block
              int2 mycand = make_int2(-1, 0);
              if (trc < 0 && (grow_g || trc == (long long)g_rc)) {
                bool found = false;
#pragma unroll 1
                for (int r = 0; r < RG && !found; ++r) {
                  const int i = tbase + tid + r * EMD_THREADS;
                  if (i >= n) break;
                  const int xy = s_xy[i];
                  const long long sxi = (xy << 16) >> 16, syi = xy >> 16, spi = (long long)s_pi[i];
                  for (int r2 = 0; r2 < rows; ++r2) {
                    const long long rc = sxi * s_row[2 * r2] + (syi * s_row[2 * r2 + 1] + (spi - s_rowp[r2]));
                    if (rc == trc) {
                      int j = j0 + r2;
                      while (j >= m) j -= m;
                      mycand = make_int2(i, j);
                      if (trc == (long long)g_rc) mykey = (unsigned)i * EMD_ROWS_G + (unsigned)r2;
                      found = true;
                      break;
                    }
                  }
                }
              }
              if (grow_g) { s_candg[tid] = mycand; have_cand = true; }
              const unsigned w_k = __reduce_min_sync(FULL_MASK, mykey);
              unsigned* wk = reinterpret_cast<unsigned*>(s_wkey + (round & 1) * EMD_WARPS);
              if (lane == 0) wk[warp] = w_k;
              __syncthreads();
              const unsigned g_key = __reduce_min_sync(FULL_MASK, lane < EMD_WARPS ? wk[lane] : 0xffffffffu);
              {
                int j = j0 + (int)(g_key % EMD_ROWS_G);
                while (j >= m) j -= m;
                int k = j * n + (int)(g_key / EMD_ROWS_G) - next_arc;
                if (k < 0) k += n_arcs;
                g_k = (unsigned)k;
              }
              ++round;
            } else {
              g_rc = PI_NONE;
            }
            }
            if (redo) { --empty; continue; }
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
            cta_lexmin64(r64, k32, round);
            g_rc = (pi_t)r64; g_k = k32;
          }
          if constexpr (!WIDE) ++round;
          }
        }
#ifdef STM_EMD_PROFILE
        { const long long _c = clock64(); c_red += _c - c_t1; c_t1 = _c; }
#endif
        const int start = next_arc;
        scanned += cnt;
        next_arc += cnt;
        if (next_arc >= n_arcs) next_arc -= n_arcs;
        if (whole_rows) { next_row += cnt == Bk ? rows_B << grow : cnt / n; if (next_row >= m) next_row -= m; }
        if (g_rc < 0) {
          best_rc = g_rc;
          best_e = start + (int)g_k;
          if (best_e >= n_arcs) best_e -= n_arcs;
          break;
        }
      }
      if constexpr (GLOBAL) {
        // nothing negative in the low words: the certificate of optimality is a sweep over all arcs in 64 bits
        if (best_e < 0 && grow_g && !wide_pr) { wide_pr = true; scanned = 0; certify = true; }
      }
      } while (certify);
      PROF_T(c_price)
#ifdef STM_EMD_PROFILE
      n_scanned += scanned;
#endif
      if (best_e < 0) break;                      // optimal
      if (it >= level_max_iter) { status = 1; break; }
      ++it;
      const int ej = cand_j >= 0 ? cand_j : best_e / n, ei = best_e - ej * n;
      const int u_i = ei, u_j = n + ej;
      // ---- cycle: the two tree paths from the entering arc's endpoints up to their join node -----------------
      // A node v is an ancestor of u iff pos[v] <= pos[u] < pos[v] + size[v].  Instead of chasing parent pointers from
      // both endpoints (two lanes, ~180 cycles per step, everyone else waiting), every thread tests a run of consecutive
      // PREORDER POSITIONS: path A = ancestors-or-self of u_i that are not ancestors of u_j, path B the converse; along
      // a path the position decreases, so an ordered compaction (one block scan) yields both paths in walking order.
      if constexpr (GLOBAL) {
        // Global-tree kernel: the tree lives in L2, where that scan costs two loads per position of a 10^5-node preorder
        // (~130k cycles).  Here two lanes climb instead — lane 0 of warp 0 from u_i (path A), lane 0 of warp 1 from u_j
        // (path B): position, size, parent and flow of a node are loaded together, so a step costs one L2 round trip,
        // ~40 steps per side.  The leaving arc is tracked on the way (same rule: smallest flow; on side A the first such
        // arc met, on side B the last, B before A).
        if (lane == 0 && warp < 2) {
          const int other = warp == 0 ? u_j : u_i;
          const int po = s_pos[other];
          idx_t* const path = warp == 0 ? s_pathA : s_pathB;
          int v = warp == 0 ? u_i : u_j, len = 0, didx = -1;
          long long dmin = LLONG_MAX;
          for (;;) {
            const int pv = s_pos[v], sv = s_size[v], par = s_parent[v];
            const long long f = FL(v);
            if ((unsigned)(po - pv) < (unsigned)sv) break;          // v is an ancestor of the other endpoint: the join
            if (len < EMD_MAXPATH) path[len] = (idx_t)v;
            if (warp == 0 ? (v < n && f < dmin) : (v >= n && f <= dmin)) { dmin = f; didx = len; }
            ++len;
            v = par;
          }
          s_i[1 + warp] = len;
          s_ll[warp] = dmin; s_ll[EMD_WARPS + warp] = didx;
        }
      } else {
        const int pu = s_pos[u_i], pw = s_pos[u_j];
        int per = (N + EMD_THREADS - 1) / EMD_THREADS;             // <= 30 * SCAN_CH for every problem that fits
        per += (2 - per) & 3;   // per = 2 (mod 4): a lane stride of per uint16 is an odd number of 4-byte words, so the
                                // loads of s_order[x0 + q] across a warp hit 32 different banks
        const int x0 = tid * per, x1 = min(N, x0 + per);
        unsigned mA[SCAN_CH], mB[SCAN_CH];                         // one bit per position of the run, 32 per chunk
#pragma unroll
        for (int c = 0; c < SCAN_CH; ++c) { mA[c] = 0u; mB[c] = 0u; }
        constexpr int PER_UNROLL = 4;                              // independent load pairs in flight
        const int xlim = min(x1, max(pu, pw) + 1);                 // ancestors precede their descendants
#pragma unroll
        for (int c = 0; c < SCAN_CH; ++c) {
          const int xc0 = x0 + 32 * c, xc1 = min(xlim, xc0 + 32);
          for (int xb = xc0; xb < xc1; xb += PER_UNROLL) {
            int vv[PER_UNROLL], sz[PER_UNROLL];
#pragma unroll
            for (int q = 0; q < PER_UNROLL; ++q) vv[q] = xb + q < xc1 ? s_order[xb + q] : root;
#pragma unroll
            for (int q = 0; q < PER_UNROLL; ++q) sz[q] = xb + q < xc1 ? (int)s_size[vv[q]] : 0;   // 0: never an ancestor
#pragma unroll
            for (int q = 0; q < PER_UNROLL; ++q) {
              const int x = xb + q;
              const unsigned inU = (unsigned)(pu - x) < (unsigned)sz[q], inW = (unsigned)(pw - x) < (unsigned)sz[q];
              mA[c] |= (inU & ~inW & 1u) << ((x - xc0) & 31);
              mB[c] |= (inW & ~inU & 1u) << ((x - xc0) & 31);
            }
          }
        }
        int cntA = 0, cntB = 0;
#pragma unroll
        for (int c = 0; c < SCAN_CH; ++c) { cntA += __popc(mA[c]); cntB += __popc(mB[c]); }
        int nA, nB, ra, rb;                                        // totals; ranks in ascending position = distance from the join
        {
          // exclusive block scan of the two counts, packed in one word
          const int cnt = cntA | (cntB << 16);
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
          nA = total & 0xffff; nB = total >> 16;
          ra = before & 0xffff; rb = before >> 16;
        }
        if (tid == 0) { s_i[1] = nA; s_i[2] = nB; }
        // ---- leaving arc, found while the paths are written: min flow over the arcs traversed against their direction;
        //      ties by the strongly-feasible rule (i side: closest to i; j side wins and takes the one closest to the join)
        //      = lexicographic (flow ascending, sequence number descending), sequence = lenA - 1 - a on path A (index a),
        //      lenA + b on path B.  Every thread looks at the entries it writes, the warps combine by shuffles and the
        //      per-warp results meet in shared memory behind the barrier the paths need anyway.
        long long dmin = LLONG_MAX;
        int dseq = -1;
        if (nA <= EMD_MAXPATH && nB <= EMD_MAXPATH) {
#pragma unroll
          for (int c = 0; c < SCAN_CH; ++c) {
            for (unsigned mm = mA[c]; mm; mm &= mm - 1) {
              const int v = s_order[x0 + 32 * c + __ffs(mm) - 1];
              s_pathA[nA - 1 - ra] = (idx_t)v;
              if (v < n) { const long long f = FL(v); if (f < dmin || (f == dmin && ra > dseq)) { dmin = f; dseq = ra; } }
              ++ra;
            }
            for (unsigned mm = mB[c]; mm; mm &= mm - 1) {
              const int v = s_order[x0 + 32 * c + __ffs(mm) - 1];
              s_pathB[nB - 1 - rb] = (idx_t)v;
              const int sq = nA + nB - 1 - rb;
              if (v >= n) { const long long f = FL(v); if (f < dmin || (f == dmin && sq > dseq)) { dmin = f; dseq = sq; } }
              ++rb;
            }
          }
        }
        {
          // flows are < 2^50 + 1 and sequence numbers < 2 * EMD_MAXPATH_S = 2^12: one 64-bit key (flow, 8191 - sequence),
          // minimised by two 32-bit REDUX steps (high word, then low word among the lanes that hold the minimum)
          const unsigned long long key = dseq < 0 ? ~0ull : ((unsigned long long)dmin << 13) | (unsigned)(0x1fff - dseq);
          const unsigned kh = (unsigned)(key >> 32), kl = (unsigned)key;
          const unsigned mh = __reduce_min_sync(FULL_MASK, kh);
          const unsigned ml = __reduce_min_sync(FULL_MASK, kh == mh ? kl : 0xffffffffu);
          if (lane == 0) s_ll[warp] = (long long)(((unsigned long long)mh << 32) | ml);
        }
      }
      __syncthreads();
      const int lenA = s_i[1], lenB = s_i[2];
      PROF_T(c_walk)
      if (lenA > EMD_MAXPATH || lenB > EMD_MAXPATH) { status = 4; break; }
      long long delta;
      int seq;
      if constexpr (!GLOBAL) {   // every warp combines the per-warp results: no second barrier
        const unsigned long long key = lane < EMD_WARPS ? (unsigned long long)s_ll[lane] : ~0ull;
        const unsigned kh = (unsigned)(key >> 32), kl = (unsigned)key;
        const unsigned mh = __reduce_min_sync(FULL_MASK, kh);
        const unsigned ml = __reduce_min_sync(FULL_MASK, kh == mh ? kl : 0xffffffffu);
        const unsigned long long best = ((unsigned long long)mh << 32) | ml;
        delta = (long long)(best >> 13);
        seq = best == ~0ull ? -1 : 0x1fff - (int)(best & 0x1fffu);
      } else {
        const long long fA = s_ll[0], fB = s_ll[1];
        const int ia = (int)s_ll[EMD_WARPS], ib = (int)s_ll[EMD_WARPS + 1];
        if (ib >= 0 && (ia < 0 || fB <= fA)) { delta = fB; seq = lenA + ib; }
        else { delta = fA; seq = ia >= 0 ? lenA - 1 - ia : -1; }
      }
      PROF_T(c_leave)
      if (seq < 0) { status = 5; break; }        // unbounded: cannot happen for a transport LP
      const int side = seq >= lenA ? 1 : 0;
      const int k = side == 0 ? lenA - 1 - seq : seq - lenA;   // leaving arc at path index k of its side
      const idx_t* pathX = side == 0 ? s_pathA : s_pathB;
      const idx_t* pathY = side == 0 ? s_pathB : s_pathA;
      const int lenX = side == 0 ? lenA : lenB, lenY = side == 0 ? lenB : lenA;
      const int q = pathX[k];
      const int t_in = side == 0 ? u_j : u_i;
      const int s = s_size[q], a0 = s_pos[q];
      const int pt = s_pos[t_in];
      const int ptr = pt < a0 ? pt : pt - s;
      const pi_t shift = side == 0 ? -best_rc : best_rc;
      // ---- flow update along the cycle + old flows / positions / sizes of the reversed path --------
      // side Y and the part of side X above the leaving arc in place; the new flows of pathX[0 .. k-1] move one node up the
      // reversed path, so they wait in registers until everyone has read (the arc of pathX[k] leaves the basis)
      constexpr int MAXT = (EMD_MAXPATH + EMD_THREADS - 1) / EMD_THREADS;
      long long fo[MAXT];
      for (int t = tid; t < lenY; t += EMD_THREADS) { const int v = pathY[t]; FL(v) += ((v < n) == (side == 1)) ? -delta : delta; }
#pragma unroll
      for (int r = 0; r < MAXT; ++r) {
        const int t = tid + r * EMD_THREADS;
        fo[r] = 0;
        if (t < lenX) {
          const int v = pathX[t];
          const long long f = FL(v) + (((v < n) == (side == 0)) ? -delta : delta);
          if (t > k) FL(v) = f; else fo[r] = f;
        }
      }
      for (int t = tid; t <= k; t += EMD_THREADS) { const int v = pathX[t]; s_P[t] = s_pos[v]; s_S[t] = s_size[v]; }
      // potential shift of the re-hung subtree (a contiguous preorder segment), copy of the affected preorder range
      const int lo = pt < a0 ? pt + 1 : a0, hi = pt < a0 ? a0 + s : pt + 1;
      if constexpr (GLOBAL) {
        // the tree lives in L2: one pass over the range (the subtree is part of it), four independent loads in flight
        constexpr int UB = 4;
        for (int x = lo + tid; x < hi; x += UB * EMD_THREADS) {
          int o[UB];
          pi_t pv[UB];
#pragma unroll
          for (int q = 0; q < UB; ++q) { const int xq = x + q * EMD_THREADS; o[q] = xq < hi ? (int)s_order[xq] : 0; }
#pragma unroll
          for (int q = 0; q < UB; ++q) { const int xq = x + q * EMD_THREADS; pv[q] = (xq >= a0 && xq < a0 + s) ? s_pi[o[q]] : 0; }
#pragma unroll
          for (int q = 0; q < UB; ++q) {
            const int xq = x + q * EMD_THREADS;
            if (xq < hi) { s_tmp[xq] = (idx_t)o[q]; if (xq >= a0 && xq < a0 + s) s_pi[o[q]] = pv[q] + shift; }
          }
        }
      } else {
        for (int x = a0 + tid; x < a0 + s; x += EMD_THREADS) s_pi[s_order[x]] += shift;
        for (int x = lo + tid; x < hi; x += EMD_THREADS) s_tmp[x] = s_order[x];
      }
      __syncthreads();
      // ---- reversed path: parents, flows, sizes (flows read before anyone overwrites them) ----------
      {
#pragma unroll
        for (int r = 0; r < MAXT; ++r) {
          const int t = tid + r * EMD_THREADS;
          if (t <= k) {
            const int v = pathX[t];
            if (t == 0) { s_parent[v] = (idx_t)t_in; FL(v) = delta; s_size[v] = (idx_t)s; }
            else { s_parent[v] = pathX[t - 1]; s_size[v] = (idx_t)(s - s_S[t - 1]); }
            if (t < k) FL(pathX[t + 1]) = fo[r];
          }
        }
      }
      for (int t = k + 1 + tid; t < lenX; t += EMD_THREADS) s_size[pathX[t]] -= (idx_t)s;
      for (int t = tid; t < lenY; t += EMD_THREADS) s_size[pathY[t]] += (idx_t)s;
      // ---- preorder re-numbering of the affected range ------------------------------------------------
      constexpr int UR = GLOBAL ? 4 : 1;         // global tree: four copied entries loaded before they are placed
      for (int xb = lo + tid; xb < hi; xb += UR * EMD_THREADS) {
        int vq[UR];
#pragma unroll
        for (int q = 0; q < UR; ++q) { const int xq = xb + q * EMD_THREADS; vq[q] = xq < hi ? (int)s_tmp[xq] : 0; }
#pragma unroll
        for (int q = 0; q < UR; ++q) {
        const int x = xb + q * EMD_THREADS;
        if (x >= hi) break;
        const int v = vq[q];
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
        s_order[f] = (idx_t)v;
        s_pos[v] = (idx_t)f;
        }
      }
      __syncthreads();
      PROF_T(c_update)
#ifdef STM_EMD_PROFILE
      sum_len += lenA + lenB; sum_s += s; sum_range += hi - lo;
#endif
    }
#ifdef STM_EMD_PROFILE
    if (tid == 0 && g < 2 && it > 0)
      printf("[emd prof] g=%d lev=%d n=%d m=%d pivots=%lld cycles/pivot: price %.0f walk %.0f leave %.0f update %.0f | arcs scanned/pivot %.0f cycle len %.1f subtree %.1f range %.1f | blocks/pivot %.2f, per block: scan %.0f reduce %.0f cycles | minor pivots %lld\n",
             g, lev, n, m, it, (double)c_price / it, (double)c_walk / it, (double)c_leave / it, (double)c_update / it,
             (double)n_scanned / it, (double)sum_len / it, (double)sum_s / it, (double)sum_range / it,
             (double)n_blocks / it, (double)c_scan / max(1LL, n_blocks), (double)c_red / max(1LL, n_blocks), n_minor);
#endif
    it_total += it;
#ifdef STM_EMD_PROFILE
    const long long c_l2 = clock64();
#endif
    if (status != 0) break;
    if (li > 0) {
      EMD_SCRATCH(sc);
      // ---- refinement: this level's optimum -> feasible forest of level lev - 1 --------------------------------
      // Stage 1, per coarse source I: north-west corner between its children (in order) and its arcs (by increasing
      // target) gives pieces (fine source, arc, flow).  Stage 2, per coarse target J: north-west corner between the
      // pieces of its arcs (by increasing source) and its children gives the fine arcs, grouped by fine target.
      const int nS = n, nT = m;
      // the finer level in use is one or two bit levels below; across a skipped level the child ranges compose (descendants
      // stay contiguous in Morton order)
      const int flev = s_lv[li - 1] >> 20;
      const bool two = lev - flev == 2;
      const int32_t* cs2 = p.hs.child + (size_t)lev * (p.hs.stride + 1);
      const int32_t* ct2 = sc.ht.child + (size_t)lev * (sc.ht.stride + 1);
      const int32_t* cs1 = p.hs.child + (size_t)(lev - 1) * (p.hs.stride + 1);
      const int32_t* ct1 = sc.ht.child + (size_t)(lev - 1) * (sc.ht.stride + 1);
      auto CS = [&](int I) { const int c = cs2[I]; return two ? cs1[c] : c; };
      auto CT = [&](int J) { const int c = ct2[J]; return two ? ct1[c] : c; };
      const int64_t* fs = p.hs.mass + (size_t)flev * p.hs.stride;
      const int64_t* ft = sc.ht.mass + (size_t)flev * sc.ht.stride;
      const int mf = s_lv[li - 1] & 0xfffff;
      int Ec = 0;
      {   // real tree arcs with positive flow, compacted in node order
        const int per = (nS + nT + EMD_THREADS - 1) / EMD_THREADS;
        const int b = min(nS + nT, tid * per), e = min(nS + nT, b + per);
        int c = 0;
        for (int v = b; v < e; ++v) c += (s_parent[v] != root && FL(v) > 0) ? 1 : 0;
        int at = cta_excl_scan<EMD_THREADS>(c, s_ws, &Ec);
        for (int v = b; v < e; ++v) {
          const int pr = s_parent[v];
          const long long f = FL(v);
          if (pr != root && f > 0) {
            sc.aI[at] = v < nS ? v : pr; sc.aJ[at] = (v < nS ? pr : v) - nS; sc.aF[at] = f;
            ++at;
          }
        }
      }
      for (int k = tid; k <= nS; k += EMD_THREADS) sc.offS[k] = 0;
      for (int k = tid; k <= nT; k += EMD_THREADS) sc.offT[k] = 0;
      __syncthreads();
      for (int e = tid; e < Ec; e += EMD_THREADS) { atomicAdd(&sc.offS[sc.aI[e] + 1], 1); atomicAdd(&sc.offT[sc.aJ[e] + 1], 1); }
      __syncthreads();
      cta_prefix_inplace<EMD_THREADS>(sc.offS, nS + 1, s_ws);
      cta_prefix_inplace<EMD_THREADS>(sc.offT, nT + 1, s_ws);
      for (int k = tid; k <= nS; k += EMD_THREADS) sc.curS[k] = sc.offS[k];
      for (int k = tid; k <= nT; k += EMD_THREADS) sc.curT[k] = sc.offT[k];
      __syncthreads();
      for (int e = tid; e < Ec; e += EMD_THREADS) {
        const int at = atomicAdd(&sc.curS[sc.aI[e]], 1);
        EMD_CHECK(at >= 0 && at < Ec && sc.aI[e] < nS && sc.aJ[e] < nT, "coarse arc list slot");
        sc.lstS[at] = e;
      }
      __syncthreads();
      for (int I = tid; I < nS; I += EMD_THREADS) {      // arcs of I by increasing target
        const int a = sc.offS[I], b = sc.offS[I + 1];
        for (int x = a + 1; x < b; ++x) {
          const int ke = sc.lstS[x], kj = sc.aJ[ke];
          int y = x - 1;
          while (y >= a && sc.aJ[sc.lstS[y]] > kj) { sc.lstS[y + 1] = sc.lstS[y]; --y; }
          sc.lstS[y + 1] = ke;
        }
      }
      if (tid == 0) sc.pcFirst[0] = 0;
      __syncthreads();
      // stage 1: count, offsets, write
      for (int pass = 0; pass < 2; ++pass) {
        for (int I = tid; I < nS; I += EMD_THREADS) {
          int a = CS(I);
          const int a_end = CS(I + 1);
          long long ra = a < a_end ? fs[a] : 0;
          for (int r = sc.offS[I]; r < sc.offS[I + 1]; ++r) {
            long long rb = sc.aF[sc.lstS[r]];
            int np = pass == 0 ? 0 : sc.pcFirst[r];
            while (rb > 0 && a < a_end) {
              const long long f = ra < rb ? ra : rb;
              if (f > 0) { if (pass == 1) { EMD_CHECK(np < 2 * cap, "piece index"); sc.pI[np] = a; sc.pF[np] = f; } ++np; }
              ra -= f; rb -= f;
              if (ra == 0) { ++a; ra = a < a_end ? fs[a] : 0; }
            }
            if (pass == 0) sc.pcFirst[r + 1] = np;
          }
        }
        __syncthreads();
        if (pass == 0) cta_prefix_inplace<EMD_THREADS>(sc.pcFirst, Ec + 1, s_ws);
      }
      // arcs (ranks in the (I, J) order) of every coarse target, by increasing source
      for (int r = tid; r < Ec; r += EMD_THREADS) sc.lstT[atomicAdd(&sc.curT[sc.aJ[sc.lstS[r]]], 1)] = r;
      __syncthreads();
      for (int J = tid; J < nT; J += EMD_THREADS) {
        const int a = sc.offT[J], b = sc.offT[J + 1];
        for (int x = a + 1; x < b; ++x) {
          const int key = sc.lstT[x];
          int y = x - 1;
          while (y >= a && sc.lstT[y] > key) { sc.lstT[y + 1] = sc.lstT[y]; --y; }
          sc.lstT[y + 1] = key;
        }
      }
      if (tid == 0) sc.aoff[0] = 0;
      __syncthreads();
      // stage 2: count, offsets, write
      for (int pass = 0; pass < 2; ++pass) {
        for (int J = tid; J < nT; J += EMD_THREADS) {
          int b = CT(J);
          const int b_end = CT(J + 1);
          long long rb = b < b_end ? ft[b] : 0;
          int E = pass == 0 ? 0 : sc.aoff[J];
          if (pass == 1 && b < b_end) sc.tfirst[b] = E;
          for (int q = sc.offT[J]; q < sc.offT[J + 1]; ++q) {
            const int r = sc.lstT[q];
            for (int pz = sc.pcFirst[r]; pz < sc.pcFirst[r + 1]; ++pz) {
              long long ra = sc.pF[pz];
              while (ra > 0 && b < b_end) {
                const long long f = ra < rb ? ra : rb;
                if (f > 0) { if (pass == 1) { EMD_CHECK(E < cap, "fine arc index"); sc.fI[E] = sc.pI[pz]; sc.fJ[E] = b; sc.fF[E] = f; } ++E; }
                ra -= f; rb -= f;
                if (rb == 0) { ++b; rb = b < b_end ? ft[b] : 0; if (pass == 1 && b < b_end) sc.tfirst[b] = E; }
              }
            }
          }
          if (pass == 0) sc.aoff[J + 1] = E;
          else for (int bb = b + 1; bb < b_end; ++bb) sc.tfirst[bb] = E;     // trailing children without demand
        }
        __syncthreads();
        if (pass == 0) cta_prefix_inplace<EMD_THREADS>(sc.aoff, nT + 1, s_ws);
      }
      if (tid == 0) { const int Eall = sc.aoff[nT]; EMD_CHECK(Eall <= cap && Eall < mf + p.src_hdr[flev], "forest size"); sc.tfirst[mf] = Eall; s_i[8] = Eall; }
      __syncthreads();
    }
#ifdef STM_EMD_PROFILE
    if (tid == 0 && g < 2)
      printf("[emd prof] g=%d lev=%d nodes %d: basis %lld, simplex %lld (%lld pivots), refine %lld cycles\n", g, lev, N,
             (long long)(c_l1 - c_l0), (long long)(c_l2 - c_l1), it, (long long)(clock64() - c_l2));
#endif
    }   // levels
    if (status >= 4 && !star) { star = true; continue; }   // a multiscale start that hit an internal limit: redo from the star
    break;
    }   // attempts
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
      if (p.out_iters) p.out_iters[g] = it_total;
    }
  }
}

inline size_t emd_node_bytes(bool wide) { return (wide ? 8 : 4) + 4 + 5 * 2; }   // potential, packed xy, five uint16 tree arrays
inline size_t emd_smem_bytes(int64_t cap_nodes, int64_t flow_nodes = 0, bool wide = false) {
  return (size_t)flow_nodes * 8 + (size_t)cap_nodes * emd_node_bytes(wide) + 4 * (size_t)EMD_MAXPATH_S * 2 + 64;
}
inline int64_t emd_ws_cap(int64_t n_src, int64_t max_tgt) { return (n_src + max_tgt + 2 + 7) & ~7LL; }
inline int64_t emd_pow2(int64_t n) { int64_t p = 1; while (p < n) p <<= 1; return p; }

// global part of the workspace shared by all CTAs (queue counter, source hierarchy), then the per-CTA slices
struct EmdGlobalLayout {
  size_t counter, hdr, src_int, sortbuf, codeA, codeB, hs_mass, hs_xy, hs_child, cta0, cta_bytes, total;
};
inline EmdGlobalLayout emd_global_layout(int64_t n_src, int64_t max_tgt, int64_t n_ctas) {
  EmdGlobalLayout g;
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t at = off; off += (bytes + 255) & ~(size_t)255; return at; };
  g.counter = take(256);
  g.hdr = take(256);
  g.src_int = take(8 * n_src);
  g.sortbuf = take(8 * emd_pow2(n_src));
  g.codeA = take(4 * n_src);
  g.codeB = take(4 * n_src);
  g.hs_mass = take(8 * (size_t)EMD_LV * n_src);
  g.hs_xy = take(4 * (size_t)EMD_LV * n_src);
  g.hs_child = take(4 * (size_t)EMD_LV * (n_src + 1));
  // the global-tree regions are sized whenever the batch might not fit shared memory (the smallest capacity, that of the
  // 64-bit kernel, is ~9.6k nodes); they sit at the end of a slice, so the shared-memory kernels never see them
  g.cta_bytes = (emd_scratch_layout(nullptr, emd_ws_cap(n_src, max_tgt), max_tgt > 0 ? max_tgt : 1, nullptr,
                                    emd_ws_cap(n_src, max_tgt) > EMD_GLOBAL_ALLOC_NODES) + 255) & ~(size_t)255;
  g.cta0 = off;
  g.total = off + g.cta_bytes * (size_t)n_ctas + 256;
  return g;
}

}  // namespace
}  // namespace stm

extern "C" int64_t stm_emd_workspace_bytes(int32_t n_src, int64_t max_tgt, int32_t n_ctas) {
  if (n_src < 0 || max_tgt < 0 || n_ctas <= 0) return -1;
  return (int64_t)stm::emd_global_layout(n_src, max_tgt, n_ctas).total;
}

namespace stm {
namespace {
int emd_launch(bool wide, const int32_t* src_x, const int32_t* src_y, const double* src_mass, int32_t n_src,
               const int64_t* tgt_ptr, const int32_t* tgt_x, const int32_t* tgt_y,
               const double* tgt_mass, int32_t G, int64_t max_tgt, const int32_t* order,
               int64_t max_iter, int32_t block_size, uint32_t flags,
               void* workspace, int64_t workspace_bytes, int32_t n_ctas,
               double* out_cost, int32_t* out_status, int64_t* out_iters, void* stream) {
  STM_CHECK_ARG(G >= 0, "G < 0");
  if (G == 0) return STM_OK;
  STM_CHECK_ARG(n_src > 0 && src_x && src_y && src_mass, "empty or null source image");
  STM_CHECK_ARG(tgt_ptr && tgt_x && tgt_y && tgt_mass && out_cost && out_status, "null pointer");
  STM_CHECK_ARG(max_tgt >= 0 && n_ctas > 0 && max_iter >= 0 && block_size >= 0, "bad size argument");
  STM_CHECK_ARG(workspace && workspace_bytes >= stm_emd_workspace_bytes(n_src, max_tgt, n_ctas), "workspace too small");
  STM_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "workspace must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int dev = 0, max_optin = 0;
  STM_CHECK_CUDA(cudaGetDevice(&dev));
  STM_CHECK_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  // node capacity of one CTA: the whole tree lives in shared memory; larger problems get status 3
  const int64_t ws_cap = emd_ws_cap(n_src, max_tgt);
  int64_t cap = ws_cap;
  const int64_t max_cap = (((int64_t)max_optin - 4096 - (int64_t)emd_smem_bytes(0)) / (int64_t)emd_node_bytes(wide)) & ~7LL;   // static smem + reserve
  if (cap > max_cap) cap = max_cap;
  char* ws = static_cast<char*>(workspace);
  const EmdGlobalLayout gl = emd_global_layout(n_src, max_tgt, n_ctas);
  EmdArgs a{};
  a.src_x = src_x; a.src_y = src_y; a.src_mass = src_mass; a.n_src = n_src;
  a.tgt_ptr = tgt_ptr; a.tgt_x = tgt_x; a.tgt_y = tgt_y; a.tgt_mass = tgt_mass;
  a.G = G; a.cap_nodes = (int32_t)cap; a.cap_tgt = (int32_t)(max_tgt > 0 ? max_tgt : 1);
  a.max_iter = max_iter; a.block_size = block_size; a.flags = flags;
  a.counter = reinterpret_cast<unsigned int*>(ws + gl.counter);
  a.src_hdr = reinterpret_cast<int32_t*>(ws + gl.hdr);
  a.src_int = reinterpret_cast<int64_t*>(ws + gl.src_int);
  a.hs.mass = reinterpret_cast<int64_t*>(ws + gl.hs_mass);
  a.hs.xy = reinterpret_cast<int32_t*>(ws + gl.hs_xy);
  a.hs.child = reinterpret_cast<int32_t*>(ws + gl.hs_child);
  a.hs.stride = n_src;
  a.cta_ws = ws + gl.cta0; a.cta_ws_bytes = (int64_t)gl.cta_bytes;
  a.order = order;
  a.out_cost = out_cost; a.out_status = out_status; a.out_iters = out_iters;
  STM_CHECK_CUDA(cudaMemsetAsync(ws + gl.counter, 0, 512, st));
  emd_prep_kernel<<<1, EMD_PREP_THREADS, 0, st>>>(src_x, src_y, src_mass, n_src, a.src_int,
                                                  reinterpret_cast<unsigned long long*>(ws + gl.sortbuf), (int)emd_pow2(n_src),
                                                  reinterpret_cast<uint32_t*>(ws + gl.codeA), reinterpret_cast<uint32_t*>(ws + gl.codeB),
                                                  a.hs, a.src_hdr);
  STM_CHECK_CUDA(cudaGetLastError());
  const int grid = n_ctas < G ? n_ctas : G;
  if (ws_cap > max_cap) {
    // ---- the batch does not fit a CTA's shared memory: tree in the CTA's global scratch (one CTA per SM) ----
    a.cap_nodes = (int32_t)ws_cap;
    const int64_t paths = 4 * (int64_t)EMD_MAXPATH_G * 4 + 64;
    int64_t fcap_g = ((int64_t)max_optin - 8192 - paths) / 8;       // static shared memory incl. the candidate list
    fcap_g = std::max<int64_t>(0, std::min<int64_t>(fcap_g, ws_cap)) & ~1LL;
    a.flow_smem_nodes = (int32_t)fcap_g;
    const size_t smem_g = (size_t)fcap_g * 8 + (size_t)paths;
    auto kern = emd_kernel<EMD_GLOBAL_THREADS, true, true, true>;
    STM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g));
    kern<<<grid, EMD_GLOBAL_THREADS, smem_g, st>>>(a);
    STM_CHECK_CUDA(cudaGetLastError());
    return STM_OK;
  }
  // the kernel carves the scratch of a CTA with its own cap (<= ws_cap), so the slices never exceed what was sized
  const size_t tree_smem = emd_smem_bytes((int)cap, 0, wide);
  // Problems that leave room for only one CTA per SM get 512 threads and the register-resident pricing path;
  // small problems run several 256-thread CTAs per SM.
  const bool one_per_sm = tree_smem * 2 + 2048 > (size_t)max_optin;
  // what is left of a CTA's share of the SM's shared memory (without lowering the CTAs per SM) holds arc flows
  const int64_t per_sm = one_per_sm ? 1 : std::min<int64_t>(8, (int64_t)max_optin / (int64_t)(tree_smem + 1024));
  const int64_t share = one_per_sm ? (int64_t)max_optin - 4096 : (int64_t)max_optin / per_sm - 1024 - 2560;   // system reserve + static
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
                                     int64_t max_iter, int32_t block_size, uint32_t flags,
                                     void* workspace, int64_t workspace_bytes, int32_t n_ctas,
                                     double* out_cost, int32_t* out_status, int64_t* out_iters, void* stream) {
  return stm::emd_launch(false, src_x, src_y, src_mass, n_src, tgt_ptr, tgt_x, tgt_y, tgt_mass, G, max_tgt, order, max_iter,
                         block_size, flags, workspace, workspace_bytes, n_ctas, out_cost, out_status, out_iters, stream);
}

extern "C" int stm_emd_spots_batched_wide(const int32_t* src_x, const int32_t* src_y, const double* src_mass, int32_t n_src,
                                          const int64_t* tgt_ptr, const int32_t* tgt_x, const int32_t* tgt_y,
                                          const double* tgt_mass, int32_t G, int64_t max_tgt, const int32_t* order,
                                          int64_t max_iter, int32_t block_size, uint32_t flags,
                                          void* workspace, int64_t workspace_bytes, int32_t n_ctas,
                                          double* out_cost, int32_t* out_status, int64_t* out_iters, void* stream) {
  return stm::emd_launch(true, src_x, src_y, src_mass, n_src, tgt_ptr, tgt_x, tgt_y, tgt_mass, G, max_tgt, order, max_iter,
                         block_size, flags, workspace, workspace_bytes, n_ctas, out_cost, out_status, out_iters, stream);
}

extern "C" int32_t stm_emd_smem_node_capacity(int32_t wide) {
  int dev = 0, max_optin = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return -1;
  const int64_t max_cap = (((int64_t)max_optin - 4096 - (int64_t)stm::emd_smem_bytes(0)) / (int64_t)stm::emd_node_bytes(wide != 0)) & ~7LL;
  return (int32_t)(max_cap - 10);     // n + m + 1 nodes, two spare entries, rounding of the workspace capacity
}

extern "C" int32_t stm_emd_pricing_threads(int32_t n_src, int64_t max_tgt, int32_t wide) {
  // threads per problem of the candidate-list pricing a batch of this size runs with (what the oracle's cand_threads mirrors)
  if (wide) return 0;
  int dev = 0, max_optin = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return -1;
  const int64_t ws_cap = stm::emd_ws_cap(n_src, max_tgt);
  const int64_t max_cap = (((int64_t)max_optin - 4096 - (int64_t)stm::emd_smem_bytes(0)) / (int64_t)stm::emd_node_bytes(false)) & ~7LL;
  if (ws_cap > max_cap) return -1;                                  // global-tree kernel (its own candidate classes and tie rule)
  return stm::emd_smem_bytes(ws_cap, 0, false) * 2 + 2048 > (size_t)max_optin ? STM_EMD_T_BIG : 256;
}

extern "C" int32_t stm_emd_global_node_capacity(void) { return (int32_t)stm::EMD_GLOBAL_MAX_NODES; }
