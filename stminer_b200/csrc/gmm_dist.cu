// GMM-to-GMM distance for sm_100a: K x K Hellinger cost + exact linear assignment, one warp per pair.
//
// Restates STMiner/Algorithm/distance.py:13-36 (distribution_distance), :50-80
// (get_hellinger_distance) and STMiner/Algorithm/algorithm.py:10-26 (linear_sum ->
// scipy.optimize.linear_sum_assignment, a shortest-augmenting-path LSAP).  All arithmetic is
// float64 like the reference; the LSAP is exact (dual potentials + Dijkstra over columns with the
// columns spread over the lanes of the warp).
#include <limits.h>
#include <math.h>

#include "stm_common.cuh"

namespace stm {
namespace {

constexpr int DIST_THREADS = 128;
constexpr int DIST_WARPS = DIST_THREADS / 32;

// per-component record in shared memory: w, mx, my, cxx, cxy, cyy, det^(1/4), pad.  A stride of 9 doubles
// (18 words) keeps the 64-bit loads of 16 consecutive components on 16 distinct bank pairs; 8 made them collide.
constexpr int COMP_STRIDE = 9;

struct DistArgs {
  const double* w; const double* mu; const double* cov;   // [G,K], [G,K,2], [G,K,2,2]
  const double* qw; const double* qmu; const double* qcov;  // one-vs-all query or NULL
  int32_t G, K;
  int64_t pair_begin, pair_end;  // triangle mode: linear pair range; one-vs-all: [0, G)
  double* out;
};

__device__ __forceinline__ void pair_to_ij(int64_t p, int G, int& i, int& j) {
  // strict upper triangle, row-major: row i starts at i*(2G-i-1)/2
  const double b = 2.0 * G - 1.0;
  int ii = (int)floor((b - sqrt(b * b - 8.0 * (double)p)) * 0.5);
  if (ii < 0) ii = 0;
  while ((int64_t)ii * (2LL * G - ii - 1) / 2 > p) --ii;
  while ((int64_t)(ii + 1) * (2LL * G - ii - 2) / 2 <= p) ++ii;
  i = ii;
  j = (int)(p - (int64_t)ii * (2LL * G - ii - 1) / 2) + ii + 1;
}

__device__ __forceinline__ void load_gmm(double* dst, const double* w, const double* mu, const double* cov,
                                         int K, int lane) {
  for (int k = lane; k < K; k += 32) {
    const double a = cov[4 * k], b = cov[4 * k + 1], c = cov[4 * k + 3];
    double* d = dst + COMP_STRIDE * k;
    d[0] = w[k]; d[1] = mu[2 * k]; d[2] = mu[2 * k + 1];
    d[3] = a; d[4] = b; d[5] = c;
    d[6] = sqrt(sqrt(a * c - b * cov[4 * k + 2]));  // det^(1/4), distance.py:78
  }
}

// distance.py:71-80 for one (i, j) component pair, times (w1_i + w2_j) (distance.py:34).
__device__ __forceinline__ double hellinger_cost(const double* c1, const double* c2) {
  const double sxx = 0.5 * (c1[3] + c2[3]), sxy = 0.5 * (c1[4] + c2[4]), syy = 0.5 * (c1[5] + c2[5]);
  const double det = sxx * syy - sxy * sxy;
  const double dx = c1[1] - c2[1], dy = c1[2] - c2[2];
  // one reciprocal square root serves both 1/det and 1/sqrt(det) (a division and a square root less per entry;
  // the result moves by an ulp or two against the reference's two separate operations)
  const double rs = rsqrt(det);
  const double quad = (syy * dx * dx - 2.0 * sxy * dx * dy + sxx * dy * dy) * (rs * rs);
  const double first = exp(-quad * 0.125);
  const double second = c1[6] * c2[6] * rs;
  return (c1[0] + c2[0]) * sqrt(fabs(1.0 - first * second));
}

__device__ __forceinline__ unsigned long long sortable_key(double v) {
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b & 0x8000000000000000ULL) ? ~b : (b | 0x8000000000000000ULL);
}

template <int CPL, typename T>
__device__ __forceinline__ T sel(const T (&a)[CPL], int c) {
  if constexpr (CPL == 1) return a[0];
  else return c ? a[1] : a[0];
}
template <int CPL>
__device__ __forceinline__ void set_bit(unsigned (&m)[CPL], int idx) {
  if constexpr (CPL == 1) m[0] |= 1u << idx;
  else { if (idx >> 5) m[1] |= 1u << (idx & 31); else m[0] |= 1u << (idx & 31); }
}
template <int CPL>
__device__ __forceinline__ void set_at(int (&a)[CPL], int idx, int value) {
  if constexpr (CPL == 1) a[0] = value;
  else { if (idx >> 5) a[1] = value; else a[0] = value; }
}
template <int CPL, typename T>
__device__ __forceinline__ T bcast(const T (&a)[CPL], int idx) {
  return __shfl_sync(FULL_MASK, sel<CPL>(a, idx >> 5), idx & 31);
}

// Exact LSAP on a K x K row-major cost matrix in shared memory; returns the optimal cost
// (NaN if the matrix holds non-finite entries — scipy raises ValueError there).
template <int CPL>
__device__ double lsap_warp(const double* C, int K, int lane, int* scratch) {
  double u[CPL], v[CPL], sh[CPL];
  int row_of_col[CPL], col_of_row[CPL], pred[CPL];
#pragma unroll
  for (int c = 0; c < CPL; ++c) { u[c] = 0.0; v[c] = 0.0; row_of_col[c] = -1; col_of_row[c] = -1; }
  // ---- column reduction (the first phase of Jonker-Volgenant): v_j = min_i C_ij keeps the duals feasible
  //      with u = 0; a column takes its arg-min row if no lower column asked for it.  Only the rows left
  //      unassigned need an augmenting path. -----------------------------------------------------------------
  for (int r = lane; r < K; r += 32) scratch[r] = INT_MAX;
  __syncwarp();
  bool invalid = false;   // NaN / -inf entries: scipy raises "matrix contains invalid numeric entries"
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    const int j = lane + 32 * c;
    int arg = -1;
    if (j < K) {
      double best = INFINITY;
      for (int i = 0; i < K; ++i) {
        const double x = C[(size_t)i * K + j];
        invalid |= !(x > -INFINITY);
        if (x < best) { best = x; arg = i; }
      }
      if (arg >= 0) { v[c] = best; atomicMin(&scratch[arg], j); }   // lowest column wins the row
    }
    pred[c] = arg;
  }
  if (__any_sync(FULL_MASK, invalid)) return NAN;
  __syncwarp();
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    const int j = lane + 32 * c;
    if (j < K && pred[c] >= 0 && scratch[pred[c]] == j) row_of_col[c] = pred[c];
    const int r = lane + 32 * c;
    if (r < K && scratch[r] != INT_MAX) col_of_row[c] = scratch[r];
  }
  __syncwarp();

  for (int cur = 0; cur < K; ++cur) {
    if (bcast<CPL>(col_of_row, cur) >= 0) continue;   // already matched by the column reduction
    unsigned sc[CPL], sr[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) { sh[c] = INFINITY; pred[c] = -1; sc[c] = 0u; sr[c] = 0u; }
    int i = cur, sink = -1;
    double min_val = 0.0;
    while (sink < 0) {
      set_bit<CPL>(sr, i);
      const double ui = bcast<CPL>(u, i);
      const double* Ci = C + (size_t)i * K;
      unsigned long long best = ~0ULL;
      int best_c = 0;
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        const int j = lane + 32 * c;
        if (j < K && !((sc[c] >> lane) & 1u)) {
          const double r = min_val + Ci[j] - ui - v[c];
          if (r < sh[c]) { sh[c] = r; pred[c] = i; }
          const unsigned long long key = sortable_key(sh[c]);
          if (key < best) { best = key; best_c = c; }
        }
      }
      const unsigned hi = (unsigned)(best >> 32), lo = (unsigned)best;
      const unsigned mh = __reduce_min_sync(FULL_MASK, hi);
      const unsigned ml = __reduce_min_sync(FULL_MASK, hi == mh ? lo : 0xffffffffu);
      const unsigned who = __ballot_sync(FULL_MASK, hi == mh && lo == ml);
      const int src = __ffs(who) - 1;
      const int jc = __shfl_sync(FULL_MASK, best_c, src);
      const int jstar = src + 32 * jc;
      min_val = bcast<CPL>(sh, jstar);
      if (!(min_val < INFINITY)) return NAN;  // NaN / inf costs: infeasible
      set_bit<CPL>(sc, jstar);
      const int owner = bcast<CPL>(row_of_col, jstar);
      if (owner < 0) sink = jstar; else i = owner;
    }
    // dual update
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int r = lane + 32 * c;
      // shortest-path cost of the column currently assigned to row r
      const int cr = col_of_row[c] < 0 ? 0 : col_of_row[c];
      double sh_cr;
      if constexpr (CPL == 1) {
        sh_cr = __shfl_sync(FULL_MASK, sh[0], cr & 31);
      } else {
        const double s0 = __shfl_sync(FULL_MASK, sh[0], cr & 31);
        const double s1 = __shfl_sync(FULL_MASK, sh[1], cr & 31);
        sh_cr = (cr >> 5) ? s1 : s0;
      }
      if (r == cur) u[c] += min_val;
      else if (r < K && ((sr[c] >> lane) & 1u)) u[c] += min_val - sh_cr;
      if ((sc[c] >> lane) & 1u) v[c] -= min_val - sh[c];
    }
    // augment along the predecessor chain
    int j = sink;
    while (true) {
      const int pi = bcast<CPL>(pred, j);
      if (lane == (j & 31)) set_at<CPL>(row_of_col, j, pi);
      const int prev_j = bcast<CPL>(col_of_row, pi);
      if (lane == (pi & 31)) set_at<CPL>(col_of_row, pi, j);
      j = prev_j;
      if (pi == cur) break;
    }
  }
  double tot = 0.0;
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    const int r = lane + 32 * c;
    if (r < K) tot += C[(size_t)r * K + col_of_row[c]];
  }
  return warp_sum(tot);
}

template <int CPL>
__global__ void __launch_bounds__(DIST_THREADS) gmm_dist_kernel(const DistArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int K = a.K;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t per_warp = (size_t)K * K + 2 * (size_t)COMP_STRIDE * K;
  double* base = reinterpret_cast<double*>(smem_raw) + warp * per_warp;
  double* C = base;
  double* g1 = base + (size_t)K * K;
  double* g2 = g1 + COMP_STRIDE * K;
  const bool ova = a.qw != nullptr;
  const int64_t n = a.pair_end - a.pair_begin;
  int last_i = -1;
  for (int64_t t = (int64_t)blockIdx.x * DIST_WARPS + warp; t < n; t += (int64_t)gridDim.x * DIST_WARPS) {
    const int64_t p = a.pair_begin + t;
    int i, j;
    if (ova) { i = -2; j = (int)p; }
    else pair_to_ij(p, a.G, i, j);
    __syncwarp();
    if (i != last_i) {
      if (ova) load_gmm(g1, a.qw, a.qmu, a.qcov, K, lane);
      else load_gmm(g1, a.w + (size_t)i * K, a.mu + (size_t)i * K * 2, a.cov + (size_t)i * K * 4, K, lane);
      last_i = i;
    }
    load_gmm(g2, a.w + (size_t)j * K, a.mu + (size_t)j * K * 2, a.cov + (size_t)j * K * 4, K, lane);
    __syncwarp();
    for (int e = lane; e < K * K; e += 32) {
      const int ci = e / K, cj = e - ci * K;
      C[e] = hellinger_cost(g1 + COMP_STRIDE * ci, g2 + COMP_STRIDE * cj);
    }
    __syncwarp();
    const double d = lsap_warp<CPL>(C, K, lane, reinterpret_cast<int*>(g2));
    if (lane == 0) a.out[t] = d;
  }
}

// Exact LSAP on a batch of caller-supplied K x K cost matrices (row-major float64), one warp each.
template <int CPL>
__global__ void __launch_bounds__(DIST_THREADS) lsap_batched_kernel(const double* __restrict__ cost, int K,
                                                                    int64_t B, double* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* C = reinterpret_cast<double*>(smem_raw) + (size_t)warp * ((size_t)K * K + K);
  int* scratch = reinterpret_cast<int*>(C + (size_t)K * K);
  for (int64_t t = (int64_t)blockIdx.x * DIST_WARPS + warp; t < B; t += (int64_t)gridDim.x * DIST_WARPS) {
    __syncwarp();
    const double* src = cost + (size_t)t * K * K;
    for (int e = lane; e < K * K; e += 32) C[e] = src[e];
    __syncwarp();
    const double d = lsap_warp<CPL>(C, K, lane, scratch);
    if (lane == 0) out[t] = d;
  }
}

__global__ void fill_square_kernel(const double* __restrict__ pairs, int G, double* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)G * G) return;
  const int i = (int)(idx / G), j = (int)(idx - (int64_t)i * G);
  if (i == j) { out[idx] = 0.0; return; }
  const int lo = i < j ? i : j, hi = i < j ? j : i;
  const int64_t p = (int64_t)lo * (2LL * G - lo - 1) / 2 + (hi - lo - 1);
  out[idx] = pairs[p];
}

int launch_dist(const DistArgs& a, cudaStream_t stream) {
  const int64_t n = a.pair_end - a.pair_begin;
  if (n <= 0) return STM_OK;
  const int K = a.K;
  const size_t smem = DIST_WARPS * ((size_t)K * K + 2 * (size_t)COMP_STRIDE * K) * sizeof(double);
  int dev = 0, sms = 0;
  STM_CHECK_CUDA(cudaGetDevice(&dev));
  STM_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int64_t blocks = (n + DIST_WARPS - 1) / DIST_WARPS;
  const int64_t cap = (int64_t)sms * 16;
  if (blocks > cap) blocks = cap;
  if (K <= 32) {
    STM_CHECK_CUDA(cudaFuncSetAttribute(gmm_dist_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gmm_dist_kernel<1><<<(unsigned)blocks, DIST_THREADS, smem, stream>>>(a);
  } else {
    STM_CHECK_CUDA(cudaFuncSetAttribute(gmm_dist_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gmm_dist_kernel<2><<<(unsigned)blocks, DIST_THREADS, smem, stream>>>(a);
  }
  STM_CHECK_CUDA(cudaGetLastError());
  return STM_OK;
}

}  // namespace
}  // namespace stm

extern "C" int stm_gmm_dist_pairs(const double* w, const double* mu, const double* cov, int32_t G, int32_t K,
                                  int64_t pair_begin, int64_t pair_end, double* out, void* stream) {
  using namespace stm;
  STM_CHECK_ARG(G >= 0 && K >= 1 && K <= 64, "need G >= 0 and 1 <= K <= 64");
  const int64_t total = (int64_t)G * (G - 1) / 2;
  STM_CHECK_ARG(pair_begin >= 0 && pair_begin <= pair_end && pair_end <= total, "pair range outside the triangle");
  if (pair_begin == pair_end) return STM_OK;
  STM_CHECK_ARG(w && mu && cov && out, "null pointer");
  DistArgs a{w, mu, cov, nullptr, nullptr, nullptr, G, K, pair_begin, pair_end, out};
  return launch_dist(a, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int stm_gmm_dist_one_vs_all(const double* qw, const double* qmu, const double* qcov,
                                       const double* w, const double* mu, const double* cov,
                                       int32_t G, int32_t K, double* out, void* stream) {
  using namespace stm;
  STM_CHECK_ARG(G >= 0 && K >= 1 && K <= 64, "need G >= 0 and 1 <= K <= 64");
  if (G == 0) return STM_OK;
  STM_CHECK_ARG(qw && qmu && qcov && w && mu && cov && out, "null pointer");
  DistArgs a{w, mu, cov, qw, qmu, qcov, G, K, 0, (int64_t)G, out};
  return launch_dist(a, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int stm_lsap_batched(const double* cost, int32_t K, int64_t B, double* out, void* stream) {
  using namespace stm;
  STM_CHECK_ARG(K >= 1 && K <= 64 && B >= 0, "need 1 <= K <= 64 and B >= 0");
  if (B == 0) return STM_OK;
  STM_CHECK_ARG(cost && out, "null pointer");
  const size_t smem = DIST_WARPS * ((size_t)K * K + K) * sizeof(double);
  int dev = 0, sms = 0;
  STM_CHECK_CUDA(cudaGetDevice(&dev));
  STM_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int64_t blocks = (B + DIST_WARPS - 1) / DIST_WARPS;
  if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (K <= 32) {
    STM_CHECK_CUDA(cudaFuncSetAttribute(lsap_batched_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lsap_batched_kernel<1><<<(unsigned)blocks, DIST_THREADS, smem, st>>>(cost, K, B, out);
  } else {
    STM_CHECK_CUDA(cudaFuncSetAttribute(lsap_batched_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lsap_batched_kernel<2><<<(unsigned)blocks, DIST_THREADS, smem, st>>>(cost, K, B, out);
  }
  STM_CHECK_CUDA(cudaGetLastError());
  return STM_OK;
}

extern "C" int stm_gmm_dist_fill_square(const double* pairs, int32_t G, double* out_square, void* stream) {
  using namespace stm;
  STM_CHECK_ARG(G >= 0, "G < 0");
  if (G == 0) return STM_OK;
  STM_CHECK_ARG(out_square && (pairs || G == 1), "null pointer");
  const int64_t n = (int64_t)G * G;
  const int threads = 256;
  fill_square_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0,
                       reinterpret_cast<cudaStream_t>(stream)>>>(pairs, G, out_square);
  STM_CHECK_CUDA(cudaGetLastError());
  return STM_OK;
}
