// Metric multidimensional scaling by SMACOF (Guttman majorisation) on a precomputed dissimilarity matrix —
// the step that follows the gene x gene distance matrix in the reference:
//   cluster()            STMiner/Algorithm/algorithm.py:29-61   (MDS(n_components, dissimilarity='precomputed'))
//   sklearn smacof       scikit-learn 1.3.0 manifold/_mds.py    (_smacof_single: the iteration restated here)
//
// One pass per SMACOF iteration, everything float64 like the reference:
//   dis_ij  = ||X_i - X_j||                      (0 -> 1e-5, _mds.py "dis[dis == 0] = 1e-5")
//   stress  = sum_ij (dis_ij - D_ij)^2 / 2       (of the configuration the pass READS)
//   X'_i    = 1/n * sum_j (D_ij / dis_ij) (X_i - X_j)        == 1/n * B(X) X, the Guttman transform
// The n x n ratio matrix B is never materialised: a warp owns one row i, its lanes stride over j with the X_j tile
// staged in shared memory ([c][j] layout, conflict-free), accumulate sum_j ratio * (X_i - X_j) in registers and
// reduce once per row.  Row partials of the stress are summed in a fixed order by a one-block kernel that also
// applies the stopping rule on the device, so the whole iteration loop is enqueued without host synchronisation;
// passes after convergence return immediately.
//
// Two stopping rules (the reference pins scikit-learn 1.3.0; the image has 1.9):
//   rule 0 (sklearn <= 1.6):  s_it = stress(X_it) / sum_i ||X_{it+1,i}||;  stop when it > 0 and s_{it-1} - s_it < eps;
//                             returns X_{it+1}, stress(X_it), n_iter = it + 1
//   rule 1 (sklearn >= 1.7):  s_it = stress(X_{it+1});  stop when it > 0 and (s_{it-1} - s_it) / (sum dis(X_{it+1})^2 / 2) < eps;
//                             returns X_{it+1}, s_it, n_iter = it + 1
#include <math.h>

#include "stm_common.cuh"

namespace stm {
namespace {

#ifndef STM_MDS_T
#define STM_MDS_T 256
#endif
#ifndef STM_MDS_MINB
#define STM_MDS_MINB 2   // 2 CTAs per SM (128 registers, a few spills) beat 1 CTA at 166 registers: 0.27 vs 0.34 ms per pass
#endif
constexpr int MDS_T = STM_MDS_T;  // threads per CTA of the row pass: one row per warp
constexpr int MDS_TJ = 256;      // columns j per shared-memory tile

struct MdsState {
  double prev;        // previous value of the monitored quantity
  double stress;      // raw stress to report
  int has_prev, done, n_iter, result_buf;
};

// X row-major [n][d] -> transposed, zero-padded [DP][n]
__global__ void mds_pack_kernel(const double* __restrict__ X, double* __restrict__ XT, int n, int d, int DP,
                                MdsState* st) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) { st->prev = 0.0; st->stress = 0.0; st->has_prev = 0; st->done = 0; st->n_iter = 0; st->result_buf = 0; }
  if (i >= n) return;
  for (int c = 0; c < DP; ++c) XT[(size_t)c * n + i] = c < d ? X[(size_t)i * d + c] : 0.0;
}

__global__ void mds_unpack_kernel(const double* __restrict__ XT0, const double* __restrict__ XT1, double* __restrict__ X,
                                  int n, int d, const MdsState* st, double* out_stress, int* out_n_iter) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const double* XT = st->result_buf ? XT1 : XT0;
  if (i == 0) { *out_stress = st->stress; *out_n_iter = st->n_iter; }
  if (i >= n) return;
  for (int c = 0; c < d; ++c) X[(size_t)i * d + c] = XT[(size_t)c * n + i];
}

template <int DP>
__global__ void __launch_bounds__(MDS_T, STM_MDS_MINB) mds_rows_kernel(const double* __restrict__ D, const double* __restrict__ Xin,
                                                         double* __restrict__ Xout, int n,
                                                         double* __restrict__ row_stress, double* __restrict__ row_ssd,
                                                         const MdsState* st) {
  if (st->done) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* Xs = reinterpret_cast<double*>(smem_raw);   // [DP][MDS_TJ]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int i = blockIdx.x * (MDS_T / 32) + warp;
  const bool row_ok = i < n;
  double xi[DP], acc[DP];
#pragma unroll
  for (int c = 0; c < DP; ++c) { xi[c] = row_ok ? Xin[(size_t)c * n + i] : 0.0; acc[c] = 0.0; }
  double stress = 0.0, ssd = 0.0;
  const double* Drow = D + (size_t)(row_ok ? i : 0) * n;
  for (int j0 = 0; j0 < n; j0 += MDS_TJ) {
    __syncthreads();
    for (int e = tid; e < DP * MDS_TJ; e += MDS_T) {
      const int c = e / MDS_TJ, t = e % MDS_TJ;
      Xs[e] = (j0 + t < n) ? Xin[(size_t)c * n + j0 + t] : 0.0;
    }
    __syncthreads();
    if (!row_ok) continue;
    // the tile's dissimilarities first: MDS_TJ / 32 independent coalesced loads in flight per lane (the pass is
    // otherwise bound by the latency of this one global load per pair)
    double dj[MDS_TJ / 32];
#pragma unroll
    for (int u = 0; u < MDS_TJ / 32; ++u) {
      const int j = j0 + lane + 32 * u;
      dj[u] = j < n ? __ldg(Drow + j) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < MDS_TJ / 32; ++u) {
      const int t = lane + 32 * u;
      if (j0 + t >= n) break;
      double d2 = 0.0;
#pragma unroll
      for (int c = 0; c < DP; ++c) { const double df = xi[c] - Xs[c * MDS_TJ + t]; d2 = fma(df, df, d2); }
      double dis = sqrt(d2);
      const double e = dis - dj[u];
      stress = fma(e, e, stress);
      ssd += d2;
      if (dis == 0.0) dis = 1e-5;
      const double ratio = dj[u] / dis;
      // the differences are recomputed rather than kept: 2*DP fewer live registers -> two CTAs per SM
#pragma unroll
      for (int c = 0; c < DP; ++c) acc[c] = fma(ratio, xi[c] - Xs[c * MDS_TJ + t], acc[c]);
    }
  }
  if (!row_ok) return;
  const double inv_n = 1.0 / (double)n;
#pragma unroll
  for (int c = 0; c < DP; ++c) {
    const double v = warp_sum(acc[c]);
    if (lane == 0) Xout[(size_t)c * n + i] = v * inv_n;
  }
  stress = warp_sum(stress); ssd = warp_sum(ssd);
  if (lane == 0) { row_stress[i] = stress; row_ssd[i] = ssd; }
}

// fixed-order block sum of v over the threads of a 1024-thread block
__device__ __forceinline__ double block_sum_1024(double v, double* s_w) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < 32; ++w) t += s_w[w];
  return t;
}

__global__ void __launch_bounds__(1024) mds_check_kernel(const double* __restrict__ row_stress,
                                                         const double* __restrict__ row_ssd,
                                                         const double* __restrict__ Xnew, int n, int DP, int pass,
                                                         int rule, double eps, int max_iter, MdsState* st) {
  if (st->done) return;
  __shared__ double s_w[32];
  double a = 0.0, b = 0.0, c = 0.0;
  for (int i = threadIdx.x; i < n; i += 1024) {
    a += row_stress[i]; b += row_ssd[i];
    if (rule == 0) {
      double q = 0.0;
      for (int k = 0; k < DP; ++k) { const double x = Xnew[(size_t)k * n + i]; q = fma(x, x, q); }
      c += sqrt(q);
    }
  }
  const double stress = 0.5 * block_sum_1024(a, s_w);
  const double ssd = block_sum_1024(b, s_w);
  const double nrm = block_sum_1024(c, s_w);
  if (threadIdx.x != 0) return;
  const int in_buf = pass & 1, out_buf = in_buf ^ 1;
  if (rule == 0) {
    const double s = stress / nrm;
    st->stress = stress; st->n_iter = pass + 1; st->result_buf = out_buf;
    if (st->has_prev && (st->prev - s) < eps) st->done = 1;
    st->prev = s; st->has_prev = 1;
    if (pass + 1 >= max_iter) st->done = 1;
  } else {
    if (pass == 0) return;                 // stress(X_0) is not monitored by this rule
    // this pass evaluated X_pass, the configuration sklearn's iteration it = pass - 1 produced
    st->stress = stress; st->n_iter = pass; st->result_buf = in_buf;
    if (st->has_prev && (st->prev - stress) / (0.5 * ssd) < eps) st->done = 1;
    st->prev = stress; st->has_prev = 1;
    if (pass >= max_iter) st->done = 1;
  }
}

constexpr int mds_dp(int d) { return d <= 2 ? 2 : d <= 4 ? 4 : d <= 8 ? 8 : d <= 12 ? 12 : d <= 16 ? 16 : d <= 20 ? 20 : d <= 24 ? 24 : 32; }

template <int DP>
int launch_rows(const double* D, const double* Xin, double* Xout, int n, double* rs, double* rq, const MdsState* st,
                cudaStream_t s) {
  const size_t smem = (size_t)DP * MDS_TJ * sizeof(double);
  auto kern = mds_rows_kernel<DP>;
  STM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(n + MDS_T / 32 - 1) / (MDS_T / 32), MDS_T, smem, s>>>(D, Xin, Xout, n, rs, rq, st);
  STM_CHECK_CUDA(cudaGetLastError());
  return STM_OK;
}

int launch_rows_dp(int DP, const double* D, const double* Xin, double* Xout, int n, double* rs, double* rq,
                   const MdsState* st, cudaStream_t s) {
  switch (DP) {
    case 2: return launch_rows<2>(D, Xin, Xout, n, rs, rq, st, s);
    case 4: return launch_rows<4>(D, Xin, Xout, n, rs, rq, st, s);
    case 8: return launch_rows<8>(D, Xin, Xout, n, rs, rq, st, s);
    case 12: return launch_rows<12>(D, Xin, Xout, n, rs, rq, st, s);
    case 16: return launch_rows<16>(D, Xin, Xout, n, rs, rq, st, s);
    case 20: return launch_rows<20>(D, Xin, Xout, n, rs, rq, st, s);
    case 24: return launch_rows<24>(D, Xin, Xout, n, rs, rq, st, s);
    default: return launch_rows<32>(D, Xin, Xout, n, rs, rq, st, s);
  }
}

}  // namespace
}  // namespace stm

extern "C" int64_t stm_mds_workspace_bytes(int32_t n, int32_t d) {
  if (n <= 0 || d <= 0 || d > 32) return -1;
  const int64_t dp = stm::mds_dp(d);
  return 2 * dp * (int64_t)n * 8 + 2 * (int64_t)n * 8 + 256;
}

extern "C" int stm_mds_smacof(const double* dissim, int32_t n, int32_t d, const double* x_init, int32_t max_iter,
                              double eps, int32_t rule, void* workspace, int64_t workspace_bytes,
                              double* x_out, double* out_stress, int32_t* out_n_iter, void* stream) {
  using namespace stm;
  STM_CHECK_ARG(n >= 1, "n < 1");
  STM_CHECK_ARG(d >= 1 && d <= 32, "n_components must be in [1, 32]");
  STM_CHECK_ARG(max_iter >= 1, "max_iter < 1");
  STM_CHECK_ARG(rule == 0 || rule == 1, "rule must be 0 (sklearn <= 1.6) or 1 (sklearn >= 1.7)");
  STM_CHECK_ARG(dissim && x_init && x_out && out_stress && out_n_iter && workspace, "null pointer");
  STM_CHECK_ARG(workspace_bytes >= stm_mds_workspace_bytes(n, d), "workspace too small");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int DP = mds_dp(d);
  double* XT[2];
  XT[0] = static_cast<double*>(workspace);
  XT[1] = XT[0] + (size_t)DP * n;
  double* rs = XT[1] + (size_t)DP * n;
  double* rq = rs + n;
  MdsState* st = reinterpret_cast<MdsState*>(rq + n);
  mds_pack_kernel<<<(n + 255) / 256, 256, 0, s>>>(x_init, XT[0], n, d, DP, st);
  STM_CHECK_CUDA(cudaGetLastError());
  const int n_pass = rule == 0 ? max_iter : max_iter + 1;
  for (int p = 0; p < n_pass; ++p) {
    const int rc = launch_rows_dp(DP, dissim, XT[p & 1], XT[(p & 1) ^ 1], n, rs, rq, st, s);
    if (rc != STM_OK) return rc;
    mds_check_kernel<<<1, 1024, 0, s>>>(rs, rq, XT[(p & 1) ^ 1], n, DP, p, rule, eps, max_iter, st);
    STM_CHECK_CUDA(cudaGetLastError());
  }
  mds_unpack_kernel<<<(n + 255) / 256, 256, 0, s>>>(XT[0], XT[1], x_out, n, d, st, out_stress, out_n_iter);
  STM_CHECK_CUDA(cudaGetLastError());
  return STM_OK;
}
