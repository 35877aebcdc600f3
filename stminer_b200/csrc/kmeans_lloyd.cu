// Lloyd iterations of the KMeans step of cluster() on the device — all n_init restarts at once, float64.
//
//   cluster()                 STMiner/Algorithm/algorithm.py:57   KMeans(n_clusters, random_state=0, max_iter=1000, n_init=20)
//   scikit-learn KMeans.fit   cluster/_kmeans.py: _kmeans_single_lloyd (restated here), lloyd_iter_chunked_dense
//
// One CTA per restart.  The start centres come from the host (scikit-learn's own k-means++ with the seeds KMeans.fit draws,
// stminer_b200/Algorithm/algorithm.py); this kernel restates the loop that follows:
//   E: label_i = argmin_j ( ||c_j||^2 - 2 x_i . c_j )          (first minimum, as np.argmin over sklearn's pairwise matrix)
//   M: c_j = mean of the points labelled j                      (an empty cluster sets status 1: the host then falls back)
//   stop: labels unchanged (strict convergence), or sum_j ||c_j_new - c_j||^2 <= tol; after a tol stop the labels are
//         recomputed for the final centres (_kmeans.py "rerun E-step so that predicted labels match cluster centers")
//   inertia = sum_i ||x_i - c_label_i||^2
// Sums run over the points in index order (deterministic).  n <= a few 10^4 points, d <= 64, k <= 64: the embedding of a
// gene set; X (n x d) is read through L2 by every CTA.
#include <float.h>
#include <math.h>

#include "stm_common.cuh"

namespace stm {
namespace {

constexpr int KM_T = 256;
constexpr int KM_MAXK = 64;
constexpr int KM_MAXD = 64;

struct KmArgs {
  const double* X; int32_t n, d, k, n_runs, max_iter; double tol;
  const double* centers_init;   // [n_runs][k][d]
  int32_t* labels_ws;           // [n_runs][2][n] scratch: current / previous labels
  double* out_centers;          // [n_runs][k][d]
  int32_t* out_labels;          // [n_runs][n]
  double* out_inertia; int32_t* out_n_iter; int32_t* out_status;
};

__global__ void __launch_bounds__(KM_T) kmeans_lloyd_kernel(const KmArgs p) {
  extern __shared__ double s_km[];
  const int n = p.n, d = p.d, k = p.k;
  double* c = s_km;                 // [k][d] centres
  double* cn = c + k * d;           // [k][d] new centres
  double* nrm = cn + k * d;         // [k]   squared norms
  double* cnt = nrm + k;            // [k]   cluster sizes
  __shared__ int s_flag[4];         // 0: labels changed, 1: empty cluster
  __shared__ double s_red[KM_T / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int run = blockIdx.x;
  int32_t* lab = p.labels_ws + (size_t)run * 2 * n;
  int32_t* lab_old = lab + n;
  for (int i = tid; i < k * d; i += KM_T) c[i] = p.centers_init[(size_t)run * k * d + i];
  for (int i = tid; i < n; i += KM_T) lab_old[i] = -1;
  if (tid == 0) { s_flag[1] = 0; }
  __syncthreads();

  auto e_step = [&]() {             // labels for the centres in c; returns (block-uniform) whether any label changed
    for (int j = tid; j < k; j += KM_T) {
      double s = 0.0;
      for (int t = 0; t < d; ++t) s += c[j * d + t] * c[j * d + t];
      nrm[j] = s;
    }
    if (tid == 0) s_flag[0] = 0;
    __syncthreads();
    int changed = 0;
    for (int i = tid; i < n; i += KM_T) {
      const double* x = p.X + (size_t)i * d;
      double best = DBL_MAX;
      int bj = 0;
      for (int j = 0; j < k; ++j) {
        double dot = 0.0;
        for (int t = 0; t < d; ++t) dot += x[t] * c[j * d + t];
        const double v = nrm[j] - 2.0 * dot;
        if (v < best) { best = v; bj = j; }
      }
      lab[i] = bj;
      changed |= bj != lab_old[i];
    }
    if (changed) s_flag[0] = 1;
    __syncthreads();
    return s_flag[0] != 0;
  };

  int n_iter = 0, strict = 0;
  for (int it = 0; it < p.max_iter; ++it) {
    const bool changed = e_step();
    // M-step on the NEW labels: sklearn's lloyd_iter does E and M in one pass, then compares the labels
    for (int q = tid; q < k * d; q += KM_T) {
      const int j = q / d, t = q - j * d;
      double s = 0.0, m = 0.0;
      for (int i = 0; i < n; ++i)
        if (lab[i] == j) { s += p.X[(size_t)i * d + t]; m += 1.0; }
      cn[q] = m > 0.0 ? s / m : c[q];
      if (t == 0) { cnt[j] = m; if (m == 0.0) s_flag[1] = 1; }
    }
    __syncthreads();
    double sh = 0.0;
    for (int q = tid; q < k * d; q += KM_T) { const double df = cn[q] - c[q]; sh += df * df; }
    sh = warp_sum(sh);
    if (lane == 0) s_red[warp] = sh;
    __syncthreads();
    double shift = 0.0;
    for (int w = 0; w < KM_T / 32; ++w) shift += s_red[w];
    for (int q = tid; q < k * d; q += KM_T) c[q] = cn[q];
    n_iter = it + 1;
    __syncthreads();
    if (s_flag[1]) break;                               // empty cluster: sklearn relocates it; the host falls back
    if (!changed) { strict = 1; break; }                // labels equal to the previous iteration's
    if (shift <= p.tol) break;
    for (int i = tid; i < n; i += KM_T) lab_old[i] = lab[i];
    __syncthreads();
  }
  if (!strict && !s_flag[1]) e_step();                  // labels that match the final centres
  // inertia and outputs
  double part = 0.0;
  for (int i = tid; i < n; i += KM_T) {
    const double* x = p.X + (size_t)i * d;
    const double* cj = c + lab[i] * d;
    double s = 0.0;
    for (int t = 0; t < d; ++t) { const double df = x[t] - cj[t]; s += df * df; }
    part += s;
    p.out_labels[(size_t)run * n + i] = lab[i];
  }
  part = warp_sum(part);
  __syncthreads();
  if (lane == 0) s_red[warp] = part;
  __syncthreads();
  for (int q = tid; q < k * d; q += KM_T) p.out_centers[(size_t)run * k * d + q] = c[q];
  if (tid == 0) {
    double tot = 0.0;
    for (int w = 0; w < KM_T / 32; ++w) tot += s_red[w];
    p.out_inertia[run] = tot;
    p.out_n_iter[run] = n_iter;
    p.out_status[run] = s_flag[1] ? 1 : 0;
  }
}

}  // namespace
}  // namespace stm

extern "C" int64_t stm_kmeans_workspace_bytes(int32_t n, int32_t n_runs) {
  if (n < 0 || n_runs < 0) return -1;
  return (int64_t)n_runs * 2 * n * 4 + 256;
}

extern "C" int stm_kmeans_lloyd(const double* X, int32_t n, int32_t d, int32_t k, const double* centers_init, int32_t n_runs,
                                int32_t max_iter, double tol, void* workspace, int64_t workspace_bytes,
                                double* out_centers, int32_t* out_labels, double* out_inertia, int32_t* out_n_iter,
                                int32_t* out_status, void* stream) {
  using namespace stm;
  STM_CHECK_ARG(n > 0 && d > 0 && k > 0 && n_runs > 0 && max_iter >= 0, "bad size");
  STM_CHECK_ARG(k <= KM_MAXK && d <= KM_MAXD, "at most 64 clusters and 64 dimensions");
  STM_CHECK_ARG(X && centers_init && workspace && out_centers && out_labels && out_inertia && out_n_iter && out_status, "null pointer");
  STM_CHECK_ARG(workspace_bytes >= stm_kmeans_workspace_bytes(n, n_runs), "workspace too small");
  KmArgs a{X, n, d, k, n_runs, max_iter, tol, centers_init, static_cast<int32_t*>(workspace), out_centers, out_labels,
           out_inertia, out_n_iter, out_status};
  const size_t smem = (size_t)(2 * k * d + 2 * k) * sizeof(double);
  STM_CHECK_CUDA(cudaFuncSetAttribute(kmeans_lloyd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kmeans_lloyd_kernel<<<n_runs, KM_T, smem, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  STM_CHECK_CUDA(cudaGetLastError());
  return STM_OK;
}
