// Device-side staging of an expression matrix into the hot path's layout (sm_100a).
//
// Replaces, for the mixture-fit path, the per-gene host work of the reference — adata[:, gene].X sliced out of the CSR matrix
// for every gene, densified, truncated to int16 and scattered on the (x, y) grid (STMiner/Algorithm/AlgUtils.py:26-36, called
// per gene from STMiner/Algorithm/distribution.py:122) — by ONE stable transposition of the spot-by-gene CSR matrix into the
// gene-major (CSC) arrays the EM kernel reads: selected genes only, values truncated toward zero and wrapped to int16
// (AlgUtils.py:35), zeros dropped, entries of a column in increasing spot index.
//
// Stable parallel transposition without atomics and without a sort: the rows (spots, taken in spot order) are cut into
// blocks; a CTA walks ITS rows in order and, inside a row, its threads take the row's entries — whose columns are distinct —
// so a per-(block, column) counter is only ever touched by one thread at a time, and positions inside a column follow
// (block, row) order = spot order.  Pass 1 counts, a per-column scan over the blocks turns counts into start positions,
// pass 2 writes.  The (block x column) counter row lives in shared memory while its CTA runs.
#include "stm_common.cuh"

namespace stm {
namespace {

constexpr int STAGE_THREADS = 256;

struct StageArgs {
  const int64_t* indptr; const int32_t* indices; const float* data;
  int32_t n_obs; int32_t n_out;
  const int32_t* col_map;      // [n_var] output column of an input column, -1 = not selected
  const int32_t* row_order;    // [n_obs] input row holding spot s (rows in spot order)
  int32_t rows_per_block;
  int32_t* H;                  // [n_blocks][n_out] counts, then start positions inside the column
  const int64_t* col_ptr;      // [n_out + 1] (pass 2)
  void* out_row; void* out_val;
};

// AlgUtils.py:35: coo_matrix(..., dtype=int16) truncates toward zero and wraps modulo 2^16
__device__ __forceinline__ int trunc_wrap_i16(float v) {
  if (!isfinite(v)) return 0;
  const double t = trunc((double)v);
  const long long q = (long long)fmod(t, 65536.0);      // exact for |v| < 2^53
  return (int)(short)(unsigned short)(q & 0xffff);
}

template <bool SCATTER, bool COMPACT, bool SMEM>
__global__ void __launch_bounds__(STAGE_THREADS) stage_rows_kernel(const StageArgs p) {
  extern __shared__ int s_cnt[];
  int* Hrow = p.H + (size_t)blockIdx.x * p.n_out;
  int* cnt = SMEM ? s_cnt : Hrow;
  const int tid = threadIdx.x;
  if (SMEM) {
    for (int c = tid; c < p.n_out; c += STAGE_THREADS) cnt[c] = SCATTER ? Hrow[c] : 0;
    __syncthreads();
  } else if (!SCATTER) {
    for (int c = tid; c < p.n_out; c += STAGE_THREADS) cnt[c] = 0;
    __syncthreads();
  }
  const int s0 = blockIdx.x * p.rows_per_block, s1 = min(p.n_obs, s0 + p.rows_per_block);
  for (int s = s0; s < s1; ++s) {
    const int r = p.row_order[s];
    const int64_t e0 = p.indptr[r], e1 = p.indptr[r + 1];
    for (int64_t e = e0 + tid; e < e1; e += STAGE_THREADS) {
      const int c = p.col_map[p.indices[e]];
      if (c < 0) continue;
      const int v = trunc_wrap_i16(p.data[e]);
      if (v == 0) continue;
      if (!SCATTER) {
        cnt[c] += 1;
      } else {
        const int64_t pos = p.col_ptr[c] + cnt[c];
        cnt[c] += 1;
        if (COMPACT) {
          static_cast<unsigned short*>(p.out_row)[pos] = (unsigned short)s;
          static_cast<short*>(p.out_val)[pos] = (short)v;
        } else {
          static_cast<int32_t*>(p.out_row)[pos] = s;
          static_cast<float*>(p.out_val)[pos] = (float)v;
        }
      }
    }
    __syncthreads();      // rows in order: the next row may hold the same columns
  }
  if (SMEM && !SCATTER) {
    for (int c = tid; c < p.n_out; c += STAGE_THREADS) Hrow[c] = cnt[c];
  }
}

// counts -> start positions inside the column (exclusive scan over the blocks), column totals
__global__ void stage_scan_blocks_kernel(int32_t* H, int n_blocks, int n_out, int64_t* totals) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_out) return;
  int run = 0;
  for (int b = 0; b < n_blocks; ++b) {
    const int t = H[(size_t)b * n_out + c];
    H[(size_t)b * n_out + c] = run;
    run += t;
  }
  totals[c] = run;
}

// col_ptr[0] = 0, col_ptr[c + 1] = sum of totals[0..c]: one CTA, sequential over tiles of 1024
__global__ void __launch_bounds__(1024) stage_col_ptr_kernel(const int64_t* totals, int n_out, int64_t* col_ptr) {
  __shared__ long long s_w[32];
  __shared__ long long s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { s_base = 0; col_ptr[0] = 0; }
  __syncthreads();
  for (int c0 = 0; c0 < n_out; c0 += 1024) {
    const int c = c0 + tid;
    long long v = c < n_out ? totals[c] : 0, incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const long long t = __shfl_up_sync(FULL_MASK, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    long long w = lane < 32 ? s_w[lane] : 0, wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const long long t = __shfl_up_sync(FULL_MASK, wi, o); if (lane >= o) wi += t; }
    const long long before = __shfl_sync(FULL_MASK, wi - w, warp);
    const long long tile_total = __shfl_sync(FULL_MASK, wi, 31);
    const long long base = s_base;
    if (c < n_out) col_ptr[c + 1] = base + before + incl;
    __syncthreads();
    if (tid == 0) s_base = base + tile_total;
    __syncthreads();
  }
}

}  // namespace
}  // namespace stm

extern "C" int32_t stm_stage_blocks(int32_t n_obs) {
  if (n_obs <= 0) return 0;
  const int target = 148 * 8;                      // CTAs in flight: 8 per SM
  const int rows = (n_obs + target - 1) / target;
  return (n_obs + rows - 1) / rows;
}

extern "C" int64_t stm_stage_workspace_bytes(int32_t n_obs, int32_t n_out) {
  if (n_obs < 0 || n_out < 0) return -1;
  return (int64_t)stm_stage_blocks(n_obs) * n_out * 4 + (int64_t)n_out * 8 + 512;
}

extern "C" int stm_stage_csr_columns(const int64_t* indptr, const int32_t* indices, const float* data, int32_t n_obs,
                                     const int32_t* col_map, int32_t n_out, const int32_t* row_order, uint32_t compact,
                                     void* workspace, int64_t workspace_bytes, int64_t out_capacity,
                                     int64_t* out_col_ptr, void* out_row_idx, void* out_val, void* stream) {
  using namespace stm;
  STM_CHECK_ARG(n_obs >= 0 && n_out >= 0, "negative size");
  STM_CHECK_ARG(out_col_ptr, "null col_ptr");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (n_obs == 0 || n_out == 0) {
    STM_CHECK_CUDA(cudaMemsetAsync(out_col_ptr, 0, sizeof(int64_t) * ((size_t)n_out + 1), st));
    return STM_OK;
  }
  STM_CHECK_ARG(indptr && indices && data && col_map && row_order && out_row_idx && out_val, "null pointer");
  STM_CHECK_ARG(!compact || n_obs <= 65536, "the compact encoding holds spot indices below 65,536");
  STM_CHECK_ARG(workspace && workspace_bytes >= stm_stage_workspace_bytes(n_obs, n_out), "workspace too small");
  (void)out_capacity;    // the caller sizes the outputs for every stored entry of the input (an upper bound)
  const int n_blocks = stm_stage_blocks(n_obs);
  StageArgs a{};
  a.indptr = indptr; a.indices = indices; a.data = data; a.n_obs = n_obs; a.n_out = n_out;
  a.col_map = col_map; a.row_order = row_order;
  a.rows_per_block = (n_obs + n_blocks - 1) / n_blocks;
  a.H = static_cast<int32_t*>(workspace);
  int64_t* totals = reinterpret_cast<int64_t*>(static_cast<char*>(workspace) + (((size_t)n_blocks * n_out * 4 + 255) & ~(size_t)255));
  a.col_ptr = out_col_ptr; a.out_row = out_row_idx; a.out_val = out_val;
  const size_t smem = (size_t)n_out * 4;
  const bool use_smem = smem <= 96 * 1024;          // two CTAs per SM keep their counter rows on chip
  auto run = [&](auto kern) -> int {
    if (use_smem) STM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<n_blocks, STAGE_THREADS, use_smem ? smem : 0, st>>>(a);
    STM_CHECK_CUDA(cudaGetLastError());
    return STM_OK;
  };
  int rc = use_smem ? run(stage_rows_kernel<false, false, true>) : run(stage_rows_kernel<false, false, false>);
  if (rc != STM_OK) return rc;
  stage_scan_blocks_kernel<<<(n_out + 255) / 256, 256, 0, st>>>(a.H, n_blocks, n_out, totals);
  STM_CHECK_CUDA(cudaGetLastError());
  stage_col_ptr_kernel<<<1, 1024, 0, st>>>(totals, n_out, out_col_ptr);
  STM_CHECK_CUDA(cudaGetLastError());
  if (compact) rc = use_smem ? run(stage_rows_kernel<true, true, true>) : run(stage_rows_kernel<true, true, false>);
  else rc = use_smem ? run(stage_rows_kernel<true, false, true>) : run(stage_rows_kernel<true, false, false>);
  return rc;
}

// ---------------------------------------------------------------------------------------------------------------------
// Pattern vote: the accumulation loop of SPFinder._genes_to_pattern (STMiner/SPFinder.py:459-489) for ALL labels at once.
// For every gene g of label l (its int16-truncated counts w_gs, AlgUtils.py:35):  total_l[s] += w_gs * 100 / sum_s w_gs  and
// votes_l[s] += 1 for every spot with w_gs != 0.  Spot-major (CSR) input: one warp per spot, lane k takes the row's entries
// k, k + 32, ... in order and keeps a private accumulator per label; a fixed shuffle tree then combines the lanes — the
// floating-point result is the same on every run (no atomics on the sums).  The per-gene totals are exact integer atomics.
// ---------------------------------------------------------------------------------------------------------------------
namespace stm {
namespace {

constexpr int PAT_THREADS = 128;      // 4 warps = 4 spots per CTA
constexpr int PAT_MAX_LABELS = 64;

__global__ void __launch_bounds__(256) pattern_gene_sums_kernel(const int64_t* indptr, const int32_t* indices, const float* data,
                                                                int n_obs, const int32_t* col_label, unsigned long long* gene_sum) {
  const int row = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n_obs) return;
  for (int64_t e = indptr[row] + lane; e < indptr[row + 1]; e += 32) {
    const int c = indices[e];
    if (col_label[c] < 0) continue;
    const int w = trunc_wrap_i16(data[e]);
    if (w != 0) atomicAdd(&gene_sum[c], (unsigned long long)(long long)w);      // two's complement: exact for negative counts too
  }
}

__global__ void __launch_bounds__(PAT_THREADS) pattern_vote_kernel(const int64_t* indptr, const int32_t* indices, const float* data,
                                                                   int n_obs, const int32_t* col_label, int n_labels,
                                                                   const long long* gene_sum, double* out_total, int32_t* out_votes) {
  extern __shared__ double s_acc[];                       // [warps][n_labels][32] sums, then votes as int
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * (PAT_THREADS / 32) + warp;
  double* acc = s_acc + (size_t)warp * n_labels * 32;
  int* vot = reinterpret_cast<int*>(s_acc + (size_t)(PAT_THREADS / 32) * n_labels * 32) + (size_t)warp * n_labels * 32;
  for (int l = 0; l < n_labels; ++l) { acc[l * 32 + lane] = 0.0; vot[l * 32 + lane] = 0; }
  __syncwarp();
  if (row < n_obs) {
    for (int64_t e = indptr[row] + lane; e < indptr[row + 1]; e += 32) {
      const int c = indices[e];
      const int l = col_label[c];
      if (l < 0) continue;
      const int w = trunc_wrap_i16(data[e]);
      if (w == 0) continue;
      const long long gs = gene_sum[c];
      if (gs != 0) acc[l * 32 + lane] += (double)w * (100.0 / (double)gs);      // scale_array: exp * (100 / sum), 0 if sum == 0
      vot[l * 32 + lane] += 1;
    }
  }
  __syncwarp();
  if (row < n_obs) {
    for (int l = 0; l < n_labels; ++l) {
      double v = acc[l * 32 + lane];
      int k = vot[l * 32 + lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { v += __shfl_xor_sync(FULL_MASK, v, o); k += __shfl_xor_sync(FULL_MASK, k, o); }
      if (lane == 0) { out_total[(size_t)l * n_obs + row] = v; out_votes[(size_t)l * n_obs + row] = k; }
    }
  }
}

}  // namespace
}  // namespace stm

extern "C" int stm_pattern_vote(const int64_t* indptr, const int32_t* indices, const float* data, int32_t n_obs, int32_t n_var,
                                const int32_t* col_label, int32_t n_labels, int64_t* gene_sum,
                                double* out_total, int32_t* out_votes, void* stream) {
  using namespace stm;
  STM_CHECK_ARG(n_obs >= 0 && n_var >= 0 && n_labels >= 0, "negative size");
  STM_CHECK_ARG(n_labels <= PAT_MAX_LABELS, "at most 64 labels per call");
  if (n_obs == 0 || n_labels == 0) return STM_OK;
  STM_CHECK_ARG(indptr && indices && data && col_label && gene_sum && out_total && out_votes, "null pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  STM_CHECK_CUDA(cudaMemsetAsync(gene_sum, 0, sizeof(int64_t) * (size_t)n_var, st));
  pattern_gene_sums_kernel<<<(n_obs + 7) / 8, 256, 0, st>>>(indptr, indices, data, n_obs, col_label,
                                                            reinterpret_cast<unsigned long long*>(gene_sum));
  STM_CHECK_CUDA(cudaGetLastError());
  const size_t smem = (size_t)(PAT_THREADS / 32) * n_labels * 32 * (sizeof(double) + sizeof(int));
  STM_CHECK_CUDA(cudaFuncSetAttribute(pattern_vote_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pattern_vote_kernel<<<(n_obs + PAT_THREADS / 32 - 1) / (PAT_THREADS / 32), PAT_THREADS, smem, st>>>(
      indptr, indices, data, n_obs, col_label, n_labels, reinterpret_cast<const long long*>(gene_sum), out_total, out_votes);
  STM_CHECK_CUDA(cudaGetLastError());
  return STM_OK;
}
