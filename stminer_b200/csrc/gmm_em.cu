// Batched count-weighted 2-D Gaussian-mixture EM for sm_100a — one CTA per gene, EM loop on chip.
//
// Arithmetic restated from scikit-learn's GaussianMixture (the reference calls it at
// STMiner/Algorithm/distribution.py:126-127): E-step log-density + logsumexp, M-step sufficient
// statistics, stop on |delta lower bound| < tol.  The count-replicated coordinate list of
// distribution.py:369-373 is replaced by integer weights on distinct spots.
//
// Kernel shape:
//   * thread group of GS lanes owns one spot at a time; lane `sub` of the group owns KT components
//     (K <= KT*GS), whose 6 E-step coefficients and 6 M-step accumulators live in registers;
//   * log2-domain whitened residuals  t = s * U^T (x - mu),  s = sqrt(log2(e)/2):
//         lp_k = c_k - t0^2 - t1^2          (5 FFMA per spot-component)
//     the M-step accumulates moments of t (n, t0, t1, t0t0, t0t1, t1t1), which stay O(1) for every
//     spot that carries responsibility, so FP32 accumulation has no cancellation; the finalize step
//     maps them back through the (float-rounded, hence exactly consistent) affine map in FP64;
//   * the lower bound and the cross-warp reduction are accumulated in FP64 so the sklearn stopping
//     rule fires on the same iteration as the float64 reference.
#include <float.h>
#include <math.h>

#include "stm_common.cuh"

namespace stm {
namespace {

constexpr int EM_THREADS = 128;
constexpr int EM_WARPS = EM_THREADS / 32;
constexpr int EM_CHUNK_PER_GROUP = 2048;  // spots per group between FP64 flushes of the accumulators

struct EmArgs {
  const int64_t* col_ptr;
  const int32_t* row_idx;
  const float* val;
  const int32_t* spot_x;
  const int32_t* spot_y;
  const int32_t* gene_order;
  const double* init_w;
  const double* init_mu;
  const double* init_cov;
  double* out_w;
  double* out_mu;
  double* out_cov;
  int32_t* out_n_iter;
  int32_t* out_converged;
  double* out_lb;
  int32_t* out_status;
  int32_t G, K;
  int32_t smem_spots;  // staging capacity (spots) of this launch
  int32_t max_iter;
  uint32_t flags;
  double reg_covar, tol;
  uint64_t seed;
};

// int16 truncation semantics of AlgUtils.py:35 (coo_matrix(..., dtype=np.int16)).
__device__ __forceinline__ int trunc_i16(float v) {
  int iv = (int)v;  // toward zero
  return (int)(short)iv;
}

// AlgUtils.py:16-17,22-23: round(max(w - mean_nonzero, 0)) when remove_low_exp_spots.
__device__ __forceinline__ int gene_weight(float v, bool remove_low, double mean_nz) {
  int w = trunc_i16(v);
  if (remove_low) {
    double d = (double)w - mean_nz;
    w = d > 0.0 ? (int)rint(d) : 0;
  }
  return w;
}

template <int N, int OFF, int NMAX>
__device__ __forceinline__ void halving_reduce(float (&v)[NMAX], int lane, int& shift) {
  if constexpr (OFF < 32) {
    constexpr int H = N / 2;
    const bool up = (lane & OFF) != 0;
#pragma unroll
    for (int i = 0; i < H; ++i) {
      const float send = up ? v[i] : v[i + H];
      const float keep = up ? v[i + H] : v[i];
      v[i] = keep + __shfl_xor_sync(FULL_MASK, send, OFF);
    }
    if (up) shift += H;
    halving_reduce<H, OFF * 2, NMAX>(v, lane, shift);
  }
}

constexpr int round_up(int a, int b) { return (a + b - 1) / b * b; }

template <int KT, int GS>
struct EmLayout {
  static constexpr int KP = KT * GS;
  static constexpr int NSTAT = KP * 6;
  // doubles
  static constexpr int D_STAT = 0;                // [NSTAT]
  static constexpr int D_MU = D_STAT + NSTAT;     // [KP*2]
  static constexpr int D_COV = D_MU + 2 * KP;     // [KP*3]
  static constexpr int D_W = D_COV + 3 * KP;      // [KP]
  static constexpr int D_LL = D_W + KP;           // [EM_WARPS]
  static constexpr int D_MISC = D_LL + EM_WARPS;  // [8]
  static constexpr int D_END = D_MISC + 8;
  // floats (after doubles)
  static constexpr int F_PAR = 0;                              // [KP*8]
  static constexpr int F_RED = F_PAR + KP * 8;                 // [EM_WARPS*NSTAT]
  static constexpr int F_SPOT = round_up(F_RED + EM_WARPS * NSTAT, 4);  // 3 * cap
  static constexpr size_t fixed_bytes() { return (size_t)D_END * 8 + (size_t)F_SPOT * 4 + 64; }
};

// Component k: (mu, cov, w) in s_d -> E-step coefficients in s_par.  Returns false if cov is not PD.
template <typename L>
__device__ __forceinline__ bool refresh_component(int k, double* s_d, float* s_par) {
  const double mx = s_d[L::D_MU + 2 * k], my = s_d[L::D_MU + 2 * k + 1];
  const double a = s_d[L::D_COV + 3 * k], b = s_d[L::D_COV + 3 * k + 1], c = s_d[L::D_COV + 3 * k + 2];
  const double w = s_d[L::D_W + k];
  bool ok = a > 0.0;
  const double l00 = sqrt(a);
  const double l10 = b / l00;
  const double t = c - l10 * l10;
  ok = ok && (t > 0.0) && isfinite(t) && isfinite(a);
  const double l11 = sqrt(t);
  const double s = 0.84932180028801904272;  // sqrt(log2(e) / 2)
  const float a00 = (float)(s / l00);
  const float a11 = (float)(s / l11);
  const float a01 = (float)(-l10 * s / (l00 * l11));
  const float b0 = (float)(-(double)a00 * mx);
  const float b1 = (float)(-((double)a01 * mx + (double)a11 * my));
  // log2( w / (2 pi l00 l11) )
  const float c2 = (float)log2(w / (6.283185307179586476925 * l00 * l11));
  float* q = s_par + 8 * k;
  q[0] = a00; q[1] = b0; q[2] = a01; q[3] = a11; q[4] = b1; q[5] = c2; q[6] = 0.f; q[7] = 0.f;
  return ok;
}

struct StagedSpots {
  const float* x; const float* y; const float* w;
  __device__ __forceinline__ void get(int i, float& xo, float& yo, float& wo) const {
    xo = x[i]; yo = y[i]; wo = w[i];
  }
};

struct StreamedSpots {
  const int32_t* row_idx; const float* val; const int32_t* spot_x; const int32_t* spot_y;
  int ox, oy; bool remove_low; double mean_nz;
  __device__ __forceinline__ void get(int i, float& xo, float& yo, float& wo) const {
    const int r = __ldg(row_idx + i);
    const int w = gene_weight(__ldg(val + i), remove_low, mean_nz);
    xo = (float)(__ldg(spot_x + r) - ox);
    yo = (float)(__ldg(spot_y + r) - oy);
    wo = w > 0 ? (float)w : 0.f;
  }
};

// One fused E+M sweep over spots [i_begin, i_end) of this CTA.
template <int KT, int GS, int S, typename Src>
__device__ __forceinline__ void em_sweep(const Src& src, int i_begin, int i_end, int grp, int sub,
                                         const float (&par)[KT][6], float* acc, double& ll) {
  constexpr int NG = EM_THREADS / GS;
  // the trip count must be uniform across the warp: the group shuffles below use the full mask
  for (int base = i_begin; base < i_end; base += NG * S) {
    float x[S], y[S], w[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int i = base + grp + s * NG;
      x[s] = 0.f; y[s] = 0.f; w[s] = 0.f;
      if (i < i_end) src.get(i, x[s], y[s], w[s]);
    }
    float t0[S][KT], t1[S][KT], lp[S][KT], m[S], sum[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
      m[s] = -INFINITY;
#pragma unroll
      for (int j = 0; j < KT; ++j) {
        t0[s][j] = fmaf(x[s], par[j][0], par[j][1]);
        t1[s][j] = fmaf(x[s], par[j][2], fmaf(y[s], par[j][3], par[j][4]));
        lp[s][j] = fmaf(-t0[s][j], t0[s][j], fmaf(-t1[s][j], t1[s][j], par[j][5]));
        m[s] = fmaxf(m[s], lp[s][j]);
      }
    }
#pragma unroll
    for (int s = 0; s < S; ++s) {
#pragma unroll
      for (int o = 1; o < GS; o <<= 1) m[s] = fmaxf(m[s], __shfl_xor_sync(FULL_MASK, m[s], o));
    }
#pragma unroll
    for (int s = 0; s < S; ++s) {
      sum[s] = 0.f;
#pragma unroll
      for (int j = 0; j < KT; ++j) {
        lp[s][j] = ex2_approx(lp[s][j] - m[s]);
        sum[s] += lp[s][j];
      }
    }
#pragma unroll
    for (int s = 0; s < S; ++s) {
#pragma unroll
      for (int o = 1; o < GS; o <<= 1) sum[s] += __shfl_xor_sync(FULL_MASK, sum[s], o);
    }
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const float scale = w[s] * rcp_approx(sum[s]);
      if (sub == 0) ll = fma((double)w[s], (double)(m[s] + lg2_approx(sum[s])), ll);
#pragma unroll
      for (int j = 0; j < KT; ++j) {
        const float r = lp[s][j] * scale;
        const float rt0 = r * t0[s][j];
        const float rt1 = r * t1[s][j];
        acc[j * 6 + 0] += r;
        acc[j * 6 + 1] += rt0;
        acc[j * 6 + 2] += rt1;
        acc[j * 6 + 3] = fmaf(rt0, t0[s][j], acc[j * 6 + 3]);
        acc[j * 6 + 4] = fmaf(rt0, t1[s][j], acc[j * 6 + 4]);
        acc[j * 6 + 5] = fmaf(rt1, t1[s][j], acc[j * 6 + 5]);
      }
    }
  }
}

template <int KT, int GS, int S, int MINB>
__global__ void __launch_bounds__(EM_THREADS, MINB) gmm_em_kernel(const EmArgs p) {
  using L = EmLayout<KT, GS>;
  constexpr int KP = L::KP;
  constexpr int NSTAT = L::NSTAT;
  constexpr int NPAD = round_up(KT * 6, 32 / GS);
  constexpr int NF = NPAD * GS / 32;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_d = reinterpret_cast<double*>(smem_raw);
  float* s_f = reinterpret_cast<float*>(s_d + L::D_END);
  float* s_par = s_f + L::F_PAR;
  float* s_red = s_f + L::F_RED;
  float* s_x = s_f + L::F_SPOT;
  float* s_y = s_x + p.smem_spots;
  float* s_wt = s_y + p.smem_spots;
  __shared__ int s_cnt[EM_WARPS];
  __shared__ int s_flag[4];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = tid % GS, grp = tid / GS;
  const int K = p.K;
  const int g = p.gene_order ? p.gene_order[blockIdx.x] : (int)blockIdx.x;
  const int64_t p0 = p.col_ptr[g], p1 = p.col_ptr[g + 1];
  const int ncol = (int)(p1 - p0);
  const int32_t* rows = p.row_idx + p0;
  const float* vals = p.val + p0;
  const bool remove_low = (p.flags & STM_FIT_REMOVE_LOW_EXP_SPOTS) != 0;

  // ---- pre-pass: weights, N, origin -------------------------------------------------------
  double mean_nz = 0.0;
  if (remove_low) {
    long long sum = 0, cnt = 0;
    for (int i = tid; i < ncol; i += EM_THREADS) {
      const int w = trunc_i16(__ldg(vals + i));
      sum += w; cnt += (w != 0);
    }
    sum = warp_sum(sum); cnt = warp_sum(cnt);
    if (lane == 0) { s_d[L::D_LL + warp] = (double)sum; s_d[L::D_STAT + warp] = (double)cnt; }
    __syncthreads();
    double ts = 0, tc = 0;
    for (int w = 0; w < EM_WARPS; ++w) { ts += s_d[L::D_LL + w]; tc += s_d[L::D_STAT + w]; }
    mean_nz = tc > 0 ? ts / tc : 0.0;
    __syncthreads();
  }
  long long npos = 0, sw = 0, swx = 0, swy = 0;
  for (int i = tid; i < ncol; i += EM_THREADS) {
    const int w = gene_weight(__ldg(vals + i), remove_low, mean_nz);
    if (w > 0) {
      const int r = __ldg(rows + i);
      npos += 1; sw += w;
      swx += (long long)w * __ldg(p.spot_x + r);
      swy += (long long)w * __ldg(p.spot_y + r);
    }
  }
  npos = warp_sum(npos); sw = warp_sum(sw); swx = warp_sum(swx); swy = warp_sum(swy);
  if (lane == 0) {
    s_d[L::D_STAT + 4 * warp + 0] = (double)npos; s_d[L::D_STAT + 4 * warp + 1] = (double)sw;
    s_d[L::D_STAT + 4 * warp + 2] = (double)swx;  s_d[L::D_STAT + 4 * warp + 3] = (double)swy;
  }
  __syncthreads();
  double t_npos = 0, t_sw = 0, t_swx = 0, t_swy = 0;
  for (int w = 0; w < EM_WARPS; ++w) {
    t_npos += s_d[L::D_STAT + 4 * w + 0]; t_sw += s_d[L::D_STAT + 4 * w + 1];
    t_swx += s_d[L::D_STAT + 4 * w + 2];  t_swy += s_d[L::D_STAT + 4 * w + 3];
  }
  __syncthreads();
  const double N_total = t_sw;

  double* o_w = p.out_w + (size_t)g * K;
  double* o_mu = p.out_mu + (size_t)g * K * 2;
  double* o_cov = p.out_cov + (size_t)g * K * 4;
  if ((int)t_npos < K) {  // distribution.py:125,129-130 -> gene dropped
    for (int i = tid; i < K; i += EM_THREADS) o_w[i] = 0.0;
    for (int i = tid; i < 2 * K; i += EM_THREADS) o_mu[i] = 0.0;
    for (int i = tid; i < 4 * K; i += EM_THREADS) o_cov[i] = 0.0;
    if (tid == 0) {
      p.out_n_iter[g] = 0; p.out_converged[g] = 0; p.out_lb[g] = 0.0; p.out_status[g] = STM_GENE_DROPPED;
    }
    return;
  }
  const int ox = (int)rint(t_swx / t_sw), oy = (int)rint(t_swy / t_sw);

  // ---- stage distinct positive spots into shared memory (ordered compaction) ----------------
  const bool staged = ncol <= p.smem_spots;
  int n_spots = ncol;
  if (staged) {
    int n_st = 0;
    for (int base = 0; base < ncol; base += EM_THREADS) {
      const int i = base + tid;
      int w = 0, r = 0;
      if (i < ncol) { w = gene_weight(__ldg(vals + i), remove_low, mean_nz); r = __ldg(rows + i); }
      const bool keep = w > 0;
      const unsigned bal = __ballot_sync(FULL_MASK, keep);
      if (lane == 0) s_cnt[warp] = __popc(bal);
      __syncthreads();
      int off = n_st, tot = 0;
#pragma unroll
      for (int ww = 0; ww < EM_WARPS; ++ww) { const int c = s_cnt[ww]; if (ww < warp) off += c; tot += c; }
      if (keep) {
        const int idx = off + __popc(bal & ((1u << lane) - 1u));
        s_x[idx] = (float)(__ldg(p.spot_x + r) - ox);
        s_y[idx] = (float)(__ldg(p.spot_y + r) - oy);
        s_wt[idx] = (float)w;
      }
      n_st += tot;
      __syncthreads();
    }
    n_spots = n_st;
  }

  // ---- initial parameters -------------------------------------------------------------------
  if (tid == 0) { s_flag[0] = 0; s_flag[1] = 0; }
  if (tid < KP) {
    const int k = tid;
    if (k < K) {
      const double* iw = p.init_w + (size_t)g * K;
      const double* imu = p.init_mu + (size_t)g * K * 2;
      const double* icov = p.init_cov + (size_t)g * K * 4;
      s_d[L::D_W + k] = iw[k];
      s_d[L::D_MU + 2 * k] = imu[2 * k] - ox;
      s_d[L::D_MU + 2 * k + 1] = imu[2 * k + 1] - oy;
      s_d[L::D_COV + 3 * k] = icov[4 * k];
      s_d[L::D_COV + 3 * k + 1] = 0.5 * (icov[4 * k + 1] + icov[4 * k + 2]);
      s_d[L::D_COV + 3 * k + 2] = icov[4 * k + 3];
    }
  }
  __syncthreads();
  if (tid < KP) {
    const int k = tid;
    if (k < K) {
      if (!refresh_component<L>(k, s_d, s_par)) atomicOr(&s_flag[1], 1);
    } else {
      float* q = s_par + 8 * k;
      q[0] = 0.f; q[1] = 0.f; q[2] = 0.f; q[3] = 0.f; q[4] = 0.f; q[5] = -1e30f; q[6] = 0.f; q[7] = 0.f;
    }
  }
  __syncthreads();

  // ---- EM loop (sklearn mixture/_base.py:265-278) ---------------------------------------------
  const StagedSpots staged_src{s_x, s_y, s_wt};
  const StreamedSpots stream_src{rows, vals, p.spot_x, p.spot_y, ox, oy, remove_low, mean_nz};
  constexpr int NG = EM_THREADS / GS;
  constexpr int CHUNK = EM_CHUNK_PER_GROUP * NG;
  double lb_prev = -INFINITY, lb = 0.0;
  int n_iter = 0, converged = 0;
  bool failed = s_flag[1] != 0;

  while (!failed && n_iter < p.max_iter) {
    ++n_iter;
    float par[KT][6];
#pragma unroll
    for (int j = 0; j < KT; ++j) {
      const float4 q0 = *reinterpret_cast<const float4*>(s_par + 8 * (sub * KT + j));
      const float2 q1 = *reinterpret_cast<const float2*>(s_par + 8 * (sub * KT + j) + 4);
      par[j][0] = q0.x; par[j][1] = q0.y; par[j][2] = q0.z; par[j][3] = q0.w; par[j][4] = q1.x; par[j][5] = q1.y;
    }
    for (int i = tid; i < NSTAT; i += EM_THREADS) s_d[L::D_STAT + i] = 0.0;
    double ll = 0.0;
    for (int c0 = 0; c0 < n_spots; c0 += CHUNK) {
      const int c1 = min(n_spots, c0 + CHUNK);
      float acc[NPAD];
#pragma unroll
      for (int i = 0; i < NPAD; ++i) acc[i] = 0.f;
      if (staged) em_sweep<KT, GS, S>(staged_src, c0, c1, grp, sub, par, acc, ll);
      else        em_sweep<KT, GS, S>(stream_src, c0, c1, grp, sub, par, acc, ll);
      int shift = 0;
      halving_reduce<NPAD, GS, NPAD>(acc, lane, shift);
#pragma unroll
      for (int i = 0; i < NF; ++i)
        if (shift + i < KT * 6) s_red[warp * NSTAT + sub * KT * 6 + shift + i] = acc[i];
      __syncthreads();
      for (int i = tid; i < NSTAT; i += EM_THREADS) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < EM_WARPS; ++w) t += (double)s_red[w * NSTAT + i];
        s_d[L::D_STAT + i] += t;
      }
      __syncthreads();
    }
    ll = warp_sum(ll);
    if (lane == 0) s_d[L::D_LL + warp] = ll;
    __syncthreads();
    // M-step finalize: whitened moments -> (w, mu, cov) (sklearn _gaussian_mixture.py:282-320,883-901)
    if (tid < K) {
      const int k = tid;
      const double* st = s_d + L::D_STAT + 6 * k;
      const double n = st[0] + 10.0 * DBL_EPSILON;  // _gaussian_mixture.py:312
      double tot = 0.0;                             // weights_ /= weights_.sum(), :898
      for (int j = 0; j < K; ++j) tot += s_d[L::D_STAT + 6 * j] + 10.0 * DBL_EPSILON;
      const double inv_n = 1.0 / n;
      const double m0 = st[1] * inv_n, m1 = st[2] * inv_n;
      const double C00 = st[3] * inv_n - m0 * m0;
      const double C01 = st[4] * inv_n - m0 * m1;
      const double C11 = st[5] * inv_n - m1 * m1;
      const float* q = s_par + 8 * k;
      const double a00 = q[0], b0 = q[1], a01 = q[2], a11 = q[3], b1 = q[4];
      const double pp = 1.0 / a00, rr = 1.0 / a11, qq = -a01 * pp * rr;
      const double mx = (m0 - b0) * pp;
      const double my = (m1 - b1 - a01 * mx) * rr;
      s_d[L::D_MU + 2 * k] = mx;
      s_d[L::D_MU + 2 * k + 1] = my;
      s_d[L::D_COV + 3 * k] = pp * pp * C00 + p.reg_covar;
      s_d[L::D_COV + 3 * k + 1] = pp * (qq * C00 + rr * C01);
      s_d[L::D_COV + 3 * k + 2] = qq * qq * C00 + 2.0 * qq * rr * C01 + rr * rr * C11 + p.reg_covar;
      s_d[L::D_W + k] = n / tot;
      if (!refresh_component<L>(k, s_d, s_par)) atomicOr(&s_flag[1], 1);
    }
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < EM_WARPS; ++w) t += s_d[L::D_LL + w];
      // log2 domain -> natural log; mean over the N replicated rows (_base.py:332)
      s_d[L::D_MISC + 1] = t * 0.69314718055994530942 / N_total;
    }
    __syncthreads();
    lb = s_d[L::D_MISC + 1];
    failed = s_flag[1] != 0;
    const double change = lb - lb_prev;
    lb_prev = lb;
    if (fabs(change) < p.tol) { converged = 1; break; }
  }

  // ---- outputs --------------------------------------------------------------------------------
  if (failed) {
    for (int i = tid; i < K; i += EM_THREADS) o_w[i] = 0.0;
    for (int i = tid; i < 2 * K; i += EM_THREADS) o_mu[i] = 0.0;
    for (int i = tid; i < 4 * K; i += EM_THREADS) o_cov[i] = 0.0;
  } else if (tid < K) {
    const int k = tid;
    o_w[k] = s_d[L::D_W + k];
    o_mu[2 * k] = s_d[L::D_MU + 2 * k] + ox;
    o_mu[2 * k + 1] = s_d[L::D_MU + 2 * k + 1] + oy;
    o_cov[4 * k] = s_d[L::D_COV + 3 * k];
    o_cov[4 * k + 1] = s_d[L::D_COV + 3 * k + 1];
    o_cov[4 * k + 2] = s_d[L::D_COV + 3 * k + 1];
    o_cov[4 * k + 3] = s_d[L::D_COV + 3 * k + 2];
  }
  if (tid == 0) {
    p.out_n_iter[g] = n_iter;
    p.out_converged[g] = failed ? 0 : converged;
    p.out_lb[g] = failed ? 0.0 : lb;
    p.out_status[g] = failed ? STM_GENE_ILL_CONDITIONED : STM_GENE_OK;
  }
}

template <int KT, int GS, int S, int MINB>
int launch_em(const EmArgs& a_in, cudaStream_t stream) {
  using L = EmLayout<KT, GS>;
  EmArgs a = a_in;
  auto kern = gmm_em_kernel<KT, GS, S, MINB>;
  int dev = 0, max_optin = 0;
  STM_CHECK_CUDA(cudaGetDevice(&dev));
  STM_CHECK_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  // MINB CTAs per SM share 227 KB (+1 KB reserved per CTA)
  size_t per_cta = ((size_t)228 * 1024) / MINB - 1024 - 64;
  if (per_cta > (size_t)max_optin) per_cta = max_optin;
  const size_t fixed = L::fixed_bytes();
  if (per_cta < fixed + 12 * 256) {
    set_last_error("stm_gmm_fit_batched: shared-memory budget too small");
    return STM_ERR_UNSUPPORTED;
  }
  a.smem_spots = (int32_t)((per_cta - fixed) / 12) & ~3;
  const size_t smem = fixed + (size_t)a.smem_spots * 12;
  STM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<a.G, EM_THREADS, smem, stream>>>(a);
  STM_CHECK_CUDA(cudaGetLastError());
  return STM_OK;
}

}  // namespace
}  // namespace stm

extern "C" int stm_gmm_fit_batched(const int64_t* col_ptr, const int32_t* row_idx, const float* val,
                                   const int32_t* spot_x, const int32_t* spot_y, int32_t n_spots,
                                   int32_t G, int32_t K, const int32_t* gene_order,
                                   const double* init_w, const double* init_mu, const double* init_cov,
                                   double reg_covar, double tol, int32_t max_iter, uint32_t flags, uint64_t seed,
                                   double* out_w, double* out_mu, double* out_cov,
                                   int32_t* out_n_iter, int32_t* out_converged, double* out_lower_bound,
                                   int32_t* out_status, void* stream) {
  using namespace stm;
  STM_CHECK_ARG(G >= 0, "G < 0");
  if (G == 0) return STM_OK;
  STM_CHECK_ARG(col_ptr && row_idx && val && spot_x && spot_y, "null input pointer");
  STM_CHECK_ARG(n_spots > 0, "n_spots <= 0");
  STM_CHECK_ARG(K >= 1 && K <= 64, "K must be in [1, 64]");
  STM_CHECK_ARG(max_iter >= 1, "max_iter < 1");
  STM_CHECK_ARG(out_w && out_mu && out_cov && out_n_iter && out_converged && out_lower_bound && out_status,
                "null output pointer");
  if (flags & STM_FIT_DEVICE_INIT) {
    set_last_error("stm_gmm_fit_batched: STM_FIT_DEVICE_INIT not available in this build");
    return STM_ERR_UNSUPPORTED;
  }
  STM_CHECK_ARG(init_w && init_mu && init_cov, "init_* required without STM_FIT_DEVICE_INIT");
  EmArgs a{};
  a.col_ptr = col_ptr; a.row_idx = row_idx; a.val = val; a.spot_x = spot_x; a.spot_y = spot_y;
  a.gene_order = gene_order; a.init_w = init_w; a.init_mu = init_mu; a.init_cov = init_cov;
  a.out_w = out_w; a.out_mu = out_mu; a.out_cov = out_cov; a.out_n_iter = out_n_iter;
  a.out_converged = out_converged; a.out_lb = out_lower_bound; a.out_status = out_status;
  a.G = G; a.K = K; a.max_iter = max_iter; a.flags = flags; a.reg_covar = reg_covar; a.tol = tol; a.seed = seed;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (K <= 5) return launch_em<5, 1, 2, 4>(a, st);
  if (K <= 10) return launch_em<5, 2, 2, 4>(a, st);
  if (K <= 20) return launch_em<5, 4, 2, 4>(a, st);
  if (K <= 32) return launch_em<8, 4, 1, 3>(a, st);
  if (K <= 40) return launch_em<5, 8, 2, 4>(a, st);
  return launch_em<8, 8, 1, 3>(a, st);
}
