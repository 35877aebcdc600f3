// Batched count-weighted 2-D Gaussian-mixture EM for sm_100a — one CTA per gene, EM loop on chip.
//
// Arithmetic restated from scikit-learn's GaussianMixture (the reference calls it at
// STMiner/Algorithm/distribution.py:126-127): E-step log-density + logsumexp, M-step sufficient
// statistics, stop on |delta lower bound| < tol.  The count-replicated coordinate list of
// distribution.py:369-373 is replaced by integer weights on distinct spots.
//
// Kernel shape:
//   * each gene's positive spots are staged ONCE into shared memory as 8-byte records
//     (int16 x, int16 y relative to the gene's integer centroid, float weight);
//   * a thread group of GS lanes owns one spot at a time; lane `sub` of the group owns KT components
//     (K <= KT*GS), whose 6 E-step coefficients and 6 M-step accumulators live in registers;
//   * log2-domain whitened residuals  t = s * U^T (x - mu),  s = sqrt(log2(e)/2):
//         lp_k = c_k - t0^2 - t1^2          (5 FFMA per spot-component)
//     the M-step accumulates moments of t (n, t0, t1, t0t0, t0t1, t1t1), which stay O(1) for every
//     spot that carries responsibility, so FP32 accumulation has no cancellation; the finalize step
//     maps them back through the (float-rounded, hence exactly consistent) affine map in FP64;
//   * the lower bound and the cross-warp reduction are accumulated in FP64 so the sklearn stopping
//     rule fires on the same iteration as the float64 reference;
//   * genes are bucketed by nnz on the host (stm_gmm_fit_plan): each bucket gets a CTA size / shared
//     memory budget that keeps its spots staged on chip; genes beyond the largest budget stream
//     their spots from L2;
//   * optional on-device initialisation (weighted k-means++ by an exponential race and Lloyd on a strided seed
//     subset, then one assignment pass over all spots) replaces the sklearn KMeans the reference runs inside
//     GaussianMixture (mixture/_base.py:119-128).
#include <float.h>
#include <limits.h>
#include <math.h>

#include <algorithm>
#include <map>
#include <numeric>
#include <vector>

#include <cooperative_groups.h>

#include "stm_common.cuh"

namespace cg = cooperative_groups;

namespace stm {
namespace {

constexpr int EM_CHUNK_PER_GROUP = 2048;  // spots per group between FP64 flushes of the accumulators
constexpr int KM_MAX_ITER = 300;          // sklearn KMeans default max_iter
constexpr int KM_LMAX = 6;                // greedy k-means++ candidates: 2 + floor(ln K) <= 6 for K <= 64

#ifndef STM_EM_S_SMALL
#define STM_EM_S_SMALL 2      // spots in flight per thread group, small-gene bucket
#endif
#ifndef STM_EM_MINB_SMALL
#define STM_EM_MINB_SMALL 4   // CTAs (of 128 threads) per SM, small-gene bucket
#endif
#ifndef STM_EM_PACKED
#define STM_EM_PACKED 1       // FP32x2 packed E/M sweep (FFMA2); 0 = scalar sweep
#endif
#ifndef STM_EM_MAGIC_I2F
#define STM_EM_MAGIC_I2F 0    // int16 -> float by magic-number add instead of I2F (XU pipe)
#endif
#ifndef STM_EM_L2_PREFETCH
#define STM_EM_L2_PREFETCH 0  // prefetch a later gene's column into L2 during the EM loop
#endif
#ifndef STM_EM_T_LARGE
#define STM_EM_T_LARGE 512    // CTA size of the one-CTA large bucket (1 CTA per SM)
#endif
#ifndef STM_EM_S_LARGE
#define STM_EM_S_LARGE 2      // spots in flight per thread group, large bucket
#endif
#ifndef STM_EM_QMOMENT
#define STM_EM_QMOMENT 0      // accumulate sum r*(t0^2+t1^2) instead of sum r*t1^2 (one FMA-pipe operation less per component)
#endif
#ifndef STM_EM_PREFETCH
#define STM_EM_PREFETCH 0     // load the next step's spots before computing the current one (manual software pipelining)
#endif
#ifndef STM_EM_PAR_SMEM
#define STM_EM_PAR_SMEM 0     // 1 = the E-step coefficients of a lane's components are re-read from shared memory (one LDS.64 per
                              // pair and coefficient, shared by the spots in flight) instead of living in 30 registers — room for
                              // a third spot in flight (STM_EM_S_LARGE=3)
#endif
#ifndef STM_EM_NOMAX
#define STM_EM_NOMAX 0        // 1 = E-step without the per-spot max over the components: the log-density offsets are shifted by their
                              // maximum over the components once per iteration (an upper bound of every log density), so ex2 cannot
                              // overflow; a spot whose densities ALL underflow (> ~12 sigma from every component) takes a slow path.
                              // 12 % fewer instructions in the sweep, parity green — and no gain (23.92 vs 23.80 ms on S50): the sweep
                              // is not issue-bound, see profiles/r2_tuning_log.txt
#endif
#ifndef STM_EM_LL_BATCH
#define STM_EM_LL_BATCH 0     // fold the FP32 log-likelihood partial into FP64 every 8 steps instead of every step
#endif
#ifndef STM_KM_DEFAULT_CAND
#define STM_KM_DEFAULT_CAND 2     // greedy k-means++ trials per centre when the flags field is 0
#endif
#ifndef STM_KM_DEFAULT_LLOYD
#define STM_KM_DEFAULT_LLOYD 4    // Lloyd iteration cap when the flags field is 0 (EM refines further)
#endif

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
  int32_t order_begin;  // this launch fits gene_order[order_begin + blockIdx.x]
  int32_t smem_spots;   // staging capacity (spots) of this launch
  int32_t seed_cap;     // k-means seed-set capacity of this launch (0 without device init)
  int32_t km_max_iter;  // Lloyd iteration cap
  int32_t km_cand;      // greedy k-means++ candidates per centre
  int32_t max_iter;
  uint32_t flags;
  double reg_covar, tol;
  uint64_t seed;
};

// One gene's column in either input encoding: {int32 row, float value} or, with STM_FIT_COMPACT_INPUT,
// {uint16 row, int16 value} — half the bytes over PCIe / HBM for matrices with < 65,536 spots whose values
// the host staging already truncated to int16 (AlgUtils.py:35).
struct Column {
  const void* rows; const void* vals; bool compact;
  __device__ __forceinline__ int row(int i) const {
    return compact ? (int)__ldg(static_cast<const unsigned short*>(rows) + i)
                   : __ldg(static_cast<const int32_t*>(rows) + i);
  }
  __device__ __forceinline__ int trunc_value(int i) const {   // int16 truncation of AlgUtils.py:35
    if (compact) return (int)__ldg(static_cast<const short*>(vals) + i);
    const int iv = (int)__ldg(static_cast<const float*>(vals) + i);   // toward zero
    return (int)(short)iv;
  }
  // AlgUtils.py:16-17,22-23: round(max(w - mean_nonzero, 0)) when remove_low_exp_spots
  __device__ __forceinline__ int weight(int i, bool remove_low, double mean_nz) const {
    int w = trunc_value(i);
    if (remove_low) {
      const double d = (double)w - mean_nz;
      w = d > 0.0 ? (int)rint(d) : 0;
    }
    return w;
  }
};

__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
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

// CL = a gene split over a thread-block cluster: every CTA keeps the same layout plus the cluster-wide totals
// (D_TOT) and a small exchange area (D_XCH) that the other CTAs of the cluster read through DSMEM.
template <int KT, int GS, int T, bool CL = false>
struct EmLayout {
  static constexpr int KT_ = KT;
  static constexpr int KP = KT * GS;
  static constexpr int NSTAT = KP * 6;
  static constexpr int WARPS = T / 32;
  // doubles
  static constexpr int D_STAT = 0;               // [NSTAT]
  static constexpr int D_MU = D_STAT + NSTAT;    // [KP*2]
  static constexpr int D_COV = D_MU + 2 * KP;    // [KP*3]
  static constexpr int D_W = D_COV + 3 * KP;     // [KP]
  static constexpr int D_LL = D_W + KP;          // [WARPS]   per-warp partials
  static constexpr int D_MISC = D_LL + WARPS;    // [8]
  static constexpr int D_SCR = D_MISC + 8;       // [WARPS * KM_LMAX + 2 * KM_LMAX] k-means++ scratch
  static constexpr int D_TOT = D_SCR + WARPS * KM_LMAX + 2 * KM_LMAX;   // [NSTAT + WARPS] cluster totals (CL only)
  static constexpr int D_XCH = D_TOT + (CL ? NSTAT + WARPS : 0);         // [8] cluster exchange (CL only)
  static constexpr int D_END = round_up(D_XCH + (CL ? 8 : 0), 2);
  // floats (after doubles)
  static constexpr int F_PAR = 0;                                   // [KP*8]
  static constexpr int F_PP = F_PAR + KP * 8;                       // [KP*6] E-step coefficients, pair-interleaved
  static constexpr int F_RED = F_PP + KP * 6;                       // [WARPS*NSTAT]
  static constexpr int F_END = round_up(F_RED + WARPS * NSTAT, 4);  // then float4 seeds[seed_cap], int2 spots[cap]
  static constexpr size_t fixed_bytes() { return (size_t)D_END * 8 + (size_t)F_END * 4 + 64; }
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
  const double i00 = 1.0 / l00, i11 = 1.0 / l11;
  const float a00 = (float)(s * i00);
  const float a11 = (float)(s * i11);
  const float a01 = (float)(-l10 * s * i00 * i11);
  const float b0 = (float)(-(double)a00 * mx);
  const float b1 = (float)(-((double)a01 * mx + (double)a11 * my));
  // log2( w / (2 pi l00 l11) ) — FP32 like the coefficient that stores it
  const float c2 = lg2_approx((float)(w * i00 * i11 * 0.15915494309189533577));
  float* q = s_par + 8 * k;
  q[0] = a00; q[1] = b0; q[2] = a01; q[3] = a11; q[4] = b1; q[5] = c2;
  // pair-interleaved copy for the packed sweep: lane `sub` reads (component 2q, component 2q+1) as one float2
  constexpr int KT = L::KT_;
  const int sub = k / KT, j = k % KT;
  float* pq = s_par + (L::F_PP - L::F_PAR) + ((sub * (KT / 2) + (j >> 1)) * 6) * 2 + (j & 1);
  if (j < 2 * (KT / 2)) { pq[0] = a00; pq[2] = b0; pq[4] = a01; pq[6] = a11; pq[8] = b1; pq[10] = c2; }
  return ok;
}

// STM_EM_NOMAX: shift the K log-density offsets c_k (s_par[8k + 5] and the pair-interleaved copy) by their maximum, kept in
// s_d[D_MISC + 7] and added back when the lower bound is formed.  lp_k(x) = c_k - |t_k(x)|^2 <= c_k, so after the shift every
// ex2 argument is <= 0.  Called by the whole CTA between two barriers.
template <typename L>
__device__ __forceinline__ void shift_log_offsets(int K, double* s_d, float* s_par, int tid) {
  constexpr int KT = L::KT_;
  if (tid < 32) {
    float c = -INFINITY;
    for (int k = tid; k < K; k += 32) c = fmaxf(c, s_par[8 * k + 5]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c = fmaxf(c, __shfl_xor_sync(FULL_MASK, c, o));
    for (int k = tid; k < K; k += 32) {
      const float v = s_par[8 * k + 5] - c;
      s_par[8 * k + 5] = v;
      const int sub = k / KT, j = k % KT;
      float* pq = s_par + (L::F_PP - L::F_PAR) + ((sub * (KT / 2) + (j >> 1)) * 6) * 2 + (j & 1);
      if (j < 2 * (KT / 2)) pq[10] = v;
    }
    if (tid == 0) s_d[L::D_MISC + 7] = (double)c;
  }
}

__device__ __forceinline__ int pack_xy(int x, int y) {
  x = max(-32768, min(32767, x)); y = max(-32768, min(32767, y));
  return (x & 0xffff) | (y << 16);
}
// int16 -> float on the ALU/FMA pipes (magic-number add) instead of I2F, which is an XU (MUFU-pipe)
// instruction — the XU pipe is the sweep's second bottleneck after the FMA pipe.
[[maybe_unused]] __device__ __forceinline__ float small_int_to_float(int v) {   // |v| < 2^22
  return __int_as_float(0x4B400000 + v) - 12582912.0f;
}
#if STM_EM_MAGIC_I2F
__device__ __forceinline__ float unpack_x(int p) { return small_int_to_float((p << 16) >> 16); }
__device__ __forceinline__ float unpack_y(int p) { return small_int_to_float(p >> 16); }
#else
__device__ __forceinline__ float unpack_x(int p) { return (float)(short)(p & 0xffff); }
__device__ __forceinline__ float unpack_y(int p) { return (float)(p >> 16); }
#endif

// spots staged in shared memory: {int16 x | int16 y, float w}
struct PackedSpots {
  unsigned s;   // shared-window address of the record array: the struct travels by value into the non-inlined sweep,
                // where a generic pointer would turn every LDS into a generic LD
  __device__ __forceinline__ explicit PackedSpots(const int2* p) : s((unsigned)__cvta_generic_to_shared(p)) {}
  __device__ __forceinline__ void get(int i, float& xo, float& yo, float& wo) const {
    int2 v;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(s + 8u * (unsigned)i));
    xo = unpack_x(v.x); yo = unpack_y(v.x); wo = __int_as_float(v.y);
  }
};

// k-means seed subset in shared memory: {x, y, w, d2}
struct SeedSpots {
  const float4* s;
  __device__ __forceinline__ void get(int i, float& xo, float& yo, float& wo) const {
    const float4 v = s[i];
    xo = v.x; yo = v.y; wo = v.z;
  }
};

// spots of a gene too long to stage: re-read from global memory (L2) every sweep
struct StreamedSpots {
  Column col; const int32_t* spot_x; const int32_t* spot_y;
  int ox, oy; bool remove_low; double mean_nz;
  __device__ __forceinline__ void get(int i, float& xo, float& yo, float& wo) const {
    const int r = col.row(i);
    const int w = col.weight(i, remove_low, mean_nz);
    xo = (float)(__ldg(spot_x + r) - ox);
    yo = (float)(__ldg(spot_y + r) - oy);
    wo = w > 0 ? (float)w : 0.f;
  }
};

// One fused E+M sweep over spots [i_begin, i_end) of this CTA.
template <int KT, int GS, int S, int T, typename Src>
__device__ __forceinline__ void em_sweep(const Src& src, int i_begin, int i_end, int grp, int sub,
                                         const float (&par)[KT][6], float* acc, double& ll) {
  constexpr int NG = T / GS;
  [[maybe_unused]] float wl_acc = 0.f;
  [[maybe_unused]] int step = 0;
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
    float wl = 0.f;  // sum_s w_s * log2-sum-exp_s of this step (S terms in FP32, then one FP64 add)
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const float scale = w[s] * rcp_approx(sum[s]);
      wl = fmaf(w[s], m[s] + lg2_approx(sum[s]), wl);
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
    // every lane of a group holds the same value; lane sub == 0 is the one summed.  The FP32 partial is
    // folded into FP64 every 8 steps (F2F is an XU instruction).
#if STM_EM_LL_BATCH
    wl_acc += wl;
    if ((++step & 7) == 0) { ll += (double)wl_acc; wl_acc = 0.f; }
#else
    ll += (double)wl;
#endif
  }
  ll += (double)wl_acc;
}

// The same sweep with the components of a lane processed in PAIRS by packed FP32x2 instructions (FFMA2 /
// FADD2 / FMUL2): half the issue slots and half the register-file operand fetches for the FMA-class work,
// which is what bounds the scalar sweep (3-operand FFMA without operand reuse sustains 0.62 of the FFMA
// peak on B200, packed 0.81 — scripts/micro/ffma2_bench.cu).  KT = 2*NP + R (R = 0 or 1 scalar leftover).
template <int KT, int GS, int S, int T, typename Src>
__device__ __forceinline__ void em_sweep_packed(const Src& src, int i_begin, int i_end, int grp, int sub,
                                                const f2 (&pp)[(KT / 2) > 0 ? (KT / 2) : 1][6], const float (&ps)[6],
                                                f2 (&ap)[(KT / 2) > 0 ? (KT / 2) : 1][6], float (&as)[6], double& ll,
                                                [[maybe_unused]] unsigned pp_saddr = 0, [[maybe_unused]] unsigned ps_saddr = 0) {
  constexpr int NG = T / GS;
  constexpr int NP = KT / 2, R = KT % 2;
  constexpr int NPA = NP > 0 ? NP : 1;
  [[maybe_unused]] float wl_acc = 0.f;
  [[maybe_unused]] int step = 0;
#if STM_EM_PREFETCH
  float xn[S], yn[S], wn[S];   // the spots of the NEXT step, loaded while this step computes
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const int i = i_begin + grp + s * NG;
    xn[s] = 0.f; yn[s] = 0.f; wn[s] = 0.f;
    if (i < i_end) src.get(i, xn[s], yn[s], wn[s]);
  }
#endif
  for (int base = i_begin; base < i_end; base += NG * S) {
    float x[S], y[S], w[S];
#if STM_EM_PREFETCH
#pragma unroll
    for (int s = 0; s < S; ++s) { x[s] = xn[s]; y[s] = yn[s]; w[s] = wn[s]; }
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int i = base + NG * S + grp + s * NG;
      xn[s] = 0.f; yn[s] = 0.f; wn[s] = 0.f;
      if (i < i_end) src.get(i, xn[s], yn[s], wn[s]);
    }
#else
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int i = base + grp + s * NG;
      x[s] = 0.f; y[s] = 0.f; w[s] = 0.f;
      if (i < i_end) src.get(i, x[s], y[s], w[s]);
    }
#endif
    f2 t0p[S][NPA], t1p[S][NPA], lpp[S][NPA];
    float t0s[S], t1s[S], lps[S], m[S], sum[S];
#if STM_EM_QMOMENT
    f2 qp[S][NPA];
    float qs[S];
#endif
#if STM_EM_NOMAX
    // the offsets are shifted so that every log density is <= 0 (shift_log_offsets): no max pass, ex2 straight away
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const f2 xx = f2_make(x[s], x[s]), yy = f2_make(y[s], y[s]);
      m[s] = 0.f;
      f2 sp = f2_make(0.f, 0.f);
#pragma unroll
      for (int q = 0; q < NP; ++q) {
        t0p[s][q] = f2_fma(xx, pp[q][0], pp[q][1]);
        t1p[s][q] = f2_fma(xx, pp[q][2], f2_fma(yy, pp[q][3], pp[q][4]));
#if STM_EM_QMOMENT
        qp[s][q] = f2_fma(t0p[s][q], t0p[s][q], f2_mul(t1p[s][q], t1p[s][q]));
        const f2 d = f2_sub(pp[q][5], qp[s][q]);
#else
        const f2 d = f2_sub(pp[q][5], f2_fma(t0p[s][q], t0p[s][q], f2_mul(t1p[s][q], t1p[s][q])));
#endif
        lpp[s][q] = f2_make(ex2_approx(d.x), ex2_approx(d.y));
        sp = f2_add(sp, lpp[s][q]);
      }
      sum[s] = sp.x + sp.y;
      if constexpr (R == 1) {
        t0s[s] = fmaf(x[s], ps[0], ps[1]);
        t1s[s] = fmaf(x[s], ps[2], fmaf(y[s], ps[3], ps[4]));
#if STM_EM_QMOMENT
        qs[s] = fmaf(t0s[s], t0s[s], t1s[s] * t1s[s]);
        lps[s] = ex2_approx(ps[5] - qs[s]);
#else
        lps[s] = ex2_approx(fmaf(-t0s[s], t0s[s], fmaf(-t1s[s], t1s[s], ps[5])));
#endif
        sum[s] += lps[s];
      }
    }
#pragma unroll
    for (int s = 0; s < S; ++s) {
#pragma unroll
      for (int o = 1; o < GS; o <<= 1) sum[s] += __shfl_xor_sync(FULL_MASK, sum[s], o);
    }
#pragma unroll
    for (int s = 0; s < S; ++s) {
      // every density of a spot underflowed (farther than ~12 sigma from all components): redo that spot with the max
      // of its log densities, as the reference's logsumexp does.  Uniform over the warp; practically never taken.
      const bool tiny = sum[s] < 1e-30f && w[s] > 0.f;
      if (__any_sync(FULL_MASK, tiny)) {
        float mx = -INFINITY;
#pragma unroll
        for (int q = 0; q < NP; ++q) {
          const f2 d = f2_sub(pp[q][5], f2_fma(t0p[s][q], t0p[s][q], f2_mul(t1p[s][q], t1p[s][q])));
          mx = fmaxf(mx, fmaxf(d.x, d.y));
        }
        if constexpr (R == 1) mx = fmaxf(mx, fmaf(-t0s[s], t0s[s], fmaf(-t1s[s], t1s[s], ps[5])));
#pragma unroll
        for (int o = 1; o < GS; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL_MASK, mx, o));
        if (!tiny) mx = 0.f;                 // the group's lanes agree on `tiny` (sum and w are group-uniform)
        float sm2 = 0.f;
#pragma unroll
        for (int q = 0; q < NP; ++q) {
          const f2 d = f2_sub(pp[q][5], f2_fma(t0p[s][q], t0p[s][q], f2_mul(t1p[s][q], t1p[s][q])));
          lpp[s][q] = f2_make(ex2_approx(d.x - mx), ex2_approx(d.y - mx));
          sm2 += lpp[s][q].x + lpp[s][q].y;
        }
        if constexpr (R == 1) { lps[s] = ex2_approx(fmaf(-t0s[s], t0s[s], fmaf(-t1s[s], t1s[s], ps[5])) - mx); sm2 += lps[s]; }
#pragma unroll
        for (int o = 1; o < GS; o <<= 1) sm2 += __shfl_xor_sync(FULL_MASK, sm2, o);
        sum[s] = sm2; m[s] = mx;
      }
      if (!(w[s] > 0.f)) sum[s] = 1.f;       // padding lanes (weight 0): keep rcp finite
    }
#elif STM_EM_PAR_SMEM
    // coefficients from shared memory, pair-major: six LDS.64 per pair serve all spots in flight
#pragma unroll
    for (int s = 0; s < S; ++s) m[s] = -INFINITY;
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      f2 c[6];
#pragma unroll
      for (int k = 0; k < 6; ++k)
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(c[k].x), "=f"(c[k].y) : "r"(pp_saddr + 8u * (unsigned)(q * 6 + k)));
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const f2 xx = f2_make(x[s], x[s]), yy = f2_make(y[s], y[s]);
        t0p[s][q] = f2_fma(xx, c[0], c[1]);
        t1p[s][q] = f2_fma(xx, c[2], f2_fma(yy, c[3], c[4]));
        lpp[s][q] = f2_sub(c[5], f2_fma(t0p[s][q], t0p[s][q], f2_mul(t1p[s][q], t1p[s][q])));
        m[s] = fmaxf(m[s], fmaxf(lpp[s][q].x, lpp[s][q].y));
      }
    }
    if constexpr (R == 1) {
      float c[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(c[k]) : "r"(ps_saddr + 4u * (unsigned)k));
#pragma unroll
      for (int s = 0; s < S; ++s) {
        t0s[s] = fmaf(x[s], c[0], c[1]);
        t1s[s] = fmaf(x[s], c[2], fmaf(y[s], c[3], c[4]));
        lps[s] = fmaf(-t0s[s], t0s[s], fmaf(-t1s[s], t1s[s], c[5]));
        m[s] = fmaxf(m[s], lps[s]);
      }
    }
#pragma unroll
    for (int s = 0; s < S; ++s) {
#pragma unroll
      for (int o = 1; o < GS; o <<= 1) m[s] = fmaxf(m[s], __shfl_xor_sync(FULL_MASK, m[s], o));
    }
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const f2 mm = f2_make(m[s], m[s]);
      f2 sp = f2_make(0.f, 0.f);
#pragma unroll
      for (int q = 0; q < NP; ++q) {
        const f2 d = f2_sub(lpp[s][q], mm);
        lpp[s][q] = f2_make(ex2_approx(d.x), ex2_approx(d.y));
        sp = f2_add(sp, lpp[s][q]);
      }
      sum[s] = sp.x + sp.y;
      if constexpr (R == 1) { lps[s] = ex2_approx(lps[s] - m[s]); sum[s] += lps[s]; }
    }
#pragma unroll
    for (int s = 0; s < S; ++s) {
#pragma unroll
      for (int o = 1; o < GS; o <<= 1) sum[s] += __shfl_xor_sync(FULL_MASK, sum[s], o);
    }
#else
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const f2 xx = f2_make(x[s], x[s]), yy = f2_make(y[s], y[s]);
      m[s] = -INFINITY;
#pragma unroll
      for (int q = 0; q < NP; ++q) {
        t0p[s][q] = f2_fma(xx, pp[q][0], pp[q][1]);
        t1p[s][q] = f2_fma(xx, pp[q][2], f2_fma(yy, pp[q][3], pp[q][4]));
#if STM_EM_QMOMENT
        qp[s][q] = f2_fma(t0p[s][q], t0p[s][q], f2_mul(t1p[s][q], t1p[s][q]));
        lpp[s][q] = f2_sub(pp[q][5], qp[s][q]);
#else
        lpp[s][q] = f2_sub(pp[q][5], f2_fma(t0p[s][q], t0p[s][q], f2_mul(t1p[s][q], t1p[s][q])));
#endif
        m[s] = fmaxf(m[s], fmaxf(lpp[s][q].x, lpp[s][q].y));
      }
      if constexpr (R == 1) {
        t0s[s] = fmaf(x[s], ps[0], ps[1]);
        t1s[s] = fmaf(x[s], ps[2], fmaf(y[s], ps[3], ps[4]));
#if STM_EM_QMOMENT
        qs[s] = fmaf(t0s[s], t0s[s], t1s[s] * t1s[s]);
        lps[s] = ps[5] - qs[s];
#else
        lps[s] = fmaf(-t0s[s], t0s[s], fmaf(-t1s[s], t1s[s], ps[5]));
#endif
        m[s] = fmaxf(m[s], lps[s]);
      }
    }
#pragma unroll
    for (int s = 0; s < S; ++s) {
#pragma unroll
      for (int o = 1; o < GS; o <<= 1) m[s] = fmaxf(m[s], __shfl_xor_sync(FULL_MASK, m[s], o));
    }
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const f2 mm = f2_make(m[s], m[s]);
      f2 sp = f2_make(0.f, 0.f);
#pragma unroll
      for (int q = 0; q < NP; ++q) {
        const f2 d = f2_sub(lpp[s][q], mm);
        lpp[s][q] = f2_make(ex2_approx(d.x), ex2_approx(d.y));
        sp = f2_add(sp, lpp[s][q]);
      }
      sum[s] = sp.x + sp.y;
      if constexpr (R == 1) { lps[s] = ex2_approx(lps[s] - m[s]); sum[s] += lps[s]; }
    }
#pragma unroll
    for (int s = 0; s < S; ++s) {
#pragma unroll
      for (int o = 1; o < GS; o <<= 1) sum[s] += __shfl_xor_sync(FULL_MASK, sum[s], o);
    }
#endif
    float wl = 0.f;
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const float scale = w[s] * rcp_approx(sum[s]);
      wl = fmaf(w[s], m[s] + lg2_approx(sum[s]), wl);
      const f2 ss = f2_make(scale, scale);
#pragma unroll
      for (int q = 0; q < NP; ++q) {
        const f2 r = f2_mul(lpp[s][q], ss);
        const f2 rt0 = f2_mul(r, t0p[s][q]);
        ap[q][0] = f2_add(ap[q][0], r);
        ap[q][1] = f2_add(ap[q][1], rt0);
        ap[q][3] = f2_fma(rt0, t0p[s][q], ap[q][3]);
        ap[q][4] = f2_fma(rt0, t1p[s][q], ap[q][4]);
#if STM_EM_QMOMENT
        // sum r*t1 by FMA, and sum r*(t0^2 + t1^2) — q is already there from the E-step — in the slot of
        // sum r*t1^2 (recovered as a difference when the accumulators are flushed): 8 instead of 9 operations
        ap[q][2] = f2_fma(r, t1p[s][q], ap[q][2]);
        ap[q][5] = f2_fma(r, qp[s][q], ap[q][5]);
#else
        const f2 rt1 = f2_mul(r, t1p[s][q]);
        ap[q][2] = f2_add(ap[q][2], rt1);
        ap[q][5] = f2_fma(rt1, t1p[s][q], ap[q][5]);
#endif
      }
      if constexpr (R == 1) {
        const float r = lps[s] * scale;
        const float rt0 = r * t0s[s];
        as[0] += r;
        as[1] += rt0;
        as[3] = fmaf(rt0, t0s[s], as[3]);
        as[4] = fmaf(rt0, t1s[s], as[4]);
#if STM_EM_QMOMENT
        as[2] = fmaf(r, t1s[s], as[2]);
        as[5] = fmaf(r, qs[s], as[5]);
#else
        const float rt1 = r * t1s[s];
        as[2] += rt1;
        as[5] = fmaf(rt1, t1s[s], as[5]);
#endif
      }
    }
#if STM_EM_LL_BATCH
    wl_acc += wl;
    if ((++step & 7) == 0) { ll += (double)wl_acc; wl_acc = 0.f; }
#else
    ll += (double)wl;
#endif
  }
  ll += (double)wl_acc;
}

// Assignment sweep with moments: every spot goes to its nearest centre (ties: lowest component
// index); the owning lane accumulates n, dx, dy, dxdx, dxdy, dydy relative to that centre.
template <int KT, int GS, int T, typename Src>
__device__ __forceinline__ void assign_moments_sweep(const Src& src, int i_begin, int i_end, int grp, int sub,
                                                     const float (&cx)[KT], const float (&cy)[KT], float* acc) {
  constexpr int NG = T / GS;
  for (int base = i_begin; base < i_end; base += NG) {
    const int i = base + grp;
    float x = 0.f, y = 0.f, w = 0.f;
    if (i < i_end) src.get(i, x, y, w);
    float dx[KT], dy[KT];
    float best = INFINITY;
    int jb = 0;
#pragma unroll
    for (int j = 0; j < KT; ++j) {
      dx[j] = x - cx[j]; dy[j] = y - cy[j];
      const float d = fmaf(dx[j], dx[j], dy[j] * dy[j]);
      if (d < best) { best = d; jb = j; }
    }
    int kb = sub * KT + jb;
#pragma unroll
    for (int o = 1; o < GS; o <<= 1) {
      const float ob = __shfl_xor_sync(FULL_MASK, best, o);
      const int ok = __shfl_xor_sync(FULL_MASK, kb, o);
      if (ob < best || (ob == best && ok < kb)) { best = ob; kb = ok; }
    }
    const bool mine = (kb / KT) == sub;
#pragma unroll
    for (int j = 0; j < KT; ++j) {
      const float wj = (mine && jb == j) ? w : 0.f;
      const float wdx = wj * dx[j];
      acc[j * 6 + 0] += wj;
      acc[j * 6 + 1] += wdx;
      acc[j * 6 + 2] = fmaf(wj, dy[j], acc[j * 6 + 2]);
      acc[j * 6 + 3] = fmaf(wdx, dx[j], acc[j * 6 + 3]);
      acc[j * 6 + 4] = fmaf(wdx, dy[j], acc[j * 6 + 4]);
      acc[j * 6 + 5] = fmaf(wj * dy[j], dy[j], acc[j * 6 + 5]);
    }
  }
}

// Lloyd iteration sweep (no moments): nearest centre = argmax_k 2 c_k.x - |c_k|^2; the component index
// rides in the 6 low mantissa bits of the score so the argmax is a plain float max (ties inside 2^-17
// relative go to the higher index — harmless for an initialiser).  Accumulates n, sum w x, sum w y.
template <int KT, int GS, int T, typename Src>
__device__ __forceinline__ void lloyd_sweep(const Src& src, int i_begin, int i_end, int grp, int sub,
                                            const float (&ax)[KT], const float (&ay)[KT], const float (&bb)[KT],
                                            float* acc) {
  constexpr int NG = T / GS;
  for (int base = i_begin; base < i_end; base += NG) {
    const int i = base + grp;
    float x = 0.f, y = 0.f, w = 0.f;
    if (i < i_end) src.get(i, x, y, w);
    float best = -INFINITY;
#pragma unroll
    for (int j = 0; j < KT; ++j) {
      const float sc = fmaf(x, ax[j], fmaf(y, ay[j], bb[j]));
      best = fmaxf(best, __int_as_float((__float_as_int(sc) & ~63) | (sub * KT + j)));
    }
#pragma unroll
    for (int o = 1; o < GS; o <<= 1) best = fmaxf(best, __shfl_xor_sync(FULL_MASK, best, o));
    const int jm = (__float_as_int(best) & 63) - sub * KT;  // in [0, KT) only on the owning lane
#pragma unroll
    for (int j = 0; j < KT; ++j) {
      const float wj = (jm == j) ? w : 0.f;
      acc[j * 6 + 0] += wj;
      acc[j * 6 + 1] = fmaf(wj, x, acc[j * 6 + 1]);
      acc[j * 6 + 2] = fmaf(wj, y, acc[j * 6 + 2]);
    }
  }
}

// Per-thread accumulators -> warp halving reduction -> cross-warp FP64 sum added into s_d[D_STAT..].
// Ends with a __syncthreads(); s_d[D_STAT..] is then visible to every thread.
template <int KT, int GS, int T, int NPAD>
__device__ __forceinline__ void flush_stats(float (&acc)[NPAD], double* s_d, float* s_red, int tid) {
  using L = EmLayout<KT, GS, T>;
  constexpr int NF = NPAD * GS / 32;
  const int lane = tid & 31, warp = tid >> 5, sub = tid % GS;
  int shift = 0;
  halving_reduce<NPAD, GS, NPAD>(acc, lane, shift);
#pragma unroll
  for (int i = 0; i < NF; ++i)
    if (shift + i < KT * 6) s_red[warp * L::NSTAT + sub * KT * 6 + shift + i] = acc[i];
  __syncthreads();
  for (int i = tid; i < L::NSTAT; i += T) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < L::WARPS; ++w) t += (double)s_red[w * L::NSTAT + i];
    s_d[L::D_STAT + i] += t;
  }
  __syncthreads();
}

// One chunk of the fused E+M sweep: coefficients from shared memory -> registers, sweep over spots [c0, c1) of this
// CTA, accumulators folded into s_d[D_STAT..] (ends with a __syncthreads()).  Returns this thread's share of
// sum_i w_i * log2-sum-exp_i.  Deliberately NOT inlined: compiled on its own, the instruction schedule of the hot loop
// does not depend on the surrounding kernel (inlined, ptxas interleaved the two in-flight spots differently from one
// kernel variant to the next, worth 3-6 % of the sweep).
template <int KT, int GS, int S, int T, typename Src>
__device__ __noinline__ double em_chunk(const Src src, int c0, int c1, double* s_d, float* s_f) {
  using L = EmLayout<KT, GS, T>;   // the float offsets and D_STAT do not depend on the cluster flag
  constexpr int NPAD = round_up(KT * 6, 32 / GS);
  const int tid = threadIdx.x, sub = tid % GS, grp = tid / GS;
  float* s_par = s_f + L::F_PAR;
  float* s_red = s_f + L::F_RED;
  double ll = 0.0;
  float acc[NPAD];
#pragma unroll
  for (int i = 0; i < NPAD; ++i) acc[i] = 0.f;
#if STM_EM_PACKED
  constexpr int NP = KT / 2, NPA = NP > 0 ? NP : 1;
  f2 pp[NPA][6];
  float ps[6];
#if STM_EM_PAR_SMEM
#pragma unroll
  for (int q = 0; q < NPA; ++q) {
#pragma unroll
    for (int c = 0; c < 6; ++c) pp[q][c] = f2_make(0.f, 0.f);      // unused: the sweep reads the coefficients from shared memory
  }
#pragma unroll
  for (int c = 0; c < 6; ++c) ps[c] = 0.f;
  const unsigned pp_saddr = (unsigned)__cvta_generic_to_shared(reinterpret_cast<const float2*>(s_f + L::F_PP) + (sub * NP) * 6);
  const unsigned ps_saddr = (unsigned)__cvta_generic_to_shared(s_par + 8 * (sub * KT + KT - 1));
#else
#pragma unroll
  for (int q = 0; q < NP; ++q) {
    const float2* qa = reinterpret_cast<const float2*>(s_f + L::F_PP) + (sub * NP + q) * 6;
#pragma unroll
    for (int c = 0; c < 6; ++c) pp[q][c] = qa[c];
  }
#pragma unroll
  for (int c = 0; c < 6; ++c) ps[c] = s_par[8 * (sub * KT + KT - 1) + c];
  const unsigned pp_saddr = 0, ps_saddr = 0;
#endif
  f2 ap[NPA][6];
  float as[6];
#pragma unroll
  for (int q = 0; q < NPA; ++q) {
#pragma unroll
    for (int c = 0; c < 6; ++c) ap[q][c] = f2_make(0.f, 0.f);
  }
#pragma unroll
  for (int c = 0; c < 6; ++c) as[c] = 0.f;
  em_sweep_packed<KT, GS, S, T>(src, c0, c1, grp, sub, pp, ps, ap, as, ll, pp_saddr, ps_saddr);
#pragma unroll
  for (int q = 0; q < NP; ++q) {
#pragma unroll
    for (int c = 0; c < 6; ++c) { acc[(2 * q) * 6 + c] = ap[q][c].x; acc[(2 * q + 1) * 6 + c] = ap[q][c].y; }
  }
  if constexpr (KT % 2 == 1) {
#pragma unroll
    for (int c = 0; c < 6; ++c) acc[(KT - 1) * 6 + c] = as[c];
  }
#if STM_EM_QMOMENT
#pragma unroll
  for (int j = 0; j < KT; ++j) acc[j * 6 + 5] -= acc[j * 6 + 3];   // sum r*t1^2 = sum r*q - sum r*t0^2
#endif
#else
  float par[KT][6];
#pragma unroll
  for (int j = 0; j < KT; ++j) {
    const float4 q0 = *reinterpret_cast<const float4*>(s_par + 8 * (sub * KT + j));
    const float2 q1 = *reinterpret_cast<const float2*>(s_par + 8 * (sub * KT + j) + 4);
    par[j][0] = q0.x; par[j][1] = q0.y; par[j][2] = q0.z; par[j][3] = q0.w; par[j][4] = q1.x; par[j][5] = q1.y;
  }
  em_sweep<KT, GS, S, T>(src, c0, c1, grp, sub, par, acc, ll);
#endif
  flush_stats<KT, GS, T, NPAD>(acc, s_d, s_red, tid);
  return ll;
}

// Cluster-wide reduction of a few per-CTA scalars through DSMEM: every CTA publishes v[0..N) in its exchange
// area, reads the areas of all ranks in rank order (so every CTA ends with bit-identical totals) and combines
// entries [0, NSUM) by + and the rest by min.
template <int N, int NSUM>
__device__ __forceinline__ void cluster_combine(cg::cluster_group& cl, double* s_xch, double (&v)[N], int csize) {
  if (threadIdx.x == 0) {
#pragma unroll
    for (int n = 0; n < N; ++n) s_xch[n] = v[n];
  }
  cl.sync();
#pragma unroll
  for (int n = 0; n < N; ++n) v[n] = n < NSUM ? 0.0 : INFINITY;
  for (int r = 0; r < csize; ++r) {
    const double* rx = cl.map_shared_rank(s_xch, r);
#pragma unroll
    for (int n = 0; n < N; ++n) v[n] = n < NSUM ? v[n] + rx[n] : fmin(v[n], rx[n]);
  }
  cl.sync();   // the exchange area may be rewritten after this point
}

// CL = true: the gene is split over the CTAs of a thread-block cluster (launch attribute, size 2..16).  CTA r of
// the cluster owns a contiguous slice of the gene's column — staged in ITS shared memory — and the per-iteration
// sufficient statistics (6K values + the log-likelihood) are summed across the cluster through distributed shared
// memory, in rank order, so that every CTA holds bit-identical parameters and takes the same branches.
template <int KT, int GS, int S, int T, int MINB, bool CL>
__global__ void __launch_bounds__(T, MINB) gmm_em_kernel(const EmArgs p) {
  using L = EmLayout<KT, GS, T, CL>;
  constexpr int KP = L::KP;
  constexpr int NSTAT = L::NSTAT;
  constexpr int WARPS = L::WARPS;
  constexpr int NPAD = round_up(KT * 6, 32 / GS);
  constexpr int NG = T / GS;
  constexpr int CHUNK = EM_CHUNK_PER_GROUP * NG;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_d = reinterpret_cast<double*>(smem_raw);
  float* s_f = reinterpret_cast<float*>(s_d + L::D_END);
  float* s_par = s_f + L::F_PAR;
  float* s_red = s_f + L::F_RED;
  float4* s_seed = reinterpret_cast<float4*>(s_f + L::F_END);
  int2* s_spot = reinterpret_cast<int2*>(s_seed + p.seed_cap);
  __shared__ int s_cnt[WARPS];
  __shared__ int s_flag[4];
  __shared__ int s_ext[4];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = tid % GS, grp = tid / GS;
  const int K = p.K;
  [[maybe_unused]] cg::cluster_group cl = cg::this_cluster();
  int crank = 0, csize = 1;
  if constexpr (CL) { crank = (int)cl.block_rank(); csize = (int)cl.num_blocks(); }
  [[maybe_unused]] double* s_xch = s_d + L::D_XCH;
  const int slot = p.order_begin + (int)blockIdx.x / csize;
  const int g = p.gene_order ? p.gene_order[slot] : slot;
  int64_t p0 = p.col_ptr[g];
  const int64_t p1_all = p.col_ptr[g + 1];
  int ncol = (int)(p1_all - p0);
  const int64_t p0_all = p0;
  const int ncol_all = ncol;
  if constexpr (CL) {
    const int slice = ((ncol + csize - 1) / csize + 3) & ~3;
    const int lo = min(ncol, crank * slice), hi = min(ncol, lo + slice);
    p0 += lo; ncol = hi - lo;
  }
  const bool compact = (p.flags & STM_FIT_COMPACT_INPUT) != 0;
  const Column col{compact ? static_cast<const void*>(reinterpret_cast<const unsigned short*>(p.row_idx) + p0)
                           : static_cast<const void*>(p.row_idx + p0),
                   compact ? static_cast<const void*>(reinterpret_cast<const short*>(p.val) + p0)
                           : static_cast<const void*>(p.val + p0),
                   compact};
  const bool remove_low = (p.flags & STM_FIT_REMOVE_LOW_EXP_SPOTS) != 0;
  const bool device_init = (p.flags & STM_FIT_DEVICE_INIT) != 0;

#ifdef STM_EM_PROFILE
  long long pf_t = clock64(), pf_pro = 0, pf_seed = 0, pf_lloyd = 0, pf_assign = 0, pf_em = 0;
#define EM_PROF(x) { const long long _c = clock64(); x += _c - pf_t; pf_t = _c; }
#else
#define EM_PROF(x)
#endif
  // ---- pre-pass: weights, N, origin, extent -------------------------------------------------
  double mean_nz = 0.0;
  if (remove_low) {
    long long sum = 0, cnt = 0;
    for (int i = tid; i < ncol; i += T) {
      const int w = col.trunc_value(i);
      sum += w; cnt += (w != 0);
    }
    sum = warp_sum(sum); cnt = warp_sum(cnt);
    if (lane == 0) { s_d[L::D_LL + warp] = (double)sum; s_d[L::D_STAT + warp] = (double)cnt; }
    __syncthreads();
    double ts = 0, tc = 0;
    for (int w = 0; w < WARPS; ++w) { ts += s_d[L::D_LL + w]; tc += s_d[L::D_STAT + w]; }
    __syncthreads();
    if constexpr (CL) {
      double v[2] = {ts, tc};
      cluster_combine<2, 2>(cl, s_xch, v, csize);
      ts = v[0]; tc = v[1];
    }
    mean_nz = tc > 0 ? ts / tc : 0.0;
  }
  if (tid < 4) s_ext[tid] = (tid & 1) ? INT_MIN : INT_MAX;  // xmin, xmax, ymin, ymax
  __syncthreads();
  long long npos = 0, sw = 0, swx = 0, swy = 0;
  int xmin = INT_MAX, xmax = INT_MIN, ymin = INT_MAX, ymax = INT_MIN;
  // The column (slice) is read from global memory ONCE when it fits the staging budget: the same pass gathers the
  // statistics and stages the positive spots (ordered ballot compaction) relative to a provisional origin — the
  // first entry's spot — which is shifted to the integer centroid afterwards, in shared memory.  Of a longer
  // column the first `smem_spots` entries are staged and only the remainder streams from L2 on every sweep.
  const int n_head = min(ncol, p.smem_spots);
  const int x0 = ncol > 0 ? __ldg(p.spot_x + col.row(0)) : 0, y0 = ncol > 0 ? __ldg(p.spot_y + col.row(0)) : 0;
  int n_st = 0;
  {
    // 4 consecutive entries per thread and step: 8 independent loads + 8 gathers in flight per thread, and a
    // quarter of the barriers of a one-entry-per-thread compaction (the loop is global-latency bound)
    constexpr int E = 4;
    for (int base = 0; base < n_head; base += E * T) {
      int w[E], x[E], y[E];
      int cnt = 0;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int i = base + E * tid + e;
        w[e] = i < n_head ? col.weight(i, remove_low, mean_nz) : 0;
      }
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int i = base + E * tid + e;
        x[e] = 0; y[e] = 0;
        if (w[e] > 0) { const int r = col.row(i); x[e] = __ldg(p.spot_x + r); y[e] = __ldg(p.spot_y + r); ++cnt; }
      }
      int incl = cnt;   // inclusive warp scan of the per-thread keep counts
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL_MASK, incl, o); if (lane >= o) incl += v; }
      if (lane == 31) s_cnt[warp] = incl;
      __syncthreads();
      int off = n_st, tot = 0;
#pragma unroll
      for (int ww = 0; ww < WARPS; ++ww) { const int c = s_cnt[ww]; if (ww < warp) off += c; tot += c; }
      int idx = off + incl - cnt;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        if (w[e] > 0) {
          npos += 1; sw += w[e];
          swx += (long long)w[e] * x[e];
          swy += (long long)w[e] * y[e];
          xmin = min(xmin, x[e]); xmax = max(xmax, x[e]); ymin = min(ymin, y[e]); ymax = max(ymax, y[e]);
          s_spot[idx++] = make_int2(pack_xy(x[e] - x0, y[e] - y0), __float_as_int((float)w[e]));
        }
      }
      n_st += tot;
      __syncthreads();
    }
    for (int i = n_head + tid; i < ncol; i += T) {
      const int w = col.weight(i, remove_low, mean_nz);
      if (w > 0) {
        const int r = col.row(i);
        const int x = __ldg(p.spot_x + r), y = __ldg(p.spot_y + r);
        npos += 1; sw += w;
        swx += (long long)w * x;
        swy += (long long)w * y;
        xmin = min(xmin, x); xmax = max(xmax, x); ymin = min(ymin, y); ymax = max(ymax, y);
      }
    }
  }
  npos = warp_sum(npos); sw = warp_sum(sw); swx = warp_sum(swx); swy = warp_sum(swy);
  xmin = __reduce_min_sync(FULL_MASK, xmin); xmax = __reduce_max_sync(FULL_MASK, xmax);
  ymin = __reduce_min_sync(FULL_MASK, ymin); ymax = __reduce_max_sync(FULL_MASK, ymax);
  if (lane == 0) {
    s_d[L::D_STAT + 4 * warp + 0] = (double)npos; s_d[L::D_STAT + 4 * warp + 1] = (double)sw;
    s_d[L::D_STAT + 4 * warp + 2] = (double)swx;  s_d[L::D_STAT + 4 * warp + 3] = (double)swy;
    atomicMin(&s_ext[0], xmin); atomicMax(&s_ext[1], xmax); atomicMin(&s_ext[2], ymin); atomicMax(&s_ext[3], ymax);
  }
  __syncthreads();
  double t_npos = 0, t_sw = 0, t_swx = 0, t_swy = 0;
  for (int w = 0; w < WARPS; ++w) {
    t_npos += s_d[L::D_STAT + 4 * w + 0]; t_sw += s_d[L::D_STAT + 4 * w + 1];
    t_swx += s_d[L::D_STAT + 4 * w + 2];  t_swy += s_d[L::D_STAT + 4 * w + 3];
  }
  xmin = s_ext[0]; xmax = s_ext[1]; ymin = s_ext[2]; ymax = s_ext[3];
  __syncthreads();
  if constexpr (CL) {   // totals and extent of the whole gene
    double v[8] = {t_npos, t_sw, t_swx, t_swy, (double)xmin, -(double)xmax, (double)ymin, -(double)ymax};
    cluster_combine<8, 4>(cl, s_xch, v, csize);
    t_npos = v[0]; t_sw = v[1]; t_swx = v[2]; t_swy = v[3];
    // an empty slice contributes +-INT_MAX/INT_MIN, which the doubles carry exactly
    xmin = (int)v[4]; xmax = (int)-v[5]; ymin = (int)v[6]; ymax = (int)-v[7];
  }
  const double N_total = t_sw;

  double* o_w = p.out_w + (size_t)g * K;
  double* o_mu = p.out_mu + (size_t)g * K * 2;
  double* o_cov = p.out_cov + (size_t)g * K * 4;
  if ((int)t_npos < K) {  // distribution.py:125,129-130 -> gene dropped
    if (crank == 0) {
      for (int i = tid; i < K; i += T) o_w[i] = 0.0;
      for (int i = tid; i < 2 * K; i += T) o_mu[i] = 0.0;
      for (int i = tid; i < 4 * K; i += T) o_cov[i] = 0.0;
      if (tid == 0) {
        p.out_n_iter[g] = 0; p.out_converged[g] = 0; p.out_lb[g] = 0.0; p.out_status[g] = STM_GENE_DROPPED;
      }
    }
    return;   // uniform over the cluster: t_npos is the cluster total
  }
  const int ox = (int)rint(t_swx / t_sw), oy = (int)rint(t_swy / t_sw);
  // the staged record holds origin-relative coordinates as int16
  const bool packable = (xmax - xmin) < 32768 && (ymax - ymin) < 32768;
  const bool staged = packable;
  if (!staged) n_st = 0;
  const int s_lo = staged ? n_head : 0;      // column entries [s_lo, ncol) stream from L2 on every sweep
  const int n_spots = n_st + (ncol - s_lo);  // positions: the staged spots, then the streamed entries
  if (staged) {
    const int dx = ox - x0, dy = oy - y0;   // provisional origin -> integer centroid
    for (int i = tid; i < n_st; i += T) {
      const int pk = s_spot[i].x;
      s_spot[i].x = pack_xy(((pk << 16) >> 16) - dx, (pk >> 16) - dy);
    }
    __syncthreads();
  }

  const PackedSpots staged_src{s_spot};
  const StreamedSpots stream_src{col, p.spot_x, p.spot_y, ox, oy, remove_low, mean_nz};

  if (tid == 0) { s_flag[0] = 0; s_flag[1] = 0; }
  // padded components never win a max / an argmin
  for (int k = K + tid; k < KP; k += T) {
    float* q = s_par + 8 * k;
    q[0] = 0.f; q[1] = 0.f; q[2] = 0.f; q[3] = 0.f; q[4] = 0.f; q[5] = -1e30f; q[6] = 1e18f; q[7] = 1e18f;
    const int psub = k / KT, pj = k % KT;
    if (pj < 2 * (KT / 2)) {
      float* pq = s_f + L::F_PP + ((psub * (KT / 2) + (pj >> 1)) * 6) * 2 + (pj & 1);
      pq[0] = 0.f; pq[2] = 0.f; pq[4] = 0.f; pq[6] = 0.f; pq[8] = 0.f; pq[10] = -1e30f;
    }
  }
  __syncthreads();

  EM_PROF(pf_pro)
  if (device_init) {
    // =========================================================================================
    // Seed subset (every stride-th spot, at most seed_cap) -> weighted greedy k-means++ -> Lloyd on the
    // subset -> one assignment pass over ALL spots -> one-hot responsibilities -> (w, mu, cov) as
    // sklearn's GaussianMixture._initialize (_gaussian_mixture.py:849-881).
    // =========================================================================================
    int n_glob = n_spots, pos0 = 0;   // spots of the whole gene / position of this CTA's first spot among them
    if constexpr (CL) {
      if (tid == 0) s_xch[0] = (double)n_spots;
      cl.sync();
      n_glob = 0;
      for (int r = 0; r < csize; ++r) {
        const int nr = (int)*cl.map_shared_rank(s_xch, r);
        if (r < crank) pos0 += nr;
        n_glob += nr;
      }
      cl.sync();
    }
    const int stride = (n_glob + p.seed_cap - 1) / p.seed_cap;
    const int ns = (n_glob + stride - 1) / stride;
    if constexpr (CL) {
      // every CTA runs the (cheap, deterministic) seeding + Lloyd on its own full copy of the seed subset:
      // the owner of a seed spot writes it into the seed array of every CTA of the cluster
      for (int j = (pos0 + stride - 1) / stride + tid; j * stride < pos0 + n_spots; j += T) {
        float x, y, w;
        const int pos = j * stride - pos0;
        if (pos < n_st) staged_src.get(pos, x, y, w); else stream_src.get(s_lo + pos - n_st, x, y, w);
        const float4 v = make_float4(x, y, w, 0.f);
        for (int r = 0; r < csize; ++r) cl.map_shared_rank(s_seed, r)[j] = v;
      }
      cl.sync();
    } else {
      for (int j = tid; j < ns; j += T) {
        float x, y, w;
        const int pos = j * stride;
        if (pos < n_st) staged_src.get(pos, x, y, w); else stream_src.get(s_lo + pos - n_st, x, y, w);
        s_seed[j] = make_float4(x, y, w, 0.f);
      }
      __syncthreads();
    }
    // k-means++ (D^2 sampling, Arthur & Vassilvitskii) by an exponential race: drawing index i with probability
    // v_i / sum(v) is arg-min_i E_i / v_i with E_i ~ Exp(1), so one pass updates the squared distances to the
    // centre just chosen AND races for the next centre, and the draw is a single block arg-min — no prefix
    // sums, no inverse-CDF walk.  E_i comes from a counter-based hash of (seed, gene, round, i).
    double* s_scr = s_d + L::D_SCR;
    unsigned long long* s_race = reinterpret_cast<unsigned long long*>(s_scr);             // [WARPS]
    // the per-gene random stream is keyed by the gene's DATA (length, first / last stored entry), not by its index in
    // this call: the same gene gets the same start whether it is fitted alone, in an upload chunk, on another rank
    // of a sharded job or split over a cluster
    uint64_t gkey = (uint64_t)(unsigned)ncol_all;
    if (ncol_all > 0) {
      const Column whole{compact ? static_cast<const void*>(reinterpret_cast<const unsigned short*>(p.row_idx) + p0_all)
                                 : static_cast<const void*>(p.row_idx + p0_all),
                         compact ? static_cast<const void*>(reinterpret_cast<const short*>(p.val) + p0_all)
                                 : static_cast<const void*>(p.val + p0_all),
                         compact};
      gkey ^= (uint64_t)(unsigned)whole.row(0) << 32;
      gkey ^= (uint64_t)(unsigned)whole.row(ncol_all - 1) * 0x9E3779B97F4A7C15ULL;
      gkey ^= (uint64_t)(unsigned)whole.trunc_value(0) << 20;
      gkey ^= (uint64_t)(unsigned)whole.row(ncol_all / 2) * 0xC2B2AE3D27D4EB4FULL;
    }
    const unsigned hseed = (unsigned)splitmix64(p.seed ^ (gkey * 0xD1342543DE82EF95ULL));
    float ccx = 0.f, ccy = 0.f;
    constexpr int SEED_R = 8;   // seeds per thread that can stay in registers over the K rounds
    if (ns <= SEED_R * T) {
      // Register-resident rounds: a thread keeps ITS seeds (x, y, w, running D^2) in registers; the arg-min over the
      // warp is two REDUX instructions, the arg-min over the CTA goes through one shared entry per warp (double-
      // buffered) and two more REDUX — one barrier, no atomics and no seed traffic per round (the rounds are
      // latency-bound).  Same keys, same (key, index) order, hence the same picks as the generic loop below.
      uint2* s_wkey = reinterpret_cast<uint2*>(s_red);   // [2 * WARPS]; s_red is idle until the first flush_stats
      static_assert(4 * WARPS <= WARPS * NSTAT, "s_wkey aliases s_red");
      float px[SEED_R], py[SEED_R], pw[SEED_R], pd[SEED_R];
#pragma unroll
      for (int r = 0; r < SEED_R; ++r) {
        const int i = tid + r * T;
        const float4 v = i < ns ? s_seed[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        px[r] = v.x; py[r] = v.y; pw[r] = v.z; pd[r] = 0.f;
      }
      for (int c = 0; c < K; ++c) {
        unsigned bk = 0xffffffffu;
        int bi = 0x7fffffff;
#pragma unroll
        for (int r = 0; r < SEED_R; ++r) {
          const int i = tid + r * T;
          float val = pw[r];                                   // round 0: P(i) ~ w_i   (0 beyond ns)
          if (c > 0) {
            const float dx = px[r] - ccx, dy = py[r] - ccy;
            const float d = fmaf(dx, dx, dy * dy);
            pd[r] = c == 1 ? d : fminf(pd[r], d);
            val = pw[r] * pd[r];                               // P(i) ~ w_i * D_i^2
          }
          if (val > 0.f) {
            unsigned h = hseed ^ ((unsigned)c * 0x85EBCA77u) ^ ((unsigned)i * 0xC2B2AE3Du);
            h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
            const float u = (float)((h >> 8) + 1u) * 5.9604645e-8f;          // (0, 1]
            const unsigned kb = __float_as_uint(-lg2_approx(u) * rcp_approx(val));   // Exp(1) / v_i >= 0: bits order as uint
            if (kb < bk) { bk = kb; bi = i; }                  // r ascending: the lowest index wins ties
          }
        }
        const unsigned wk = __reduce_min_sync(FULL_MASK, bk);
        const int wi = __reduce_min_sync(FULL_MASK, bk == wk ? bi : 0x7fffffff);
        uint2* wkey = s_wkey + (c & 1) * WARPS;
        if (lane == 0) wkey[warp] = make_uint2(wk, (unsigned)wi);
        __syncthreads();
        const uint2 e = lane < WARPS ? wkey[lane] : make_uint2(0xffffffffu, 0x7fffffffu);
        const unsigned gk = __reduce_min_sync(FULL_MASK, e.x);
        const unsigned gi = __reduce_min_sync(FULL_MASK, e.x == gk ? e.y : 0x7fffffffu);
        const int pick = gk == 0xffffffffu ? 0 : (int)gi;   // all remaining mass on chosen centres: reuse seed 0
        const float4 pv = s_seed[pick];
        ccx = pv.x; ccy = pv.y;
        if (tid == 0) { s_par[8 * c + 6] = ccx; s_par[8 * c + 7] = ccy; }
      }
      __syncthreads();
    } else {
    for (int c = 0; c < K; ++c) {
      unsigned long long best = ~0ULL;
      for (int i = tid; i < ns; i += T) {
        float4 v = s_seed[i];
        float val = v.z;                                   // round 0: P(i) ~ w_i
        if (c > 0) {
          const float dx = v.x - ccx, dy = v.y - ccy;
          const float d = fmaf(dx, dx, dy * dy);
          v.w = c == 1 ? d : fminf(v.w, d);
          s_seed[i].w = v.w;
          val = v.z * v.w;                                 // P(i) ~ w_i * D_i^2
        }
        if (val > 0.f) {
          unsigned h = hseed ^ ((unsigned)c * 0x85EBCA77u) ^ ((unsigned)i * 0xC2B2AE3Du);
          h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
          const float u = (float)((h >> 8) + 1u) * 5.9604645e-8f;          // (0, 1]
          const float key = -lg2_approx(u) * rcp_approx(val);                // Exp(1) / v_i, >= 0
          const unsigned long long kk = ((unsigned long long)__float_as_uint(key) << 32) | (unsigned)i;
          best = min(best, kk);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(FULL_MASK, best, o));
      if (lane == 0) s_race[warp] = best;
      __syncthreads();
      best = s_race[0];
#pragma unroll
      for (int w = 1; w < WARPS; ++w) best = min(best, s_race[w]);
      const int pick = best == ~0ULL ? 0 : (int)(unsigned)best;   // all remaining mass on chosen centres: reuse seed 0
      const float4 pv = s_seed[pick];
      ccx = pv.x; ccy = pv.y;
      if (tid == 0) { s_par[8 * c + 6] = ccx; s_par[8 * c + 7] = ccy; }
      __syncthreads();   // s_race is rewritten and s_seed[].w updated in the next round
    }
    }
    if (tid < K) { s_d[L::D_MU + 2 * tid] = (double)s_par[8 * tid + 6]; s_d[L::D_MU + 2 * tid + 1] = (double)s_par[8 * tid + 7]; }
    __syncthreads();

    EM_PROF(pf_seed)
    // tolerance of sklearn KMeans: tol * mean(var(X, axis=0)) with tol = 1e-4 (cluster/_kmeans.py _tolerance),
    // estimated on the seed subset
    const SeedSpots seed_src{s_seed};
    double vx = 0.0, vy = 0.0, vw = 0.0;
    {
      const double mxo = t_swx / t_sw - ox, myo = t_swy / t_sw - oy;
      for (int i = tid; i < ns; i += T) {
        const float4 v = s_seed[i];
        const double dx = v.x - mxo, dy = v.y - myo;
        vx += v.z * dx * dx; vy += v.z * dy * dy; vw += v.z;
      }
      vx = warp_sum(vx); vy = warp_sum(vy); vw = warp_sum(vw);
      if (lane == 0) { s_scr[3 * warp] = vx; s_scr[3 * warp + 1] = vy; s_scr[3 * warp + 2] = vw; }
      __syncthreads();
      vx = 0.0; vy = 0.0; vw = 0.0;
      for (int w = 0; w < WARPS; ++w) { vx += s_scr[3 * w]; vy += s_scr[3 * w + 1]; vw += s_scr[3 * w + 2]; }
      __syncthreads();
    }
    const double km_tol = 1e-4 * 0.5 * (vx + vy) / fmax(vw, 1.0);

    for (int it = 0; it < p.km_max_iter; ++it) {
      float ax[KT], ay[KT], bb[KT];
#pragma unroll
      for (int j = 0; j < KT; ++j) {
        const float cx = s_par[8 * (sub * KT + j) + 6], cy = s_par[8 * (sub * KT + j) + 7];
        ax[j] = 2.f * cx; ay[j] = 2.f * cy; bb[j] = -fmaf(cx, cx, cy * cy);
      }
      for (int i = tid; i < NSTAT; i += T) s_d[L::D_STAT + i] = 0.0;
      __syncthreads();
      {
        float acc[NPAD];
#pragma unroll
        for (int i = 0; i < NPAD; ++i) acc[i] = 0.f;
        lloyd_sweep<KT, GS, T>(seed_src, 0, ns, grp, sub, ax, ay, bb, acc);
        flush_stats<KT, GS, T, NPAD>(acc, s_d, s_red, tid);
      }
      if (tid < 64) {   // K <= 64: warps 0 and 1; the centre shift is summed in a fixed order (no atomics), so the
                        // iteration count is reproducible — and identical on every CTA of a cluster
        const int k = tid;
        double shift2 = 0.0;
        if (k < K) {
          const double* st = s_d + L::D_STAT + 6 * k;
          if (st[0] > 0.0) {
            const double nx = st[1] / st[0], ny = st[2] / st[0];
            const double ddx = nx - s_d[L::D_MU + 2 * k], ddy = ny - s_d[L::D_MU + 2 * k + 1];
            s_d[L::D_MU + 2 * k] = nx;
            s_d[L::D_MU + 2 * k + 1] = ny;
            s_par[8 * k + 6] = (float)nx;
            s_par[8 * k + 7] = (float)ny;
            shift2 = ddx * ddx + ddy * ddy;
          }
        }
        shift2 = warp_sum(shift2);
        if (lane == 0) s_d[L::D_MISC + 2 + warp] = shift2;
      }
      __syncthreads();
      const bool done = s_d[L::D_MISC + 2] + s_d[L::D_MISC + 3] <= km_tol;
      __syncthreads();
      if (done) break;
    }
    EM_PROF(pf_lloyd)
    // assignment pass with second moments -> initial (w, mu, cov): one-hot responsibilities as sklearn's
    // _initialize, evaluated on the seed subset (every stride-th spot) — the first EM iteration sees all spots
    {
      float cx[KT], cy[KT];
#pragma unroll
      for (int j = 0; j < KT; ++j) { cx[j] = s_par[8 * (sub * KT + j) + 6]; cy[j] = s_par[8 * (sub * KT + j) + 7]; }
      for (int i = tid; i < NSTAT; i += T) s_d[L::D_STAT + i] = 0.0;
      __syncthreads();
      {
        float acc[NPAD];
#pragma unroll
        for (int i = 0; i < NPAD; ++i) acc[i] = 0.f;
        assign_moments_sweep<KT, GS, T>(seed_src, 0, ns, grp, sub, cx, cy, acc);
        flush_stats<KT, GS, T, NPAD>(acc, s_d, s_red, tid);
      }
      if (tid < K) {
        const int k = tid;
        const double* st = s_d + L::D_STAT + 6 * k;
        const double n = st[0] + 10.0 * DBL_EPSILON;  // _gaussian_mixture.py:312
        const double ddx = st[1] / n, ddy = st[2] / n;
        // measured from the float centre the sweep used, not the FP64 master copy
        s_d[L::D_MU + 2 * k] = (double)s_par[8 * k + 6] + ddx;
        s_d[L::D_MU + 2 * k + 1] = (double)s_par[8 * k + 7] + ddy;
        s_d[L::D_COV + 3 * k] = st[3] / n - ddx * ddx + p.reg_covar;
        s_d[L::D_COV + 3 * k + 1] = st[4] / n - ddx * ddy;
        s_d[L::D_COV + 3 * k + 2] = st[5] / n - ddy * ddy + p.reg_covar;
        s_d[L::D_W + k] = n / fmax(vw, 1.0);          // weights /= n_samples (of the subset), _gaussian_mixture.py:866
      }
    }
  } else if (tid < K) {
    const int k = tid;
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
  __syncthreads();
  if (tid < K) {
    if (!refresh_component<L>(tid, s_d, s_par)) atomicOr(&s_flag[1], 1);
  }
  __syncthreads();
#if STM_EM_NOMAX && STM_EM_PACKED
  shift_log_offsets<L>(K, s_d, s_par, tid);
  __syncthreads();
#endif

  EM_PROF(pf_assign)
#if STM_EM_L2_PREFETCH
  // While this gene iterates, pull the column of a gene that will start about one wave of CTAs later into L2, so its
  // prologue (bound by the latency of its first loads) finds the data on chip.
  if constexpr (!CL) {
    const int ahead = slot + (int)gridDim.x / 8 + 148;
    if (ahead < p.order_begin + (int)gridDim.x) {
      const int ga = p.gene_order ? p.gene_order[ahead] : ahead;
      const int64_t q0 = p.col_ptr[ga], q1 = p.col_ptr[ga + 1];
      const char* r0 = compact ? reinterpret_cast<const char*>(reinterpret_cast<const unsigned short*>(p.row_idx) + q0)
                               : reinterpret_cast<const char*>(p.row_idx + q0);
      const char* v0 = compact ? reinterpret_cast<const char*>(reinterpret_cast<const short*>(p.val) + q0)
                               : reinterpret_cast<const char*>(p.val + q0);
      const int64_t bytes = (q1 - q0) * (compact ? 2 : 4);
      for (int64_t o = (int64_t)tid * 128; o < bytes; o += (int64_t)T * 128) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(r0 + o));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(v0 + o));
      }
    }
  }
#endif
  // ---- EM loop (sklearn mixture/_base.py:265-278) ---------------------------------------------
  double lb_prev = -INFINITY, lb = 0.0;
  int n_iter = 0, converged = 0;
  bool failed = s_flag[1] != 0;

  while (!failed && n_iter < p.max_iter) {
    ++n_iter;
    for (int i = tid; i < NSTAT; i += T) s_d[L::D_STAT + i] = 0.0;
    __syncthreads();
    double ll = 0.0;
    // part 0: the staged spots (shared memory); part 1: the column entries that did not fit (L2)
    for (int c0 = 0; c0 < n_st; c0 += CHUNK)
      ll += em_chunk<KT, GS, S, T>(staged_src, c0, min(n_st, c0 + CHUNK), s_d, s_f);
    for (int c0 = s_lo; c0 < ncol; c0 += CHUNK)
      ll += em_chunk<KT, GS, S, T>(stream_src, c0, min(ncol, c0 + CHUNK), s_d, s_f);
    ll = warp_sum(sub == 0 ? ll : 0.0);
    if (lane == 0) s_d[L::D_LL + warp] = ll;
    const double* st_all = s_d + L::D_STAT;   // statistics of the whole gene
    const double* ll_all = s_d + L::D_LL;
    if constexpr (CL) {
      cl.sync();   // every CTA's partial statistics are complete and visible cluster-wide
      for (int i = tid; i < NSTAT + WARPS; i += T) {
        const double* src = i < NSTAT ? s_d + L::D_STAT + i : s_d + L::D_LL + (i - NSTAT);
        double t = 0.0;
        for (int r = 0; r < csize; ++r) t += *cl.map_shared_rank(src, r);   // rank order: identical on every CTA
        s_d[L::D_TOT + i] = t;
      }
      cl.sync();   // all remote reads done before any CTA zeroes its partials for the next iteration
      st_all = s_d + L::D_TOT;
      ll_all = s_d + L::D_TOT + NSTAT;
    } else {
      __syncthreads();
    }
    // M-step finalize: whitened moments -> (w, mu, cov) (sklearn _gaussian_mixture.py:282-320,883-901)
    if (tid < K) {
      const int k = tid;
      const double* st = st_all + 6 * k;
      const double n = st[0] + 10.0 * DBL_EPSILON;  // _gaussian_mixture.py:312
      double tot = 0.0;                             // weights_ /= weights_.sum(), :898
      for (int j = 0; j < K; ++j) tot += st_all[6 * j] + 10.0 * DBL_EPSILON;
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
    if (tid == T - 1) {
      double t = 0.0;
      for (int w = 0; w < WARPS; ++w) t += ll_all[w];
      // log2 domain -> natural log; mean over the N replicated rows (_base.py:332)
#if STM_EM_NOMAX && STM_EM_PACKED
      t += s_d[L::D_MISC + 7] * N_total;     // the offset the sweep's coefficients were shifted by (every spot: w * cmax)
#endif
      s_d[L::D_MISC + 1] = t * 0.69314718055994530942 / N_total;
    }
    __syncthreads();
#if STM_EM_NOMAX && STM_EM_PACKED
    shift_log_offsets<L>(K, s_d, s_par, tid);   // the refreshed coefficients, for the next sweep
    __syncthreads();
#endif
    lb = s_d[L::D_MISC + 1];
    failed = s_flag[1] != 0;
    const double change = lb - lb_prev;
    lb_prev = lb;
    if (fabs(change) < p.tol) { converged = 1; break; }
  }

  EM_PROF(pf_em)
#ifdef STM_EM_PROFILE
  if (tid == 0 && ((blockIdx.x % 997) == 3 || (CL && blockIdx.x % 61 == 0)))
    printf("[em prof] T=%d gene=%d ncol=%d n_spots=%d staged=%d n_iter=%d cycles: prologue %lld seed %lld lloyd %lld assign+refresh %lld em %lld (%.0f/iter)\n",
           T, g, ncol, n_spots, (int)staged, n_iter, pf_pro, pf_seed, pf_lloyd, pf_assign, pf_em, (double)pf_em / max(n_iter, 1));
#endif
  // ---- outputs --------------------------------------------------------------------------------
  if (crank != 0) return;   // every CTA of a cluster holds the same result; no remote access is pending
  if (failed) {
    for (int i = tid; i < K; i += T) o_w[i] = 0.0;
    for (int i = tid; i < 2 * K; i += T) o_mu[i] = 0.0;
    for (int i = tid; i < 4 * K; i += T) o_cov[i] = 0.0;
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

// ------------------------------------------------------------------------------------------------
// Launch configurations.  Genes are sorted by decreasing nnz and cut into STM_FIT_N_BUCKETS contiguous buckets:
//   0..3  cluster buckets: one gene per thread-block cluster of 16 / 8 / 4 / 2 CTAs (512 threads, 1 CTA per SM);
//         a gene goes to the smallest cluster whose CTAs can each stage their slice, genes beyond 16 slices run
//         in bucket 0 and stream the part of each slice that does not fit from L2;
//   4..6  one CTA per gene: large (512 threads x 1 CTA/SM), medium (256 x 2), small (128 x 4).
// The staging capacity of a bucket follows from the 228 KB of shared memory an SM has.
// ------------------------------------------------------------------------------------------------
constexpr int N_BUCKETS = STM_FIT_N_BUCKETS;
constexpr int N_CLUSTER_BUCKETS = 4;
constexpr int CLUSTER_SIZE[N_CLUSTER_BUCKETS] = {16, 8, 4, 2};
constexpr int SPOT_BYTES = 8, SEED_BYTES = 16;
constexpr int DEFAULT_MAX_OPTIN = 232448;  // sm_100: 227 KB per CTA

inline int32_t cap_for(size_t fixed, int minb, int max_optin, int seed_cap) {
  size_t per_cta = ((size_t)228 * 1024) / minb - 1024 - 256;  // 1 KB reserved per CTA + static smem
  if (per_cta > (size_t)max_optin) per_cta = max_optin;
  const size_t need = fixed + (size_t)seed_cap * SEED_BYTES;
  if (per_cta < need + SPOT_BYTES * 64) return 0;
  return (int32_t)(((per_cta - need) / SPOT_BYTES) & ~(size_t)3);
}

// c[0] describes the cluster buckets (same CTA shape for every cluster size), c[1..3] the one-CTA buckets
struct BucketCfg { int T, minb; size_t fixed; int seed_cap; };
constexpr int N_CFG = 4;

template <int KT, int GS>
inline void fill_cfg(BucketCfg (&c)[N_CFG]) {
  c[0] = {512, 1, EmLayout<KT, GS, 512, true>::fixed_bytes(), 2048};
  c[1] = {STM_EM_T_LARGE, 1, EmLayout<KT, GS, STM_EM_T_LARGE>::fixed_bytes(), 2048};
  c[2] = {256, 2, EmLayout<KT, GS, 256>::fixed_bytes(), 1024};
  c[3] = {128, STM_EM_MINB_SMALL, EmLayout<KT, GS, 128>::fixed_bytes(), 1024};
}

inline void class_cfg(int K, uint32_t flags, BucketCfg (&c)[N_CFG]) {
  if (K <= 5) fill_cfg<5, 1>(c);
  else if (K <= 10) fill_cfg<5, 2>(c);
  else if (K <= 20) fill_cfg<5, 4>(c);
  else if (K <= 40) fill_cfg<5, 8>(c);
  else for (int b = 0; b < N_CFG; ++b) c[b] = {128, 3, EmLayout<8, 8, 128>::fixed_bytes(), 1024};
  const int user = (int)((flags >> 20) & 0xff) * 256;  // seed-cap field of the flags
  for (int b = 0; b < N_CFG; ++b) {
    if (!(flags & STM_FIT_DEVICE_INIT)) c[b].seed_cap = 0;
    else if (user) c[b].seed_cap = user;
  }
}
inline bool class_has_clusters(int K) { return K <= 40; }   // K in (40, 64] has a single kernel shape

template <int KT, int GS, int S, int T, int MINB, bool CL>
int launch_em(EmArgs a, const BucketCfg& cfg, int n_genes, int cluster, cudaStream_t stream) {
  using L = EmLayout<KT, GS, T, CL>;
  auto kern = gmm_em_kernel<KT, GS, S, T, MINB, CL>;
  int dev = 0, max_optin = 0;
  STM_CHECK_CUDA(cudaGetDevice(&dev));
  STM_CHECK_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  const size_t fixed = L::fixed_bytes();
  a.seed_cap = cfg.seed_cap;
  a.smem_spots = cap_for(fixed, MINB, max_optin, a.seed_cap);
  if (a.smem_spots <= 0) {
    set_last_error("stm_gmm_fit_batched: shared-memory budget too small");
    return STM_ERR_UNSUPPORTED;
  }
  const size_t smem = fixed + (size_t)a.seed_cap * SEED_BYTES + (size_t)a.smem_spots * SPOT_BYTES;
  STM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if constexpr (!CL) {
    kern<<<n_genes, T, smem, stream>>>(a);
    STM_CHECK_CUDA(cudaGetLastError());
    return STM_OK;
  } else {
    // clusters above the portable size of 8 need an opt-in and may not be schedulable at this shared-memory
    // footprint (one CTA per SM, a cluster lives inside one GPC): fall back to the next smaller size, whose
    // slices are then longer than the staging capacity and stream from L2
    STM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t lc{};
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    lc.blockDim = dim3(T, 1, 1);
    lc.dynamicSmemBytes = smem;
    lc.stream = stream;
    lc.attrs = at;
    lc.numAttrs = 1;
    for (; cluster >= 2; cluster /= 2) {
      at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      lc.gridDim = dim3((unsigned)n_genes * cluster, 1, 1);
      int n_active = 0;
      if (cudaOccupancyMaxActiveClusters(&n_active, kern, &lc) == cudaSuccess && n_active >= 1) break;
      (void)cudaGetLastError();
    }
    if (cluster < 2) {
      set_last_error("stm_gmm_fit_batched: no thread-block cluster size is schedulable");
      return STM_ERR_UNSUPPORTED;
    }
    STM_CHECK_CUDA(cudaLaunchKernelEx(&lc, kern, a));
    return STM_OK;
  }
}

template <int KT, int GS>
int launch_bucket(const EmArgs& a, const BucketCfg (&c)[N_CFG], int bucket, int n_genes, cudaStream_t st) {
  if (bucket < N_CLUSTER_BUCKETS) return launch_em<KT, GS, 2, 512, 1, true>(a, c[0], n_genes, CLUSTER_SIZE[bucket], st);
  if (bucket == 4) return launch_em<KT, GS, STM_EM_S_LARGE, STM_EM_T_LARGE, 1, false>(a, c[1], n_genes, 1, st);
  if (bucket == 5) return launch_em<KT, GS, 2, 256, 2, false>(a, c[2], n_genes, 1, st);
  return launch_em<KT, GS, STM_EM_S_SMALL, 128, STM_EM_MINB_SMALL, false>(a, c[3], n_genes, 1, st);
}

int launch_class(const EmArgs& a, int bucket, int n_genes, cudaStream_t st) {
  const int K = a.K;
  BucketCfg c[N_CFG];
  class_cfg(K, a.flags, c);
  if (K <= 5) return launch_bucket<5, 1>(a, c, bucket, n_genes, st);
  if (K <= 10) return launch_bucket<5, 2>(a, c, bucket, n_genes, st);
  if (K <= 20) return launch_bucket<5, 4>(a, c, bucket, n_genes, st);
  if (K <= 40) return launch_bucket<5, 8>(a, c, bucket, n_genes, st);
  return launch_em<8, 8, 1, 128, 3, false>(a, c[3], n_genes, 1, st);
}

// auxiliary streams so the buckets of one call overlap (fork/join on the caller's stream); stream priorities
// were measured to make no difference
struct AuxStreams {
  cudaStream_t s[N_BUCKETS - 1] = {};
  cudaEvent_t fork = nullptr, join[N_BUCKETS - 1] = {};
  int dev = -1;
};
thread_local std::map<int, AuxStreams> g_aux;   // per calling thread and device, created on first use, kept for the
                                                // life of the thread (a rank drives one device)

int get_aux(AuxStreams** out) {
  int dev = 0;
  STM_CHECK_CUDA(cudaGetDevice(&dev));
  AuxStreams& a = g_aux[dev];
  if (a.dev != dev) {
    for (int i = 0; i < N_BUCKETS - 1; ++i) {
      STM_CHECK_CUDA(cudaStreamCreateWithFlags(&a.s[i], cudaStreamNonBlocking));
      STM_CHECK_CUDA(cudaEventCreateWithFlags(&a.join[i], cudaEventDisableTiming));
    }
    STM_CHECK_CUDA(cudaEventCreateWithFlags(&a.fork, cudaEventDisableTiming));
    a.dev = dev;
  }
  *out = &a;
  return STM_OK;
}

}  // namespace
}  // namespace stm

extern "C" int stm_gmm_fit_plan(const int64_t* col_ptr_host, int32_t G, int32_t K, uint32_t flags,
                                int32_t* gene_order_host, int32_t* bucket_counts_host) {
  using namespace stm;
  STM_CHECK_ARG(G >= 0, "G < 0");
  STM_CHECK_ARG(K >= 1 && K <= 64, "K must be in [1, 64]");
  STM_CHECK_ARG(gene_order_host && bucket_counts_host && (col_ptr_host || G == 0), "null pointer");
  std::vector<int32_t> order(G);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
    return (col_ptr_host[a + 1] - col_ptr_host[a]) > (col_ptr_host[b + 1] - col_ptr_host[b]);
  });
  BucketCfg cfg[N_CFG];
  class_cfg(K, flags, cfg);
  int64_t cap[N_CFG];
  for (int b = 0; b < N_CFG; ++b) cap[b] = cap_for(cfg[b].fixed, cfg[b].minb, DEFAULT_MAX_OPTIN, cfg[b].seed_cap);
  const bool clusters = class_has_clusters(K) && !(flags & STM_FIT_NO_CLUSTERS);
  for (int b = 0; b < N_BUCKETS; ++b) bucket_counts_host[b] = 0;
  for (int32_t i = 0; i < G; ++i) {
    const int64_t n = col_ptr_host[order[i] + 1] - col_ptr_host[order[i]];
    int b;
    if (n <= cap[3]) b = 6;
    else if (n <= cap[2]) b = 5;
    else if (n <= cap[1] || !clusters) b = 4;
    else if (n <= 2 * cap[0]) b = 3;
    else if (n <= 4 * cap[0]) b = 2;
    else if (n <= 8 * cap[0]) b = 1;
    else b = 0;
    ++bucket_counts_host[b];
    gene_order_host[i] = order[i];
  }
  return STM_OK;
}

extern "C" int stm_gmm_fit_batched(const int64_t* col_ptr, const int32_t* row_idx, const float* val,
                                   const int32_t* spot_x, const int32_t* spot_y, int32_t n_spots,
                                   int32_t G, int32_t K, const int32_t* gene_order,
                                   const int32_t* bucket_counts_host,
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
  STM_CHECK_ARG(!(flags & STM_FIT_COMPACT_INPUT) || n_spots <= 65536, "compact input needs n_spots <= 65536");
  STM_CHECK_ARG(K >= 1 && K <= 64, "K must be in [1, 64]");
  STM_CHECK_ARG(max_iter >= 1, "max_iter < 1");
  STM_CHECK_ARG(out_w && out_mu && out_cov && out_n_iter && out_converged && out_lower_bound && out_status,
                "null output pointer");
  STM_CHECK_ARG((flags & STM_FIT_DEVICE_INIT) || (init_w && init_mu && init_cov),
                "init_* required without STM_FIT_DEVICE_INIT");
  int32_t counts[N_BUCKETS] = {};
  counts[N_BUCKETS - 1] = G;
  if (bucket_counts_host) {
    STM_CHECK_ARG(gene_order != nullptr, "bucket_counts_host needs gene_order");
    int64_t s = 0;
    for (int b = 0; b < N_BUCKETS; ++b) { counts[b] = bucket_counts_host[b]; STM_CHECK_ARG(counts[b] >= 0, "negative bucket"); s += counts[b]; }
    STM_CHECK_ARG(s == G, "bucket counts do not sum to G");
    for (int b = 0; b < N_CLUSTER_BUCKETS; ++b)
      STM_CHECK_ARG(counts[b] == 0 || class_has_clusters(K), "cluster buckets need K <= 40");
  }
  EmArgs a{};
  a.col_ptr = col_ptr; a.row_idx = row_idx; a.val = val; a.spot_x = spot_x; a.spot_y = spot_y;
  a.gene_order = gene_order; a.init_w = init_w; a.init_mu = init_mu; a.init_cov = init_cov;
  a.out_w = out_w; a.out_mu = out_mu; a.out_cov = out_cov; a.out_n_iter = out_n_iter;
  a.out_converged = out_converged; a.out_lb = out_lower_bound; a.out_status = out_status;
  a.G = G; a.K = K; a.max_iter = max_iter; a.flags = flags; a.reg_covar = reg_covar; a.tol = tol; a.seed = seed;
  const int user_lloyd = (int)((flags >> 8) & 0xff);
  a.km_max_iter = user_lloyd ? (user_lloyd == 255 ? KM_MAX_ITER : user_lloyd) : STM_KM_DEFAULT_LLOYD;
  a.km_cand = 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int n_active = 0;
  for (int b = 0; b < N_BUCKETS; ++b) n_active += counts[b] > 0;
  AuxStreams* aux = nullptr;
  if (n_active > 1) {
    int rc = get_aux(&aux);
    if (rc != STM_OK) return rc;
    STM_CHECK_CUDA(cudaEventRecord(aux->fork, st));
  }
  // buckets in order of decreasing gene length; the first active one runs on the caller's stream, the others on
  // auxiliary streams (measured on B200: 0.4 ms per S50 step better than the other way round)
  int begin = 0, used_aux = 0;
  for (int b = 0; b < N_BUCKETS; ++b) {
    if (counts[b] == 0) continue;
    a.order_begin = begin;
    cudaStream_t s = st;
    const int my_aux = begin > 0 ? used_aux++ : -1;
    if (my_aux >= 0) {
      s = aux->s[my_aux];
      STM_CHECK_CUDA(cudaStreamWaitEvent(s, aux->fork, 0));
    }
    const int rc = launch_class(a, b, counts[b], s);
    if (rc != STM_OK) return rc;
    if (my_aux >= 0) {
      STM_CHECK_CUDA(cudaEventRecord(aux->join[my_aux], s));
      STM_CHECK_CUDA(cudaStreamWaitEvent(st, aux->join[my_aux], 0));
    }
    begin += counts[b];
  }
  return STM_OK;
}
