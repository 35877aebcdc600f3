/*
 * stminer_b200.h — C ABI of the B200-native STMiner spatial-pattern hot path.
 *
 * The reference (xjtu-omics/STMiner v1.1.4) has no FFI layer: its boundary is the
 * Python API.  Each entry point below replaces the arithmetic of
 * one reference function and is what a reference-side ctypes binding would call
 * (INTEGRATION.md shows those stubs).
 *
 * Conventions
 *   - all array arguments are caller-owned DEVICE pointers unless the name ends
 *     in `_host`; nothing is allocated, freed or retained by the library,
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream),
 *   - calls only enqueue work on `stream`; there is no host synchronisation inside,
 *   - return value: 0 on success, <0 on error (STM_ERR_*); the message of the
 *     last error of the calling thread is available from stm_last_error_string(),
 *   - the library keeps no state between calls except, per calling thread and device, a few auxiliary CUDA
 *     streams / events created on first use (the nnz buckets of one fit run concurrently on them, forked from and
 *     joined back into `stream`); it is thread-compatible (one caller thread per GPU rank).
 */
#ifndef STMINER_B200_H_
#define STMINER_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STM_OK 0
#define STM_ERR_INVALID_ARG (-1)
#define STM_ERR_CUDA (-2)
#define STM_ERR_UNSUPPORTED (-3)

/* per-gene status written by stm_gmm_fit_batched */
#define STM_GENE_OK 0
#define STM_GENE_DROPPED 1          /* fewer than K distinct positive spots: reference returns None
                                       (STMiner/Algorithm/distribution.py:125,129-130) */
#define STM_GENE_ILL_CONDITIONED 2  /* a covariance lost positive-definiteness: sklearn raises ValueError
                                       (_gaussian_mixture.py:361-367 -> distribution.py:200-202) */

#define STM_FIT_N_BUCKETS 7

/* flags of stm_gmm_fit_batched */
#define STM_FIT_REMOVE_LOW_EXP_SPOTS 1u /* AlgUtils.py:16-17: w <- round(max(w - mean(nonzero w), 0)) */
#define STM_FIT_DEVICE_INIT 2u          /* ignore init_*; run the on-device weighted k-means++/Lloyd init
                                           (replaces sklearn KMeans inside GaussianMixture, _base.py:119-128) */
#define STM_FIT_COMPACT_INPUT 4u        /* row_idx points to uint16 and val to int16 (values already truncated by the
                                           host, AlgUtils.py:35); needs n_spots <= 65536.  4 B per stored value
                                           instead of 8: halves the host->device transfer the e2e path is bound by */
#define STM_FIT_NO_CLUSTERS 8u          /* stm_gmm_fit_plan: never split a gene over a thread-block cluster (long genes
                                           then stream from L2 inside one CTA) */
/* optional tuning fields of the device init, 0 = library default */
#define STM_FIT_LLOYD_ITERS(n) (((uint32_t)(n) & 0xffu) << 8)   /* Lloyd iteration cap (default 4; the EM that follows
                                                                    refines further), 255 = sklearn's 300 */
#define STM_FIT_SEED_CAP_256(n) (((uint32_t)(n) & 0xffu) << 20)  /* seed-subset capacity in units of 256 spots */

/* Library / build identification. */
int stm_version(void);
const char* stm_last_error_string(void);

/*
 * Batched count-weighted 2-D Gaussian-mixture EM fit — one CTA per gene, whole EM loop on chip.
 *
 * Replaces: fit_gmms / fit_gmm        STMiner/Algorithm/distribution.py:148-205, :87-130
 *           get_exp_array             STMiner/Algorithm/AlgUtils.py:8-36       (int16 truncation of the column)
 *           array_to_list             STMiner/Algorithm/distribution.py:369-373 (count-replicated list == integer weights)
 *           GaussianMixture.fit       scikit-learn 1.3.0 mixture/_base.py:203-312 (E/M loop, tol on the lower bound)
 *
 * Input is the expression matrix in CSC form over DISTINCT spots:
 *   col_ptr[G+1]   int64   column (gene) pointers
 *   row_idx[nnz]   int32   spot index of each stored value          (uint16 with STM_FIT_COMPACT_INPUT)
 *   val[nnz]       float   expression value; truncated toward zero to int16 inside (AlgUtils.py:35);
 *                          entries whose truncated value is <= 0 are ignored   (int16 with STM_FIT_COMPACT_INPUT)
 *   spot_x, spot_y [n_spots] int32 integer grid coordinates of each spot (adata.obs['x'], ['y'])
 *   gene_order[G]  int32   optional (NULL = 0..G-1): CTA b fits gene gene_order[b]; pass genes sorted by
 *                          decreasing nnz so long genes start first (LPT)
 *   bucket_counts_host[STM_FIT_N_BUCKETS]  HOST int32, optional (NULL = one bucket): how many consecutive entries
 *                          of gene_order belong to each nnz bucket, longest genes first, as stm_gmm_fit_plan
 *                          computes them: buckets 0..3 fit one gene per thread-block CLUSTER of 16 / 8 / 4 / 2 CTAs
 *                          (each CTA stages a slice of the gene's spots in its shared memory; the per-iteration
 *                          statistics are summed through distributed shared memory), buckets 4..6 one gene per
 *                          CTA of 512 / 256 / 128 threads.  Each bucket's shared-memory budget keeps its genes'
 *                          spots staged on chip; the buckets of one call run concurrently.  (The reference has
 *                          no counterpart: it fits genes one after another, distribution.py:184.)
 * Initial parameters (ignored with STM_FIT_DEVICE_INIT; required otherwise), float64, original coordinates:
 *   init_w[G,K], init_mu[G,K,2], init_cov[G,K,2,2]
 * Outputs, float64, same layouts: out_w, out_mu, out_cov; plus per gene
 *   out_n_iter int32 (sklearn n_iter_), out_converged int32, out_lower_bound float64 (sklearn lower_bound_),
 *   out_status int32 (STM_GENE_*).  Outputs of a gene whose status != STM_GENE_OK are zero-filled.
 * K <= 64.  reg_covar/tol/max_iter as sklearn (reference: 1e-3 / 1e-3 / 100, distribution.py:148-156).
 * seed is used only by the device init.
 */
int stm_gmm_fit_batched(const int64_t* col_ptr, const int32_t* row_idx, const float* val,
                        const int32_t* spot_x, const int32_t* spot_y, int32_t n_spots,
                        int32_t G, int32_t K, const int32_t* gene_order,
                        const int32_t* bucket_counts_host,
                        const double* init_w, const double* init_mu, const double* init_cov,
                        double reg_covar, double tol, int32_t max_iter, uint32_t flags, uint64_t seed,
                        double* out_w, double* out_mu, double* out_cov,
                        int32_t* out_n_iter, int32_t* out_converged, double* out_lower_bound,
                        int32_t* out_status, void* stream);

/*
 * Host-side launch plan for stm_gmm_fit_batched (no device work): sorts genes by decreasing nnz (the
 * reference fits them in list order, distribution.py:184; order only affects scheduling here) and splits
 * them into the nnz buckets.  col_ptr_host[G+1], gene_order_host[G] and bucket_counts_host[STM_FIT_N_BUCKETS] are
 * HOST arrays; upload gene_order_host and pass both to stm_gmm_fit_batched.  `flags` must be the flags
 * of the fit call (the device init reserves shared memory, which moves the bucket boundaries).
 */
int stm_gmm_fit_plan(const int64_t* col_ptr_host, int32_t G, int32_t K, uint32_t flags,
                     int32_t* gene_order_host, int32_t* bucket_counts_host);

/*
 * GMM-to-GMM distance: K x K cost C[i][j] = (w1_i + w2_j) * Hellinger(N1_i, N2_j), solved as an exact
 * linear assignment (transport LP with unit marginals).  One warp per pair, float64.
 *
 * Replaces: distribution_distance     STMiner/Algorithm/distance.py:13-36
 *           get_hellinger_distance    STMiner/Algorithm/distance.py:50-80
 *           linear_sum                STMiner/Algorithm/algorithm.py:10-26 (scipy.optimize.linear_sum_assignment)
 *           build_gmm_distance_array  STMiner/Algorithm/distance.py:83-99  (all i<j pairs)
 *
 * w[G,K], mu[G,K,2], cov[G,K,2,2] float64.  Pairs are the linearised strict upper triangle in row-major
 * order: p = 0 is (0,1), p = 1 is (0,2) ... ; the call computes pairs [pair_begin, pair_end) and writes
 * out[p - pair_begin] (float64).  A rank of a multi-GPU job passes its own slice.
 */
int stm_gmm_dist_pairs(const double* w, const double* mu, const double* cov, int32_t G, int32_t K,
                       int64_t pair_begin, int64_t pair_end, double* out, void* stream);

/*
 * Exact linear-sum-assignment cost of B caller-supplied K x K cost matrices (row-major float64, device),
 * one warp per matrix.  Replaces linear_sum STMiner/Algorithm/algorithm.py:10-26
 * (scipy.optimize.linear_sum_assignment + cost[row, col].sum()).  out[B] float64; NaN when a matrix holds
 * non-finite entries (scipy raises ValueError there).
 */
int stm_lsap_batched(const double* cost, int32_t K, int64_t B, double* out, void* stream);

/* Mirror a full pair vector (length G(G-1)/2) into a dense symmetric G x G float64 matrix with zero
 * diagonal (distance.py:93-98). */
int stm_gmm_dist_fill_square(const double* pairs, int32_t G, double* out_square, void* stream);

/*
 * One-vs-all variant.  Replaces compare_gmm_distance STMiner/Algorithm/distance.py:102-111.
 * Query GMM (qw[K], qmu[K,2], qcov[K,2,2]) against G GMMs; out[G] float64.
 */
int stm_gmm_dist_one_vs_all(const double* qw, const double* qmu, const double* qcov,
                            const double* w, const double* mu, const double* cov,
                            int32_t G, int32_t K, double* out, void* stream);

/*
 * Batched exact spot-level EMD (squared-Euclidean cost on integer grid coordinates) between one
 * source image and G target images.  One CTA per target; integer network simplex, costs computed on the fly.
 *
 * Replaces: calculate_ot_distance     STMiner/Algorithm/distance.py:191-202 (ot.dist + ot.emd2, POT 0.9.1)
 *           the loop of spatial_high_variable_genes  STMiner/SPFinder.py:285-299
 *
 *   src_x, src_y [n] int32, src_mass[n] float64 (positive, any scale — normalised inside as a/sum(a))
 *   tgt_ptr[G+1] int64, tgt_x, tgt_y int32, tgt_mass float64 (positive; normalised per target)
 *   max_tgt      largest target support in the batch; order[G] optional processing order (largest first)
 *   max_iter     pivot limit on the problem itself (reference: numItermax=200000); block_size: arcs priced per block
 *                search step, 0 = about 32 arcs per thread rounded to whole target rows (LEMON/POT use sqrt(n*m))
 *   flags        0 = multiscale start basis: the grid is coarsened one Morton bit (half a cell) per level, the coarsest problem
 *                is solved from the artificial star and every finer level starts from the refinement of the coarser
 *                optimum; the finest level is the problem itself, solved to optimality by the same simplex, with about
 *                ten times fewer pivots (max_iter applies to that level; out_iters counts all levels).
 *                STM_EMD_STAR_START = the artificial star POT/LEMON start from (one level).  The optimum — the value
 *                the reference returns — does not depend on the start basis.
 *                STM_EMD_NO_CANDIDATES = plain block search: one pricing pass per pivot.  By default the shared-memory
 *                kernels keep every thread's best arc of a successful pass as a candidate list and re-price only that
 *                list until it holds no negative arc (about a third fewer block scans, ~10 % fewer pivots); the
 *                optimum is the same.
 *   workspace    16-byte aligned device scratch of at least stm_emd_workspace_bytes(n_src, max_tgt, n_ctas) bytes;
 *                n_ctas CTAs (one problem at a time each) pull problems from a queue
 *   out_cost[G] float64 = sum_ij G_ij * M_ij;  out_iters[G] (nullable) pivots;  out_status[G]:
 *     0 optimal, 1 iteration limit reached (POT returns the current, sub-optimal value with a warning),
 *     2 empty image, 3 problem does not fit one CTA's shared memory / int16 coordinate span / the range of the
 *     potentials (out_cost = NaN), 4-5 internal limits (cycle longer than 2048 arcs, unbounded ray).
 *
 * A batch whose largest problem exceeds a CTA's shared memory is solved with the tree in global memory (see
 * stm_emd_smem_node_capacity), whichever of the two entry points is called.
 * stm_emd_spots_batched keeps int32 potentials (18 B of shared memory per node, problems up to ~11.9k nodes); it needs
 * 2 * (span_x^2 + span_y^2 + 1) * (n + m + 1) < 2^31 and reports status 3 otherwise.  stm_emd_spots_batched_wide is the
 * same solver (same pivot sequence) with 64-bit potentials for problems beyond that bound — scattered integer
 * coordinates, long thin tissues — at 22 B per node (~9.8k nodes) and a slower pricing loop.
 */
#define STM_EMD_STAR_START 1u
#define STM_EMD_NO_CANDIDATES 2u
int64_t stm_emd_workspace_bytes(int32_t n_src, int64_t max_tgt, int32_t n_ctas);
/* Largest n + m + 1 a batch may reach and still run with the tree in shared memory (wide = the 64-bit kernel) on the
 * current device; a batch beyond it (n_src + max_tgt + 1) runs the same solver with the tree in the CTA's global scratch
 * (one CTA per SM, 64-bit potentials, ~10x slower per pivot) up to stm_emd_global_node_capacity() nodes — Stereo-seq
 * bin50 slides (40,000 source bins).  Callers that mix sizes should batch the small problems separately. */
int32_t stm_emd_smem_node_capacity(int32_t wide);
int32_t stm_emd_global_node_capacity(void);
/* Threads per problem of the candidate-list pricing for a batch of this size: 512 (one CTA per SM), 256 (small
 * problems), 0 = plain block search (the 64-bit shared-memory kernel), -1 = the global-tree kernel (512 threads, sources
 * dealt by index mod 512, ties to the lowest source index).  The CPU oracle takes the value as cand_threads. */
int32_t stm_emd_pricing_threads(int32_t n_src, int64_t max_tgt, int32_t wide);
int stm_emd_spots_batched(const int32_t* src_x, const int32_t* src_y, const double* src_mass, int32_t n_src,
                          const int64_t* tgt_ptr, const int32_t* tgt_x, const int32_t* tgt_y,
                          const double* tgt_mass, int32_t G, int64_t max_tgt, const int32_t* order,
                          int64_t max_iter, int32_t block_size, uint32_t flags,
                          void* workspace, int64_t workspace_bytes, int32_t n_ctas,
                          double* out_cost, int32_t* out_status, int64_t* out_iters, void* stream);
int stm_emd_spots_batched_wide(const int32_t* src_x, const int32_t* src_y, const double* src_mass, int32_t n_src,
                               const int64_t* tgt_ptr, const int32_t* tgt_x, const int32_t* tgt_y,
                               const double* tgt_mass, int32_t G, int64_t max_tgt, const int32_t* order,
                               int64_t max_iter, int32_t block_size, uint32_t flags,
                               void* workspace, int64_t workspace_bytes, int32_t n_ctas,
                               double* out_cost, int32_t* out_status, int64_t* out_iters, void* stream);

/*
 * Lloyd iterations of KMeans on the device, all restarts at once (one CTA per restart), float64.
 *
 * Replaces: the KMeans half of cluster()   STMiner/Algorithm/algorithm.py:57
 *           (sklearn.cluster.KMeans(n_clusters, random_state=0, max_iter=1000, n_init=20).fit -> _kmeans_single_lloyd,
 *            scikit-learn cluster/_kmeans.py; the k-means++ start centres of every restart come from the caller, who draws
 *            them exactly as KMeans.fit does)
 *
 *   X[n,d] row-major, mean-centred as KMeans.fit centres it; centers_init[n_runs,k,d]; k, d <= 64
 *   tol = 1e-4 * mean(var(X, axis=0)) (KMeans' _tolerance); stop on unchanged labels or squared centre shift <= tol
 *   out_centers[n_runs,k,d], out_labels[n_runs,n] int32, out_inertia[n_runs], out_n_iter[n_runs],
 *   out_status[n_runs]: 0 ok, 1 a cluster became empty (scikit-learn relocates it; not restated here)
 *   workspace: stm_kmeans_workspace_bytes(n, n_runs) bytes
 */
int64_t stm_kmeans_workspace_bytes(int32_t n, int32_t n_runs);
int stm_kmeans_lloyd(const double* X, int32_t n, int32_t d, int32_t k, const double* centers_init, int32_t n_runs,
                     int32_t max_iter, double tol, void* workspace, int64_t workspace_bytes,
                     double* out_centers, int32_t* out_labels, double* out_inertia, int32_t* out_n_iter,
                     int32_t* out_status, void* stream);

/*
 * Device-side staging: spot-by-gene CSR matrix -> the gene-major (CSC) arrays stm_gmm_fit_batched reads.
 *
 * Replaces the per-gene host staging of the reference's fit path: get_exp_array / _preprocess
 * STMiner/Algorithm/AlgUtils.py:8-36 (adata[:, gene].X sliced per gene, coo_matrix(..., dtype=int16)), called per gene from
 * STMiner/Algorithm/distribution.py:122.
 *
 *   indptr[n_obs+1] int64, indices int32, data float32: canonical CSR (distinct columns inside a row), rows = observations
 *   col_map[n_var]   output column of every input column, -1 = gene not selected (one-to-one)
 *   row_order[n_obs] the input row that holds spot s, spots in the order the output indexes them (row-major (x, y));
 *                    the rows must sit on distinct spots
 *   compact != 0     outputs are {uint16 spot, int16 value} (n_obs <= 65,536), else {int32 spot, float32 value}
 *   out_col_ptr[n_out+1] int64; out_row_idx / out_val sized for every stored entry of the input (an upper bound)
 * Values are truncated toward zero and wrapped to int16 (AlgUtils.py:35), zeros are dropped, the entries of a column are in
 * increasing spot index — the arrays SpotMatrix.from_matrix builds on the host.  Stable transposition without atomics:
 * stm_stage_blocks(n_obs) row blocks, per-(block, column) counters in the workspace.
 */
int32_t stm_stage_blocks(int32_t n_obs);
int64_t stm_stage_workspace_bytes(int32_t n_obs, int32_t n_out);
int stm_stage_csr_columns(const int64_t* indptr, const int32_t* indices, const float* data, int32_t n_obs,
                          const int32_t* col_map, int32_t n_out, const int32_t* row_order, uint32_t compact,
                          void* workspace, int64_t workspace_bytes, int64_t out_capacity,
                          int64_t* out_col_ptr, void* out_row_idx, void* out_val, void* stream);

/*
 * Pattern vote: the accumulation loop of _genes_to_pattern (STMiner/SPFinder.py:459-489, called per label from
 * get_pattern_array :441-457) for all labels in one pass over the spot-by-gene CSR matrix.
 *
 *   col_label[n_var]  label (0 .. n_labels-1, at most 64) of every input column, -1 = gene not in any label
 *   gene_sum[n_var]   int64 scratch: per-gene totals of the int16-truncated counts (zeroed and filled here)
 *   out_total[n_labels][n_obs] float64:  sum over the label's genes of  w_gs * 100 / sum_s w_gs   (scale_array, :32-38)
 *   out_votes[n_labels][n_obs] int32:    number of the label's genes with w_gs != 0 at the spot
 * The caller scatters the rows onto the (x, y) grid and applies the vote_rate mask.  Deterministic (no floating-point
 * atomics).
 */
int stm_pattern_vote(const int64_t* indptr, const int32_t* indices, const float* data, int32_t n_obs, int32_t n_var,
                     const int32_t* col_label, int32_t n_labels, int64_t* gene_sum,
                     double* out_total, int32_t* out_votes, void* stream);

/*
 * Metric MDS by SMACOF on a precomputed dissimilarity matrix, float64, whole iteration loop on the device.
 *
 * Replaces: the MDS half of cluster()   STMiner/Algorithm/algorithm.py:54-55
 *           (sklearn.manifold.MDS(n_components, dissimilarity='precomputed').fit_transform -> smacof ->
 *            _smacof_single, scikit-learn 1.3.0 manifold/_mds.py; one call here = one _smacof_single run from a
 *            caller-supplied initial configuration, the n_init restarts are the caller's loop)
 *
 *   dissim[n,n]  symmetric, zero diagonal (the gene x gene distance matrix);  x_init[n,d] row-major start
 *   max_iter, eps as sklearn (MDS defaults 300 / 1e-3);  d = n_components <= 32
 *   rule 0: stopping rule of scikit-learn <= 1.6 (the version the reference pins):
 *           stress(X_it) / sum_i ||X_{it+1,i}|| decreased by less than eps
 *   rule 1: scikit-learn >= 1.7: (stress(X_it) - stress(X_{it+1})) / (sum_ij dis_ij^2 / 2) < eps
 *   workspace    device scratch of stm_mds_workspace_bytes(n, d) bytes
 *   x_out[n,d], out_stress (raw stress, sklearn's `stress_`), out_n_iter — device pointers
 */
int64_t stm_mds_workspace_bytes(int32_t n, int32_t d);
int stm_mds_smacof(const double* dissim, int32_t n, int32_t d, const double* x_init, int32_t max_iter,
                   double eps, int32_t rule, void* workspace, int64_t workspace_bytes,
                   double* x_out, double* out_stress, int32_t* out_n_iter, void* stream);

/*
 * Measurement utility: FP32 FMA-pipe peak micro-benchmark used as the roofline denominator of the EM
 * kernel (bench.py).  Launches `blocks` CTAs of 256 threads, each thread running `iters` rounds of 16
 * independent FFMA chains: FLOPs = blocks * 256 * iters * 32.  `out` is a device float (never written).
 */
int stm_bench_fma(int32_t blocks, int32_t iters, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STMINER_B200_H_ */
