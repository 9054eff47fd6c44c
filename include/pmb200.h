/* pmb200.h -- C ABI of libpmb200.so: B200 (sm_100a) kernels for pmarlo's MSM
 * estimation hot path (featurize -> TICA -> k-means -> lagged counts ->
 * reversible MSM -> implied timescales).
 *
 * pmarlo (the reference) is pure Python and has no FFI layer of its own: the
 * arithmetic of this path lives in mdtraj / deeptime / scikit-learn.  Each
 * entry point below therefore cites the reference CALL SITE it replaces
 * (paths relative to the reference checkout, src/pmarlo/...).  The Python
 * host code in pmarlo_b200/ binds these symbols with ctypes and keeps
 * pmarlo's own call signatures on top.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in _host;
 *    matrices are row-major; outputs are caller-allocated;
 *  - no allocation, no host synchronisation (except where stated), no hidden
 *    state: re-entrant per stream;
 *  - every function returns 0 on success or a negative PMB_E* code; the
 *    message is available from pmb_last_error() (thread-local);
 *  - functions that need scratch memory take (ws, ws_bytes); the required
 *    size comes from the matching pmb_*_ws_bytes() query.
 */
#ifndef PMB200_H
#define PMB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* pmb_stream_t; /* cudaStream_t */

#define PMB_API __attribute__((visibility("default")))

#define PMB_OK 0
#define PMB_EINVAL (-1)   /* bad argument */
#define PMB_ECUDA (-2)    /* CUDA runtime / launch error */
#define PMB_EWORKSPACE (-3) /* workspace too small */
#define PMB_EUNSUPPORTED (-4)

PMB_API const char* pmb_last_error(void);
PMB_API int pmb_version(void);
/* number of kernels this library has launched since load (all threads) */
PMB_API int64_t pmb_launch_count(void);

/* ---- K1 featurize ---------------------------------------------------------
 * Replaces mdtraj.compute_phi/psi/compute_distances as called from
 * features/featurize.py:41-53, features/builtins.py:17-24,51-76 and
 * markov_state_model/_features.py:131-171, plus the cos/sin expansion of
 * api/features.py:138-180.
 * units: n_units x 8 int32 {kind(0 dihedral,1 distance), a0,a1,a2,a3,
 *        col_value, col_cos, col_sin}; a column index of -1 means "not
 *        written".  For a dihedral col_value receives the angle in (-pi,pi].
 * xyz: (n_frames, n_atoms, 3) float32 nm.  out: (n_frames, ld_out) float32. */
PMB_API int pmb_featurize(const float* xyz, int64_t n_frames, int n_atoms,
                  const int32_t* units, int n_units, int n_cols,
                  float* out, int64_t ld_out, pmb_stream_t stream);

/* cos/sin expansion of periodic columns (api/features.py:138-180), fp64:
 * out_col[j] = first output column of input column j; periodic columns write
 * cos at out_col[j] and sin at out_col[j]+1, others are copied. */
PMB_API int pmb_trig_expand(const double* X, int64_t n, int F, const uint8_t* periodic,
                    const int32_t* out_col, double* Xe, int Fe, pmb_stream_t stream);

/* ---- K2 column moments ----------------------------------------------------
 * Replaces SimpleImputer(mean) + StandardScaler statistics of
 * markov_state_model/reduction.py:13-40 and the mean/std of
 * analysis/discretize.py:446-447.
 * out: 6 x d doubles over the non-NaN entries of each column:
 *   {n_valid, shift, sum(x-shift), sum((x-shift)^2), sum e(x-shift), sum e}
 * with shift = shift_in[c] when shift_in != NULL (frame shards on several GPUs
 * share rank 0's shift so that their moments add), else X[0][c] (0 if NaN), and e_g = 2 - popcount(mask[g]&3) (0 when
 * mask is NULL): the edge-frame sums from which the TICA pair means follow. */
PMB_API size_t pmb_col_moments_ws_bytes(int d);
PMB_API int pmb_col_moments(const float* X, int64_t n, int d, int64_t ld, const uint8_t* mask,
                    const double* shift_in, double* out, void* ws, size_t ws_bytes,
                    pmb_stream_t stream);

/* ---- pair mask --------------------------------------------------------------
 * Shard semantics: lagged pairs never cross a trajectory boundary
 * (deeptime Covariance / TransitionCountEstimator over a list of arrays,
 * reduction.py:103-110, _features.py:182-202).
 * mask[g] bit0: frame g has a partner g+lag in its own segment;
 *         bit1: frame g is the partner of g-lag. */
PMB_API int pmb_pair_mask(const int64_t* seg_offsets, int n_seg, int64_t n, int lag,
                  uint8_t* mask, pmb_stream_t stream);

/* ---- K3 weighted Gram -------------------------------------------------------
 * Replaces deeptime.covariance.Covariance(lagtime, reversible=True) running
 * moments behind TICA.fit (reduction.py:103-110).
 * z = NaN ? 0 : (x - shift) * scale (fp32 conditioning transform).
 * mode 0: G += sum_g popcount(mask[g]&3) z_g z_g^T          (= X0^T X0 + Xt^T Xt)
 * mode 1: G += sum_g (mask[g]&1) (z_g - z_{g+lag})(z_g - z_{g+lag})^T
 * G: d x d fp64, overwritten.  impl: 0 auto, 1 SIMT fp32, 2 tcgen05 3xTF32, 4 tcgen05 3xTF32 on CTA pairs
 * (cta_group::2, d = 256 only; correct but measured slower than 2). */
PMB_API size_t pmb_gram_ws_bytes(int d);
PMB_API int pmb_gram(const float* X, int64_t n, int d, int64_t ld, const uint8_t* mask,
             int lag, int mode, const float* shift, const float* scale,
             double* G, void* ws, size_t ws_bytes, int impl, pmb_stream_t stream);

/* ---- scaler + covariance assembly (fp64, d-sized) -----------------------------
 * pmb_scaler_from_moments: from the K2 moments derive, per column,
 *   stats[0]=m (imputation mean == column mean), stats[1]=sc (population std,
 *   ddof 0, 1 for constant columns -- sklearn StandardScaler rule; all ones
 *   when with_std == 0), stats[2]=sample std (ddof 1, discretize.py:447),
 *   and the fp32 conditioning transform cond[0]=shift32, cond[1]=scale32 fed
 *   to pmb_gram (always 1/std so that tensor-core operands are O(1)).
 *   semantic: 1 -> preprocessing is z = (x_imputed - m)/sc (reduction.py:13-40),
 *             0 -> none (z = x; _features.py:181-231 feeds raw features).
 * pmb_tica_covariances: mean-free symmetrised C00, C0t and the pair mean mu
 *   (in z units) from the two Gram matrices, undoing the fp32 conditioning
 *   exactly: C00 = G0z/2T - mu mu^T, C0t = (G0z - G1z)/2T - mu mu^T. */
PMB_API int pmb_scaler_from_moments(const double* moments, int64_t n, int d, int semantic, int with_std,
                            double* stats /*3 x d*/, float* cond /*2 x d*/, pmb_stream_t stream);
PMB_API int pmb_tica_covariances(const double* G0, const double* G1, const double* moments,
                         const double* stats, const float* cond, int64_t n, int64_t n_pairs,
                         int d, int semantic, double* C00, double* C0t, double* mu,
                         pmb_stream_t stream);

/* ---- K4 TICA eigenproblem ----------------------------------------------------
 * Replaces deeptime.numeric.eig_corr / spd_inv_split (scipy eigh/schur) behind
 * TICA.fit: C00 = V S V^T, drop |s| <= eps, L = V S^-1/2, eigh(L^T C0t L),
 * sort by |lambda| desc, canonical signs.  evals: d, evecs: d x d (first
 * `rank` columns valid, NOT kinetic-map scaled), rank: int32[4] = {rank, Jacobi
 * sweeps of the C00 problem, sweeps of the reduced problem, reserved}. */
PMB_API size_t pmb_tica_solve_ws_bytes(int d);
PMB_API int pmb_tica_solve(const double* C00, const double* C0t, int d, double eps,
                   double* evals, double* evecs, int32_t* rank,
                   void* ws, size_t ws_bytes, pmb_stream_t stream);

/* projection operands of the fitted model (CovarianceKoopmanModel.transform,
 * reduction.py:108-110, composed with the scaler of reduction.py:13-40):
 * a = m + sc*mu (d), nanfill = imputation mean (d), W = diag(1/sc) R[:, :m],
 * columns scaled by their eigenvalue when kinetic_map != 0 (deeptime default
 * scaling="kinetic_map"); columns >= rank are zero. */
PMB_API int pmb_tica_finalize(const double* evals, const double* evecs, const double* moments,
                      const double* stats, const double* mu, int d, int m, int kinetic_map,
                      double* a, double* nanfill, double* W, pmb_stream_t stream);

/* batched symmetric eigenvalues (cyclic Jacobi), one CTA per matrix.
 * A: batch x n x n (destroyed), evals: batch x n sorted by magnitude desc. */
PMB_API size_t pmb_sym_eigvals_ws_bytes(int n, int batch);
PMB_API int pmb_sym_eigvals_batched(double* A, int n, int batch, double* evals,
                            void* ws, size_t ws_bytes, pmb_stream_t stream);

/* ---- K5 projection -----------------------------------------------------------
 * Replaces CovarianceKoopmanModel.transform (reduction.py:108-110):
 * y[r][c] = sum_j ((NaN ? nanfill_j : x[r][j]) - a_j) * W[j][c], fp64 math.
 * out_f64 = 0 -> Y float32, 1 -> Y float64; ldy in elements. */
PMB_API int pmb_project(const float* X, int64_t n, int d, int64_t ld, const double* a,
                const double* nanfill, const double* W, int m, void* Y,
                int64_t ldy, int out_f64, pmb_stream_t stream);

/* ---- K6 k-means assign + accumulate -------------------------------------------
 * Replaces deeptime KMeans cluster_loop/assign (clustering.py:349-355,605-609)
 * and sklearn KMeans.predict (analysis/discretize.py:493).
 * labels = argmin_k sum_d (y_d - c_kd)^2 evaluated in fp64, first minimum
 * wins (fp32 screening + fp64 re-check of near ties => bit-exact labels).
 * sums (K x D), counts (K), inertia (1) are ACCUMULATED (caller zeroes);
 * pass NULL for sums/counts/inertia to only assign.  y_f64: Y element type.
 * impl: 0 = auto, 1 = SIMT register-tiled kernel (kmeans.cu), 2 = tcgen05 TF32 score GEMM with the
 * argmin fused into the TMEM epilogue (kmeans_tc.cu; float32 Y, needs the workspace).  Both paths
 * return the same labels: near ties are re-evaluated in fp64.
 * hints (n, may be NULL, may alias labels): labels of a previous assignment, e.g. the last Lloyd
 * iteration.  They never change the result; the tcgen05 epilogue uses the distance to the hinted centre
 * to skip 32-centre blocks that cannot contain the winner. */
PMB_API size_t pmb_kmeans_assign_ws_bytes(int64_t n, int D, int K);
PMB_API int pmb_kmeans_assign(const void* Y, int y_f64, int64_t n, int D, int64_t ld,
                      const double* centers, int K, int32_t* labels,
                      double* sums, int64_t* counts, double* inertia,
                      int64_t* n_rechecked, const int32_t* hints, void* ws, size_t ws_bytes,
                      int impl, pmb_stream_t stream);
/* Test hook of the tcgen05 path: also writes the (n x Kpad, Kpad = K rounded up to 256) fp32 score
 * matrix |y|^2 + |c_k|^2 - 2 y.c_k as it leaves TMEM, so that the error envelope the certainty test
 * relies on can be measured (tests/test_gpu_parity.py). */
/* Debug: phase cycle counters of the last block-Jacobi solve (load, local sweeps, store, grid barrier,
 * block rounds, total) as seen by CTA 0. */
PMB_API int pmb_debug_counters_tica(int64_t* out8);
/* trace of the last pmb_tica_solve: [0..15] SM id of CTAs 0..15, then per Jacobi sweep (up to 32) the pair
 * (SM cycles, nanoseconds of the global timer) on CTA 0. */
PMB_API int pmb_debug_trace_tica(int64_t* out80);
/* Debug: role-timing counters of the last fp16 tcgen05 Gram launch (zeros unless built with -DPMB_GH_PROF). */
PMB_API int pmb_debug_counters_gram(int64_t* out16);
/* Debug: role-timing counters of the last tcgen05 assignment (zeros unless built with -DPMB_KM_PROF). */
PMB_API int pmb_debug_counters_kmeans(int64_t* out16);
PMB_API int pmb_kmeans_tc_scores(const float* Y, int64_t n, int D, int64_t ld, const double* centers,
                      int K, int32_t* labels, float* scores, void* ws, size_t ws_bytes,
                      pmb_stream_t stream);

/* centres <- sums / counts (empty cluster keeps the old centre), also returns
 * shift2 = sum_k counts_k |new_k - old_k|^2 in out_shift2 (may be NULL). */
/* Per-sample silhouette coefficients (sklearn.metrics.silhouette_samples, Euclidean) behind the automatic
 * choice of n_states: `_auto_select_n_states`, markov_state_model/clustering.py:155-250.
 * Y: n x D fp64 row-major, labels in [0, K), K <= 64, sizes[k] = members of cluster k, out: n fp64. */
PMB_API int pmb_silhouette_samples(const double* Y, int64_t n, int D, const int32_t* labels, int K,
                                   const int64_t* sizes, double* out, pmb_stream_t stream);

PMB_API int pmb_kmeans_update(double* centers, const double* sums, const int64_t* counts,
                      int K, int D, double* out_shift2, pmb_stream_t stream);

/* ---- K7 lagged transition counts ------------------------------------------------
 * Replaces deeptime TransitionCountEstimator(count_mode="sliding") at
 * _msm_utils.py:238-246, bridge.py:95-104, _estimation.py:150-156 and the Python
 * loops ck_its_selector.py:70-83, debug_export.py:385-409, discretize.py:609-645.
 * C[i][j] += #{t in segment, t % step == 0 : s_t = i, s_{t+lag} = j}, pairs with
 * an endpoint outside [0,K) are skipped.  C: K x K int64, ACCUMULATED. */
PMB_API int pmb_count_lagged(const int32_t* labels, int64_t n, const int64_t* seg_offsets,
                     int n_seg, int K, int lag, int step, int64_t* C,
                     pmb_stream_t stream);
/* weighted variant (discretize.py:631-640): C[i][j] += w[t], fp64. */
PMB_API int pmb_count_lagged_weighted(const int32_t* labels, const double* weights, int64_t n,
                              const int64_t* seg_offsets, int n_seg, int K, int lag,
                              int step, double* C, pmb_stream_t stream);

/* ---- state relabelling with frame removal (Chapman-Kolmogorov test) -------------
 * Replaces the per-frame Python loops
 *   [state_map[s] for s in traj if s in state_map]   ck_runner.py:150-153, :235-238, _ck.py:86-90
 *   macro_labels[traj]                                 ck_runner.py:204
 * labels_out receives, in order, lut[labels[t]] for every frame with 0 <= labels[t] < n_lut and
 * lut[labels[t]] >= 0; the other frames are dropped.  new_seg_offsets (n_seg + 1, device) are the shard
 * offsets of the shortened trajectories, new_seg_offsets[n_seg] = frames kept.  seg_offsets must start at
 * 0 and end at n.  labels_out has room for n entries and must not alias labels. */
PMB_API size_t pmb_relabel_compact_ws_bytes(int64_t n);
PMB_API int pmb_relabel_compact(const int32_t* labels, int64_t n, const int64_t* seg_offsets, int n_seg,
                        const int32_t* lut, int n_lut, int32_t* labels_out,
                        int64_t* new_seg_offsets, void* workspace, size_t workspace_bytes,
                        pmb_stream_t stream);

/* int64 counts -> fp64 matrix + active mask (utils/msm_utils.py:150-156):
 * active[i] = (rowsum_i + colsum_i > eps).  Cf: K x K fp64, active: K bytes. */
PMB_API int pmb_counts_active(const int64_t* C, int K, double eps, double* Cf, uint8_t* active,
                      pmb_stream_t stream);

/* Self-test of the tcgen05 plumbing (descriptor encodings, TMEM round trip): one
 * 128 x N x Kdim TF32 tile product.  mode 0: A (128 x Kdim), B (N x Kdim) row-major
 * (K-major, no swizzle: the k-means score layout); mode 1: A (Kdim x 128),
 * B (Kdim x N) row-major (MN-major, 128 B swizzle with 32 B base: the Gram layout).  D: 128 x N fp32. */
PMB_API int pmb_tc_selftest(const float* A, const float* B, int N, int Kdim, int mode, float* D,
                            pmb_stream_t stream);
/* Same product from caller-built shared-memory operand images (a_bytes / b_bytes, copied verbatim to
 * 1024-byte aligned shared memory) and explicit descriptor fields: leading / stride byte offsets, layout type
 * (0 none, 1 128B swizzle with 32B base, 2 128B, 4 64B, 6 32B), descriptor address advance per K step for
 * each operand, the 32-bit instruction descriptor, kind (0 tf32, 1 f16).  Test-only: validates the 16-bit
 * operand layouts of K6 (fp16, K-major) and K3 (bf16, MN-major) in isolation. */
PMB_API int pmb_tc_selftest_raw(const void* Aimg, uint32_t a_bytes, const void* Bimg, uint32_t b_bytes, int N,
                                int nsteps, uint32_t lbo, uint32_t sbo, uint32_t layout, uint32_t step_a,
                                uint32_t step_b, uint32_t idesc, int kind, float* D, pmb_stream_t stream);

/* Weighted 2-D histogram with np.histogram2d semantics (last bin closed, out-of-range and NaN dropped): the
 * per-frame part of generate_2d_fes, markov_state_model/free_energy.py:417-865.  x, y, w (may be NULL): n fp64;
 * H: bx x by fp64, ACCUMULATED into (zero it first); bx * by <= 16384. */
PMB_API int pmb_hist2d(const double* x, const double* y, const double* w, int64_t n, double xlo, double xhi, int bx,
                       double ylo, double yhi, int by, double* H, pmb_stream_t stream);

/* ---- Bayesian MSM: Gibbs sampler of reversible transition matrices ----------------
 * Replaces deeptime BayesianMSM(lagtime, n_samples).fit(dtrajs) / TransitionMatrixSampler(reversible=True)
 * (_its.py:272-357, 670-740).  C: batch x K x K counts (fp64, zero rows / columns outside the connected set),
 * X: batch x K x K, in = diag(pi) T of the reversible MLE, out = final chain state; Tout: batch x n_samples x K x K
 * transition-matrix samples (identity rows for states without counts), pi_out: batch x n_samples x K their
 * stationary vectors (may be NULL).  One CTA per chain; n_steps Gibbs sweeps per sample (deeptime: sqrt(K)). */
PMB_API int pmb_bayes_rev_sample(const double* C, double* X, int K, int batch, int n_samples, int n_steps,
                                 uint64_t seed, double* Tout, double* pi_out, pmb_stream_t stream);

/* ---- K8 reversible maximum-likelihood MSM -----------------------------------------
 * Replaces deeptime MaximumLikelihoodMSM(reversible=True).fit(counts)
 * (_msm_utils.py:255-261, ck_its_selector.py:397-399): fixed point on the
 * row-sum vector, err = max_i |x_i-x'_i| / (0.5(x_i+x'_i)).
 * C: batch x K x K fp64.  active: batch x K bytes or NULL; states with
 * active == 0 are excluded from the estimate (T_ii = 1, pi_i = 0) -- the way
 * the caller restricts to the largest connected set (utils/scc.py:69-130).
 * alpha is added to every cell of the active block first (the Dirichlet
 * pseudocount of utils/msm_utils.py:129-167, 1e-3 in build_simple_msm; 0 for
 * the raw-count ITS path).
 * The active block must be strongly connected with positive row sums
 * (info[1] = -1 otherwise).  T: batch x K x K, pi: batch x K,
 * info: batch x {iterations, converged} int64. */
PMB_API size_t pmb_mle_rev_ws_bytes(int K, int batch);
PMB_API int pmb_mle_rev(const double* C, const uint8_t* active, int K, int batch, double alpha,
                double maxerr, int64_t maxiter, double* T, double* pi, int64_t* info,
                void* ws, size_t ws_bytes, pmb_stream_t stream);

/* Diagnostics: cycle counters of the last cooperative MLE run, HOST pointer to 8 int64:
 * {iterations, q phase, row phase, exchange, norm/err phase, CTAs, register path, 0}.
 * Synchronises the device. */
PMB_API int pmb_debug_counters(int64_t* out8_host);

/* ---- K9 leading eigenvalues of a reversible T ----------------------------------------
 * Replaces deeptime.markov.tools.analysis.eigenvalues(T, k, reversible=True, mu)
 * (_its.py:561-588,745-786): spectrum of D^1/2 T D^-1/2, sorted by magnitude.
 * K <= 256: batched Jacobi (one CTA per matrix); larger K: Lanczos with full
 * reorthogonalisation (max_steps Lanczos steps, 0 = automatic).
 * T: batch x K x K, pi: batch x K, evals: batch x k.
 * info: batch x {lanczos steps (0 = Jacobi), converged} int64. */
PMB_API size_t pmb_eig_rev_topk_ws_bytes(int K, int k, int batch, int max_steps);
PMB_API int pmb_eig_rev_topk(const double* T, const double* pi, int K, int k, int batch, int max_steps,
                     double* evals, int64_t* info,
                     void* ws, size_t ws_bytes, pmb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PMB200_H */
