/*
 * dlsa_b200 — C ABI of the B200-native DLSA hot path (libdlsa_b200.so).
 *
 * Drop-in boundary for the per-partition fit -> one-round combine -> LARS/IC path of feng-li/dlsa.
 * Every entry point names the reference interface it replaces (paths relative to the reference tree).
 *
 * Conventions
 *   - plain C types only; all array arguments are DEVICE pointers unless the name ends in _host;
 *   - float64 everywhere (the reference's arithmetic type), row-major, int64 row offsets;
 *   - the caller owns every buffer, including workspaces (sizes from the *_workspace_bytes queries);
 *     the library never allocates device memory, never synchronises the device and never touches
 *     the default stream: work is enqueued on `stream` (a cudaStream_t passed as void*);
 *   - return value: 0 = OK, negative = argument error (DLSA_E_*), positive = cudaError_t;
 *     dlsa_last_error() gives a thread-local message for the last non-zero return;
 *   - per-partition outcomes are reported in device `status[K]` ints (DLSA_ST_*), never by exceptions.
 */
#ifndef DLSA_B200_H
#define DLSA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DLSA_ABI_VERSION 1

/* argument errors */
#define DLSA_E_BADARG (-1)
#define DLSA_E_UNSUPPORTED (-2)
#define DLSA_E_WORKSPACE (-3)

/* per-partition status codes */
#define DLSA_ST_OK 0
#define DLSA_ST_NOT_CONVERGED 1
#define DLSA_ST_NOT_SPD 2
#define DLSA_ST_SKIPPED 4

/* model family of the fused pass */
#define DLSA_FAMILY_LOGISTIC 0
#define DLSA_FAMILY_LINEAR 1
#define DLSA_FAMILY_POISSON 2  /* log link: mu = exp(eta), w = mu, r = y - mu (R glm(family = "poisson"), dlsa/R/wlse_func.R:155-168) */

/* LARS path type (dlsa/lsa.py:90 `type`) */
#define DLSA_LARS_LAR 0
#define DLSA_LARS_LASSO 1

int dlsa_abi_version(void);
const char *dlsa_last_error(void);
/* number of kernels this library has launched since it was loaded (bench.py's gpu_launches) */
unsigned long long dlsa_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * K1  fused Newton / IRLS pass over all partitions — one launch (+ a small slice-reduce when a
 *     partition is split over several CTAs).
 * Replaces, per partition k with rows [part_off[k], part_off[k+1]) of X:
 *   sklearn LogisticRegression(newton-cg).fit's gradient/Hessian sweeps  dlsa/models.py:110-113
 *   predict_proba + Sig_inv = X' diag(p(1-p)) X                          dlsa/models.py:114,130
 *   R logistic(): grad = colSums((Y-prob)*X), hess = t(X)%*%(prob*(1-prob)*X)   dlsa/R/wlse_func.R:52-54
 * family = LINEAR gives X'X, X'y, y'y instead (R lm per block, dlsa/R/wlse_func.R:169-182).
 * family = POISSON: prob -> mu = exp(eta): g = X'(y - mu), H = X' diag(mu) X (= solve(vcov(glm(y~x-1, poisson))),
 *   dlsa/R/wlse_func.R:155-168), loglik = sum y*eta - mu (without the -log(y!) constant).
 *
 *   X        [n, ldx] float64 row-major, px feature columns used (px <= ldx)
 *   y        [n] float64 labels (0/1) or responses
 *   part_off [K+1] int64 row offsets (device)
 *   fit_intercept: p = px + 1 and a synthesised all-ones column is coefficient 0 (dlsa/models.py:116-122)
 *   beta     [K, p] current coefficients (ignored for LINEAR)
 *   active   [K] int or NULL; partitions with active[k]==0 are skipped and their outputs left untouched
 *   g        [K, p]     out: X'(y - prob)            (LINEAR: X'y)
 *   H        [K, p, p]  out: X' diag(prob(1-prob)) X (LINEAR: X'X), full symmetric
 *   loglik   [K]        out: sum y*eta - log(1+exp(eta)) (LINEAR: y'y); NULL to skip computing it
 *   max_part_rows: upper bound of rows in any partition (host int; sizes the grid)
 *   workspace: partial sums of partitions that are split over several CTAs and, for logistic fits whose Gram triangle
 *     needs sibling CTAs (p > ~170), w / r / log-likelihood term of every row (24 bytes per row, sized for
 *     K * max_part_rows rows): there the link stage runs as a streaming kernel of its own ahead of the Gram kernel.
 */
size_t dlsa_irls_pass_workspace_bytes(int K, int p, int64_t max_part_rows);
int dlsa_irls_pass_f64(const double *X, int64_t ldx, const double *y, const int64_t *part_off, int K, int px,
                       int fit_intercept, int family, const double *beta, const int *active, double *g, double *H,
                       double *loglik, int64_t max_part_rows, void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * K1c the same fused pass for a design with one-hot (dummy) blocks, computed from the COMPACT representation
 *     [numeric | one code per categorical column] — the expanded 0/1 columns are never materialised.
 * Replaces the dummy branch of logistic_model: pd.get_dummies(sparse=True) -> dense design -> fit -> Sig_inv
 *                                                                            dlsa/models.py:56-91,110-114,130
 * Outputs are those of dlsa_irls_pass_f64 on the matrix dlsa_onehot_expand_f64 would build, in the same column
 * order: [intercept] numeric columns, then for c = 0..n_cat-1 the n_levels[c] level columns of categorical c
 * (dlsa/models.py:71-79).  Level-by-numeric blocks run on DMMA with a one-hot operand (level tiles untouched by a
 * 4-row step are skipped), level-by-level counts and the level gradient are 64-bit fixed-point sums (order-independent).
 *
 *   numeric  [n, ldn] float64, n_num <= 7 columns used; mean/std [n_num] or both NULL: x <- (x-mean)/std (:99-101)
 *   codes    [n, n_cat] int32, codes[i,c] in [0, n_levels[c]) or -1 for the baseline level (no column)
 *   n_levels_host [n_cat] HOST ints (1 <= n_cat <= 8, sum <= 192)
 *   p = fit_intercept + n_num + sum(n_levels);  beta/g [K,p], H [K,p,p], loglik [K] or NULL; active as in K1.
 * DLSA_E_UNSUPPORTED when the shape limits above are exceeded (expand and use dlsa_irls_pass_f64 instead).
 */
size_t dlsa_irls_pass_cat_workspace_bytes(int K, int n_cat, const int *n_levels_host, int64_t max_part_rows);
int dlsa_irls_pass_cat_f64(const double *numeric, int64_t ldn, int n_num, const double *mean, const double *std,
                           const int *codes, int n_cat, const int *n_levels_host, const double *y,
                           const int64_t *part_off, int K, int fit_intercept, const double *beta, const int *active,
                           double *g, double *H, double *loglik, int64_t max_part_rows, void *workspace,
                           size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * K2  batched p x p Cholesky solve + Newton update for every active partition.
 * Replaces  del = solve(hess)%*%grad; damping; beta = beta + del; stop rule   dlsa/R/wlse_func.R:55-62
 *           (and sklearn's inner CG + line search, dlsa/models.py:110-113).
 *   delta = H^-1 g;  md = max|delta|;  if md > damp_above: delta *= damp_factor;
 *   md >  tol: beta += delta, partition stays active;
 *   md <= tol: converged, active[k] = 0 and beta is LEFT at the point H and g were evaluated at (the step is below
 *              the solver's resolution), so H is already Sig_inv at the returned coefficients (dlsa/models.py:114,130);
 *   iters[k] += 1;  H not SPD / non-finite => status NOT_SPD, active 0.
 *   n_active (device int, caller zeroes it) is incremented once per partition still active afterwards.
 */
size_t dlsa_newton_update_workspace_bytes(int K, int p);
int dlsa_newton_update_f64(const double *H, const double *g, int K, int p, double *beta, double *max_delta,
                           int *status, int *active, int *iters, int *n_active, double tol, double damp_above,
                           double damp_factor, void *workspace, size_t workspace_bytes, void *stream);

/* K5  plain batched SPD solve  x_k = A_k^-1 b_k  (np.linalg.lstsq on the SPD sum, dlsa/dlsa.py:48;
 *     R solve(Sig_inv_sum), dlsa/R/wlse_func.R:200).  status[k] = NOT_SPD when the factorisation fails. */
int dlsa_spd_solve_f64(const double *A, const double *b, int K, int p, double *x, int *status, void *workspace,
                       size_t workspace_bytes, void *stream);

/* K3  batched symmetric matrix-vector product  out_k = H_k beta_k   (Sig_invMcoef, dlsa/models.py:131) */
int dlsa_batched_symv_f64(const double *H, const double *beta, int K, int p, double *out, void *stream);

/* OLS block epilogue (R lm + solve(vcov), dlsa/R/wlse_func.R:169-182): given XtX, Xty, yty, rows n_k and
 * theta_k = XtX^-1 Xty (from dlsa_spd_solve_f64): sigma2_k = (yty - theta'Xty)/(n_k - p);
 * Sig_inv_k = XtX/sigma2_k (in place), Sig_invMcoef_k = Xty/sigma2_k (in place). */
int dlsa_ols_scale_f64(double *XtX, double *Xty, const double *yty, const double *theta, const int64_t *part_off,
                       int K, int p, double *sigma2, void *stream);

/* ---------------------------------------------------------------------------------------------
 * K4  one-round combine, local part: fixed-order sum over this rank's partitions.
 * Replaces Spark  groupby('par_id').sum(...)                                  dlsa/dlsa.py:30-34
 *   packed_out [p*p + 2p + 2] = [ sum_k H_k | sum_k Hbeta_k | sum_k beta_k | K_eff | K_total ]
 *   partitions with status != OK contribute nothing to the first three blocks or to K_eff
 *   (the reference's all-zero frame for a skipped chunk, dlsa/models.py:84-91).
 * The cross-GPU step is ONE allreduce(sum) of packed_out, issued by the caller on the same stream.
 */
int dlsa_partition_reduce_f64(const double *H, const double *Hbeta, const double *beta, const int *status, int K,
                              int p, double *packed_out, void *stream);

/* ---------------------------------------------------------------------------------------------
 * K6+K7  LARS / lasso path on the least-squares approximation and its information criteria.
 * Replaces lars_lsa(Sigma0, b0, intercept, n, type, eps, max_steps)          dlsa/lsa.py:90-212
 *          (intercept branch as in the R original, dlsa/R/dlsa_alasso_func.R:115-125,243-245)
 *      and the first-argmin AIC/BIC selection of dlsa()                       dlsa/dlsa.py:90-91
 *   Sigma0 [p0,p0], b0 [p0]; m = p0 - (intercept?1:0); max_steps < 0 means 8*m (0 = only the all-zero first row)
 *   beta_path [(max_steps+1), m], beta0_path/rss/aic/bic [(max_steps+1)], dof [(max_steps+1)] int
 *   info [4] int: {n_rows (= steps+1), argmin AIC, argmin BIC, error flag}
 */
size_t dlsa_lars_path_workspace_bytes(int p0, int max_steps);
int dlsa_lars_path_f64(const double *Sigma0, const double *b0, int p0, int intercept, double n, int type, double eps,
                       int max_steps, double *beta_path, double *beta0_path, double *rss, double *aic, double *bic,
                       int *dof, int *info, void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * f1  log-likelihood of q candidate coefficient vectors over all partitions
 * Replaces logistic_model_eval                                               dlsa/models.py:151-225
 *   par [p, q] (column i is one estimator), out [K, q] per-partition log-likelihoods.
 */
size_t dlsa_logistic_loglik_workspace_bytes(int K, int q);
int dlsa_logistic_loglik_f64(const double *X, int64_t ldx, const double *y, const int64_t *part_off, int K, int px,
                             int fit_intercept, const double *par, int q, double *out, void *workspace,
                             size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * a1  synthetic data with the law of simulate_logistic (dlsa/models.py:6-40), generated on device by
 *     a counter-based generator keyed by (seed, global row index) so any sharding sees the same rows.
 *   law 0: X ~ U(-0.5, 0.5), y ~ Bernoulli(sigmoid(X beta_true))        (dlsa/models.py:22-30)
 *   law 1: X ~ N(0,1),       y = X beta_true + N(0,1)                    (dlsa/R/simulator.R, linear)
 *   rows [row0, row0 + n) of the global stream are written to X [n, ldx], y [n].
 */
int dlsa_simulate_f64(double *X, int64_t ldx, double *y, int64_t row0, int64_t n, int px, const double *beta_true,
                      uint64_t seed, int law, void *stream);

/* Measurement aid (bench.py): a registers-only DMMA.8x8x4 stream, 16 warps per CTA, `ctas` CTAs (one per SM), `iters`
 * x 8 independent MMAs per warp.  sink [ctas * 512] doubles; *flops_out (host) = floating-point operations the launch
 * performs.  Timed with CUDA events it gives the FP64 tensor-pipe peak of THIS device at THIS moment — the roofline
 * denominator of the fused pass (MEASURED_PEAKS.json carries no FP64 entry). */
int dlsa_probe_fp64_dmma(int ctas, int iters, double *sink, double *flops_out, void *stream);

/* Benchmark input for BASELINE.json configs[2] (airline-shaped data, SURVEY.md section 8d "C3"): the COMPACT design
 *   numeric [n, ldn] (n_num columns ~ N(0,1)), codes [n, n_cat] int32 (level - 1; -1 = the baseline level 0), y [n],
 *   rows [row0, row0 + n) of a counter-based stream keyed by (seed, global row), as dlsa_simulate_f64.
 *   levels_dev [n_cat] = levels per column incl. the baseline; cdf_dev = their cumulative shares, concatenated;
 *   eta = beta0 + numeric . beta_num + sum_c beta_lev[level_c]  (beta_lev: non-baseline levels, concatenated);
 *   y ~ Bernoulli(sigmoid(eta)).  Stands in for the cleaned airline frame of projects/logistic_dlsa.py:213-237
 *   after dummies.select_dummy_factors; it is input generation, not part of the timed path. */
int dlsa_simulate_airline_f64(double *numeric, int64_t ldn, int *codes, double *y, int64_t row0, int64_t n, int n_num,
                              int n_cat, const int *levels_dev, const double *cdf_dev, double beta0,
                              const double *beta_num_dev, const double *beta_lev_dev, uint64_t seed, void *stream);

/* a11  dense one-hot expansion with the layout of logistic_model's dummy branch (dlsa/models.py:56-79):
 *   X[i, col_base[c] + codes[i, c]] = 1 for every categorical column c whose code is >= 0
 *   (code -1 = baseline level, no column).  numeric [n, n_num] is copied (optionally standardised with
 *   mean/std, dlsa/models.py:99-101) to columns [0, n_num).  X [n, ldx] must be zero-filled by the caller. */
int dlsa_onehot_expand_f64(const double *numeric, int n_num, const double *mean, const double *std, const int *codes,
                           int n_cat, const int *col_base, int64_t n, double *X, int64_t ldx, void *stream);

/* Diagnostic (host only, no GPU needed): the block schedule of the tiled fused pass for p = px + fit_intercept in
 *   57..512.  sched_out [16 * 12 * 3] words, entry (sibling CTA, MMA warp, slot): bits 0-7 = valid-tile mask of a block
 *   of 2 tile rows x 4 tile columns (bit 4*ii + jj), 8-15 = row pair m (tile rows 2m, 2m+1), 16-23 = column group J
 *   (tile columns 4J..4J+3), 24 = owns the gradient of its column group, 0 = empty slot.
 *   info [8]: tile rows NT, sibling CTAs TG, rows per chunk, ring stages, streaming-link flag, row pitch (doubles),
 *   dynamic shared memory (bytes), slots per CTA. */
int dlsa_irls_tiled_schedule(int px, int fit_intercept, uint32_t *sched_out, int *info);

/* Host-only diagnostic (no GPU work): the shared-memory layout the categorical pass (f-2) chooses for a shape.
 *   info [12]: slices per partition, warps per CTA, 1 if warp pairs share one set of level tables, wide (scattered)
 *   categoricals, scatter rounds, DMMA level tiles, bytes of one set of tables, bytes of one ring stage, bytes of a warp's
 *   ring + row scratch, dynamic shared memory per CTA, cells of the pair histogram, padded levels.
 *   cat_out [n_cat][4]: wide slot (-1 = narrow: on DMMA tiles), row of the per-row integer scratch, byte offset of the
 *   categorical's level 0 inside the tables, byte offset of its sink.  tile_out [48][3]: for each of the 6 x 8 tile
 *   slots the scratch row of its categorical, the code it stands for (-2 = unused slot) and its level (level space). */
int dlsa_irls_cat_plan(int K, int n_num, int n_cat, const int *n_levels_host, int64_t max_part_rows, int *info,
                       int *cat_out, int *tile_out);

/* f-3  describe(): count, mean, sample standard deviation, min, max of every column of X [n, ldx] in one streaming
 *   pass — the statistics the reference takes from Spark's data_sdf_i.describe() (projects/logistic_dlsa.py:279-283)
 *   and applies as (x - mean) / stddev in logistic_model (dlsa/models.py:99-101).
 *   stats [5, px] out: rows count, mean, stddev (n-1 denominator), min, max.  n >= 1. */
size_t dlsa_column_stats_workspace_bytes(int px);
int dlsa_column_stats_f64(const double *X, int64_t ldx, int64_t n, int px, double *stats, void *workspace,
                          size_t workspace_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DLSA_B200_H */
