// extern "C" surface of libdlsa_b200.so — see include/dlsa_b200.h for the contract of every symbol.
#include <atomic>
#include <cstdarg>
#include <cstring>

#include "common.cuh"

namespace dlsa {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s launch failed: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

// implemented in the kernel translation units
int irls_small_launch(const double*, int64_t, const double*, const int64_t*, int, int, int, int, const double*,
                      const int*, double*, double*, double*, int64_t, void*, size_t, cudaStream_t);
size_t irls_small_workspace_bytes(int K, int p, int64_t max_part_rows);
int irls_tiled_launch(const double*, int64_t, const double*, const int64_t*, int, int, int, int, const double*,
                      const int*, double*, double*, double*, int64_t, void*, size_t, cudaStream_t);
size_t irls_tiled_workspace_bytes(int K, int p, int64_t max_part_rows);
size_t chol_workspace_bytes(int K, int p);
int chol_launch(const double*, const double*, int, int, int, double*, double*, double*, int*, int*, int*, int*, double,
                double, double, void*, size_t, cudaStream_t);
int symv_launch(const double*, const double*, int, int, double*, cudaStream_t);
int ols_scale_launch(double*, double*, const double*, const double*, const int64_t*, int, int, double*, cudaStream_t);
int partition_reduce_launch(const double*, const double*, const double*, const int*, int, int, double*, cudaStream_t);
size_t lars_workspace_bytes(int p0, int max_steps);
int lars_launch(const double*, const double*, int, int, double, int, double, int, double*, double*, double*, double*,
                double*, int*, int*, void*, size_t, cudaStream_t);
int simulate_launch(double*, int64_t, double*, int64_t, int64_t, int, const double*, uint64_t, int, cudaStream_t);
int airline_launch(double*, int64_t, int*, double*, int64_t, int64_t, int, int, const int*, const double*, double,
                   const double*, const double*, uint64_t, cudaStream_t);
int fp64_probe_launch(int, int, double*, cudaStream_t);
int loglik_slices(int K);
int loglik_launch(const double*, int64_t, const double*, const int64_t*, int, int, int, const double*, int, double*,
                  double*, cudaStream_t);
int onehot_launch(const double*, int, const double*, const double*, const int*, int, const int*, int64_t, double*,
                  int64_t, cudaStream_t);

int irls_tiled_schedule_info(int px, int icpt, uint32_t* sched_out, int* info);
size_t colstats_workspace_bytes(int px);
int colstats_launch(const double*, int64_t, int64_t, int, double*, double*, cudaStream_t);

size_t irls_cat_workspace_bytes(int K, int n_cat, const int* n_levels, int64_t max_part_rows);
int irls_cat_plan_info(int K, int n_num, int n_cat, const int* n_levels, int64_t max_part_rows, int* info, int* cat_out,
                       int* tile_out);
int irls_cat_launch(const double*, int64_t, int, const double*, const double*, const int*, int, const int*, const double*,
                    const int64_t*, int, int, const double*, const int*, double*, double*, double*, int64_t, void*,
                    size_t, cudaStream_t);

constexpr int kSmallMaxP = 56;

}  // namespace dlsa

using namespace dlsa;

extern "C" {

int dlsa_abi_version(void) { return DLSA_ABI_VERSION; }
const char* dlsa_last_error(void) { return g_err; }
unsigned long long dlsa_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

size_t dlsa_irls_pass_workspace_bytes(int K, int p, int64_t max_part_rows) {
    if (K <= 0 || p <= 0) return 0;
    return p <= kSmallMaxP ? irls_small_workspace_bytes(K, p, max_part_rows)
                           : irls_tiled_workspace_bytes(K, p, max_part_rows);
}

int dlsa_irls_pass_f64(const double* X, int64_t ldx, const double* y, const int64_t* part_off, int K, int px,
                       int fit_intercept, int family, const double* beta, const int* active, double* g, double* H,
                       double* loglik, int64_t max_part_rows, void* workspace, size_t workspace_bytes, void* stream) {
    DLSA_REQUIRE(K >= 0 && px >= 1 && ldx >= px, DLSA_E_BADARG, "irls pass: bad shape K=%d px=%d ldx=%lld", K, px,
                 (long long)ldx);
    DLSA_REQUIRE(family == DLSA_FAMILY_LOGISTIC || family == DLSA_FAMILY_LINEAR || family == DLSA_FAMILY_POISSON, DLSA_E_BADARG,
                 "irls pass: unknown family %d", family);
    DLSA_REQUIRE(X && y && part_off && g && H, DLSA_E_BADARG, "irls pass: null pointer");
    DLSA_REQUIRE(family == DLSA_FAMILY_LINEAR || beta != nullptr, DLSA_E_BADARG, "irls pass: beta is null");
    DLSA_REQUIRE(max_part_rows >= 0, DLSA_E_BADARG, "irls pass: max_part_rows < 0");
    if (K == 0) return 0;
    const int p = px + (fit_intercept ? 1 : 0);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (p <= kSmallMaxP)
        return irls_small_launch(X, ldx, y, part_off, K, px, fit_intercept, family, beta, active, g, H, loglik,
                                 max_part_rows, workspace, workspace_bytes, st);
    return irls_tiled_launch(X, ldx, y, part_off, K, px, fit_intercept, family, beta, active, g, H, loglik,
                             max_part_rows, workspace, workspace_bytes, st);
}

size_t dlsa_irls_pass_cat_workspace_bytes(int K, int n_cat, const int* n_levels_host, int64_t max_part_rows) {
    if (K <= 0 || n_cat <= 0 || n_cat > 8 || n_levels_host == nullptr) return 0;
    return irls_cat_workspace_bytes(K, n_cat, n_levels_host, max_part_rows);
}

int dlsa_irls_pass_cat_f64(const double* numeric, int64_t ldn, int n_num, const double* mean, const double* std,
                           const int* codes, int n_cat, const int* n_levels_host, const double* y,
                           const int64_t* part_off, int K, int fit_intercept, const double* beta, const int* active,
                           double* g, double* H, double* loglik, int64_t max_part_rows, void* workspace,
                           size_t workspace_bytes, void* stream) {
    DLSA_REQUIRE(K >= 0 && n_num >= 0 && ldn >= n_num && n_cat >= 0 && max_part_rows >= 0, DLSA_E_BADARG,
                 "categorical pass: bad shape K=%d n_num=%d n_cat=%d", K, n_num, n_cat);
    DLSA_REQUIRE(n_num <= 7 && n_cat >= 1 && n_cat <= 8, DLSA_E_UNSUPPORTED,
                 "categorical pass: needs 1..8 categorical and at most 7 numeric columns (got %d, %d); expand with "
                 "dlsa_onehot_expand_f64 and use dlsa_irls_pass_f64",
                 n_cat, n_num);
    DLSA_REQUIRE((n_num == 0 || numeric) && codes && n_levels_host && y && part_off && beta && g && H, DLSA_E_BADARG,
                 "categorical pass: null pointer");
    DLSA_REQUIRE((mean == nullptr) == (std == nullptr), DLSA_E_BADARG, "categorical pass: mean/std must both be given");
    for (int c = 0; c < n_cat; ++c)
        DLSA_REQUIRE(n_levels_host[c] >= 1, DLSA_E_BADARG, "categorical pass: column %d has %d levels", c, n_levels_host[c]);
    if (K == 0) return 0;
    return irls_cat_launch(numeric, ldn, n_num, mean, std, codes, n_cat, n_levels_host, y, part_off, K, fit_intercept, beta,
                           active, g, H, loglik, max_part_rows, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

size_t dlsa_newton_update_workspace_bytes(int K, int p) { return (K <= 0 || p <= 0) ? 0 : chol_workspace_bytes(K, p); }

int dlsa_newton_update_f64(const double* H, const double* g, int K, int p, double* beta, double* max_delta,
                           int* status, int* active, int* iters, int* n_active, double tol, double damp_above,
                           double damp_factor, void* workspace, size_t workspace_bytes, void* stream) {
    DLSA_REQUIRE(K >= 0 && p >= 1, DLSA_E_BADARG, "newton update: bad shape");
    DLSA_REQUIRE(H && g && beta && status && active && iters && n_active, DLSA_E_BADARG, "newton update: null pointer");
    return chol_launch(H, g, K, p, 0, beta, nullptr, max_delta, status, active, iters, n_active, tol, damp_above,
                       damp_factor, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int dlsa_spd_solve_f64(const double* A, const double* b, int K, int p, double* x, int* status, void* workspace,
                       size_t workspace_bytes, void* stream) {
    DLSA_REQUIRE(K >= 0 && p >= 1, DLSA_E_BADARG, "spd solve: bad shape");
    DLSA_REQUIRE(A && b && x, DLSA_E_BADARG, "spd solve: null pointer");
    return chol_launch(A, b, K, p, 1, nullptr, x, nullptr, status, nullptr, nullptr, nullptr, 0.0, 0.0, 0.0, workspace,
                       workspace_bytes, static_cast<cudaStream_t>(stream));
}

int dlsa_batched_symv_f64(const double* H, const double* beta, int K, int p, double* out, void* stream) {
    DLSA_REQUIRE(K >= 0 && p >= 1 && H && beta && out, DLSA_E_BADARG, "symv: bad argument");
    return symv_launch(H, beta, K, p, out, static_cast<cudaStream_t>(stream));
}

int dlsa_ols_scale_f64(double* XtX, double* Xty, const double* yty, const double* theta, const int64_t* part_off,
                       int K, int p, double* sigma2, void* stream) {
    DLSA_REQUIRE(K >= 0 && p >= 1 && XtX && Xty && yty && theta && part_off, DLSA_E_BADARG, "ols scale: bad argument");
    return ols_scale_launch(XtX, Xty, yty, theta, part_off, K, p, sigma2, static_cast<cudaStream_t>(stream));
}

int dlsa_partition_reduce_f64(const double* H, const double* Hbeta, const double* beta, const int* status, int K,
                              int p, double* packed_out, void* stream) {
    DLSA_REQUIRE(K >= 0 && p >= 1 && packed_out, DLSA_E_BADARG, "partition reduce: bad argument");
    DLSA_REQUIRE(K == 0 || (H && Hbeta && beta), DLSA_E_BADARG, "partition reduce: null pointer");
    return partition_reduce_launch(H, Hbeta, beta, status, K, p, packed_out, static_cast<cudaStream_t>(stream));
}

size_t dlsa_lars_path_workspace_bytes(int p0, int max_steps) { return p0 <= 0 ? 0 : lars_workspace_bytes(p0, max_steps); }

int dlsa_lars_path_f64(const double* Sigma0, const double* b0, int p0, int intercept, double n, int type, double eps,
                       int max_steps, double* beta_path, double* beta0_path, double* rss, double* aic, double* bic,
                       int* dof, int* info, void* workspace, size_t workspace_bytes, void* stream) {
    DLSA_REQUIRE(p0 >= 1 && Sigma0 && b0 && beta_path && beta0_path && rss && aic && bic && dof && info, DLSA_E_BADARG,
                 "lars: bad argument");
    DLSA_REQUIRE(type == DLSA_LARS_LAR || type == DLSA_LARS_LASSO, DLSA_E_BADARG, "lars: unknown type %d", type);
    return lars_launch(Sigma0, b0, p0, intercept, n, type, eps, max_steps, beta_path, beta0_path, rss, aic, bic, dof,
                       info, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

size_t dlsa_logistic_loglik_workspace_bytes(int K, int q) {
    if (K <= 0 || q <= 0) return 0;
    return (size_t)K * loglik_slices(K) * q * sizeof(double);
}

int dlsa_logistic_loglik_f64(const double* X, int64_t ldx, const double* y, const int64_t* part_off, int K, int px,
                             int fit_intercept, const double* par, int q, double* out, void* workspace,
                             size_t workspace_bytes, void* stream) {
    DLSA_REQUIRE(K >= 0 && px >= 1 && q >= 0 && X && y && part_off && par && out, DLSA_E_BADARG, "loglik: bad argument");
    DLSA_REQUIRE(workspace && workspace_bytes >= dlsa_logistic_loglik_workspace_bytes(K, q), DLSA_E_WORKSPACE,
                 "loglik: workspace too small");
    return loglik_launch(X, ldx, y, part_off, K, px, fit_intercept ? 1 : 0, par, q, out, static_cast<double*>(workspace),
                         static_cast<cudaStream_t>(stream));
}

int dlsa_simulate_f64(double* X, int64_t ldx, double* y, int64_t row0, int64_t n, int px, const double* beta_true,
                      uint64_t seed, int law, void* stream) {
    DLSA_REQUIRE(X && y && beta_true && n >= 0 && px >= 1 && ldx >= px && (law == 0 || law == 1), DLSA_E_BADARG,
                 "simulate: bad argument");
    return simulate_launch(X, ldx, y, row0, n, px, beta_true, seed, law, static_cast<cudaStream_t>(stream));
}

int dlsa_simulate_airline_f64(double* numeric, int64_t ldn, int* codes, double* y, int64_t row0, int64_t n, int n_num,
                              int n_cat, const int* levels_dev, const double* cdf_dev, double beta0,
                              const double* beta_num_dev, const double* beta_lev_dev, uint64_t seed, void* stream) {
    DLSA_REQUIRE(codes && y && levels_dev && cdf_dev && beta_lev_dev && n >= 0 && n_num >= 0 && n_cat >= 1 &&
                     ldn >= n_num && (n_num == 0 || (numeric && beta_num_dev)),
                 DLSA_E_BADARG, "simulate airline: bad argument");
    return airline_launch(numeric, ldn, codes, y, row0, n, n_num, n_cat, levels_dev, cdf_dev, beta0, beta_num_dev,
                          beta_lev_dev, seed, static_cast<cudaStream_t>(stream));
}

int dlsa_onehot_expand_f64(const double* numeric, int n_num, const double* mean, const double* std, const int* codes,
                           int n_cat, const int* col_base, int64_t n, double* X, int64_t ldx, void* stream) {
    DLSA_REQUIRE(X && n >= 0 && n_num >= 0 && n_cat >= 0, DLSA_E_BADARG, "onehot: bad argument");
    DLSA_REQUIRE((n_num == 0 || numeric) && (n_cat == 0 || (codes && col_base)), DLSA_E_BADARG, "onehot: null pointer");
    DLSA_REQUIRE((mean == nullptr) == (std == nullptr), DLSA_E_BADARG, "onehot: mean/std must both be given");
    return onehot_launch(numeric, n_num, mean, std, codes, n_cat, col_base, n, X, ldx, static_cast<cudaStream_t>(stream));
}

int dlsa_probe_fp64_dmma(int ctas, int iters, double* sink, double* flops_out, void* stream) {
    DLSA_REQUIRE(ctas >= 1 && iters >= 1 && sink && flops_out, DLSA_E_BADARG, "fp64 probe: bad argument");
    *flops_out = (double)ctas * 16.0 * (double)iters * 8.0 * 512.0;   // 8x8x4 FMAs = 512 flops per warp-level DMMA
    return fp64_probe_launch(ctas, iters, sink, static_cast<cudaStream_t>(stream));
}

int dlsa_irls_tiled_schedule(int px, int fit_intercept, uint32_t* sched_out, int* info) {
    DLSA_REQUIRE(sched_out && info, DLSA_E_BADARG, "tiled schedule: null pointer");
    DLSA_REQUIRE(irls_tiled_schedule_info(px, fit_intercept ? 1 : 0, sched_out, info) == 0, DLSA_E_UNSUPPORTED,
                 "tiled schedule: p must be in 57..512");
    return 0;
}

int dlsa_irls_cat_plan(int K, int n_num, int n_cat, const int* n_levels_host, int64_t max_part_rows, int* info,
                       int* cat_out, int* tile_out) {
    DLSA_REQUIRE(n_levels_host && info && cat_out && tile_out, DLSA_E_BADARG, "categorical plan: null pointer");
    DLSA_REQUIRE(irls_cat_plan_info(K, n_num, n_cat, n_levels_host, max_part_rows, info, cat_out, tile_out) == 0,
                 DLSA_E_UNSUPPORTED, "categorical plan: 1..8 categorical columns, 0..7 numeric columns");
    return 0;
}

size_t dlsa_column_stats_workspace_bytes(int px) { return px <= 0 ? 0 : colstats_workspace_bytes(px); }

int dlsa_column_stats_f64(const double* X, int64_t ldx, int64_t n, int px, double* stats, void* workspace,
                          size_t workspace_bytes, void* stream) {
    DLSA_REQUIRE(X && stats && n >= 1 && px >= 1 && ldx >= px, DLSA_E_BADARG, "column stats: bad argument");
    DLSA_REQUIRE(workspace && workspace_bytes >= colstats_workspace_bytes(px), DLSA_E_WORKSPACE,
                 "column stats: workspace too small");
    return colstats_launch(X, ldx, n, px, stats, static_cast<double*>(workspace), static_cast<cudaStream_t>(stream));
}

}  // extern "C"
