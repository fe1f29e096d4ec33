// K2/K3/K4/K5: batched p x p Cholesky solve + Newton update, symv, OLS scaling, partition reduce.
//
// Reference arithmetic replaced:
//   del = solve(hess) %*% grad; damping; beta = beta + del        dlsa/R/wlse_func.R:55-62
//   Sig_invMcoef = Sig_inv.dot(coef)                               dlsa/models.py:131
//   groupby('par_id').sum(...)                                     dlsa/dlsa.py:30-34
//   np.linalg.lstsq(Sig_inv_sum, Sig_invMcoef_sum)                 dlsa/dlsa.py:48
//   lm(): SI = solve(vcov) = X'X / sigma2                          dlsa/R/wlse_func.R:169-182
#include "common.cuh"

namespace dlsa {

constexpr int kCholThreads = 256;
constexpr int kCholSmemMaxP = 160;  // p*p*8 <= 204,800 B fits one CTA's shared memory

// In-place Cholesky A = L L' on the lower triangle of L[p][ld] (right-looking, one CTA), then
// x = A^-1 b by forward/back substitution.  Returns false (in *ok_s) if a pivot is not positive/finite.
__device__ void chol_solve_cta(double* L, int ld, int p, double* x, int* ok_s) {
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) *ok_s = 1;
    __syncthreads();
    for (int j = 0; j < p; j++) {
        const double d = L[(size_t)j * ld + j];
        if (!(d > 0.0) || !isfinite(d)) {
            if (tid == 0) *ok_s = 0;
            __syncthreads();
            return;
        }
        const double dj = sqrt(d);
        __syncthreads();  // everyone has read the pivot before it is overwritten
        for (int i = j + tid; i < p; i += nt) L[(size_t)i * ld + j] = (i == j) ? dj : L[(size_t)i * ld + j] / dj;
        __syncthreads();
        // trailing update of the lower triangle: L[i][c] -= L[i][j] * L[c][j], j < c <= i < p
        const int m = p - j - 1;
        for (int e = tid; e < m * m; e += nt) {
            const int ii = e / m, cc = e - ii * m;
            if (cc <= ii) {
                const int i = j + 1 + ii, c = j + 1 + cc;
                L[(size_t)i * ld + c] = fma(-L[(size_t)i * ld + j], L[(size_t)c * ld + j], L[(size_t)i * ld + c]);
            }
        }
        __syncthreads();
    }
    // forward: L z = b
    for (int j = 0; j < p; j++) {
        if (tid == 0) x[j] = x[j] / L[(size_t)j * ld + j];
        __syncthreads();
        const double xj = x[j];
        for (int i = j + 1 + tid; i < p; i += nt) x[i] = fma(-L[(size_t)i * ld + j], xj, x[i]);
        __syncthreads();
    }
    // backward: L' x = z
    for (int j = p - 1; j >= 0; j--) {
        if (tid == 0) x[j] = x[j] / L[(size_t)j * ld + j];
        __syncthreads();
        const double xj = x[j];
        for (int i = tid; i < j; i += nt) x[i] = fma(-L[(size_t)j * ld + i], xj, x[i]);
        __syncthreads();
    }
}

// mode 0: Newton update (beta += delta with damping / convergence bookkeeping); mode 1: plain solve
template <bool SMEM>
__global__ void __launch_bounds__(kCholThreads) chol_kernel(const double* __restrict__ A, const double* __restrict__ b,
                                                            int K, int p, int mode, double* beta, double* x_out,
                                                            double* max_delta, int* status, int* active, int* iters,
                                                            int* n_active, double tol, double damp_above,
                                                            double damp_factor, double* gws) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int ok_s;
    __shared__ double md_s[kCholThreads / 32];
    const int k = blockIdx.x;
    if (mode == 0 && active != nullptr && active[k] == 0) return;
    double* xs = reinterpret_cast<double*>(smem_raw);  // [p]
    double* L;
    int ld;
    if (SMEM) {
        L = xs + ((p + 1) & ~1);
        ld = p | 1;  // odd leading dimension: column walks hit distinct banks
    } else {
        L = gws + (size_t)k * p * p;
        ld = p;
    }
    const double* Ak = A + (size_t)k * p * p;
    for (int e = threadIdx.x; e < p * p; e += blockDim.x) {
        const int i = e / p, c = e - i * p;
        if (c <= i) L[(size_t)i * ld + c] = Ak[e];
    }
    for (int i = threadIdx.x; i < p; i += blockDim.x) xs[i] = b[(size_t)k * p + i];
    __syncthreads();
    chol_solve_cta(L, ld, p, xs, &ok_s);
    const bool ok = ok_s != 0;
    if (mode == 1) {
        for (int i = threadIdx.x; i < p; i += blockDim.x) x_out[(size_t)k * p + i] = ok ? xs[i] : 0.0;
        if (threadIdx.x == 0 && status != nullptr) status[k] = ok ? DLSA_ST_OK : DLSA_ST_NOT_SPD;
        return;
    }
    // Newton bookkeeping
    double md = 0.0;
    bool finite = true;
    for (int i = threadIdx.x; i < p; i += blockDim.x) {
        const double v = fabs(xs[i]);
        finite = finite && isfinite(v);
        md = fmax(md, v);
    }
    if (!finite) md = INFINITY;
    for (int m = 16; m >= 1; m >>= 1) md = fmax(md, __shfl_xor_sync(0xffffffffu, md, m));
    if ((threadIdx.x & 31) == 0) md_s[threadIdx.x >> 5] = md;
    __syncthreads();
    md = md_s[0];
    for (int w = 1; w < kCholThreads / 32; w++) md = fmax(md, md_s[w]);
    if (!ok || !isfinite(md)) {
        if (threadIdx.x == 0) {
            status[k] = DLSA_ST_NOT_SPD;
            active[k] = 0;
            if (max_delta) max_delta[k] = INFINITY;
        }
        return;
    }
    const double scale = (md > damp_above) ? damp_factor : 1.0;  // dlsa/R/wlse_func.R:56-60
    // A step with max|delta| <= tol is below the solver's resolution: it is NOT applied, so that the H of this
    // pass is Sig_inv evaluated exactly at the returned coefficients (dlsa/models.py:114,130) and the caller
    // needs no extra pass for it.
    if (md > tol)
        for (int i = threadIdx.x; i < p; i += blockDim.x) beta[(size_t)k * p + i] += scale * xs[i];
    if (threadIdx.x == 0) {
        if (max_delta) max_delta[k] = md;
        iters[k] += 1;
        if (md <= tol) {
            active[k] = 0;
            status[k] = DLSA_ST_OK;
        } else {
            status[k] = DLSA_ST_NOT_CONVERGED;
            atomicAdd(n_active, 1);
        }
    }
}

static size_t chol_smem_bytes(int p) { return ((size_t)((p + 1) & ~1) + (size_t)p * (p | 1)) * sizeof(double); }

size_t chol_workspace_bytes(int K, int p) {
    return p <= kCholSmemMaxP ? 0 : (size_t)K * p * p * sizeof(double);
}

int chol_launch(const double* A, const double* b, int K, int p, int mode, double* beta, double* x_out,
                double* max_delta, int* status, int* active, int* iters, int* n_active, double tol, double damp_above,
                double damp_factor, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    if (K == 0) return 0;
    if (p <= kCholSmemMaxP) {
        const size_t smem = chol_smem_bytes(p);
        DLSA_CUDA(cudaFuncSetAttribute(chol_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        chol_kernel<true><<<K, kCholThreads, smem, st>>>(A, b, K, p, mode, beta, x_out, max_delta, status, active, iters,
                                                         n_active, tol, damp_above, damp_factor, nullptr);
    } else {
        const size_t need = chol_workspace_bytes(K, p);
        DLSA_REQUIRE(workspace != nullptr && workspace_bytes >= need, DLSA_E_WORKSPACE,
                     "cholesky: workspace too small (%zu < %zu)", workspace_bytes, need);
        const size_t smem = (size_t)((p + 1) & ~1) * sizeof(double);
        chol_kernel<false><<<K, kCholThreads, smem, st>>>(A, b, K, p, mode, beta, x_out, max_delta, status, active,
                                                          iters, n_active, tol, damp_above, damp_factor,
                                                          static_cast<double*>(workspace));
    }
    count_launch();
    return check_launch("chol_kernel");
}

// out_k = H_k beta_k  (one warp per output row)
__global__ void symv_kernel(const double* __restrict__ H, const double* __restrict__ beta, int K, int p,
                            double* __restrict__ out) {
    const int k = blockIdx.y;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= p) return;
    const double* row = H + ((size_t)k * p + warp) * p;
    const double* bk = beta + (size_t)k * p;
    double s = 0.0;
    for (int j = lane; j < p; j += 32) s = fma(row[j], bk[j], s);
    for (int m = 16; m >= 1; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
    if (lane == 0) out[(size_t)k * p + warp] = s;
}

int symv_launch(const double* H, const double* beta, int K, int p, double* out, cudaStream_t st) {
    if (K == 0) return 0;
    dim3 grid((p * 32 + 255) / 256, K);
    symv_kernel<<<grid, 256, 0, st>>>(H, beta, K, p, out);
    count_launch();
    return check_launch("symv_kernel");
}

// sigma2_k = (yty - theta'Xty)/(n_k - p); XtX /= sigma2; Xty /= sigma2     (dlsa/R/wlse_func.R:169-182)
__global__ void ols_scale_kernel(double* XtX, double* Xty, const double* __restrict__ yty,
                                 const double* __restrict__ theta, const int64_t* __restrict__ part_off, int K, int p,
                                 double* sigma2) {
    __shared__ double s2_s;
    __shared__ double red[8];
    const int k = blockIdx.x;
    double s = 0.0;
    for (int i = threadIdx.x; i < p; i += blockDim.x) s = fma(theta[(size_t)k * p + i], Xty[(size_t)k * p + i], s);
    for (int m = 16; m >= 1; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) tot += red[w];
        const double nk = (double)(part_off[k + 1] - part_off[k]);
        s2_s = (yty[k] - tot) / (nk - (double)p);
        if (sigma2) sigma2[k] = s2_s;
    }
    __syncthreads();
    const double inv = 1.0 / s2_s;
    for (int e = threadIdx.x; e < p * p; e += blockDim.x) XtX[(size_t)k * p * p + e] *= inv;
    for (int e = threadIdx.x; e < p; e += blockDim.x) Xty[(size_t)k * p + e] *= inv;
}

int ols_scale_launch(double* XtX, double* Xty, const double* yty, const double* theta, const int64_t* part_off, int K,
                     int p, double* sigma2, cudaStream_t st) {
    if (K == 0) return 0;
    ols_scale_kernel<<<K, 256, 0, st>>>(XtX, Xty, yty, theta, part_off, K, p, sigma2);
    count_launch();
    return check_launch("ols_scale_kernel");
}

// packed = [ sum_k H_k | sum_k Hbeta_k | sum_k beta_k | K_eff | K ], fixed summation order k = 0..K-1
__global__ void partition_reduce_kernel(const double* __restrict__ H, const double* __restrict__ Hb,
                                        const double* __restrict__ beta, const int* __restrict__ status, int K, int p,
                                        double* __restrict__ packed) {
    const int total = p * p + 2 * p + 2;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        double v = 0.0;
        if (e < p * p) {
            for (int k = 0; k < K; k++)
                if (status == nullptr || status[k] == DLSA_ST_OK) v += H[(size_t)k * p * p + e];
        } else if (e < p * p + p) {
            const int i = e - p * p;
            for (int k = 0; k < K; k++)
                if (status == nullptr || status[k] == DLSA_ST_OK) v += Hb[(size_t)k * p + i];
        } else if (e < p * p + 2 * p) {
            const int i = e - p * p - p;
            for (int k = 0; k < K; k++)
                if (status == nullptr || status[k] == DLSA_ST_OK) v += beta[(size_t)k * p + i];
        } else if (e == p * p + 2 * p) {
            for (int k = 0; k < K; k++) v += (status == nullptr || status[k] == DLSA_ST_OK) ? 1.0 : 0.0;
        } else {
            v = (double)K;
        }
        packed[e] = v;
    }
}

int partition_reduce_launch(const double* H, const double* Hb, const double* beta, const int* status, int K, int p,
                            double* packed, cudaStream_t st) {
    const int total = p * p + 2 * p + 2;
    const int blocks = (total + 127) / 128;
    partition_reduce_kernel<<<blocks, 128, 0, st>>>(H, Hb, beta, status, K, p, packed);
    count_launch();
    return check_launch("partition_reduce_kernel");
}

}  // namespace dlsa
