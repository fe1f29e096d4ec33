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

constexpr int kCholThreads = 256;          // p <= 160: factor in shared memory
constexpr int kCholThreadsBlocked = 512;   // p > 160: 16 warps hide the global-memory latency of the trailing update
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


// ---- p > 160: the factor does not fit shared memory; blocked right-looking Cholesky on the copy in global memory ------
// Panels of 32 columns: (1) the 32x32 diagonal block is factored in shared memory by one warp, (2) every row below
// solves its 32-column slice against it (one thread per row, the row in registers) and leaves it in a shared-memory
// panel P, (3) the trailing lower triangle gets  A22 -= P P'  as 8x8 tiles on DMMA (mma.sync m8n8k4 f64; fragments from
// P with a conflict-free pitch, C tiles straight from / to global memory, kTIF independent tiles in flight per warp).
// 2/3 of the flops of a p=500 factorisation are step (3).  The unblocked rank-1 version this replaces took 68 ms for
// 1024 matrices of p=500 (C5); see profiles/r01_c5_breakdown.md.
constexpr int kNB = 32;
constexpr int kDPitch = 33;
constexpr int kTIF = 8;      // 8x8 tiles in flight per warp in the trailing update (independent C loads and DMMA chains)
constexpr int kPPitch = 36;  // == 4 (mod 16) doubles: the (row a8, column q) fragment loads hit 16 distinct banks per half-warp

__device__ void chol_blocked_cta(double* L, int ld, int p, double* x, double* sm, int* ok_s) {
    const int tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5, lane = tid & 31;
    const int a8 = lane >> 2, q = lane & 3;
    double* D = sm;
    double* invd = D + kNB * kDPitch;
    double* P = invd + kNB;
    if (tid == 0) *ok_s = 1;
    __syncthreads();
    for (int kb = 0; kb < p; kb += kNB) {
        const int nb = (p - kb < kNB) ? (p - kb) : kNB;
        const int m = p - kb - nb;            // rows below the panel
        const int mpad = (m + 7) & ~7;
        // (1) diagonal block -> shared memory (identity beyond nb), factored by warp 0
        for (int e = tid; e < kNB * kNB; e += nt) {
            const int r = e >> 5, c = e & 31;
            double v = (r == c) ? 1.0 : 0.0;
            if (r < nb && c <= r) v = L[(size_t)(kb + r) * ld + kb + c];
            D[r * kDPitch + c] = v;
        }
        __syncthreads();
        if (warp == 0) {
            bool good = true;
            for (int j = 0; j < kNB; j++) {
                const double d = D[j * kDPitch + j];
                if (!(d > 0.0) || !isfinite(d)) { good = false; break; }  // warp-uniform
                const double dj = sqrt(d);
                const double lr = (lane > j) ? D[lane * kDPitch + j] / dj : 0.0;
                __syncwarp();
                if (lane == j) D[j * kDPitch + j] = dj;
                if (lane > j) D[lane * kDPitch + j] = lr;
                __syncwarp();
                if (lane > j)
                    for (int c = j + 1; c <= lane; c++)
                        D[lane * kDPitch + c] = fma(-lr, D[c * kDPitch + j], D[lane * kDPitch + c]);
                __syncwarp();
            }
            if (good) invd[lane] = 1.0 / D[lane * kDPitch + lane];
            else if (lane == 0) *ok_s = 0;
        }
        __syncthreads();
        if (*ok_s == 0) return;
        for (int e = tid; e < kNB * kNB; e += nt) {
            const int r = e >> 5, c = e & 31;
            if (r < nb && c <= r) L[(size_t)(kb + r) * ld + kb + c] = D[r * kDPitch + c];
        }
        if (m == 0) break;
        // (2) panel: row i solves  x L11' = a  (forward substitution over the 32 columns), x -> global and P
        for (int ri = tid; ri < mpad; ri += nt) {
            double xr[kNB];
            if (ri < m) {
                double* row = L + (size_t)(kb + nb + ri) * ld + kb;
#pragma unroll
                for (int c = 0; c < kNB; c++) xr[c] = (c < nb) ? row[c] : 0.0;
#pragma unroll
                for (int c = 0; c < kNB; c++) {
                    double sacc = xr[c];
#pragma unroll
                    for (int t = 0; t < c; t++) sacc = fma(-xr[t], D[c * kDPitch + t], sacc);
                    xr[c] = sacc * invd[c];
                }
#pragma unroll
                for (int c = 0; c < kNB; c++)
                    if (c < nb) row[c] = xr[c];
            } else {
#pragma unroll
                for (int c = 0; c < kNB; c++) xr[c] = 0.0;
            }
#pragma unroll
            for (int c = 0; c < kNB; c++) P[ri * kPPitch + c] = xr[c];
        }
        __syncthreads();
        // (3) trailing update on DMMA: tile (ti, tc), tc <= ti, rows kb+nb+8ti.., columns kb+nb+8tc..
        const int mt = mpad >> 3;
        const int ntile = mt * (mt + 1) / 2;
        const int base = kb + nb;
        for (int t0 = warp * kTIF; t0 < ntile; t0 += (nt >> 5) * kTIF) {
            double c0[kTIF], c1[kTIF];
            int gi[kTIF], gc[kTIF], pi[kTIF], pc[kTIF];
            bool in0[kTIF], in1[kTIF];
#pragma unroll
            for (int u = 0; u < kTIF; u++) {
                const int t = t0 + u;
                int ti = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
                while ((ti + 1) * (ti + 2) / 2 <= t) ti++;
                while (ti * (ti + 1) / 2 > t) ti--;
                const int tc = t - ti * (ti + 1) / 2;
                const bool live = t < ntile;
                pi[u] = (live ? 8 * ti : 0) + a8;
                pc[u] = (live ? 8 * tc : 0) + a8;
                gi[u] = base + (live ? 8 * ti : 0) + a8;
                gc[u] = base + (live ? 8 * tc : 0) + 2 * q;
                in0[u] = live && gi[u] < p && gc[u] <= gi[u];
                in1[u] = live && gi[u] < p && gc[u] + 1 <= gi[u];
                c0[u] = in0[u] ? L[(size_t)gi[u] * ld + gc[u]] : 0.0;
                c1[u] = in1[u] ? L[(size_t)gi[u] * ld + gc[u] + 1] : 0.0;
            }
#pragma unroll
            for (int kk = 0; kk < kNB; kk += 4) {
#pragma unroll
                for (int u = 0; u < kTIF; u++) {
                    const double av = -P[pi[u] * kPPitch + kk + q];
                    const double bv = P[pc[u] * kPPitch + kk + q];
                    dmma884(c0[u], c1[u], av, bv);
                }
            }
#pragma unroll
            for (int u = 0; u < kTIF; u++) {
                if (in0[u]) L[(size_t)gi[u] * ld + gc[u]] = c0[u];
                if (in1[u]) L[(size_t)gi[u] * ld + gc[u] + 1] = c1[u];
            }
        }
        __syncthreads();
    }
}

// x = (L L')^-1 x with the factor in global memory: blocked forward and backward substitution (diagonal block solved by
// warp 0 with the lane's row / column of the block in registers, the rest of x updated by all threads)
__device__ void chol_blocked_solve_cta(const double* L, int ld, int p, double* x) {
    const int tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5, lane = tid & 31;
    __syncthreads();  // the last diagonal block was written back by all threads
    for (int kb = 0; kb < p; kb += kNB) {  // forward: L z = b
        const int nb = (p - kb < kNB) ? (p - kb) : kNB;
        if (warp == 0) {
            double lrow[kNB];
#pragma unroll
            for (int j = 0; j < kNB; j++) lrow[j] = (lane < nb && j <= lane) ? L[(size_t)(kb + lane) * ld + kb + j] : 1.0;
            double xr = (lane < nb) ? x[kb + lane] : 0.0;
#pragma unroll
            for (int j = 0; j < kNB; j++) {
                const double djj = __shfl_sync(0xffffffffu, lrow[j], j);
                const double xj = __shfl_sync(0xffffffffu, xr, j) / djj;
                if (lane == j) xr = xj;
                if (lane > j) xr = fma(-lrow[j], xj, xr);
            }
            if (lane < nb) x[kb + lane] = xr;
        }
        __syncthreads();
        for (int i = kb + nb + tid; i < p; i += nt) {
            const double* row = L + (size_t)i * ld + kb;
            double sacc = 0.0;
            for (int t = 0; t < nb; t++) sacc = fma(row[t], x[kb + t], sacc);
            x[i] -= sacc;
        }
        __syncthreads();
    }
    const int nblk = (p + kNB - 1) / kNB;
    for (int bi = nblk - 1; bi >= 0; bi--) {  // backward: L' x = z
        const int kb = bi * kNB;
        const int nb = (p - kb < kNB) ? (p - kb) : kNB;
        if (warp == 0) {
            double lcol[kNB];  // column `lane` of the block: L[kb+j][kb+lane], j >= lane
#pragma unroll
            for (int j = 0; j < kNB; j++) lcol[j] = (j < nb && lane <= j) ? L[(size_t)(kb + j) * ld + kb + lane] : 1.0;
            double xr = (lane < nb) ? x[kb + lane] : 0.0;
#pragma unroll
            for (int j = kNB - 1; j >= 0; j--) {
                const double djj = __shfl_sync(0xffffffffu, lcol[j], j);
                const double xj = __shfl_sync(0xffffffffu, xr, j) / djj;
                if (lane == j) xr = xj;
                if (lane < j && j < nb) xr = fma(-lcol[j], xj, xr);
            }
            if (lane < nb) x[kb + lane] = xr;
        }
        __syncthreads();
        for (int c = tid; c < kb; c += nt) {
            double sacc = 0.0;
            for (int t = 0; t < nb; t++) sacc = fma(L[(size_t)(kb + t) * ld + c], x[kb + t], sacc);
            x[c] -= sacc;
        }
        __syncthreads();
    }
}

// mode 0: Newton update (beta += delta with damping / convergence bookkeeping); mode 1: plain solve
template <bool SMEM>
__global__ void __launch_bounds__(SMEM ? kCholThreads : kCholThreadsBlocked) chol_kernel(const double* __restrict__ A, const double* __restrict__ b,
                                                            int K, int p, int mode, double* beta, double* x_out,
                                                            double* max_delta, int* status, int* active, int* iters,
                                                            int* n_active, double tol, double damp_above,
                                                            double damp_factor, double* gws) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int ok_s;
    __shared__ double md_s[kCholThreadsBlocked / 32];
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
    if (SMEM) {
        chol_solve_cta(L, ld, p, xs, &ok_s);
    } else {
        chol_blocked_cta(L, ld, p, xs, xs + ((p + 1) & ~1), &ok_s);
        if (ok_s != 0) chol_blocked_solve_cta(L, ld, p, xs);
    }
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
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) md = fmax(md, md_s[w]);
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
        const int mmax = (p > kNB) ? ((p - kNB + 7) & ~7) : 8;
        const size_t smem = ((size_t)((p + 1) & ~1) + (size_t)kNB * kDPitch + kNB + (size_t)mmax * kPPitch) * sizeof(double);
        DLSA_CUDA(cudaFuncSetAttribute(chol_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        chol_kernel<false><<<K, kCholThreadsBlocked, smem, st>>>(A, b, K, p, mode, beta, x_out, max_delta, status, active,
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
        if (e < p * p) {  // 8 loads in flight, added in the order k = 0..K-1 (x + 0.0 == x: skipped partitions add 0)
            int k = 0;
            for (; k + 8 <= K; k += 8) {
                double t[8];
#pragma unroll
                for (int u = 0; u < 8; u++)
                    t[u] = (status == nullptr || status[k + u] == DLSA_ST_OK) ? H[(size_t)(k + u) * p * p + e] : 0.0;
#pragma unroll
                for (int u = 0; u < 8; u++) v += t[u];
            }
            for (; k < K; k++)
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
