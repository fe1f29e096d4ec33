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
__device__ __forceinline__ void chol_solve_cta(double* L, int ld, int p, double* x, int* ok_s) {
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


// ---- p > 160: the factor does not fit shared memory; blocked LEFT-looking Cholesky on the copy in global memory -------
// Panels of 32 columns.  For panel kb the 16 warps each take 32-row blocks of the panel and form
//     U = A[rows, kb:kb+32] - L[rows, 0:kb] L[kb:kb+32, 0:kb]'
// as 4x4 tiles of 8x8 on DMMA with the 16 accumulator tiles in registers for the whole k loop (fragments straight from
// global memory, 16 bytes per lane = two k-steps per load, the next load issued before the current MMAs), then leave U in a
// shared-memory panel P.  The 32x32 diagonal block of P is factored by one warp, every row below solves its 32-column
// slice against it (one thread per row, the row in registers) and the finished columns go back to global memory.
// Unlike the right-looking form this replaces (profiles/r01_k2_chol_p500_ncu_summary.csv: DMMA pipe 10 % busy, the warps
// waiting on the read-modify-write of the trailing matrix in global memory and on CTA barriers), no element of the
// trailing matrix is touched before its own panel: every entry of A is read once and every entry of L written once.
constexpr int kNB = 32;
constexpr int kDPitch = 33;
constexpr int kPPitch = 36;  // == 4 (mod 16) doubles: 16-byte aligned rows, (row a8, column pair q) stores spread over the banks

// The right-hand side rides along as ROW p of the matrix: its "factor row" L[p][0:p] is z = L^-1 b, so the forward
// substitution (16 serial block steps with every other warp waiting, ~10 % of the kernel in the r02 ncu capture) is part
// of the panel GEMMs.  Rows of the panel therefore run to p inclusive; L has p + 1 rows.
__device__ __forceinline__ void chol_blocked_cta(const double* __restrict__ A, double* L, int ld, int p, double* x,
                                                 double* sm, int* ok_s) {
    const int tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5, lane = tid & 31, nwarps = nt >> 5;
    const int a8 = lane >> 2, q = lane & 3;
    double* D = sm;
    double* invd = D + kNB * kDPitch;
    double* W = invd + kNB;                 // inverse of the diagonal block's factor, [32][kPPitch]
    double* P = W + kNB * kPPitch;
    if (tid == 0) *ok_s = 1;
    __syncthreads();
    const int pcap = (p + 32) >> 5;   // 32-row blocks P has room for (rows 0..p)
    for (int kb = 0; kb < p; kb += kNB) {
        const int nb = (p - kb < kNB) ? (p - kb) : kNB;
        const int rows_total = p + 1 - kb;             // rows of the panel, the diagonal block first, the rhs row last
        const int nrb = (rows_total + 31) >> 5;
        // (1) panel update on DMMA: 32 rows x 32 columns per warp and turn, k over the finished columns 0..kb-1.  Late panels
        // have few row blocks and long k loops: the k range is then SPLIT over the idle warps (nsp parts per block), the
        // partial products parked in the unused rows of P and summed in a fixed order afterwards.
        const int kiters = kb >> 3;
        int nsp = nwarps / nrb;
        if (nsp > pcap / nrb) nsp = pcap / nrb;
        if (nsp > (kiters >> 1)) nsp = kiters >> 1;
        if (nsp < 1) nsp = 1;
        for (int wi = warp; wi < nrb * nsp; wi += nwarps) {
            const int b = wi % nrb, sp = wi / nrb;
            const int r0 = kb + 32 * b;
            const int k_lo = ((kiters * sp) / nsp) << 3, k_hi = ((kiters * (sp + 1)) / nsp) << 3;
            double c0[4][4], c1[4][4];
#pragma unroll
            for (int ti = 0; ti < 4; ti++)
#pragma unroll
                for (int tj = 0; tj < 4; tj++) { c0[ti][tj] = 0.0; c1[ti][tj] = 0.0; }
            int ao[4], bo[4];
            bool aok[4], bok[4];
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const int ra = r0 + 8 * t + a8, rb = kb + 8 * t + a8;
                aok[t] = ra <= p;
                bok[t] = rb < p;
                ao[t] = (aok[t] ? ra : p - 1) * ld + 2 * q;
                bo[t] = (bok[t] ? rb : p - 1) * ld + 2 * q;
            }
#pragma unroll 1
            for (int kk = k_lo; kk < k_hi; kk += 8) {
                double2 ac[4], bc[4];
                // (prefetch.global.L1 of the next one or two k blocks: p = 500 -2 %, p = 300 +3 % — not kept)
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    ac[t] = *reinterpret_cast<const double2*>(L + ao[t] + kk);
                    bc[t] = *reinterpret_cast<const double2*>(L + bo[t] + kk);
                    if (!aok[t]) ac[t] = make_double2(0.0, 0.0);
                    if (!bok[t]) bc[t] = make_double2(0.0, 0.0);
                }
                // lane q carries columns kk + 2q (.x) and kk + 2q + 1 (.y) of BOTH operands: two k-steps whose k order is
                // a permutation of kk..kk+7 — the sum over k does not care
#pragma unroll
                for (int ti = 0; ti < 4; ti++)
#pragma unroll
                    for (int tj = 0; tj < 4; tj++) {
                        dmma884(c0[ti][tj], c1[ti][tj], ac[ti].x, bc[tj].x);
                        dmma884(c0[ti][tj], c1[ti][tj], ac[ti].y, bc[tj].y);
                    }
            }
            // U = A - (L L')  ->  P; entries above the diagonal of the diagonal block are never read.  With a split k range
            // the partial product goes to block sp * nrb + b of P instead and U is formed below.
#pragma unroll
            for (int ti = 0; ti < 4; ti++) {
                const int r = r0 + 8 * ti + a8;
#pragma unroll
                for (int tj = 0; tj < 4; tj++) {
                    const int c = kb + 8 * tj + 2 * q;
                    double v0 = c0[ti][tj], v1 = c1[ti][tj];
                    if (nsp == 1) {
                        v0 = 0.0, v1 = 0.0;
                        if (r <= p) {   // the entries of A (row p: of b) are read here, once; L only ever holds finished columns
                            const double* Ar = r < p ? A + (size_t)r * p : x;
                            if (c < kb + nb && c <= r) v0 = Ar[c] - c0[ti][tj];
                            if (c + 1 < kb + nb && c + 1 <= r) v1 = Ar[c + 1] - c1[ti][tj];
                        }
                    }
                    *reinterpret_cast<double2*>(P + (size_t)(32 * (sp * nrb + b) + 8 * ti + a8) * kPPitch + 8 * tj + 2 * q) =
                        make_double2(v0, v1);
                }
            }
        }
        __syncthreads();
        if (nsp > 1) {
            for (int e = tid; e < nrb * 32 * 16; e += nt) {   // (row of the panel, column pair)
                const int rr = e >> 4, cp = (e & 15) << 1;
                const int r = kb + rr, c = kb + cp;
                double s0 = 0.0, s1 = 0.0;
                for (int sp = 0; sp < nsp; sp++) {
                    const double2 v = *reinterpret_cast<const double2*>(P + (size_t)(32 * sp * nrb + rr) * kPPitch + cp);
                    s0 += v.x;
                    s1 += v.y;
                }
                double v0 = 0.0, v1 = 0.0;
                if (r <= p) {
                    const double* Ar = r < p ? A + (size_t)r * p : x;
                    if (c < kb + nb && c <= r) v0 = Ar[c] - s0;
                    if (c + 1 < kb + nb && c + 1 <= r) v1 = Ar[c + 1] - s1;
                }
                *reinterpret_cast<double2*>(P + (size_t)rr * kPPitch + cp) = make_double2(v0, v1);
            }
            __syncthreads();
        }
        // (2) diagonal block (rows 0..31 of P) -> D (identity beyond nb), factored by warp 0
        for (int e = tid; e < kNB * kNB; e += nt) {
            const int r = e >> 5, c = e & 31;
            double v = (r == c) ? 1.0 : 0.0;
            if (r < nb && c <= r) v = P[r * kPPitch + c];
            D[r * kDPitch + c] = v;
        }
        __syncthreads();
        if (warp == 0) {
            // lane i holds row i of the block in registers; column j of the factor travels by shuffle
            double row[kNB];
#pragma unroll
            for (int c = 0; c < kNB; c++) row[c] = D[lane * kDPitch + c];
            bool good = true;
#pragma unroll
            for (int j = 0; j < kNB; j++) {
                const double d = __shfl_sync(0xffffffffu, row[j], j);
                good = good && (d > 0.0) && isfinite(d);
                const double rs = rsqrt(d);                          // one reciprocal square root instead of sqrt + divide on the
                const double lj = row[j] * rs;                       // 32-step dependent chain: column j of the factor (lane j: d/sqrt(d))
                row[j] = lj;
                // column j goes round through shared memory (W is free until the inverse is formed; two buffers, so a
                // fast lane's next column cannot overwrite what a slow lane still reads): one broadcast load per element
                // instead of two shuffles
                double* colbuf = W + (j & 1) * kNB;
                colbuf[lane] = lj;
                __syncwarp();
#pragma unroll
                for (int c = j + 1; c < kNB; c++) row[c] = fma(-lj, colbuf[c], row[c]);   // only c <= lane is ever read
            }
            if (good) {
#pragma unroll
                for (int c = 0; c < kNB; c++) D[lane * kDPitch + c] = row[c];
                // M = L11^-1 (lower triangular), lane j owns column j in registers:
                //   M[i][j] = -(sum_{t=j}^{i-1} L11[i][t] M[t][j]) / L11[i][i];  L11[i][t] is read back from D (one broadcast
                //   load instead of two shuffles, and the lane's row of the factor need not stay in registers next to m[])
                double m[kNB];
                const double dinv = 1.0 / row[lane];
                invd[lane] = dinv;
                __syncwarp();
#pragma unroll
                for (int i = 0; i < kNB; i++) {
                    double s0 = 0.0, s1 = 0.0;
#pragma unroll
                    for (int t = 0; t < i; t++) {
                        const double lit = D[i * kDPitch + t];
                        if (t & 1) s1 = fma(lit, m[t], s1); else s0 = fma(lit, m[t], s0);
                    }
                    const double di = invd[i];
                    m[i] = (lane == i) ? dinv : ((lane < i) ? -(s0 + s1) * di : 0.0);
                    W[i * kPPitch + lane] = m[i];
                }
            } else if (lane == 0) {
                *ok_s = 0;
            }
        }
        __syncthreads();
        if (*ok_s == 0) return;
        for (int e = tid; e < kNB * kNB; e += nt) {
            const int r = e >> 5, c = e & 31;
            if (r < nb && c <= r) L[(size_t)(kb + r) * ld + kb + c] = D[r * kDPitch + c];
        }
        // (3) panel rows below the diagonal block:  X = U L11^-T = U M'  — one more 32-wide GEMM on DMMA per 32-row block
        //     (A fragments from P, B[k][n] = M[n][k] from W), straight to global memory
        //     (the last, ragged panel keeps the rhs row inside block 0, right under the nb rows of the diagonal block)
        for (int b = (nb < kNB ? 0 : 1) + warp; b < nrb; b += nwarps) {
            double c0[4][4], c1[4][4];
#pragma unroll
            for (int ti = 0; ti < 4; ti++)
#pragma unroll
                for (int tj = 0; tj < 4; tj++) { c0[ti][tj] = 0.0; c1[ti][tj] = 0.0; }
            const double* Pa = P + (size_t)(32 * b + a8) * kPPitch + q;
            const double* Wb = W + (size_t)a8 * kPPitch + q;
#pragma unroll 2
            for (int kk = 0; kk < kNB; kk += 4) {
                double av[4], bv[4];
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    av[t] = Pa[(size_t)(8 * t) * kPPitch + kk];
                    bv[t] = Wb[(size_t)(8 * t) * kPPitch + kk];
                }
#pragma unroll
                for (int ti = 0; ti < 4; ti++)
#pragma unroll
                    for (int tj = 0; tj < 4; tj++) dmma884(c0[ti][tj], c1[ti][tj], av[ti], bv[tj]);
            }
#pragma unroll
            for (int ti = 0; ti < 4; ti++) {
                const int r = kb + 32 * b + 8 * ti + a8;
                if (r <= p && r >= kb + nb) {
#pragma unroll
                    for (int tj = 0; tj < 4; tj++) {
                        const int c = 8 * tj + 2 * q;   // a ragged panel has nb < 32 columns: nothing past them (row p is
                        double* dst = L + (size_t)r * ld + kb + c;   // the LAST row of this matrix's workspace)
                        if (c + 1 < nb) *reinterpret_cast<double2*>(dst) = make_double2(c0[ti][tj], c1[ti][tj]);
                        else if (c < nb) *dst = c0[ti][tj];
                    }
                }
            }
        }
        __syncthreads();   // the finished columns are in global memory before the next panel reads them
    }
}

// x = (L L')^-1 b with the factor in global memory: z = L^-1 b is row p of the factor (see above); blocked backward
// substitution (diagonal block solved by warp 0 with the lane's column of the block in registers, the rest of x updated by
// all threads)
__device__ __forceinline__ void chol_blocked_solve_cta(const double* L, int ld, int p, double* x) {
    const int tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5, lane = tid & 31;
    __syncthreads();  // the last diagonal block was written back by all threads
    // forward substitution: done by the factorisation (row p of L is z = L^-1 b)
    for (int i = tid; i < p; i += nt) x[i] = L[(size_t)p * ld + i];
    __syncthreads();
    const int nblk = (p + kNB - 1) / kNB;
    for (int bi = nblk - 1; bi >= 0; bi--) {  // backward: L' x = z
        const int kb = bi * kNB;
        const int nb = (p - kb < kNB) ? (p - kb) : kNB;
        if (warp == 0) {
            double lcol[kNB];  // column `lane` of the block: L[kb+j][kb+lane], j >= lane
#pragma unroll
            for (int j = 0; j < kNB; j++) lcol[j] = (j < nb && lane <= j) ? L[(size_t)(kb + j) * ld + kb + lane] : 1.0;
            double xr = (lane < nb) ? x[kb + lane] : 0.0;
            double rinv = 1.0;
#pragma unroll
            for (int j = 0; j < kNB; j++) rinv = (lane == j) ? 1.0 / lcol[j] : rinv;
#pragma unroll
            for (int j = kNB - 1; j >= 0; j--) {
                const double xj = __shfl_sync(0xffffffffu, xr * rinv, j);
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
        ld = (p + 1) & ~1;  // even: every row of the factor starts 16-byte aligned (two doubles per fragment load)
        L = gws + (size_t)k * (p + 1) * ld;
    }
    const double* Ak = A + (size_t)k * p * p;
    if (SMEM) {
        for (int e = threadIdx.x; e < p * p; e += blockDim.x) {
            const int i = e / p, c = e - i * p;
            if (c <= i) L[(size_t)i * ld + c] = Ak[e];
        }
    }
    for (int i = threadIdx.x; i < p; i += blockDim.x) xs[i] = b[(size_t)k * p + i];
    __syncthreads();
    if (SMEM) {
        chol_solve_cta(L, ld, p, xs, &ok_s);
    } else {
        chol_blocked_cta(Ak, L, ld, p, xs, xs + ((p + 1) & ~1), &ok_s);
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

// ---- p <= 64: ONE WARP per matrix ------------------------------------------------------------------------------------
// The CTA version above spends its time in block-wide barriers (3 per column, 2 per substitution step: ~110 us for the
// 1024 50x50 systems of C2).  Here a warp owns one system in its own slice of shared
// memory (odd row pitch: a lane per row walks distinct banks) and only ever needs __syncwarp:
//   * left-looking factorisation, column j: lane r takes rows j + r, j + r + 32 (and j + r + 64): the dot products over the
//     finished columns read the lane's own row and broadcast row j;
//   * the right-hand side is row p of the matrix, so its factor row is z = L^-1 b (as in the blocked kernel);
//   * back substitution with z in registers (lane owns entries lane and lane + 32), x_j broadcast by shuffle, and the
//     reciprocal diagonal kept from the factorisation: no division on the chain.
constexpr int kCholWarpMaxP = 64;

static int chol_warp_per_cta(int p) {
    const size_t per_warp = ((size_t)(p + 1) * (p | 1) + kCholWarpMaxP) * sizeof(double);
    int w = (int)((size_t)(200 * 1024) / per_warp);
    return w > 8 ? 8 : (w < 1 ? 1 : w);
}

__global__ void __launch_bounds__(256) chol_warp_kernel(const double* __restrict__ A, const double* __restrict__ b, int K, int p,
                                                        int mode, double* beta, double* x_out, double* max_delta, int* status,
                                                        int* active, int* iters, int* n_active, double tol, double damp_above,
                                                        double damp_factor) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k = blockIdx.x * (blockDim.x >> 5) + warp;
    if (k >= K) return;   // no block-wide barrier below
    if (mode == 0 && active != nullptr && active[k] == 0) return;
    const int ld = p | 1;
    double* L = reinterpret_cast<double*>(smem_raw) + (size_t)warp * ((size_t)(p + 1) * ld + kCholWarpMaxP);
    double* dinv = L + (size_t)(p + 1) * ld;
    const double* Ak = A + (size_t)k * p * p;
    // the whole p x p block in batches of 16 coalesced loads per lane, all in flight before the first store (a row-by-row
    // copy of the lower triangle serialises p global-memory latencies: 83 us per call instead of 2x this kernel's arithmetic)
    const int n2 = p * p;
    for (int e0 = 0; e0 < n2; e0 += 32 * 16) {
        double v[16];
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const int e = e0 + lane + 32 * q;
            v[q] = (e < n2) ? __ldg(Ak + e) : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const int e = e0 + lane + 32 * q;
            if (e < n2) {
                const int i = e / p, c = e - i * p;
                if (c <= i) L[i * ld + c] = v[q];
            }
        }
    }
    for (int c = lane; c < p; c += 32) L[p * ld + c] = b[(size_t)k * p + c];
    __syncwarp();
    bool good = true;
    for (int j = 0; j < p; j++) {
        const double* Lj = L + j * ld;
        const int i0 = j + lane, i1 = i0 + 32, i2 = i0 + 64;
        const bool ok0 = i0 <= p, ok1 = i1 <= p, ok2 = i2 <= p;
        const double* R0 = L + (ok0 ? i0 : j) * ld;
        const double* R1 = L + (ok1 ? i1 : j) * ld;
        const double* R2 = L + (ok2 ? i2 : j) * ld;
        double a0 = R0[j], a1 = R1[j], a2 = R2[j];
        if (j + 64 <= p) {   // warp-uniform: a third row per lane only while more than 64 rows are left (p = 64, j = 0)
            for (int t = 0; t < j; t++) {
                const double ljt = Lj[t];
                a0 = fma(-R0[t], ljt, a0);
                a1 = fma(-R1[t], ljt, a1);
                a2 = fma(-R2[t], ljt, a2);
            }
        } else {
            for (int t = 0; t < j; t++) {   // (four columns per trip with split partial sums: no faster, 0.066 ms either way)
                const double ljt = Lj[t];
                a0 = fma(-R0[t], ljt, a0);
                a1 = fma(-R1[t], ljt, a1);
            }
        }
        const double d = __shfl_sync(0xffffffffu, a0, 0);   // lane 0 holds row j: the pivot
        good = good && (d > 0.0) && isfinite(d);
        const double rs = rsqrt(d);
        __syncwarp();   // every lane has read row j before lane 0 rewrites its entry (j, j)
        if (ok0) L[i0 * ld + j] = a0 * rs;
        if (ok1) L[i1 * ld + j] = a1 * rs;
        if (ok2) L[i2 * ld + j] = a2 * rs;
        if (lane == 0) dinv[j] = rs;
        __syncwarp();
    }
    // z = row p of the factor; x = L^-T z, entries lane and lane + 32 in registers
    double z0 = (lane < p) ? L[p * ld + lane] : 0.0;
    double z1 = (lane + 32 < p) ? L[p * ld + lane + 32] : 0.0;
    for (int j = p - 1; j >= 0; j--) {
        const double zj = __shfl_sync(0xffffffffu, (j < 32) ? z0 : z1, j & 31);
        const double xj = zj * dinv[j];
        if (lane == (j & 31)) { if (j < 32) z0 = xj; else z1 = xj; }
        const double* Lj = L + j * ld;
        if (lane < j) z0 = fma(-Lj[lane], xj, z0);
        if (lane + 32 < j) z1 = fma(-Lj[lane + 32], xj, z1);
    }
    const bool ok = good;
    if (mode == 1) {
        if (lane < p) x_out[(size_t)k * p + lane] = ok ? z0 : 0.0;
        if (lane + 32 < p) x_out[(size_t)k * p + lane + 32] = ok ? z1 : 0.0;
        if (lane == 0 && status != nullptr) status[k] = ok ? DLSA_ST_OK : DLSA_ST_NOT_SPD;
        return;
    }
    // Newton bookkeeping (same rules as chol_kernel)
    double md = fmax((lane < p) ? fabs(z0) : 0.0, (lane + 32 < p) ? fabs(z1) : 0.0);
    bool finite = isfinite(md);
    if (!finite) md = INFINITY;
    for (int m = 16; m >= 1; m >>= 1) md = fmax(md, __shfl_xor_sync(0xffffffffu, md, m));
    if (!ok || !isfinite(md)) {
        if (lane == 0) {
            status[k] = DLSA_ST_NOT_SPD;
            active[k] = 0;
            if (max_delta) max_delta[k] = INFINITY;
        }
        return;
    }
    const double scale = (md > damp_above) ? damp_factor : 1.0;  // dlsa/R/wlse_func.R:56-60
    if (md > tol) {   // a sub-tolerance step is not applied (see chol_kernel)
        if (lane < p) beta[(size_t)k * p + lane] += scale * z0;
        if (lane + 32 < p) beta[(size_t)k * p + lane + 32] += scale * z1;
    }
    if (lane == 0) {
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
    return p <= kCholSmemMaxP ? 0 : (size_t)K * (p + 1) * ((p + 1) & ~1) * sizeof(double);   // p + 1 rows: the rhs row
}

int chol_launch(const double* A, const double* b, int K, int p, int mode, double* beta, double* x_out,
                double* max_delta, int* status, int* active, int* iters, int* n_active, double tol, double damp_above,
                double damp_factor, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    if (K == 0) return 0;
    if (p <= kCholWarpMaxP) {
        const int wpc = chol_warp_per_cta(p);
        const size_t smem = (size_t)wpc * ((size_t)(p + 1) * (p | 1) + kCholWarpMaxP) * sizeof(double);
        DLSA_CUDA(cudaFuncSetAttribute(chol_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        chol_warp_kernel<<<(K + wpc - 1) / wpc, wpc * 32, smem, st>>>(A, b, K, p, mode, beta, x_out, max_delta, status, active,
                                                                      iters, n_active, tol, damp_above, damp_factor);
    } else if (p <= kCholSmemMaxP) {
        const size_t smem = chol_smem_bytes(p);
        DLSA_CUDA(cudaFuncSetAttribute(chol_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        chol_kernel<true><<<K, kCholThreads, smem, st>>>(A, b, K, p, mode, beta, x_out, max_delta, status, active, iters,
                                                         n_active, tol, damp_above, damp_factor, nullptr);
    } else {
        const size_t need = chol_workspace_bytes(K, p);
        DLSA_REQUIRE(workspace != nullptr && workspace_bytes >= need, DLSA_E_WORKSPACE,
                     "cholesky: workspace too small (%zu < %zu)", workspace_bytes, need);
        const int prow = (p + 32) & ~31;   // the panel holds every row from the diagonal block down to the rhs row, in 32-row blocks
        const size_t smem = ((size_t)((p + 1) & ~1) + (size_t)kNB * kDPitch + kNB + (size_t)kNB * kPPitch +
                             (size_t)prow * kPPitch) * sizeof(double);
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
