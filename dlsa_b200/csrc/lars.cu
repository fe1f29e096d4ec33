// K6 + K7: LARS / lasso path on the least-squares approximation, then RSS / dof / AIC / BIC and argmin.
//
// Reference replaced: lars_lsa (dlsa/lsa.py:90-212; updateR :12-32, backsolvet :8-9, delcol :35-71,
// downdateR :74-80); intercept branch as in the R original (dlsa/R/dlsa_alasso_func.R:115-125, 243-245);
// first-argmin selection of dlsa() (dlsa/dlsa.py:90-91).
//
// The path is sequential (each knot depends on the previous one), so K6 is ONE persistent CTA: the O(m a)
// correlation update runs on all threads against the L2-resident scaled Sigma.  Until a lasso drop happens nothing in
// the step is a sequential triangular solve: besides the upper Cholesky factor R the kernel keeps R^-1 (packed like R),
// z = R^-T Sign and G = R^-1 z, and appending a variable is two triangular MATRIX-VECTOR products on all threads
// (v = R^-T x_old, the new column of R; c = -R^-1 v / r_pp, the new column of R^-1) plus O(a) updates of z and G —
// the reference's updateR / backsolve chain (dlsa/lsa.py:12-32,151) computes the same vectors by substitution.
// After the first drop (Givens deletion on R, dlsa/lsa.py:35-71) the step falls back to blocked triangular solves on R
// (diagonal blocks on warp 0 with shuffles, the rest on all threads).
// K7 evaluates the quadratic form of every knot in parallel (one CTA per knot) and K7b picks the minima.
#include "common.cuh"

namespace dlsa {

constexpr int kLarsThreads = 512;
constexpr int kLarsMaxM = 1024;          // vectors live in shared memory
constexpr int kLarsRSmemDoubles = 16384;  // packed R kept in shared memory up to m = 180 (128 KB)

struct LarsShared {
    double Cmax, A, gamhat, zmin, red[kLarsThreads / 32];
    int n_act, n_join, k, drops_pending, any_drop, err, fast;
    int ired[kLarsThreads / 32];
};

__device__ __forceinline__ size_t rpk(int l, int i, int m) {  // packed upper-triangular rows, i >= l
    return (size_t)l * m - (size_t)l * (l - 1) / 2 + (i - l);
}

__device__ double block_reduce(double v, bool is_max, double* red) {
    for (int s = 16; s >= 1; s >>= 1) {
        const double o = __shfl_xor_sync(0xffffffffu, v, s);
        v = is_max ? fmax(v, o) : fmin(v, o);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = red[0];
    for (int w = 1; w < kLarsThreads / 32; w++) r = is_max ? fmax(r, red[w]) : fmin(r, red[w]);
    __syncthreads();
    return r;
}

__device__ double block_sum(double v, double* red) {
    for (int s = 16; s >= 1; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    for (int w = 0; w < kLarsThreads / 32; w++) r += red[w];
    __syncthreads();
    return r;
}

// warp 0: solve R' r = x in place (R upper, packed rows).  Column-oriented: row l of R is contiguous.
__device__ void forward_solve_w0(const double* R, int a, int m, double* x) {
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        for (int l = 0; l < a; l++) {
            const double rl = x[l] / R[rpk(l, l, m)];
            __syncwarp();
            if (lane == 0) x[l] = rl;
            for (int i = l + 1 + lane; i < a; i += 32) x[i] = fma(-R[rpk(l, i, m)], rl, x[i]);
            __syncwarp();
        }
    }
    __syncthreads();
}
// warp 0: solve R z = x in place (row-oriented dot products)
__device__ void backward_solve_w0(const double* R, int a, int m, double* x) {
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        for (int l = a - 1; l >= 0; l--) {
            double d = 0.0;
            for (int i = l + 1 + lane; i < a; i += 32) d = fma(R[rpk(l, i, m)], x[i], d);
            for (int s = 16; s >= 1; s >>= 1) d += __shfl_xor_sync(0xffffffffu, d, s);
            __syncwarp();
            if (lane == 0) x[l] = (x[l] - d) / R[rpk(l, l, m)];
            __syncwarp();
        }
    }
    __syncthreads();
}

// ---- blocked versions for a large active set (R in global memory, a > kBlkMin): 32-column blocks; the diagonal block is
// solved by warp 0 with the lane's column / row of the block in registers, the rest of x is updated by ALL threads with
// 32 independent loads each.  The warp-0 versions above walk the packed factor one dependent global load per column:
// ~200 us per solve at a = 500, three solves per LARS step (C5: 415 ms for the 501-knot path).
constexpr int kBlk = 32;
constexpr int kBlkMin = 64;

// solve R' r = x in place
__device__ void forward_solve_blk(const double* R, int a, int m, double* x) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    for (int kb = 0; kb < a; kb += kBlk) {
        const int nb = (a - kb < kBlk) ? (a - kb) : kBlk;
        if (tid < 32) {
            double col[kBlk];  // column `lane` of the diagonal block: R[kb+t][kb+lane], t <= lane
#pragma unroll
            for (int t = 0; t < kBlk; t++) col[t] = (lane < nb && t <= lane) ? R[rpk(kb + t, kb + lane, m)] : 1.0;
            double xr = (lane < nb) ? x[kb + lane] : 0.0;
            double rd = 1.0;  // 1 / R[kb+lane][kb+lane]: one division per lane instead of one per column on the chain
#pragma unroll
            for (int t = 0; t < kBlk; t++)
                if (lane == t) rd = 1.0 / col[t];
#pragma unroll
            for (int t = 0; t < kBlk; t++) {
                const double xt = __shfl_sync(0xffffffffu, xr * rd, t);
                if (lane == t) xr = xt;
                if (lane > t) xr = fma(-col[t], xt, xr);
            }
            if (lane < nb) x[kb + lane] = xr;
        }
        __syncthreads();
        for (int i = kb + nb + tid; i < a; i += nt) {
            double sacc = 0.0;
#pragma unroll 8
            for (int t = 0; t < nb; t++) sacc = fma(R[rpk(kb + t, i, m)], x[kb + t], sacc);
            x[i] -= sacc;
        }
        __syncthreads();
    }
}

// solve R z = x in place
__device__ void backward_solve_blk(const double* R, int a, int m, double* x) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    const int nblk = (a + kBlk - 1) / kBlk;
    for (int bi = nblk - 1; bi >= 0; bi--) {
        const int kb = bi * kBlk;
        const int nb = (a - kb < kBlk) ? (a - kb) : kBlk;
        if (tid < 32) {
            double row[kBlk];  // row `lane` of the diagonal block: R[kb+lane][kb+s], s >= lane
#pragma unroll
            for (int sidx = 0; sidx < kBlk; sidx++)
                row[sidx] = (sidx < nb && lane <= sidx) ? R[rpk(kb + lane, kb + sidx, m)] : 1.0;
            double xr = (lane < nb) ? x[kb + lane] : 0.0;
            double rd = 1.0;
#pragma unroll
            for (int sidx = 0; sidx < kBlk; sidx++)
                if (lane == sidx) rd = 1.0 / row[sidx];
#pragma unroll
            for (int sidx = kBlk - 1; sidx >= 0; sidx--) {
                const double zs = __shfl_sync(0xffffffffu, xr * rd, sidx);
                if (lane == sidx) xr = zs;
                if (lane < sidx && sidx < nb) xr = fma(-row[sidx], zs, xr);
            }
            if (lane < nb) x[kb + lane] = xr;
        }
        __syncthreads();
        for (int l = tid; l < kb; l += nt) {
            const double* rl = R + rpk(l, kb, m);
            double sacc = 0.0;
#pragma unroll 8
            for (int t = 0; t < nb; t++) sacc = fma(rl[t], x[kb + t], sacc);
            x[l] -= sacc;
        }
        __syncthreads();
    }
}

__device__ __forceinline__ void forward_solve(const double* R, int a, int m, double* x) {
    if (a > kBlkMin) forward_solve_blk(R, a, m, x);
    else forward_solve_w0(R, a, m, x);
}
__device__ __forceinline__ void backward_solve(const double* R, int a, int m, double* x) {
    if (a > kBlkMin) backward_solve_blk(R, a, m, x);
    else backward_solve_w0(R, a, m, x);
}

// warp 0: delete column kdel of R (a x a) and restore triangularity with Givens rotations (dlsa/lsa.py:35-71)
__device__ void delete_column_w0(double* R, int a, int m, int kdel) {
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        // rows l <= kdel: shift columns kdel+1.. left by one
        for (int l = 0; l <= kdel; l++) {
            for (int j0 = kdel; j0 < a - 1; j0 += 32) {
                const int j = j0 + lane;
                double v = 0.0;
                if (j < a - 1) v = R[rpk(l, j + 1, m)];
                __syncwarp();
                if (j < a - 1 && j >= l) R[rpk(l, j, m)] = v;
                __syncwarp();
            }
        }
        for (int i = kdel + 1; i < a; i++) {
            const double bb = R[rpk(i, i, m)];  // becomes the sub-diagonal entry (i, i-1) after the shift
            __syncwarp();
            for (int j0 = i; j0 < a - 1; j0 += 32) {  // shift row i left by one
                const int j = j0 + lane;
                double v = 0.0;
                if (j < a - 1) v = R[rpk(i, j + 1, m)];
                __syncwarp();
                if (j < a - 1) R[rpk(i, j, m)] = v;
                __syncwarp();
            }
            const double aa = R[rpk(i - 1, i - 1, m)];
            double c = 1.0, s = 0.0;
            if (bb != 0.0) {
                if (!(fabs(bb) > fabs(aa))) {
                    const double tau = -bb / aa;
                    c = 1.0 / sqrt(1.0 + tau * tau);
                    s = c * tau;
                } else {
                    const double tau = -aa / bb;
                    s = 1.0 / sqrt(1.0 + tau * tau);
                    c = s * tau;
                }
                __syncwarp();
                if (lane == 0) R[rpk(i - 1, i - 1, m)] = c * aa - s * bb;
                for (int j = i + lane; j < a - 1; j += 32) {
                    const double x1 = R[rpk(i - 1, j, m)], x2 = R[rpk(i, j, m)];
                    R[rpk(i - 1, j, m)] = c * x1 - s * x2;
                    R[rpk(i, j, m)] = s * x1 + c * x2;
                }
            }
            __syncwarp();
        }
    }
    __syncthreads();
}

struct LarsArgs {
    const double* Sigma0;
    const double* b0;
    int p0, intercept, type, max_steps;
    double n, eps;
    double* beta_path;   // [(max_steps+1)][m]  (unscaled during the path; scaled by |b| in K7)
    double* beta0_path;  // [(max_steps+1)]
    double *rss, *aic, *bic;
    int* dof;
    int* info;
    double* Sig;   // workspace [m*m] scaled Sigma
    double* Rg;    // workspace packed R when it does not fit in shared memory
    double* Rinv;  // workspace packed R^-1 (same layout)
    double* scal;  // workspace [4]: a11, beta0_init, (unused)
};

__global__ void __launch_bounds__(kLarsThreads, 1) lars_path_kernel(const LarsArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ LarsShared sh;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int p0 = a.p0, off = a.intercept ? 1 : 0, m = p0 - off;
    const double eps = a.eps;

    double* Cvec = reinterpret_cast<double*>(smem_raw);
    double* u = Cvec + m;       // Sig[:, act] w
    double* w = u + m;          // direction over active (ordered like act)
    double* Sign = w + m;
    double* tmp = Sign + m;     // triangular-solve scratch
    double* sgn = tmp + m;      // sign(b)
    double* absb = sgn + m;
    double* zf = absb + m;      // R^-T Sign, kept up to date as columns are appended (one O(a) update per join instead of
                                // an O(a^2) forward solve per step); rebuilt after a lasso drop
    double* gi = zf + m;        // G = R^-1 z (the un-normalised equiangular direction), maintained while sh.fast
    int* act = reinterpret_cast<int*>(gi + m);
    int* state = act + m;       // 0 inactive, 1 active, 2 ignored
    int* joiners = state + m;
    unsigned char* dropflag = reinterpret_cast<unsigned char*>(joiners + m);
    double* Rs = reinterpret_cast<double*>(smem_raw + ((size_t)(9 * m) * 8 + (size_t)(3 * m) * 4 + m + 15) / 16 * 16);
    const bool r_in_smem = (size_t)m * (m + 1) / 2 <= kLarsRSmemDoubles;
    double* R = r_in_smem ? Rs : a.Rg;

    // ---- set-up: intercept profiling, adaptive rescaling (dlsa/lsa.py:98-109; R :115-131) ----------
    const double a11 = a.intercept ? a.Sigma0[0] : 1.0;
    for (int i = tid; i < m; i += nt) {
        const double bi = a.b0[i + off];
        absb[i] = fabs(bi);
        sgn[i] = (bi > 0.0) ? 1.0 : ((bi < 0.0) ? -1.0 : 0.0);
        state[i] = 0;
    }
    __syncthreads();
    for (int e = tid; e < m * m; e += nt) {
        const int i = e / m, j = e - i * m;
        double v = a.Sigma0[(size_t)(i + off) * p0 + (j + off)];
        if (a.intercept) v -= a.Sigma0[(size_t)(i + 1) * p0] * a.Sigma0[(size_t)(j + 1) * p0] / a11;
        a.Sig[e] = absb[i] * v * absb[j];
    }
    if (a.intercept) {
        double part = 0.0;
        for (int i = tid; i < m; i += nt) part = fma(a.Sigma0[(size_t)(i + 1) * p0], a.b0[i + 1], part);
        const double tot = block_sum(part, sh.red);
        if (tid == 0) { a.scal[0] = a11; a.scal[1] = tot / a11; }
    } else if (tid == 0) {
        a.scal[0] = 1.0; a.scal[1] = 0.0;
    }
    __syncthreads();
    for (int j = tid; j < m; j += nt) {
        double c = 0.0;
        for (int i = 0; i < m; i++) c = fma(sgn[i], a.Sig[(size_t)i * m + j], c);
        Cvec[j] = c;
        a.beta_path[j] = 0.0;
    }
    if (tid == 0) { sh.n_act = 0; sh.k = 0; sh.drops_pending = 0; sh.err = 0; sh.fast = 1; }
    __syncthreads();

    const int max_steps = a.max_steps;
    while (true) {
        if (!(sh.k < max_steps && sh.n_act < m)) break;
        __syncthreads();
        if (tid == 0) sh.k += 1;
        __syncthreads();
        const int k = sh.k;
        // Cmax over the inactive set (dlsa/lsa.py:128-129)
        double cm = 0.0;
        bool any_inactive = false;
        for (int j = tid; j < m; j += nt)
            if (state[j] == 0) { cm = fmax(cm, fabs(Cvec[j])); any_inactive = true; }
        const double Cmax = block_reduce(cm, true, sh.red);
        (void)any_inactive;
        if (!sh.drops_pending) {
            // joiners in ascending index order (:131-133): ballot per warp, prefix over the warps
            int jbase = 0;
            for (int j0 = 0; j0 < m; j0 += nt) {
                const int j = j0 + tid;
                const bool f = j < m && state[j] == 0 && fabs(Cvec[j]) >= Cmax - eps;
                const unsigned bal = __ballot_sync(0xffffffffu, f);
                if ((tid & 31) == 0) sh.ired[tid >> 5] = __popc(bal);
                __syncthreads();
                int before = 0, tot = 0;
                for (int wq = 0; wq < kLarsThreads / 32; wq++) {
                    const int cq = sh.ired[wq];
                    before += (wq < (tid >> 5)) ? cq : 0;
                    tot += cq;
                }
                if (f) joiners[jbase + before + __popc(bal & ((1u << (tid & 31)) - 1u))] = j;
                jbase += tot;
                __syncthreads();
            }
            if (tid == 0) sh.n_join = jbase;
            __syncthreads();
            for (int jn = 0; jn < sh.n_join; jn++) {
                const int inew = joiners[jn];
                const int na = sh.n_act;
                const double xnew = a.Sig[(size_t)inew * m + inew];
                const double nx = sqrt(xnew);
                if (na == 0) {
                    if (tid == 0) {
                        R[0] = nx;
                        act[0] = inew; state[inew] = 1;
                        Sign[0] = (Cvec[inew] > 0.0) ? 1.0 : ((Cvec[inew] < 0.0) ? -1.0 : 0.0);
                        zf[0] = Sign[0] / nx;
                        a.Rinv[0] = 1.0 / nx;
                        gi[0] = zf[0] / nx;
                        sh.n_act = 1;
                    }
                    __syncthreads();
                    continue;
                }
                for (int l = tid; l < na; l += nt) tmp[l] = a.Sig[(size_t)inew * m + act[l]];
                __syncthreads();
                const bool fast = sh.fast != 0;
                if (fast) {  // updateR: r = backsolvet(R, xold) (:19) as the product R^-T xold; thread i owns entry i
                    for (int i = tid; i < na; i += nt) {
                        double acc0 = 0.0, acc1 = 0.0;
                        int l = 0;
                        for (; l + 1 <= i; l += 2) {
                            acc0 = fma(a.Rinv[rpk(l, i, m)], tmp[l], acc0);
                            acc1 = fma(a.Rinv[rpk(l + 1, i, m)], tmp[l + 1], acc1);
                        }
                        if (l <= i) acc0 = fma(a.Rinv[rpk(l, i, m)], tmp[l], acc0);
                        u[i] = acc0 + acc1;
                    }
                    __syncthreads();
                    for (int i = tid; i < na; i += nt) tmp[i] = u[i];
                    __syncthreads();
                } else {
                    forward_solve(R, na, m, tmp);
                }
                double part = 0.0, part2 = 0.0;
                for (int l = tid; l < na; l += nt) {
                    part = fma(tmp[l], tmp[l], part);
                    part2 = fma(tmp[l], zf[l], part2);   // new column of R against R^-T Sign
                }
                const double ss = block_sum(part, sh.red);
                const double rz = block_sum(part2, sh.red);
                double rpp = nx * nx - ss;  // (:20)
                if (rpp <= eps) {
                    // machine-singular: back out and ignore the variable (:138-143; R :168-173)
                    if (tid == 0) state[inew] = 2;
                } else {
                    rpp = sqrt(rpp);
                    for (int l = tid; l < na; l += nt) R[rpk(l, na, m)] = tmp[l];
                    const double sgn_new = (Cvec[inew] > 0.0) ? 1.0 : ((Cvec[inew] < 0.0) ? -1.0 : 0.0);
                    const double z_new = (sgn_new - rz) / rpp;   // last row of the forward substitution R' z = Sign
                    if (fast) {
                        // new column of R^-1: c = -R^-1 r / r_pp, one warp per row (row l of R^-1 is contiguous), and
                        // G <- [G + c z_new ; z_new / r_pp]
                        const int wrp = tid >> 5, lane = tid & 31;
                        for (int l = wrp; l < na; l += (nt >> 5)) {
                            const double* rrow = a.Rinv + rpk(l, l, m);
                            double acc = 0.0;
                            for (int i = l + lane; i < na; i += 32) acc = fma(rrow[i - l], tmp[i], acc);
                            for (int sft = 16; sft >= 1; sft >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sft);
                            if (lane == 0) {
                                const double cl = -acc / rpp;
                                a.Rinv[rpk(l, na, m)] = cl;
                                gi[l] = fma(cl, z_new, gi[l]);
                            }
                        }
                    }
                    if (tid == 0) {
                        R[rpk(na, na, m)] = rpp;
                        act[na] = inew; state[inew] = 1;
                        Sign[na] = sgn_new;
                        zf[na] = z_new;
                        if (fast) {
                            a.Rinv[rpk(na, na, m)] = 1.0 / rpp;
                            gi[na] = z_new / rpp;
                        }
                        sh.n_act = na + 1;
                    }
                }
                __syncthreads();
            }
        }
        const int na = sh.n_act;
        if (na == 0) {  // nothing could join (all remaining variables singular): stop
            if (tid == 0) { sh.k -= 1; sh.err = 1; }
            __syncthreads();
            break;
        }
        // equiangular direction: Gi1 = R^-1 R^-T Sign, A = 1/sqrt(Gi1.Sign), w = A Gi1 (:151-153)
        if (sh.fast) {
            for (int l = tid; l < na; l += nt) tmp[l] = gi[l];   // == R^-1 R^-T Sign
            __syncthreads();
        } else {
            for (int l = tid; l < na; l += nt) tmp[l] = zf[l];   // == R^-T Sign
            __syncthreads();
            backward_solve(R, na, m, tmp);
        }
        double part = 0.0;
        for (int l = tid; l < na; l += nt) part = fma(tmp[l], Sign[l], part);
        const double gs = block_sum(part, sh.red);
        const double A = 1.0 / sqrt(gs);
        for (int l = tid; l < na; l += nt) w[l] = A * tmp[l];
        __syncthreads();
        // u_j = sum_l w_l Sig[act_l][j]  (serves both `a` (:157-160) and the Cvec update (:177))
        for (int j = tid; j < m; j += nt) {
            double s = 0.0;
            for (int l = 0; l < na; l++) s = fma(w[l], a.Sig[(size_t)act[l] * m + j], s);
            u[j] = s;
        }
        __syncthreads();
        double gam = INFINITY;
        if (na < m) {
            for (int j = tid; j < m; j += nt) {
                if (state[j] == 0) {
                    const double cj = Cvec[j], aj = u[j];
                    const double g1 = (Cmax - cj) / (A - aj), g2 = (Cmax + cj) / (A + aj);
                    if (g1 > eps) gam = fmin(gam, g1);
                    if (g2 > eps) gam = fmin(gam, g2);
                }
            }
        }
        double gamhat = block_reduce(gam, false, sh.red);
        gamhat = fmin(gamhat, Cmax / A);  // (:155,162)
        const double* bprev = a.beta_path + (size_t)(k - 1) * m;
        double* bcur = a.beta_path + (size_t)k * m;
        bool drop_now = false;
        if (a.type == DLSA_LARS_LASSO) {  // (:164-173)
            double zm = INFINITY;
            for (int l = tid; l < na; l += nt) {
                const double z1 = -bprev[act[l]] / w[l];
                if (z1 > eps) zm = fmin(zm, z1);
            }
            double zmin = block_reduce(zm, false, sh.red);
            zmin = fmin(zmin, gamhat);
            if (zmin < gamhat) {
                gamhat = zmin;
                drop_now = true;
                for (int l = tid; l < na; l += nt) dropflag[l] = ((-bprev[act[l]] / w[l]) == zmin) ? 1 : 0;
            }
        }
        for (int j = tid; j < m; j += nt) {
            bcur[j] = bprev[j];
            Cvec[j] = fma(-gamhat, u[j], Cvec[j]);
        }
        __syncthreads();
        for (int l = tid; l < na; l += nt) bcur[act[l]] += gamhat * w[l];
        __syncthreads();
        if (drop_now) {  // (:179-186)
            for (int l = na - 1; l >= 0; l--) {
                if (dropflag[l]) {
                    delete_column_w0(R, sh.n_act, m, l);
                    if (tid == 0) {
                        const int na2 = sh.n_act;
                        const int v = act[l];
                        bcur[v] = 0.0;
                        state[v] = 0;
                        for (int t = l; t < na2 - 1; t++) { act[t] = act[t + 1]; Sign[t] = Sign[t + 1]; }
                        sh.n_act = na2 - 1;
                    }
                    __syncthreads();
                }
            }
        }
        if (drop_now) {  // columns were deleted from R: from here on substitution on R; rebuild R^-T Sign for the shrunken factor
            if (tid == 0) sh.fast = 0;
            const int na3 = sh.n_act;
            for (int l = tid; l < na3; l += nt) zf[l] = Sign[l];
            __syncthreads();
            if (na3 > 0) forward_solve(R, na3, m, zf);
        }
        if (tid == 0) sh.drops_pending = drop_now ? 1 : 0;
        __syncthreads();
    }
    __syncthreads();
    if (tid == 0) {
        a.info[0] = sh.k + 1;
        a.info[3] = sh.err;
    }
}

// K7: one CTA per knot: RSS_k = (s - beta_k)' Sig (s - beta_k) (:191-192); scale beta by |b| (:198-201);
// beta0 (R :243-245); dof (:208); AIC/BIC (:209-210).
__global__ void __launch_bounds__(256) lars_ic_kernel(const LarsArgs a) {
    __shared__ double red[8];
    const int row = blockIdx.x;
    if (row >= a.info[0]) return;
    const int off = a.intercept ? 1 : 0, m = a.p0 - off;
    extern __shared__ double d_s[];  // [m] s - beta
    double* brow = a.beta_path + (size_t)row * m;
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
        const double bj = a.b0[j + off];
        const double sg = (bj > 0.0) ? 1.0 : ((bj < 0.0) ? -1.0 : 0.0);
        d_s[j] = sg - brow[j];
    }
    __syncthreads();
    double part = 0.0;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        double s = 0.0;
        const double* srow = a.Sig + (size_t)i * m;
        for (int j = 0; j < m; j++) s = fma(srow[j], d_s[j], s);
        part = fma(d_s[i], s, part);
    }
    for (int s = 16; s >= 1; s >>= 1) part += __shfl_xor_sync(0xffffffffu, part, s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    // scale, dof, beta0
    double b0part = 0.0;
    int dofpart = 0;
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
        const double v = brow[j] * fabs(a.b0[j + off]);
        brow[j] = v;
        if (fabs(v) > a.eps) dofpart++;
        if (a.intercept) b0part = fma(a.Sigma0[(size_t)(j + 1) * a.p0], v, b0part);
    }
    __shared__ double red2[8];
    __shared__ int red3[8];
    for (int s = 16; s >= 1; s >>= 1) {
        b0part += __shfl_xor_sync(0xffffffffu, b0part, s);
        dofpart += __shfl_xor_sync(0xffffffffu, dofpart, s);
    }
    if ((threadIdx.x & 31) == 0) { red2[threadIdx.x >> 5] = b0part; red3[threadIdx.x >> 5] = dofpart; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double rss = 0.0, bb = 0.0;
        int dof = 0;
        for (int w_ = 0; w_ < 8; w_++) { rss += red[w_]; bb += red2[w_]; dof += red3[w_]; }
        a.rss[row] = rss;
        a.dof[row] = dof;
        a.bic[row] = rss + log(a.n) * (double)dof;
        a.aic[row] = rss + 2.0 * (double)dof;
        a.beta0_path[row] = a.intercept ? (a.scal[1] - bb / a.scal[0]) : 0.0;
    }
}

// K7b: first-argmin of AIC and BIC (np.argmin, dlsa/dlsa.py:90-91)
__global__ void lars_select_kernel(const LarsArgs a) {
    if (threadIdx.x == 0) {
        const int n = a.info[0];
        int ia = 0, ib = 0;
        for (int r = 1; r < n; r++) {
            if (a.aic[r] < a.aic[ia]) ia = r;
            if (a.bic[r] < a.bic[ib]) ib = r;
        }
        a.info[1] = ia;
        a.info[2] = ib;
    }
}

static size_t lars_vec_bytes(int m) { return ((size_t)(9 * m) * 8 + (size_t)(3 * m) * 4 + m + 15) / 16 * 16; }

size_t lars_workspace_bytes(int p0, int max_steps) {
    (void)max_steps;
    const size_t m = (size_t)p0;
    return (m * m + 2 * (m * (m + 1) / 2) + 8) * sizeof(double);
}

int lars_launch(const double* Sigma0, const double* b0, int p0, int intercept, double n, int type, double eps,
                int max_steps, double* beta_path, double* beta0_path, double* rss, double* aic, double* bic, int* dof,
                int* info, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    const int m = p0 - (intercept ? 1 : 0);
    DLSA_REQUIRE(m >= 1 && m <= kLarsMaxM, DLSA_E_UNSUPPORTED, "lars: m=%d outside [1,%d]", m, kLarsMaxM);
    DLSA_REQUIRE(workspace != nullptr && workspace_bytes >= lars_workspace_bytes(p0, max_steps), DLSA_E_WORKSPACE,
                 "lars: workspace too small");
    LarsArgs a;
    a.Sigma0 = Sigma0; a.b0 = b0; a.p0 = p0; a.intercept = intercept ? 1 : 0; a.type = type;
    a.max_steps = max_steps >= 0 ? max_steps : 8 * m;  // 0 steps = the zero row only, as the reference's loop (dlsa/lsa.py:126)
    a.n = n; a.eps = eps;
    a.beta_path = beta_path; a.beta0_path = beta0_path; a.rss = rss; a.aic = aic; a.bic = bic; a.dof = dof; a.info = info;
    double* ws = static_cast<double*>(workspace);
    a.scal = ws;
    a.Sig = ws + 8;
    a.Rg = a.Sig + (size_t)p0 * p0;
    a.Rinv = a.Rg + (size_t)p0 * (p0 + 1) / 2;
    const bool r_in_smem = (size_t)m * (m + 1) / 2 <= kLarsRSmemDoubles;
    const size_t smem = lars_vec_bytes(m) + (r_in_smem ? (size_t)m * (m + 1) / 2 * 8 : 0);
    DLSA_CUDA(cudaFuncSetAttribute(lars_path_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lars_path_kernel<<<1, kLarsThreads, smem, st>>>(a);
    count_launch();
    int rc = check_launch("lars_path_kernel");
    if (rc) return rc;
    // rounded up to 16 bytes: the compiler reads d_s two doubles at a time (LDS.128), also across the end of an odd m
    lars_ic_kernel<<<a.max_steps + 1, 256, (size_t)((m + 1) & ~1) * 8, st>>>(a);
    count_launch();
    rc = check_launch("lars_ic_kernel");
    if (rc) return rc;
    lars_select_kernel<<<1, 32, 0, st>>>(a);
    count_launch();
    return check_launch("lars_select_kernel");
}

}  // namespace dlsa
