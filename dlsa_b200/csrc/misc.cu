// a1 synthetic data generator, f1 log-likelihood evaluation, a11 one-hot expansion.
#include "common.cuh"

namespace dlsa {

// ---- counter-based generator: value = f(seed, global row, column) ---------------------------------------
// splitmix64 finaliser applied to a keyed counter; dlsa_b200/synth.py holds the identical NumPy restatement.
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ double u01(uint64_t seed, uint64_t row, uint64_t col) {
    uint64_t z = mix64(seed + row * 0x9E3779B97F4A7C15ULL);
    z = mix64(z + col * 0xD1B54A32D192ED03ULL + 0x8CB92BA72F3D8DD7ULL);
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);  // [0, 1), 53 bits
}
__device__ __forceinline__ double gauss(uint64_t seed, uint64_t row, uint64_t col) {
    const double u1 = u01(seed, row, 2 * col), u2 = u01(seed, row, 2 * col + 1);
    return sqrt(-2.0 * log(1.0 - u1)) * cospi(2.0 * u2);
}

// one warp per row: X coalesced, label from the row's linear predictor
__global__ void simulate_kernel(double* __restrict__ X, int64_t ldx, double* __restrict__ y, int64_t row0, int64_t n,
                                int px, const double* __restrict__ beta_true, uint64_t seed, int law) {
    const int lane = threadIdx.x & 31;
    const int64_t wpg = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n; r += wpg) {
        const uint64_t grow = (uint64_t)(row0 + r);
        double dot = 0.0;
        for (int c = lane; c < px; c += 32) {
            const double v = (law == 0) ? (u01(seed, grow, (uint64_t)c) - 0.5) : gauss(seed, grow, (uint64_t)c);
            X[r * ldx + c] = v;
            dot = fma(v, beta_true[c], dot);
        }
        for (int m = 16; m >= 1; m >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, m);
        if (lane == 0) {
            if (law == 0) {
                const double u = u01(seed, grow, 0xFFFFFFFFULL);
                y[r] = (u < sigmoid_f64(dot)) ? 1.0 : 0.0;
            } else {
                y[r] = dot + gauss(seed, grow, 0x7FFFFFFFULL);
            }
        }
    }
}

int simulate_launch(double* X, int64_t ldx, double* y, int64_t row0, int64_t n, int px, const double* beta_true,
                    uint64_t seed, int law, cudaStream_t st) {
    if (n == 0) return 0;
    int64_t blocks = (n + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    simulate_kernel<<<(int)blocks, 256, 0, st>>>(X, ldx, y, row0, n, px, beta_true, seed, law);
    count_launch();
    return check_launch("simulate_kernel");
}

// ---- FP64 peak probe: what this GPU's DMMA pipe delivers right now (the roofline denominator bench.py reports) ------
// 16 warps per SM, 8 independent DMMA.8x8x4 accumulator chains per warp, registers only.
__global__ void __launch_bounds__(512, 1) fp64_probe_kernel(double* __restrict__ sink, int iters) {
    double c0[8], c1[8];
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; i++) { c0[i] = 0.0; c1[i] = 0.0; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) dmma884(c0[i], c1[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c0[i] + c1[i];
    sink[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int fp64_probe_launch(int ctas, int iters, double* sink, cudaStream_t st) {
    fp64_probe_kernel<<<ctas, 512, 0, st>>>(sink, iters);
    count_launch();
    return check_launch("fp64_probe_kernel");
}

// ---- airline-shaped compact design (BASELINE.json configs[2]), counter-based like simulate_kernel ----------------
// One thread per row: n_num numerics ~ N(0,1) (columns 0..n_num-1 of the counter space), one uniform per categorical
// column (column 1000 + c) pushed through that column's cumulative level shares (level 0 = baseline -> code -1), the
// label from the row's linear predictor (column 0xFFFFFFFF).  Rows depend on (seed, global row) only, so any sharding
// of the rows over GPUs sees the same data.
struct AirlineArgs {
    double* numeric; int64_t ldn; int* codes; double* y;
    int64_t row0, n;
    int n_num, n_cat;
    const int* levels;      // [n_cat] levels per column INCLUDING the baseline
    const double* cdf;      // concatenated cumulative shares, levels[c] entries per column
    const double* beta_num; // [n_num]
    const double* beta_lev; // concatenated, levels[c]-1 entries per column (non-baseline levels)
    double beta0;
    uint64_t seed;
};
__global__ void __launch_bounds__(256) airline_kernel(const AirlineArgs a) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < a.n; r += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t grow = (uint64_t)(a.row0 + r);
        double eta = a.beta0;
        for (int c = 0; c < a.n_num; c++) {
            const double v = gauss(a.seed, grow, (uint64_t)c);
            a.numeric[r * a.ldn + c] = v;
            eta = fma(v, a.beta_num[c], eta);
        }
        int cb = 0, lb = 0;
        for (int c = 0; c < a.n_cat; c++) {
            const int L = a.levels[c];
            const double u = u01(a.seed, grow, 1000ULL + (uint64_t)c);
            int lo = 0, hi = L - 1;        // first level whose cumulative share exceeds u (searchsorted, side = right)
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (u < a.cdf[cb + mid]) hi = mid; else lo = mid + 1;
            }
            a.codes[r * a.n_cat + c] = lo - 1;
            if (lo > 0) eta += a.beta_lev[lb + lo - 1];
            cb += L;
            lb += L - 1;
        }
        const double u = u01(a.seed, grow, 0xFFFFFFFFULL);
        a.y[r] = (u < sigmoid_f64(eta)) ? 1.0 : 0.0;
    }
}

int airline_launch(double* numeric, int64_t ldn, int* codes, double* y, int64_t row0, int64_t n, int n_num, int n_cat,
                   const int* levels, const double* cdf, double beta0, const double* beta_num, const double* beta_lev,
                   uint64_t seed, cudaStream_t st) {
    if (n == 0) return 0;
    AirlineArgs a;
    a.numeric = numeric; a.ldn = ldn; a.codes = codes; a.y = y; a.row0 = row0; a.n = n; a.n_num = n_num; a.n_cat = n_cat;
    a.levels = levels; a.cdf = cdf; a.beta_num = beta_num; a.beta_lev = beta_lev; a.beta0 = beta0; a.seed = seed;
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    airline_kernel<<<(int)blocks, 256, 0, st>>>(a);
    count_launch();
    return check_launch("airline_kernel");
}

// ---- f1: per-partition log-likelihood of q coefficient vectors (dlsa/models.py:216-222) -------------------
// y log(prob) + (1-y) log(1-prob) == y*eta - log(1+exp(eta)), evaluated in the overflow-free form.
constexpr int kEvalMaxQ = 8;
__global__ void __launch_bounds__(256) loglik_kernel(const double* __restrict__ X, int64_t ldx,
                                                     const double* __restrict__ y,
                                                     const int64_t* __restrict__ part_off, int px, int icpt,
                                                     const double* __restrict__ par, int q, int q0, int slices,
                                                     double* __restrict__ partial) {
    extern __shared__ double par_s[];  // [qq][p]
    __shared__ double red[8][kEvalMaxQ];
    const int p = px + icpt;
    const int qq = (q - q0 < kEvalMaxQ) ? (q - q0) : kEvalMaxQ;
    for (int e = threadIdx.x; e < qq * p; e += blockDim.x) {
        const int c = e / p, i = e - c * p;
        par_s[e] = par[(size_t)i * q + q0 + c];
    }
    __syncthreads();
    const int k = blockIdx.x / slices, sl = blockIdx.x % slices;
    const int64_t lo = part_off[k], n_k = part_off[k + 1] - lo;
    const int64_t per = (n_k + slices - 1) / slices;
    const int64_t r_lo = sl * per, r_hi = (r_lo + per < n_k) ? r_lo + per : n_k;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double acc[kEvalMaxQ];
#pragma unroll
    for (int c = 0; c < kEvalMaxQ; c++) acc[c] = 0.0;
    for (int64_t r = r_lo + warp; r < r_hi; r += 8) {
        const double* xr = X + (lo + r) * ldx;
        double s[kEvalMaxQ];
#pragma unroll
        for (int c = 0; c < kEvalMaxQ; c++) s[c] = 0.0;
        for (int j = lane; j < px; j += 32) {
            const double xv = xr[j];
#pragma unroll
            for (int c = 0; c < kEvalMaxQ; c++)
                if (c < qq) s[c] = fma(xv, par_s[c * p + j + icpt], s[c]);
        }
        const double yv = y[lo + r];
#pragma unroll
        for (int c = 0; c < kEvalMaxQ; c++) {
            if (c < qq) {
                double v = s[c];
                for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
                if (icpt) v += par_s[c * p];
                acc[c] += yv * v - log1pexp_f64(v);
            }
        }
    }
    if (lane == 0)
        for (int c = 0; c < kEvalMaxQ; c++) red[warp][c] = acc[c];
    __syncthreads();
    if (threadIdx.x < qq) {
        double v = 0.0;
        for (int w = 0; w < 8; w++) v += red[w][threadIdx.x];
        partial[((size_t)blockIdx.x) * q + q0 + threadIdx.x] = v;
    }
}

// out[k][c] = sum over slices (fixed order)
__global__ void loglik_reduce_kernel(const double* __restrict__ partial, int K, int slices, int q,
                                     double* __restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= K * q) return;
    const int k = e / q, c = e - k * q;
    double v = 0.0;
    for (int s = 0; s < slices; s++) v += partial[((size_t)k * slices + s) * q + c];
    out[e] = v;
}

int loglik_slices(int K) {
    int s = (148 * 4 + K - 1) / K;
    return s < 1 ? 1 : (s > 64 ? 64 : s);
}

int loglik_launch(const double* X, int64_t ldx, const double* y, const int64_t* part_off, int K, int px, int icpt,
                  const double* par, int q, double* out, double* partial, cudaStream_t st) {
    if (K == 0 || q == 0) return 0;
    const int p = px + icpt;
    const int slices = loglik_slices(K);
    for (int q0 = 0; q0 < q; q0 += kEvalMaxQ) {
        const size_t smem = (size_t)kEvalMaxQ * p * sizeof(double);
        DLSA_CUDA(cudaFuncSetAttribute(loglik_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        loglik_kernel<<<K * slices, 256, smem, st>>>(X, ldx, y, part_off, px, icpt, par, q, q0, slices, partial);
        count_launch();
        int rc = check_launch("loglik_kernel");
        if (rc) return rc;
    }
    loglik_reduce_kernel<<<(K * q + 255) / 256, 256, 0, st>>>(partial, K, slices, q, out);
    count_launch();
    return check_launch("loglik_reduce_kernel");
}

// ---- a11: dense one-hot expansion in the reference's column layout (dlsa/models.py:56-79, 99-104) ----------
__global__ void onehot_kernel(const double* __restrict__ numeric, int n_num, const double* __restrict__ mean,
                              const double* __restrict__ sd, const int* __restrict__ codes, int n_cat,
                              const int* __restrict__ col_base, int64_t n, double* __restrict__ X, int64_t ldx) {
    const int width = n_num + n_cat;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n * width;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / width;
        const int c = (int)(e - r * width);
        if (c < n_num) {
            double v = numeric[r * n_num + c];
            if (mean != nullptr) v = (v - mean[c]) / sd[c];
            X[r * ldx + c] = v;
        } else {
            const int cc = c - n_num;
            const int code = codes[r * n_cat + cc];
            if (code >= 0) X[r * ldx + col_base[cc] + code] = 1.0;
        }
    }
}

int onehot_launch(const double* numeric, int n_num, const double* mean, const double* sd, const int* codes, int n_cat,
                  const int* col_base, int64_t n, double* X, int64_t ldx, cudaStream_t st) {
    if (n == 0 || n_num + n_cat == 0) return 0;
    int64_t blocks = (n * (n_num + n_cat) + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    onehot_kernel<<<(int)blocks, 256, 0, st>>>(numeric, n_num, mean, sd, codes, n_cat, col_base, n, X, ldx);
    count_launch();
    return check_launch("onehot_kernel");
}

// ---- f-3: describe() — count / mean / sample stddev / min / max of every column in ONE streaming pass -------------
// Reference: data_info = data_sdf_i.describe().toPandas() (projects/logistic_dlsa.py:279-283), used as
// (x - mean) / stddev in logistic_model (dlsa/models.py:99-101).  HBM-bound: 8*px bytes per row, read once.
// Sums are taken about column c's first value (shifted data: no cancellation for columns like Year ~ 2000 +- 5),
// per CTA in a fixed order, and the CTA partials are merged in a fixed order: bitwise reproducible.
constexpr int kStatCtas = 148 * 4;
constexpr int kStatWarps = 8;
// kStatNS (template): 32-column stripes per sweep — a lane owns columns lane and lane+32 of a sweep (1 stripe for p <= 32)
constexpr int kStatRU = 4;   // rows in flight per lane and stripe
// CTA b takes the 32-row tiles b, b + gridDim.x, ... (all CTAs sweep the matrix front together: sequential DRAM
// streaming); a sweep covers 64 columns, so p <= 64 reads every row once, whole, in one go.
template <int kStatNS>
__global__ void __launch_bounds__(kStatWarps * 32, 4) colstats_partial_kernel(const double* __restrict__ X, int64_t ldx,
                                                                           int64_t n, int px,
                                                                           double* __restrict__ partial) {
    __shared__ double red[kStatWarps][32][4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kTile = kStatWarps * kStatRU;  // rows per CTA step
    for (int c0 = 0; c0 < px; c0 += 32 * kStatNS) {
        double shift[kStatNS], s1[kStatNS], s2[kStatNS], mn[kStatNS], mx[kStatNS];
        bool in[kStatNS];
#pragma unroll
        for (int s = 0; s < kStatNS; s++) {
            const int c = c0 + 32 * s + lane;
            in[s] = c < px;
            shift[s] = in[s] ? __ldg(X + c) : 0.0;
            s1[s] = s2[s] = 0.0;
            mn[s] = INFINITY;
            mx[s] = -INFINITY;
        }
        const double* xc = X + c0 + lane;
        for (int64_t t = blockIdx.x; t * kTile < n; t += gridDim.x) {
            const int64_t r0 = t * kTile + warp;
            double v[kStatRU][kStatNS];
#pragma unroll
            for (int u = 0; u < kStatRU; u++) {
                const int64_t r = r0 + (int64_t)u * kStatWarps;
#pragma unroll
                for (int s = 0; s < kStatNS; s++)
                    v[u][s] = (in[s] && r < n) ? __ldg(xc + r * ldx + 32 * s) : shift[s];
            }
#pragma unroll
            for (int u = 0; u < kStatRU; u++) {
                const bool rv = r0 + (int64_t)u * kStatWarps < n;
#pragma unroll
                for (int s = 0; s < kStatNS; s++) {
                    const double d = v[u][s] - shift[s];  // 0 for rows past the end
                    s1[s] += d;
                    s2[s] = fma(d, d, s2[s]);
                    mn[s] = rv ? fmin(mn[s], v[u][s]) : mn[s];
                    mx[s] = rv ? fmax(mx[s], v[u][s]) : mx[s];
                }
            }
        }
#pragma unroll
        for (int s = 0; s < kStatNS; s++) {
            red[warp][lane][0] = s1[s]; red[warp][lane][1] = s2[s]; red[warp][lane][2] = mn[s]; red[warp][lane][3] = mx[s];
            __syncthreads();
            if (warp == 0 && in[s]) {
                double a1 = 0.0, a2 = 0.0, amn = INFINITY, amx = -INFINITY;
                for (int w = 0; w < kStatWarps; w++) {
                    a1 += red[w][lane][0]; a2 += red[w][lane][1];
                    amn = fmin(amn, red[w][lane][2]); amx = fmax(amx, red[w][lane][3]);
                }
                double* o = partial + ((size_t)blockIdx.x * px + c0 + 32 * s + lane) * 4;
                o[0] = a1; o[1] = a2; o[2] = amn; o[3] = amx;
            }
            __syncthreads();
        }
    }
}

// stats [5][px]: count, mean, stddev (n-1 denominator, Spark's stddev_samp), min, max.  One warp per column: lane l sums
// the partials of CTAs l, l+32, ... in order, then a fixed shuffle tree.
__global__ void __launch_bounds__(128) colstats_final_kernel(const double* __restrict__ partial,
                                                             const double* __restrict__ X, int ctas, int64_t n, int px,
                                                             double* __restrict__ stats) {
    const int c = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c >= px) return;
    double s1 = 0.0, s2 = 0.0, mn = INFINITY, mx = -INFINITY;
    for (int b = lane; b < ctas; b += 32) {
        const double* o = partial + ((size_t)b * px + c) * 4;
        s1 += o[0]; s2 += o[1];
        mn = fmin(mn, o[2]); mx = fmax(mx, o[3]);
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        s1 += shfl_xor_f64(s1, m); s2 += shfl_xor_f64(s2, m);
        mn = fmin(mn, shfl_xor_f64(mn, m)); mx = fmax(mx, shfl_xor_f64(mx, m));
    }
    if (lane != 0) return;
    const double nn = (double)n;
    const double shift = X[c];
    const double var = (n > 1) ? fmax(s2 - s1 * s1 / nn, 0.0) / (nn - 1.0) : NAN;
    stats[c] = nn;
    stats[px + c] = shift + s1 / nn;
    stats[2 * px + c] = sqrt(var);
    stats[3 * px + c] = mn;
    stats[4 * px + c] = mx;
}

size_t colstats_workspace_bytes(int px) { return (size_t)kStatCtas * px * 4 * sizeof(double); }

int colstats_launch(const double* X, int64_t ldx, int64_t n, int px, double* stats, double* partial, cudaStream_t st) {
    int64_t ctas = (n + kStatWarps * kStatRU - 1) / (kStatWarps * kStatRU);
    if (ctas > kStatCtas) ctas = kStatCtas;
    if (px <= 32) colstats_partial_kernel<1><<<(int)ctas, kStatWarps * 32, 0, st>>>(X, ldx, n, px, partial);
    else colstats_partial_kernel<2><<<(int)ctas, kStatWarps * 32, 0, st>>>(X, ldx, n, px, partial);
    count_launch();
    int rc = check_launch("colstats_partial_kernel");
    if (rc) return rc;
    colstats_final_kernel<<<(px + 3) / 4, 128, 0, st>>>(partial, X, (int)ctas, n, px, stats);
    count_launch();
    return check_launch("colstats_final_kernel");
}

}  // namespace dlsa
