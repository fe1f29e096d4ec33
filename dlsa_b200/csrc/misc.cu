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

}  // namespace dlsa
