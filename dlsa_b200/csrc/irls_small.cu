// K1 (p <= 56): fused Newton/IRLS pass, one warp owns the whole lower-triangular Gram.
//
// Reference arithmetic replaced (per partition): eta = X beta, prob = sigmoid(eta), w = prob(1-prob),
// g = X'(y-prob), H = X' diag(w) X  — dlsa/R/wlse_func.R:47-54, dlsa/models.py:114,130.
//
// Design: CTA = 8 independent warps on one partition slice.  Each warp streams its own
// 32-row (or 16-row) chunks of X through a private multi-stage ring in shared memory, filled by TMA 1-D
// bulk copies (one 8*px-byte copy per row into a padded row, stride == 4 mod 8 doubles, so every DMMA
// fragment load is bank-conflict free) and signalled on per-stage mbarriers.  Per chunk: eta by DFMA on a
// (row, column-quarter) lane grid, sigmoid/weights one row per lane, then H += (wX)'X on FP64 tensor cores
// (DMMA.8x8x4) with all NT(NT+1)/2 lower-triangle 8x8 accumulator tiles held in that warp's registers.
// X is read from HBM exactly once per pass.  The 8 warps' accumulators are summed in a fixed order through
// shared memory, so results are bitwise reproducible for a given launch configuration.
#include "common.cuh"

namespace dlsa {

struct IrlsArgs {
    const double* X;
    int64_t ldx;
    const double* y;
    const int64_t* part_off;
    const double* beta;
    const int* active;
    double* outH;   // [K*S][p*p]   (S == 1: the caller's H)
    double* outg;   // [K*S][p]
    double* outll;  // [K*S] or nullptr
    int K, px, p, icpt, S, Sd, aligned, want_ll;
};

constexpr int kWarps = 8;

template <int NT>
__device__ __forceinline__ constexpr int tri_index(int i, int j) { return i * (i + 1) / 2 + j; }

template <int NT, int CH, int STAGES, int FAMILY>
__global__ void __launch_bounds__(kWarps * 32, 1) irls_small_kernel(const IrlsArgs a) {
    constexpr int NTRI = NT * (NT + 1) / 2;
    constexpr int NG = CH / 8;  // 8-row groups per chunk
    constexpr int NRED = NTRI * 64 + NT * 8 + 8;
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int a8 = lane >> 2, q = lane & 3;
    const int k = blockIdx.x / a.S, slice = blockIdx.x % a.S;
    if (a.active != nullptr && a.active[k] == 0) return;

    const int64_t row_lo = a.part_off[k];
    const int64_t n_k = a.part_off[k + 1] - row_lo;
    const int64_t nchunks = (n_k + CH - 1) / CH;
    const int64_t cps = (nchunks + a.S - 1) / a.S;
    const int64_t c_lo = (int64_t)slice * cps;
    const int64_t c_hi = (c_lo + cps < nchunks) ? (c_lo + cps) : nchunks;
    // this warp's chunks: c_lo + warp + i*kWarps, i = 0..my_n-1
    const int64_t first = c_lo + warp;
    const int my_n = (first < c_hi) ? (int)((c_hi - first + kWarps - 1) / kWarps) : 0;

    const int Sd = a.Sd, p = a.p, px = a.px;
    const int stage_doubles = CH * Sd;
    double* tiles = reinterpret_cast<double*>(smem_raw);
    double* my_tiles = tiles + (size_t)warp * STAGES * stage_doubles;
    size_t off = (size_t)kWarps * STAGES * stage_doubles;
    double* ws = tiles + off + (size_t)warp * CH * 2;
    double* rs = ws + CH;
    off += (size_t)kWarps * CH * 2;
    double* beta_s = tiles + off;
    off += NT * 8;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + off) + warp * STAGES;

    // ---- one-time setup ------------------------------------------------------------------------
    if (threadIdx.x < NT * 8) {
        const int c = threadIdx.x;  // internal column: features 0..px-1, then the intercept
        double bv = 0.0;
        if (FAMILY == DLSA_FAMILY_LOGISTIC) {
            if (c < px) bv = a.beta[(size_t)k * p + c + a.icpt];
            else if (c == px && a.icpt) bv = a.beta[(size_t)k * p];
        }
        beta_s[c] = bv;
    }
    if (a.icpt) {
        for (int r = lane; r < STAGES * CH; r += 32) my_tiles[(size_t)r * Sd + px] = 1.0;
    }
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) mbar_init(smem_u32(&bars[s]), a.aligned ? 1u : 32u);
        fence_mbar_init();
    }
    fence_proxy_async();
    __syncthreads();

    const double* Xp = a.X + row_lo * a.ldx;
    const double* yp = a.y + row_lo;
    const uint32_t row_bytes = (uint32_t)px * 8u;

    auto issue = [&](int ci) {
        if (ci >= my_n) return;
        const int st = ci % STAGES;
        const int64_t r0 = (first + (int64_t)ci * kWarps) * CH;
        const int valid = (n_k - r0 < CH) ? (int)(n_k - r0) : CH;
        const uint32_t bar = smem_u32(&bars[st]);
        double* dst = my_tiles + (size_t)st * stage_doubles;
        if (a.aligned) {
            if (lane == 0) mbar_arrive_expect_tx(bar, row_bytes * CH);
            __syncwarp();
            if (lane < CH) {
                const int rr = lane < valid ? lane : valid - 1;  // tail rows re-read a valid row; their weight is 0
                bulk_g2s(smem_u32(dst + (size_t)lane * Sd), Xp + (r0 + rr) * a.ldx, row_bytes, bar);
            }
        } else {
            for (int e = lane; e < CH * px; e += 32) {
                const int row = e / px, col = e - row * px;
                const int rr = row < valid ? row : valid - 1;
                cp_async8(smem_u32(dst + (size_t)row * Sd + col), Xp + (r0 + rr) * a.ldx + col);
            }
            cp_async_mbar_arrive_noinc(bar);
        }
    };
    // row this lane owns in the per-row (sigmoid) stage: rho = 8*q + a8 when q < NG
    const int rho = 8 * q + a8;
    auto load_y = [&](int ci) -> double {
        if (ci >= my_n || q >= NG) return 0.0;
        const int64_t r = (first + (int64_t)ci * kWarps) * CH + rho;
        return (r < n_k) ? __ldg(yp + r) : 0.0;
    };

    double C[NTRI][2];
    double gacc[NT];
#pragma unroll
    for (int t = 0; t < NTRI; t++) C[t][0] = C[t][1] = 0.0;
#pragma unroll
    for (int t = 0; t < NT; t++) gacc[t] = 0.0;
    double llacc = 0.0;

    // column masks of the last (ragged) 8-column tile
    const bool ok_last = (8 * (NT - 1) + a8) < p;
    const int col_last = ok_last ? 8 * (NT - 1) + a8 : 0;

#pragma unroll
    for (int s = 0; s < STAGES; s++) issue(s);
    double y_next = load_y(0);

    for (int ci = 0; ci < my_n; ci++) {
        const int st = ci % STAGES;
        const double y_cur = y_next;
        y_next = load_y(ci + 1);
        const int64_t r0 = (first + (int64_t)ci * kWarps) * CH;
        const int valid = (n_k - r0 < CH) ? (int)(n_k - r0) : CH;
        mbar_wait(smem_u32(&bars[st]), (uint32_t)((ci / STAGES) & 1));
        const double* T = my_tiles + (size_t)st * stage_doubles;

        // ---- eta = x . beta on a (row = 8g + a8, column quarter q) lane grid ---------------------
        double wv_row = 0.0, rv_row = 0.0;
        if (FAMILY == DLSA_FAMILY_LOGISTIC) {
            double s[NG];
#pragma unroll
            for (int g = 0; g < NG; g++) s[g] = 0.0;
#pragma unroll
            for (int t = 0; t < 2 * NT; t++) {
                const int col = 4 * t + q;
                const bool ok = (t < 2 * NT - 2) || (col < p);
                const int colc = ok ? col : 0;
                const double bv = beta_s[col];
#pragma unroll
                for (int g = 0; g < NG; g++) {
                    double xv = T[(size_t)(8 * g + a8) * Sd + colc];
                    xv = ok ? xv : 0.0;
                    s[g] = fma(xv, bv, s[g]);
                }
            }
            double eta = 0.0;
#pragma unroll
            for (int g = 0; g < NG; g++) {
                s[g] += shfl_xor_f64(s[g], 1);
                s[g] += shfl_xor_f64(s[g], 2);
                eta = (q == g) ? s[g] : eta;
            }
            const double pr = sigmoid_f64(eta);
            wv_row = pr * (1.0 - pr);
            rv_row = y_cur - pr;
            if (a.want_ll && q < NG && rho < valid) llacc += y_cur * eta - log1pexp_f64(eta);
        } else {
            wv_row = 1.0;
            rv_row = y_cur;
            if (q < NG && rho < valid) llacc = fma(y_cur, y_cur, llacc);
        }
        if (q < NG) {
            const bool rowvalid = rho < valid;
            ws[rho] = rowvalid ? wv_row : 0.0;
            rs[rho] = rowvalid ? rv_row : 0.0;
        }
        __syncwarp();

        // ---- H += (wX)'X on DMMA, g += X'r on DFMA ----------------------------------------------
#pragma unroll 2
        for (int s4 = 0; s4 < CH / 4; s4++) {
            const int row = 4 * s4 + q;
            const double wv = ws[row], rv = rs[row];
            const double* Tr = T + (size_t)row * Sd;
            double f[NT], av[NT];
#pragma unroll
            for (int t = 0; t < NT - 1; t++) f[t] = Tr[8 * t + a8];
            {
                const double v = Tr[col_last];
                f[NT - 1] = ok_last ? v : 0.0;
            }
#pragma unroll
            for (int t = 0; t < NT; t++) {
                av[t] = wv * f[t];
                gacc[t] = fma(rv, f[t], gacc[t]);
            }
#pragma unroll
            for (int i = 0; i < NT; i++)
#pragma unroll
                for (int j = 0; j <= i; j++) dmma884(C[i * (i + 1) / 2 + j][0], C[i * (i + 1) / 2 + j][1], av[i], f[j]);
        }
        __syncwarp();
        // refill this stage with the chunk STAGES ahead (generic reads above -> async writes: proxy fence)
        if (ci + STAGES < my_n) {
            fence_proxy_async();
            issue(ci + STAGES);
        }
    }

    // ---- CTA reduction in a fixed order, then emit ------------------------------------------------
#pragma unroll
    for (int t = 0; t < NT; t++) {
        gacc[t] += shfl_xor_f64(gacc[t], 1);
        gacc[t] += shfl_xor_f64(gacc[t], 2);
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) llacc += shfl_xor_f64(llacc, m);

    __syncthreads();  // every warp is done with its tiles; reuse them as the reduction buffer
    double* red = tiles + (size_t)warp * NRED;
#pragma unroll
    for (int t = 0; t < NTRI; t++) {
        red[t * 64 + a8 * 8 + 2 * q] = C[t][0];
        red[t * 64 + a8 * 8 + 2 * q + 1] = C[t][1];
    }
    if (q == 0) {
#pragma unroll
        for (int t = 0; t < NT; t++) red[NTRI * 64 + 8 * t + a8] = gacc[t];
    }
    if (lane == 0) red[NTRI * 64 + NT * 8] = llacc;
    __syncthreads();

    double* oH = a.outH + (size_t)blockIdx.x * p * p;
    double* og = a.outg + (size_t)blockIdx.x * p;
    for (int e = threadIdx.x; e < NTRI * 64 + NT * 8 + 1; e += kWarps * 32) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kWarps; w++) v += tiles[(size_t)w * NRED + e];
        if (e < NTRI * 64) {
            const int tt = e >> 6, m = (e >> 3) & 7, n = e & 7;
            int i = 0;
            while ((i + 1) * (i + 2) / 2 <= tt) i++;
            const int j = tt - i * (i + 1) / 2;
            const int r_ = 8 * i + m, c_ = 8 * j + n;
            if (r_ < p && c_ < p && (i != j || n <= m)) {
                const int ro = a.icpt ? (r_ == px ? 0 : r_ + 1) : r_;
                const int co = a.icpt ? (c_ == px ? 0 : c_ + 1) : c_;
                oH[(size_t)ro * p + co] = v;
                oH[(size_t)co * p + ro] = v;
            }
        } else if (e < NTRI * 64 + NT * 8) {
            const int c_ = e - NTRI * 64;
            if (c_ < p) {
                const int co = a.icpt ? (c_ == px ? 0 : c_ + 1) : c_;
                og[co] = v;
            }
        } else if (a.outll != nullptr) {
            a.outll[blockIdx.x] = v;
        }
    }
}

// sum the S slice partials of every partition in a fixed order
__global__ void slice_reduce_kernel(const double* __restrict__ pH, const double* __restrict__ pg,
                                    const double* __restrict__ pll, const int* __restrict__ active, int K, int S,
                                    int p, double* __restrict__ H, double* __restrict__ g, double* __restrict__ ll) {
    const int k = blockIdx.y;
    if (active != nullptr && active[k] == 0) return;
    const int per = p * p + p + 1;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < per; e += gridDim.x * blockDim.x) {
        double v = 0.0;
        if (e < p * p) {
            for (int s = 0; s < S; s++) v += pH[((size_t)k * S + s) * p * p + e];
            H[(size_t)k * p * p + e] = v;
        } else if (e < p * p + p) {
            for (int s = 0; s < S; s++) v += pg[((size_t)k * S + s) * p + (e - p * p)];
            g[(size_t)k * p + (e - p * p)] = v;
        } else if (ll != nullptr) {
            for (int s = 0; s < S; s++) v += pll[(size_t)k * S + s];
            ll[k] = v;
        }
    }
}

// ---- host side ---------------------------------------------------------------------------------------
static int row_stride_doubles(int p) {
    int s = p;
    while (s % 8 != 4) s++;
    return s;
}

struct SmallCfg {
    int NT, CH, STAGES, Sd;
    size_t smem;
};

static SmallCfg small_cfg(int p) {
    SmallCfg c;
    c.NT = (p + 7) / 8;
    c.Sd = row_stride_doubles(p);
    if (c.NT <= 4) { c.CH = 32; c.STAGES = 3; }
    else if (c.Sd <= 52) { c.CH = 32; c.STAGES = 2; }
    else { c.CH = 16; c.STAGES = 3; }
    c.smem = ((size_t)kWarps * c.STAGES * c.CH * c.Sd + (size_t)kWarps * c.CH * 2 + c.NT * 8) * 8 + kWarps * c.STAGES * 8 + 64;
    return c;
}

int irls_small_slices(int K, int p, int64_t max_part_rows) {
    SmallCfg c = small_cfg(p);
    int64_t chunks = (max_part_rows + c.CH - 1) / c.CH;
    int64_t s_want = (148 * 4 + K - 1) / K;
    int64_t s_max = chunks / (kWarps * 2);
    if (s_max < 1) s_max = 1;
    int64_t s = s_want < s_max ? s_want : s_max;
    return (int)(s < 1 ? 1 : s);
}

template <int NT, int CH, int STAGES>
static int launch_cfg(const IrlsArgs& a, int family, size_t smem, int grid, cudaStream_t st) {
    if (family == DLSA_FAMILY_LOGISTIC) {
        auto kern = irls_small_kernel<NT, CH, STAGES, DLSA_FAMILY_LOGISTIC>;
        DLSA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, kWarps * 32, smem, st>>>(a);
    } else {
        auto kern = irls_small_kernel<NT, CH, STAGES, DLSA_FAMILY_LINEAR>;
        DLSA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, kWarps * 32, smem, st>>>(a);
    }
    count_launch();
    return check_launch("irls_small_kernel");
}

int irls_small_launch(const double* X, int64_t ldx, const double* y, const int64_t* part_off, int K, int px,
                      int icpt, int family, const double* beta, const int* active, double* g, double* H,
                      double* loglik, int64_t max_part_rows, void* workspace, size_t workspace_bytes,
                      cudaStream_t st) {
    const int p = px + (icpt ? 1 : 0);
    SmallCfg c = small_cfg(p);
    const int S = irls_small_slices(K, p, max_part_rows);
    IrlsArgs a;
    a.X = X; a.ldx = ldx; a.y = y; a.part_off = part_off; a.beta = beta; a.active = active;
    a.K = K; a.px = px; a.p = p; a.icpt = icpt ? 1 : 0; a.S = S; a.Sd = c.Sd;
    a.aligned = ((reinterpret_cast<uintptr_t>(X) & 15) == 0 && (ldx % 2) == 0 && (px % 2) == 0) ? 1 : 0;
    a.want_ll = loglik != nullptr;
    double *pH = nullptr, *pg = nullptr, *pll = nullptr;
    if (S == 1) {
        a.outH = H; a.outg = g; a.outll = loglik;
    } else {
        const size_t need = (size_t)K * S * ((size_t)p * p + p + 1) * sizeof(double);
        DLSA_REQUIRE(workspace != nullptr && workspace_bytes >= need, DLSA_E_WORKSPACE,
                     "irls pass: workspace too small (%zu < %zu)", workspace_bytes, need);
        pH = static_cast<double*>(workspace);
        pg = pH + (size_t)K * S * p * p;
        pll = pg + (size_t)K * S * p;
        a.outH = pH; a.outg = pg; a.outll = pll;
    }
    const int grid = K * S;
    int rc = 0;
#define DLSA_CASE(NT_, CH_, ST_) rc = launch_cfg<NT_, CH_, ST_>(a, family, c.smem, grid, st)
    if (c.CH == 32 && c.STAGES == 3) {
        switch (c.NT) {
            case 1: DLSA_CASE(1, 32, 3); break;
            case 2: DLSA_CASE(2, 32, 3); break;
            case 3: DLSA_CASE(3, 32, 3); break;
            case 4: DLSA_CASE(4, 32, 3); break;
            default: DLSA_REQUIRE(false, DLSA_E_UNSUPPORTED, "irls small: bad config");
        }
    } else if (c.CH == 32 && c.STAGES == 2) {
        switch (c.NT) {
            case 5: DLSA_CASE(5, 32, 2); break;
            case 6: DLSA_CASE(6, 32, 2); break;
            case 7: DLSA_CASE(7, 32, 2); break;
            default: DLSA_REQUIRE(false, DLSA_E_UNSUPPORTED, "irls small: bad config");
        }
    } else {
        DLSA_REQUIRE(c.NT == 7, DLSA_E_UNSUPPORTED, "irls small: bad config");
        DLSA_CASE(7, 16, 3);
    }
#undef DLSA_CASE
    if (rc != 0) return rc;
    if (S > 1) {
        dim3 grid2((p * p + p + 1 + 255) / 256, K);
        slice_reduce_kernel<<<grid2, 256, 0, st>>>(pH, pg, loglik ? pll : nullptr, active, K, S, p, H, g, loglik);
        count_launch();
        return check_launch("slice_reduce_kernel");
    }
    return 0;
}

size_t irls_small_workspace_bytes(int K, int p, int64_t max_part_rows) {
    const int S = irls_small_slices(K, p, max_part_rows);
    if (S == 1) return 0;
    return (size_t)K * S * ((size_t)p * p + p + 1) * sizeof(double);
}

}  // namespace dlsa
