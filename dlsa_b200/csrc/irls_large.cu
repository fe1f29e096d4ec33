// K1 for 56 < p <= 512: fused Newton/IRLS pass with the Gram tiled across the warps of a CTA.
//
// Same arithmetic as irls_small.cu (dlsa/R/wlse_func.R:47-54, dlsa/models.py:114,130).  The lower triangle of
// H has NT(NT+1)/2 8x8 tiles (NT = ceil(p/8)), too many for one warp's registers, so all 8 warps of a CTA work
// on the SAME 16-row chunk (staged by one TMA bulk copy per chunk into a 3-stage ring) and each warp owns up to
// NSEG "segments" = runs of <= SEG consecutive tiles of one tile row: a segment needs one weighted A fragment and
// SEG B fragments per k-step.  When the triangle does not fit one CTA's registers (p > ~180) the tile set is
// split into TG tile-groups, one CTA each per row slice (X is then re-read TG times, mostly from L2; the pass is
// FP64-bound by a wide margin at those p).  Every output tile is owned by exactly one warp of one CTA, so there
// is no cross-warp reduction; slices of a partition are summed in a fixed order by slice_reduce_large_kernel.
#include "common.cuh"

namespace dlsa {

constexpr int kLW = 8;     // warps per CTA
constexpr int kLCH = 16;   // rows per chunk
constexpr int kLST = 3;    // ring stages
constexpr int kNSEG = 4;   // segments per warp
constexpr int kSEG = 8;    // tiles per segment
constexpr int kLargeMaxP = 512;

struct LargeArgs {
    const double* X;
    int64_t ldx;
    const double* y;
    const int64_t* part_off;
    const double* beta;
    const int* active;
    double* outH;   // [K*S][p*p]
    double* outg;   // [K*S][p]
    double* outll;  // [K*S] or nullptr
    int K, px, p, icpt, S, Sd, aligned, want_ll, NT, TG, rstep;
};

__host__ __device__ inline int segs_in_row(int i) { return (i + 1 + kSEG - 1) / kSEG; }

template <int FAMILY>
__global__ void __launch_bounds__(kLW * 32, 1) irls_large_kernel(const LargeArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int a8 = lane >> 2, q = lane & 3;
    const int item = blockIdx.x;           // (partition, slice)
    const int tg = blockIdx.y;             // tile group
    const int k = item / a.S, slice = item % a.S;
    if (a.active != nullptr && a.active[k] == 0) return;

    const int p = a.p, px = a.px, Sd = a.Sd, NT = a.NT;
    const int64_t row_lo = a.part_off[k];
    const int64_t n_k = a.part_off[k + 1] - row_lo;
    const int64_t nchunks = (n_k + kLCH - 1) / kLCH;
    const int64_t cps = (nchunks + a.S - 1) / a.S;
    const int64_t c_lo = (int64_t)slice * cps;
    const int64_t c_hi = (c_lo + cps < nchunks) ? (c_lo + cps) : nchunks;
    const int my_n = (c_hi > c_lo) ? (int)(c_hi - c_lo) : 0;

    const int stage_doubles = kLCH * Sd;
    double* tiles = reinterpret_cast<double*>(smem_raw);             // [kLST][kLCH*Sd]
    double* ws = tiles + (size_t)kLST * stage_doubles;               // [kLCH]
    double* rs = ws + kLCH;                                          // [kLCH]
    double* beta_s = rs + kLCH;                                      // [NT*8] internal column order
    uint64_t* bars = reinterpret_cast<uint64_t*>(beta_s + NT * 8);   // [kLST]

    for (int c = tid; c < NT * 8; c += blockDim.x) {
        double bv = 0.0;
        if (FAMILY == DLSA_FAMILY_LOGISTIC) {
            if (c < px) bv = a.beta[(size_t)k * p + c + a.icpt];
            else if (c == px && a.icpt) bv = a.beta[(size_t)k * p];
        }
        beta_s[c] = bv;
    }
    // zero the ring once: tail chunks leave stale (finite) rows behind, never uninitialised ones
    for (int e = tid; e < kLST * stage_doubles; e += blockDim.x) tiles[e] = 0.0;
    if (tid == 0) {
        for (int s = 0; s < kLST; s++) mbar_init(smem_u32(&bars[s]), a.aligned ? 1u : (uint32_t)(kLW * 32));
        fence_mbar_init();
    }
    fence_proxy_async();
    __syncthreads();

    // ---- this warp's segments: id = (tg*kNSEG + slot)*kLW + warp over the row-major segment list ---------
    int seg_i[kNSEG], seg_j0[kNSEG], seg_len[kNSEG];
#pragma unroll
    for (int s = 0; s < kNSEG; s++) {
        const int id = (tg * kNSEG + s) * kLW + warp;
        int i = 0, cum = 0;
        while (i < NT && cum + segs_in_row(i) <= id) { cum += segs_in_row(i); i++; }
        if (i < NT) {
            seg_i[s] = i;
            seg_j0[s] = (id - cum) * kSEG;
            const int rem = i + 1 - seg_j0[s];
            seg_len[s] = rem < kSEG ? rem : kSEG;
        } else {
            seg_i[s] = 0; seg_j0[s] = 0; seg_len[s] = 0;
        }
    }

    const double* Xp = a.X + row_lo * a.ldx;
    const double* yp = a.y + row_lo;
    const uint32_t row_bytes = (uint32_t)px * 8u;

    auto issue = [&](int ci) {
        if (ci >= my_n) return;
        const int st = ci % kLST;
        const int64_t r0 = (c_lo + ci) * kLCH;
        const int valid = (n_k - r0 < kLCH) ? (int)(n_k - r0) : kLCH;
        const uint32_t bar = smem_u32(&bars[st]);
        double* dst = tiles + (size_t)st * stage_doubles;
        if (a.aligned) {  // dense rows (Sd == px): ONE bulk copy per chunk
            if (tid == 0) {
                mbar_arrive_expect_tx(bar, row_bytes * (uint32_t)valid);
                bulk_g2s(smem_u32(dst), Xp + r0 * a.ldx, row_bytes * (uint32_t)valid, bar);
            }
        } else {
            for (int e = tid; e < valid * px; e += blockDim.x) {
                const int row = e / px, col = e - row * px;
                cp_async8(smem_u32(dst + (size_t)row * Sd + col), Xp + (r0 + row) * a.ldx + col);
            }
            cp_async_mbar_arrive_noinc(bar);
        }
    };

    double C[kNSEG][kSEG][2];
#pragma unroll
    for (int s = 0; s < kNSEG; s++)
#pragma unroll
        for (int e = 0; e < kSEG; e++) C[s][e][0] = C[s][e][1] = 0.0;
    double gacc[2] = {0.0, 0.0};  // columns tid and tid + 256 (tile group 0 only)
    double llacc = 0.0;

    // k-step row map: rows {base + rstep*q}; rstep = 2 when Sd % 4 == 2 keeps DMMA fragment loads conflict-free
    const int rstep = a.rstep;

    for (int s = 0; s < kLST; s++) issue(s);

    for (int ci = 0; ci < my_n; ci++) {
        const int st = ci % kLST;
        const int64_t r0 = (c_lo + ci) * kLCH;
        const int valid = (n_k - r0 < kLCH) ? (int)(n_k - r0) : kLCH;
        mbar_wait(smem_u32(&bars[st]), (uint32_t)((ci / kLST) & 1));
        const double* T = tiles + (size_t)st * stage_doubles;

        // ---- per-row stage: warp w owns rows 2w, 2w+1 ----------------------------------------------------
#pragma unroll
        for (int rr = 0; rr < 2; rr++) {
            const int row = 2 * warp + rr;
            double wv = 0.0, rv = 0.0;
            const bool rowvalid = row < valid;
            const double yv = rowvalid ? __ldg(yp + r0 + row) : 0.0;
            if (FAMILY == DLSA_FAMILY_LOGISTIC) {
                double s = 0.0;
                for (int j = lane; j < px; j += 32) s = fma(T[(size_t)row * Sd + j], beta_s[j], s);
#pragma unroll
                for (int m = 16; m >= 1; m >>= 1) s += shfl_xor_f64(s, m);
                if (a.icpt) s += beta_s[px];
                const double pr = sigmoid_f64(s);
                wv = pr * (1.0 - pr);
                rv = yv - pr;
                if (a.want_ll && rowvalid && lane == 0 && tg == 0) llacc += yv * s - log1pexp_f64(s);
            } else {
                wv = 1.0;
                rv = yv;
                if (rowvalid && lane == 0 && tg == 0) llacc = fma(yv, yv, llacc);
            }
            if (lane == 0) {
                ws[row] = rowvalid ? wv : 0.0;
                rs[row] = rowvalid ? rv : 0.0;
            }
        }
        __syncthreads();

        // ---- g += X'r (tile group 0): thread <-> column -------------------------------------------------
        if (tg == 0) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int c = tid + h * kLW * 32;
                if (c < p) {
                    double acc = gacc[h];
                    if (c < px) {
#pragma unroll
                        for (int row = 0; row < kLCH; row++) acc = fma(rs[row], T[(size_t)row * Sd + c], acc);
                    } else {  // the synthesised intercept column
#pragma unroll
                        for (int row = 0; row < kLCH; row++) acc += rs[row];
                    }
                    gacc[h] = acc;
                }
            }
        }

        // ---- H tiles on DMMA -----------------------------------------------------------------------------
#pragma unroll
        for (int ks = 0; ks < kLCH / 4; ks++) {
            const int base = (rstep == 1) ? 4 * ks : (8 * (ks >> 1) + (ks & 1));
            const int row = base + rstep * q;
            const double wv = ws[row];
            const double* Tr = T + (size_t)row * Sd;
#pragma unroll
            for (int s = 0; s < kNSEG; s++) {
                if (seg_len[s] > 0) {  // warp-uniform
                    const int ca = 8 * seg_i[s] + a8;
                    double av = (ca < px) ? Tr[ca] : ((ca == px && a.icpt) ? 1.0 : 0.0);
                    av *= wv;
#pragma unroll
                    for (int e = 0; e < kSEG; e++) {
                        if (e < seg_len[s]) {  // warp-uniform
                            const int cb = 8 * (seg_j0[s] + e) + a8;
                            const double fv = (cb < px) ? Tr[cb] : ((cb == px && a.icpt) ? 1.0 : 0.0);
                            dmma884(C[s][e][0], C[s][e][1], av, fv);
                        }
                    }
                }
            }
        }
        __syncthreads();  // everyone is done with this stage and with ws/rs
        if (ci + kLST < my_n) {
            if (tid == 0) fence_proxy_async();
            issue(ci + kLST);
        }
    }

    // ---- emit: every tile has exactly one owner ----------------------------------------------------------
    double* oH = a.outH + (size_t)item * p * p;
    double* og = a.outg + (size_t)item * p;
#pragma unroll
    for (int s = 0; s < kNSEG; s++) {
#pragma unroll
        for (int e = 0; e < kSEG; e++) {
            if (e < seg_len[s]) {
                const int i = seg_i[s], j = seg_j0[s] + e;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int r_ = 8 * i + a8, c_ = 8 * j + 2 * q + h;
                    if (r_ < p && c_ < p && (i != j || c_ <= r_)) {
                        const int ro = a.icpt ? (r_ == px ? 0 : r_ + 1) : r_;
                        const int co = a.icpt ? (c_ == px ? 0 : c_ + 1) : c_;
                        oH[(size_t)ro * p + co] = C[s][e][h];
                        oH[(size_t)co * p + ro] = C[s][e][h];
                    }
                }
            }
        }
    }
    if (tg == 0) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int c = tid + h * kLW * 32;
            if (c < p) og[a.icpt ? (c == px ? 0 : c + 1) : c] = gacc[h];
        }
        if (a.outll != nullptr) {
            __shared__ double llred[kLW];
            if (lane == 0) llred[warp] = llacc;
            __syncthreads();
            if (tid == 0) {
                double v = 0.0;
                for (int w = 0; w < kLW; w++) v += llred[w];
                a.outll[item] = v;
            }
        }
    }
}

__global__ void slice_reduce_large_kernel(const double* __restrict__ pH, const double* __restrict__ pg,
                                          const double* __restrict__ pll, const int* __restrict__ active, int K, int S,
                                          int p, double* __restrict__ H, double* __restrict__ g,
                                          double* __restrict__ ll) {
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

// ---- host side ---------------------------------------------------------------------------------------------
struct LargeCfg {
    int NT, TG, Sd, aligned_ok, rstep;
    size_t smem;
};

static LargeCfg large_cfg(int px, int p, bool aligned) {
    LargeCfg c;
    c.NT = (p + 7) / 8;
    int nseg = 0;
    for (int i = 0; i < c.NT; i++) nseg += segs_in_row(i);
    c.TG = (nseg + kLW * kNSEG - 1) / (kLW * kNSEG);
    if (aligned) {
        c.Sd = px;  // dense rows, one bulk copy per chunk
        c.rstep = (px % 4 == 2) ? 2 : 1;
    } else {
        c.Sd = px;
        while (c.Sd % 8 != 4) c.Sd++;
        c.rstep = 1;
    }
    c.smem = ((size_t)kLST * kLCH * c.Sd + 2 * kLCH + (size_t)c.NT * 8) * 8 + kLST * 8 + 64;
    return c;
}

static int large_slices(int K, int TG, int64_t max_part_rows) {
    const int64_t chunks = (max_part_rows + kLCH - 1) / kLCH;
    int64_t s_want = (148 * 3 + (int64_t)K * TG - 1) / ((int64_t)K * TG);
    int64_t s_max = chunks / 8;
    if (s_max < 1) s_max = 1;
    int64_t s = s_want < s_max ? s_want : s_max;
    return (int)(s < 1 ? 1 : s);
}

size_t irls_large_workspace_bytes(int K, int p, int64_t max_part_rows) {
    if (p > kLargeMaxP) return 0;
    LargeCfg c = large_cfg(p, p, true);
    const int S = large_slices(K, c.TG, max_part_rows);
    if (S == 1) return 0;
    return (size_t)K * S * ((size_t)p * p + p + 1) * sizeof(double);
}

int irls_large_launch(const double* X, int64_t ldx, const double* y, const int64_t* part_off, int K, int px, int icpt,
                      int family, const double* beta, const int* active, double* g, double* H, double* loglik,
                      int64_t max_part_rows, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    const int p = px + (icpt ? 1 : 0);
    DLSA_REQUIRE(p <= kLargeMaxP, DLSA_E_UNSUPPORTED, "irls pass: p=%d > %d not supported", p, kLargeMaxP);
    const bool aligned = ((reinterpret_cast<uintptr_t>(X) & 15) == 0 && (ldx % 2) == 0 && (px % 2) == 0);
    LargeCfg c = large_cfg(px, p, aligned);
    const int S = large_slices(K, c.TG, max_part_rows);
    LargeArgs a;
    a.X = X; a.ldx = ldx; a.y = y; a.part_off = part_off; a.beta = beta; a.active = active;
    a.K = K; a.px = px; a.p = p; a.icpt = icpt ? 1 : 0; a.S = S; a.Sd = c.Sd; a.aligned = aligned ? 1 : 0;
    a.want_ll = loglik != nullptr; a.NT = c.NT; a.TG = c.TG; a.rstep = c.rstep;
    // dense single-copy staging needs the chunk's rows contiguous in global memory
    if (aligned && ldx != px) {
        a.aligned = 0;
        LargeCfg c2 = large_cfg(px, p, false);
        a.Sd = c2.Sd; a.rstep = c2.rstep; c.smem = c2.smem;
    }
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
    dim3 grid(K * S, c.TG);
    if (family == DLSA_FAMILY_LOGISTIC) {
        DLSA_CUDA(cudaFuncSetAttribute(irls_large_kernel<DLSA_FAMILY_LOGISTIC>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
        irls_large_kernel<DLSA_FAMILY_LOGISTIC><<<grid, kLW * 32, c.smem, st>>>(a);
    } else {
        DLSA_CUDA(cudaFuncSetAttribute(irls_large_kernel<DLSA_FAMILY_LINEAR>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
        irls_large_kernel<DLSA_FAMILY_LINEAR><<<grid, kLW * 32, c.smem, st>>>(a);
    }
    count_launch();
    int rc = check_launch("irls_large_kernel");
    if (rc) return rc;
    if (S > 1) {
        dim3 grid2((p * p + p + 1 + 255) / 256, K);
        slice_reduce_large_kernel<<<grid2, 256, 0, st>>>(pH, pg, loglik ? pll : nullptr, active, K, S, p, H, g, loglik);
        count_launch();
        return check_launch("slice_reduce_large_kernel");
    }
    return 0;
}

}  // namespace dlsa
