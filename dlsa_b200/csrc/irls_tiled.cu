// K1 for 56 < p <= 512: fused Newton/IRLS pass with the Gram triangle tiled over the warps of a CTA (and, when the
// triangle does not fit one CTA's registers, over TG sibling CTAs that stream the same rows).
//
// Same arithmetic as irls_small.cu (dlsa/R/wlse_func.R:47-54, dlsa/models.py:114,130).  Design:
//   * The lower triangle of H is cut into BLOCKS of 2 tile rows x 4 tile columns of 8x8 DMMA tiles (<= 8 tiles,
//     32 accumulator registers per lane).  A block needs 2 weighted A fragments and 4 B fragments per 4-row k-step
//     (0.75 fragment loads per DMMA).  The host builds the block list, sorts it by valid-tile count and deals it
//     to (sibling CTA, sub-partition, warp, slot) so that every sub-partition's FP64 pipe carries the same number of
//     DMMAs (the pipe is per sub-partition; an unbalanced deal idles it).  Every tile has exactly one owner: no
//     cross-warp reduction, bitwise reproducible.
//   * CTA = 16 warps, one CTA per SM.  12 MMA warps (setmaxnreg 144) own up to NBW blocks each and do nothing but
//     LDS + DMUL + DMMA, in a straight-line body specialised on (blocks owned, any partial block).  3 row warps
//     take the 8-row half-chunks round-robin: eta = x'beta (lane = column stripe, 4 rows in flight, transposed
//     butterfly reduction), the sigmoid, w and r ahead of the MMA warps, and g = X'r.  1 producer warp keeps a ring
//     of 16-row chunks full with one TMA bulk copy per row (UBLKCP; rows land with a pitch == 4 mod 16 doubles so
//     that fragment loads are bank-conflict free), gated by full / wr / empty mbarriers per stage.
//   * Sibling CTAs (tile group tg = blockIdx.x % TG) are adjacent in the grid, run at the same time and read the
//     same rows: X comes from HBM once and from L2 TG-1 times.
#include <algorithm>
#include <type_traits>
#include <vector>

#include "common.cuh"

namespace dlsa {

constexpr int kTMma = 12;                       // MMA warps
constexpr int kTRow = 3;                        // row warps
constexpr int kTWarps = kTMma + kTRow + 1;      // + producer = 16; setmaxnreg acts on whole warpgroups (4 aligned warps)
constexpr int kTThreads = kTWarps * 32;         // 512 -> 128 registers per thread at launch
constexpr int kTCH = 16;                        // rows per chunk
constexpr int kTMaxTG = 16;
constexpr int kTMaxNBW = 3;
constexpr int kTRegsMma = 144;
constexpr int kTRegsAux = 80;
static_assert(kTMma * kTRegsMma + (kTWarps - kTMma) * kTRegsAux == kTWarps * 128, "register pool");
static_assert(kTMma % 4 == 0 && kTWarps % 4 == 0, "setmaxnreg needs whole warpgroups");
constexpr int kTiledMaxP = 512;

struct TiledArgs {
    const double* X;
    int64_t ldx;
    const double* y;
    const int64_t* part_off;
    const double* beta;
    const int* active;
    double* outH;   // [K*S][p*p]
    double* outg;   // [K*S][p]
    double* outll;  // [K*S] or nullptr
    int K, px, p, icpt, S, Sd, aligned, want_ll, NT, TG, stages, copy_cols;
    // block of MMA warp w, slot b of sibling tg: bits 0-7 valid-tile mask (bit 4*ii+jj), 8-15 row pair m, 16-23 column
    // group J; 0 = empty slot
    uint32_t sched[kTMaxTG * kTMma * kTMaxNBW];
};

template <int N>
__device__ __forceinline__ void t_setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void t_setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ void t_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

template <int FAMILY, int NBW, int JMAX>
__global__ void __launch_bounds__(kTThreads, 1) irls_tiled_kernel(const __grid_constant__ TiledArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int a8 = lane >> 2, q = lane & 3;
    const int item = blockIdx.x / a.TG, tg = blockIdx.x % a.TG;  // siblings adjacent: co-scheduled, share L2 lines
    const int k = item / a.S, slice = item % a.S;
    if (a.active != nullptr && a.active[k] == 0) return;

    const int p = a.p, px = a.px, Sd = a.Sd, ST = a.stages;
    const int64_t row_lo = a.part_off[k];
    const int64_t n_k = a.part_off[k + 1] - row_lo;
    const int64_t nchunks = (n_k + kTCH - 1) / kTCH;
    const int64_t cps = (nchunks + a.S - 1) / a.S;
    const int64_t c_lo = (int64_t)slice * cps;
    const int64_t c_hi = (c_lo + cps < nchunks) ? (c_lo + cps) : nchunks;
    const int my_n = (c_hi > c_lo) ? (int)(c_hi - c_lo) : 0;

    // ---- shared memory: tiles [ST][CH*Sd] | ws [ST][CH] | rs [ST][CH] | beta_s [Sd] | gred [2][JMAX*32+2] | bars
    const int stage_doubles = kTCH * Sd;
    double* tiles = reinterpret_cast<double*>(smem_raw);
    double* ws_all = tiles + (size_t)ST * stage_doubles;
    double* rs_all = ws_all + ST * kTCH;
    double* beta_s = rs_all + ST * kTCH;
    double* gred = beta_s + Sd;
    uint64_t* bars = reinterpret_cast<uint64_t*>(gred + 2 * (JMAX * 32 + 2));  // [ST][3]: full, wr, empty

    for (int c = tid; c < Sd; c += kTThreads) {
        double bv = 0.0;
        if (FAMILY == DLSA_FAMILY_LOGISTIC) {
            if (c < px) bv = a.beta[(size_t)k * p + c + a.icpt];
            else if (c == px && a.icpt) bv = a.beta[(size_t)k * p];
        }
        beta_s[c] = bv;
    }
    // zero the ring once (tail chunks leave stale finite rows behind, never uninitialised ones; columns >= px are
    // never written by the copies of an even-px matrix), then plant the intercept column
    for (int e = tid; e < ST * stage_doubles; e += kTThreads) tiles[e] = 0.0;
    __syncthreads();
    if (a.icpt)
        for (int r = tid; r < ST * kTCH; r += kTThreads) tiles[(size_t)r * Sd + px] = 1.0;
    if (tid < ST) {
        uint64_t* b = bars + tid * 3;
        mbar_init(smem_u32(&b[0]), a.aligned ? 1u : 32u);
        mbar_init(smem_u32(&b[1]), 2u);                       // the two half-chunks
        mbar_init(smem_u32(&b[2]), (uint32_t)(kTMma + 2));
        fence_mbar_init();
    }
    fence_proxy_async();
    __syncthreads();

    const double* Xp = a.X + row_lo * a.ldx;
    const double* yp = a.y + row_lo;

    if (warp == kTWarps - 1) {
        // =========================== producer warp: one bulk copy per row ==================================
        t_setmaxnreg_dec<kTRegsAux>();
        const uint32_t row_bytes = (uint32_t)a.copy_cols * 8u;
        int st = 0;
        uint32_t par = 1;  // parity of the empty phase a refill waits for; flips to 0 at the first wrap
        for (int ci = 0; ci < my_n; ci++) {
            if (ci >= ST) {
                mbar_wait(smem_u32(&bars[st * 3 + 2]), par);
                fence_proxy_async();
            }
            const int64_t r0 = (c_lo + ci) * kTCH;
            const int valid = (n_k - r0 < kTCH) ? (int)(n_k - r0) : kTCH;
            const uint32_t bar = smem_u32(&bars[st * 3 + 0]);
            double* dst = tiles + (size_t)st * stage_doubles;
            if (a.aligned) {
                if (lane == 0) mbar_arrive_expect_tx(bar, row_bytes * (uint32_t)valid);
                __syncwarp();
                if (lane < valid)
                    bulk_g2s(smem_u32(dst + (size_t)lane * Sd), Xp + (r0 + lane) * a.ldx, row_bytes, bar);
            } else {
                for (int e = lane; e < valid * px; e += 32) {
                    const int rw_ = e / px, col = e - rw_ * px;
                    cp_async8(smem_u32(dst + (size_t)rw_ * Sd + col), Xp + (r0 + rw_) * a.ldx + col);
                }
                cp_async_mbar_arrive_noinc(bar);
            }
            if (++st == ST) { st = 0; par ^= 1u; }
        }
    } else if (warp >= kTMma) {
        // =========================== row warps: eta, sigmoid, w, r (+ g, log-likelihood on sibling 0) ========
        // The 3 row warps take the half-chunks (8 rows) of the stream round-robin; a half-chunk is two 4-row batches
        // whose sigmoid chains interleave.
        t_setmaxnreg_dec<kTRegsAux>();
        const int rw = warp - kTMma;
        const int bit4 = (lane >> 4) & 1, bit3 = (lane >> 3) & 1;
        const int rr = 2 * bit4 + bit3;        // row of a 4-row batch this lane finishes
        const bool lead = (lane & 7) == 0;
        const bool fix_pad = a.copy_cols > px;  // odd px: the copy drags one foreign column along
        const double pad_val = a.icpt ? 1.0 : 0.0;
        const double b0 = a.icpt ? beta_s[px] : 0.0;
        double gacc[JMAX];
#pragma unroll
        for (int j = 0; j < JMAX; j++) gacc[j] = 0.0;
        double g_icpt = 0.0, llacc = 0.0;
        const int n_half = 2 * my_n;
        auto load_y = [&](int h, int bt) -> double {
            const int64_t r = (c_lo + (h >> 1)) * kTCH + (h & 1) * 8 + bt * 4 + rr;
            return (h < n_half && r < n_k) ? __ldg(yp + r) : 0.0;
        };
        double y_next[2] = {load_y(rw, 0), load_y(rw, 1)};
        int st = 0, cur = 0;
        uint32_t par = 0;

        for (int h = rw; h < n_half; h += kTRow) {
            const int ci = h >> 1, hb = (h & 1) * 8;
            while (cur < ci) { cur++; if (++st == ST) { st = 0; par ^= 1u; } }
            const int64_t r0 = (c_lo + ci) * kTCH;
            const int valid = (n_k - r0 < kTCH) ? (int)(n_k - r0) : kTCH;
            const double y_cur[2] = {y_next[0], y_next[1]};
            y_next[0] = load_y(h + kTRow, 0);
            y_next[1] = load_y(h + kTRow, 1);
            double* T = tiles + (size_t)st * stage_doubles;
            mbar_wait(smem_u32(&bars[st * 3 + 0]), par);
            double rv_b[2];
#pragma unroll
            for (int bt = 0; bt < 2; bt++) {
                const int rb = hb + bt * 4;
                const int row = rb + rr;
                const bool rowvalid = row < valid;
                const double yv = y_cur[bt];
                double wv, rv;
                if (FAMILY == DLSA_FAMILY_LOGISTIC) {
                    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                    const double* Tb = T + (size_t)rb * Sd + lane;
#pragma unroll
                    for (int j = 0; j < JMAX; j++) {
                        const int c = lane + 32 * j;
                        if (c < px) {
                            const double bv = beta_s[c];
                            s0 = fma(Tb[32 * j], bv, s0);
                            s1 = fma(Tb[32 * j + Sd], bv, s1);
                            s2 = fma(Tb[32 * j + 2 * Sd], bv, s2);
                            s3 = fma(Tb[32 * j + 3 * Sd], bv, s3);
                        }
                    }
                    // transposed butterfly: 4 rows x 32 lanes -> every group of 8 lanes holds one row's sum
                    const double t0 = (bit4 ? s2 : s0) + shfl_xor_f64(bit4 ? s0 : s2, 16);
                    const double t1 = (bit4 ? s3 : s1) + shfl_xor_f64(bit4 ? s1 : s3, 16);
                    double u = (bit3 ? t1 : t0) + shfl_xor_f64(bit3 ? t0 : t1, 8);
                    u += shfl_xor_f64(u, 4);
                    u += shfl_xor_f64(u, 2);
                    u += shfl_xor_f64(u, 1);
                    const double eta = u + b0;
                    const double e = exp_nonpos(-fabs(eta));
                    const double pr = div_pos((eta >= 0.0) ? 1.0 : e, 1.0 + e);
                    wv = pr * (1.0 - pr);
                    rv = yv - pr;
                    if (a.want_ll && rowvalid && lead) llacc += yv * eta - (fmax(eta, 0.0) + log1p_01(e));
                } else {
                    wv = 1.0;
                    rv = yv;
                    if (rowvalid && lead) llacc = fma(yv, yv, llacc);
                }
                wv = rowvalid ? wv : 0.0;
                rv = rowvalid ? rv : 0.0;
                rv_b[bt] = rv;
                if (lead) {
                    ws_all[st * kTCH + row] = wv;
                    rs_all[st * kTCH + row] = rv;
                    if (fix_pad) T[(size_t)row * Sd + px] = pad_val;
                }
            }
            if (fix_pad) fence_proxy_async();  // generic-proxy writes into a stage the TMA will overwrite later
            __syncwarp();
            if (lane == 0) t_mbar_arrive(smem_u32(&bars[st * 3 + 1]));  // w/r of these 8 rows are ready
            if (tg == 0) {
#pragma unroll
                for (int bt = 0; bt < 2; bt++) {
                    const double* Tb = T + (size_t)(hb + bt * 4) * Sd + lane;
                    const double r0v = __shfl_sync(0xffffffffu, rv_b[bt], 0);
                    const double r1v = __shfl_sync(0xffffffffu, rv_b[bt], 8);
                    const double r2v = __shfl_sync(0xffffffffu, rv_b[bt], 16);
                    const double r3v = __shfl_sync(0xffffffffu, rv_b[bt], 24);
#pragma unroll
                    for (int j = 0; j < JMAX; j++) {
                        const int c = lane + 32 * j;
                        if (c < px) {
                            double acc = gacc[j];
                            acc = fma(r0v, Tb[32 * j], acc);
                            acc = fma(r1v, Tb[32 * j + Sd], acc);
                            acc = fma(r2v, Tb[32 * j + 2 * Sd], acc);
                            acc = fma(r3v, Tb[32 * j + 3 * Sd], acc);
                            gacc[j] = acc;
                        }
                    }
                    if (lead) g_icpt += rv_b[bt];
                }
            }
            __syncwarp();
            if (lane == 0) t_mbar_arrive(smem_u32(&bars[st * 3 + 2]));  // done reading the stage
        }
        if (tg == 0) {
            // intercept gradient and log-likelihood live on the 4 group leaders: sum them, then the 3 row warps
            g_icpt += shfl_xor_f64(g_icpt, 8);
            g_icpt += shfl_xor_f64(g_icpt, 16);
            llacc += shfl_xor_f64(llacc, 8);
            llacc += shfl_xor_f64(llacc, 16);
            if (rw > 0) {
                double* gr = gred + (size_t)(rw - 1) * (JMAX * 32 + 2);
#pragma unroll
                for (int j = 0; j < JMAX; j++) gr[lane + 32 * j] = gacc[j];
                if (lane == 0) { gr[JMAX * 32] = g_icpt; gr[JMAX * 32 + 1] = llacc; }
            }
            asm volatile("bar.sync 1, 96;" ::: "memory");
            if (rw == 0) {
                double* og = a.outg + (size_t)item * p;
                const double* g1 = gred;
                const double* g2 = gred + (JMAX * 32 + 2);
#pragma unroll
                for (int j = 0; j < JMAX; j++) {
                    const int c = lane + 32 * j;
                    if (c < px) og[c + a.icpt] = (gacc[j] + g1[c]) + g2[c];
                }
                if (lane == 0) {
                    if (a.icpt) og[0] = (g_icpt + g1[JMAX * 32]) + g2[JMAX * 32];
                    if (a.outll != nullptr) a.outll[item] = (llacc + g1[JMAX * 32 + 1]) + g2[JMAX * 32 + 1];
                }
            }
        }
    } else {
        // =========================== MMA warps: H blocks on DMMA ==========================================
        t_setmaxnreg_inc<kTRegsMma>();
        uint32_t bmask[NBW];
        int offA[NBW], offB[NBW];
        int nb = 0;
        bool pred = false;
#pragma unroll
        for (int b = 0; b < NBW; b++) {
            const uint32_t e = a.sched[(tg * kTMma + warp) * kTMaxNBW + b];
            bmask[b] = e & 0xffu;
            offA[b] = 16 * (int)((e >> 8) & 0xffu) + a8;    // column of tile row 2m (tile row 2m+1 is +8)
            offB[b] = 32 * (int)((e >> 16) & 0xffu) + a8;   // column of tile column 4J (+8 per tile)
            if (bmask[b]) nb = b + 1;                       // slots fill in order
            if (bmask[b] != 0u && bmask[b] != 0xffu) pred = true;
        }
        double C[NBW][2][4][2];
#pragma unroll
        for (int b = 0; b < NBW; b++)
#pragma unroll
            for (int ii = 0; ii < 2; ii++)
#pragma unroll
                for (int jj = 0; jj < 4; jj++) C[b][ii][jj][0] = C[b][ii][jj][1] = 0.0;

        // one straight-line body per (number of blocks, any partial block): loads are unconditional (every column a
        // block can name exists in shared memory), so ptxas can hoist them across the DMMAs
        auto run = [&](auto nb_c, auto pred_c) {
            constexpr int NB = decltype(nb_c)::value;
            constexpr bool PRED = decltype(pred_c)::value;
            int st = 0;
            uint32_t par = 0;
            for (int ci = 0; ci < my_n; ci++) {
                mbar_wait(smem_u32(&bars[st * 3 + 0]), par);  // rows landed (async-proxy writes visible here)
                mbar_wait(smem_u32(&bars[st * 3 + 1]), par);  // w written, pad column repaired
                const double* T = tiles + (size_t)st * stage_doubles + (size_t)q * Sd;
                const double* ws = ws_all + st * kTCH + q;
#pragma unroll
                for (int ks = 0; ks < kTCH / 4; ks++) {
                    const double wv = ws[4 * ks];
                    const double* Tr = T + (size_t)(4 * ks) * Sd;  // row 4ks + q
#pragma unroll
                    for (int b = 0; b < NB; b++) {
                        const double a0 = Tr[offA[b]] * wv, a1 = Tr[offA[b] + 8] * wv;
                        double f[4];
#pragma unroll
                        for (int jj = 0; jj < 4; jj++) f[jj] = Tr[offB[b] + 8 * jj];
#pragma unroll
                        for (int jj = 0; jj < 4; jj++) {
                            if (!PRED || (bmask[b] & (1u << jj))) dmma884(C[b][0][jj][0], C[b][0][jj][1], a0, f[jj]);
                            if (!PRED || (bmask[b] & (16u << jj))) dmma884(C[b][1][jj][0], C[b][1][jj][1], a1, f[jj]);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) t_mbar_arrive(smem_u32(&bars[st * 3 + 2]));  // done with the stage
                if (++st == ST) { st = 0; par ^= 1u; }
            }
        };
        using std::integral_constant;
        if (nb == 0) run(integral_constant<int, 0>{}, integral_constant<bool, false>{});
        else if (nb == 1 && !pred) run(integral_constant<int, 1>{}, integral_constant<bool, false>{});
        else if (nb == 1) run(integral_constant<int, 1>{}, integral_constant<bool, true>{});
        else if (nb == 2 && !pred) run(integral_constant<int, 2>{}, integral_constant<bool, false>{});
        else if (nb == 2) run(integral_constant<int, 2>{}, integral_constant<bool, true>{});
        else if constexpr (NBW >= 3) {
            if (!pred) run(integral_constant<int, NBW>{}, integral_constant<bool, false>{});
            else run(integral_constant<int, NBW>{}, integral_constant<bool, true>{});
        }

        // ---- emit: every tile has exactly one owner --------------------------------------------------------
        double* oH = a.outH + (size_t)item * p * p;
#pragma unroll
        for (int b = 0; b < NBW; b++) {
            const uint32_t e = a.sched[(tg * kTMma + warp) * kTMaxNBW + b];
            const int m = (int)((e >> 8) & 0xffu), J = (int)((e >> 16) & 0xffu);
#pragma unroll
            for (int ii = 0; ii < 2; ii++)
#pragma unroll
                for (int jj = 0; jj < 4; jj++) {
                    if (bmask[b] & (1u << (4 * ii + jj))) {
                        const int i = 2 * m + ii, j = 4 * J + jj;
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const int r_ = 8 * i + a8, c_ = 8 * j + 2 * q + h;
                            if (r_ < p && c_ < p && (i != j || c_ <= r_)) {
                                const int ro = a.icpt ? (r_ == px ? 0 : r_ + 1) : r_;
                                const int co = a.icpt ? (c_ == px ? 0 : c_ + 1) : c_;
                                oH[(size_t)ro * p + co] = C[b][ii][jj][h];
                                oH[(size_t)co * p + ro] = C[b][ii][jj][h];
                            }
                        }
                    }
                }
        }
    }
}

__global__ void slice_reduce_tiled_kernel(const double* __restrict__ pH, const double* __restrict__ pg,
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
struct TiledCfg {
    int NT, TG, NBW, JMAX, Sd, stages;
    size_t smem;
    std::vector<uint32_t> sched;
};

static inline int popcount8(uint32_t m) { return __builtin_popcount(m & 0xffu); }

static TiledCfg tiled_cfg(int px, int p) {
    TiledCfg c;
    c.NT = (p + 7) / 8;
    const int MR = (c.NT + 1) / 2;
    struct Blk { int m, J; uint32_t mask; };
    std::vector<Blk> blocks;
    for (int m = 0; m < MR; m++)
        for (int J = 0; 4 * J <= 2 * m + 1 && 4 * J < c.NT; J++) {
            uint32_t mask = 0;
            for (int ii = 0; ii < 2; ii++)
                for (int jj = 0; jj < 4; jj++) {
                    const int i = 2 * m + ii, j = 4 * J + jj;
                    if (i < c.NT && j <= i) mask |= 1u << (4 * ii + jj);
                }
            if (mask) blocks.push_back({m, J, mask});
        }
    const int nb = (int)blocks.size();
    const int tg2 = (nb + kTMma * 2 - 1) / (kTMma * 2), tg3 = (nb + kTMma * 3 - 1) / (kTMma * 3);
    c.NBW = (tg3 < tg2) ? 3 : 2;
    c.TG = (c.NBW == 3) ? tg3 : tg2;
    // deal the blocks, heaviest first, to the least-loaded (sibling, sub-partition) bin with a free slot
    std::stable_sort(blocks.begin(), blocks.end(),
                     [](const Blk& x, const Blk& y) { return popcount8(x.mask) > popcount8(y.mask); });
    const int nbins = c.TG * 4, cap = 3 * c.NBW;
    std::vector<int> load(nbins, 0), cnt(nbins, 0);
    c.sched.assign((size_t)kTMaxTG * kTMma * kTMaxNBW, 0u);
    for (const Blk& b : blocks) {
        int best = -1;
        for (int i = 0; i < nbins; i++)
            if (cnt[i] < cap && (best < 0 || load[i] < load[best])) best = i;
        const int tg = best / 4, smsp = best % 4;
        const int idx = cnt[best]++;
        load[best] += popcount8(b.mask);
        const int warp = smsp + 4 * (idx % 3), slot = idx / 3;   // MMA warp w runs on sub-partition w % 4
        c.sched[((size_t)tg * kTMma + warp) * kTMaxNBW + slot] = b.mask | ((uint32_t)b.m << 8) | ((uint32_t)b.J << 16);
    }
    c.JMAX = px <= 128 ? 4 : (px <= 256 ? 8 : 16);
    int cols = 8 * 4 * ((c.NT + 3) / 4);   // every fragment column a block can touch exists in shared memory
    if (cols < px + 1) cols = px + 1;
    c.Sd = cols;
    while (c.Sd % 16 != 4) c.Sd++;
    const size_t fixed = ((size_t)c.Sd + 2 * ((size_t)c.JMAX * 32 + 2)) * 8 + 256;
    const size_t per_stage = ((size_t)kTCH * c.Sd + 2 * kTCH) * 8 + 3 * 8;
    int st = (int)((220 * 1024 - fixed) / per_stage);
    c.stages = st > 8 ? 8 : st;
    c.smem = fixed + per_stage * c.stages;
    return c;
}

static int tiled_slices(int K, int TG, int64_t max_part_rows) {
    const int64_t chunks = (max_part_rows + kTCH - 1) / kTCH;
    int64_t s_cap = chunks / 8;
    if (s_cap < 1) s_cap = 1;
    if (s_cap > 32) s_cap = 32;
    const int64_t sms = 148;
    double best_eff = -1.0;
    int best_s = 1;
    for (int64_t s = 1; s <= s_cap; s++) {
        const int64_t u = (int64_t)K * TG * s;
        const double eff = (double)u / (double)(sms * ((u + sms - 1) / sms));
        if (eff > best_eff + 0.01) { best_eff = eff; best_s = (int)s; }  // a larger S must buy > 1 % to be taken
    }
    return best_s;
}

size_t irls_tiled_workspace_bytes(int K, int p, int64_t max_part_rows) {
    if (p > kTiledMaxP) return 0;
    TiledCfg c = tiled_cfg(p, p);
    const int S = tiled_slices(K, c.TG, max_part_rows);
    if (S == 1) return 0;
    return (size_t)K * S * ((size_t)p * p + p + 1) * sizeof(double);
}

template <int FAMILY, int NBW>
static int tiled_launch_j(const TiledArgs& a, int jmax, size_t smem, int grid, cudaStream_t st) {
#define DLSA_TJ(J_)                                                                                                  \
    {                                                                                                                \
        auto kern = irls_tiled_kernel<FAMILY, NBW, J_>;                                                              \
        DLSA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));               \
        kern<<<grid, kTThreads, smem, st>>>(a);                                                                      \
    }
    if (jmax == 4) DLSA_TJ(4)
    else if (jmax == 8) DLSA_TJ(8)
    else DLSA_TJ(16)
#undef DLSA_TJ
    count_launch();
    return check_launch("irls_tiled_kernel");
}

int irls_tiled_launch(const double* X, int64_t ldx, const double* y, const int64_t* part_off, int K, int px, int icpt,
                      int family, const double* beta, const int* active, double* g, double* H, double* loglik,
                      int64_t max_part_rows, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    const int p = px + (icpt ? 1 : 0);
    DLSA_REQUIRE(p <= kTiledMaxP, DLSA_E_UNSUPPORTED, "irls pass: p=%d > %d not supported", p, kTiledMaxP);
    TiledCfg c = tiled_cfg(px, p);
    DLSA_REQUIRE(c.TG <= kTMaxTG && c.stages >= 2, DLSA_E_UNSUPPORTED, "irls pass: p=%d does not fit (TG=%d, stages=%d)",
                 p, c.TG, c.stages);
    const int S = tiled_slices(K, c.TG, max_part_rows);
    TiledArgs a;
    a.X = X; a.ldx = ldx; a.y = y; a.part_off = part_off; a.beta = beta; a.active = active;
    a.K = K; a.px = px; a.p = p; a.icpt = icpt ? 1 : 0; a.S = S; a.Sd = c.Sd;
    a.want_ll = loglik != nullptr; a.NT = c.NT; a.TG = c.TG; a.stages = c.stages;
    // one 16-byte-aligned bulk copy per row: even row pitch; an odd px drags the next (padding) column along
    a.aligned = ((reinterpret_cast<uintptr_t>(X) & 15) == 0 && (ldx % 2) == 0) ? 1 : 0;
    a.copy_cols = (a.aligned && (px % 2)) ? px + 1 : px;
    for (size_t i = 0; i < c.sched.size(); i++) a.sched[i] = c.sched[i];
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
    const int grid = K * S * c.TG;
    int rc;
    if (family == DLSA_FAMILY_LOGISTIC)
        rc = (c.NBW == 3) ? tiled_launch_j<DLSA_FAMILY_LOGISTIC, 3>(a, c.JMAX, c.smem, grid, st)
                          : tiled_launch_j<DLSA_FAMILY_LOGISTIC, 2>(a, c.JMAX, c.smem, grid, st);
    else
        rc = (c.NBW == 3) ? tiled_launch_j<DLSA_FAMILY_LINEAR, 3>(a, c.JMAX, c.smem, grid, st)
                          : tiled_launch_j<DLSA_FAMILY_LINEAR, 2>(a, c.JMAX, c.smem, grid, st);
    if (rc) return rc;
    if (S > 1) {
        dim3 grid2((p * p + p + 1 + 255) / 256, K);
        slice_reduce_tiled_kernel<<<grid2, 256, 0, st>>>(pH, pg, loglik ? pll : nullptr, active, K, S, p, H, g, loglik);
        count_launch();
        return check_launch("slice_reduce_tiled_kernel");
    }
    return 0;
}

}  // namespace dlsa
