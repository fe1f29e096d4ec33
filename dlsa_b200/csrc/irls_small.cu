// K1 (p <= 56): fused Newton/IRLS pass, warp-specialised.
//
// Reference arithmetic replaced (per partition): eta = X beta, prob = sigmoid(eta), w = prob(1-prob),
// g = X'(y-prob), H = X' diag(w) X  — dlsa/R/wlse_func.R:47-54, dlsa/models.py:114,130.
//
// Design.  FP64 DMMA and DFMA share one 64 FMA/clk/SM pipe on B200, so the pass is bound by
// that pipe; the kernel is built so the pipe never waits for anything else and executes as few non-DMMA FP64
// instructions as possible:
//   * CTA = 16 warps.  Warps 0..7 are MMA warps: each owns a private ring of 32-row chunks of X and keeps ALL
//     lower-triangle 8x8 accumulator tiles of H in registers (setmaxnreg raises them to 184 regs).
//   * The gradient rides the DMMAs that are issued anyway: in a DIAGONAL tile only the lower triangle is needed, so
//     row 0 of its A operand carries the residual r instead of w*x — that row of the accumulator is then g for the
//     tile's 8 columns (no division: the A element is r itself), and the one element it displaces, H[8t][8t], is
//     accumulated by a single DFMA for all tiles at once (lane a8 = t).
//   * Ragged columns p % 8 in 1..4 stay off the DMMA pipe (REM): their rows of H are accumulated by DFMA on the
//     fragments already in registers; their share of g shares an instruction with the ragged corner.
//   * Fragments (and w, r) of k-step s+1 — across chunk boundaries too — are loaded while k-step s is being multiplied,
//     and the w*x scaling of tile row i+2 is issued before the DMMAs of tile row i, so a DMMA never waits for a load
//     or a DMUL of its own warp.
//   * Warps 8..15 are row warps (72 regs).  Row warp r feeds MMA warp r: it issues the TMA bulk copies (4 per chunk),
//     computes eta (DFMA on a conflict-free lane grid), the sigmoid, the weights w and residuals r one chunk ahead,
//     and hands them over through shared memory + mbarriers.
//   * Chunk layout in shared memory: 4 groups of CH/4 dense rows (one bulk copy each), groups 4 doubles apart
//     modulo 16, so that the 4 rows of a DMMA k-step (one per group) and the 4x4 lane grid of the eta stage hit
//     16 distinct 8-byte banks for every even px; (w, r) pairs are stored in k-step order, one 16-byte load per lane.
//   * X is read from HBM exactly once per pass; the 8 MMA warps' accumulators are summed in a fixed order.
#include <cstdlib>

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
    int K, px, p, icpt, S, aligned, want_ll, stages;
    int poisson;  // GLM instantiation only: 0 = logistic link, 1 = log link (DLSA_FAMILY_POISSON)
};

constexpr int kMmaWarps = 8;
constexpr int kRowWarps = 8;
constexpr int kThreads = (kMmaWarps + kRowWarps) * 32;
#ifndef DLSA_K1_REGS_ROW   // tools/k1_variants.sh builds other splits side by side; 56/200 and 80/176 measured slower
#define DLSA_K1_REGS_ROW 72
#define DLSA_K1_REGS_MMA 184
#endif
constexpr int kRegsRow = DLSA_K1_REGS_ROW;
constexpr int kRegsMma = DLSA_K1_REGS_MMA;  // 72*256 + 184*256 = 65536 = 128*512 registers per SM
static_assert(kRegsRow * 256 + kRegsMma * 256 <= 65536, "register budget of the 16-warp CTA");

template <int N>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// REM > 0: p = 8(NT-1) + REM with REM <= 4: the last (ragged) tile row of H is accumulated by DFMA on the fragments
// already in registers (REM*(NT-1) + REM DFMAs per k-step) instead of NT mostly-padding DMMAs; REM == 0: all tiles on DMMA.
template <int NT, int REM, int CH, int FAMILY>
__global__ void __launch_bounds__(kThreads, 1) irls_small_kernel(const IrlsArgs a) {
    const int STAGES = a.stages;  // ring depth per MMA warp: as many CH-row stages as shared memory holds
    constexpr int ND = (REM > 0) ? NT - 1 : NT;  // tile rows on DMMA
    // Gradient on row 0 of the diagonal DMMA tiles (see the header).  Measured: p = 50 2.31 -> 2.27 ms per 2e7 rows; at
    // p = 20, where DFMAs already outnumber DMMAs, the extra selects cost more than the DFMAs they replace (4.30 -> 4.57).
    constexpr bool GSAC = ND >= 5;
    constexpr int NE = (REM > 0) ? REM : 1;
    constexpr int NTRI = ND * (ND + 1) / 2;
    // reduction-buffer layout: DMMA tiles | ragged rows x DMMA columns | ragged corner (+ its g) | displaced diagonal
    constexpr int EOFF = NTRI * 64;
    constexpr int TOFF = EOFF + REM * ND * 8;
    constexpr int DOFF = TOFF + REM * 8;
    constexpr int NOUT = DOFF + 8;
    constexpr int NRED = NOUT + 8;
    constexpr int GR = CH / 4;   // rows per group (= k-steps per chunk); even
    constexpr int NG = CH / 8;   // 8-row groups of the eta stage
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int a8 = lane >> 2, q = lane & 3;
    const int k = blockIdx.x / a.S, slice = blockIdx.x % a.S;
    if (a.active != nullptr && a.active[k] == 0) return;

    const int p = a.p, px = a.px;
    const int64_t row_lo = a.part_off[k];
    const int64_t n_k = a.part_off[k + 1] - row_lo;
    const int64_t nchunks = (n_k + CH - 1) / CH;
    const int64_t cps = (nchunks + a.S - 1) / a.S;
    const int64_t c_lo = (int64_t)slice * cps;
    const int64_t c_hi = (c_lo + cps < nchunks) ? (c_lo + cps) : nchunks;

    // ---- shared memory carve-up -------------------------------------------------------------------
    const int S = px + (px & 1);          // dense row stride (even); odd px only on the cp.async path
    const int gstride = GR * S + 4;       // doubles between row groups: == 4 (mod 16) apart for even S
    const int stage_doubles = 4 * gstride;
    double* ring = reinterpret_cast<double*>(smem_raw);                      // [8][STAGES][stage_doubles]
    size_t off = (size_t)kMmaWarps * STAGES * stage_doubles;
    double* wr_all = ring + off;                                             // [8][STAGES][CH] (w, r) pairs, k-step order
    off += (size_t)kMmaWarps * STAGES * 2 * CH;
    double* beta_s = ring + off;                                             // [NT*8]
    off += NT * 8;
    double* ll_s = ring + off;                                               // [kRowWarps]
    off += 8;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + off);                // [8][STAGES][3]: x_full, wr_full, empty

    if (threadIdx.x < NT * 8) {
        const int c = threadIdx.x;  // internal column: features 0..px-1, then the intercept
        double bv = 0.0;
        if (FAMILY == DLSA_FAMILY_LOGISTIC) {
            if (c < px) bv = a.beta[(size_t)k * p + c + a.icpt];
            else if (c == px && a.icpt) bv = a.beta[(size_t)k * p];
        }
        beta_s[c] = bv;
    }
    // zero the ring once: tail chunks leave stale finite rows behind, never uninitialised ones
    for (int e = threadIdx.x; e < kMmaWarps * STAGES * stage_doubles; e += kThreads) ring[e] = 0.0;
    if (threadIdx.x < kMmaWarps * STAGES) {
        uint64_t* b = bars + threadIdx.x * 3;
        mbar_init(smem_u32(&b[0]), a.aligned ? 1u : 32u);
        mbar_init(smem_u32(&b[1]), 1u);
        mbar_init(smem_u32(&b[2]), 1u);
        fence_mbar_init();
    }
    fence_proxy_async();
    __syncthreads();

    if (warp >= kMmaWarps) {
        // =========================== row warp: TMA producer + eta / sigmoid / weights ==================
        setmaxnreg_dec<kRegsRow>();
        const int rw = warp - kMmaWarps;  // row warp index 0..7; serves MMA warp rw (same sub-partition)
        const double* Xp = a.X + row_lo * a.ldx;
        const double* yp = a.y + row_lo;
        const int64_t first = c_lo + rw;
        const int n_iter = (first < c_hi) ? (int)((c_hi - first + kMmaWarps - 1) / kMmaWarps) : 0;

        auto issue = [&](int ci, int st) {
            const int64_t r0 = (first + (int64_t)ci * kMmaWarps) * CH;
            const int valid = (n_k - r0 < CH) ? (int)(n_k - r0) : CH;
            const uint32_t bar = smem_u32(&bars[(rw * STAGES + st) * 3 + 0]);
            double* dst = ring + (size_t)(rw * STAGES + st) * stage_doubles;
            if (a.aligned) {
                if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)valid * (uint32_t)px * 8u);
                __syncwarp();
                if (lane < 4) {
                    const int rows = valid - lane * GR < GR ? valid - lane * GR : GR;
                    if (rows > 0)
                        bulk_g2s(smem_u32(dst + (size_t)lane * gstride), Xp + (r0 + lane * GR) * a.ldx,
                                 (uint32_t)rows * (uint32_t)px * 8u, bar);
                }
            } else {
                for (int e = lane; e < valid * px; e += 32) {
                    const int row = e / px, col = e - row * px;
                    cp_async8(smem_u32(dst + (size_t)(row / GR) * gstride + (size_t)(row % GR) * S + col),
                              Xp + (r0 + row) * a.ldx + col);
                }
                cp_async_mbar_arrive_noinc(bar);
            }
        };
        // eta-stage row of this lane in 8-row group g: one row from each of the 4 copy groups per half-warp
        // row(g, a8) = GR*(a8&3) + 2g + (a8>>2)  (GR = 8: CH = 32;  GR = 4: CH = 16)
        const int erow_base = (a8 & 3) * gstride + (a8 >> 2) * S;   // + 2g*S
        const int erow_idx = GR * (a8 & 3) + (a8 >> 2);             // + 2g
        const int rho = erow_idx + 2 * q;      // row owned in the per-row stage (when q < NG): group a8&3, k-step rho % GR
        const int wr_slot = (rho % GR) * 4 + (rho / GR);            // (w, r) pairs are stored in k-step order
        auto load_y = [&](int ci) -> double {
            if (ci >= n_iter || q >= NG) return 0.0;
            const int64_t r = (first + (int64_t)ci * kMmaWarps) * CH + rho;
            return (r < n_k) ? __ldg(yp + r) : 0.0;
        };

        for (int s = 0; s < STAGES; s++)
            if (s < n_iter) issue(s, s);
        double y_next = load_y(0);
        double llacc = 0.0;

        int st = 0, stp = STAGES - 1;          // stage of chunk ci / of chunk ci - 1
        uint32_t par = 0u, parp = 1u;          // phase parity of chunk ci / of chunk ci - 1 (unused while ci == 0)
        for (int ci = 0; ci < n_iter; ci++) {
            const double y_cur = y_next;
            y_next = load_y(ci + 1);
            {
                const int64_t r0 = (first + (int64_t)ci * kMmaWarps) * CH;
                const int valid = (n_k - r0 < CH) ? (int)(n_k - r0) : CH;
                uint64_t* b = bars + (rw * STAGES + st) * 3;
                const double* T = ring + (size_t)(rw * STAGES + st) * stage_doubles;
                double2* wrp = reinterpret_cast<double2*>(wr_all + (size_t)(rw * STAGES + st) * 2 * CH);
                double wv = 0.0, rv = 0.0;
                const bool rowvalid = (q < NG) && (rho < valid);
                if (FAMILY == DLSA_FAMILY_LOGISTIC) {
                    mbar_wait(smem_u32(&b[0]), par);
                    double s[NG];
#pragma unroll
                    for (int g = 0; g < NG; g++) s[g] = 0.0;
#pragma unroll
                    for (int t = 0; t < 2 * NT; t++) {
                        const int col = 4 * t + q;
                        const bool in_x = (t < 2 * NT - 2) || (col < px);
                        const bool is_one = (t >= 2 * NT - 2) && (col == px) && a.icpt;
                        const int colc = in_x ? col : 0;
                        const double bv = beta_s[col];
#pragma unroll
                        for (int g = 0; g < NG; g++) {
                            double xv = T[erow_base + 2 * g * S + colc];
                            xv = in_x ? xv : (is_one ? 1.0 : 0.0);
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
                    if (a.poisson) {  // warp-uniform: mu = exp(eta), w = mu, r = y - mu (dlsa/R/wlse_func.R:155-168)
                        const double mu = exp_clamped(eta);
                        wv = mu;
                        rv = y_cur - mu;
                        if (a.want_ll && rowvalid) llacc += fma(y_cur, eta, -mu);
                    } else {
                        const double pr = sigmoid_f64(eta);
                        wv = pr * (1.0 - pr);
                        rv = y_cur - pr;
                        if (a.want_ll && rowvalid) llacc += y_cur * eta - log1pexp_f64(eta);
                    }
                } else {
                    wv = 1.0;
                    rv = y_cur;
                    if (rowvalid) llacc = fma(y_cur, y_cur, llacc);
                }
                if (q < NG) wrp[wr_slot] = make_double2(rowvalid ? wv : 0.0, rowvalid ? rv : 0.0);
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&b[1]));  // w/r of this chunk are ready
            }
            // refill the stage the MMA warp finished last (chunk ci-1) with chunk ci-1+STAGES
            if (ci >= 1) {
                const int cn = ci - 1 + STAGES;
                if (cn < n_iter) {
                    mbar_wait(smem_u32(&bars[(rw * STAGES + stp) * 3 + 2]), parp);
                    fence_proxy_async();
                    issue(cn, stp);
                }
            }
            stp = st; parp = par;
            if (++st == STAGES) { st = 0; par ^= 1u; }
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) llacc += shfl_xor_f64(llacc, m);
        if (lane == 0) ll_s[rw] = llacc;
        named_bar_sync(1, kThreads);  // (A) every warp is done with the ring
        named_bar_sync(2, kThreads);  // (B) partial accumulators are in shared memory
    } else {
        // =========================== MMA warp: H += (wX)'X on DMMA, g rides row 0 of the diagonal tiles ==========
        setmaxnreg_inc<kRegsMma>();
        const int m = warp;
        const int64_t first = c_lo + m;
        const int my_n = (first < c_hi) ? (int)((c_hi - first + kMmaWarps - 1) / kMmaWarps) : 0;

        double C[NTRI > 0 ? NTRI : 1][2];
        double E[NE][ND > 0 ? ND : 1];   // ragged rows x DMMA tile columns (per-lane partial sums over the 4 rows of a k-step)
        double Et[NE];                   // ragged corner; lanes a8 >= 4 of Et[0] carry g of the ragged columns
        double Dg = 0.0;                 // lane a8 = t < ND: H[8t][8t], displaced from diagonal tile t by the g row
        double gacc[ND > 0 ? ND : 1];    // !GSAC: g of the DMMA tile columns by DFMA
#pragma unroll
        for (int t = 0; t < ND; t++) gacc[t] = 0.0;
#pragma unroll
        for (int t = 0; t < NTRI; t++) C[t][0] = C[t][1] = 0.0;
#pragma unroll
        for (int e = 0; e < NE; e++) {
            Et[e] = 0.0;
#pragma unroll
            for (int t = 0; t < ND; t++) E[e][t] = 0.0;
        }
        // ragged-edge columns 8 ND + e, e < REM: shared-memory column (or the constant intercept / padding)
        int ecol[NE];
        bool e_in_x[NE];
        double e_const[NE];
#pragma unroll
        for (int e = 0; e < NE; e++) {
            const int c = 8 * ND + e;
            e_in_x[e] = c < px;
            e_const[e] = (c == px && a.icpt) ? 1.0 : 0.0;
            ecol[e] = e_in_x[e] ? c : 0;
        }

        // fragment element of this lane in k-step ks: row = GR*q + ks (group q), column 8t + a8.  The last tile column:
        // a DMMA tile (REM == 0: column 8(NT-1) + a8) or the ragged corner (REM > 0: column 8 ND + (a8 & 3), so that
        // lanes a8 and a8 + 4 hold the same column: a8 < 4 multiplies it by w*x_edge, a8 >= 4 by r).
        const int frag_base = q * gstride;  // + ks*S + 8t + a8
        const int col_last = 8 * (NT - 1) + ((REM > 0) ? (a8 & 3) : a8);
        const bool last_in_x = col_last < px;
        const double last_const = (col_last == px && a.icpt) ? 1.0 : 0.0;
        const int col_last_c = last_in_x ? col_last : 0;
        const int col_diag = (a8 < ND) ? 8 * a8 : 0;   // the displaced diagonal element of tile a8
        const bool is_g_row = (a8 == 0);
        const bool lane_lo = (a8 < 4);

        struct Frag {
            double f[NT];
            double2 wr;      // (w, r) of this lane's row
            double xe[NE];   // ragged-edge columns of this lane's row (broadcast loads)
            double xd;       // column 8*a8 of this lane's row
        };
        auto load_frag = [&](Frag& F, const double* T, const double2* wrp, int ks) {
            const double* Tr = T + ks * S;
#pragma unroll
            for (int t = 0; t < NT - 1; t++) F.f[t] = Tr[8 * t + a8];
            {
                const double v = Tr[col_last_c];
                F.f[NT - 1] = last_in_x ? v : last_const;
            }
            F.wr = wrp[ks * 4 + q];
            if (REM > 0) {
#pragma unroll
                for (int e = 0; e < NE; e++) {
                    const double v = Tr[ecol[e]];  // same address for the 8 lanes of a row: broadcast load
                    F.xe[e] = e_in_x[e] ? v : e_const[e];
                }
            }
            if (GSAC) {
                const double v = Tr[col_diag];
                F.xd = (a8 < ND) ? v : 0.0;
            } else {
                F.xd = 0.0;
            }
        };
        auto compute = [&](const Frag& F) {
            const double wv = F.wr.x, rv = F.wr.y;
            double av[ND > 0 ? ND : 1];
            double ae[NE];
            av[0] = wv * F.f[0];
            if (ND > 1) av[1] = wv * F.f[1];
            if (REM > 0) {
#pragma unroll
                for (int e = 0; e < NE; e++) ae[e] = wv * F.xe[e];
            }
#pragma unroll
            for (int i = 0; i < ND; i++) {
                if (i + 2 < ND) av[i + 2] = wv * F.f[i + 2];  // two tile rows ahead of its first DMMA
#pragma unroll
                for (int j = 0; j < i; j++)
                    dmma884(C[i * (i + 1) / 2 + j][0], C[i * (i + 1) / 2 + j][1], av[i], F.f[j]);
                const double adiag = (GSAC && is_g_row) ? rv : av[i];   // row 0 of the diagonal tile accumulates g[8i .. 8i+7]
                if (!GSAC) gacc[i] = fma(rv, F.f[i], gacc[i]);
                dmma884(C[i * (i + 1) / 2 + i][0], C[i * (i + 1) / 2 + i][1], adiag, F.f[i]);
                if (REM > 0) {
#pragma unroll
                    for (int e = 0; e < NE; e++) E[e][i] = fma(ae[e], F.f[i], E[e][i]);
                }
                if (GSAC && i == 0) Dg = fma(wv * F.xd, F.xd, Dg);
            }
            if (REM > 0) {
                Et[0] = fma(lane_lo ? ae[0] : rv, F.f[NT - 1], Et[0]);
#pragma unroll
                for (int e = 1; e < NE; e++) Et[e] = fma(ae[e], F.f[NT - 1], Et[e]);
            }
        };

        // Loads, scalings and MMAs of the GR k-steps of a chunk are unrolled into one block: ptxas hoists the fragment
        // loads of k-step s+1 above the DMMAs of k-step s on its own (explicit double buffers measured 5-9 % slower).
        {
            int st = 0;
            uint32_t par = 0u;
            for (int ci = 0; ci < my_n; ci++) {
                uint64_t* b = bars + (m * STAGES + st) * 3;
                const double* T = ring + (size_t)(m * STAGES + st) * stage_doubles + frag_base;
                const double2* wrp = reinterpret_cast<const double2*>(wr_all + (size_t)(m * STAGES + st) * 2 * CH);
                mbar_wait(smem_u32(&b[0]), par);  // X chunk landed (async-proxy writes visible to this warp)
                mbar_wait(smem_u32(&b[1]), par);  // weights / residuals written by the row warp
#pragma unroll
                for (int ks = 0; ks < GR; ks++) {
                    Frag F;
                    load_frag(F, T, wrp, ks);
                    compute(F);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&b[2]));  // stage free
                if (++st == STAGES) { st = 0; par ^= 1u; }
            }
        }
        named_bar_sync(1, kThreads);  // (A) every warp is done with the ring: reuse it as the reduction buffer
        double* red = ring + (size_t)m * NRED;
#pragma unroll
        for (int t = 0; t < NTRI; t++) {
            red[t * 64 + a8 * 8 + 2 * q] = C[t][0];
            red[t * 64 + a8 * 8 + 2 * q + 1] = C[t][1];
        }
        if (!GSAC) {  // put g and the diagonal element where the emit loop expects them
            __syncwarp();
#pragma unroll
            for (int t = 0; t < ND; t++) {
                double v = gacc[t];
                v += shfl_xor_f64(v, 1);
                v += shfl_xor_f64(v, 2);
                const int tt = t * (t + 1) / 2 + t;
                if (a8 == 0 && q == 0) Dg = red[tt * 64];          // H[8t][8t] as the DMMA computed it
                __syncwarp();
                if (q == 0) red[tt * 64 + a8] = v;                  // row 0 of the diagonal tile <- g[8t + a8]
                Dg = __shfl_sync(0xffffffffu, Dg, 0);
                if (a8 == t && q == 0) red[DOFF + t] = Dg;
                __syncwarp();
            }
        }
        if (REM > 0) {
#pragma unroll
            for (int e = 0; e < NE; e++) {
#pragma unroll
                for (int t = 0; t < ND; t++) {
                    double v = E[e][t];
                    v += shfl_xor_f64(v, 1);
                    v += shfl_xor_f64(v, 2);
                    if (q == 0) red[EOFF + (e * ND + t) * 8 + a8] = v;
                }
                double v = Et[e];
                v += shfl_xor_f64(v, 1);
                v += shfl_xor_f64(v, 2);
                if (q == 0) red[TOFF + e * 8 + a8] = v;
            }
        }
        if (GSAC) {
            double v = Dg;
            v += shfl_xor_f64(v, 1);
            v += shfl_xor_f64(v, 2);
            if (q == 0) red[DOFF + a8] = v;
        }
        named_bar_sync(2, kThreads);  // (B)
    }

    // ---- fixed-order reduction over the 8 MMA warps, then emit (all 16 warps) ------------------------
    double* oH = a.outH + (size_t)blockIdx.x * p * p;
    double* og = a.outg + (size_t)blockIdx.x * p;
    for (int e = threadIdx.x; e < NOUT + 1; e += kThreads) {
        double v = 0.0;
        if (e < NOUT) {
#pragma unroll
            for (int w = 0; w < kMmaWarps; w++) v += ring[(size_t)w * NRED + e];
        }
        int r_ = -1, c_ = -1;   // r_ >= 0: element (r_, c_) of H;  r_ == -2: element c_ of g
        if (e < EOFF) {  // DMMA tile (i, j): element (mm, nn)
            const int tt = e >> 6, mm = (e >> 3) & 7, nn = e & 7;
            int i = 0;
            while ((i + 1) * (i + 2) / 2 <= tt) i++;
            const int j = tt - i * (i + 1) / 2;
            if (i != j) { r_ = 8 * i + mm; c_ = 8 * j + nn; }
            else if (mm == 0) { r_ = -2; c_ = 8 * i + nn; }          // the g row of a diagonal tile
            else if (nn <= mm) { r_ = 8 * i + mm; c_ = 8 * i + nn; }
        } else if (e < TOFF) {  // ragged row 8 ND + ee against DMMA tile column t
            const int ee = (e - EOFF) / (ND * 8), cc = (e - EOFF) % (ND * 8);
            r_ = 8 * ND + ee; c_ = cc;
        } else if (e < DOFF) {  // ragged corner; lanes a8 >= 4 of row 0 carry g
            const int ee = (e - TOFF) >> 3, l8 = (e - TOFF) & 7;
            if (l8 <= ee && l8 < 4) { r_ = 8 * ND + ee; c_ = 8 * ND + l8; }
            else if (ee == 0 && l8 >= 4) { r_ = -2; c_ = 8 * ND + (l8 - 4); }
        } else if (e < NOUT) {  // displaced diagonal elements
            const int t = e - DOFF;
            if (t < ND) { r_ = 8 * t; c_ = 8 * t; }
        }
        if (e < NOUT) {
            if (r_ >= 0 && r_ < p && c_ < p) {
                const int ro = a.icpt ? (r_ == px ? 0 : r_ + 1) : r_;
                const int co = a.icpt ? (c_ == px ? 0 : c_ + 1) : c_;
                oH[(size_t)ro * p + co] = v;
                oH[(size_t)co * p + ro] = v;
            } else if (r_ == -2 && c_ < p) {
                const int co = a.icpt ? (c_ == px ? 0 : c_ + 1) : c_;
                og[co] = v;
            }
        } else if (a.outll != nullptr) {
            a.outll[blockIdx.x] = (((ll_s[0] + ll_s[1]) + (ll_s[2] + ll_s[3])) + ((ll_s[4] + ll_s[5]) + (ll_s[6] + ll_s[7])));
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
struct SmallCfg {
    int NT, CH, STAGES;
    size_t smem;
};

static size_t small_smem(int NT, int S, int CH, int stages) {
    const int gstride = (CH / 4) * S + 4;
    return ((size_t)kMmaWarps * stages * 4 * gstride + (size_t)kMmaWarps * stages * 2 * CH + NT * 8 + 8) * 8 +
           (size_t)kMmaWarps * stages * 3 * 8 + 128;
}

// Ring geometry: 32-row chunks, as many stages per MMA warp as shared memory holds (2 at p = 50, 3+ for p <= 32); 16-row
// chunks only when two 32-row stages do not fit (px = 54..56).  Measured (profiles/r02_k1_small_experiments.md): 16-row
// chunks with 3-4 stages are 10-15 % slower at p = 50 (the per-row stage of a row warp then fills 16 of its 32 lanes),
// and more than 3 stages change nothing at p = 20 / 32.
static SmallCfg small_cfg(int px, int p) {
    SmallCfg c;
    c.NT = (p + 7) / 8;
    const int S = px + (px & 1);
    const size_t cap = 227 * 1024;
    c.CH = 32;
    if (const char* ev = std::getenv("DLSA_K1_CH")) {  // tuning override (tools/k1_sweep.py): 16 or 32
        const int v = std::atoi(ev);
        if (v == 16 || v == 32) c.CH = v;
    }
    int max_st = 8;
    if (const char* ev = std::getenv("DLSA_K1_STAGES")) {
        const int v = std::atoi(ev);
        if (v >= 2 && v <= 16) max_st = v;
    }
    if (small_smem(c.NT, S, c.CH, 2) > cap) c.CH = 16;
    c.STAGES = 2;
    while (c.STAGES < max_st && small_smem(c.NT, S, c.CH, c.STAGES + 1) <= cap) c.STAGES++;
    c.smem = small_smem(c.NT, S, c.CH, c.STAGES);
    return c;
}

int irls_small_slices(int K, int p, int64_t max_part_rows) {
    const int CH = 32;
    int64_t chunks = (max_part_rows + CH - 1) / CH;
    int64_t s_want = (148 * 4 + K - 1) / K;
    int64_t s_max = chunks / (kMmaWarps * 2);
    if (s_max < 1) s_max = 1;
    int64_t s = s_want < s_max ? s_want : s_max;
    (void)p;
    return (int)(s < 1 ? 1 : s);
}

template <int NT, int REM, int CH>
static int launch_cfg(const IrlsArgs& a, int family, size_t smem, int grid, cudaStream_t st) {
    if (family != DLSA_FAMILY_LINEAR) {
        auto kern = irls_small_kernel<NT, REM, CH, DLSA_FAMILY_LOGISTIC>;
        DLSA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, kThreads, smem, st>>>(a);
    } else {
        auto kern = irls_small_kernel<NT, REM, CH, DLSA_FAMILY_LINEAR>;
        DLSA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, kThreads, smem, st>>>(a);
    }
    count_launch();
    return check_launch("irls_small_kernel");
}

template <int NT, int CH>
static int launch_rem(const IrlsArgs& a, int rem, int family, size_t smem, int grid, cudaStream_t st) {
    if (NT >= 2) {  // ragged last tile row (p % 8 in 1..4) goes to DFMA
        switch (rem) {
            case 1: return launch_cfg<NT, (NT >= 2 ? 1 : 0), CH>(a, family, smem, grid, st);
            case 2: return launch_cfg<NT, (NT >= 2 ? 2 : 0), CH>(a, family, smem, grid, st);
            // NT = 7 with 3 or 4 ragged columns: the ragged accumulators no longer fit beside 21 tiles (spills)
            case 3: return launch_cfg<NT, ((NT >= 2 && NT <= 6) ? 3 : 0), CH>(a, family, smem, grid, st);
            case 4: return launch_cfg<NT, ((NT >= 2 && NT <= 6) ? 4 : 0), CH>(a, family, smem, grid, st);
            default: break;
        }
    }
    return launch_cfg<NT, 0, CH>(a, family, smem, grid, st);
}

int irls_small_launch(const double* X, int64_t ldx, const double* y, const int64_t* part_off, int K, int px,
                      int icpt, int family, const double* beta, const int* active, double* g, double* H,
                      double* loglik, int64_t max_part_rows, void* workspace, size_t workspace_bytes,
                      cudaStream_t st) {
    const int p = px + (icpt ? 1 : 0);
    SmallCfg c = small_cfg(px, p);
    const int S = irls_small_slices(K, p, max_part_rows);
    IrlsArgs a;
    a.X = X; a.ldx = ldx; a.y = y; a.part_off = part_off; a.beta = beta; a.active = active;
    a.K = K; a.px = px; a.p = p; a.icpt = icpt ? 1 : 0; a.S = S; a.stages = c.STAGES;
    // TMA staging copies GR consecutive rows with one bulk copy: rows must be contiguous (ldx == px), 16-byte aligned
    a.aligned = ((reinterpret_cast<uintptr_t>(X) & 15) == 0 && ldx == px && (px % 2) == 0) ? 1 : 0;
    a.want_ll = loglik != nullptr;
    a.poisson = family == DLSA_FAMILY_POISSON ? 1 : 0;
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
    const int rem = p % 8;
    int rc = 0;
#define DLSA_CASE(NT_)                                                              \
    case NT_:                                                                       \
        rc = (c.CH == 32) ? launch_rem<NT_, 32>(a, rem, family, c.smem, grid, st)   \
                          : launch_rem<NT_, 16>(a, rem, family, c.smem, grid, st);  \
        break;
    switch (c.NT) {
        DLSA_CASE(1) DLSA_CASE(2) DLSA_CASE(3) DLSA_CASE(4) DLSA_CASE(5) DLSA_CASE(6) DLSA_CASE(7)
        default: DLSA_REQUIRE(false, DLSA_E_UNSUPPORTED, "irls small: bad config");
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
