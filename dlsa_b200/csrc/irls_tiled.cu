// K1 for 56 < p <= 512: fused Newton/IRLS pass with the Gram triangle tiled over the warps of a CTA (and, when the
// triangle does not fit one CTA's registers, over TG sibling CTAs that stream the same rows).
//
// Same arithmetic as irls_small.cu (dlsa/R/wlse_func.R:47-54, dlsa/models.py:114,130).  Design:
//   * The lower triangle of H is cut into BLOCKS of 2 tile rows x 4 tile columns of 8x8 DMMA tiles (<= 8 tiles,
//     32 accumulator registers per lane).  A block needs 2 weighted A fragments and 4 B fragments per 4-row k-step
//     (0.75 fragment loads per DMMA).  The host builds the block list and deals it to (sibling CTA, sub-partition,
//     warp, slot) so that every sub-partition's FP64 pipe carries the same number of DMMAs (the pipe is per
//     sub-partition; an unbalanced deal idles it).  Every tile has exactly one owner: no cross-warp reduction,
//     bitwise reproducible.  A diagonal block is PARTIAL (only tiles on or below the diagonal exist): a warp gets
//     at most one, in its last slot, and runs a body compiled for that block's shape — a predicated-off DMMA
//     still occupies the FP64 pipe for its 16 cycles (measured), so masking by predicate is not an option.
//   * The 12 MMA warps (setmaxnreg 152) do the Gram blocks and g = X'r (4 DFMAs per k-step in the warp that owns a column
//     group's diagonal block, reusing its B fragments; the intercept column is an ordinary column of ones in shared
//     memory) and nothing else: their loop is wait(full, wr) -> straight-line DMMA body -> arrive(empty).
//   * 3 link warps take whole chunks round-robin and run AHEAD of the MMA warps by up to the ring depth: eta = x'beta of
//     one row per lane (two columns per LDS.128, rotated so a quarter-warp wavefront is conflict-free), the sigmoid,
//     w, r, the log-likelihood, the padding-column repair for odd px.  (The previous version computed the eta partial
//     sums inside the MMA warps, LA chunks ahead: ncu showed them spending 18 % of their time there on a 2 000-cycle
//     serial stretch per call and the link warps idle 91 % of the time — profiles/r01_k1_tiled_p100_logit_before.md.)
//     1 producer warp keeps a ring of 16- or 32-row chunks full with one TMA bulk copy per row (UBLKCP; rows land with
//     a pitch == 4 mod 16 doubles so that fragment loads are bank-conflict free).  Stages are gated by full / wr / empty
//     mbarriers.
//   * Sibling CTAs (tile group tg = blockIdx.x % TG) are adjacent in the grid, run at the same time and read the
//     same rows: X comes from HBM once and from L2 TG-1 times.
#include <algorithm>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "common.cuh"

namespace dlsa {

constexpr int kTMma = 12;                       // MMA warps
constexpr int kTLink = 3;                       // link (sigmoid) warps
constexpr int kTWarps = kTMma + kTLink + 1;     // + producer = 16; setmaxnreg acts on whole warpgroups (4 aligned warps)
constexpr int kTThreads = kTWarps * 32;         // 512 -> 128 registers per thread at launch
// rows per chunk: 32 when at least 4 stages of 32 rows fit in shared memory (p <= ~200), else 16.  The per-chunk costs of
// an MMA warp (barrier waits, stage bookkeeping, the switch into its body, its share of the eta partial sums) are ~40 % of
// a 16-row chunk's time at p=100 (ncu, profiles/r01_k1_tiled_p100_logit_*): 32-row chunks halve them per row.
constexpr int kTMaxTG = 16;
constexpr int kTNBW = 3;                        // block slots per MMA warp
constexpr int kTRegsMma = 152;
constexpr int kTRegsAux = 56;
static_assert(kTMma * kTRegsMma + (kTWarps - kTMma) * kTRegsAux == kTWarps * 128, "register pool");
static_assert(kTMma % 4 == 0 && kTWarps % 4 == 0, "setmaxnreg needs whole warpgroups");
constexpr int kTiledMaxP = 512;

// shape of a partial block: valid tile columns in its first / second tile row
enum : int { kPatNone = 0, kPat34, kPat12, kPat40, kPat30, kPat10, kPatCount };
__host__ __device__ constexpr int pat_n0(int pat) {
    return pat == kPat34 ? 3 : pat == kPat12 ? 1 : pat == kPat40 ? 4 : pat == kPat30 ? 3 : pat == kPat10 ? 1 : 4;
}
__host__ __device__ constexpr int pat_n1(int pat) {
    return pat == kPat34 ? 4 : pat == kPat12 ? 2 : (pat == kPat40 || pat == kPat30 || pat == kPat10) ? 0 : 4;
}

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
    const double* pre_w;   // PRE mode: w, r (and the per-row log-likelihood terms) of every row, written by
    const double* pre_r;   // tiled_link_kernel, indexed by global row
    const double* pre_ll;
    int K, px, p, icpt, S, Sd, aligned, want_ll, NT, TG, stages, copy_cols;
    int poisson;  // GLM instantiation only: 0 = logistic link, 1 = log link (DLSA_FAMILY_POISSON)
    int ycol;     // linear family: y rides the Gram as column `ycol` (= p) of the staged rows, so g = X'y is row p of the
                  // augmented Gram and no warp issues gradient DFMAs; -1 = off (GLM families, p = 512)
    // block of MMA warp w, slot b of sibling tg: bits 0-7 valid-tile mask (bit 4*ii+jj), 8-15 row pair m, 16-23 column
    // group J, 24 g-owner flag, 25-27 partial pattern; 0 = empty slot.  Full blocks first, the partial one last.
    uint32_t sched[kTMaxTG * kTMma * kTNBW];
};

#ifndef DLSA_TILED_KS_UNROLL
#define DLSA_TILED_KS_UNROLL 1   // k-steps of a 16-row half chunk unrolled in the MMA bodies.  Rolled (1) beats 2 and 4 on every shape (r02 same-run A/B: c4 24.9 / 23.0 / 21.6 ms for 4 / 2 / 1; c5 26.2 / 25.2 / 24.7): the 12 MMA warps run different bodies and the unrolled ones do not fit the instruction caches (no_instruction stalls 0.3-1.1 per issue)
#endif
constexpr int kTKsUnroll = DLSA_TILED_KS_UNROLL;

template <int N>
__device__ __forceinline__ void t_setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void t_setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ void t_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// 4 k-steps (16 rows) of a chunk for an MMA warp that owns NF full blocks and (PAT != kPatNone) one partial block.
// HASG: this warp's partial block owns g of its 4 tile columns.  A template parameter, not a run-time flag: ptxas turns a
// warp-uniform "if (has_g)" around 4 DFMAs into predicated DFMAs, and a predicated-off DFMA still takes its turn on the
// FP64 pipe (ncu, profiles/r01_tiled_p100_lin_after: 0.44 DFMA issued per DMMA in EVERY warp, 16 of 40 per k-step real).
template <bool LOGIT, int NF, int PAT, int NH, bool HASG>
__device__ __forceinline__ void tiled_chunk_body(const double* __restrict__ T, const double* __restrict__ ws,
                                                 const double* __restrict__ rs, int Sd, const int (&offA)[kTNBW],
                                                 const int (&offB)[kTNBW], bool has_g, double (&C)[kTNBW][2][4][2],
                                                 double (&gacc)[4]) {
    constexpr int N0 = pat_n0(PAT), N1 = pat_n1(PAT);
#pragma unroll 1
    for (int hb = 0; hb < NH; hb++, T += (size_t)16 * Sd, ws += 16, rs += 16)
#pragma unroll(kTKsUnroll)
    for (int ks = 0; ks < 4; ks++) {
        const double wv = LOGIT ? ws[4 * ks] : 1.0;
        const double* Tr = T + (size_t)(4 * ks) * Sd;  // row 4ks + q
        // (Timing experiment, r02: one scaled A fragment pair shared by all full blocks of the warp — the upper bound of a
        // schedule that deals row strips — is SLOWER for the GLM bodies, p=500 24.75 -> 25.63 ms, and gains 3 % for the
        // linear ones: the w-scaling DMULs are not what the GLM bodies wait for.)
#pragma unroll
        for (int b = 0; b < NF; b++) {
            double a0 = Tr[offA[b]], a1 = Tr[offA[b] + 8];
            if (LOGIT) { a0 *= wv; a1 *= wv; }
            double f[4];
#pragma unroll
            for (int jj = 0; jj < 4; jj++) f[jj] = Tr[offB[b] + 8 * jj];
#pragma unroll
            for (int jj = 0; jj < 4; jj++) {
                dmma884(C[b][0][jj][0], C[b][0][jj][1], a0, f[jj]);
                dmma884(C[b][1][jj][0], C[b][1][jj][1], a1, f[jj]);
            }
        }
        if (PAT != kPatNone) {
            constexpr int b = NF < kTNBW ? NF : kTNBW - 1;
            double a0 = Tr[offA[b]], a1 = Tr[offA[b] + 8];
            if (LOGIT) { a0 *= wv; a1 *= wv; }
            double f[4];
#pragma unroll
            for (int jj = 0; jj < 4; jj++) f[jj] = Tr[offB[b] + 8 * jj];
#pragma unroll
            for (int jj = 0; jj < 4; jj++) {
                if (jj < N0) dmma884(C[b][0][jj][0], C[b][0][jj][1], a0, f[jj]);
                if (jj < N1) dmma884(C[b][1][jj][0], C[b][1][jj][1], a1, f[jj]);
            }
            if (HASG && has_g) {  // has_g: run-time owner flag (GLM instantiation, see the dispatch below)
                const double rv = rs[4 * ks];
#pragma unroll
                for (int jj = 0; jj < 4; jj++) gacc[jj] = fma(rv, f[jj], gacc[jj]);
            }
        }
    }
}

template <int FAMILY, int kTCH, bool PRE>
__global__ void __launch_bounds__(kTThreads, 1) irls_tiled_kernel(const __grid_constant__ TiledArgs a) {
    constexpr bool LOGIT = FAMILY == DLSA_FAMILY_LOGISTIC;
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

    // ---- shared memory: tiles [ST][CH*Sd] | ws, rs, eta [ST][CH] each | beta_s [Sd] | llred [4] | bars [ST][4]
    // (beta_s holds the feature coefficients only: 0 at the intercept / padding columns, so the eta partial sums
    // can run over whole 32-column stripes without a bounds test; the intercept coefficient is added as b0)
    const int stage_doubles = kTCH * Sd;
    double* tiles = reinterpret_cast<double*>(smem_raw);
    double* ws_all = tiles + (size_t)ST * stage_doubles;
    double* rs_all = ws_all + ST * kTCH;
    double* beta_s = rs_all + ST * kTCH;
    double* llred = beta_s + Sd;
    uint64_t* bars = reinterpret_cast<uint64_t*>(llred + 4);  // [ST][4]: full, wr, empty, (unused)

    for (int c = tid; c < Sd; c += kTThreads) {
        double bv = 0.0;
        if (LOGIT && c < px) bv = a.beta[(size_t)k * p + c + a.icpt];
        beta_s[c] = bv;
    }
    // zero the ring once (tail chunks leave stale finite rows behind, never uninitialised ones; columns >= px are
    // never written by the copies of an even-px matrix), then plant the intercept column
    for (int e = tid; e < ST * stage_doubles; e += kTThreads) tiles[e] = 0.0;
    __syncthreads();
    if (a.icpt)
        for (int r = tid; r < ST * kTCH; r += kTThreads) tiles[(size_t)r * Sd + px] = 1.0;
    if (tid < ST) {
        uint64_t* b = bars + tid * 4;
        mbar_init(smem_u32(&b[0]), a.aligned ? 1u : 32u);
        mbar_init(smem_u32(&b[1]), 1u);               // the link warp that owns the chunk
        mbar_init(smem_u32(&b[2]), (uint32_t)kTMma);
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
                mbar_wait(smem_u32(&bars[st * 4 + 2]), par);
                fence_proxy_async();
            }
            const int64_t r0 = (c_lo + ci) * kTCH;
            const int valid = (n_k - r0 < kTCH) ? (int)(n_k - r0) : kTCH;
            const uint32_t bar = smem_u32(&bars[st * 4 + 0]);
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
        // =========================== link warps: sigmoid, w, r, log-likelihood; one row per lane ============
        t_setmaxnreg_dec<kTRegsAux>();
        const int lw = warp - kTMma;            // takes chunks ci == lw (mod 3)
        const int row = lane & (kTCH - 1);
        const bool act = lane < kTCH;
        const bool fix_pad = a.copy_cols > px;  // odd px: the copy drags one foreign column along
        const double pad_val = a.icpt ? 1.0 : 0.0;
        // eta = x'beta of this lane's row (CH = 32) or half row (CH = 16: lanes l and l+16 share row l), two columns per
        // LDS.128.  The 8 lanes of a quarter-warp wavefront read rows 8 doubles apart modulo 32 banks (Sd == 4 mod 16), so
        // rows 4..7 of each octet start one column pair later (rotation inside the lane's pair range): conflict-free.
        const int npairs = (px + 1) / 2;
        const int nparts = 32 / kTCH;
        const int per_part = (npairs + nparts - 1) / nparts;
        const int pb = (lane / kTCH) * per_part;
        const int plen = (npairs - pb < per_part) ? (npairs - pb) : per_part;
        const int skew = (row >> 2) & 1;
        const double b0 = (LOGIT && a.icpt) ? a.beta[(size_t)k * p] : 0.0;
        double llacc = 0.0;
        auto load_y = [&](int ci) -> double {
            const int64_t r = (c_lo + ci) * kTCH + row;
            return (ci < my_n && r < n_k) ? __ldg(yp + r) : 0.0;
        };
        // PRE: the link stage ran as its own kernel; this warp only moves w, r (and the log-likelihood terms) of its
        // chunk from global to shared memory
        auto load_pre = [&](const double* src, int ci) -> double {
            const int64_t r = (c_lo + ci) * kTCH + row;
            return (ci < my_n && r < n_k) ? __ldg(src + (row_lo - a.part_off[0]) + r) : 0.0;
        };
        double y_next = PRE ? 0.0 : load_y(lw);
        double w_next = PRE ? load_pre(a.pre_w, lw) : 0.0, r_next = PRE ? load_pre(a.pre_r, lw) : 0.0;
        double l_next = (PRE && a.want_ll) ? load_pre(a.pre_ll, lw) : 0.0;
        int st = lw % ST;
        uint32_t par = (uint32_t)((lw / ST) & 1);
        for (int ci = lw; ci < my_n; ci += kTLink) {
            const int64_t r0 = (c_lo + ci) * kTCH;
            const int valid = (n_k - r0 < kTCH) ? (int)(n_k - r0) : kTCH;
            const bool rowvalid = act && row < valid;
            const double yv = y_next;
            if (!PRE) y_next = load_y(ci + kTLink);
            double wv, rv;
            if (PRE) {
                wv = w_next;
                rv = r_next;
                if (rowvalid) llacc += l_next;
                w_next = load_pre(a.pre_w, ci + kTLink);
                r_next = load_pre(a.pre_r, ci + kTLink);
                if (a.want_ll) l_next = load_pre(a.pre_ll, ci + kTLink);
                mbar_wait(smem_u32(&bars[st * 4 + 0]), par);  // the stage belongs to this chunk now
            } else if (LOGIT) {
                mbar_wait(smem_u32(&bars[st * 4 + 0]), par);  // rows landed
                const double* Tr = tiles + (size_t)st * stage_doubles + (size_t)row * Sd + 2 * pb;
                const double* bp = beta_s + 2 * pb;
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                int pj = skew < plen ? skew : 0;
#pragma unroll 2
                for (int j = 0; j < per_part; j++) {
                    if (j < plen) {
                        double2 xv = *reinterpret_cast<const double2*>(Tr + 2 * pj);
                        const double2 bv = *reinterpret_cast<const double2*>(bp + 2 * pj);
                        if (2 * (pb + pj) + 1 == px) xv.y = 0.0;  // the padding column of an odd px may hold anything
                        if (j & 1) { a2 = fma(xv.x, bv.x, a2); a3 = fma(xv.y, bv.y, a3); }
                        else       { a0 = fma(xv.x, bv.x, a0); a1 = fma(xv.y, bv.y, a1); }
                        if (++pj == plen) pj = 0;
                    }
                }
                double eta = (a0 + a1) + (a2 + a3);
                if (kTCH == 16) eta += shfl_xor_f64(eta, 16);
                eta += b0;
                if (a.poisson) {  // warp-uniform: mu = exp(eta), w = mu, r = y - mu (dlsa/R/wlse_func.R:155-168)
                    const double mu = exp_clamped(eta);
                    wv = mu;
                    rv = yv - mu;
                    if (a.want_ll && rowvalid) llacc += fma(yv, eta, -mu);
                } else {
                    const double e = exp_nonpos(-fabs(eta));
                    const double pr = div_pos((eta >= 0.0) ? 1.0 : e, 1.0 + e);
                    wv = pr * (1.0 - pr);
                    rv = yv - pr;
                    if (a.want_ll && rowvalid) llacc += yv * eta - (fmax(eta, 0.0) + log1p_01(e));
                }
            } else {
                mbar_wait(smem_u32(&bars[st * 4 + 0]), par);  // the stage belongs to this chunk now
                wv = 1.0;
                rv = yv;
                if (rowvalid) llacc = fma(yv, yv, llacc);
                // The linear body does not scale its A fragments by w, so the rows past `valid` of a tail chunk (what an
                // earlier chunk left in this stage) must not stay: blank them before the MMA warps are released.
                if (valid < kTCH) {
                    double* stale = tiles + (size_t)st * stage_doubles + (size_t)valid * Sd;
                    for (int e = lane; e < (kTCH - valid) * Sd; e += 32) stale[e] = 0.0;
                    fence_proxy_async();
                }
            }
            if (act) {
                ws_all[st * kTCH + row] = rowvalid ? wv : 0.0;
                rs_all[st * kTCH + row] = rowvalid ? rv : 0.0;
                // (linear family: the A fragments are not scaled by w, so an intercept 1.0 in a row past `valid` would be
                // counted — found by tests/test_gpu_parity.py::test_linear_pass_y_column_edge_widths, odd px + intercept)
                if (fix_pad) tiles[(size_t)st * stage_doubles + (size_t)row * Sd + px] = (LOGIT || rowvalid) ? pad_val : 0.0;
                if (!LOGIT && a.ycol >= 0) tiles[(size_t)st * stage_doubles + (size_t)row * Sd + a.ycol] = rowvalid ? yv : 0.0;
            }
            if (fix_pad || (!LOGIT && a.ycol >= 0)) fence_proxy_async();  // generic-proxy writes into a stage the TMA will overwrite later
            __syncwarp();
            if (lane == 0) t_mbar_arrive(smem_u32(&bars[st * 4 + 1]));
            st += kTLink;
            if (st >= ST) { st -= ST; par ^= 1u; }
        }
        if (tg == 0 && a.outll != nullptr) {
#pragma unroll
            for (int m = kTCH / 2; m >= 1; m >>= 1) llacc += shfl_xor_f64(llacc, m);  // lanes 0..CH-1 hold the rows
            if (lane == 0) llred[lw] = llacc;
            asm volatile("bar.sync 1, 96;" ::: "memory");
            if (lw == 0 && lane == 0) a.outll[item] = (llred[0] + llred[1]) + llred[2];
        }
    } else {
        // =========================== MMA warps: H blocks on DMMA, g, eta partial sums =======================
        t_setmaxnreg_inc<kTRegsMma>();
        int offA[kTNBW], offB[kTNBW];
        int nfull = 0, pat = kPatNone;
        bool has_g = false;
#pragma unroll
        for (int b = 0; b < kTNBW; b++) {
            const uint32_t e = a.sched[(tg * kTMma + warp) * kTNBW + b];
            offA[b] = 16 * (int)((e >> 8) & 0xffu) + a8;    // column of tile row 2m (tile row 2m+1 is +8)
            offB[b] = 32 * (int)((e >> 16) & 0xffu) + a8;   // column of tile column 4J (+8 per tile)
            if ((e & 0xffu) == 0xffu) nfull = b + 1;        // full blocks come first
            else if (e & 0xffu) { pat = (int)((e >> 25) & 7u); has_g = ((e >> 24) & 1u) != 0u; }
        }
        double C[kTNBW][2][4][2];
#pragma unroll
        for (int b = 0; b < kTNBW; b++)
#pragma unroll
            for (int ii = 0; ii < 2; ii++)
#pragma unroll
                for (int jj = 0; jj < 4; jj++) C[b][ii][jj][0] = C[b][ii][jj][1] = 0.0;
        double gacc[4] = {0.0, 0.0, 0.0, 0.0};

        const uint32_t bars_u32 = smem_u32(bars);

        // Per chunk: ONE shared loop (barriers, stage counters, eta) and a switch into the k-step body compiled for
        // this warp's (number of full blocks, shape of the partial block).  The bodies are straight-line code with
        // unconditional fragment loads (every column a block can name exists in shared memory); sharing the rest
        // keeps the instruction footprint of the 12 MMA warps inside the 32 KB instruction cache.
        // Linear family: owners and non-owners run different bodies (same-run A/B, 5e7 rows x p=100: 23.35 -> 21.70 ms).
        // GLM family: every warp runs the owner body with the run-time flag — the compiled-out variant measured SLOWER
        // there (p=100 fused link 25.4 -> 27.3 ms, p=500 28.3 -> 29.2 ms; only p=182 gained, 23.6 -> 22.1), and again with the
        // rolled k-step loop (p=182 22.6 -> 23.1, p=100 25.5 -> 25.8, p=500 unchanged), so it keeps the round-1 code;
        // profiles/r02_k1_tiled_ab.md.
        const int path = pat * 8 + ((LOGIT || has_g) ? 4 : 0) + nfull;
        int st = 0;
        uint32_t par = 0;
        for (int ci = 0; ci < my_n; ci++) {
            mbar_wait(bars_u32 + (uint32_t)(st * 4 + 0) * 8u, par);  // rows landed (async-proxy writes visible here)
            mbar_wait(bars_u32 + (uint32_t)(st * 4 + 1) * 8u, par);  // w, r written, pad column repaired
            const double* T = tiles + (size_t)st * stage_doubles + (size_t)q * Sd;
            const double* ws = ws_all + st * kTCH + q;
            const double* rs = rs_all + st * kTCH + q;
#define DLSA_BODY(NF_, PAT_)                                                                                     \
    case (PAT_) * 8 + (NF_): tiled_chunk_body<LOGIT, NF_, PAT_, kTCH / 16, false>(T, ws, rs, Sd, offA, offB, false, C, gacc); break; \
    case (PAT_) * 8 + 4 + (NF_): tiled_chunk_body<LOGIT, NF_, PAT_, kTCH / 16, true>(T, ws, rs, Sd, offA, offB, has_g, C, gacc); break;
            switch (path) {
                DLSA_BODY(0, kPatNone) DLSA_BODY(1, kPatNone) DLSA_BODY(2, kPatNone) DLSA_BODY(3, kPatNone)
                DLSA_BODY(0, kPat34) DLSA_BODY(1, kPat34) DLSA_BODY(2, kPat34)
                DLSA_BODY(0, kPat12) DLSA_BODY(1, kPat12) DLSA_BODY(2, kPat12)
                DLSA_BODY(0, kPat40) DLSA_BODY(1, kPat40) DLSA_BODY(2, kPat40)
                DLSA_BODY(0, kPat30) DLSA_BODY(1, kPat30) DLSA_BODY(2, kPat30)
                DLSA_BODY(0, kPat10) DLSA_BODY(1, kPat10) DLSA_BODY(2, kPat10)
                default: break;
            }
#undef DLSA_BODY
            __syncwarp();
            if (lane == 0) t_mbar_arrive(bars_u32 + (uint32_t)(st * 4 + 2) * 8u);  // done with the stage
            if (++st == ST) { st = 0; par ^= 1u; }
        }

        // ---- emit: every tile (and every column of g) has exactly one owner -------------------------------------
        double* oH = a.outH + (size_t)item * p * p;
#pragma unroll
        for (int b = 0; b < kTNBW; b++) {
            const uint32_t e = a.sched[(tg * kTMma + warp) * kTNBW + b];
            const uint32_t bm = e & 0xffu;
            const int m = (int)((e >> 8) & 0xffu), J = (int)((e >> 16) & 0xffu);
#pragma unroll
            for (int ii = 0; ii < 2; ii++)
#pragma unroll
                for (int jj = 0; jj < 4; jj++) {
                    if (bm & (1u << (4 * ii + jj))) {
                        const int i = 2 * m + ii, j = 4 * J + jj;
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const int r_ = 8 * i + a8, c_ = 8 * j + 2 * q + h;
                            if (!LOGIT && r_ == a.ycol && c_ < p) {   // row p of the augmented Gram: g = X'y
                                a.outg[(size_t)item * p + (a.icpt ? (c_ == px ? 0 : c_ + 1) : c_)] = C[b][ii][jj][h];
                            } else if (r_ < p && c_ < p && (i != j || c_ <= r_)) {
                                const int ro = a.icpt ? (r_ == px ? 0 : r_ + 1) : r_;
                                const int co = a.icpt ? (c_ == px ? 0 : c_ + 1) : c_;
                                oH[(size_t)ro * p + co] = C[b][ii][jj][h];
                                oH[(size_t)co * p + ro] = C[b][ii][jj][h];
                            }
                        }
                    }
                }
            if (bm != 0u && bm != 0xffu && ((e >> 24) & 1u)) {
                double* og = a.outg + (size_t)item * p;
#pragma unroll
                for (int jj = 0; jj < 4; jj++) {
                    double v = gacc[jj];
                    v += shfl_xor_f64(v, 1);
                    v += shfl_xor_f64(v, 2);
                    const int c = 32 * J + 8 * jj + a8;
                    if (q == 0 && c < p) og[a.icpt ? (c == px ? 0 : c + 1) : c] = v;
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

// The link stage as its own streaming kernel (PRE mode, used when the Gram needs sibling CTAs or p is large): for every
// row eta = x'beta_k, prob = sigmoid(eta), w = prob(1-prob), r = y - prob and the log-likelihood term, written to
// w[n], r[n], ll[n] (row index relative to part_off[0]: the caller may pass a window of a longer offsets array).  With TG sibling CTAs per partition slice the fused link stage is computed TG
// times over, by 3 warps per CTA whose FP64 instructions queue behind 12 DMMA streams (~100 cycles each, measured): at
// p = 500 that is slower than reading X a second time at HBM speed.  One warp takes 32 rows: 8 batches of 4 rows with
// lane = column stripe (coalesced), a transposed butterfly per batch, then ONE sigmoid per lane for the 32 rows.
constexpr int kLinkWarps = 8;
__global__ void __launch_bounds__(kLinkWarps * 32) tiled_link_kernel(const double* __restrict__ X, int64_t ldx,
                                                                     const double* __restrict__ y,
                                                                     const int64_t* __restrict__ part_off,
                                                                     const double* __restrict__ beta,
                                                                     const int* __restrict__ active, int px, int icpt,
                                                                     int poisson, double* __restrict__ wout,
                                                                     double* __restrict__ rout, double* __restrict__ llout) {
    extern __shared__ double link_s[];  // beta [32 * jn] | eta [kLinkWarps][32]
    const int k = blockIdx.y;
    if (active != nullptr && active[k] == 0) return;
    const int p = px + icpt;
    const int jn = (px + 31) / 32;
    double* beta_s = link_s;
    double* eta_s = link_s + 32 * jn;
    for (int c = threadIdx.x; c < 32 * jn; c += blockDim.x) beta_s[c] = (c < px) ? beta[(size_t)k * p + c + icpt] : 0.0;
    __syncthreads();
    const double b0 = icpt ? beta[(size_t)k * p] : 0.0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bit4 = (lane >> 4) & 1, bit3 = (lane >> 3) & 1;
    const int64_t row_lo = part_off[k];
    const int64_t n_k = part_off[k + 1] - row_lo;
    const double* Xp = X + row_lo * ldx;
    y += row_lo;
    const int64_t out_lo = row_lo - part_off[0];
    wout += out_lo;
    rout += out_lo;
    if (llout != nullptr) llout += out_lo;
    for (int64_t base = ((int64_t)blockIdx.x * kLinkWarps + warp) * 32; base < n_k;
         base += (int64_t)gridDim.x * kLinkWarps * 32) {
#pragma unroll 1
        for (int b = 0; b < 8; b++) {
            const int64_t r = base + 4 * b;
            if (r >= n_k) break;  // warp-uniform
            const double* x0 = Xp + r * ldx + lane;
            // rows past the end of the partition re-read its last row: finite values, results discarded
            const int64_t last = n_k - 1 - r;
            const double* x1 = x0 + (last >= 1 ? 1 : last) * ldx;
            const double* x2 = x0 + (last >= 2 ? 2 : last) * ldx;
            const double* x3 = x0 + (last >= 3 ? 3 : last) * ldx;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll 2
            for (int j = 0; j < jn; j++) {
                const bool in = lane + 32 * j < px;
                const double bv = beta_s[lane + 32 * j];
                const double v0 = in ? __ldg(x0 + 32 * j) : 0.0, v1 = in ? __ldg(x1 + 32 * j) : 0.0;
                const double v2 = in ? __ldg(x2 + 32 * j) : 0.0, v3 = in ? __ldg(x3 + 32 * j) : 0.0;
                s0 = fma(v0, bv, s0);
                s1 = fma(v1, bv, s1);
                s2 = fma(v2, bv, s2);
                s3 = fma(v3, bv, s3);
            }
            const double t0 = (bit4 ? s2 : s0) + shfl_xor_f64(bit4 ? s0 : s2, 16);
            const double t1 = (bit4 ? s3 : s1) + shfl_xor_f64(bit4 ? s1 : s3, 16);
            double u = (bit3 ? t1 : t0) + shfl_xor_f64(bit3 ? t0 : t1, 8);
            u += shfl_xor_f64(u, 4);
            u += shfl_xor_f64(u, 2);
            u += shfl_xor_f64(u, 1);
            if ((lane & 7) == 0) eta_s[warp * 32 + 4 * b + 2 * bit4 + bit3] = u;
        }
        __syncwarp();
        const int64_t r = base + lane;
        if (r < n_k) {
            const double eta = eta_s[warp * 32 + lane] + b0;
            const double yv = __ldg(y + r);
            if (poisson) {
                const double mu = exp_clamped(eta);
                wout[r] = mu;
                rout[r] = yv - mu;
                if (llout != nullptr) llout[r] = fma(yv, eta, -mu);
            } else {
                const double e = exp_nonpos(-fabs(eta));
                const double pr = div_pos((eta >= 0.0) ? 1.0 : e, 1.0 + e);
                wout[r] = pr * (1.0 - pr);
                rout[r] = yv - pr;
                if (llout != nullptr) llout[r] = yv * eta - (fmax(eta, 0.0) + log1p_01(e));
            }
        }
        __syncwarp();
    }
}

// ---- host side ---------------------------------------------------------------------------------------------
struct TiledCfg {
    int NT, TG, Sd, stages, CH, pre;
    size_t smem;
    std::vector<uint32_t> sched;
};

static inline int popcount8(uint32_t m) { return __builtin_popcount(m & 0xffu); }

static int pattern_of(uint32_t mask) {
    const int n0 = popcount8(mask & 0x0fu), n1 = popcount8(mask & 0xf0u);
    if (n0 == 4 && n1 == 4) return kPatNone;
    if (n0 == 3 && n1 == 4) return kPat34;
    if (n0 == 1 && n1 == 2) return kPat12;
    if (n0 == 4 && n1 == 0) return kPat40;
    if (n0 == 3 && n1 == 0) return kPat30;
    if (n0 == 1 && n1 == 0) return kPat10;
    return -1;
}

// ycol: the linear family's Gram is that of [X | y] (p + 1 columns) and nobody owns g; GLM: p columns, g owned per column group
static TiledCfg tiled_cfg(int px, int p, bool ycol = false) {
    TiledCfg c;
    c.NT = (p + (ycol ? 1 : 0) + 7) / 8;
    const int MR = (c.NT + 1) / 2;
    struct Blk { int m, J; uint32_t mask; bool g; };
    std::vector<Blk> full, part;
    for (int m = 0; m < MR; m++)
        for (int J = 0; 4 * J <= 2 * m + 1 && 4 * J < c.NT; J++) {
            uint32_t mask = 0;
            for (int ii = 0; ii < 2; ii++)
                for (int jj = 0; jj < 4; jj++) {
                    const int i = 2 * m + ii, j = 4 * J + jj;
                    if (i < c.NT && j <= i) mask |= 1u << (4 * ii + jj);
                }
            if (!mask) continue;
            // g of tile columns 4J..4J+3 belongs to the lowest diagonal block of that column group
            const int mg = (2 * J + 1 < MR) ? 2 * J + 1 : MR - 1;
            if (mask == 0xffu) full.push_back({m, J, mask, false});
            else part.push_back({m, J, mask, !ycol && m == mg});
        }
    // (Tried and measured slower: one body per BLOCK instead of per (full-block count, partial shape), blocks run one
    // after the other, so that any mix of shapes can share a warp and the deal can equalise tiles per warp (p=100: 7..8
    // instead of 4..11).  C4 22.9 -> 25.2 ms, p=100 logistic 25.4 -> 29.1, C5 28.2 -> 29.8: the interleaved body keeps 24
    // independent DMMAs and all fragment loads of a k-step in flight, which is worth more than the balance.)
    // Siblings: enough MMA warps for one partial block each and enough slots for everything.  The FP64 pipe is per
    // sub-partition, so what has to be equal is the number of DMMAs per (sibling, sub-partition) bin: all blocks,
    // heaviest first (full blocks, then partial ones by tile count), each to the least-loaded bin that can still take
    // it (<= 9 blocks = 3 warps x 3 slots, <= 3 partial blocks = one per warp).  Measured with the previous deal
    // (partial blocks first, then full ones): p=100 -> bins 27/19/26/19 tiles, the 27 set the pace (ncu: 23 % of the
    // MMA warps' time waiting on stage barriers); this deal gives 23/23/23/22.
    const int nb = (int)(full.size() + part.size());
    c.TG = (nb + kTMma * kTNBW - 1) / (kTMma * kTNBW);
    if (c.TG < 1) c.TG = 1;
    while (c.TG * kTMma < (int)part.size()) c.TG++;
    auto heavier = [](const Blk& x, const Blk& y) { return popcount8(x.mask) > popcount8(y.mask); };
    std::stable_sort(part.begin(), part.end(), heavier);
    std::vector<std::vector<Blk>> bfull, bpart;
    for (;; c.TG++) {
        const int nbins = c.TG * 4;   // (sibling, sub-partition); MMA warp w runs on sub-partition w % 4
        std::vector<int> load(nbins, 0);
        bfull.assign(nbins, {});
        bpart.assign(nbins, {});
        bool ok = true;
        auto place = [&](const Blk& b, bool partial) {
            int best = -1;
            for (int i = 0; i < nbins; i++) {
                if ((int)(bfull[i].size() + bpart[i].size()) >= 3 * kTNBW) continue;
                if (partial && (int)bpart[i].size() >= 3) continue;
                if (best < 0 || load[i] < load[best]) best = i;
            }
            if (best < 0) { ok = false; return; }
            (partial ? bpart : bfull)[best].push_back(b);
            load[best] += popcount8(b.mask);
        };
        for (const Blk& b : full) place(b, false);
        for (const Blk& b : part) place(b, true);
        if (ok || c.TG >= kTMaxTG) break;
    }
    // inside a bin: partial blocks to distinct warps (a warp runs at most one, in its last slot), full blocks to the
    // least-loaded warp with a free slot
    std::vector<std::vector<Blk>> wfull(c.TG * kTMma);
    std::vector<Blk> wpart(c.TG * kTMma, Blk{0, 0, 0u, false});
    std::vector<int> wload(c.TG * kTMma, 0);
    for (int i = 0; i < c.TG * 4; i++) {
        const int tg = i / 4, smsp = i % 4;
        for (size_t t = 0; t < bpart[i].size(); t++) {
            const int w = tg * kTMma + smsp + 4 * (int)t;
            wpart[w] = bpart[i][t];
            wload[w] += popcount8(bpart[i][t].mask);
        }
        for (const Blk& b : bfull[i]) {
            int wsel = -1;
            for (int t = 0; t < 3; t++) {
                const int w = tg * kTMma + smsp + 4 * t;
                const int used = (int)wfull[w].size() + (wpart[w].mask ? 1 : 0);
                if (used < kTNBW && (wsel < 0 || wload[w] < wload[wsel])) wsel = w;
            }
            wfull[wsel].push_back(b);
            wload[wsel] += 8;
        }
    }
    c.sched.assign((size_t)kTMaxTG * kTMma * kTNBW, 0u);
    for (int w = 0; w < c.TG * kTMma; w++) {
        int slot = 0;
        auto enc = [&](const Blk& b) {
            return b.mask | ((uint32_t)b.m << 8) | ((uint32_t)b.J << 16) | ((b.g ? 1u : 0u) << 24) |
                   ((uint32_t)pattern_of(b.mask) << 25);
        };
        for (const Blk& b : wfull[w]) c.sched[(size_t)w * kTNBW + slot++] = enc(b);
        if (wpart[w].mask) c.sched[(size_t)w * kTNBW + slot++] = enc(wpart[w]);
    }
    int cols = 8 * 4 * ((c.NT + 3) / 4);   // every fragment column a block can touch exists in shared memory
    if (cols < p + 1) cols = p + 1;
    c.Sd = cols;
    while (c.Sd % 16 != 4) c.Sd++;
    const size_t fixed = ((size_t)c.Sd + 4) * 8 + 256;
    auto stage_bytes = [&](int ch) { return ((size_t)ch * c.Sd + 2 * ch) * 8 + 4 * 8; };
    c.CH = 32;
    if ((220 * 1024 - fixed) / stage_bytes(32) < 3) c.CH = 16;
    if (const char* ev = std::getenv("DLSA_TILED_CH")) {  // tuning override (tools/k1_sweep.py): 16 or 32
        const int v = std::atoi(ev);
        if (v == 16 || (v == 32 && (220 * 1024 - fixed) / stage_bytes(32) >= 3)) c.CH = v;
    }
    // link stage: fused into the Gram kernel while one CTA holds the whole triangle, as a streaming kernel of its own
    // (X read twice, the second time far below the FP64 time of the Gram) once sibling CTAs would each repeat it
    c.pre = c.TG >= 2 ? 1 : 0;
    if (const char* ev = std::getenv("DLSA_TILED_PRE")) c.pre = std::atoi(ev) ? 1 : 0;  // tuning override
    const size_t per_stage = stage_bytes(c.CH);
    int st = (int)((220 * 1024 - fixed) / per_stage);
    // A multiple of 3 (the number of link warps): link warp l then owns the stages == l (mod 3) and sees every phase of
    // their barriers in order.  With any other ring size a link warp comes round to a stage whose previous phase belongs
    // to another link warp's chunk and may still be pending: a parity wait cannot tell "not yet" from "one phase ahead".
    st = st > 12 ? 12 : st - st % 3;
    c.stages = st;
    c.smem = fixed + per_stage * c.stages;
    return c;
}

static int tiled_slices(int K, int TG, int CH, int64_t max_part_rows) {
    // Slices per partition: S > 1 fills the SMs when K * TG CTAs do not, at a price per CTA — ring fill, accumulator
    // flush, and a partial H that a second kernel has to sum.  Model: time ~ waves(S) x (chunks per CTA + overhead), the
    // overhead taken as 8 chunks (r02 ncu at p = 500, 31-chunk CTAs: 10 % of the samples in prologue + epilogue, plus
    // 10 GB of partials written and read back per pass; the old rule — best wave fill, 1 % hysteresis — took S = 5 there
    // for 98.8 -> 99.9 % fill).
    const int64_t chunks = (max_part_rows + CH - 1) / CH;
    int64_t s_cap = chunks / 8;
    if (s_cap < 1) s_cap = 1;
    if (s_cap > 32) s_cap = 32;
    const int64_t sms = 148;
    const double ovh = 8.0;
    double best_cost = 0.0;
    int best_s = 1;
    for (int64_t s = 1; s <= s_cap; s++) {
        const int64_t u = (int64_t)K * TG * s;
        const double cost = (double)((u + sms - 1) / sms) * ((double)((chunks + s - 1) / s) + ovh);
        if (s == 1 || cost < 0.98 * best_cost) { best_cost = cost; best_s = (int)s; }  // a larger S must buy > 2 %
    }
    return best_s;
}

static size_t tiled_slice_bytes(int K, int S, int p) {
    return S == 1 ? 0 : (size_t)K * S * ((size_t)p * p + p + 1) * sizeof(double);
}

// The streaming link kernel addresses its w / r / log-lik buffers by global row, and the host only knows the bound
// K * max_part_rows for the row count.  With very ragged partitions that bound can dwarf the data; past 16 GB the fused
// link warps are used instead (same results, slower at large p).
// Host-only diagnostic behind dlsa_irls_tiled_schedule (include/dlsa_b200.h): the block schedule and ring geometry
// the tiled kernel would use for this shape.  tests/test_abi_and_host.py checks on the CPU, for every p, that each
// lower-triangle tile has exactly one owner and that the ring fits shared memory.
int irls_tiled_schedule_info(int px, int icpt, uint32_t* sched_out, int* info) {
    const int p = px + (icpt ? 1 : 0);
    if (p <= 56 || p > kTiledMaxP) return -1;
    TiledCfg c = tiled_cfg(px, p);
    for (size_t i = 0; i < c.sched.size(); i++) sched_out[i] = c.sched[i];
    info[0] = c.NT; info[1] = c.TG; info[2] = c.CH; info[3] = c.stages; info[4] = c.pre; info[5] = c.Sd;
    info[6] = (int)c.smem; info[7] = kTMma * kTNBW;
    return 0;
}

static bool tiled_linear_ycol(int p) { return p + 1 <= kTiledMaxP; }

static bool tiled_use_pre(const TiledCfg& c, int K, int64_t max_part_rows) {
    return c.pre && 3.0 * sizeof(double) * (double)K * (double)max_part_rows <= 16.0 * 1024 * 1024 * 1024;
}

// workspace = [slice partials | w, r, ll of every row (PRE mode; rows <= K * max_part_rows)]
size_t irls_tiled_workspace_bytes(int K, int p, int64_t max_part_rows) {
    if (p > kTiledMaxP) return 0;
    TiledCfg c = tiled_cfg(p, p);
    int S = tiled_slices(K, c.TG, c.CH, max_part_rows);
    if (tiled_linear_ycol(p)) {   // the family is not known here: room for whichever layout slices more
        TiledCfg cl = tiled_cfg(p, p, true);
        const int Sl = tiled_slices(K, cl.TG, cl.CH, max_part_rows);
        if (Sl > S) S = Sl;
    }
    size_t need = tiled_slice_bytes(K, S, p);
    if (tiled_use_pre(c, K, max_part_rows)) need += 3 * (size_t)K * (size_t)max_part_rows * sizeof(double);
    return need;
}

int irls_tiled_launch(const double* X, int64_t ldx, const double* y, const int64_t* part_off, int K, int px, int icpt,
                      int family, const double* beta, const int* active, double* g, double* H, double* loglik,
                      int64_t max_part_rows, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    const int p = px + (icpt ? 1 : 0);
    DLSA_REQUIRE(p <= kTiledMaxP, DLSA_E_UNSUPPORTED, "irls pass: p=%d > %d not supported", p, kTiledMaxP);
    const bool ycol = family == DLSA_FAMILY_LINEAR && tiled_linear_ycol(p);
    TiledCfg c = tiled_cfg(px, p, ycol);
    DLSA_REQUIRE(c.TG <= kTMaxTG && c.stages >= 3, DLSA_E_UNSUPPORTED, "irls pass: p=%d does not fit (TG=%d, stages=%d)",
                 p, c.TG, c.stages);
    const int S = tiled_slices(K, c.TG, c.CH, max_part_rows);
    TiledArgs a;
    a.X = X; a.ldx = ldx; a.y = y; a.part_off = part_off; a.beta = beta; a.active = active;
    a.K = K; a.px = px; a.p = p; a.icpt = icpt ? 1 : 0; a.S = S; a.Sd = c.Sd;
    a.want_ll = loglik != nullptr; a.NT = c.NT; a.TG = c.TG; a.stages = c.stages;
    // one 16-byte-aligned bulk copy per row: even row pitch; an odd px drags the next (padding) column along
    a.aligned = ((reinterpret_cast<uintptr_t>(X) & 15) == 0 && (ldx % 2) == 0) ? 1 : 0;
    a.copy_cols = (a.aligned && (px % 2)) ? px + 1 : px;
    for (size_t i = 0; i < c.sched.size(); i++) a.sched[i] = c.sched[i];
    double *pH = nullptr, *pg = nullptr, *pll = nullptr;
    const bool glm = family != DLSA_FAMILY_LINEAR;  // logistic or Poisson: the GLM instantiation, link chosen at run time
    a.poisson = family == DLSA_FAMILY_POISSON ? 1 : 0;
    a.ycol = ycol ? p : -1;
    const bool pre = glm && tiled_use_pre(c, K, max_part_rows);
    const size_t slice_bytes = tiled_slice_bytes(K, S, p);
    const size_t need = slice_bytes + (pre ? 3 * (size_t)K * (size_t)max_part_rows * sizeof(double) : 0);
    DLSA_REQUIRE(need == 0 || (workspace != nullptr && workspace_bytes >= need), DLSA_E_WORKSPACE,
                 "irls pass: workspace too small (%zu < %zu)", workspace_bytes, need);
    a.pre_w = a.pre_r = a.pre_ll = nullptr;
    if (pre) {
        double* wbuf = reinterpret_cast<double*>(static_cast<unsigned char*>(workspace) + slice_bytes);
        double* rbuf = wbuf + (size_t)K * max_part_rows;
        double* lbuf = rbuf + (size_t)K * max_part_rows;
        a.pre_w = wbuf; a.pre_r = rbuf; a.pre_ll = a.want_ll ? lbuf : nullptr;
        const size_t lsm = ((size_t)32 * ((px + 31) / 32) + kLinkWarps * 32) * sizeof(double);
        // one wave: as many CTAs per partition as keep every SM's resident slots filled, no partial second wave
        int resident = 4;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, tiled_link_kernel, kLinkWarps * 32, lsm) != cudaSuccess ||
            resident < 1)
            resident = 4;
        const int gx_max = (int)((max_part_rows + kLinkWarps * 32 - 1) / (kLinkWarps * 32));
        const int slots = 148 * resident;
        int gx = 1;
        double best = -1.0;
        for (int cand = 1; cand <= gx_max && (int64_t)K * cand <= 4 * (int64_t)slots + K; cand++) {
            const int64_t ctas = (int64_t)K * cand;
            const double eff = (double)ctas / (double)(((ctas + slots - 1) / slots) * slots);
            if (eff > best + 0.02) { best = eff; gx = cand; }  // more CTAs per partition must buy > 2 % of wave efficiency
        }
        tiled_link_kernel<<<dim3(gx, K), kLinkWarps * 32, lsm, st>>>(X, ldx, y, part_off, beta, active, px, a.icpt,
                                                                     a.poisson, wbuf, rbuf, a.want_ll ? lbuf : nullptr);
        count_launch();
        int rc0 = check_launch("tiled_link_kernel");
        if (rc0) return rc0;
    }
    if (S == 1) {
        a.outH = H; a.outg = g; a.outll = loglik;
    } else {
        pH = static_cast<double*>(workspace);
        pg = pH + (size_t)K * S * p * p;
        pll = pg + (size_t)K * S * p;
        a.outH = pH; a.outg = pg; a.outll = pll;
    }
    const int grid = K * S * c.TG;
#define DLSA_TILED_LAUNCH(FAM_, CH_, PRE_)                                                                        \
    {                                                                                                             \
        auto kern = irls_tiled_kernel<FAM_, CH_, PRE_>;                                                           \
        DLSA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));          \
        kern<<<grid, kTThreads, c.smem, st>>>(a);                                                                 \
    }
    if (glm && pre) {
        if (c.CH == 32) DLSA_TILED_LAUNCH(DLSA_FAMILY_LOGISTIC, 32, true) else DLSA_TILED_LAUNCH(DLSA_FAMILY_LOGISTIC, 16, true)
    } else if (glm) {
        if (c.CH == 32) DLSA_TILED_LAUNCH(DLSA_FAMILY_LOGISTIC, 32, false) else DLSA_TILED_LAUNCH(DLSA_FAMILY_LOGISTIC, 16, false)
    } else {
        if (c.CH == 32) DLSA_TILED_LAUNCH(DLSA_FAMILY_LINEAR, 32, false) else DLSA_TILED_LAUNCH(DLSA_FAMILY_LINEAR, 16, false)
    }
#undef DLSA_TILED_LAUNCH
    count_launch();
    int rc = check_launch("irls_tiled_kernel");
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
