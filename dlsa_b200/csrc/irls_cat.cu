// K1c — the fused Newton / IRLS pass for a design matrix with one-hot (dummy) blocks, computed from the COMPACT
// representation [numeric columns | one integer code per categorical column] instead of the expanded dense X.
//
// Replaces, for the dummy branch of logistic_model (dlsa/models.py:56-91: pd.get_dummies -> dense 0/1 columns ->
// sklearn fit -> Sig_inv = X' diag(p(1-p)) X, :130), the same quantities as dlsa_irls_pass_f64 on the expanded X:
//   g = X'(y - prob),  H = X' diag(prob(1-prob)) X,  loglik.
// A row of the expanded X is z = [1 | x_1..x_q] (q <= 7 numerics) plus ONE 1 per categorical column.  The row stage
// (lane = row) leaves b = (r, w z_1, .., w z_7) per row in shared memory; then
//   H[z, z], g[z]      = sum z (x) b                        one DMMA per 4 rows: column 0 of the product is g[z], row 0 gives
//                                                           the intercept row by symmetry, sum w is a scalar accumulator
//   H[level, z_j],     = sum_{rows at that level} b          NARROW categoricals (<= 16 levels, <= 6 tiles of 8 levels in
//   g[level]                                                 total): one-hot(level)' b on DMMA — the FP64 pipe is idle here.
//                                                           WIDE categoricals: scatter-add into a WARP-PRIVATE shared-memory
//                                                           table, two categoricals per table, level l of the first in
//                                                           bytes 0..63 of row l and of the second in bytes 64..127: a
//                                                           half-warp (2 categoricals x 8 columns) then touches 16 different
//                                                           bank pairs whatever the levels are, and the two half-warps work
//                                                           on two consecutive rows at once (rows that share a level are
//                                                           applied one after the other) — no atomics, fixed summation order;
//                                                           baseline levels point at a sink row
//   H[level, level']   = sum w over rows holding both        (different categorical columns) — a weighted 2-D histogram:
//                                                           64-bit fixed-point RED.ADD into an L2-resident table whose
//                                                           (c, d) block has a sink row and a sink column for baselines
//   H[level, level]    = H[level, 1] = sum_j block(c, d)[level][j]   every row holds exactly one level (or the baseline) of the
//                                                           partner column d, so the row sums of a pair block — sinks
//                                                           included — are the level's own weight: no ninth scatter column
// Algorithmic traffic per row: 8 q + 4 c + 8 bytes (airline shape: 84 B instead of 1456 B).  Fixed-point sums are
// exact integer additions => the histogram part is bitwise independent of the order in which rows arrive; resolution
// 2^-44 (w <= 1/4) against sums of at most 2^20 rows per CTA.
//
// History (profiles/): v1 did every H[level, z] tile on DMMA with per-tile skipping — a predicated-off DMMA still holds
// the FP64 pipe for 16 cycles (r01).  v2 scattered nine columns of every categorical with lane = (categorical, column):
// 26 shared-memory wavefronts per row, LSU data pipe 70-80 % busy (r02_cat_c3_ncu_summary.csv).  This version: 8 columns,
// conflict-free by construction, two rows per instruction.
#include "common.cuh"

namespace dlsa {

namespace {

constexpr int kCatMaxC = 8;      // categorical columns
constexpr int kCatMaxNum = 7;    // numeric columns (z = [1 | numerics] fills one 8-wide tile)
constexpr int kCatMaxWarps = 16;
constexpr int kCatStages = 2;
constexpr int kCatMaxRowsPerCta = 1 << 20;
constexpr int kCatMaxTiles = 6;     // 8-level DMMA tiles for the narrow categoricals
constexpr int kTabPitch = 9;        // doubles per level in the PARTIALS: [unused | w z_1..7 | r]
constexpr double kWScale = 17592186044416.0;  // 2^44
constexpr int kIPitch = 40;                               // ints per row of the per-row integer scratch: 32 + 8 (see the tile loads)
constexpr int kBPitch = 10;                                // doubles per row of b: r | w z_1..7 | 2 pad (bank spread, see the DMMA stage)

struct CatArgs {
    const double* num;
    int64_t ldn;
    int n_num;
    const int* codes;
    int n_cat;
    int n_pair_cats;                     // max(n_cat, 2): a lone categorical gets an empty partner for its weight sums
    const double* mean;
    const double* sd;
    const double* y;
    const int64_t* part_off;
    int K, S;
    int fit_intercept, want_ll;
    const double* beta;
    const int* active;
    int Ltot, Lpad, npairs;
    int cat_lo[kCatMaxC + 1];            // level-space offsets of each categorical column
    int pair_off[kCatMaxC * kCatMaxC];   // offset of the (c, c') block, c < c', in the pair table
    int pair_ld[kCatMaxC];               // levels of the column + 1: cell (0, .) and (., 0) of a block are sinks
    int partner[kCatMaxC];               // the column whose pair block yields this column's level weights
    double* part_d;                      // [K*S, stride_d]: Dz 64 | gz 8 | ll 1 | pad to 80 | table [Lpad][9]
    unsigned long long* part_i;          // [K*S, stride_i]: pairs [npairs]
    int64_t stride_d, stride_i;
    int n_wide;                          // scattered categoricals; wide slot v = 2 * round + half
    int cat_wide[kCatMaxC];              // categorical -> wide slot, or -1 (narrow)
    int cat_row[kCatMaxC];               // categorical -> row of the per-row int scratch (wide slots first, then narrow)
    int cat_tab[kCatMaxC];               // wide: byte offset of level 0 of this categorical inside the warp's tables
    int cat_sink[kCatMaxC];              // wide: byte offset of its half of the round's sink row
    int tab_bytes;                       // per-warp tables, all rounds
    int share_tables;                    // 1: warps 2 i and 2 i + 1 take turns on one set of tables; 0: one set per warp
    int stage_bytes, front_bytes;        // one ring stage (numerics | codes | labels); ring + row scratch (tables follow)
    int ntiles;                          // DMMA level tiles in use
    // a tile is 8 SLOTS, each one (narrow categorical, code): the narrow levels are packed densely, categoricals mixed
    short slot_row[kCatMaxTiles * 8];    // row of the per-row int scratch holding the slot's categorical's codes
    short slot_code[kCatMaxTiles * 8];   // the code the slot stands for; -2 = unused slot
    short slot_lev[kCatMaxTiles * 8];    // its level-space index
};

__device__ __forceinline__ void cp_async4_zfill(uint32_t dst, const void* src, bool pred) {
    const int sz = pred ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async8_zfill(uint32_t dst, const void* src, bool pred) {
    const int sz = pred ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void red_add_u64(unsigned long long* p, long long v) {
    asm volatile("red.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

template <int N, int NC>
__device__ __forceinline__ void pair_reds(const CatArgs& a, unsigned long long* pairs, const int (&li)[NC], long long wfix) {
    if constexpr (N <= NC) {
#pragma unroll
        for (int c = 0; c < N; ++c) {
#pragma unroll
            for (int d = c + 1; d < N; ++d)
                red_add_u64(pairs + (unsigned)(a.pair_off[c * kCatMaxC + d] + li[c] * a.pair_ld[d] + li[d]), wfix);
        }
    }
}

// ROUNDS = ceil(wide categoricals / 2) (0..4).  NC = categorical columns compiled into the row stage (8).
// NTILE = compiled DMMA level tiles (0, 2, 4, 6).
template <int ROUNDS, int NC, int NTILE>
__global__ void __launch_bounds__(kCatMaxWarps * 32, 1) irls_cat_kernel(const CatArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int k = blockIdx.x / a.S;
    const int s = blockIdx.x - k * a.S;
    if (a.active != nullptr && a.active[k] == 0) return;
    const int nwarps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_num = a.n_num, n_cat = a.n_cat;

    // shared: per warp [ring | b rows | int rows]; the level tables, one set per warp or — when that buys at least four more
    // warps per SM (shared memory is what limits them, and a table is in use for a fifth of a warp's time) — one set per
    // warp PAIR, warps 2 i and 2 i + 1 taking turns (see the scatter stage); then beta (8 + Lpad), two mbarriers per pair
    const int stage_z = 256 * n_num, stage_c = 128 * n_cat;   // 32 rows of numerics, of codes (then 32 labels)
    const int tshift = a.share_tables ? 1 : 0;
    const int npair = (nwarps + tshift) >> tshift;   // sets of tables
    unsigned char* mine = smem_raw + (size_t)warp * a.front_bytes;
    unsigned char* ring = mine;
    double* bs = reinterpret_cast<double*>(mine + kCatStages * a.stage_bytes);                          // [32][10]
    int* irow = reinterpret_cast<int*>(mine + kCatStages * a.stage_bytes + 32 * kBPitch * 8);           // [n_cat][32]
    unsigned char* tab_all = smem_raw + (size_t)nwarps * a.front_bytes;
    unsigned char* tab = tab_all + (size_t)(warp >> tshift) * a.tab_bytes;
    double* bz = reinterpret_cast<double*>(tab_all + (size_t)npair * a.tab_bytes);
    double* bl = bz + 8;
    double* mean_s = bl + a.Lpad;   // [8] column means, [8] reciprocal standard deviations (standardisation on load)
    double* rsd_s = mean_s + 8;
    const uint32_t turn = smem_u32(rsd_s + 8) + 16u * (warp >> 1);   // [0]: the even warp is done, [1]: the odd warp is done
    const bool paired = a.share_tables && (warp | 1) < nwarps;         // the last warp of an odd count has its tables alone

    // coefficients in internal order: [intercept | numerics (7) | levels]; zero this warp's tables
    {
        const int p = a.fit_intercept + n_num + a.Ltot;
        const double* bk = a.beta + (size_t)k * p;
        for (int i = threadIdx.x; i < 8 + a.Lpad; i += blockDim.x) {
            double v = 0.0;
            if (i == 0) v = a.fit_intercept ? bk[0] : 0.0;
            else if (i < 8) v = (i - 1 < n_num) ? bk[a.fit_intercept + i - 1] : 0.0;
            else if (i - 8 < a.Ltot) v = bk[a.fit_intercept + n_num + (i - 8)];
            bz[i] = v;
        }
        if (threadIdx.x < 8) {
            const bool on = a.mean != nullptr && (int)threadIdx.x < n_num;
            mean_s[threadIdx.x] = on ? a.mean[threadIdx.x] : 0.0;
            rsd_s[threadIdx.x] = on ? 1.0 / a.sd[threadIdx.x] : 1.0;
        }
        double* t = reinterpret_cast<double*>(tab_all);
        for (int i = threadIdx.x; i < npair * (a.tab_bytes / 8); i += blockDim.x) t[i] = 0.0;
        if (lane == 0 && (warp & 1) == 0) {
            mbar_init(turn, 1);
            mbar_init(turn + 8, 1);
        }
    }
    __syncthreads();

    // this CTA's rows: slice s of partition k, in 32-row groups dealt round-robin to the warps
    const int64_t p0 = a.part_off[k], p1 = a.part_off[k + 1];
    const int64_t rows = p1 - p0;
    const int64_t per = ((rows + a.S - 1) / a.S + 31) / 32 * 32;
    const int64_t r0 = p0 + (int64_t)s * per;
    const int64_t r1 = (r0 + per < p1) ? r0 + per : p1;
    const int64_t ngroups = r1 > r0 ? (r1 - r0 + 31) / 32 : 0;
    const int64_t my_groups = ngroups > warp ? (ngroups - warp + nwarps - 1) / nwarps : 0;
    const int64_t mate_groups = ngroups > (warp ^ 1) ? (ngroups - (warp ^ 1) + nwarps - 1) / nwarps : 0;   // even >= odd

    unsigned long long* pairs = a.part_i + (size_t)blockIdx.x * a.stride_i;

    auto issue = [&](int64_t gi) {  // cp.async the gi-th group of this warp into stage gi % kCatStages
        if (gi < my_groups) {
            const int64_t R0 = r0 + (warp + gi * nwarps) * 32;
            const int nv = (int)((r1 - R0 < 32) ? (r1 - R0) : 32);
            unsigned char* st = ring + (size_t)(gi % kCatStages) * a.stage_bytes;
            const uint32_t zs = smem_u32(st), cs = zs + stage_z, ys = cs + stage_c;
            if (nv == 32 && a.ldn == n_num) {  // the common case: three contiguous runs, no predicates, one address each
                const char* zsrc = reinterpret_cast<const char*>(a.num + R0 * n_num) + 8 * lane;
                const char* csrc = reinterpret_cast<const char*>(a.codes + R0 * n_cat) + 4 * lane;
#pragma unroll
                for (int i = 0; i < kCatMaxNum; ++i)
                    if (i < n_num) cp_async8(zs + 8 * lane + 256 * i, zsrc + 256 * i);
#pragma unroll
                for (int i = 0; i < kCatMaxC; ++i)
                    if (i < n_cat) cp_async4(cs + 4 * lane + 128 * i, csrc + 128 * i);
                cp_async8(ys + 8 * lane, a.y + R0 + lane);
                cp_async_commit();
                return;
            }
            if (a.ldn == n_num) {
                const double* src = a.num + R0 * n_num;
                for (int i = 0; i < n_num; ++i) {
                    const int e = lane + 32 * i;
                    const bool ok = e < nv * n_num;
                    cp_async8_zfill(zs + 8 * e, src + (ok ? e : 0), ok);
                }
            } else {
                for (int i = 0; i < n_num; ++i) {
                    const int e = lane + 32 * i;
                    const int rr = e / n_num, cc = e - rr * n_num;
                    const bool ok = rr < nv;
                    cp_async8_zfill(zs + 8 * e, a.num + (R0 + (ok ? rr : 0)) * a.ldn + cc, ok);
                }
            }
            const int* csrc = a.codes + R0 * n_cat;
            for (int i = 0; i < n_cat; ++i) {
                const int e = lane + 32 * i;
                const bool ok = e < nv * n_cat;
                cp_async4_zfill(cs + 4 * e, csrc + (ok ? e : 0), ok);
            }
            cp_async8_zfill(ys + 8 * lane, a.y + R0 + (lane < nv ? lane : 0), lane < nv);
        }
        cp_async_commit();  // one (possibly empty) group per call keeps the wait_group arithmetic uniform
    };

    // scatter lanes: half-warp h works on row 2 rp + h; inside it lane = (which categorical of the round's two, column j);
    // column 0 is r (-> g[level]), column j >= 1 is w z_j
    const uint32_t bs_u32 = smem_u32(bs), irow_u32 = smem_u32(irow), tab_u32 = smem_u32(tab);
    const int half = lane >> 4, colj = lane & 7;
    uint32_t sc_to[ROUNDS > 0 ? ROUNDS : 1];   // shared address of this lane's categorical's offsets of rows 0, 1
    bool sc_on[ROUNDS > 0 ? ROUNDS : 1];
#pragma unroll
    for (int q = 0; q < ROUNDS; ++q) {
        const int v = 2 * q + ((lane >> 3) & 1);
        sc_on[q] = v < a.n_wide && colj <= n_num;
        sc_to[q] = irow_u32 + (uint32_t)(v < a.n_wide ? v : 0) * (kIPitch * 4u);
    }
    const uint32_t sc_b = bs_u32 + (uint32_t)half * (kBPitch * 8u) + 8u * colj;   // b[row 0 + half][j]
    // DMMA level tiles: lane (m = lane >> 2, k = lane & 3) holds one-hot[code code0 + m][row] of the A operand
    int tile_cmp[NTILE > 0 ? NTILE : 1];
    uint32_t tile_to[NTILE > 0 ? NTILE : 1];
    double td0[NTILE > 0 ? NTILE : 1], td1[NTILE > 0 ? NTILE : 1];
#pragma unroll
    for (int t = 0; t < NTILE; ++t) {
        const bool on = t < a.ntiles;
        // slots of one tile may belong to different categoricals: with 40 ints per scratch row the 4 rows x up to 4
        // categoricals a tile load touches fall into different banks (8 c + row)
        tile_cmp[t] = on ? (int)a.slot_code[t * 8 + (lane >> 2)] : -2;   // codes are >= -1: -2 never matches
        tile_to[t] = irow_u32 + (uint32_t)(on ? (int)a.slot_row[t * 8 + (lane >> 2)] : 0) * (kIPitch * 4u) + 8u * (lane & 3);
        td0[t] = 0.0;
        td1[t] = 0.0;
    }

    double dz0 = 0.0, dz1 = 0.0, ll = 0.0, sw = 0.0;

    for (int64_t gi = 0; gi < kCatStages - 1; ++gi) issue(gi);

    const int jcol = lane >> 2;  // the z column this lane holds in the MMA fragments
    for (int64_t gi = 0; gi < my_groups; ++gi) {
        issue(gi + kCatStages - 1);
        cp_async_wait<kCatStages - 1>();
        __syncwarp();
        const int64_t R0 = r0 + (warp + gi * nwarps) * 32;
        const int nv = (int)((r1 - R0 < 32) ? (r1 - R0) : 32);
        unsigned char* st = ring + (size_t)(gi % kCatStages) * a.stage_bytes;
        double* zs = reinterpret_cast<double*>(st);
        const int* cs = reinterpret_cast<const int*>(st + stage_z);
        const double* ys = reinterpret_cast<const double*>(st + stage_z + stage_c);

        // ---- row stage: lane = row -----------------------------------------------------------------------
        const bool valid = lane < nv;
        double eta = bz[0];
        double z[kCatMaxNum];
        if (a.mean != nullptr) {   // one warp-uniform branch around the whole loop: inside it ptxas predicates the extra
                                   // FP64 instructions, and a predicated-off FP64 instruction still takes its pipe slot
#pragma unroll
            for (int j = 0; j < kCatMaxNum; ++j) {
                z[j] = 0.0;
                if (j < n_num) {
                    // (x - mean) / std of dlsa/models.py:99-101 with the reciprocal taken once per CTA (<= 1 ulp from the
                    // division; 7 FP64 divisions per row and pass otherwise)
                    const double v = (zs[lane * n_num + j] - mean_s[j]) * rsd_s[j];
                    zs[lane * n_num + j] = v;
                    z[j] = v;
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < kCatMaxNum; ++j) z[j] = (j < n_num) ? zs[lane * n_num + j] : 0.0;
        }
#pragma unroll
        for (int j = 0; j < kCatMaxNum; ++j) eta = fma(z[j], bz[1 + j], eta);   // bz is 0 beyond n_num
        int li[NC];   // 1 + level within the column; 0 = baseline level, no such column, or no such row
        {
            // all loads first (the level coefficients are gathered independently, then summed), stores afterwards: the
            // compiler cannot move a shared load across a shared store on its own
            double bv[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                li[c] = 0;
                bv[c] = 0.0;
                if (c < n_cat) {
                    const int code = valid ? cs[lane * n_cat + c] : -1;
                    if (code >= 0) {
                        li[c] = code + 1;
                        bv[c] = bl[a.cat_lo[c] + code];
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < NC; ++c) eta += bv[c];
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                if (c < n_cat) {
                    // wide: byte offset of the level's half row in the warp's tables; narrow: the code itself
                    const int code = li[c] - 1;
                    const int wide_to = code >= 0 ? a.cat_tab[c] + code * 128 : a.cat_sink[c];
                    irow[a.cat_row[c] * kIPitch + lane] = a.cat_wide[c] >= 0 ? wide_to : code;
                }
            }
        }
        double w = 0.0, r = 0.0;
        if (valid) {
            const double yy = ys[lane];
            const double prob = sigmoid_f64(eta);
            w = prob * (1.0 - prob);
            r = yy - prob;
            if (a.want_ll) ll += yy * eta - log1pexp_f64(eta);
        }
        sw += w;
        {   // b = (r, w z_1..7) of this row
            double2* brow = reinterpret_cast<double2*>(bs + lane * kBPitch);
            brow[0] = make_double2(r, w * z[0]);
            brow[1] = make_double2(w * z[1], w * z[2]);
            brow[2] = make_double2(w * z[3], w * z[4]);
            brow[3] = make_double2(w * z[5], w * z[6]);
        }
        {   // weighted pair histogram: every (c, d) block carries a sink row and a sink column (index 0) that take the rows
            // with a baseline level, so the cell index is two multiply-adds and the REDs are unconditional
            const long long wfix = __double2ll_rn(w * kWScale);
            switch (a.n_pair_cats) {
                case 2: pair_reds<2, NC>(a, pairs, li, wfix); break;
                case 3: pair_reds<3, NC>(a, pairs, li, wfix); break;
                case 4: pair_reds<4, NC>(a, pairs, li, wfix); break;
                case 5: pair_reds<5, NC>(a, pairs, li, wfix); break;
                case 6: pair_reds<6, NC>(a, pairs, li, wfix); break;
                case 7: pair_reds<7, NC>(a, pairs, li, wfix); break;
                case 8: pair_reds<8, NC>(a, pairs, li, wfix); break;
                default: break;
            }
        }
        __syncwarp();

        // ---- dense z block: 8 k-steps of 4 rows, one DMMA each: A = z, B = b; the narrow categoricals' tiles reuse B
        // with A = one-hot(level).  The 4 rows of a k-step are 2 apart: with the 10-double pitch of b a half-warp's 16
        // operands then fall into 16 different bank pairs (consecutive rows collide: r02 ncu, 4 wavefronts instead of 2).
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            const int base = 8 * (ks >> 1) + (ks & 1);
            const int src = base + 2 * (lane & 3);
            double zz = 1.0;
            if (jcol > 0) zz = (jcol - 1 < n_num) ? zs[src * n_num + jcol - 1] : 0.0;
            const double bw = bs[src * kBPitch + jcol];
            dmma884(dz0, dz1, zz, bw);
#pragma unroll
            for (int t = 0; t < NTILE; ++t) {
                int o;
                asm volatile("ld.shared.s32 %0, [%1];" : "=r"(o) : "r"(tile_to[t] + (uint32_t)base * 4u));
                dmma884(td0[t], td1[t], (o == tile_cmp[t]) ? 1.0 : 0.0, bw);
            }
        }

        // ---- wide categoricals: scatter-add b into the warp's tables, two rows per step ---------------------------
        if (ROUNDS > 0) {
            // turns on the pair's tables: even warp's group gi, odd warp's group gi, even warp's group gi + 1, ... — a fixed
            // order, so the sums do not depend on timing
            if (paired) {
                if (warp & 1) mbar_wait(turn, (uint32_t)gi & 1u);
                else if (gi > 0 && gi - 1 < mate_groups) mbar_wait(turn + 8, (uint32_t)(gi - 1) & 1u);
            }
            int2 to[ROUNDS > 0 ? ROUNDS : 1];
            double xb;
            auto load_rows = [&](int rp) {   // offsets of rows 2 rp, 2 rp + 1 for this lane's categoricals; this lane's addend
#pragma unroll
                for (int q = 0; q < ROUNDS; ++q)
                    asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(to[q].x), "=r"(to[q].y) : "r"(sc_to[q] + 8u * rp));
                asm volatile("ld.shared.f64 %0, [%1];" : "=d"(xb) : "r"(sc_b + (uint32_t)rp * (2u * kBPitch * 8u)));
            };
            load_rows(0);
#pragma unroll
            for (int rp = 0; rp < 16; ++rp) {
                uint32_t addr[ROUNDS > 0 ? ROUNDS : 1];
                bool clash = false;
#pragma unroll
                for (int q = 0; q < ROUNDS; ++q) {
                    addr[q] = tab_u32 + (uint32_t)(half ? to[q].y : to[q].x) + 8u * colj;
                    clash = clash || (sc_on[q] && to[q].x == to[q].y);
                }
                const double x = xb;
                if (!__any_sync(0xffffffffu, clash)) {
                    double t[ROUNDS > 0 ? ROUNDS : 1];
#pragma unroll
                    for (int q = 0; q < ROUNDS; ++q)
                        if (sc_on[q]) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(t[q]) : "r"(addr[q]) : "memory");
                    if (rp + 1 < 16) load_rows(rp + 1);   // the row scratch is not written inside this loop
#pragma unroll
                    for (int q = 0; q < ROUNDS; ++q)
                        if (sc_on[q]) asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr[q]), "d"(t[q] + x) : "memory");
                } else {   // the two rows share a level somewhere: first row, then second row
                    if (rp + 1 < 16) load_rows(rp + 1);
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        if (half == hh) {
#pragma unroll
                            for (int q = 0; q < ROUNDS; ++q) {
                                if (sc_on[q]) {
                                    double t;
                                    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(t) : "r"(addr[q]) : "memory");
                                    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr[q]), "d"(t + x) : "memory");
                                }
                            }
                        }
                        __syncwarp();
                    }
                }
                __syncwarp();   // the next step's rows may hit the levels other lanes have just written
            }
            if (paired && lane == 0) {
                if (warp & 1) { if (gi + 1 < mate_groups) mbar_arrive(turn + 8); }
                else if (gi < mate_groups) mbar_arrive(turn);
            }
        }
        __syncwarp();  // the stage and the row scratch are re-filled by the next iteration
    }
    cp_async_wait<0>();

    // ---- CTA reduction in a fixed warp order, straight into the workspace partial ----------------------------------
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) ll += shfl_xor_f64(ll, m);
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) sw += shfl_xor_f64(sw, m);
    __syncthreads();  // every warp is done with its ring: reuse it for the small per-warp results
    {
        // Dz 64 (row i = z_i, column 0 = g[z_i], column j >= 1 = H[z_i, z_j]) | ll at 72 | tiles [ntiles][8][8] from 80
        double* sm = reinterpret_cast<double*>(ring);
        const int di = lane >> 2, dj = 2 * (lane & 3);
        sm[di * 8 + dj] = dz0;
        sm[di * 8 + dj + 1] = dz1;
        if (lane == 8) sm[72] = ll;
        if (lane == 9) sm[73] = sw;
#pragma unroll
        for (int t = 0; t < NTILE; ++t) {
            sm[80 + t * 64 + di * 8 + dj] = td0[t];
            sm[80 + t * 64 + di * 8 + dj + 1] = td1[t];
        }
    }
    __syncthreads();
    double* out = a.part_d + (size_t)blockIdx.x * a.stride_d;
    for (int i = threadIdx.x; i < 80 + a.Lpad * kTabPitch; i += blockDim.x) {
        // where entry i of the partial lives in a warp's block: a double index, or -1 for a structural zero
        int srci = -1;
        bool in_tab = false;
        if (i < 64) {
            const int u1 = i >> 3, u2 = i & 7;            // the assembly reads u1 <= u2 only
            if (i == 0) srci = 73;                        // sum w
            else if (u2 >= 1 && u1 <= u2) srci = u1 * 8 + u2;
        } else if (i < 72) {
            srci = (i - 64) * 8;                          // g[z_i] = column 0
        } else if (i == 72) {
            srci = 72;
        } else if (i >= 80) {
            const int level = (i - 80) / kTabPitch, col = (i - 80) - level * kTabPitch;
            if (level < a.Ltot && col >= 1 && (col == 8 || col <= n_num)) {
                int c = 0;
                while (c + 1 < n_cat && level >= a.cat_lo[c + 1]) ++c;
                const int code = level - a.cat_lo[c], jj = col == 8 ? 0 : col;
                if (a.cat_wide[c] >= 0) {
                    srci = (a.cat_tab[c] + code * 128) / 8 + jj;
                    in_tab = true;
                } else {
                    for (int sl = 0; sl < a.ntiles * 8; ++sl)
                        if (a.slot_code[sl] >= 0 && a.slot_lev[sl] == level) srci = 80 + sl * 8 + jj;
                }
            }
        }
        double v = 0.0;
        if (srci >= 0) {
            if (!in_tab)
                for (int wv = 0; wv < nwarps; ++wv)
                    v += reinterpret_cast<const double*>(smem_raw + (size_t)wv * a.front_bytes)[srci];
            else
                for (int pr = 0; pr < npair; ++pr) v += reinterpret_cast<const double*>(tab_all + (size_t)pr * a.tab_bytes)[srci];
        }
        out[i] = v;
    }
}

// Assemble the full symmetric H [p, p], g [p] and loglik of every active partition from its S partials, slices summed in a
// fixed order.  Output order = the expanded design: [intercept] numerics, then the level columns (dlsa/models.py:71-79).
__global__ void irls_cat_finalize_kernel(const CatArgs a, double* __restrict__ g, double* __restrict__ H,
                                         double* __restrict__ loglik) {
    const int k = blockIdx.x;
    if (a.active != nullptr && a.active[k] == 0) return;
    const int q = a.n_num, ic = a.fit_intercept;
    const int p = ic + q + a.Ltot;
    const double* pd = a.part_d + (size_t)k * a.S * a.stride_d;
    const unsigned long long* pi = a.part_i + (size_t)k * a.S * a.stride_i;
    auto internal = [&](int o) {  // output coefficient index -> internal index (0 = intercept, 1..7 numerics, 8.. levels)
        if (ic) {
            if (o == 0) return 0;
            return (o - 1 < q) ? o : 8 + (o - 1 - q);
        }
        return (o < q) ? o + 1 : 8 + (o - q);
    };
    auto cat_of = [&](int level) {
        int c = 0;
        while (c + 1 < a.n_cat && level >= a.cat_lo[c + 1]) ++c;
        return c;
    };
    // sum of w over the rows at a level = the level's row (or column) of the pair block with its partner column, sinks
    // included; exact integer sums per slice, converted per slice (a slice sum is < 2^63, the partition sum may not be)
    auto level_weight = [&](int level) {
        const int c = cat_of(level), d = a.partner[c];
        const int code1 = level - a.cat_lo[c] + 1;
        double v = 0.0;
        for (int sl = 0; sl < a.S; ++sl) {
            const unsigned long long* ps = pi + sl * a.stride_i;
            long long acc = 0;
            if (c < d) {
                const unsigned long long* row = ps + a.pair_off[c * kCatMaxC + d] + code1 * a.pair_ld[d];
                for (int j = 0; j < a.pair_ld[d]; ++j) acc += (long long)row[j];
            } else {
                const unsigned long long* col = ps + a.pair_off[d * kCatMaxC + c] + code1;
                for (int j = 0; j < a.pair_ld[d]; ++j) acc += (long long)col[j * a.pair_ld[c]];
            }
            v += (double)acc * (1.0 / kWScale);
        }
        return v;
    };
    double* Hk = H + (size_t)k * p * p;
    for (int e = threadIdx.x; e < p * p; e += blockDim.x) {
        const int o1 = e / p, o2 = e - o1 * p;
        if (o2 > o1) continue;
        int u1 = internal(o1), u2 = internal(o2);
        if (u1 > u2) { const int t = u1; u1 = u2; u2 = t; }
        double v = 0.0;
        if (u2 < 8) {
            for (int sl = 0; sl < a.S; ++sl) v += pd[sl * a.stride_d + u1 * 8 + u2];
        } else if (u1 == 0) {
            v = level_weight(u2 - 8);
        } else if (u1 < 8) {
            for (int sl = 0; sl < a.S; ++sl) v += pd[sl * a.stride_d + 80 + (u2 - 8) * kTabPitch + u1];
        } else {
            const int l1 = u1 - 8, l2 = u2 - 8;
            if (l1 == l2) {
                v = level_weight(l1);
            } else {
                const int c1 = cat_of(l1), c2 = cat_of(l2);
                if (c1 != c2) {
                    const int idx = a.pair_off[c1 * kCatMaxC + c2] + (l1 - a.cat_lo[c1] + 1) * a.pair_ld[c2] +
                                    (l2 - a.cat_lo[c2] + 1);
                    for (int sl = 0; sl < a.S; ++sl)  // per-slice conversion, as above
                        v += (double)(long long)pi[sl * a.stride_i + idx] * (1.0 / kWScale);
                }
            }
        }
        Hk[o1 * p + o2] = v;
        Hk[o2 * p + o1] = v;
    }
    for (int o = threadIdx.x; o < p; o += blockDim.x) {
        const int u = internal(o);
        double v = 0.0;
        if (u < 8) {
            for (int sl = 0; sl < a.S; ++sl) v += pd[sl * a.stride_d + 64 + u];
        } else {
            for (int sl = 0; sl < a.S; ++sl) v += pd[sl * a.stride_d + 80 + (u - 8) * kTabPitch + 8];
        }
        g[(size_t)k * p + o] = v;
    }
    if (loglik != nullptr && threadIdx.x == 0) {
        double v = 0.0;
        for (int sl = 0; sl < a.S; ++sl) v += pd[sl * a.stride_d + 72];
        loglik[k] = v;
    }
}

struct CatPlan {
    int Ltot, Lpad, npairs, S, warps;
    int64_t stride_d, stride_i;
    size_t bytes_d, bytes_i, smem;
    // layout of the level work (see CatArgs)
    int n_wide, rounds, ntiles, tab_bytes, stage_bytes, front_bytes, share_tables;
    int cat_wide[kCatMaxC], cat_row[kCatMaxC], cat_tab[kCatMaxC], cat_sink[kCatMaxC];
    short slot_row[kCatMaxTiles * 8], slot_code[kCatMaxTiles * 8], slot_lev[kCatMaxTiles * 8];
};

int cat_sm_count() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
            sms <= 0)
            sms = 148;
    }
    return sms;
}

void cat_plan(int K, int n_num, int n_cat, const int* n_levels, int64_t max_part_rows, CatPlan& pl) {
    pl.Ltot = 0;
    pl.npairs = 0;
    int cat_lo[kCatMaxC];
    for (int c = 0; c < n_cat; ++c) {
        for (int d = 0; d < c; ++d) pl.npairs += (n_levels[d] + 1) * (n_levels[c] + 1);  // + sink row and column
        cat_lo[c] = pl.Ltot;
        pl.Ltot += n_levels[c];
    }
    if (n_cat == 1) pl.npairs = n_levels[0] + 1;  // the block with the empty partner column
    pl.Lpad = (pl.Ltot + 7) / 8 * 8;
    const int sms = cat_sm_count();
    // slices per partition: one CTA per SM at a time, so K * S should fill whole waves of `sms` CTAs (K = 120, S = 1 leaves
    // 19 % of the SMs idle; S = 6 gives 4.86 waves = 97 %).  The smallest S wins unless a larger one fills the waves 6 % better;
    // every slice keeps at least 4096 rows and S <= 16 bounds the partial buffers.
    const int64_t s_rows = (max_part_rows + 4095) / 4096;
    int64_t s_cap = K <= 16 ? (2 * sms + K - 1) / K : 16;
    if (s_cap > s_rows) s_cap = s_rows;
    if (s_cap < 1) s_cap = 1;
    double best = -1.0;
    int64_t S = 1;
    for (int64_t cand = 1; cand <= s_cap; ++cand) {
        const int64_t u = (int64_t)K * cand;
        const double eff = (double)u / (double)(((u + sms - 1) / sms) * sms);
        if (eff > best + 0.06) { best = eff; S = cand; }
    }
    const int64_t s_min = (max_part_rows + kCatMaxRowsPerCta - 1) / kCatMaxRowsPerCta;  // fixed-point headroom
    if (S < s_min) S = s_min;
    if (S < 1) S = 1;
    pl.S = (int)S;
    pl.stride_d = 80 + (int64_t)pl.Lpad * kTabPitch;
    pl.stride_i = pl.npairs;
    pl.bytes_d = (size_t)K * S * pl.stride_d * 8;
    pl.bytes_i = (size_t)K * S * pl.stride_i * 8;

    // narrow categoricals (<= 16 levels, narrowest first, <= tile_budget tiles in total) go to DMMA tiles, the others are
    // scattered, two per table: widest with second widest, and so on (a table has max(L, L') + 1 rows of 128 bytes)
    static const int tile_budget = [] {
        const char* e = getenv("DLSA_CAT_DMMA_TILES");
        int v = e ? atoi(e) : kCatMaxTiles;
        return v < 0 ? 0 : (v > kCatMaxTiles ? kCatMaxTiles : v);
    }();
    bool narrow[kCatMaxC] = {};
    int n_narrow = 0, n_slots = 0;
    for (;;) {
        int pick = -1;
        for (int c = 0; c < n_cat; ++c)
            if (!narrow[c] && n_levels[c] >= 1 && n_levels[c] <= 16 && (pick < 0 || n_levels[c] < n_levels[pick])) pick = c;
        if (pick < 0 || n_slots + n_levels[pick] > 8 * tile_budget) break;
        narrow[pick] = true;
        ++n_narrow;
        n_slots += n_levels[pick];
    }
    pl.ntiles = (n_slots + 7) / 8;
    int wide[kCatMaxC];
    pl.n_wide = 0;
    for (int c = 0; c < n_cat; ++c)
        if (!narrow[c]) wide[pl.n_wide++] = c;
    for (int i = 1; i < pl.n_wide; ++i)  // insertion sort, widest first
        for (int j = i; j > 0 && n_levels[wide[j]] > n_levels[wide[j - 1]]; --j) {
            const int t = wide[j]; wide[j] = wide[j - 1]; wide[j - 1] = t;
        }
    pl.rounds = (pl.n_wide + 1) / 2;
    pl.tab_bytes = 0;
    for (int c = 0; c < kCatMaxC; ++c) pl.cat_wide[c] = -1, pl.cat_row[c] = 0, pl.cat_tab[c] = 0, pl.cat_sink[c] = 0;
    for (int v = 0; v < pl.n_wide; ++v) {
        const int c = wide[v], rq = v >> 1, which = v & 1;
        const int rows = n_levels[wide[2 * rq]] + 1;   // the round's wider column + the sink row
        if (which == 0 && v > 0) pl.tab_bytes += (n_levels[wide[2 * (rq - 1)]] + 1) * 128;
        pl.cat_wide[c] = v;
        pl.cat_row[c] = v;
        pl.cat_tab[c] = pl.tab_bytes + which * 64;
        pl.cat_sink[c] = pl.tab_bytes + (rows - 1) * 128 + which * 64;
    }
    if (pl.n_wide > 0) pl.tab_bytes += (n_levels[wide[2 * (pl.rounds - 1)]] + 1) * 128;
    int sl = 0, nrow = pl.n_wide;
    for (int c = 0; c < n_cat; ++c) {
        if (!narrow[c]) continue;
        pl.cat_row[c] = nrow;
        for (int code = 0; code < n_levels[c]; ++code, ++sl) {
            pl.slot_row[sl] = (short)nrow;
            pl.slot_code[sl] = (short)code;
            pl.slot_lev[sl] = (short)(cat_lo[c] + code);
        }
        ++nrow;
    }
    for (; sl < kCatMaxTiles * 8; ++sl) pl.slot_row[sl] = 0, pl.slot_code[sl] = -2, pl.slot_lev[sl] = -1;

    // as many warps as shared memory allows (each owns a ring, a row scratch and its level tables)
    // (the ring + row scratch also stage the warp's small results at the end: 80 + 64 per tile doubles)
    pl.stage_bytes = 256 * n_num + 128 * n_cat + 256;
    pl.front_bytes = kCatStages * pl.stage_bytes + 32 * kBPitch * 8 + kIPitch * 4 * n_cat;
    if (pl.front_bytes < (80 + 64 * kCatMaxTiles) * 8) pl.front_bytes = (80 + 64 * kCatMaxTiles) * 8;
    const size_t fixed = (size_t)(8 + pl.Lpad + 16) * 8 + 16 * (kCatMaxWarps / 2);   // beta | mean, 1/std | mbarriers
    const size_t budget = 227 * 1024;
    auto smem_for = [&](int w, int share) {
        return (size_t)w * pl.front_bytes + (size_t)(share ? (w + 1) / 2 : w) * pl.tab_bytes + fixed;
    };
    int w_own = kCatMaxWarps, w_pair = kCatMaxWarps;
    while (w_own > 1 && smem_for(w_own, 0) > budget) --w_own;
    while (w_pair > 1 && smem_for(w_pair, 1) > budget) --w_pair;
    // measured on the airline shape (r02): 12 warps with own tables 9.98 ms, 16 warps in pairs 9.54 ms, 14 in pairs 10.5 ms —
    // taking turns costs about as much as two warps
    static const int share_env = [] {
        const char* e = getenv("DLSA_CAT_SHARE_TABLES");
        return e ? atoi(e) : -1;
    }();
    pl.share_tables = share_env >= 0 ? (share_env != 0) : (pl.tab_bytes > 0 && w_pair >= w_own + 4);
    int warps = pl.share_tables ? w_pair : w_own;
    static const int warp_cap = [] {
        const char* e = getenv("DLSA_CAT_WARPS");
        return e ? atoi(e) : kCatMaxWarps;
    }();
    if (warps > warp_cap && warp_cap >= 4) warps = warp_cap;
    pl.warps = warps;
    pl.smem = smem_for(warps, pl.share_tables);
}

}  // namespace

size_t irls_cat_workspace_bytes(int K, int n_cat, const int* n_levels, int64_t max_part_rows) {
    CatPlan pl;
    cat_plan(K, 0, n_cat, n_levels, max_part_rows, pl);   // the partial buffers do not depend on the numeric columns
    return pl.bytes_d + pl.bytes_i + 256;
}

// Host-only: the layout cat_plan chooses, for tests.  info[12]: slices per partition, warps, tables shared by warp pairs,
// wide categoricals, scatter rounds, DMMA tiles, table bytes (one set), ring stage bytes, front bytes, dynamic shared
// memory, pair cells, padded levels.  cat_out [n_cat][4]: wide slot (-1 = narrow), scratch row, byte offset of level 0 in
// the tables, byte offset of the sink.  tile_out [48][3]: per tile slot its scratch row, code (-2 = unused), level (level space).
int irls_cat_plan_info(int K, int n_num, int n_cat, const int* n_levels, int64_t max_part_rows, int* info, int* cat_out,
                       int* tile_out) {
    if (n_cat < 1 || n_cat > kCatMaxC || n_num < 0 || n_num > kCatMaxNum || K < 1) return -1;
    CatPlan pl;
    cat_plan(K, n_num, n_cat, n_levels, max_part_rows, pl);
    const int v[12] = {pl.S, pl.warps, pl.share_tables, pl.n_wide, pl.rounds, pl.ntiles, pl.tab_bytes, pl.stage_bytes,
                       pl.front_bytes, (int)pl.smem, pl.npairs, pl.Lpad};
    for (int i = 0; i < 12; ++i) info[i] = v[i];
    for (int c = 0; c < n_cat; ++c) {
        cat_out[4 * c + 0] = pl.cat_wide[c];
        cat_out[4 * c + 1] = pl.cat_row[c];
        cat_out[4 * c + 2] = pl.cat_tab[c];
        cat_out[4 * c + 3] = pl.cat_sink[c];
    }
    for (int sl = 0; sl < kCatMaxTiles * 8; ++sl) {
        tile_out[3 * sl + 0] = pl.slot_row[sl];
        tile_out[3 * sl + 1] = pl.slot_code[sl];
        tile_out[3 * sl + 2] = pl.slot_lev[sl];
    }
    return 0;
}

template <int ROUNDS, int NC, int NTILE>
static int cat_launch_t(const CatArgs& a, const CatPlan& pl, cudaStream_t st) {
    auto kern = irls_cat_kernel<ROUNDS, NC, NTILE>;
    DLSA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    kern<<<a.K * a.S, pl.warps * 32, pl.smem, st>>>(a);
    count_launch();
    return check_launch("irls_cat_kernel");
}

int irls_cat_launch(const double* numeric, int64_t ldn, int n_num, const double* mean, const double* sd, const int* codes,
                    int n_cat, const int* n_levels, const double* y, const int64_t* part_off, int K, int fit_intercept,
                    const double* beta, const int* active, double* g, double* H, double* loglik, int64_t max_part_rows,
                    void* workspace, size_t workspace_bytes, cudaStream_t st) {
    CatPlan pl;
    cat_plan(K, n_num, n_cat, n_levels, max_part_rows, pl);
    DLSA_REQUIRE(pl.Ltot <= 192, DLSA_E_UNSUPPORTED, "categorical pass: %d levels in total (max 192)", pl.Ltot);
    DLSA_REQUIRE(pl.warps >= 4, DLSA_E_UNSUPPORTED, "categorical pass: level tables do not fit shared memory");
    const size_t need = pl.bytes_d + pl.bytes_i + 256;
    DLSA_REQUIRE(workspace != nullptr && workspace_bytes >= need, DLSA_E_WORKSPACE,
                 "categorical pass: workspace %zu < %zu bytes", workspace_bytes, need);
    CatArgs a{};
    a.num = numeric;
    a.ldn = ldn;
    a.n_num = n_num;
    a.codes = codes;
    a.n_cat = n_cat;
    a.n_pair_cats = n_cat < 2 ? 2 : n_cat;
    a.mean = mean;
    a.sd = sd;
    a.y = y;
    a.part_off = part_off;
    a.K = K;
    a.S = pl.S;
    a.fit_intercept = fit_intercept ? 1 : 0;
    a.want_ll = loglik != nullptr;
    a.beta = beta;
    a.active = active;
    a.Ltot = pl.Ltot;
    a.Lpad = pl.Lpad;
    a.npairs = pl.npairs;
    int lo = 0, off = 0;
    for (int c = 0; c < kCatMaxC; ++c) a.pair_ld[c] = 1;
    for (int c = 0; c < n_cat; ++c) {
        a.cat_lo[c] = lo;
        for (int d = 0; d < c; ++d) {
            a.pair_off[d * kCatMaxC + c] = off;
            off += (n_levels[d] + 1) * (n_levels[c] + 1);
        }
        a.pair_ld[c] = n_levels[c] + 1;
        lo += n_levels[c];
    }
    for (int c = n_cat; c <= kCatMaxC; ++c) a.cat_lo[c] = lo;
    for (int c = 0; c < n_cat; ++c) {  // partner = the other column with the fewest levels (a lone column: the empty one)
        int d = -1;
        for (int e = 0; e < n_cat; ++e)
            if (e != c && (d < 0 || n_levels[e] < n_levels[d])) d = e;
        a.partner[c] = d < 0 ? 1 : d;
    }
    const uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255);
    a.part_d = reinterpret_cast<double*>(base);
    a.part_i = reinterpret_cast<unsigned long long*>(base + pl.bytes_d);
    a.stride_d = pl.stride_d;
    a.stride_i = pl.stride_i;
    DLSA_CUDA(cudaMemsetAsync(a.part_i, 0, pl.bytes_i, st));
    a.n_wide = pl.n_wide;
    a.tab_bytes = pl.tab_bytes;
    a.share_tables = pl.share_tables;
    a.stage_bytes = pl.stage_bytes;
    a.front_bytes = pl.front_bytes;
    a.ntiles = pl.ntiles;
    for (int c = 0; c < kCatMaxC; ++c) {
        a.cat_wide[c] = pl.cat_wide[c];
        a.cat_row[c] = pl.cat_row[c];
        a.cat_tab[c] = pl.cat_tab[c];
        a.cat_sink[c] = pl.cat_sink[c];
    }
    for (int sl = 0; sl < kCatMaxTiles * 8; ++sl) {
        a.slot_row[sl] = pl.slot_row[sl];
        a.slot_code[sl] = pl.slot_code[sl];
        a.slot_lev[sl] = pl.slot_lev[sl];
    }
    const int rounds = pl.rounds;              // 0..4
    const int ntile = pl.ntiles == 5 ? 6 : pl.ntiles;    // 0, 1, 2, 3, 4, 6
    int rc = DLSA_E_UNSUPPORTED;
#define DLSA_CAT_CASE8(R, T) \
    if (rounds == R && ntile == T) rc = cat_launch_t<R, 8, T>(a, pl, st);
#define DLSA_CAT_ROW8(R) DLSA_CAT_CASE8(R, 0) DLSA_CAT_CASE8(R, 1) DLSA_CAT_CASE8(R, 2) DLSA_CAT_CASE8(R, 3) DLSA_CAT_CASE8(R, 4) DLSA_CAT_CASE8(R, 6)
    DLSA_CAT_CASE8(0, 1) DLSA_CAT_CASE8(0, 2) DLSA_CAT_CASE8(0, 3) DLSA_CAT_CASE8(0, 4) DLSA_CAT_CASE8(0, 6)
    DLSA_CAT_ROW8(1) DLSA_CAT_ROW8(2) DLSA_CAT_ROW8(3) DLSA_CAT_ROW8(4)
#undef DLSA_CAT_ROW8
#undef DLSA_CAT_CASE8
    DLSA_REQUIRE(rc != DLSA_E_UNSUPPORTED, DLSA_E_UNSUPPORTED, "categorical pass: no kernel for %d scatter rounds x %d tiles",
                 rounds, ntile);
    if (rc != 0) return rc;
    irls_cat_finalize_kernel<<<K, 256, 0, st>>>(a, g, H, loglik);
    count_launch();
    return check_launch("irls_cat_finalize_kernel");
}

}  // namespace dlsa
