// K1c — the fused Newton / IRLS pass for a design matrix with one-hot (dummy) blocks, computed from the COMPACT
// representation [numeric columns | one integer code per categorical column] instead of the expanded dense X.
//
// Replaces, for the dummy branch of logistic_model (dlsa/models.py:56-91: pd.get_dummies -> dense 0/1 columns ->
// sklearn fit -> Sig_inv = X' diag(p(1-p)) X, :130), the same quantities as dlsa_irls_pass_f64 on the expanded X:
//   g = X'(y - prob),  H = X' diag(prob(1-prob)) X,  loglik.
// A row of the expanded X is z = [1 | x_1..x_q] (q <= 7 numerics) plus ONE 1 per categorical column, so
//   H[z, z]           = sum w z z'                         one DMMA per 4 rows (8x8 tile)
//   H[level, z]       = sum_{rows at that level} w z'      scatter-add into a WARP-PRIVATE shared-memory table
//   g[level]          = sum_{rows at that level} r          [level][9]: lane = (categorical, column); the row stage leaves
//                                                          (w z', r) and the table offset of every (row, categorical) in
//                                                          shared memory, so a row pair costs 4 loads + 2 adds + 2 stores
//                                                          per lane, the second row's sum forwarded in registers when both
//                                                          hit the same level — no atomics, fixed summation order; codes
//                                                          of baseline levels point at a sink row (no predicates)
//                                                          Categoricals with at most 16 levels (up to 6 tiles of 8 levels
//                                                          in total) skip the scatter: their H[level, z] block is
//                                                          one-hot(level)' (w z') on DMMA, 8 levels per tile, the one-hot
//                                                          operand built from the row's table offset by one compare —
//                                                          the FP64 pipe is idle here and the shared-memory pipe is the
//                                                          bound (r02: 26 wavefronts per row before, profiles/).  Only
//                                                          their g column stays in the scatter.
//   H[level, level]   = H[level, 1]  (the intercept column of the block above)
//   H[level, level']  = sum w over rows holding both       (different categorical columns) — a weighted 2-D histogram:
//                                                          64-bit fixed-point RED.ADD into an L2-resident table
// Algorithmic traffic per row: 8 q + 4 c + 8 bytes (airline shape: 84 B instead of 1456 B).  Fixed-point sums are
// exact integer additions => the histogram part is bitwise independent of the order in which rows arrive; resolution
// 2^-44 (w <= 1/4) against sums of at most 2^20 rows per CTA.
//
// v1 of this kernel did H[level, z] on DMMA with a one-hot operand and per-tile skipping (profiles/r01_k1c_*_v1.*):
// ptxas predicates the skipped DMMAs and a predicated-off DMMA still holds the FP64 pipe for 16 cycles, so all 25
// tile slots per 4 rows were paid (3.07 ms for 15M rows); an indirect-branch dispatch was slower still (5.0 ms).
#include "common.cuh"

namespace dlsa {

namespace {

constexpr int kCatMaxC = 8;      // categorical columns
constexpr int kCatMaxNum = 7;    // numeric columns (z = [1 | numerics] fills one 8-wide tile)
constexpr int kCatMaxWarps = 12;
constexpr int kCatStages = 2;
constexpr int kCatMaxRowsPerCta = 1 << 20;
constexpr int kCatMaxTiles = 6;     // 8-level DMMA tiles for the narrow categoricals
constexpr int kCatMaxLevels = 248;  // one byte per level index, 0xFF = no level
constexpr int kTabPitch = 9;        // doubles per level: w*z (8) | r
constexpr double kWScale = 17592186044416.0;  // 2^44
// per-warp, per-stage shared memory: 32 rows of numerics (7 doubles), codes (8 ints), labels
constexpr int kStageZ = 32 * kCatMaxNum * 8;
constexpr int kStageC = 32 * kCatMaxC * 4;
constexpr int kStageY = 32 * 8;
constexpr int kStageBytes = kStageZ + kStageC + kStageY;  // 3072
constexpr int kWzPitch = 10;                               // doubles per row: w | w*z_1..7 | r | 0
constexpr int kToPitch = 12;                                // ints per row: offsets of categoricals 0..7 | 4 x sink (idle lanes)
constexpr int kRowScratch = 32 * kWzPitch * 8 + 32 * kToPitch * 4;  // (w z', r) per row | table byte offset per (row, categorical)

struct CatArgs {
    const double* num;
    int64_t ldn;
    int n_num;
    const int* codes;
    int n_cat;
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
    double* part_d;                      // [K*S, stride_d]: Dz 64 | gz 8 | ll 1 | pad to 80 | table [Lpad][9]
    unsigned long long* part_i;          // [K*S, stride_i]: pairs [npairs]
    int64_t stride_d, stride_i;
    int nslots;                          // scatter slots in use (<= 96)
    unsigned char slot[96];              // slot -> categorical * 16 + column (0..7 = w z', 8 = r)
    int ntiles;                          // DMMA level tiles in use
    int tile_cat[kCatMaxTiles];          // categorical column of the tile
    int tile_lev0[kCatMaxTiles];         // first level (level-space index) of the tile
    int tile_nlev[kCatMaxTiles];         // levels of the tile that exist (1..8)
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

// ROUNDS = ceil(9 * n_cat / 32): scatter slots (categorical c, column j in 0..8) are dealt to the 32 lanes round by round.
// NC = categorical columns compiled into the row stage (4 or 8).  NTILE = compiled DMMA level tiles (0, 2, 4, 6).
template <int ROUNDS, int NC, int NTILE>
__global__ void __launch_bounds__(kCatMaxWarps * 32, 1) irls_cat_kernel(const CatArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int k = blockIdx.x / a.S;
    const int s = blockIdx.x - k * a.S;
    if (a.active != nullptr && a.active[k] == 0) return;
    const int nwarps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_num = a.n_num, n_cat = a.n_cat;
    const int tab_doubles = a.Lpad * kTabPitch;          // what the CTA reduction reads; the sink row follows it
    const int tab_alloc = ((a.Lpad + 1) * kTabPitch + 1) & ~1;   // even: every warp's block stays 16-byte aligned

    // shared: per warp [ring | row scratch | table + sink row] ; then beta (8 + Lpad)
    const size_t per_warp = (size_t)kCatStages * kStageBytes + kRowScratch + (size_t)tab_alloc * 8;
    unsigned char* mine = smem_raw + (size_t)warp * per_warp;
    unsigned char* ring = mine;
    double* wz = reinterpret_cast<double*>(mine + kCatStages * kStageBytes);                              // [32][10]
    int* toff = reinterpret_cast<int*>(mine + kCatStages * kStageBytes + 32 * kWzPitch * 8);              // [32][12]
    double* tab = reinterpret_cast<double*>(mine + kCatStages * kStageBytes + kRowScratch);
    double* bz = reinterpret_cast<double*>(smem_raw + (size_t)nwarps * per_warp);
    double* bl = bz + 8;

    // coefficients in internal order: [intercept | numerics (7) | levels]; zero this warp's table
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
        for (int i = lane; i < tab_alloc; i += 32) tab[i] = 0.0;
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

    unsigned long long* pairs = a.part_i + (size_t)blockIdx.x * a.stride_i;

    auto issue = [&](int64_t gi) {  // cp.async the gi-th group of this warp into stage gi % kCatStages
        if (gi < my_groups) {
            const int64_t R0 = r0 + (warp + gi * nwarps) * 32;
            const int nv = (int)((r1 - R0 < 32) ? (r1 - R0) : 32);
            unsigned char* st = ring + (size_t)(gi % kCatStages) * kStageBytes;
            const uint32_t zs = smem_u32(st), cs = zs + kStageZ, ys = cs + kStageC;
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

    // per-lane scatter slots: round q handles slot 32 q + lane = (categorical c, column j); j = 8 is the g column.  Idle
    // lanes take "categorical 8", whose offset is always the sink row, and column 9 of the (w z', r) row, which is 0.0.
    const uint32_t wz_u32 = smem_u32(wz), toff_u32 = smem_u32(toff), tab_u32 = smem_u32(tab);
    const int sink_off = a.Lpad * kTabPitch * 8;
    uint32_t slot_to[ROUNDS];   // shared address of toff[0][c]
    uint32_t slot_wz[ROUNDS];   // shared address of wz[0][j]
    uint32_t slot_tb[ROUNDS];   // shared address of tab[0][j]
#pragma unroll
    for (int q = 0; q < ROUNDS; ++q) {
        const int sidx = 32 * q + lane;
        int c = 8, j = 9;       // idle
        if (sidx < a.nslots) { c = a.slot[sidx] >> 4; j = a.slot[sidx] & 15; }
        slot_to[q] = toff_u32 + 4 * c;
        slot_wz[q] = wz_u32 + 8 * j;
        slot_tb[q] = tab_u32 + 8 * (j == 9 ? 0 : j);
    }
    // DMMA level tiles: lane (m = lane >> 2, k = lane & 3) holds one-hot[level lev0 + m][row 4 ks + k] of the A operand
    int tile_cmp[NTILE > 0 ? NTILE : 1];
    uint32_t tile_to[NTILE > 0 ? NTILE : 1];
    double td0[NTILE > 0 ? NTILE : 1], td1[NTILE > 0 ? NTILE : 1];
#pragma unroll
    for (int t = 0; t < NTILE; ++t) {
        const bool on = t < a.ntiles;
        tile_cmp[t] = on ? (a.tile_lev0[t] + (lane >> 2)) * (kTabPitch * 8) : -1;   // offsets are >= 0: -1 never matches
        tile_to[t] = toff_u32 + 4 * (on ? a.tile_cat[t] : 8) + 2 * (lane & 3) * (kToPitch * 4);
        td0[t] = 0.0;
        td1[t] = 0.0;
    }

    double dz0 = 0.0, dz1 = 0.0, ll = 0.0;
    double gz[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) gz[j] = 0.0;

    for (int64_t gi = 0; gi < kCatStages - 1; ++gi) issue(gi);

    const int jcol = lane >> 2;  // the z column this lane holds in the MMA fragments
    for (int64_t gi = 0; gi < my_groups; ++gi) {
        issue(gi + kCatStages - 1);
        cp_async_wait<kCatStages - 1>();
        __syncwarp();
        const int64_t R0 = r0 + (warp + gi * nwarps) * 32;
        const int nv = (int)((r1 - R0 < 32) ? (r1 - R0) : 32);
        unsigned char* st = ring + (size_t)(gi % kCatStages) * kStageBytes;
        double* zs = reinterpret_cast<double*>(st);
        const int* cs = reinterpret_cast<const int*>(st + kStageZ);
        const double* ys = reinterpret_cast<const double*>(st + kStageZ + kStageC);

        // ---- row stage: lane = row -----------------------------------------------------------------------
        const bool valid = lane < nv;
        double eta = bz[0];
        double z[kCatMaxNum];
#pragma unroll
        for (int j = 0; j < kCatMaxNum; ++j) {
            z[j] = 0.0;
            if (j < n_num) {
                double v = zs[lane * n_num + j];
                if (a.mean != nullptr) {  // dlsa/models.py:99-101, same expression as dlsa_onehot_expand_f64
                    v = (v - a.mean[j]) / a.sd[j];
                    zs[lane * n_num + j] = v;
                }
                z[j] = v;
                eta = fma(v, bz[1 + j], eta);
            }
        }
        int lev[NC];
        int to[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) to[c] = sink_off;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            lev[c] = -1;
            if (c < n_cat && valid) {
                const int code = cs[lane * n_cat + c];
                if (code >= 0) {
                    lev[c] = a.cat_lo[c] + code;
                    eta += bl[lev[c]];
                    to[c] = lev[c] * (kTabPitch * 8);
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
        gz[0] += r;
#pragma unroll
        for (int j = 0; j < kCatMaxNum; ++j) gz[1 + j] = fma(r, z[j], gz[1 + j]);
        {   // (w, w z_1..7, r, 0) and the table offsets of this row
            double2* wrow = reinterpret_cast<double2*>(wz + lane * kWzPitch);
            wrow[0] = make_double2(w, w * z[0]);
            wrow[1] = make_double2(w * z[1], w * z[2]);
            wrow[2] = make_double2(w * z[3], w * z[4]);
            wrow[3] = make_double2(w * z[5], w * z[6]);
            wrow[4] = make_double2(r, 0.0);
            int4* trow = reinterpret_cast<int4*>(toff + lane * kToPitch);
            trow[0] = make_int4(to[0], to[1], to[2], to[3]);
            trow[1] = make_int4(to[4], to[5], to[6], to[7]);
            trow[2] = make_int4(sink_off, sink_off, sink_off, sink_off);
        }
        {   // weighted pair histogram: every (c, d) block carries a sink row and a sink column (index 0) that take the rows
            // with a baseline level, so the cell index is two multiply-adds and the REDs are unconditional
            const long long wfix = __double2ll_rn(w * kWScale);
            int li[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) li[c] = lev[c] < 0 ? 0 : lev[c] - a.cat_lo[c < n_cat ? c : 0] + 1;
            switch (n_cat) {
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

        // ---- dense z block: 8 k-steps of 4 rows, one DMMA each: A = z, B = w z (already in shared memory); the narrow
        // categoricals' H[level, z] tiles reuse B with A = one-hot(level) ---------------------------------------------
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            // the 4 rows of a k-step are 2 apart: with the 10-double row pitch of (w z', r) a half-warp's 16 operands then
            // fall into 16 different bank pairs (consecutive rows collide: r02 ncu, 4 wavefronts per load instead of 2)
            const int src = 8 * (ks >> 1) + (ks & 1) + 2 * (lane & 3);
            double zz = 1.0;
            if (jcol > 0) zz = (jcol - 1 < n_num) ? zs[src * n_num + jcol - 1] : 0.0;
            const double bw = wz[src * kWzPitch + jcol];
            dmma884(dz0, dz1, zz, bw);
#pragma unroll
            for (int t = 0; t < NTILE; ++t) {
                int o;
                asm volatile("ld.shared.s32 %0, [%1];" : "=r"(o) : "r"(tile_to[t] + (uint32_t)(8 * (ks >> 1) + (ks & 1)) * (kToPitch * 4u)));
                dmma884(td0[t], td1[t], (o == tile_cmp[t]) ? 1.0 : 0.0, bw);
            }
        }

        // ---- level blocks: scatter-add into the warp's table, two rows per step ------------------------------------
        // Lane (c, j) always owns column j of categorical c's levels, so successive rows that hit the same level are
        // ordered by program order of ONE thread; inside a row pair the second row's sum is forwarded in registers.
        // All loads of a pair (both rounds: their slots are disjoint) are issued before its stores.
        uint32_t pa[ROUNDS], pb[ROUNDS];
        double va[ROUNDS], vb[ROUNDS];
        auto load_pair = [&](int rp) {   // table addresses and addends of rows 2 rp, 2 rp + 1 for this lane's slots
            const uint32_t ra_t = (uint32_t)(2 * rp) * (kToPitch * 4u), ra_w = (uint32_t)(2 * rp) * (kWzPitch * 8u);
#pragma unroll
            for (int q = 0; q < ROUNDS; ++q) {
                int oa, ob;
                asm volatile("ld.shared.s32 %0, [%1];" : "=r"(oa) : "r"(slot_to[q] + ra_t));
                asm volatile("ld.shared.s32 %0, [%1];" : "=r"(ob) : "r"(slot_to[q] + ra_t + kToPitch * 4u));
                asm volatile("ld.shared.f64 %0, [%1];" : "=d"(va[q]) : "r"(slot_wz[q] + ra_w));
                asm volatile("ld.shared.f64 %0, [%1];" : "=d"(vb[q]) : "r"(slot_wz[q] + ra_w + kWzPitch * 8u));
                pa[q] = slot_tb[q] + (uint32_t)oa;
                pb[q] = slot_tb[q] + (uint32_t)ob;
            }
        };
        load_pair(0);
#pragma unroll
        for (int rp = 0; rp < 16; ++rp) {
            uint32_t ca[ROUNDS], cb[ROUNDS];
            double ta[ROUNDS], tb[ROUNDS], xa[ROUNDS], xb[ROUNDS];
#pragma unroll
            for (int q = 0; q < ROUNDS; ++q) {
                ca[q] = pa[q]; cb[q] = pb[q]; xa[q] = va[q]; xb[q] = vb[q];
                asm volatile("ld.shared.f64 %0, [%1];" : "=d"(ta[q]) : "r"(ca[q]) : "memory");
                asm volatile("ld.shared.f64 %0, [%1];" : "=d"(tb[q]) : "r"(cb[q]) : "memory");
            }
            if (rp + 1 < 16) load_pair(rp + 1);   // the row scratch is not written inside this loop: safe to read ahead
#pragma unroll
            for (int q = 0; q < ROUNDS; ++q) {
                const double na = ta[q] + xa[q];
                const double nb = ((ca[q] == cb[q]) ? na : tb[q]) + xb[q];
                asm volatile("st.shared.f64 [%0], %1;" ::"r"(ca[q]), "d"(na) : "memory");
                asm volatile("st.shared.f64 [%0], %1;" ::"r"(cb[q]), "d"(nb) : "memory");
            }
        }
        __syncwarp();  // the stage and the row scratch are re-filled by the next iteration
    }
    cp_async_wait<0>();
    // the DMMA tiles' sums go where the scatter would have put them (those cells were never touched)
#pragma unroll
    for (int t = 0; t < NTILE; ++t) {
        if (t < a.ntiles && (lane >> 2) < a.tile_nlev[t]) {
            double* cell = tab + (a.tile_lev0[t] + (lane >> 2)) * kTabPitch + 2 * (lane & 3);
            cell[0] = td0[t];
            cell[1] = td1[t];
        }
    }

    // ---- CTA reduction in a fixed warp order, straight into the workspace partial ----------------------------------
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) gz[j] += shfl_xor_f64(gz[j], m);
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) ll += shfl_xor_f64(ll, m);
    __syncthreads();  // every warp is done with its ring: reuse stage 0 of each ring for the small per-warp results
    {
        double* sm = reinterpret_cast<double*>(ring);  // Dz 64 | gz 8 | ll
        const int di = lane >> 2, dj = 2 * (lane & 3);
        sm[di * 8 + dj] = dz0;
        sm[di * 8 + dj + 1] = dz1;
        if (lane < 8) {
            double v = gz[0];
#pragma unroll
            for (int j = 1; j < 8; ++j) v = (lane == j) ? gz[j] : v;
            sm[64 + lane] = v;
        }
        if (lane == 8) sm[72] = ll;
    }
    __syncthreads();
    double* out = a.part_d + (size_t)blockIdx.x * a.stride_d;
    for (int i = threadIdx.x; i < 80 + tab_doubles; i += blockDim.x) {
        double v = 0.0;
        if (i < 73) {
            for (int wv = 0; wv < nwarps; ++wv) v += reinterpret_cast<const double*>(smem_raw + (size_t)wv * per_warp)[i];
        } else if (i >= 80) {
            for (int wv = 0; wv < nwarps; ++wv)
                v += reinterpret_cast<const double*>(smem_raw + (size_t)wv * per_warp + kCatStages * kStageBytes +
                                                     kRowScratch)[i - 80];
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
    double* Hk = H + (size_t)k * p * p;
    for (int e = threadIdx.x; e < p * p; e += blockDim.x) {
        const int o1 = e / p, o2 = e - o1 * p;
        if (o2 > o1) continue;
        int u1 = internal(o1), u2 = internal(o2);
        if (u1 > u2) { const int t = u1; u1 = u2; u2 = t; }
        double v = 0.0;
        if (u2 < 8) {
            for (int sl = 0; sl < a.S; ++sl) v += pd[sl * a.stride_d + u1 * 8 + u2];
        } else if (u1 < 8) {
            for (int sl = 0; sl < a.S; ++sl) v += pd[sl * a.stride_d + 80 + (u2 - 8) * kTabPitch + u1];
        } else {
            const int l1 = u1 - 8, l2 = u2 - 8;
            if (l1 == l2) {
                for (int sl = 0; sl < a.S; ++sl) v += pd[sl * a.stride_d + 80 + l1 * kTabPitch];
            } else {
                const int c1 = cat_of(l1), c2 = cat_of(l2);
                if (c1 != c2) {
                    const int idx = a.pair_off[c1 * kCatMaxC + c2] + (l1 - a.cat_lo[c1] + 1) * a.pair_ld[c2] +
                                    (l2 - a.cat_lo[c2] + 1);
                    for (int sl = 0; sl < a.S; ++sl)  // per-slice conversion: a slice sum is < 2^63, the partition sum may not be
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

void cat_plan(int K, int n_cat, const int* n_levels, int64_t max_part_rows, CatPlan& pl) {
    pl.Ltot = 0;
    pl.npairs = 0;
    for (int c = 0; c < n_cat; ++c) {
        for (int d = 0; d < c; ++d) pl.npairs += (n_levels[d] + 1) * (n_levels[c] + 1);  // + sink row and column
        pl.Ltot += n_levels[c];
    }
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
    // as many warps as shared memory allows (each owns a ring, a row scratch and a level table)
    const size_t per_warp = (size_t)kCatStages * kStageBytes + kRowScratch +
                            (size_t)((((pl.Lpad + 1) * kTabPitch + 1) & ~1)) * 8;
    const size_t fixed = (size_t)(8 + pl.Lpad) * 8;
    const size_t budget = 227 * 1024;
    int warps = (int)((budget - fixed) / per_warp);
    if (warps > kCatMaxWarps) warps = kCatMaxWarps;
    pl.warps = warps;
    pl.smem = (size_t)warps * per_warp + fixed;
}

}  // namespace

size_t irls_cat_workspace_bytes(int K, int n_cat, const int* n_levels, int64_t max_part_rows) {
    CatPlan pl;
    cat_plan(K, n_cat, n_levels, max_part_rows, pl);
    return pl.bytes_d + pl.bytes_i + 256;
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
    cat_plan(K, n_cat, n_levels, max_part_rows, pl);
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
    const uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255);
    a.part_d = reinterpret_cast<double*>(base);
    a.part_i = reinterpret_cast<unsigned long long*>(base + pl.bytes_d);
    a.stride_d = pl.stride_d;
    a.stride_i = pl.stride_i;
    DLSA_CUDA(cudaMemsetAsync(a.part_i, 0, pl.bytes_i, st));
    // narrow categoricals (<= 16 levels, narrowest first, <= kCatMaxTiles tiles in total) go to DMMA tiles; the others —
    // and every g column — to the scatter slots
    static const int tile_budget = [] {
        const char* e = getenv("DLSA_CAT_DMMA_TILES");
        int v = e ? atoi(e) : kCatMaxTiles;
        return v < 0 ? 0 : (v > kCatMaxTiles ? kCatMaxTiles : v);
    }();
    bool on_dmma[kCatMaxC] = {};
    for (;;) {
        int best = -1;
        for (int c = 0; c < n_cat; ++c)
            if (!on_dmma[c] && n_levels[c] >= 1 && n_levels[c] <= 16 && (best < 0 || n_levels[c] < n_levels[best])) best = c;
        if (best < 0 || a.ntiles + (n_levels[best] + 7) / 8 > tile_budget) break;
        on_dmma[best] = true;
        for (int b = 0; b < n_levels[best]; b += 8) {
            a.tile_cat[a.ntiles] = best;
            a.tile_lev0[a.ntiles] = a.cat_lo[best] + b;
            a.tile_nlev[a.ntiles] = n_levels[best] - b < 8 ? n_levels[best] - b : 8;
            ++a.ntiles;
        }
    }
    // Slot layout: 64-bit shared accesses are served one HALF-warp at a time, so a scattered categorical's 2 + n_num
    // columns (one run of consecutive doubles in its level's row) are kept inside one half-warp and the single g
    // columns of the DMMA categoricals fill the gaps — two runs in one half-warp always overlap in some bank.
    {
        const int run = n_num + 2;                   // w | w z_1..n_num | r
        int fill[6] = {0, 0, 0, 0, 0, 0};            // lanes used per half-warp (3 rounds x 2)
        unsigned char lay[96];
        for (int i = 0; i < 96; ++i) lay[i] = 0x89;  // idle: categorical 8 (always the sink row), column 9 (always 0.0)
        auto place = [&](int c, int j0, int cnt, bool r_col) -> bool {
            for (int h = 0; h < 6; ++h) {
                if (fill[h] + cnt + (r_col ? 1 : 0) > 16) continue;
                for (int j = 0; j < cnt; ++j) lay[16 * h + fill[h]++] = (unsigned char)(c * 16 + j0 + j);
                if (r_col) lay[16 * h + fill[h]++] = (unsigned char)(c * 16 + 8);
                return true;
            }
            return false;
        };
        bool ok = true;
        for (int c = 0; c < n_cat && ok; ++c)
            if (!on_dmma[c]) ok = place(c, 0, run - 1, true);
        for (int c = 0; c < n_cat && ok; ++c)
            if (on_dmma[c]) ok = place(c, 8, 1, false);
        if (ok) {
            int last = 0;
            for (int h = 0; h < 6; ++h)
                if (fill[h] > 0) last = 16 * h + 16;
            a.nslots = last;
            for (int i = 0; i < last; ++i) a.slot[i] = lay[i];
        } else {  // more runs than half-warps (7+ scattered categoricals): dense packing, 9 n_cat <= 72 slots always fit
            a.nslots = 0;
            for (int c = 0; c < n_cat; ++c)
                if (!on_dmma[c])
                    for (int j = 0; j <= n_num; ++j) a.slot[a.nslots++] = (unsigned char)(c * 16 + j);
            for (int c = 0; c < n_cat; ++c) a.slot[a.nslots++] = (unsigned char)(c * 16 + 8);
        }
    }
    const int rounds = (a.nslots + 31) / 32;  // 1..3
    const int ntile = (a.ntiles + 1) & ~1;
    const bool c4 = n_cat <= 4;
    int rc = DLSA_E_UNSUPPORTED;
#define DLSA_CAT_CASE(R, T)                                                                     \
    if (rounds == R && ntile == T) rc = c4 ? cat_launch_t<R, 4, T>(a, pl, st) : cat_launch_t<R, 8, T>(a, pl, st);
    DLSA_CAT_CASE(1, 0) DLSA_CAT_CASE(1, 2) DLSA_CAT_CASE(1, 4) DLSA_CAT_CASE(1, 6)
    DLSA_CAT_CASE(2, 0) DLSA_CAT_CASE(2, 2) DLSA_CAT_CASE(2, 4) DLSA_CAT_CASE(2, 6)
    DLSA_CAT_CASE(3, 0) DLSA_CAT_CASE(3, 2) DLSA_CAT_CASE(3, 4) DLSA_CAT_CASE(3, 6)
#undef DLSA_CAT_CASE
    if (rc != 0) return rc;
    irls_cat_finalize_kernel<<<K, 256, 0, st>>>(a, g, H, loglik);
    count_launch();
    return check_launch("irls_cat_finalize_kernel");
}

}  // namespace dlsa
