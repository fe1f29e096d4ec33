// Microbenchmark: how the B200 FP64 pipe behaves when DMMA.8x8x4 streams are mixed with DFMA / DMUL from the same or from
// other warps of a sub-partition — the question behind K1's 80 % pipe utilisation (profiles/r02_k1_small_experiments.md).
// One CTA per SM, W warps per sub-partition; every warp runs `iters` k-steps of 21 DMMAs (distinct accumulators) plus a
// variant-specific set of scalar FP64 instructions.  Output: cycles per k-step per sub-partition / (W * 21 * 16) = the
// pipe's utilisation by DMMA alone, and the implied cost per scalar instruction.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void smsp_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// MODE 0: DMMA only
//      1: + NF DFMAs (shared multiplier) grouped after the DMMAs
//      2: + NF DFMAs interleaved one per DMMA
//      3: as 1, and the W warps of a sub-partition meet at a named barrier before and after the DFMA group
//      4: + 7 DMULs, each feeding the DMMAs of "its tile row" and issued immediately before them (dependent)
//      5: + 7 DMULs issued two tile rows ahead of their first use
//      6: as 4 + NF grouped DFMAs (the K1 mix: 21 DMMA + 7 DMUL + NF DFMA)
//      7: DFMA only (NF per step), no DMMA
//      8: A operands loaded from shared memory every k-step (LDS.64 -> DMMA, no scalar FP64 at all)
//      9: as 8, and the A operand of each tile row's last DMMA is a per-lane select (2 x FSEL -> DMMA)
//     10: as 9 + NF grouped DFMAs (the pre-scaled K1 mix: LDS + 21 DMMA + 6 selects + NF DFMA)
template <int MODE, int NF, int W, int R>
__global__ void __launch_bounds__(128 * W, 1) k_mix(double* out, const double* in, int iters, int warps_per_smsp) {
    constexpr int NTL = R * (R + 1) / 2;   // DMMA tiles per k-step: R tile rows of 1..R tiles
    double C0[NTL], C1[NTL], f[8], av[8], e[NF > 0 ? NF : 1];
#pragma unroll
    for (int i = 0; i < NTL; i++) { C0[i] = 0; C1[i] = 0; }
#pragma unroll
    for (int i = 0; i < 8; i++) { f[i] = in[i] + threadIdx.x * 1e-12; av[i] = in[8 + i] + threadIdx.x * 2e-12; }
#pragma unroll
    for (int i = 0; i < NF; i++) e[i] = in[16 + i];
    const double rv = in[100] + threadIdx.x * 1e-12;
    double wv = in[101] + threadIdx.x * 1e-13;
    const int smsp = (threadIdx.x >> 5) & 3;
    __shared__ double sm[512 * 8 + 64];
    for (int i = threadIdx.x; i < 512 * 8 + 64; i += blockDim.x) sm[i] = in[i & 127];
    __syncthreads();
    const bool sel = (threadIdx.x & 28) == 0;
    for (int it = 0; it < iters; it++) {
        if (MODE >= 8) {
            const double* src = sm + (threadIdx.x & 511) + (it & 7) * 8;
#pragma unroll
            for (int i = 0; i < R; i++) { av[i] = src[i * 512]; }
            int t = 0;
#pragma unroll
            for (int i = 0; i < R; i++) {
#pragma unroll
                for (int j = 0; j < i; j++) { dmma884(C0[t], C1[t], av[i], av[j]); t++; }
                const double ad = (MODE >= 9 && sel) ? rv : av[i];
                dmma884(C0[t], C1[t], ad, av[i]);
                t++;
            }
            if (MODE == 10) {
#pragma unroll
                for (int i = 0; i < NF; i++) e[i] = fma(rv, av[i % R], e[i]);
            }
            continue;
        }
        if (MODE == 7) {
#pragma unroll
            for (int i = 0; i < NF; i++) e[i] = fma(rv, f[i & 7], e[i]);
            continue;
        }
        if (MODE == 5) { av[0] = wv * f[0]; av[1] = wv * f[1]; }
        int t = 0;
#pragma unroll
        for (int i = 0; i < R; i++) {          // tile rows 0..R-1: i+1 tiles each
            if (MODE == 4 || MODE == 6) av[i] = wv * f[i];
            if (MODE == 5 && i + 2 < R) av[i + 2] = wv * f[i + 2];
#pragma unroll
            for (int j = 0; j <= i; j++) {
                dmma884(C0[t], C1[t], av[i], f[j]);
                if (MODE == 2 && t < NF) e[t] = fma(rv, f[t & 7], e[t]);
                t++;
            }
        }
        if (MODE == 3) smsp_bar(1 + smsp, 32 * warps_per_smsp);
        if (MODE == 1 || MODE == 3 || MODE == 6) {
#pragma unroll
            for (int i = 0; i < NF; i++) e[i] = fma(rv, f[i & 7], e[i]);
        }
        if (MODE == 3) smsp_bar(1 + smsp, 32 * warps_per_smsp);
        wv += 1e-13;   // keep the DMULs loop-variant
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NTL; i++) s += C0[i] + C1[i];
#pragma unroll
    for (int i = 0; i < NF; i++) s += e[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + av[7];
}

template <typename F>
float timeit(F f, int reps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    int clk_khz; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    const double clk = clk_khz * 1e3;
    double *out, *in; CK(cudaMalloc(&out, sizeof(double) * sms * 1024)); CK(cudaMalloc(&in, sizeof(double) * 256));
    double hin[256]; for (int i = 0; i < 256; i++) hin[i] = 1.0 + 1e-9 * i;
    CK(cudaMemcpy(in, hin, sizeof(hin), cudaMemcpyHostToDevice));
    const int iters = 20000;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"unit\": \"SM clocks per k-step per sub-partition, all warps together\"", prop.name, sms);
#define RUN(MODE, NF, W, R, name) { \
        float ms = timeit([&] { k_mix<MODE, NF, W, R><<<sms, 128 * (W)>>>(out, in, iters, W); }, 3); \
        printf(", \"%s_r%d_w%d\": %.1f", name, R, W, ms * 1e-3 * clk / (double)iters); }
#define SET(W, R, NF) \
    RUN(0, 0, W, R, "dmma") RUN(1, NF, W, R, "dmma_dfma_grouped") RUN(2, NF, W, R, "dmma_dfma_interleaved") \
    RUN(3, NF, W, R, "dmma_dfma_grouped_phase_aligned") RUN(4, 0, W, R, "dmma_dmul_adjacent") RUN(5, 0, W, R, "dmma_dmul_ahead") \
    RUN(6, NF, W, R, "k1mix") RUN(7, NF, W, R, "dfma_only") RUN(8, 0, W, R, "lds_dmma") RUN(9, 0, W, R, "lds_fsel_dmma") \
    RUN(10, NF, W, R, "prescaled_k1mix")
    // 21 tiles + 12 DFMAs fit 255 registers (W = 1, 2); W = 4 (128 registers per thread) runs 15 tiles + 12 DFMAs,
    // and W = 2 repeats that shape for comparison.  DFMAs per k-step: 12; DMULs per k-step: R.
    SET(1, 6, 12) SET(2, 6, 12) SET(2, 5, 12) SET(4, 5, 12) SET(1, 5, 12)
    printf(", \"sm_clock_khz_attr\": %d}\n", clk_khz);
    return 0;
}
