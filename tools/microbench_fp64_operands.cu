// Microbenchmark: what a DFMA / DMUL costs on the B200 FP64 pipe as a function of its operand pattern (register
// reuse) and of how it is interleaved with DMMA.8x8x4 in the same warp.  Motivation: profiles/r01_k1_ablation.md —
// inside K1 every edge-row / gradient DFMA costs ~1.7x its nominal 2 pipe cycles while DMULs cost nominal.
// Output: one JSON object on stdout; "cyc" figures are SM clocks per warp-instruction per sub-partition.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// MODE 0: c[i] = fma(a[i], b[i], c[i])   (three distinct register pairs per DFMA, no operand can be reused)
// MODE 1: c[i] = fma(a0,   b[i], c[i])   (one operand shared by consecutive DFMAs: reuse cache)
// MODE 2: c[i] = fma(a0,   b0,   c[i])   (two shared)
// MODE 3: c[i] = c[i] * a[i]             (DMUL, two distinct operands)
// MODE 4: c[i] = fma(a[i], b[i&3], c[i]) (groups of 4 share nothing consecutive; B repeats every 4)
template <int MODE, int N>
__global__ void k_dfma(double* out, const double* in, int iters) {
    double a[N], b[N], c[N];
#pragma unroll
    for (int i = 0; i < N; i++) { a[i] = in[i] + threadIdx.x * 1e-12; b[i] = in[N + i] + threadIdx.x * 2e-12; c[i] = in[2 * N + i] + threadIdx.x * 3e-12; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < N; i++) {
            if (MODE == 0) c[i] = fma(a[i], b[i], c[i]);
            if (MODE == 1) c[i] = fma(a[0], b[i], c[i]);
            if (MODE == 2) c[i] = fma(a[0], b[0], c[i]);
            if (MODE == 3) c[i] = c[i] * a[i];
            if (MODE == 4) c[i] = fma(a[i], b[i & 3], c[i]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < N; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// K1-like k-step in one warp: ND DMMAs (distinct accumulators, A from NA weighted fragments, B from fragments) and
// NF DFMAs arranged by MODE:
//   0: none (DMMA only)
//   1: DFMAs with distinct operands, issued after the DMMAs of the k-step (source order)
//   2: DFMAs sharing the multiplier (g[t] = fma(rv, f[t], g[t]) pattern), after the DMMAs
//   3: as 2 but interleaved one DFMA after each DMMA (source order)
//   4: as 1 but interleaved
template <int MODE, int ND, int NF>
__global__ void __launch_bounds__(256, 1) k_step(double* out, const double* in, int iters) {
    double C0[ND], C1[ND], f[8], av[8], e[NF > 0 ? NF : 1], x[NF > 0 ? NF : 1];
#pragma unroll
    for (int i = 0; i < ND; i++) { C0[i] = 0; C1[i] = 0; }
#pragma unroll
    for (int i = 0; i < 8; i++) { f[i] = in[i] + threadIdx.x * 1e-12; av[i] = in[8 + i] + threadIdx.x * 2e-12; }
#pragma unroll
    for (int i = 0; i < NF; i++) { e[i] = in[16 + i]; x[i] = in[48 + i] + threadIdx.x * 1e-12; }
    const double rv = in[100] + threadIdx.x * 1e-12;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ND; i++) {
            dmma884(C0[i], C1[i], av[i & 7], f[(i * 3) & 7]);
            if (MODE == 3 && i < NF) e[i] = fma(rv, f[i & 7], e[i]);
            if (MODE == 4 && i < NF) e[i] = fma(x[i], f[i & 7], e[i]);
        }
        if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < NF; i++) e[i] = fma(x[i], f[i & 7], e[i]);
        }
        if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < NF; i++) e[i] = fma(rv, f[i & 7], e[i]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ND; i++) s += C0[i] + C1[i];
#pragma unroll
    for (int i = 0; i < NF; i++) s += e[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
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
    printf("{\"gpu\": \"%s\", \"sms\": %d", prop.name, sms);
    // cycles per warp-level DFMA per sub-partition, at `wps` warps per SM (wps/4 per sub-partition)
#define RUN_DFMA(MODE, N, wps, name) { \
        const int threads = (wps) * 32; \
        float ms = timeit([&] { k_dfma<MODE, N><<<sms, threads>>>(out, in, iters); }, 3); \
        const double instr_per_smsp = (double)iters * (N) * ((wps) / 4.0); \
        printf(", \"%s_w%d_cyc\": %.3f", name, wps, ms * 1e-3 * clk / instr_per_smsp); }
    for (int wps : {8, 16}) {
        if (wps == 8) {
            RUN_DFMA(0, 16, 8, "dfma_distinct") RUN_DFMA(1, 16, 8, "dfma_share1") RUN_DFMA(2, 16, 8, "dfma_share2")
            RUN_DFMA(3, 16, 8, "dmul_distinct") RUN_DFMA(4, 16, 8, "dfma_bmod4")
        } else {
            RUN_DFMA(0, 16, 16, "dfma_distinct") RUN_DFMA(1, 16, 16, "dfma_share1") RUN_DFMA(2, 16, 16, "dfma_share2")
            RUN_DFMA(3, 16, 16, "dmul_distinct") RUN_DFMA(4, 16, 16, "dfma_bmod4")
        }
    }
    // K1-like k-step: cycles per k-step per sub-partition with 2 warps per sub-partition (256 threads, 1 CTA / SM)
#define RUN_STEP(MODE, ND, NF, name) { \
        float ms = timeit([&] { k_step<MODE, ND, NF><<<sms, 256>>>(out, in, iters); }, 3); \
        printf(", \"%s_cyc_per_step_per_warp\": %.2f", name, ms * 1e-3 * clk / ((double)iters * 2.0)); }
    RUN_STEP(0, 21, 0, "step_21dmma")
    RUN_STEP(1, 21, 21, "step_21dmma_21dfma_distinct_after")
    RUN_STEP(2, 21, 21, "step_21dmma_21dfma_shared_after")
    RUN_STEP(3, 21, 21, "step_21dmma_21dfma_shared_interleaved")
    RUN_STEP(4, 21, 21, "step_21dmma_21dfma_distinct_interleaved")
    RUN_STEP(0, 28, 0, "step_28dmma")
    printf(", \"sm_clock_khz_attr\": %d}\n", clk_khz);
    return 0;
}
