// Shared device/host helpers for libdlsa_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dlsa_b200.h"

namespace dlsa {

// ---- error plumbing (no C++ exceptions cross the ABI) --------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int check_launch(const char* what);  // cudaGetLastError -> return code (+ message)

#define DLSA_REQUIRE(cond, code, ...)      \
    do {                                   \
        if (!(cond)) {                     \
            ::dlsa::set_error(__VA_ARGS__); \
            return (code);                 \
        }                                  \
    } while (0)

#define DLSA_CUDA(call)                                                                   \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            ::dlsa::set_error("%s failed: %s", #call, cudaGetErrorString(e__));           \
            return (int)e__;                                                              \
        }                                                                                 \
    } while (0)

// ---- PTX wrappers ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// 8-byte cp.async (LDGSTS) for rows that are not 16-byte aligned
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

// FP64 tensor-core MMA: D(8x8) += A(8x4) * B(4x8).  SASS: DMMA.8x8x4
// lane l holds A[l/4][l%4], B[l%4][l/4], D[l/4][2*(l%4) + {0,1}]
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ double shfl_xor_f64(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

// logistic link pieces, written exactly as the reference evaluates them:
//   prob = 1/(1+exp(-eta)) (sklearn predict_proba / R wlse_func.R:52), w = prob*(1-prob) (dlsa/models.py:130)
__device__ __forceinline__ double sigmoid_f64(double eta) {
    if (eta >= 0) return 1.0 / (1.0 + exp(-eta));
    double e = exp(eta);
    return e / (1.0 + e);
}
__device__ __forceinline__ double log1pexp_f64(double eta) { return fmax(eta, 0.0) + log1p(exp(-fabs(eta))); }

}  // namespace dlsa
