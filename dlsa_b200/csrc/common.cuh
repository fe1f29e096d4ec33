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
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {  // non-blocking
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
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

// ---- branch-free FP64 link arithmetic ---------------------------------------------------------------------
// The per-row stage is software-pipelined between DMMAs of the same warp, which only works if it is straight-line
// code: CUDA's exp() and '/' carry slow-path branches/calls that split the basic block.  These versions are
// branch-free and accurate to ~1 ulp on the ranges used here.

// exp(x) for x in [-708, 709] (clamped: below -708, 1 + exp(x) == 1 anyway; 709 is the last finite exp), branch-free.
// exp_nonpos is the same routine under its historical name: the logistic link only calls it with x <= 0.
__device__ __forceinline__ double exp_clamped(double x);
__device__ __forceinline__ double exp_nonpos(double x) { return exp_clamped(x); }
__device__ __forceinline__ double exp_clamped(double x) {
    x = fmin(fmax(x, -708.0), 709.0);
    const double kShift = 6755399441055744.0;  // 1.5 * 2^52: adding it rounds to the nearest integer
    const double t = fma(x, 1.4426950408889634074, kShift);
    const int n = __double2loint(t);
    const double nf = t - kShift;
    double r = fma(nf, -6.93147180369123816490e-01, x);  // Cody-Waite: ln2_hi
    r = fma(nf, -1.90821492927058770002e-10, r);         //             ln2_lo
    double p = 1.6059043836821613e-10;                   // 1/13!   (|r| <= 0.3466: truncation < 1e-17)
    p = fma(p, r, 2.08767569878681e-09);                 // 1/12!
    p = fma(p, r, 2.505210838544172e-08);                // 1/11!
    p = fma(p, r, 2.755731922398589e-07);                // 1/10!
    p = fma(p, r, 2.7557319223985893e-06);               // 1/9!
    p = fma(p, r, 2.48015873015873e-05);                 // 1/8!
    p = fma(p, r, 1.984126984126984e-04);                // 1/7!
    p = fma(p, r, 1.388888888888889e-03);                // 1/6!
    p = fma(p, r, 8.333333333333333e-03);                // 1/5!
    p = fma(p, r, 4.1666666666666664e-02);               // 1/4!
    p = fma(p, r, 1.6666666666666666e-01);               // 1/3!
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return p * __hiloint2double((n + 1023) << 20, 0);    // * 2^n, n in [-1022, 1023]
}

// num / den for a positive normal den of moderate size: MUFU.RCP64H seed, two Newton steps, one residual
// correction (correctly rounded or 1 ulp off).
__device__ __forceinline__ double div_pos(double num, double den) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
    double e = fma(-den, r, 1.0);
    r = fma(r, e, r);
    e = fma(-den, r, 1.0);
    r = fma(r, e, r);
    const double qv = num * r;
    const double rem = fma(-den, qv, num);
    return fma(rem, r, qv);
}

// prob = 1/(1+exp(-eta)) for eta >= 0, exp(eta)/(1+exp(eta)) otherwise — the overflow-free form of
// sklearn's predict_proba / R wlse_func.R:52; w = prob*(1-prob) is formed by the caller (dlsa/models.py:130).
__device__ __forceinline__ double sigmoid_f64(double eta) {
    const double e = exp_nonpos(-fabs(eta));
    const double num = (eta >= 0.0) ? 1.0 : e;
    return div_pos(num, 1.0 + e);
}

// log(1 + e) for e in [0, 1], branch-free: u = 1 + e is halved when it exceeds sqrt(2), then
// log(u) = 2 atanh(s), s = (u-1)/(u+1), |s| <= 0.1716 (odd series to s^23: truncation < 2e-19), plus the
// first-order correction (e - (u-1))/u for the rounding of 1 + e (what makes log1p accurate for tiny e).
__device__ __forceinline__ double log1p_01(double e) {
    const double u = 1.0 + e;
    const double c = e - (u - 1.0);             // rounding error of u (exact)
    const bool big = u > 1.4142135623730951;
    const double v = big ? 0.5 * u : u;         // [0.7071, 1.4142]
    // v - 1 without cancellation error: for !big it is e - c; for big it is 0.5*u - 1 (exact: u in (1.41, 2])
    const double vm1 = big ? fma(0.5, u, -1.0) : (u - 1.0);
    const double s = div_pos(vm1, v + 1.0);
    const double z = s * s;
    double p = 1.0 / 23.0;
    p = fma(p, z, 1.0 / 21.0);
    p = fma(p, z, 1.0 / 19.0);
    p = fma(p, z, 1.0 / 17.0);
    p = fma(p, z, 1.0 / 15.0);
    p = fma(p, z, 1.0 / 13.0);
    p = fma(p, z, 1.0 / 11.0);
    p = fma(p, z, 1.0 / 9.0);
    p = fma(p, z, 1.0 / 7.0);
    p = fma(p, z, 1.0 / 5.0);
    p = fma(p, z, 1.0 / 3.0);
    p = p * z;                                  // atanh(s)/s - 1
    double r = fma(2.0 * s, p, 2.0 * s);        // log(v)
    r += big ? 6.93147180559945286227e-01 : 0.0;
    double ru;                                  // d log(u)/du * c = c / u: c <= 2^-53, a 20-bit reciprocal is plenty
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(ru) : "d"(u));
    return fma(c, ru, r);
}
__device__ __forceinline__ double log1pexp_f64(double eta) {
    return fmax(eta, 0.0) + log1p_01(exp_nonpos(-fabs(eta)));
}

}  // namespace dlsa
