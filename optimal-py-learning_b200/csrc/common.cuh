// Shared helpers for the opl_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "opl_b200.h"

namespace opl {

// thread-local error text returned by opl_last_error()
std::string &last_error();
int fail(int code, const char *fmt, ...);

#define OPL_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return ::opl::fail(OPL_E_CUDA, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, \
                               cudaGetErrorString(_e));                                        \
    } while (0)

#define OPL_LAUNCH_CHECK()                                                                     \
    do {                                                                                       \
        cudaError_t _e = cudaGetLastError();                                                   \
        if (_e != cudaSuccess)                                                                 \
            return ::opl::fail(OPL_E_CUDA, "kernel launch failed at %s:%d: %s", __FILE__,      \
                               __LINE__, cudaGetErrorString(_e));                              \
    } while (0)

constexpr double kBig = 1.7976931348623157e308;   // numpy.nan_to_num's finite stand-in for inf

int sm_count();   // multiprocessors of the current device (cached)
int current_device_slot();   // cudaGetDevice() clamped to [0, 64): index of per-device one-time setup flags

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------
// device-side reductions
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum; result valid in every thread.  `scratch` holds >= 33 doubles.
__device__ __forceinline__ double block_sum(double v, double *scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    if (warp == 0) {
        double t = lane < nwarps ? scratch[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}

// round-to-nearest FP32 -> TF32 (10-bit mantissa); the result is an FP32 bit pattern a tf32 tensor core reads exactly
__device__ __forceinline__ float round_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

// Softmax back-propagation for one output k: the reference contracts the upstream gradient g with the Jacobian
// J_kk = a_k (1 - a_k), J_jk = -a_j a_k (calculate.py:165-174, mlp.py:295):
//     delta_k = g_k a_k (1 - a_k) - a_k sum_{j != k} g_j a_j.
// With t = sum_j g_j a_j and p = g_k a_k this is a_k (g_k (1 - a_k) - (t - p)).  The algebraically equal short form
// a_k (g_k - t) must NOT be used: with a one-hot target and a confident prediction (a_k -> 1) g_k - t cancels and loses
// eps / (1 - a_k) of relative accuracy (measured 5.7e-7 on a saturated iris network), whereas 1 - a_k is exact and
// t - p is exactly zero for the target class of a one-hot row (every other g_j a_j is +-0).
// The product p and the difference t - p are rounded as SEPARATE operations on purpose: contracted into fma(-g, a, t) the
// difference would return the rounding residual of g a instead of 0 and bring exactly the lost accuracy back (measured).
__device__ __forceinline__ double softmax_delta(double a, double g, double t) {
    const double p = __dmul_rn(g, a);
    return a * (g * (1.0 - a) - __dsub_rn(t, p));
}
__device__ __forceinline__ float softmax_delta(float a, float g, float t) {
    const float p = __fmul_rn(g, a);
    return a * (g * (1.0f - a) - __fsub_rn(t, p));
}

// numpy.nan_to_num for float64: nan -> 0, +inf -> DBL_MAX, -inf -> -DBL_MAX
__device__ __forceinline__ double nan_to_num(double v) {
    if (isnan(v)) return 0.0;
    if (isinf(v)) return v > 0 ? kBig : -kBig;
    return v;
}

}  // namespace opl
