// Shared helpers for the libcnerf kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/cnerf_debug.h"

namespace cnerf {

// Per-thread error message (never throws across the C boundary).
int set_error(int code, const char* fmt, ...);

inline int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return CNERF_OK;
    return set_error(CNERF_ECUDA, "%s: %s", what, cudaGetErrorString(e));
}

#define CNERF_REQUIRE(cond, ...)                                   \
    do {                                                           \
        if (!(cond)) return ::cnerf::set_error(CNERF_EINVAL, __VA_ARGS__); \
    } while (0)

#define CNERF_LAUNCH_CHECK(name)                                   \
    do {                                                           \
        cudaError_t e__ = cudaGetLastError();                      \
        if (e__ != cudaSuccess) return ::cnerf::check_cuda(e__, name); \
    } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

constexpr int kNumSMs = 148;   // B200

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Small camera matrices travel by value in the kernel parameter space.
struct Mat3 { float m[9]; };
struct Mat34 { float m[12]; };

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// 128-bit streaming load through the read-only path.
__device__ __forceinline__ float4 ldg_f4(const float* p) {
    return __ldg(reinterpret_cast<const float4*>(p));
}

}  // namespace cnerf
