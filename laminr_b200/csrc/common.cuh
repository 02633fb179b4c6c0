// Shared helpers for the laminr_b200 CUDA kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "laminr_b200.h"

namespace lmr {

void set_error(const char* fmt, ...);
int num_sms();

#define LMR_CHECK_ARG(cond, ...)           \
    do {                                   \
        if (!(cond)) {                     \
            lmr::set_error(__VA_ARGS__);   \
            return LMR_ERR_INVALID;        \
        }                                  \
    } while (0)

#define LMR_CUDA_OK(expr)                                                                          \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            lmr::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return LMR_ERR_CUDA;                                                                   \
        }                                                                                          \
    } while (0)

#define LMR_LAUNCH_OK() LMR_CUDA_OK(cudaGetLastError())

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum; every thread gets the result.  `scratch` needs >= 33 floats.  Safe to call back to back.
__device__ __forceinline__ float block_sum(float v, float* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    if (warp == 0) {
        float t = lane < nw ? scratch[lane] : 0.f;
        t = warp_sum(t);
        if (lane == 0) scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}

}  // namespace lmr
