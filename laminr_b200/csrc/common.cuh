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

// state of a torch.optim.Adam step fused into another launch (lmr_inr_backward_adam): parameters, both moments, the int32[2] step word of lmr_adam_step
struct AdamArgs {
    float *p, *m, *v;
    int* step;
    float lr, b1, b2, eps;
};

// One element of torch.optim.Adam's single-tensor update (exp_avg.lerp_(grad, 1 - beta1); exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value =
// 1 - beta2); param.addcdiv_(exp_avg, sqrt(exp_avg_sq) / sqrt(bias_correction2) + eps, value = -lr / bias_correction1)).  Every rounding is spelled
// out with intrinsics: the expression is compiled into several kernels (lmr_adam_step, the fused learn tail), and left to the compiler each copy
// may contract its multiply-adds differently -- equal to the last bit only as long as the second moment is still zero.
__device__ __forceinline__ void adam_element(float g, float m0, float v0, float p0, float b1, float b2, float eps, float step_size, float bc2_sqrt,
                                             float* m_out, float* v_out, float* p_out) {
    const float mi = __fmaf_rn(__fsub_rn(g, m0), __fsub_rn(1.f, b1), m0);
    const float vi = __fmaf_rn(__fmul_rn(__fsub_rn(1.f, b2), g), g, __fmul_rn(v0, b2));
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vi), bc2_sqrt), eps);
    *m_out = mi, *v_out = vi;
    *p_out = __fmaf_rn(-step_size, __fdiv_rn(mi, denom), p0);
}
// bias corrections of step t (1-based): step_size = lr / (1 - beta1^t), sqrt(1 - beta2^t)
__device__ __forceinline__ void adam_step_consts(int t, float lr, float b1, float b2, float* step_size, float* bc2_sqrt) {
    const float bc1 = __fsub_rn(1.f, powf(b1, (float)t)), bc2 = __fsub_rn(1.f, powf(b2, (float)t));
    *step_size = __fdiv_rn(lr, bc1);
    *bc2_sqrt = __fsqrt_rn(bc2);
}

// Division of a small non-negative index by a run-time constant as one multiply-high (the flattened loops of the latency-bound kernels spent
// 20-40 % of their issued instructions in the ~20-instruction integer division sequence; ncu source page, round 2).  m = ceil(2^32 / n);
// floor(i m / 2^32) == i / n for every i with i * n < 2^32 (the indices here stay below 2^20, n below 2^11).
struct FastDiv {
    unsigned int n, m;
};
static inline FastDiv make_fastdiv(int n) {
    FastDiv d;
    d.n = (unsigned int)(n < 1 ? 1 : n);
    d.m = d.n == 1 ? 0u : (unsigned int)(((1ull << 32) + d.n - 1) / d.n);
    return d;
}
__device__ __forceinline__ int fd_div(int i, const FastDiv& d) { return d.n == 1 ? i : (int)__umulhi((unsigned int)i, d.m); }
// i -> (i / n, i % n)
__device__ __forceinline__ void fd_divmod(int i, const FastDiv& d, int* q, int* r) {
    *q = fd_div(i, d);
    *r = i - *q * (int)d.n;
}

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
