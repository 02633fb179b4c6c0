// C ABI glue: error reporting, device check, INR entry points (dispatch on precision).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace lmr {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int num_sms() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached = n;
        else return 148;  // B200
    }
    return cached;
}

namespace simt {
int forward(const lmr_inr_desc*, const float*, const float*, const float*, const float*, int, const float*, int, const float*, const lmr_affine*, int,
            float*, void*, size_t, cudaStream_t);
int backward(const lmr_inr_desc*, const float*, const float*, const float*, const float*, int, const float*, int, const float*, const lmr_affine*, int,
             const float*, float*, float*, float*, float*, float*, void*, size_t, cudaStream_t);
size_t workspace_bytes(const lmr_inr_desc*, int, int, int, bool);
size_t params_count(const lmr_inr_desc*);
}  // namespace simt

namespace tc {
int forward(const lmr_inr_desc*, const float*, const float*, const float*, const float*, int, const float*, int, const float*, const lmr_affine*, int,
            float*, void*, size_t, int, cudaStream_t);
int backward(const lmr_inr_desc*, const float*, const float*, const float*, const float*, int, const float*, int, const float*, const lmr_affine*, int,
             const float*, float*, float*, float*, float*, float*, void*, size_t, int, cudaStream_t, const void*, size_t, const AdamArgs*);
size_t workspace_bytes(const lmr_inr_desc*, int, int, int, bool, bool keep_act = false);
}  // namespace tc

}  // namespace lmr

using namespace lmr;

extern "C" int lmr_abi_version(void) { return LMR_ABI_VERSION; }
extern "C" const char* lmr_last_error(void) { return g_err; }

extern "C" int lmr_check_device(void) {
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
        set_error("no CUDA device");
        return LMR_ERR_NO_DEVICE;
    }
    if (prop.major != 10) {
        set_error("device %s is sm_%d%d; laminr_b200 kernels are built for sm_100a only", prop.name, prop.major, prop.minor);
        return LMR_ERR_NO_DEVICE;
    }
    return 0;
}

extern "C" size_t lmr_inr_params_count(const lmr_inr_desc* desc) { return simt::params_count(desc); }

static size_t max_sz(size_t a, size_t b) { return a > b ? a : b; }

extern "C" size_t lmr_inr_forward_workspace_bytes(const lmr_inr_desc* desc, int N, int P, int D) {
    return max_sz(simt::workspace_bytes(desc, N, P, D, false), tc::workspace_bytes(desc, N, P, D, false));
}
extern "C" size_t lmr_inr_forward_workspace_bytes_keep(const lmr_inr_desc* desc, int N, int P, int D) {
    return max_sz(simt::workspace_bytes(desc, N, P, D, false), tc::workspace_bytes(desc, N, P, D, false, true));
}
extern "C" size_t lmr_inr_backward_workspace_bytes(const lmr_inr_desc* desc, int N, int P, int D) {
    return max_sz(simt::workspace_bytes(desc, N, P, D, true), tc::workspace_bytes(desc, N, P, D, true));
}

extern "C" int lmr_inr_forward(const lmr_inr_desc* desc, const float* params, const float* B, const float* auxB, const float* grid, int N,
                               const float* coords, int P, const float* coord_shift, const lmr_affine* affine, int D, float* img, void* ws,
                               size_t ws_bytes, int precision, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int mode = precision & ~(LMR_PREC_KEEP_ACTIVATIONS | LMR_PREC_REUSE_COORD_TABLE);  // the flags are passed on to the tensor-core path, ignored by the fp32 path
    if (mode == LMR_PREC_FP32) return simt::forward(desc, params, B, auxB, grid, N, coords, P, coord_shift, affine, D, img, ws, ws_bytes, st);
    if (mode == LMR_PREC_BF16X3_FAST || mode == LMR_PREC_BF16X3)
        return tc::forward(desc, params, B, auxB, grid, N, coords, P, coord_shift, affine, D, img, ws, ws_bytes, precision, st);
    set_error("inr_forward: unknown precision %d", precision);
    return LMR_ERR_INVALID;
}

extern "C" int lmr_inr_backward(const lmr_inr_desc* desc, const float* params, const float* B, const float* auxB, const float* grid, int N,
                                const float* coords, int P, const float* coord_shift, const lmr_affine* affine, int D, const float* g_img,
                                float* g_params, float* g_angles, float* g_scalings, float* g_shears, float* g_translation, void* ws,
                                size_t ws_bytes, int precision, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (precision == LMR_PREC_FP32)
        return simt::backward(desc, params, B, auxB, grid, N, coords, P, coord_shift, affine, D, g_img, g_params, g_angles, g_scalings, g_shears,
                              g_translation, ws, ws_bytes, st);
    if (precision == LMR_PREC_BF16X3_FAST || precision == LMR_PREC_BF16X3)
        return tc::backward(desc, params, B, auxB, grid, N, coords, P, coord_shift, affine, D, g_img, g_params, g_angles, g_scalings, g_shears,
                            g_translation, ws, ws_bytes, precision, st, nullptr, 0, nullptr);
    set_error("inr_backward: unknown precision %d", precision);
    return LMR_ERR_INVALID;
}

extern "C" int lmr_inr_backward_after_forward(const lmr_inr_desc* desc, const float* params, const float* B, const float* auxB, const float* grid,
                                              int N, const float* coords, int P, const float* coord_shift, const lmr_affine* affine, int D,
                                              const float* g_img, float* g_params, float* g_angles, float* g_scalings, float* g_shears,
                                              float* g_translation, void* ws, size_t ws_bytes, const void* forward_workspace,
                                              size_t forward_workspace_bytes, int precision, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int mode = precision & ~LMR_PREC_KEEP_ACTIVATIONS;
    if (mode == LMR_PREC_BF16X3_FAST || mode == LMR_PREC_BF16X3)
        return tc::backward(desc, params, B, auxB, grid, N, coords, P, coord_shift, affine, D, g_img, g_params, g_angles, g_scalings, g_shears,
                            g_translation, ws, ws_bytes, precision, st, forward_workspace, forward_workspace_bytes, nullptr);
    return lmr_inr_backward(desc, params, B, auxB, grid, N, coords, P, coord_shift, affine, D, g_img, g_params, g_angles, g_scalings, g_shears,
                            g_translation, ws, ws_bytes, mode, stream);  // the fp32 path recomputes its tables
}

extern "C" int lmr_inr_backward_adam(const lmr_inr_desc* desc, float* params, const float* B, const float* auxB, const float* grid, int N,
                                     const float* coords, int P, const float* coord_shift, const float* g_img, float* g_params, float* adam_m,
                                     float* adam_v, int* adam_step, float lr, float beta1, float beta2, float eps, void* ws, size_t ws_bytes,
                                     const void* forward_workspace, size_t forward_workspace_bytes, int precision, void* stream) {
    LMR_CHECK_ARG(params && g_params && adam_m && adam_v && adam_step, "inr_backward_adam: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int mode = precision & ~LMR_PREC_KEEP_ACTIVATIONS;
    if ((mode == LMR_PREC_BF16X3_FAST || mode == LMR_PREC_BF16X3) && forward_workspace != nullptr) {
        const AdamArgs ad{params, adam_m, adam_v, adam_step, lr, beta1, beta2, eps};
        return tc::backward(desc, params, B, auxB, grid, N, coords, P, coord_shift, nullptr, 1, g_img, g_params, nullptr, nullptr, nullptr, nullptr, ws,
                            ws_bytes, precision, st, forward_workspace, forward_workspace_bytes, &ad);
    }
    // fp32 path (or no forward workspace): the two calls it stands for
    if (int rc = lmr_inr_backward_after_forward(desc, params, B, auxB, grid, N, coords, P, coord_shift, nullptr, 1, g_img, g_params, nullptr, nullptr,
                                                nullptr, nullptr, ws, ws_bytes, forward_workspace, forward_workspace_bytes, precision, stream))
        return rc;
    const size_t n = lmr_inr_params_count(desc);
    return lmr_adam_step(params, g_params, adam_m, adam_v, (int)n, lr, beta1, beta2, eps, adam_step, stream);
}
