// INR generator, tcgen05 tensor-core path (LMR_PREC_TF32 / LMR_PREC_BF16X3).  Placeholder until the kernels land.
#include "common.cuh"

namespace lmr {
namespace tc {

int forward(const lmr_inr_desc*, const float*, const float*, const float*, const float*, int, const float*, int, const float*, const lmr_affine*, int,
            float*, void*, size_t, int, cudaStream_t) {
    set_error("inr_forward: tensor-core precision modes are not built yet");
    return LMR_ERR_INVALID;
}
int backward(const lmr_inr_desc*, const float*, const float*, const float*, const float*, int, const float*, int, const float*, const lmr_affine*, int,
             const float*, float*, float*, float*, float*, float*, void*, size_t, int, cudaStream_t) {
    set_error("inr_backward: tensor-core precision modes are not built yet");
    return LMR_ERR_INVALID;
}
size_t workspace_bytes(const lmr_inr_desc*, int, int, int, bool) { return 0; }

}  // namespace tc
}  // namespace lmr
