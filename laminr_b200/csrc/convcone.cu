// Stacked-conv core + Gaussian readout response of ONE selected neuron per image, forward and input gradient in one launch
// (neuralpredictors Stacked2dCore.forward conv2d.py:246-253, FullGaussian2d.forward gaussian.py:532-581 with the eval
// grid of sample_grid :396-429, FiringRateEncoder.forward firing_rate.py:36-52, SingleNeuronModel laminr/utils/general.py:5-12).
//
// The reference evaluates every layer on the full frame for all neurons and then selects one output.  The selected neuron reads the last
// feature map at 4 bilinear taps only, so its response (and the gradient that flows back to the image) depends on a receptive-field CONE:
// a 2x2 window of the last layer, (2+k-1)^2 of the layer below, ... ((2 + sum(k_l - 1))^2 input pixels; 22x22 for kernels 9/7/7).
// One CTA evaluates the cone of one (image, neuron) pair entirely in shared memory: same arithmetic as the full-frame network on those
// positions (zero padding honoured by forcing every position outside the frame to 0), ~1000x fewer FLOPs at 200x200.
// fp32 CUDA-core FMAs (the cone is ~10 MFMA per image: far below the INR generator's cost, and exact fp32 keeps parity tight).
//
// Convolutions run as register-blocked items (CB output channels x XB adjacent pixels per thread, sliding input window in registers,
// weights as warp-uniform 128-bit loads that stay in L1), split over input-channel ranges when a layer has few pixels; the partial sums are
// combined in a fixed order (deterministic).  The data-gradient convolutions reuse the same routine on flipped, transposed weights with a
// virtual zero border.
#include "common.cuh"

namespace lmr {
namespace cone {

constexpr int NT = 512;
constexpr int MAXL = LMR_CONV_MAX_LAYERS;

__host__ __device__ inline int pick_cb(int co) { return (co % 4 == 0) ? 4 : (co % 2 == 0) ? 2 : 1; }
__host__ __device__ inline int ksplit_for(int So, int Co, int Cin) {
    const int cb = pick_cb(Co), xb = 16 / cb;
    const int items = So * ((So + xb - 1) / xb) * (Co / cb);
    int ks = NT / items;
    if (ks < 1) ks = 1;
    if (ks > Cin) ks = Cin;
    return ks;
}

struct Plan {
    int S[MAXL];  // output window edge of layer l
    int Sin;      // input window edge
    int off_in, off_a[MAXL], off_g[MAXL], off_gin, off_scratch, total;  // float offsets into dynamic shared memory
};

__host__ __device__ inline void make_plan(const lmr_convcore_desc& d, bool with_grad, Plan& p) {
    const int L = d.n_layers;
    p.S[L - 1] = 2;
    for (int l = L - 2; l >= 0; --l) p.S[l] = p.S[l + 1] + d.kernel[l + 1] - 1;
    p.Sin = p.S[0] + d.kernel[0] - 1;
    int off = 0;
    auto take = [&](int n) { int o = off; off += (n + 3) & ~3; return o; };
    p.off_in = take(d.c_in * p.Sin * p.Sin);
    for (int l = 0; l < L; ++l) p.off_a[l] = take(d.hidden * p.S[l] * p.S[l]);
    int scratch = 0;
    for (int l = 0; l < L; ++l) {  // forward convolutions
        const int cin = l == 0 ? d.c_in : d.hidden;
        const int n = ksplit_for(p.S[l], d.hidden, cin) * d.hidden * p.S[l] * p.S[l];
        scratch = n > scratch ? n : scratch;
    }
    if (with_grad) {
        for (int l = 0; l < L; ++l) p.off_g[l] = take(d.hidden * p.S[l] * p.S[l]);
        p.off_gin = take(d.c_in * p.Sin * p.Sin);
        for (int l = 0; l < L; ++l) {  // data-gradient convolutions: out = input window of layer l
            const int cout = l == 0 ? d.c_in : d.hidden, So = l == 0 ? p.Sin : p.S[l - 1];
            const int n = ksplit_for(So, cout, d.hidden) * cout * So * So;
            scratch = n > scratch ? n : scratch;
        }
    } else {
        for (int l = 0; l < L; ++l) p.off_g[l] = 0;
        p.off_gin = 0;
    }
    p.off_scratch = take(scratch);
    p.total = off;
}

struct Args {
    lmr_convcore_desc d;
    lmr_convcore_weights w;
    const float* img;
    long long group_stride;
    int NB, G;
    const int* neuron_of_group;
    const float* g_act;  // nullable
    float* act;
    float* g_img;  // nullable
    int accumulate;
};

// Partial sums of a stride-1 correlation: out[co, y, x] = sum_{ci in split, ky, kx} in[ci, y + ky - vp, x + kx - vp] * Wt[ci, ky, kx, co], input taken
// as 0 outside [0,Sin)^2 (vp = 0: "valid" forward convolution of the window; vp = k-1: "full" data-gradient convolution).
// scratch[ks][co][y*So + x] receives the partial sum of input-channel range ks.
template <int CB>
__device__ __forceinline__ void conv_items(const float* __restrict__ in, int Cin, int Sin, int vp, int So, int Co, int k,
                                           const float* __restrict__ Wt, float* __restrict__ scratch, int ksplit) {
    constexpr int XB = 16 / CB;
    const int nseg = (So + XB - 1) / XB, nsp = So * nseg, ncb = Co / CB;
    const int nitems = nsp * ncb * ksplit;
    for (int item = threadIdx.x; item < nitems; item += NT) {
        const int sp = item % nsp, t = item / nsp, cb = t % ncb, ks = t / ncb;
        const int y = sp / nseg, x0 = (sp - y * nseg) * XB;
        const int c_lo = ks * Cin / ksplit, c_hi = (ks + 1) * Cin / ksplit;
        float acc[CB][XB];
#pragma unroll
        for (int c = 0; c < CB; ++c)
#pragma unroll
            for (int j = 0; j < XB; ++j) acc[c][j] = 0.f;
        const int kx_lo = max(0, vp - x0 - (XB - 1)), kx_hi = min(k, vp - x0 + Sin);
        const int ky_lo = max(0, vp - y), ky_hi = min(k, vp - y + Sin);
        for (int ci = c_lo; ci < c_hi; ++ci) {
            const float* plane = in + ci * Sin * Sin;
            for (int ky = ky_lo; ky < ky_hi; ++ky) {
                const float* row = plane + (y + ky - vp) * Sin;
                const float* wrow = Wt + ((size_t)(ci * k + ky) * k) * Co + cb * CB;
                float r[XB];
                const int base = x0 + kx_lo - vp;
#pragma unroll
                for (int j = 0; j < XB - 1; ++j) {
                    const int ix = base + j;
                    r[j + 1] = (ix >= 0 && ix < Sin) ? row[ix] : 0.f;
                }
#pragma unroll 4
                for (int kx = kx_lo; kx < kx_hi; ++kx) {
#pragma unroll
                    for (int j = 0; j < XB - 1; ++j) r[j] = r[j + 1];
                    const int ix = x0 + kx - vp + XB - 1;
                    r[XB - 1] = (ix >= 0 && ix < Sin) ? row[ix] : 0.f;
                    float wv[CB];
                    if (CB == 4) {
                        const float4 q = __ldg(reinterpret_cast<const float4*>(wrow + kx * Co));
                        wv[0] = q.x, wv[1 % CB] = q.y, wv[2 % CB] = q.z, wv[3 % CB] = q.w;
                    } else if (CB == 2) {
                        const float2 q = __ldg(reinterpret_cast<const float2*>(wrow + kx * Co));
                        wv[0] = q.x, wv[1 % CB] = q.y;
                    } else {
                        wv[0] = __ldg(wrow + kx * Co);
                    }
#pragma unroll
                    for (int c = 0; c < CB; ++c)
#pragma unroll
                        for (int j = 0; j < XB; ++j) acc[c][j] = fmaf(wv[c], r[j], acc[c][j]);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < CB; ++c) {
            float* o = scratch + ((size_t)(ks * Co + cb * CB + c) * So + y) * So + x0;
#pragma unroll
            for (int j = 0; j < XB; ++j)
                if (x0 + j < So) o[j] = acc[c][j];
        }
    }
}

__device__ __forceinline__ void conv_dispatch(const float* in, int Cin, int Sin, int vp, int So, int Co, int k, const float* Wt, float* scratch, int ksplit) {
    const int cb = pick_cb(Co);
    if (cb == 4) conv_items<4>(in, Cin, Sin, vp, So, Co, k, Wt, scratch, ksplit);
    else if (cb == 2) conv_items<2>(in, Cin, Sin, vp, So, Co, k, Wt, scratch, ksplit);
    else conv_items<1>(in, Cin, Sin, vp, So, Co, k, Wt, scratch, ksplit);
}

__device__ __forceinline__ float elu_shifted(float v, float xs, float ys) {
    const float t = v - xs;
    return (t > 0.f ? t : expm1f(t)) + ys;
}
// derivative of elu(v - xs) + ys recovered from the stored output
__device__ __forceinline__ float elu_grad_from_out(float out, float ys) {
    const float t = out - ys;
    return t > 0.f ? 1.f : t + 1.f;
}

__global__ void __launch_bounds__(NT, 1) convcone_kernel(const Args a) {
    extern __shared__ __align__(16) float sm[];
    __shared__ float s_tap[4];
    __shared__ float s_red[40];
    __shared__ float s_gy;
    const lmr_convcore_desc& d = a.d;
    const int L = d.n_layers, H = d.height, W = d.width, P = H * W, Hc = d.hidden, Cin = d.c_in;
    const bool with_grad = a.g_img != nullptr;
    Plan pl;
    make_plan(d, with_grad, pl);
    const int g = blockIdx.x / a.NB, b = blockIdx.x - g * a.NB;
    const int neuron = a.neuron_of_group[g];
    const size_t img_off = (size_t)g * (size_t)a.group_stride + (size_t)b * Cin * P;
    const float* img = a.img + img_off;
    const int tid = threadIdx.x;

    // ---- readout position -> 2x2 tap window of the last layer (torch grid_sample, bilinear, zero padding)
    float mx = fminf(fmaxf(a.w.mu[2 * neuron], -1.f), 1.f), my = fminf(fmaxf(a.w.mu[2 * neuron + 1], -1.f), 1.f);
    float fx, fy;
    if (d.align_corners) {
        fx = ((mx + 1.f) / 2.f) * (float)(W - 1);
        fy = ((my + 1.f) / 2.f) * (float)(H - 1);
    } else {
        fx = ((mx + 1.f) * (float)W - 1.f) / 2.f;
        fy = ((my + 1.f) * (float)H - 1.f) / 2.f;
    }
    const float flx = floorf(fx), fly = floorf(fy);
    const int X0 = (int)flx, Y0 = (int)fly;
    const float wx1 = fx - flx, wy1 = fy - fly, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
    if (tid == 0) s_tap[0] = wy0 * wx0, s_tap[1] = wy0 * wx1, s_tap[2] = wy1 * wx0, s_tap[3] = wy1 * wx1;

    // window origins (frame coordinates) of every layer's output and of the input
    int oy[MAXL], ox[MAXL];
    oy[L - 1] = Y0, ox[L - 1] = X0;
    for (int l = L - 2; l >= 0; --l) oy[l] = oy[l + 1] - d.kernel[l + 1] / 2, ox[l] = ox[l + 1] - d.kernel[l + 1] / 2;
    const int iy0 = oy[0] - d.kernel[0] / 2, ix0 = ox[0] - d.kernel[0] / 2;
    const int Sin = pl.Sin;

    // ---- input window
    float* win = sm + pl.off_in;
    for (int i = tid; i < Cin * Sin * Sin; i += NT) {
        const int c = i / (Sin * Sin), r = i - c * Sin * Sin, y = r / Sin, x = r - y * Sin;
        const int gy = iy0 + y, gx = ix0 + x;
        win[i] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? __ldg(img + (size_t)c * P + gy * W + gx) : 0.f;
    }
    __syncthreads();

    // ---- forward through the layers
    float* scratch = sm + pl.off_scratch;
    for (int l = 0; l < L; ++l) {
        const int cin = l == 0 ? Cin : Hc, So = pl.S[l], Si = l == 0 ? Sin : pl.S[l - 1], k = d.kernel[l];
        const float* in = l == 0 ? win : sm + pl.off_a[l - 1];
        const int ks = ksplit_for(So, Hc, cin);
        conv_dispatch(in, cin, Si, 0, So, Hc, k, a.w.w_fwd[l], scratch, ks);
        __syncthreads();
        float* out = sm + pl.off_a[l];
        const int n = Hc * So * So;
        for (int i = tid; i < n; i += NT) {
            const int c = i / (So * So), r = i - c * So * So, y = r / So, x = r - y * So;
            float v = 0.f;
            for (int s = 0; s < ks; ++s) v += scratch[s * n + i];
            v = fmaf(v, __ldg(a.w.scale[l] + c), __ldg(a.w.shift[l] + c));
            if (d.nonlin[l]) v = elu_shifted(v, d.elu_xshift, d.elu_yshift);
            const int gy = oy[l] + y, gx = ox[l] + x;
            out[i] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? v : 0.f;  // outside the frame = the next layer's zero padding
        }
        __syncthreads();
    }

    // ---- readout + encoder nonlinearity
    const float* top = sm + pl.off_a[L - 1];  // [Hc][2][2]
    float part = 0.f;
    for (int i = tid; i < Hc * 4; i += NT) part += top[i] * s_tap[i & 3] * __ldg(a.w.features + (size_t)neuron * Hc + (i >> 2));
    const float ysum = block_sum(part, s_red) + (a.w.bias ? __ldg(a.w.bias + neuron) : 0.f);
    const float pre = ysum + d.elu_offset;
    if (tid == 0) a.act[blockIdx.x] = (pre > 0.f ? pre : expm1f(pre)) + 1.f;
    if (!with_grad) return;

    // ---- backward: d(g_act * act) / d image
    const float gy_ = (a.g_act ? a.g_act[blockIdx.x] : 1.f) * (pre > 0.f ? 1.f : expf(pre));
    {
        float* gl = sm + pl.off_g[L - 1];
        for (int i = tid; i < Hc * 4; i += NT) {
            const int c = i >> 2, tap = i & 3;
            const int gyy = Y0 + (tap >> 1), gxx = X0 + (tap & 1);
            float v = gy_ * __ldg(a.w.features + (size_t)neuron * Hc + c) * s_tap[tap];
            v *= __ldg(a.w.scale[L - 1] + c) * (d.nonlin[L - 1] ? elu_grad_from_out(top[i], d.elu_yshift) : 1.f);
            gl[i] = (gyy >= 0 && gyy < H && gxx >= 0 && gxx < W) ? v : 0.f;
        }
    }
    __syncthreads();
    for (int l = L - 1; l >= 0; --l) {
        // gradient wrt the input window of layer l = full correlation of g[l] (pre-batch-norm gradient) with the flipped kernel
        const int cout = l == 0 ? Cin : Hc, So = l == 0 ? Sin : pl.S[l - 1], Si = pl.S[l], k = d.kernel[l];
        const int ks = ksplit_for(So, cout, Hc);
        conv_dispatch(sm + pl.off_g[l], Hc, Si, k - 1, So, cout, k, a.w.w_bwd[l], scratch, ks);
        __syncthreads();
        const int n = cout * So * So;
        if (l > 0) {
            float* gout = sm + pl.off_g[l - 1];
            const float* aout = sm + pl.off_a[l - 1];
            for (int i = tid; i < n; i += NT) {
                const int c = i / (So * So), r = i - c * So * So, y = r / So, x = r - y * So;
                float v = 0.f;
                for (int s = 0; s < ks; ++s) v += scratch[s * n + i];
                v *= __ldg(a.w.scale[l - 1] + c) * (d.nonlin[l - 1] ? elu_grad_from_out(aout[i], d.elu_yshift) : 1.f);
                const int gyy = oy[l - 1] + y, gxx = ox[l - 1] + x;
                gout[i] = (gyy >= 0 && gyy < H && gxx >= 0 && gxx < W) ? v : 0.f;
            }
        } else {
            float* gin = sm + pl.off_gin;
            for (int i = tid; i < n; i += NT) {
                float v = 0.f;
                for (int s = 0; s < ks; ++s) v += scratch[s * n + i];
                gin[i] = v;
            }
        }
        __syncthreads();
    }
    // ---- write the image gradient (this CTA owns the whole image: no atomics)
    float* gimg = a.g_img + img_off;
    const float* gin = sm + pl.off_gin;
    if (a.accumulate) {
        for (int i = tid; i < Cin * Sin * Sin; i += NT) {
            const int c = i / (Sin * Sin), r = i - c * Sin * Sin, y = r / Sin, x = r - y * Sin;
            const int gyy = iy0 + y, gxx = ix0 + x;
            if (gyy >= 0 && gyy < H && gxx >= 0 && gxx < W) gimg[(size_t)c * P + gyy * W + gxx] += gin[i];
        }
    } else {
        for (int i = tid; i < Cin * P; i += NT) {
            const int c = i / P, r = i - c * P, gyy = r / W, gxx = r - gyy * W;
            const int y = gyy - iy0, x = gxx - ix0;
            gimg[i] = (y >= 0 && y < Sin && x >= 0 && x < Sin) ? gin[(c * Sin + y) * Sin + x] : 0.f;
        }
    }
}

}  // namespace cone
}  // namespace lmr

using namespace lmr;

extern "C" int lmr_convcore_response(const lmr_convcore_desc* desc, const lmr_convcore_weights* weights, const float* img,
                                     long long img_group_stride, int NB, int G, const int* neuron_of_group, const float* g_act, float* act,
                                     float* g_img, int accumulate, void* stream) {
    LMR_CHECK_ARG(desc && weights && img && neuron_of_group && act, "convcore_response: null argument");
    const lmr_convcore_desc& d = *desc;
    LMR_CHECK_ARG(d.n_layers >= 1 && d.n_layers <= LMR_CONV_MAX_LAYERS, "convcore_response: n_layers %d not in 1..%d", d.n_layers, LMR_CONV_MAX_LAYERS);
    LMR_CHECK_ARG(d.c_in >= 1 && d.c_in <= 16 && d.hidden >= 1 && d.hidden <= 256, "convcore_response: unsupported channel counts (%d -> %d)", d.c_in, d.hidden);
    LMR_CHECK_ARG(d.height >= 1 && d.width >= 1 && NB >= 1 && G >= 1, "convcore_response: bad sizes");
    for (int l = 0; l < d.n_layers; ++l) {
        LMR_CHECK_ARG(d.kernel[l] >= 1 && (d.kernel[l] & 1) && d.kernel[l] <= 31, "convcore_response: kernel size %d of layer %d must be odd and <= 31", d.kernel[l], l);
        LMR_CHECK_ARG(weights->w_fwd[l] && weights->scale[l] && weights->shift[l] && (!g_img || weights->w_bwd[l]), "convcore_response: null weights of layer %d", l);
    }
    LMR_CHECK_ARG(weights->mu && weights->features, "convcore_response: null readout");
    LMR_CHECK_ARG(!g_img || img_group_stride != 0 || G == 1, "convcore_response: the gradient needs one image set per group (img_group_stride != 0)");
    cone::Plan pl;
    cone::make_plan(d, g_img != nullptr, pl);
    const size_t smem = (size_t)pl.total * sizeof(float);
    LMR_CHECK_ARG(smem <= 220 * 1024, "convcore_response: the receptive-field cone needs %zu bytes of shared memory (> 220 KB)", smem);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    LMR_CUDA_OK(cudaFuncSetAttribute(cone::convcone_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cone::Args a;
    a.d = d, a.w = *weights, a.img = img, a.group_stride = img_group_stride, a.NB = NB, a.G = G, a.neuron_of_group = neuron_of_group;
    a.g_act = g_act, a.act = act, a.g_img = g_img, a.accumulate = accumulate;
    cone::convcone_kernel<<<G * NB, cone::NT, smem, st>>>(a);
    LMR_LAUNCH_OK();
    return 0;
}
