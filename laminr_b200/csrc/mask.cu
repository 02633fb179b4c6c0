// RF-mask construction on the device, batched over images -- replaces `create_mask` (laminr/utils/mask.py:25-64), which the reference runs
// on the host (numpy / scipy.ndimage / skimage.morphology) once per neuron between the MEI ascent and learn / match (mei.py:56).  At population
// scale that host loop (~30 ms per neuron) dominates get_mei_dict once the ascent itself is batched on the GPU.
//
//   norm  = (mei - mean) / std                                  (population std)                         mask.py:50
//   thr   = |norm| > zscore_thresh                                                                       mask.py:51
//   close = binary_closing(thr, iterations)   = `iterations` dilations then erosions, 3x3 cross, outside = 0 (scipy.ndimage)   mask.py:52
//   obj   = largest 8-connected component (ties: the one met first in raster order, like argmax(bincount))  mask.py:53-55
//   hull  = convex_hull_image(obj): hull of the 4 edge mid-points of every pixel, tested at pixel centres (closed)  mask.py:56
//   mask  = gaussian_filter(hull, sigma) (separable, reflect, truncate 4, float64 accumulation, float32 between the axes)  mask.py:57
//   centroid = mean row / column index of the hull pixels + 0.5                                           mask.py:61
//
// Kernel 1 (one CTA per image, everything in shared memory): statistics, threshold, morphology, connected components by min-label
// propagation, component sizes, the hull as two monotone chains over per-row extreme points in exact integer arithmetic (coordinates
// doubled, so the closed-hull test at pixel centres has no rounding), hull bitmap + centroid sums.  Kernel 2: one 1-D reflect correlation
// per axis in the summation order of scipy's symmetric correlate1d (bit-compatible doubles).
#include "common.cuh"

namespace lmr {
namespace mask {

constexpr int NT = 512;

__device__ __forceinline__ double block_sum_d(double v, double* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = NT / 32;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < nw; ++w) t += scratch[w];  // fixed order
    return t;
}

struct Seg {  // the chain segment (r1,c1)-(r2,c2), r1 < r2, spanning a pixel row (doubled, shifted coordinates)
    int r1, c1, r2, c2;
};

__global__ void __launch_bounds__(NT) mask_hull_kernel(const float* __restrict__ mei, int H, int W, float zthr, int iters, unsigned char* __restrict__ hull,
                                                        unsigned* __restrict__ stats) {
    extern __shared__ __align__(16) unsigned char smraw[];
    __shared__ double dscratch[NT / 32];
    __shared__ unsigned long long best_key[NT / 32];
    __shared__ int s_n[2];
    __shared__ unsigned s_acc[3];
    const int HW = H * W, t = threadIdx.x, g = blockIdx.x;
    const int HWp = (HW + 3) & ~3;
    unsigned char* A = smraw;                                            // [HW]  binary plane
    unsigned char* B = A + HWp;                                          // [HW]  binary plane (ping-pong)
    unsigned* cnt = reinterpret_cast<unsigned*>(smraw);                  // [(HW+1)/2] packed 16-bit component sizes; aliases A,B once the labels exist
    unsigned short* lab = reinterpret_cast<unsigned short*>(smraw + 2 * HWp);  // [HW]
    int* ibase = reinterpret_cast<int*>(smraw + 2 * HWp + 2 * (size_t)HWp);
    int* cmin = ibase;                 // [H]
    int* cmax = cmin + H;              // [H]
    int* lo = cmax + H;                // [2H+1]
    int* hi = lo + 2 * H + 1;          // [2H+1]
    int* Lr = hi + 2 * H + 1;          // chains: [2H+1] each
    int* Lc = Lr + 2 * H + 1;
    int* Ur = Lc + 2 * H + 1;
    int* Uc = Ur + 2 * H + 1;
    Seg* segL = reinterpret_cast<Seg*>(Uc + 2 * H + 1);  // [H]
    Seg* segU = segL + H;                                // [H]
    const float* x = mei + (size_t)g * HW;

    // ---- statistics (double) and threshold
    double s = 0.0;
    for (int i = t; i < HW; i += NT) s += (double)x[i];
    const double mean = block_sum_d(s, dscratch) / (double)HW;
    double q = 0.0;
    for (int i = t; i < HW; i += NT) {
        const double u = (double)x[i] - mean;
        q += u * u;
    }
    const double var = block_sum_d(q, dscratch) / (double)HW;
    const float mean_f = (float)mean, std_f = (float)sqrt(var);
    for (int i = t; i < HW; i += NT) A[i] = fabsf((x[i] - mean_f) / std_f) > zthr;
    __syncthreads();

    // ---- binary closing: `iters` dilations, then `iters` erosions (3x3 cross; everything outside the frame is 0)
    unsigned char* src = A;
    unsigned char* dst = B;
    for (int pass = 0; pass < 2 * iters; ++pass) {
        const bool dil = pass < iters;
        for (int i = t; i < HW; i += NT) {
            const int r = i / W, c = i - r * W;
            const int up = r > 0 ? src[i - W] : 0, dn = r < H - 1 ? src[i + W] : 0, lf = c > 0 ? src[i - 1] : 0, rt = c < W - 1 ? src[i + 1] : 0;
            dst[i] = dil ? (src[i] | up | dn | lf | rt) : (src[i] & up & dn & lf & rt);
        }
        __syncthreads();
        unsigned char* tmp = src;
        src = dst, dst = tmp;
    }

    // ---- 8-connected components: every foreground pixel converges to the smallest (raster index + 1) of its component
    for (int i = t; i < HW; i += NT) lab[i] = src[i] ? (unsigned short)(i + 1) : 0;
    __syncthreads();
    for (;;) {
        int changed = 0;
        for (int i = t; i < HW; i += NT) {
            unsigned m = lab[i];
            if (m == 0) continue;
            const int r = i / W, c = i - r * W;
            const unsigned m0 = m;
#pragma unroll
            for (int dr = -1; dr <= 1; ++dr) {
                const int rr = r + dr;
                if (rr < 0 || rr >= H) continue;
#pragma unroll
                for (int dc = -1; dc <= 1; ++dc) {
                    const int cc = c + dc;
                    if (cc < 0 || cc >= W) continue;
                    const unsigned v = lab[rr * W + cc];
                    if (v != 0 && v < m) m = v;
                }
            }
            if (m < m0) lab[i] = (unsigned short)m, changed = 1;  // only the owner thread writes a pixel: stale reads just cost an extra sweep
        }
        if (!__syncthreads_or(changed)) break;
    }

    // ---- component sizes (packed 16-bit counters: sizes <= HW < 65536, no carry) and the largest component
    for (int i = t; i < (HW + 1) / 2; i += NT) cnt[i] = 0u;
    __syncthreads();
    for (int i = t; i < HW; i += NT) {
        const unsigned m = lab[i];
        if (m) atomicAdd(&cnt[(m - 1) >> 1], ((m - 1) & 1) ? 0x10000u : 1u);
    }
    __syncthreads();
    unsigned long long key = 0ull;
    for (int i = t; i < HW; i += NT) {
        if (lab[i] == (unsigned short)(i + 1)) {
            const unsigned n = (cnt[i >> 1] >> ((i & 1) * 16)) & 0xffffu;
            const unsigned long long k = ((unsigned long long)n << 32) | (unsigned long long)(0xffffffffu - (unsigned)i);
            key = k > key ? k : key;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
        key = other > key ? other : key;
    }
    if ((t & 31) == 0) best_key[t >> 5] = key;
    __syncthreads();
    key = 0ull;
    for (int w = 0; w < NT / 32; ++w) key = best_key[w] > key ? best_key[w] : key;
    unsigned char* out = hull + (size_t)g * HW;
    if (key == 0ull) {  // no pixel above threshold: the reference fails here (argmax of an empty sequence); report it
        for (int i = t; i < HW; i += NT) out[i] = 0;
        if (t == 0) stats[4 * g] = 0u, stats[4 * g + 1] = 0u, stats[4 * g + 2] = 0u, stats[4 * g + 3] = 1u;
        return;
    }
    const unsigned short best = (unsigned short)((0xffffffffu - (unsigned)(key & 0xffffffffull)) + 1u);

    // ---- per-row extent of the component, then per doubled row the extreme mid-points: pixel (r,c) contributes (2r+1 +- 1, 2c+1), (2r+1, 2c+1 +- 1)
    for (int r = t; r < H; r += NT) {
        int mn = 1 << 30, mx = -1;
        for (int c = 0; c < W; ++c)
            if (lab[r * W + c] == best) mn = min(mn, c), mx = c;
        cmin[r] = mn, cmax[r] = mx;
    }
    __syncthreads();
    for (int R = t; R <= 2 * H; R += NT) {
        int l = 1 << 30, h = -1;
        if (R & 1) {  // the pixel row r itself: columns 2c+1 +- 1
            const int r = R >> 1;
            if (cmax[r] >= 0) l = 2 * cmin[r], h = 2 * cmax[r] + 2;
        } else {  // between rows R/2 - 1 and R/2: columns 2c+1
            const int r1 = (R >> 1) - 1, r2 = R >> 1;
            if (r1 >= 0 && cmax[r1] >= 0) l = min(l, 2 * cmin[r1] + 1), h = max(h, 2 * cmax[r1] + 1);
            if (r2 < H && cmax[r2] >= 0) l = min(l, 2 * cmin[r2] + 1), h = max(h, 2 * cmax[r2] + 1);
        }
        lo[R] = l, hi[R] = h;
    }
    __syncthreads();
    // monotone chains: thread 0 the left boundary (convex function of the row), thread 32 the right boundary (concave)
    if (t == 0 || t == 32) {
        const bool left = t == 0;
        int* Cr = left ? Lr : Ur;
        int* Cc = left ? Lc : Uc;
        const int* val = left ? lo : hi;
        int n = 0;
        for (int R = 0; R <= 2 * H; ++R) {
            if (hi[R] < 0) continue;
            const int c = val[R];
            while (n >= 2) {
                const long long cross = (long long)(Cr[n - 1] - Cr[n - 2]) * (c - Cc[n - 2]) - (long long)(Cc[n - 1] - Cc[n - 2]) * (R - Cr[n - 2]);
                if (left ? cross <= 0 : cross >= 0) --n;
                else break;
            }
            Cr[n] = R, Cc[n] = c, ++n;
        }
        s_n[left ? 0 : 1] = n;
    }
    if (t < 3) s_acc[t] = 0u;
    __syncthreads();
    for (int r = t; r < H; r += NT) {
        const int Rc = 2 * r + 1;
        Seg a = {0, 0, 0, 0}, b = {0, 0, 0, 0};
        if (Rc >= Lr[0] && Rc <= Lr[s_n[0] - 1]) {
            int k = 0;
            while (Lr[k + 1] < Rc) ++k;
            a = {Lr[k], Lc[k], Lr[k + 1], Lc[k + 1]};
            k = 0;
            while (Ur[k + 1] < Rc) ++k;
            b = {Ur[k], Uc[k], Ur[k + 1], Uc[k + 1]};
        }
        segL[r] = a, segU[r] = b;
    }
    __syncthreads();
    unsigned n_in = 0, sr = 0, sc = 0;
    for (int i = t; i < HW; i += NT) {
        const int r = i / W, c = i - r * W;
        const Seg a = segL[r], b = segU[r];
        bool in = false;
        if (a.r2 > a.r1) {
            const int Rc = 2 * r + 1, Cc = 2 * c + 1;
            in = (long long)(Cc - a.c1) * (a.r2 - a.r1) >= (long long)(a.c2 - a.c1) * (Rc - a.r1) &&
                 (long long)(Cc - b.c1) * (b.r2 - b.r1) <= (long long)(b.c2 - b.c1) * (Rc - b.r1);
        }
        out[i] = in;
        if (in) ++n_in, sr += (unsigned)r, sc += (unsigned)c;
    }
    atomicAdd(&s_acc[0], n_in), atomicAdd(&s_acc[1], sr), atomicAdd(&s_acc[2], sc);
    __syncthreads();
    if (t == 0) stats[4 * g] = s_acc[0], stats[4 * g + 1] = s_acc[1], stats[4 * g + 2] = s_acc[2], stats[4 * g + 3] = 0u;
}

// 1-D correlation along `axis` (0: rows, 1: columns) with a symmetric kernel, scipy 'reflect' boundary (d c b a | a b c d | d c b a);
// doubles, summed in the order of scipy's symmetric correlate1d: centre tap, then the pairs from the outermost inwards.
// w[0..radius]: w[j] = weight of offset +-(radius - j), w[radius] = centre.
template <typename TIn>
__global__ void __launch_bounds__(256) blur_axis_kernel(const TIn* __restrict__ in, float* __restrict__ out, int H, int W, int axis, int radius,
                                                         const double* __restrict__ w) {
    const int HW = H * W;
    const size_t base = (size_t)blockIdx.y * HW;
    const int n = axis == 0 ? H : W;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
        const int r = i / W, c = i - r * W;
        const int p = axis == 0 ? r : c;
        auto at = [&](int q) -> double {
            while (q < 0 || q >= n) q = q < 0 ? -q - 1 : 2 * n - q - 1;
            return (double)in[base + (axis == 0 ? (size_t)q * W + c : (size_t)r * W + q)];
        };
        double acc = at(p) * w[radius];
        for (int j = 0; j < radius; ++j) {
            const int off = radius - j;
            acc += (at(p - off) + at(p + off)) * w[j];
        }
        out[base + i] = (float)acc;
    }
}

// ---- featurevis.ops.GaussianBlur (featurevis/ops.py:378-423): F.pad(mode) followed by a depthwise 1-D correlation per axis, float32.
// One launch per axis; taps accumulate in window order (k = -h .. h) in float32 like a direct convolution.
//   pad_mode 0: constant 0;  1: torch 'reflect' (mirror WITHOUT repeating the edge pixel: -1 -> 1, n -> n - 2);  2: 'replicate' (clamp)
__global__ void __launch_bounds__(256) gblur_axis_kernel(const float* __restrict__ in, float* __restrict__ out, int H, int W, int axis, int h,
                                                          const float* __restrict__ w, int pad_mode, float scale) {
    const int HW = H * W;
    const size_t base = (size_t)blockIdx.y * HW;
    const int n = axis == 0 ? H : W;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
        const int r = i / W, c = i - r * W;
        const int p = axis == 0 ? r : c;
        float acc = 0.f;
        for (int k = -h; k <= h; ++k) {
            int q = p + k;
            float v = 0.f;
            bool inside = q >= 0 && q < n;
            if (!inside && pad_mode == 1) q = q < 0 ? -q : 2 * n - 2 - q, inside = true;
            if (!inside && pad_mode == 2) q = q < 0 ? 0 : n - 1, inside = true;
            if (inside) v = in[base + (axis == 0 ? (size_t)q * W + c : (size_t)r * W + q)];
            acc = fmaf(w[k + h], v, acc);
        }
        out[base + i] = acc * scale;
    }
}

}  // namespace mask
}  // namespace lmr

using namespace lmr;

static size_t mask_smem_bytes(int H, int W) {
    const size_t HWp = ((size_t)H * W + 3) & ~(size_t)3;
    return 2 * HWp + 2 * HWp + sizeof(int) * (size_t)(2 * H + 6 * (2 * H + 1)) + 2 * sizeof(mask::Seg) * (size_t)H + 64;
}

extern "C" size_t lmr_create_mask_workspace_bytes(int G, int H, int W) {
    const size_t HW = (size_t)H * W;
    return align_up((size_t)G * HW, 256) + align_up((size_t)G * HW * sizeof(float), 256);  // hull bitmap + the float plane between the two blur axes
}

extern "C" int lmr_create_mask(const float* mei, int G, int H, int W, float zscore_thresh, int closing_iters, const double* blur_weights,
                               int blur_radius, float* mask_out, unsigned* stats, void* ws, size_t ws_bytes, void* stream) {
    LMR_CHECK_ARG(mei && mask_out && stats && blur_weights, "create_mask: null argument");
    LMR_CHECK_ARG(G >= 1 && G <= 65535 && H >= 2 && W >= 2 && (long long)H * W < 65536, "create_mask: unsupported sizes G=%d H=%d W=%d (H*W < 65536)", G, H, W);
    LMR_CHECK_ARG(closing_iters >= 0 && blur_radius >= 0 && blur_radius <= 64, "create_mask: bad closing_iters / blur radius");
    if (ws_bytes < lmr_create_mask_workspace_bytes(G, H, W)) {
        set_error("create_mask: workspace too small");
        return LMR_ERR_WORKSPACE;
    }
    const size_t smem = mask_smem_bytes(H, W);
    LMR_CHECK_ARG(smem <= 220 * 1024, "create_mask: a %dx%d image needs %zu bytes of shared memory (> 220 KB)", H, W, smem);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    LMR_CUDA_OK(cudaFuncSetAttribute(mask::mask_hull_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unsigned char* hull = static_cast<unsigned char*>(ws);
    float* tmp = reinterpret_cast<float*>(static_cast<unsigned char*>(ws) + align_up((size_t)G * H * W, 256));
    mask::mask_hull_kernel<<<G, mask::NT, smem, st>>>(mei, H, W, zscore_thresh, closing_iters, hull, stats);
    LMR_LAUNCH_OK();
    dim3 grid((unsigned)((H * W + 255) / 256), (unsigned)G);
    mask::blur_axis_kernel<unsigned char><<<grid, 256, 0, st>>>(hull, tmp, H, W, 0, blur_radius, blur_weights);
    LMR_LAUNCH_OK();
    mask::blur_axis_kernel<float><<<grid, 256, 0, st>>>(tmp, mask_out, H, W, 1, blur_radius, blur_weights);
    LMR_LAUNCH_OK();
    return 0;
}

extern "C" int lmr_gaussian_blur(const float* x, int B, int H, int W, const float* wy, int hy, const float* wx, int hx, int pad_mode, float inv_norm,
                                 float* out, void* ws, size_t ws_bytes, void* stream) {
    LMR_CHECK_ARG(x && wy && wx && out && ws, "gaussian_blur: null argument");
    LMR_CHECK_ARG(B >= 1 && B <= 65535 && H >= 1 && W >= 1 && hy >= 0 && hx >= 0, "gaussian_blur: bad sizes B=%d H=%d W=%d hy=%d hx=%d", B, H, W, hy, hx);
    LMR_CHECK_ARG(pad_mode >= 0 && pad_mode <= 2, "gaussian_blur: pad_mode %d (0 constant, 1 reflect, 2 replicate)", pad_mode);
    // torch.nn.functional.pad(mode='reflect') raises when the padding is not smaller than the dimension
    LMR_CHECK_ARG(pad_mode != 1 || (hy < H && hx < W), "gaussian_blur: reflect padding (%d, %d) must be smaller than the image (%d, %d)", hy, hx, H, W);
    if (ws_bytes < (size_t)B * H * W * sizeof(float)) {
        set_error("gaussian_blur: workspace too small");
        return LMR_ERR_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* tmp = static_cast<float*>(ws);
    dim3 grid((unsigned)((H * W + 255) / 256), (unsigned)B);
    mask::gblur_axis_kernel<<<grid, 256, 0, st>>>(x, tmp, H, W, 0, hy, wy, pad_mode, 1.f);
    LMR_LAUNCH_OK();
    mask::gblur_axis_kernel<<<grid, 256, 0, st>>>(tmp, out, H, W, 1, hx, wx, pad_mode, inv_norm);
    LMR_LAUNCH_OK();
    return 0;
}
