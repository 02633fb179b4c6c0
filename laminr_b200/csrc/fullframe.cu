// Full-frame evaluation of a frozen conv response model: every neuron's response to every image, and the gradient of a weighted sum of
// responses back to the images.  Serves the response models whose layers cannot be cut down to one neuron's receptive-field cone
// (csrc/convcone.cu): squeeze-and-excitation blocks pool over the whole frame, and `model(x)` of all neurons shares one core evaluation.
//
// Replaces, for eval-mode models:
//   * SE2dCore.forward            laminr/neuron_models/utils/monkey_model.py:114-317 (first conv, DepthSeparableConv2d hidden layers
//                                 neuralpredictors/layers/conv.py:4-30 with dilation, BatchNorm, ELU, SQ_EX_Block :95-108)
//   * Stacked2dCore.forward       neuralpredictors/layers/cores/conv2d.py:246-253 (dense hidden convolutions)
//   * PointPooled2d.forward       neuralpredictors/layers/readouts/point_pooled.py:121-169 (bilinear reads of the feature map and of its
//                                 average-pooled copies at one position per neuron)
//   * FullGaussian2d.forward      neuralpredictors/layers/readouts/gaussian.py:532-581 in eval mode (= the same read with one scale)
//   * elu(readout + offset) + 1   monkey_models.py:292-294 / firing_rate.py:36-52
// and torch autograd's gradient of all of the above with respect to the images (parameters are frozen: no weight gradients).
//
// A model is a list of layers (dense / depthwise / pointwise convolution with an optional per-channel affine and shifted ELU behind it, or a
// squeeze-excitation block) followed by the multi-scale point readout.  Kernels are fp32 CUDA-core code: the goldens are fp32 and these
// response models are small (32 channels, 93 x 93 images), so the work is latency / L2 bound rather than a tensor-core problem.
//   * dense convolution: each thread owns 8 consecutive output pixels x CO output channels and slides a 16-float register window along
//     the input row, so one 8-tap block costs 2 input + 8 (CO=4) weight shared-memory loads for 256 FMAs
//   * depthwise (dilated) convolution: same 8-pixel register window, one channel plane per CTA tile
//   * pointwise convolution: 2 pixels x 32 output channels per thread, weights broadcast from shared memory
// The data gradient of a convolution is the same kernel run with flipped / transposed weights (packed once on the host side).
#include "common.cuh"

namespace lmr {
namespace ff {

constexpr int NT = 256;  // threads per CTA for the convolution kernels (8 warps)

struct Epi {
    const float* scale;  // [c_out] or nullptr
    const float* shift;
    int elu;
    float xs, ys;
};

__device__ __forceinline__ float epilogue(float v, const Epi& e, int c) {
    if (e.scale) v = fmaf(v, __ldg(e.scale + c), __ldg(e.shift + c));
    if (e.elu) {
        const float t = v - e.xs;
        v = (t > 0.f ? t : expm1f(t)) + e.ys;
    }
    return v;
}

__device__ __forceinline__ void ld4(float* dst, const float* src) {
    const float4 v = *reinterpret_cast<const float4*>(src);
    dst[0] = v.x, dst[1] = v.y, dst[2] = v.z, dst[3] = v.w;
}

// ---------------------------------------------------------------------------------------------------------------------------------------
// dense convolution (stride 1, dilation 1, zero padding `pad`)
// ---------------------------------------------------------------------------------------------------------------------------------------
struct DenseArgs {
    const float* in;  // [B, cin, Hin, Win]
    float* out;       // [B, cout, Hout, Wout]
    const float* w;   // [cin][cout / CO][k][kp][CO], taps kx >= k are zero
    int cin, cout, k, kp, pad, Hin, Win, Hout, Wout;
    int XG, TR;      // a warp covers XG groups of 8 pixels x TR rows (XG * TR <= 32)
    int ncog_blk;    // channel groups per CTA: warp -> (cog = warp % ncog_blk, row block = warp / ncog_blk)
    int tiles_x;
    int tile_w;      // shared-memory row stride = XG * 8 + kp
    Epi epi;
};

template <int CO>
__global__ void __launch_bounds__(NT) ff_dense_kernel(DenseArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int nrb = (NT / 32) / a.ncog_blk;
    const int tile_rows = a.TR * nrb + a.k - 1;
    float* tile = smem;
    float* wsm = smem + tile_rows * a.tile_w;
    const int wcount = a.ncog_blk * a.k * a.kp * CO;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cog = warp % a.ncog_blk, rb = warp / a.ncog_blk;
    const int xg = lane % a.XG, trow = lane / a.XG;
    const bool live = trow < a.TR;
    const int tx = blockIdx.x % a.tiles_x, ty = blockIdx.x / a.tiles_x;
    const int ox0 = tx * a.XG * 8, oy0 = ty * a.TR * nrb;
    const int b = blockIdx.z;
    const int ncog_total = a.cout / CO;
    const int cog0 = blockIdx.y * a.ncog_blk;

    float acc[8][CO];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int c = 0; c < CO; ++c) acc[j][c] = 0.f;

    for (int ci = 0; ci < a.cin; ++ci) {
        __syncthreads();
        const float* plane = a.in + ((size_t)b * a.cin + ci) * a.Hin * a.Win;
        for (int i = tid; i < tile_rows * a.tile_w; i += NT) {
            const int r = i / a.tile_w, c = i - r * a.tile_w;
            const int iy = oy0 + r - a.pad, ix = ox0 + c - a.pad;
            tile[i] = (iy >= 0 && iy < a.Hin && ix >= 0 && ix < a.Win) ? __ldg(plane + (size_t)iy * a.Win + ix) : 0.f;
        }
        const float* wsrc = a.w + ((size_t)ci * ncog_total + cog0) * a.k * a.kp * CO;
        for (int i = tid; i < wcount; i += NT) wsm[i] = __ldg(wsrc + i);
        __syncthreads();
        if (!live) continue;
        const float* wp = wsm + cog * a.k * a.kp * CO;
        for (int ky = 0; ky < a.k; ++ky) {
            const float* irow = tile + (rb * a.TR + trow + ky) * a.tile_w + xg * 8;
            float win[16];
            ld4(win, irow);
            ld4(win + 4, irow + 4);
            for (int kb = 0; kb < a.kp; kb += 8) {
                ld4(win + 8, irow + kb + 8);
                ld4(win + 12, irow + kb + 12);
                const float* wk = wp + (ky * a.kp + kb) * CO;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    float wv[CO];
                    if (CO == 4) {
                        ld4(wv, wk + u * 4);
                    } else {
#pragma unroll
                        for (int c = 0; c < CO; ++c) wv[c] = wk[u * CO + c];
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j)
#pragma unroll
                        for (int c = 0; c < CO; ++c) acc[j][c] = fmaf(win[j + u], wv[c], acc[j][c]);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) win[j] = win[j + 8];
            }
        }
    }
    if (!live) return;
    const int oy = oy0 + rb * a.TR + trow;
    if (oy >= a.Hout) return;
#pragma unroll
    for (int c = 0; c < CO; ++c) {
        const int co = (cog0 + cog) * CO + c;
        float* orow = a.out + (((size_t)b * a.cout + co) * a.Hout + oy) * a.Wout;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int ox = ox0 + xg * 8 + j;
            if (ox < a.Wout) orow[ox] = epilogue(acc[j][c], a.epi, co);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------------------------
// depthwise convolution with dilation (stride 1, zero padding)
// ---------------------------------------------------------------------------------------------------------------------------------------
struct DwArgs {
    const float* in;  // [B*C, Hin, Win]
    float* out;       // [B*C, Hout, Wout]
    const float* w;   // [C][K][K]
    int C, pad, Hin, Win, Hout, Wout;
    int XG, TR, tiles_x, tile_w;
    Epi epi;
};

template <int K, int DIL>
__global__ void __launch_bounds__(NT) ff_dw_kernel(DwArgs a) {
    constexpr int HALO = DIL * (K - 1);
    constexpr int WIN = 8 + ((HALO + 3) / 4) * 4;
    extern __shared__ __align__(16) float smem[];
    const int tile_rows = a.TR * (NT / 32) + HALO;
    float* tile = smem;
    float* wsm = smem + tile_rows * a.tile_w;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int xg = lane % a.XG, trow = lane / a.XG;
    const int tx = blockIdx.x % a.tiles_x, ty = blockIdx.x / a.tiles_x;
    const int ox0 = tx * a.XG * 8, oy0 = ty * a.TR * (NT / 32);
    const int bc = blockIdx.y, c = bc % a.C;
    const float* plane = a.in + (size_t)bc * a.Hin * a.Win;
    for (int i = tid; i < tile_rows * a.tile_w; i += NT) {
        const int r = i / a.tile_w, cc = i - r * a.tile_w;
        const int iy = oy0 + r - a.pad, ix = ox0 + cc - a.pad;
        tile[i] = (iy >= 0 && iy < a.Hin && ix >= 0 && ix < a.Win) ? __ldg(plane + (size_t)iy * a.Win + ix) : 0.f;
    }
    if (tid < K * K) wsm[tid] = __ldg(a.w + c * K * K + tid);
    __syncthreads();
    if (trow >= a.TR) return;
    const int oy = oy0 + warp * a.TR + trow;
    if (oy >= a.Hout) return;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll 1
    for (int ky = 0; ky < K; ++ky) {
        const float* irow = tile + (warp * a.TR + trow + ky * DIL) * a.tile_w + xg * 8;
        float win[WIN];
#pragma unroll
        for (int i = 0; i < WIN; i += 4) ld4(win + i, irow + i);
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
            const float wv = wsm[ky * K + kx];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(win[j + kx * DIL], wv, acc[j]);
        }
    }
    float* orow = a.out + ((size_t)bc * a.Hout + oy) * a.Wout;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int ox = ox0 + xg * 8 + j;
        if (ox < a.Wout) orow[ox] = epilogue(acc[j], a.epi, c);
    }
}

// ---------------------------------------------------------------------------------------------------------------------------------------
// pointwise (1 x 1) convolution: 2 pixels x 32 output channels per thread
// ---------------------------------------------------------------------------------------------------------------------------------------
struct PwArgs {
    const float* in;  // [B, cin, HW]
    float* out;       // [B, cout, HW]
    const float* w;   // [cin][cout]
    int cin, cout, HW;
    Epi epi;
};

__global__ void __launch_bounds__(NT) ff_pw_kernel(PwArgs a) {
    extern __shared__ __align__(16) float wsm[];  // [cin][32]
    const int tid = threadIdx.x, b = blockIdx.z, co0 = blockIdx.y * 32;
    for (int i = tid; i < a.cin * 32; i += NT) {
        const int ci = i >> 5, c = i & 31;
        wsm[i] = co0 + c < a.cout ? __ldg(a.w + ci * a.cout + co0 + c) : 0.f;
    }
    __syncthreads();
    const int p0 = blockIdx.x * (2 * NT) + tid, p1 = p0 + NT;
    const bool v0 = p0 < a.HW, v1 = p1 < a.HW;
    float acc0[32], acc1[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) acc0[c] = acc1[c] = 0.f;
    const float* src = a.in + (size_t)b * a.cin * a.HW;
    for (int ci = 0; ci < a.cin; ++ci) {
        const float x0 = v0 ? __ldg(src + (size_t)ci * a.HW + p0) : 0.f;
        const float x1 = v1 ? __ldg(src + (size_t)ci * a.HW + p1) : 0.f;
        const float* wr = wsm + ci * 32;
#pragma unroll
        for (int c4 = 0; c4 < 32; c4 += 4) {
            float wv[4];
            ld4(wv, wr + c4);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                acc0[c4 + c] = fmaf(x0, wv[c], acc0[c4 + c]);
                acc1[c4 + c] = fmaf(x1, wv[c], acc1[c4 + c]);
            }
        }
    }
    float* dst = a.out + (size_t)b * a.cout * a.HW;
#pragma unroll
    for (int c = 0; c < 32; ++c) {
        if (co0 + c < a.cout) {
            if (v0) dst[(size_t)(co0 + c) * a.HW + p0] = epilogue(acc0[c], a.epi, co0 + c);
            if (v1) dst[(size_t)(co0 + c) * a.HW + p1] = epilogue(acc1[c], a.epi, co0 + c);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------------------------
// squeeze-and-excitation block:  y = x * sigmoid(W2 relu(W1 mean_hw(x) + b1) + b2)
// ---------------------------------------------------------------------------------------------------------------------------------------
// out[bc] = mult * sum_hw a[bc, :] * (b ? b[bc, :] : 1)
__global__ void __launch_bounds__(NT) ff_plane_dot_kernel(const float* a, const float* b, float* out, int HW, float mult) {
    __shared__ float scratch[33];
    const size_t base = (size_t)blockIdx.x * HW;
    float s = 0.f;
    for (int i = threadIdx.x; i < HW; i += NT) s += b ? a[base + i] * b[base + i] : a[base + i];
    s = block_sum(s, scratch);
    if (threadIdx.x == 0) out[blockIdx.x] = s * mult;
}

// one CTA per image; C <= 1024, R <= 64
__global__ void ff_se_gate_kernel(const float* mean, const float* w1, const float* b1, const float* w2, const float* b2, int C, int R, float* hidden,
                                  float* gate) {
    __shared__ float h[64];
    const int b = blockIdx.x;
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        float s = b1[r];
        for (int c = 0; c < C; ++c) s = fmaf(w1[r * C + c], mean[b * C + c], s);
        h[r] = fmaxf(s, 0.f);
        hidden[b * R + r] = h[r];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = b2[c];
        for (int r = 0; r < R; ++r) s = fmaf(w2[c * R + r], h[r], s);
        gate[b * C + c] = 1.f / (1.f + expf(-s));
    }
}

// gm[b,c] = (1/HW) * d(sum_c' dots[c'] * gate[c']) / d mean[c]
__global__ void ff_se_gate_bwd_kernel(const float* dots, const float* gate, const float* hidden, const float* w1, const float* w2, int C, int R,
                                      float inv_hw, float* gm) {
    __shared__ float dh[64];
    const int b = blockIdx.x;
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        float s = 0.f;
        if (hidden[b * R + r] > 0.f)
            for (int c = 0; c < C; ++c) {
                const float g = gate[b * C + c];
                s = fmaf(w2[c * R + r], dots[b * C + c] * g * (1.f - g), s);
            }
        dh[r] = s;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int r = 0; r < R; ++r) s = fmaf(w1[r * C + c], dh[r], s);
        gm[b * C + c] = s * inv_hw;
    }
}

// out[bc, i] = in[bc, i] * gate[bc] + (add ? add[bc] : 0)
__global__ void __launch_bounds__(NT) ff_plane_affine_kernel(const float* in, const float* gate, const float* add, float* out, int HW) {
    const int bc = blockIdx.y;
    const float g = gate[bc], s = add ? add[bc] : 0.f;
    const size_t base = (size_t)bc * HW;
    for (int i = blockIdx.x * NT + threadIdx.x; i < HW; i += gridDim.x * NT) out[base + i] = fmaf(in[base + i], g, s);
}

// gradient through the affine + shifted ELU behind a convolution, in place:  g *= elu'(.) * scale[c]   (elu' recovered from the kept output)
__global__ void __launch_bounds__(NT) ff_act_bwd_kernel(float* g, const float* out, Epi e, int C, int HW) {
    const int bc = blockIdx.y, c = bc % C;
    const float sc = e.scale ? e.scale[c] : 1.f;
    const size_t base = (size_t)bc * HW;
    for (int i = blockIdx.x * NT + threadIdx.x; i < HW; i += gridDim.x * NT) {
        float d = sc;
        if (e.elu) {
            const float t = out[base + i] - e.ys;  // = elu(pre - xs): > 0 on the linear side, expm1(.) <= 0 otherwise
            if (t <= 0.f) d *= t + 1.f;
        }
        g[base + i] *= d;
    }
}

// ---------------------------------------------------------------------------------------------------------------------------------------
// average pooling k x k, stride k, no padding (nn.AvgPool2d(k, stride=k)): Hc = H / k (floor)
// ---------------------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) ff_pool_kernel(const float* in, float* out, int H, int W, int k, int Hc, int Wc) {
    const int bc = blockIdx.y;
    const float inv = 1.f / (float)(k * k);
    for (int i = blockIdx.x * NT + threadIdx.x; i < Hc * Wc; i += gridDim.x * NT) {
        const int y = i / Wc, x = i - y * Wc;
        const float* p = in + ((size_t)bc * H + y * k) * W + x * k;
        float s = 0.f;
        for (int dy = 0; dy < k; ++dy)
            for (int dx = 0; dx < k; ++dx) s += p[dy * W + dx];
        out[(size_t)bc * Hc * Wc + i] = s * inv;
    }
}

// g_fine += unpool(g_coarse)
__global__ void __launch_bounds__(NT) ff_pool_bwd_kernel(const float* gc, float* gf, int H, int W, int k, int Hc, int Wc) {
    const int bc = blockIdx.y;
    const float inv = 1.f / (float)(k * k);
    for (int i = blockIdx.x * NT + threadIdx.x; i < Hc * k * W; i += gridDim.x * NT) {
        const int y = i / W, x = i - y * W;
        if (x < Wc * k) gf[((size_t)bc * H + y) * W + x] += gc[((size_t)bc * Hc + y / k) * Wc + x / k] * inv;
    }
}

// ---------------------------------------------------------------------------------------------------------------------------------------
// multi-scale point readout: one warp per (image, neuron), lanes over channels
// ---------------------------------------------------------------------------------------------------------------------------------------
constexpr int MAX_SCALES = 4;
struct ReadoutArgs {
    const float* maps[MAX_SCALES];  // [B, C, Hs, Ws]
    float* gmaps[MAX_SCALES];       // gradients (backward)
    int Hs[MAX_SCALES], Ws[MAX_SCALES];
    int n_scales, C, n_neurons, B, align_corners;
    const float* grid;      // [n, 2] (x, y) in [-1, 1]
    const float* features;  // [n][n_scales * C]
    const float* bias;      // [n] or nullptr
    float offset;
    float* pre;             // [B, n] readout + offset (kept for the backward)
    float* act;             // [B, n]
    const float* g_act;     // [B, n]
};

struct Taps {
    int x0, y0;
    float wx1, wy1;
};
__device__ __forceinline__ Taps bilinear(float gx, float gy, int H, int W, int align) {
    gx = fminf(fmaxf(gx, -1.f), 1.f), gy = fminf(fmaxf(gy, -1.f), 1.f);
    const float fx = align ? (gx + 1.f) * 0.5f * (float)(W - 1) : ((gx + 1.f) * (float)W - 1.f) * 0.5f;
    const float fy = align ? (gy + 1.f) * 0.5f * (float)(H - 1) : ((gy + 1.f) * (float)H - 1.f) * 0.5f;
    Taps t;
    const float flx = floorf(fx), fly = floorf(fy);
    t.x0 = (int)flx, t.y0 = (int)fly;
    t.wx1 = fx - flx, t.wy1 = fy - fly;
    return t;
}

template <bool BWD>
__global__ void __launch_bounds__(NT) ff_readout_kernel(ReadoutArgs a) {
    const int gw = (blockIdx.x * NT + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (gw >= a.B * a.n_neurons) return;
    const int b = gw / a.n_neurons, n = gw - b * a.n_neurons;
    float gy = 0.f;
    if (BWD) {
        const float g = a.g_act[gw];
        if (g == 0.f) return;
        const float pre = a.pre[gw];
        gy = g * (pre > 0.f ? 1.f : expf(pre));
    }
    const float px = a.grid[2 * n], py = a.grid[2 * n + 1];
    float acc = 0.f;
    for (int s = 0; s < a.n_scales; ++s) {
        const int H = a.Hs[s], W = a.Ws[s];
        const Taps t = bilinear(px, py, H, W, a.align_corners);
        const float wgt[4] = {(1.f - t.wx1) * (1.f - t.wy1), t.wx1 * (1.f - t.wy1), (1.f - t.wx1) * t.wy1, t.wx1 * t.wy1};
        for (int c = lane; c < a.C; c += 32) {
            const float f = a.features[(size_t)n * a.n_scales * a.C + s * a.C + c];
            const size_t plane = ((size_t)b * a.C + c) * H * W;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int x = t.x0 + (q & 1), y = t.y0 + (q >> 1);
                if (x < 0 || x >= W || y < 0 || y >= H || wgt[q] == 0.f) continue;
                if (BWD)
                    atomicAdd(a.gmaps[s] + plane + (size_t)y * W + x, gy * f * wgt[q]);
                else
                    acc = fmaf(a.maps[s][plane + (size_t)y * W + x] * wgt[q], f, acc);
            }
        }
    }
    if (!BWD) {
        acc = warp_sum(acc);
        if (lane == 0) {
            const float pre = acc + (a.bias ? a.bias[n] : 0.f) + a.offset;
            a.pre[gw] = pre;
            a.act[gw] = (pre > 0.f ? pre : expm1f(pre)) + 1.f;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------------------------
struct Dims {
    int C, H, W;
};

static void pick_tiling(int Wout, int Hout, int rows_per_lane_block, int* XG, int* TR) {
    // a warp covers XG*8 pixels x TR rows; fewest CTAs first, then most live lanes
    long best = -1;
    for (int xg = 1; xg <= 32; ++xg) {
        const int tr = 32 / xg;
        const long tiles = (long)((Wout + xg * 8 - 1) / (xg * 8)) * ((Hout + tr * rows_per_lane_block - 1) / (tr * rows_per_lane_block));
        const long cost = tiles * 64 - xg * tr;
        if (best < 0 || cost < best) best = cost, *XG = xg, *TR = tr;
    }
}

static Epi make_epi(const lmr_ff_layer& L) {
    Epi e;
    e.scale = L.affine ? L.scale : nullptr, e.shift = L.affine ? L.shift : nullptr;
    e.elu = L.elu, e.xs = L.elu_xshift, e.ys = L.elu_yshift;
    return e;
}
static Epi no_epi() {
    Epi e;
    e.scale = e.shift = nullptr, e.elu = 0, e.xs = e.ys = 0.f;
    return e;
}

static int dense_co(int cout) { return cout % 4 == 0 ? 4 : 1; }

static int launch_dense(const float* in, float* out, const float* w, int B, int cin, int cout, int k, int pad, Dims di, Dims dout, Epi epi,
                        cudaStream_t st) {
    DenseArgs a;
    a.in = in, a.out = out, a.w = w, a.cin = cin, a.cout = cout, a.k = k, a.kp = (k + 7) / 8 * 8, a.pad = pad;
    a.Hin = di.H, a.Win = di.W, a.Hout = dout.H, a.Wout = dout.W, a.epi = epi;
    const int CO = dense_co(cout), ncog = cout / CO;
    int nb = 1;
    while (nb * 2 <= 8 && ncog % (nb * 2) == 0) nb *= 2;
    a.ncog_blk = nb;
    const int nrb = (NT / 32) / nb;
    pick_tiling(dout.W, dout.H, nrb, &a.XG, &a.TR);
    a.tile_w = a.XG * 8 + a.kp;
    a.tiles_x = (dout.W + a.XG * 8 - 1) / (a.XG * 8);
    const int tiles_y = (dout.H + a.TR * nrb - 1) / (a.TR * nrb);
    const size_t smem = ((size_t)(a.TR * nrb + k - 1) * a.tile_w + (size_t)nb * k * a.kp * CO) * sizeof(float);
    LMR_CHECK_ARG(smem <= 200 * 1024, "full-frame dense convolution: kernel %d x %d with %d channels per CTA needs %zu bytes of shared memory", k, k,
                  nb * CO, smem);
    const dim3 grid(a.tiles_x * tiles_y, ncog / nb, B);
    if (CO == 4) {
        LMR_CUDA_OK(cudaFuncSetAttribute(ff_dense_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ff_dense_kernel<4><<<grid, NT, smem, st>>>(a);
    } else {
        LMR_CUDA_OK(cudaFuncSetAttribute(ff_dense_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ff_dense_kernel<1><<<grid, NT, smem, st>>>(a);
    }
    LMR_LAUNCH_OK();
    return LMR_OK;
}

template <int K, int DIL>
static int launch_dw_t(DwArgs a, int B, cudaStream_t st) {
    constexpr int HALO = DIL * (K - 1);
    pick_tiling(a.Wout, a.Hout, NT / 32, &a.XG, &a.TR);
    a.tile_w = a.XG * 8 + (HALO + 3) / 4 * 4;
    a.tiles_x = (a.Wout + a.XG * 8 - 1) / (a.XG * 8);
    const int tiles_y = (a.Hout + a.TR * (NT / 32) - 1) / (a.TR * (NT / 32));
    const size_t smem = ((size_t)(a.TR * (NT / 32) + HALO) * a.tile_w + K * K) * sizeof(float);
    LMR_CHECK_ARG(smem <= 200 * 1024, "full-frame depthwise convolution needs %zu bytes of shared memory", smem);
    LMR_CUDA_OK(cudaFuncSetAttribute(ff_dw_kernel<K, DIL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ff_dw_kernel<K, DIL><<<dim3(a.tiles_x * tiles_y, B * a.C), NT, smem, st>>>(a);
    LMR_LAUNCH_OK();
    return LMR_OK;
}

static bool dw_supported(int k, int dil) {
    return (dil == 1 && (k == 3 || k == 5 || k == 7 || k == 9)) || (dil == 2 && (k == 3 || k == 5 || k == 7 || k == 9));
}

static int launch_dw(const float* in, float* out, const float* w, int B, int C, int k, int dil, int pad, Dims di, Dims dout, Epi epi, cudaStream_t st) {
    DwArgs a;
    a.in = in, a.out = out, a.w = w, a.C = C, a.pad = pad, a.Hin = di.H, a.Win = di.W, a.Hout = dout.H, a.Wout = dout.W, a.epi = epi;
#define LMR_DW_CASE(K_, D_) \
    if (k == K_ && dil == D_) return launch_dw_t<K_, D_>(a, B, st);
    LMR_DW_CASE(3, 1) LMR_DW_CASE(5, 1) LMR_DW_CASE(7, 1) LMR_DW_CASE(9, 1) LMR_DW_CASE(3, 2) LMR_DW_CASE(5, 2) LMR_DW_CASE(7, 2) LMR_DW_CASE(9, 2)
#undef LMR_DW_CASE
    LMR_CHECK_ARG(false, "full-frame depthwise convolution: kernel %d with dilation %d is not instantiated (3/5/7/9 with dilation 1 or 2)", k, dil);
}

static int launch_pw(const float* in, float* out, const float* w, int B, int cin, int cout, int HW, Epi epi, cudaStream_t st) {
    PwArgs a;
    a.in = in, a.out = out, a.w = w, a.cin = cin, a.cout = cout, a.HW = HW, a.epi = epi;
    const size_t smem = (size_t)cin * 32 * sizeof(float);
    LMR_CHECK_ARG(smem <= 200 * 1024, "full-frame pointwise convolution: %d input channels do not fit shared memory", cin);
    LMR_CUDA_OK(cudaFuncSetAttribute(ff_pw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ff_pw_kernel<<<dim3((HW + 2 * NT - 1) / (2 * NT), (cout + 31) / 32, B), NT, smem, st>>>(a);
    LMR_LAUNCH_OK();
    return LMR_OK;
}

static inline int plane_blocks(int HW) { return (HW + NT * 4 - 1) / (NT * 4); }

// workspace plan: kept layer outputs, pooled maps, squeeze-excitation vectors, readout pre-activations, two gradient buffers + pooled gradients
struct Plan {
    Dims in, out[LMR_FF_MAX_LAYERS];
    size_t out_off[LMR_FF_MAX_LAYERS];                                   // floats
    size_t se_off[LMR_FF_MAX_LAYERS];                                    // mean | gate | hidden | dots | gm   (per SE layer)
    Dims scale_dims[MAX_SCALES];
    size_t pool_off[MAX_SCALES], gpool_off[MAX_SCALES];                  // scale 0 = the last layer's output / gradient buffer A
    size_t pre_off, gA_off, gB_off, total;
};

static int make_plan(const lmr_ff_net* net, int B, Plan* p) {
    LMR_CHECK_ARG(net && net->n_layers >= 1 && net->n_layers <= LMR_FF_MAX_LAYERS, "full-frame model: 1..%d layers", LMR_FF_MAX_LAYERS);
    LMR_CHECK_ARG(net->n_scales >= 1 && net->n_scales <= MAX_SCALES, "full-frame readout: 1..%d scales", MAX_SCALES);
    LMR_CHECK_ARG(B >= 1 && net->n_neurons >= 1, "full-frame model: empty batch or readout");
    p->in = Dims{net->c_in, net->height, net->width};
    Dims d = p->in;
    size_t off = 0, max_elems = (size_t)d.C * d.H * d.W;
    for (int l = 0; l < net->n_layers; ++l) {
        const lmr_ff_layer& L = net->layers[l];
        LMR_CHECK_ARG(L.c_in == d.C, "full-frame layer %d expects %d input channels, gets %d", l, L.c_in, d.C);
        Dims o = d;
        switch (L.kind) {
            case LMR_FF_DENSE:
                LMR_CHECK_ARG(L.dilation == 1, "full-frame dense convolution: dilation 1 only (layer %d)", l);
                o = Dims{L.c_out, d.H + 2 * L.pad - (L.k - 1), d.W + 2 * L.pad - (L.k - 1)};
                LMR_CHECK_ARG(L.pad <= L.k - 1, "full-frame dense convolution: padding above k - 1 (layer %d)", l);
                break;
            case LMR_FF_DEPTHWISE:
                LMR_CHECK_ARG(L.c_out == L.c_in && dw_supported(L.k, L.dilation), "full-frame depthwise layer %d: kernel %d / dilation %d unsupported", l, L.k,
                              L.dilation);
                o = Dims{L.c_out, d.H + 2 * L.pad - L.dilation * (L.k - 1), d.W + 2 * L.pad - L.dilation * (L.k - 1)};
                LMR_CHECK_ARG(L.pad <= L.dilation * (L.k - 1), "full-frame depthwise convolution: padding above the kernel extent (layer %d)", l);
                break;
            case LMR_FF_POINTWISE:
                o = Dims{L.c_out, d.H, d.W};
                break;
            case LMR_FF_SE:
                LMR_CHECK_ARG(L.c_out == L.c_in && L.se_hidden >= 1 && L.se_hidden <= 64 && L.c_in <= 1024, "full-frame squeeze-excitation layer %d: sizes", l);
                break;
            default:
                LMR_CHECK_ARG(false, "full-frame layer %d: unknown kind %d", l, L.kind);
        }
        LMR_CHECK_ARG(o.H >= 1 && o.W >= 1, "full-frame layer %d leaves no pixels", l);
        p->out[l] = o, p->out_off[l] = off;
        off += align_up((size_t)B * o.C * o.H * o.W, 64);
        if (L.kind == LMR_FF_SE) {
            p->se_off[l] = off;
            off += align_up((size_t)B * (4 * o.C + L.se_hidden), 64);
        }
        max_elems = max_elems > (size_t)o.C * o.H * o.W ? max_elems : (size_t)o.C * o.H * o.W;
        d = o;
    }
    p->scale_dims[0] = d, p->pool_off[0] = p->out_off[net->n_layers - 1];
    for (int s = 1; s < net->n_scales; ++s) {
        const Dims f = p->scale_dims[s - 1];
        LMR_CHECK_ARG(net->pool_kern >= 1 && f.H / net->pool_kern >= 1 && f.W / net->pool_kern >= 1, "full-frame readout: pooling step %d leaves no pixels", s);
        p->scale_dims[s] = Dims{f.C, f.H / net->pool_kern, f.W / net->pool_kern};
        p->pool_off[s] = off;
        off += align_up((size_t)B * f.C * p->scale_dims[s].H * p->scale_dims[s].W, 64);
    }
    for (int s = 1; s < net->n_scales; ++s) {
        p->gpool_off[s] = off;
        off += align_up((size_t)B * d.C * p->scale_dims[s].H * p->scale_dims[s].W, 64);
    }
    p->pre_off = off, off += align_up((size_t)B * net->n_neurons, 64);
    p->gA_off = off, off += align_up((size_t)B * max_elems, 64);
    p->gB_off = off, off += align_up((size_t)B * max_elems, 64);
    p->gpool_off[0] = p->gA_off;
    p->total = off;
    return LMR_OK;
}

static int forward(const lmr_ff_net* net, const float* img, int B, float* act, float* ws, const Plan& p, cudaStream_t st) {
    const float* cur = img;
    Dims d = p.in;
    for (int l = 0; l < net->n_layers; ++l) {
        const lmr_ff_layer& L = net->layers[l];
        float* out = ws + p.out_off[l];
        const Dims o = p.out[l];
        int rc = LMR_OK;
        switch (L.kind) {
            case LMR_FF_DENSE:
                rc = launch_dense(cur, out, L.w_fwd, B, L.c_in, L.c_out, L.k, L.pad, d, o, make_epi(L), st);
                break;
            case LMR_FF_DEPTHWISE:
                rc = launch_dw(cur, out, L.w_fwd, B, L.c_in, L.k, L.dilation, L.pad, d, o, make_epi(L), st);
                break;
            case LMR_FF_POINTWISE:
                rc = launch_pw(cur, out, L.w_fwd, B, L.c_in, L.c_out, d.H * d.W, make_epi(L), st);
                break;
            case LMR_FF_SE: {
                float* mean = ws + p.se_off[l];
                float *gate = mean + (size_t)B * o.C, *hidden = mean + (size_t)4 * B * o.C;
                const int HW = o.H * o.W;
                ff_plane_dot_kernel<<<B * o.C, NT, 0, st>>>(cur, nullptr, mean, HW, 1.f / (float)HW);
                ff_se_gate_kernel<<<B, 128, 0, st>>>(mean, L.se_w1, L.se_b1, L.se_w2, L.se_b2, o.C, L.se_hidden, hidden, gate);
                ff_plane_affine_kernel<<<dim3(plane_blocks(HW), B * o.C), NT, 0, st>>>(cur, gate, nullptr, out, HW);
                LMR_LAUNCH_OK();
                break;
            }
        }
        if (rc != LMR_OK) return rc;
        cur = out, d = o;
    }
    ReadoutArgs r;
    for (int s = 0; s < net->n_scales; ++s) {
        const Dims sd = p.scale_dims[s];
        r.maps[s] = ws + p.pool_off[s], r.gmaps[s] = nullptr, r.Hs[s] = sd.H, r.Ws[s] = sd.W;
        if (s > 0) {
            const Dims f = p.scale_dims[s - 1];
            ff_pool_kernel<<<dim3(plane_blocks(sd.H * sd.W), B * sd.C), NT, 0, st>>>(ws + p.pool_off[s - 1], ws + p.pool_off[s], f.H, f.W, net->pool_kern, sd.H,
                                                                                    sd.W);
        }
    }
    r.n_scales = net->n_scales, r.C = d.C, r.n_neurons = net->n_neurons, r.B = B, r.align_corners = net->align_corners;
    r.grid = net->grid, r.features = net->features, r.bias = net->bias, r.offset = net->elu_offset;
    r.pre = ws + p.pre_off, r.act = act, r.g_act = nullptr;
    const long warps = (long)B * net->n_neurons;
    ff_readout_kernel<false><<<(unsigned)((warps * 32 + NT - 1) / NT), NT, 0, st>>>(r);
    LMR_LAUNCH_OK();
    return LMR_OK;
}

static int backward(const lmr_ff_net* net, const float* g_act, int B, float* g_img, float* ws, const Plan& p, cudaStream_t st) {
    const int Lr = net->n_layers;
    const Dims dl = p.out[Lr - 1];
    // readout: scatter into the per-scale gradient maps, then fold the pooled gradients down to the finest map
    ReadoutArgs r;
    for (int s = 0; s < net->n_scales; ++s) {
        const Dims sd = p.scale_dims[s];
        r.maps[s] = nullptr, r.gmaps[s] = ws + p.gpool_off[s], r.Hs[s] = sd.H, r.Ws[s] = sd.W;
        LMR_CUDA_OK(cudaMemsetAsync(ws + p.gpool_off[s], 0, (size_t)B * sd.C * sd.H * sd.W * sizeof(float), st));
    }
    r.n_scales = net->n_scales, r.C = dl.C, r.n_neurons = net->n_neurons, r.B = B, r.align_corners = net->align_corners;
    r.grid = net->grid, r.features = net->features, r.bias = net->bias, r.offset = net->elu_offset;
    r.pre = ws + p.pre_off, r.act = nullptr, r.g_act = g_act;
    const long warps = (long)B * net->n_neurons;
    ff_readout_kernel<true><<<(unsigned)((warps * 32 + NT - 1) / NT), NT, 0, st>>>(r);
    for (int s = net->n_scales - 1; s >= 1; --s) {
        const Dims f = p.scale_dims[s - 1], c = p.scale_dims[s];
        ff_pool_bwd_kernel<<<dim3(plane_blocks(c.H * net->pool_kern * f.W), B * f.C), NT, 0, st>>>(ws + p.gpool_off[s], ws + p.gpool_off[s - 1], f.H, f.W,
                                                                                                 net->pool_kern, c.H, c.W);
    }
    LMR_LAUNCH_OK();
    float *g = ws + p.gA_off, *other = ws + p.gB_off;
    for (int l = Lr - 1; l >= 0; --l) {
        const lmr_ff_layer& L = net->layers[l];
        const Dims o = p.out[l], di = l ? p.out[l - 1] : p.in;
        const float* x_in = l ? ws + p.out_off[l - 1] : nullptr;
        float* dst = l ? other : g_img;
        int rc = LMR_OK;
        if (L.kind != LMR_FF_SE && (L.affine || L.elu)) {
            ff_act_bwd_kernel<<<dim3(plane_blocks(o.H * o.W), B * o.C), NT, 0, st>>>(g, ws + p.out_off[l], make_epi(L), o.C, o.H * o.W);
            LMR_LAUNCH_OK();
        }
        switch (L.kind) {
            case LMR_FF_DENSE:
                LMR_CHECK_ARG(L.w_bwd, "full-frame backward: layer %d has no gradient weights", l);
                rc = launch_dense(g, dst, L.w_bwd, B, L.c_out, L.c_in, L.k, L.k - 1 - L.pad, o, di, no_epi(), st);
                break;
            case LMR_FF_DEPTHWISE:
                LMR_CHECK_ARG(L.w_bwd, "full-frame backward: layer %d has no gradient weights", l);
                rc = launch_dw(g, dst, L.w_bwd, B, L.c_in, L.k, L.dilation, L.dilation * (L.k - 1) - L.pad, o, di, no_epi(), st);
                break;
            case LMR_FF_POINTWISE:
                LMR_CHECK_ARG(L.w_bwd, "full-frame backward: layer %d has no gradient weights", l);
                rc = launch_pw(g, dst, L.w_bwd, B, L.c_out, L.c_in, o.H * o.W, no_epi(), st);
                break;
            case LMR_FF_SE: {
                LMR_CHECK_ARG(l > 0, "full-frame backward: a squeeze-excitation block cannot be the first layer");
                float* mean = ws + p.se_off[l];
                float *gate = mean + (size_t)B * o.C, *dots = mean + (size_t)2 * B * o.C, *gm = mean + (size_t)3 * B * o.C, *hidden = mean + (size_t)4 * B * o.C;
                const int HW = o.H * o.W;
                ff_plane_dot_kernel<<<B * o.C, NT, 0, st>>>(g, x_in, dots, HW, 1.f);
                ff_se_gate_bwd_kernel<<<B, 128, 0, st>>>(dots, gate, hidden, L.se_w1, L.se_w2, o.C, L.se_hidden, 1.f / (float)HW, gm);
                ff_plane_affine_kernel<<<dim3(plane_blocks(HW), B * o.C), NT, 0, st>>>(g, gate, gm, dst, HW);
                LMR_LAUNCH_OK();
                break;
            }
        }
        if (rc != LMR_OK) return rc;
        other = g, g = dst;
    }
    return LMR_OK;
}

}  // namespace ff
}  // namespace lmr

extern "C" {

size_t lmr_ff_workspace_bytes(const lmr_ff_net* net, int B) {
    lmr::ff::Plan p;
    if (lmr::ff::make_plan(net, B, &p) != LMR_OK) return 0;
    return p.total * sizeof(float);
}

int lmr_ff_forward(const lmr_ff_net* net, const float* img, int B, float* act, void* workspace, size_t workspace_bytes, void* stream) {
    lmr::ff::Plan p;
    const int rc = lmr::ff::make_plan(net, B, &p);
    if (rc != LMR_OK) return rc;
    LMR_CHECK_ARG(img && act && workspace, "lmr_ff_forward: null pointer");
    LMR_CHECK_ARG(workspace_bytes >= p.total * sizeof(float), "lmr_ff_forward: workspace of %zu bytes, %zu needed", workspace_bytes, p.total * sizeof(float));
    return lmr::ff::forward(net, img, B, act, (float*)workspace, p, (cudaStream_t)stream);
}

int lmr_ff_backward(const lmr_ff_net* net, const float* g_act, int B, float* g_img, void* workspace, size_t workspace_bytes, void* stream) {
    lmr::ff::Plan p;
    const int rc = lmr::ff::make_plan(net, B, &p);
    if (rc != LMR_OK) return rc;
    LMR_CHECK_ARG(g_act && g_img && workspace, "lmr_ff_backward: null pointer");
    LMR_CHECK_ARG(workspace_bytes >= p.total * sizeof(float), "lmr_ff_backward: workspace of %zu bytes, %zu needed", workspace_bytes, p.total * sizeof(float));
    return lmr::ff::backward(net, g_act, B, g_img, (float*)workspace, p, (cudaStream_t)stream);
}

}  // extern "C"
