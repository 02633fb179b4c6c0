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
// Convolutions run as register-blocked items (CB output channels x XB adjacent pixels per thread; the K weight vectors and the XB+K-1 input
// values of a kernel row are loaded up front and the K*CB*XB FMAs run from registers), split over input-channel ranges when a layer has few
// pixels; the partial sums are combined in a fixed order (deterministic).  The data-gradient convolutions reuse the same routine on flipped,
// transposed weights with a virtual zero border.
//
// When there are fewer (image, neuron) pairs than SMs (the learn stage: 20 images) each pair is spread over a thread-block CLUSTER: every CTA
// of the cluster computes a contiguous slice of each layer's (pixel block, channel block) items and writes its finished activations /
// gradients into the shared memory of ALL CTAs of the cluster (distributed shared memory), with one cluster barrier per layer.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace lmr {
namespace cone {

constexpr int NT = 512;
constexpr int MAXL = LMR_CONV_MAX_LAYERS;
constexpr int TILE = 16;  // outputs per item: CB channels x XB pixels
#ifndef LMR_CONE_UNROLL
#define LMR_CONE_UNROLL 1
#endif
constexpr int ROW_UNROLL = LMR_CONE_UNROLL;
#ifndef LMR_CONE_ITEMS_INLINE
#define LMR_CONE_ITEMS_INLINE __device__ __forceinline__
#endif

__host__ __device__ inline int pick_cb(int co) { return (co % 4 == 0) ? 4 : (co % 2 == 0) ? 2 : 1; }
// number of (pixel block, channel block) items of a convolution with So x So outputs and Co channels
__host__ __device__ inline int flat_items(int So, int Co) {
    const int cb = pick_cb(Co), xb = TILE / cb;
    return So * ((So + xb - 1) / xb) * (Co / cb);
}

struct Plan {  // built on the host, handed to the kernel inside Args (parameter space: indexed without a local-memory copy)
    int S[MAXL];     // output window edge of layer l
    int halo[MAXL];  // window origin of layer l's output = top-left readout tap - halo[l]  (sum of k/2 of the layers above)
    int halo_in;     // the same for the input window
    int Sin;         // input window edge
    int GS[MAXL];    // row stride of layer l's gradient map: S[l] + 2 (k_l - 1) -- k_l - 1 zeros on either side of every row, so that the
                     // data-gradient ("full") correlation reads its rows without bounds checks -- or S[l] (unpadded, see make_plan)
    int off_in, off_a[MAXL], off_g[MAXL], off_gin, off_scratch, total;  // float offsets into dynamic shared memory
};

__host__ __device__ inline void make_plan(const lmr_convcore_desc& d, bool with_grad, Plan& p, bool allow_pad = true) {
    const int L = d.n_layers;
    p.S[L - 1] = 2, p.halo[L - 1] = 0;
    for (int l = L - 2; l >= 0; --l) p.S[l] = p.S[l + 1] + d.kernel[l + 1] - 1, p.halo[l] = p.halo[l + 1] + d.kernel[l + 1] / 2;
    p.Sin = p.S[0] + d.kernel[0] - 1, p.halo_in = p.halo[0] + d.kernel[0] / 2;
    int off = 0;
    auto take = [&](int n) { int o = off; off += (n + 3) & ~3; return o; };
    p.off_in = take(d.c_in * p.Sin * p.Sin);
    for (int l = 0; l < L; ++l) p.off_a[l] = take(d.hidden * p.S[l] * p.S[l]);
    int items = NT;  // partial-sum scratch: TILE floats per (item, channel split); splits never push the count beyond max(NT, items)
    for (int l = 0; l < L; ++l) {
        const int n = flat_items(p.S[l], d.hidden);
        items = n > items ? n : items;
    }
    if (with_grad) {
        for (int l = 0; l < L; ++l) {
            // padded rows where the data-gradient convolution that reads the map runs 4-channel blocks (the bulk of the work); the others keep
            // plain rows and check their bounds: shared memory left over is L1 cache, and the weight rows are shared between warps through L1
            const bool pad = allow_pad && pick_cb(l == 0 ? d.c_in : d.hidden) == 4;
            p.GS[l] = p.S[l] + (pad ? 2 * (d.kernel[l] - 1) : 0), p.off_g[l] = take(d.hidden * p.S[l] * p.GS[l]);
        }
        p.off_gin = take(d.c_in * p.Sin * p.Sin);
        for (int l = 0; l < L; ++l) {
            const int n = flat_items(l == 0 ? p.Sin : p.S[l - 1], l == 0 ? d.c_in : d.hidden);
            items = n > items ? n : items;
        }
    } else {
        for (int l = 0; l < L; ++l) p.off_g[l] = 0, p.GS[l] = 0;
        p.off_gin = 0;
    }
    p.off_scratch = take(items * TILE);
    p.total = off + 64;  // slack: unchecked forward reads of ragged pixel blocks may run a few floats past the last buffer
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
    Plan pl;
};

// Partial sums of a stride-1 correlation: out[co, y, x] = sum_{ci in split, ky, kx} in[ci][y + ky - vp][x + kx] * W[ci, ky, kx, co] over an input of
// `rows` rows with row stride RS and plane stride PS; rows outside [0, rows) count as 0 (vp = 0: "valid" forward convolution of the window;
// vp = k-1: "full" data-gradient convolution; its input rows either carry k-1 zeros on both sides (xlim = 0: no bounds in x at all) or are
// unpadded rows of xlim values, bounds-checked per load).
// This CTA computes the items [lo, hi) of the flat (pixel block, channel block) space (channel block fastest), each split over `ksplit`
// input-channel ranges; scratch[(ks*(hi-lo) + item-lo)*TILE + c*XB + j] receives the partial sums.
// Weights: Wt[ci][ky][co / CB][kx][CB] -- the K*CB values a thread needs for one (ci, ky) row are contiguous: one base pointer, immediate offsets.
// K > 0: kernel size known at compile time: the weights of row r+1 are loaded (second register set) before the K*CB*XB FMAs of row r run;
// K == 0: any size (sliding window).
template <int CB, int K>
__device__ __forceinline__ void load_wrow(const float* __restrict__ wp, float (&w)[K > 0 ? K : 1][CB]) {
#pragma unroll
    for (int kx = 0; kx < K; ++kx) {
        if constexpr (CB == 4) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(wp) + kx);
            w[kx][0] = q.x, w[kx][1] = q.y, w[kx][2] = q.z, w[kx][3] = q.w;
        } else if constexpr (CB == 2) {
            const float2 q = __ldg(reinterpret_cast<const float2*>(wp) + kx);
            w[kx][0] = q.x, w[kx][1] = q.y;
        } else {
            w[kx][0] = __ldg(wp + kx);
        }
    }
}
// row = first input value of the block (x index xb of an unpadded row when xlim > 0: values outside [0, xlim) count as 0)
template <int CB, int K>
__device__ __forceinline__ void fma_row(const float* __restrict__ row, int xb, int xlim, const float (&w)[K > 0 ? K : 1][CB], float (&acc)[CB][TILE / CB]) {
    constexpr int XB = TILE / CB, RW = XB + K - 1;
    float r[RW];  // ragged pixel blocks read (finite or not) junk that only reaches accumulators which are never stored
    if (xlim > 0) {
#pragma unroll
        for (int j = 0; j < RW; ++j) r[j] = (unsigned)(xb + j) < (unsigned)xlim ? row[j] : 0.f;
    } else {
#pragma unroll
        for (int j = 0; j < RW; ++j) r[j] = row[j];
    }
#pragma unroll
    for (int kx = 0; kx < K; ++kx)
#pragma unroll
        for (int c = 0; c < CB; ++c)
#pragma unroll
            for (int j = 0; j < XB; ++j) acc[c][j] = fmaf(w[kx][c], r[j + kx], acc[c][j]);
}

template <int CB, int K>
LMR_CONE_ITEMS_INLINE void conv_items(const float* __restrict__ in, int Cin, int RS, int PS, int rows, int vp, int xlim, int So, int Co, int kdyn,
                                           const float* __restrict__ Wt, float* __restrict__ scratch, int lo, int hi, int ksplit) {
    constexpr int XB = TILE / CB;
    const int k = K > 0 ? K : kdyn;
    const int nseg = (So + XB - 1) / XB, ncb = Co / CB;
    const int nloc = hi - lo, nitems = nloc * ksplit;
    const int wstep = ncb * k * CB;  // floats from one (ci, ky) weight row to the next
    for (int item = threadIdx.x; item < nitems; item += NT) {
        const int ks = item / nloc, fl = lo + item - ks * nloc;
        const int sp = fl / ncb, cb = fl - sp * ncb;
        const int y = sp / nseg, x0 = (sp - y * nseg) * XB;
        const int c_lo = ks * Cin / ksplit, c_hi = (ks + 1) * Cin / ksplit;
        float acc[CB][XB];
#pragma unroll
        for (int c = 0; c < CB; ++c)
#pragma unroll
            for (int j = 0; j < XB; ++j) acc[c][j] = 0.f;
        const int ky_lo = max(0, vp - y), ky_hi = min(k, vp - y + rows), nky = ky_hi - ky_lo;
        const int nrows = nky > 0 ? (c_hi - c_lo) * nky : 0;
        const float* wp = Wt + ((size_t)(c_lo * k + ky_lo) * ncb + cb) * (k * CB);
        const int xb = xlim > 0 ? x0 - vp : x0;  // unpadded rows (xlim > 0): the block starts vp values to the left, bounds checked per value
        const float* rp = in + c_lo * PS + (y + ky_lo - vp) * RS + xb;
        const int wskip = (k - nky) * wstep, rskip = PS - nky * RS;  // extra steps when the row index wraps to the next input channel
        int kyc = 0;
        auto advance = [&]() {
            wp += wstep, rp += RS;
            if (++kyc == nky) kyc = 0, wp += wskip, rp += rskip;
        };
        if constexpr (K > 0) {
#pragma unroll ROW_UNROLL
            for (int rr = 0; rr < nrows; ++rr) {
                float w0[K][CB];
                load_wrow<CB, K>(wp, w0);
                fma_row<CB, K>(rp, xb, xlim, w0, acc);
                advance();
            }
        } else {
#pragma unroll 1
            for (int rr = 0; rr < nrows; ++rr) {
                float r[XB];
                auto ld = [&](int j) { return (xlim <= 0 || (unsigned)(xb + j) < (unsigned)xlim) ? rp[j] : 0.f; };
#pragma unroll
                for (int j = 0; j < XB - 1; ++j) r[j + 1] = ld(j);
#pragma unroll 2
                for (int kx = 0; kx < k; ++kx) {
#pragma unroll
                    for (int j = 0; j < XB - 1; ++j) r[j] = r[j + 1];
                    r[XB - 1] = ld(kx + XB - 1);
                    float wv[CB];
#pragma unroll
                    for (int c = 0; c < CB; ++c) wv[c] = __ldg(wp + kx * CB + c);
#pragma unroll
                    for (int c = 0; c < CB; ++c)
#pragma unroll
                        for (int j = 0; j < XB; ++j) acc[c][j] = fmaf(wv[c], r[j], acc[c][j]);
                }
                advance();
            }
        }
        float4* o = reinterpret_cast<float4*>(scratch + (size_t)item * TILE);
        const float* af = &acc[0][0];
#pragma unroll
        for (int u = 0; u < TILE / 4; ++u) o[u] = make_float4(af[4 * u], af[4 * u + 1], af[4 * u + 2], af[4 * u + 3]);
    }
}

template <int CB>
__device__ __forceinline__ void conv_by_k(const float* in, int Cin, int RS, int PS, int rows, int vp, int xlim, int So, int Co, int k, const float* Wt,
                                          float* scratch, int lo, int hi, int ksplit) {
    switch (k) {
        case 3: conv_items<CB, 3>(in, Cin, RS, PS, rows, vp, xlim, So, Co, k, Wt, scratch, lo, hi, ksplit); break;
        case 5: conv_items<CB, 5>(in, Cin, RS, PS, rows, vp, xlim, So, Co, k, Wt, scratch, lo, hi, ksplit); break;
        case 7: conv_items<CB, 7>(in, Cin, RS, PS, rows, vp, xlim, So, Co, k, Wt, scratch, lo, hi, ksplit); break;
        case 9: conv_items<CB, 9>(in, Cin, RS, PS, rows, vp, xlim, So, Co, k, Wt, scratch, lo, hi, ksplit); break;
        default: conv_items<CB, 0>(in, Cin, RS, PS, rows, vp, xlim, So, Co, k, Wt, scratch, lo, hi, ksplit);
    }
}

__device__ __forceinline__ void conv_dispatch(const float* in, int Cin, int RS, int PS, int rows, int vp, int xlim, int So, int Co, int k, const float* Wt,
                                              float* scratch, int lo, int hi, int ksplit) {
    const int cb = pick_cb(Co);
    if (cb == 4) conv_by_k<4>(in, Cin, RS, PS, rows, vp, xlim, So, Co, k, Wt, scratch, lo, hi, ksplit);
    else if (cb == 2) conv_by_k<2>(in, Cin, RS, PS, rows, vp, xlim, So, Co, k, Wt, scratch, lo, hi, ksplit);
    else conv_by_k<1>(in, Cin, RS, PS, rows, vp, xlim, So, Co, k, Wt, scratch, lo, hi, ksplit);
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

enum { EPI_FWD = 0, EPI_BWD = 1, EPI_PLAIN = 2 };

// Combines the channel-split partial sums of this CTA's items in a fixed order, applies the layer epilogue and writes the results into `dst`
// (layout [Co][So][So]) of EVERY CTA of the cluster.
//   EPI_FWD:   v*scale[c] + shift[c], optional shifted ELU, 0 outside the frame (= the next layer's zero padding)
//   EPI_BWD:   v * scale[c] * elu'(aout) , 0 outside the frame   (gradient wrt the pre-batch-norm output of the layer below)
//   EPI_PLAIN: v
// `dst` rows have stride DS and start dpad floats in (activations: DS = So, dpad = 0; gradient maps: Plan::GS, k - 1).
__device__ __forceinline__ void reduce_store(cg::cluster_group& cl, int CS, const float* scratch, int lo, int hi, int ksplit, int So, int Co, float* dst, int DS,
                                             int dpad, int mode,
                                             const float* __restrict__ scale, const float* __restrict__ shift, int nonlin, float xs, float ys,
                                             const float* aout, int oy, int ox, int H, int W) {
    const int CB = pick_cb(Co), XB = TILE / CB;
    const int nseg = (So + XB - 1) / XB, ncb = Co / CB, nloc = hi - lo;
    for (int e = threadIdx.x; e < nloc * TILE; e += NT) {
        const int li = e / TILE, q = e - li * TILE, cc = q / XB, j = q - cc * XB;
        const int fl = lo + li, sp = fl / ncb, cb = fl - sp * ncb;
        const int y = sp / nseg, x = (sp - y * nseg) * XB + j;
        if (x >= So) continue;
        const int c = cb * CB + cc;
        float v = 0.f;
        for (int s = 0; s < ksplit; ++s) v += scratch[(size_t)(s * nloc + li) * TILE + q];
        const int idx = (c * So + y) * So + x;
        if (mode != EPI_PLAIN) {
            if (mode == EPI_FWD) {
                v = fmaf(v, __ldg(scale + c), __ldg(shift + c));
                if (nonlin) v = elu_shifted(v, xs, ys);
            } else {
                v *= __ldg(scale + c) * (nonlin ? elu_grad_from_out(aout[idx], ys) : 1.f);
            }
            const int gy = oy + y, gx = ox + x;
            if (!(gy >= 0 && gy < H && gx >= 0 && gx < W)) v = 0.f;
        }
        const int didx = (c * So + y) * DS + x + dpad;  // (gradient maps: padded rows)
        if (CS == 1) dst[didx] = v;
        else
            for (int r = 0; r < CS; ++r) *cl.map_shared_rank(dst + didx, r) = v;
    }
}

__device__ __forceinline__ int ksplit_for(int nloc, int Cin) {
    int ks = nloc > 0 ? NT / nloc : 1;
    if (ks < 1) ks = 1;
    if (ks > Cin) ks = Cin;
    return ks;
}

__global__ void __launch_bounds__(NT, 1) convcone_kernel(const Args a) {
    extern __shared__ __align__(16) float sm[];
    __shared__ float s_tap[4];
    __shared__ float s_red[40];
    cg::cluster_group cl = cg::this_cluster();
    const int CS = (int)cl.num_blocks(), rank = (int)cl.block_rank();
    const lmr_convcore_desc& d = a.d;
    const int L = d.n_layers, H = d.height, W = d.width, P = H * W, Hc = d.hidden, Cin = d.c_in;
    const bool with_grad = a.g_img != nullptr;
    const Plan& pl = a.pl;
    const int pair = blockIdx.x / CS;  // (group, image) pair of this cluster
    const int g = pair / a.NB, b = pair - g * a.NB;
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

    // window origins (frame coordinates): layer l's output starts at (Y0 - halo[l], X0 - halo[l]), the input window at (iy0, ix0)
    const int iy0 = Y0 - pl.halo_in, ix0 = X0 - pl.halo_in;
    const int Sin = pl.Sin;

    // ---- input window (every CTA of the cluster keeps its own copy)
    float* win = sm + pl.off_in;
    for (int i = tid; i < Cin * Sin * Sin; i += NT) {
        const int c = i / (Sin * Sin), r = i - c * Sin * Sin, y = r / Sin, x = r - y * Sin;
        const int gy = iy0 + y, gx = ix0 + x;
        win[i] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? __ldg(img + (size_t)c * P + gy * W + gx) : 0.f;
    }
    for (int i = tid; i < 64; i += NT) sm[pl.total - 64 + i] = 0.f;
    if (with_grad)  // gradient maps: the row borders must read as zeros (the interiors are overwritten layer by layer)
        for (int i = pl.off_g[0] + tid; i < pl.off_gin; i += NT) sm[i] = 0.f;
    cl.sync();  // also: every CTA of the cluster is running before the first remote shared-memory write

    // ---- forward through the layers
    float* scratch = sm + pl.off_scratch;
    for (int l = 0; l < L; ++l) {
        const int cin = l == 0 ? Cin : Hc, So = pl.S[l], Si = l == 0 ? Sin : pl.S[l - 1], k = d.kernel[l];
        const float* in = l == 0 ? win : sm + pl.off_a[l - 1];
        const int nflat = flat_items(So, Hc), lo = rank * nflat / CS, hi = (rank + 1) * nflat / CS;
        const int ks = ksplit_for(hi - lo, cin);
        conv_dispatch(in, cin, Si, Si * Si, Si, 0, 0, So, Hc, k, a.w.w_fwd[l], scratch, lo, hi, ks);
        __syncthreads();
        reduce_store(cl, CS, scratch, lo, hi, ks, So, Hc, sm + pl.off_a[l], So, 0, EPI_FWD, a.w.scale[l], a.w.shift[l], d.nonlin[l], d.elu_xshift, d.elu_yshift, nullptr,
                     Y0 - pl.halo[l], X0 - pl.halo[l], H, W);
        cl.sync();
    }

    // ---- readout + encoder nonlinearity (every CTA holds the full top window; rank 0 reports)
    const float* top = sm + pl.off_a[L - 1];  // [Hc][2][2]
    float part = 0.f;
    for (int i = tid; i < Hc * 4; i += NT) part += top[i] * s_tap[i & 3] * __ldg(a.w.features + (size_t)neuron * Hc + (i >> 2));
    const float ysum = block_sum(part, s_red) + (a.w.bias ? __ldg(a.w.bias + neuron) : 0.f);
    const float pre = ysum + d.elu_offset;
    if (tid == 0 && rank == 0) a.act[pair] = (pre > 0.f ? pre : expm1f(pre)) + 1.f;
    if (!with_grad) return;

    // ---- backward: d(g_act * act) / d image
    const float gy_ = (a.g_act ? a.g_act[pair] : 1.f) * (pre > 0.f ? 1.f : expf(pre));
    {
        float* gl = sm + pl.off_g[L - 1];
        for (int i = tid; i < Hc * 4; i += NT) {
            const int c = i >> 2, tap = i & 3;
            const int gyy = Y0 + (tap >> 1), gxx = X0 + (tap & 1);
            float v = gy_ * __ldg(a.w.features + (size_t)neuron * Hc + c) * s_tap[tap];
            v *= __ldg(a.w.scale[L - 1] + c) * (d.nonlin[L - 1] ? elu_grad_from_out(top[i], d.elu_yshift) : 1.f);
            gl[(c * 2 + (tap >> 1)) * pl.GS[L - 1] + (tap & 1) + ((pl.GS[L - 1] - 2) >> 1)] = (gyy >= 0 && gyy < H && gxx >= 0 && gxx < W) ? v : 0.f;
        }
    }
    __syncthreads();
    for (int l = L - 1; l >= 0; --l) {
        // gradient wrt the input window of layer l = full correlation of g[l] (pre-batch-norm gradient) with the flipped kernel
        const int cout = l == 0 ? Cin : Hc, So = l == 0 ? Sin : pl.S[l - 1], Si = pl.S[l], k = d.kernel[l];
        const int nflat = flat_items(So, cout), lo = rank * nflat / CS, hi = (rank + 1) * nflat / CS;
        const int ks = ksplit_for(hi - lo, Hc);
        conv_dispatch(sm + pl.off_g[l], Hc, pl.GS[l], Si * pl.GS[l], Si, k - 1, pl.GS[l] == Si ? Si : 0, So, cout, k, a.w.w_bwd[l], scratch, lo, hi, ks);
        __syncthreads();
        if (l > 0)
            reduce_store(cl, CS, scratch, lo, hi, ks, So, cout, sm + pl.off_g[l - 1], pl.GS[l - 1], (pl.GS[l - 1] - So) >> 1, EPI_BWD, a.w.scale[l - 1], nullptr, d.nonlin[l - 1], d.elu_xshift,
                         d.elu_yshift, sm + pl.off_a[l - 1], Y0 - pl.halo[l - 1], X0 - pl.halo[l - 1], H, W);
        else
            reduce_store(cl, CS, scratch, lo, hi, ks, So, cout, sm + pl.off_gin, So, 0, EPI_PLAIN, nullptr, nullptr, 0, 0.f, 0.f, nullptr, 0, 0, H, W);
        cl.sync();
    }
    // ---- write the image gradient (this cluster owns the whole image: no atomics; the CTAs interleave)
    float* gimg = a.g_img + img_off;
    const float* gin = sm + pl.off_gin;
    if (a.accumulate) {  // 1: add the window into the image; 2: overwrite the window only (the caller keeps the rest of the image at zero)
        const bool add = a.accumulate == 1;
        for (int i = rank * NT + tid; i < Cin * Sin * Sin; i += NT * CS) {
            const int c = i / (Sin * Sin), r = i - c * Sin * Sin, y = r / Sin, x = r - y * Sin;
            const int gyy = iy0 + y, gxx = ix0 + x;
            if (gyy >= 0 && gyy < H && gxx >= 0 && gxx < W) {
                float* dst = gimg + (size_t)c * P + gyy * W + gxx;
                *dst = add ? *dst + gin[i] : gin[i];
            }
        }
    } else {
        // one warp per image row (the cluster's warps interleave): rows that miss the window are zero-filled with 16-byte stores, the others
        // test the column only -- no division per pixel (the flat-index version of this loop was 10 % of the kernel's instructions)
        const int wg = rank * (NT / 32) + (tid >> 5), nwg = CS * (NT / 32), lane = tid & 31;
        const bool vec = (W & 3) == 0 && (reinterpret_cast<size_t>(gimg) & 15) == 0;
        for (int rw = wg; rw < Cin * H; rw += nwg) {
            const int c = rw / H, gyy = rw - c * H, y = gyy - iy0;
            float* dst = gimg + (size_t)c * P + (size_t)gyy * W;
            if (y < 0 || y >= Sin) {
                if (vec)
                    for (int x4 = lane; x4 < (W >> 2); x4 += 32) reinterpret_cast<float4*>(dst)[x4] = make_float4(0.f, 0.f, 0.f, 0.f);
                else
                    for (int gxx = lane; gxx < W; gxx += 32) dst[gxx] = 0.f;
            } else {
                const float* src = gin + (c * Sin + y) * Sin - ix0;
                for (int gxx = lane; gxx < W; gxx += 32) dst[gxx] = (gxx >= ix0 && gxx < ix0 + Sin) ? src[gxx] : 0.f;
            }
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
    if ((size_t)pl.total * sizeof(float) > 220 * 1024) cone::make_plan(d, g_img != nullptr, pl, false);  // large cones: plain gradient rows, bounds-checked
    const size_t smem = (size_t)pl.total * sizeof(float);
    LMR_CHECK_ARG(smem <= 220 * 1024, "convcore_response: the receptive-field cone needs %zu bytes of shared memory (> 220 KB)", smem);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    LMR_CUDA_OK(cudaFuncSetAttribute(cone::convcone_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // cluster size: spread each (image, neuron) pair over as many SMs as still leaves every pair resident at once (a cluster lives inside one
    // GPC, so the driver is asked how many clusters of each size fit; cached per shared-memory size)
    const int pairs = G * NB;
    static int cached_smem = -1, max_active[9];
    if (cached_smem != (int)smem) {
        for (int c = 2; c <= 8; ++c) {
            cudaLaunchConfig_t q = {};
            q.gridDim = dim3((unsigned)(c * 64)), q.blockDim = dim3(cone::NT), q.dynamicSmemBytes = smem;
            cudaLaunchAttribute qa[1];
            qa[0].id = cudaLaunchAttributeClusterDimension;
            qa[0].val.clusterDim.x = (unsigned)c, qa[0].val.clusterDim.y = 1, qa[0].val.clusterDim.z = 1;
            q.attrs = qa, q.numAttrs = 1;
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, cone::convcone_kernel, &q) != cudaSuccess) n = 0, (void)cudaGetLastError();
            max_active[c] = n;
        }
        cached_smem = (int)smem;
    }
    int cs = 1;
    for (int c = 8; c >= 2; --c)
        if (max_active[c] >= pairs) {
            cs = c;
            break;
        }
    cone::Args a;
    a.d = d, a.w = *weights, a.img = img, a.group_stride = img_group_stride, a.NB = NB, a.G = G, a.neuron_of_group = neuron_of_group;
    a.g_act = g_act, a.act = act, a.g_img = g_img, a.accumulate = accumulate, a.pl = pl;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(pairs * cs)), cfg.blockDim = dim3(cone::NT), cfg.dynamicSmemBytes = smem, cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cs, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    LMR_CUDA_OK(cudaLaunchKernelEx(&cfg, cone::convcone_kernel, a));
    return 0;
}
