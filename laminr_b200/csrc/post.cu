// Bandwidth-bound stages of the LAMINR hot path: image-constraint projection, simulated response model,
// masked contrastive objective, Adam, batched MEI ascent.  All fp32, coalesced / vectorised, warp-shuffle
// reductions, deterministic two-stage reductions (no float atomics).
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace lmr {
namespace post {

// ===============================================================================================================
// Image-constraint projection (reference laminr/utils/img_transforms.py:19-65; featurevis/ops.py:441-481)
// ===============================================================================================================
constexpr int CCH = 2048;  // elements per CTA chunk (256 threads x 8)

// partial layout per (image b, chunk s): 4 floats
//   fwd NORM: [sum (x-mean)^2]                 fwd STD: [count, chunk mean, chunk M2]
//   bwd NORM: [sum gy*u]                       bwd STD: [sum gy, sum gy*(x-mu)]
__global__ void __launch_bounds__(256) constrain_stats_kernel(int mode, const float* __restrict__ x, int n, float mean, float* __restrict__ part,
                                                              const float* __restrict__ g = nullptr, float step = 0.f) {
    __shared__ float scratch[40];
    const int b = blockIdx.y, s = blockIdx.x, S = gridDim.x;
    const float* xb = x + (size_t)b * n;
    const float* gb = g ? g + (size_t)b * n : nullptr;  // ascent update: statistics of x + step*g
    const int i0 = s * CCH, i1 = min(n, i0 + CCH);
    float v[CCH / 256];
    int cnt = 0;
#pragma unroll
    for (int q = 0; q < CCH / 256; ++q) {
        const int i = i0 + q * 256 + threadIdx.x;
        v[q] = i < i1 ? (gb ? fmaf(step, gb[i], xb[i]) : xb[i]) : 0.f;
        cnt += i < i1;
    }
    float* dst = part + ((size_t)b * S + s) * 4;
    if (mode == LMR_CONSTRAIN_NORM) {
        float acc = 0.f;
#pragma unroll
        for (int q = 0; q < CCH / 256; ++q) {
            const int i = i0 + q * 256 + threadIdx.x;
            const float u = v[q] - mean;
            acc += i < i1 ? u * u : 0.f;
        }
        acc = block_sum(acc, scratch);
        if (threadIdx.x == 0) dst[0] = acc;
    } else {
        float sum = 0.f;
#pragma unroll
        for (int q = 0; q < CCH / 256; ++q) sum += v[q];
        sum = block_sum(sum, scratch);
        const float m = sum / (float)(i1 - i0);
        float m2 = 0.f;
#pragma unroll
        for (int q = 0; q < CCH / 256; ++q) {
            const int i = i0 + q * 256 + threadIdx.x;
            const float u = v[q] - m;
            m2 += i < i1 ? u * u : 0.f;
        }
        m2 = block_sum(m2, scratch);
        if (threadIdx.x == 0) dst[0] = (float)(i1 - i0), dst[1] = m, dst[2] = m2;
    }
    (void)cnt;
}

// combine the S chunk partials of image b -> (s or std, mu).  Chan et al. pairwise update for the variance.
__device__ inline void combine_stats(int mode, const float* part, int S, int n, float* out_s, float* out_mu) {
    if (mode == LMR_CONSTRAIN_NORM) {
        float acc = 0.f;
        for (int s = 0; s < S; ++s) acc += part[s * 4];
        *out_s = sqrtf(acc);
        *out_mu = 0.f;
    } else {
        float cnt = 0.f, m = 0.f, m2 = 0.f;
        for (int s = 0; s < S; ++s) {
            const float cb = part[s * 4], mb = part[s * 4 + 1], m2b = part[s * 4 + 2];
            const float tot = cnt + cb, delta = mb - m;
            m2 += m2b + delta * delta * cnt * cb / tot;
            m += delta * cb / tot;
            cnt = tot;
        }
        *out_s = sqrtf(m2 / (float)(n - 1));  // unbiased (Tensor.std default)
        *out_mu = m;
    }
}

__global__ void __launch_bounds__(256) constrain_apply_kernel(int mode, const float* x, int n, float target, float mean, int has_min,
                                                              float pmin, int has_max, float pmax, float eps, const float* __restrict__ part,
                                                              float* z, float* __restrict__ stats, const float* __restrict__ g = nullptr,
                                                              float step = 0.f) {
    __shared__ float sh[2];
    const int b = blockIdx.y, s = blockIdx.x, S = gridDim.x;
    if (threadIdx.x == 0) {
        combine_stats(mode, part + (size_t)b * S * 4, S, n, &sh[0], &sh[1]);
        if (s == 0 && stats) stats[2 * b] = sh[0], stats[2 * b + 1] = sh[1];
    }
    __syncthreads();
    const float sd = sh[0], mu = sh[1];
    const float* xb = x + (size_t)b * n;
    const float* gb = g ? g + (size_t)b * n : nullptr;
    float* zb = z + (size_t)b * n;  // may alias x (each element is read and written by the same thread)
    const int i0 = s * CCH, i1 = min(n, i0 + CCH);
    for (int i = i0 + threadIdx.x; i < i1; i += 256) {
        const float xv = gb ? fmaf(step, gb[i], xb[i]) : xb[i];
        float y;
        if (mode == LMR_CONSTRAIN_NORM) y = (xv - mean) / (sd + eps) * target + mean;
        else y = target * (xv - mu) / (sd + eps) + mean;
        if (has_min) y = fmaxf(y, pmin);
        if (has_max) y = fminf(y, pmax);
        zb[i] = y;
    }
}

// The ascent update x <- clip(renorm(x + step*g)) of ONE image per CTA, the image held in shared memory between the statistics and the apply pass
// (one launch and one read of x and g instead of two of each; the conv-encoder MEI ascent runs it 1000 times behind the cone kernel).  Four groups
// of 256 threads walk the same 2048-element chunks as constrain_stats_kernel, with the same per-thread sums and the same reduction tree per chunk,
// and one thread combines the chunk partials with combine_stats: bit-identical to the two-launch path.
constexpr int AUF_NT = 1024, AUF_G = AUF_NT / 256;
__device__ __forceinline__ float group_sum256(float v, float* scratch /* [AUF_G][40] */, int grp, int lt) {
    const int lane = lt & 31, warp = lt >> 5;
    float* sc = scratch + grp * 40;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sc[warp] = v;
    __syncthreads();
    if (warp == 0) {
        float t = lane < 8 ? sc[lane] : 0.f;
        t = warp_sum(t);
        if (lane == 0) sc[32] = t;
    }
    __syncthreads();
    return sc[32];
}
__global__ void __launch_bounds__(AUF_NT, 1) ascent_update_fused_kernel(int mode, float* x, const float* __restrict__ g, float step, int n, float target,
                                                                        float mean, int has_min, float pmin, int has_max, float pmax, float eps) {
    extern __shared__ __align__(16) float aus[];
    __shared__ float scratch[AUF_G * 40];
    __shared__ float sh[2];
    const int S = (n + CCH - 1) / CCH;
    float* v = aus;                        // [n]  x + step*g
    float* part = aus + ((n + 3) & ~3);    // [S][4]
    const int b = blockIdx.x, t = threadIdx.x, grp = t >> 8, lt = t & 255;
    float* xb = x + (size_t)b * n;
    const float* gb = g + (size_t)b * n;
    for (int it = 0; it * AUF_G < S; ++it) {
        const int s = it * AUF_G + grp;
        const int i0 = s * CCH, i1 = min(n, i0 + CCH);  // (s >= S: empty chunk, the group only keeps the barriers company)
        float w[CCH / 256];
#pragma unroll
        for (int q = 0; q < CCH / 256; ++q) {
            const int i = i0 + q * 256 + lt;
            w[q] = i < i1 ? fmaf(step, gb[i], xb[i]) : 0.f;
            if (i < i1) v[i] = w[q];
        }
        if (mode == LMR_CONSTRAIN_NORM) {
            float acc = 0.f;
#pragma unroll
            for (int q = 0; q < CCH / 256; ++q) {
                const int i = i0 + q * 256 + lt;
                const float u = w[q] - mean;
                acc += i < i1 ? u * u : 0.f;
            }
            acc = group_sum256(acc, scratch, grp, lt);
            if (lt == 0 && s < S) part[s * 4] = acc;
        } else {
            float sum = 0.f;
#pragma unroll
            for (int q = 0; q < CCH / 256; ++q) sum += w[q];
            sum = group_sum256(sum, scratch, grp, lt);
            const float m = sum / (float)(i1 - i0);
            float m2 = 0.f;
#pragma unroll
            for (int q = 0; q < CCH / 256; ++q) {
                const int i = i0 + q * 256 + lt;
                const float u = w[q] - m;
                m2 += i < i1 ? u * u : 0.f;
            }
            m2 = group_sum256(m2, scratch, grp, lt);
            if (lt == 0 && s < S) part[s * 4] = (float)(i1 - i0), part[s * 4 + 1] = m, part[s * 4 + 2] = m2;
        }
    }
    __syncthreads();
    if (t == 0) combine_stats(mode, part, S, n, &sh[0], &sh[1]);
    __syncthreads();
    const float sd = sh[0], mu = sh[1];
    for (int i = t; i < n; i += AUF_NT) {
        const float xv = v[i];
        float y;
        if (mode == LMR_CONSTRAIN_NORM) y = (xv - mean) / (sd + eps) * target + mean;
        else y = target * (xv - mu) / (sd + eps) + mean;
        if (has_min) y = fmaxf(y, pmin);
        if (has_max) y = fminf(y, pmax);
        xb[i] = y;
    }
}

__device__ __forceinline__ float clamp_pass(float y, int has_min, float pmin, int has_max, float pmax) {
    // torch.clamp backward passes the gradient where pmin <= y <= pmax (inclusive)
    return ((!has_min || y >= pmin) && (!has_max || y <= pmax)) ? 1.f : 0.f;
}

__global__ void __launch_bounds__(256) constrain_bwd_stats_kernel(int mode, const float* __restrict__ x, const float* __restrict__ gz,
                                                                  const float* __restrict__ stats, int n, float target, float mean, int has_min,
                                                                  float pmin, int has_max, float pmax, float eps, float* __restrict__ part) {
    __shared__ float scratch[40];
    const int b = blockIdx.y, s = blockIdx.x, S = gridDim.x;
    const float sd = stats[2 * b], mu = stats[2 * b + 1];
    const float* xb = x + (size_t)b * n;
    const float* gb = gz + (size_t)b * n;
    const int i0 = s * CCH, i1 = min(n, i0 + CCH);
    float a0 = 0.f, a1 = 0.f;
    for (int i = i0 + threadIdx.x; i < i1; i += 256) {
        float u, y;
        if (mode == LMR_CONSTRAIN_NORM) u = xb[i] - mean, y = u / (sd + eps) * target + mean;
        else u = xb[i] - mu, y = target * u / (sd + eps) + mean;
        const float gy = gb[i] * clamp_pass(y, has_min, pmin, has_max, pmax);
        a0 += gy;
        a1 = fmaf(gy, u, a1);
    }
    a0 = block_sum(a0, scratch);
    a1 = block_sum(a1, scratch);
    if (threadIdx.x == 0) part[((size_t)b * S + s) * 4] = a0, part[((size_t)b * S + s) * 4 + 1] = a1;
}

__global__ void __launch_bounds__(256) constrain_bwd_apply_kernel(int mode, const float* __restrict__ x, const float* __restrict__ gz,
                                                                  const float* __restrict__ stats, int n, float target, float mean, int has_min,
                                                                  float pmin, int has_max, float pmax, float eps, const float* __restrict__ part,
                                                                  float* __restrict__ gx) {
    __shared__ float sh[2];
    const int b = blockIdx.y, s = blockIdx.x, S = gridDim.x;
    if (threadIdx.x == 0) {
        float a0 = 0.f, a1 = 0.f;
        for (int q = 0; q < S; ++q) a0 += part[((size_t)b * S + q) * 4], a1 += part[((size_t)b * S + q) * 4 + 1];
        sh[0] = a0, sh[1] = a1;
    }
    __syncthreads();
    const float sd = stats[2 * b], mu = stats[2 * b + 1];
    const float sum_gy = sh[0], dot = sh[1];
    const float k = target / (sd + eps);
    // NORM: gx = k gy - target dot / ((s+eps)^2 s) u ;  STD: gx = k (gy - mean(gy)) - target dot / ((sd+eps)^2 (n-1) sd) (x-mu)
    const float denom = (sd + eps) * (sd + eps) * sd * (mode == LMR_CONSTRAIN_NORM ? 1.f : (float)(n - 1));
    const float c2 = target * dot / denom;
    const float gmean = mode == LMR_CONSTRAIN_NORM ? 0.f : sum_gy / (float)n;
    const float* xb = x + (size_t)b * n;
    const float* gb = gz + (size_t)b * n;
    float* ob = gx + (size_t)b * n;
    const int i0 = s * CCH, i1 = min(n, i0 + CCH);
    for (int i = i0 + threadIdx.x; i < i1; i += 256) {
        float u, y;
        if (mode == LMR_CONSTRAIN_NORM) u = xb[i] - mean, y = u / (sd + eps) * target + mean;
        else u = xb[i] - mu, y = target * u / (sd + eps) + mean;
        const float gy = gb[i] * clamp_pass(y, has_min, pmin, has_max, pmax);
        ob[i] = k * (gy - gmean) - c2 * u;
    }
}

// ===============================================================================================================
// Simulated response model (reference laminr/neuron_models/utils/simulated_model.py:8-29)
// ===============================================================================================================
constexpr int RCH = 512;  // pixels per CTA chunk
constexpr int RIB = 8;    // images per register block

// Rpart[g][s][b][f] = sum_{p in chunk s} (sum_c img[g,b,c,p]) * w[f][p]
__global__ void __launch_bounds__(256) simresp_partial_kernel(const float* __restrict__ img, long long gstride, int NB, int C, int HW,
                                                              const float* __restrict__ filters, int Fmax, const int* __restrict__ nfilt,
                                                              const int* __restrict__ neuron_of_group, float* __restrict__ Rpart, int rch /* pixels per CTA, <= RCH */) {
    __shared__ float xs[RIB][RCH];
    const int s = blockIdx.x, S = gridDim.x, g = blockIdx.y;
    const int neuron = neuron_of_group[g];
    const int F = nfilt[neuron];
    const float* bank = filters + (size_t)neuron * Fmax * HW;
    const float* gimg = img + (size_t)g * gstride;
    const int p0 = s * rch, np = min(rch, HW - p0);
    const int rsh = 31 - __clz(rch);  // rch is a power of two
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* out = Rpart + ((size_t)g * S + s) * NB * Fmax;
    for (int b0 = 0; b0 < NB; b0 += RIB) {
        const int nb = min(RIB, NB - b0);
        __syncthreads();
        for (int i = threadIdx.x; i < RIB * rch; i += 256) {
            const int bb = i >> rsh, p = i & (rch - 1);
            float v = 0.f;
            if (bb < nb && p < np) {
                const float* src = gimg + ((size_t)(b0 + bb) * C) * HW + p0 + p;
                for (int c = 0; c < C; ++c) v += src[(size_t)c * HW];
            }
            xs[bb][p] = v;
        }
        __syncthreads();
        for (int f = warp; f < F; f += 8) {
            const float* w = bank + (size_t)f * HW + p0;
            float acc[RIB];
#pragma unroll
            for (int bb = 0; bb < RIB; ++bb) acc[bb] = 0.f;
            for (int p = lane; p < np; p += 32) {
                const float wv = __ldg(w + p);
#pragma unroll
                for (int bb = 0; bb < RIB; ++bb) acc[bb] = fmaf(xs[bb][p], wv, acc[bb]);
            }
#pragma unroll
            for (int bb = 0; bb < RIB; ++bb) acc[bb] = warp_sum(acc[bb]);
            if (lane == 0)
                for (int bb = 0; bb < nb; ++bb) out[(size_t)(b0 + bb) * Fmax + f] = acc[bb];
        }
    }
}

// one warp per (group, image): lanes own the filters f = lane, lane + 32, ..; each sums its filter's chunk partials in chunk order (deterministic),
// then the warp takes the maximum with the lowest filter index on ties (torch.max returns the first maximum)
__global__ void __launch_bounds__(128) simresp_finalize_kernel(const float* __restrict__ Rpart, int S, int NB, int G, int Fmax, const int* __restrict__ nfilt,
                                                               const int* __restrict__ neuron_of_group, float eps, float* __restrict__ act,
                                                               float* __restrict__ rmax, int* __restrict__ argmax) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= G * NB) return;  // whole warps leave together
    const int g = i / NB, b = i % NB;
    const int F = nfilt[neuron_of_group[g]];
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int f = lane; f < F; f += 32) {
        const float* src = Rpart + (((size_t)g * S) * NB + b) * Fmax + f;
        float r4[4] = {0.f, 0.f, 0.f, 0.f};  // four independent chains (the loads are L2 round trips), combined in a fixed order
        int sidx = 0;
        for (; sidx + 16 <= S; sidx += 16) {  // 16 round trips in flight, added in the order of the 4-wide loop below (same bits); few groups
            float v[16];                      // mean many chunks: 79 per response for two targets at 100x100, a 20 us kernel with 4 in flight
#pragma unroll
            for (int u = 0; u < 16; ++u) v[u] = src[(size_t)(sidx + u) * NB * Fmax];
#pragma unroll
            for (int u = 0; u < 16; ++u) r4[u & 3] += v[u];
        }
        for (; sidx + 4 <= S; sidx += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) r4[u] += src[(size_t)(sidx + u) * NB * Fmax];
        }
        for (int u = 0; sidx < S; ++sidx, ++u) r4[u] += src[(size_t)sidx * NB * Fmax];
        const float r = (r4[0] + r4[1]) + (r4[2] + r4[3]);
        if (r > best) best = r, bi = f;  // strict: the lane's first maximum
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) best = ob, bi = oi;
    }
    if (lane == 0) {
        rmax[i] = best;
        argmax[i] = bi;
        const float e = best > 0.f ? best : expm1f(best);
        act[i] = (e + 1.f) / 2.f + eps;
    }
}

// g_img[g,b,c,p] (+)= g_act * 1/2 * elu'(m) * w_{f*}[p]
__global__ void __launch_bounds__(256) simresp_bwd_kernel(const float* __restrict__ g_act, const float* __restrict__ rmax, const int* __restrict__ argmax,
                                                          long long gstride, int NB, int G, int C, int HW, const float* __restrict__ filters, int Fmax,
                                                          const int* __restrict__ neuron_of_group, float* __restrict__ g_img, int accumulate) {
    const int b = blockIdx.y;
    const int gout = blockIdx.z;  // == group when images are per-group; single slice when shared
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    float v = 0.f;
    const int g_lo = gstride ? gout : 0, g_hi = gstride ? gout + 1 : G;
    for (int g = g_lo; g < g_hi; ++g) {
        const int i = g * NB + b;
        const float m = rmax[i];
        const float coef = g_act[i] * 0.5f * (m > 0.f ? 1.f : expf(m));
        v = fmaf(coef, __ldg(filters + ((size_t)neuron_of_group[g] * Fmax + argmax[i]) * HW + p), v);
    }
    float* dst = g_img + (size_t)gout * gstride + ((size_t)b * C) * HW + p;
    for (int c = 0; c < C; ++c) {
        if (accumulate) dst[(size_t)c * HW] += v;
        else dst[(size_t)c * HW] = v;
    }
}

// ===============================================================================================================
// Masked contrastive objective (reference laminr/utils/mask.py:7-22, laminr/contrastive_loss.py:139-154)
// ===============================================================================================================
constexpr int GCH = 256;   // elements per CTA chunk
constexpr int GMAXN = 64;  // max grid points

__device__ __forceinline__ float masked_value(float x, float m, float c0, float g1, float g2) {
    // rescale(x, pmin, pmax, -1, 1) * mask, then rescale back (same op order as mask.py:7-22)
    const float r = (x - c0) * g1 + 0.f;
    return (r * m - 0.f) * g2 + c0;
}

// Gpart[s][i][j] = sum_{e in chunk s} xm_i[e] xm_j[e]
__global__ void __launch_bounds__(256) simclr_gram_kernel(const float* __restrict__ img, const float* __restrict__ mask, int N, int C, int HW,
                                                          float c0, float g1, float g2, float* __restrict__ Gpart) {
    extern __shared__ float xm[];  // [N][GCH+1]
    const int n = C * HW;
    const int e0 = blockIdx.x * GCH, ne = min(GCH, n - e0);
    for (int i = threadIdx.x; i < N * GCH; i += 256) {
        const int r = i / GCH, e = i % GCH;
        float v = 0.f;
        if (e < ne) v = masked_value(img[(size_t)r * n + e0 + e], mask[(e0 + e) % HW], c0, g1, g2);
        xm[r * (GCH + 1) + e] = v;
    }
    __syncthreads();
    float* out = Gpart + (size_t)blockIdx.x * N * N;
    for (int pr = threadIdx.x; pr < N * N; pr += 256) {
        const int i = pr / N, j = pr % N;
        if (j < i) continue;
        const float* a = xm + i * (GCH + 1);
        const float* b = xm + j * (GCH + 1);
        float acc = 0.f;
        for (int e = 0; e < GCH; ++e) acc = fmaf(a[e], b[e], acc);
        out[i * N + j] = acc;
        out[j * N + i] = acc;
    }
}

// one CTA: reduce the gram, evaluate reg and the backward coefficient matrix K
__global__ void __launch_bounds__(1024) simclr_finalize_kernel(const float* __restrict__ Gpart, int S, int N, const float* __restrict__ pos,
                                                              const float* __restrict__ neg, float T, float eps, float* __restrict__ reg,
                                                              float* __restrict__ K) {
    extern __shared__ float sm[];
    float* cosm = sm;             // [N][N] cos, later Lambda
    float* Sm = cosm + N * N;     // [N][N] exp(cos/T), later Gamma
    float* rho = Sm + N * N;      // [N]
    float* rowc = rho + N;        // [N][4]: a_pos, a_neg (row coefficients), log term, kappa
    float* posS = rowc + 4 * N;   // [N][N] shared copies of the neighbour masks (the row loops below are serial chains: keep their loads on chip)
    float* negS = posS + N * N;
    __shared__ float scratch[40];
    // gram entries: the S chunk partials of an entry are split over G thread groups (coalesced across consecutive entries), each thread keeps eight
    // independent chains in flight (the loads are L2 round trips); fixed-order combines -> deterministic
    {
        const int NN = N * N;
        int G = (int)blockDim.x / NN;
        G = G < 1 ? 1 : (G > 4 ? 4 : G);
        auto part = [&](int g) { return g == 0 ? cosm : (g == 1 ? Sm : (g == 2 ? posS : negS)); };  // scratch until the masks are staged below
        for (int item = threadIdx.x; item < G * NN; item += blockDim.x) {
            const int g = item / NN, pr = item - g * NN;
            float a8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            int s = g;
            for (; s + 7 * G < S; s += 8 * G) {
#pragma unroll
                for (int u = 0; u < 8; ++u) a8[u] += Gpart[(size_t)(s + u * G) * NN + pr];
            }
            for (int u = 0; s < S; s += G, ++u) a8[u] += Gpart[(size_t)s * NN + pr];
            part(g)[pr] = ((a8[0] + a8[1]) + (a8[2] + a8[3])) + ((a8[4] + a8[5]) + (a8[6] + a8[7]));
        }
        __syncthreads();
        for (int pr = threadIdx.x; pr < NN; pr += blockDim.x) {
            float tot = cosm[pr];
            for (int g = 1; g < G; ++g) tot += part(g)[pr];
            cosm[pr] = tot;
        }
        __syncthreads();
        for (int pr = threadIdx.x; pr < NN; pr += blockDim.x) posS[pr] = pos[pr], negS[pr] = neg[pr];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) rho[i] = sqrtf(cosm[i * N + i]);
    __syncthreads();
    for (int pr = threadIdx.x; pr < N * N; pr += blockDim.x) {
        const int i = pr / N, j = pr % N;
        const float c = cosm[pr] / (rho[i] * rho[j]);
        cosm[pr] = c;
        Sm[pr] = expf(c / T);
    }
    __syncthreads();
    float logsum = 0.f;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        float ps = 0.f, np_ = 0.f, ns = 0.f, nn = 0.f;
        for (int j = 0; j < N; ++j) {
            ps = fmaf(Sm[i * N + j], posS[i * N + j], ps), np_ += posS[i * N + j];
            ns = fmaf(Sm[i * N + j], negS[i * N + j], ns), nn += negS[i * N + j];
        }
        const float num = (ps == 0.f ? 0.f : ps / np_) + eps;
        const float den = (ns == 0.f ? 0.f : ns / nn) + eps;
        logsum += logf(num / den);
        rowc[i * 4 + 0] = ps == 0.f ? 0.f : (T / (2.f * N)) / (np_ * num);
        rowc[i * 4 + 1] = ns == 0.f ? 0.f : (T / (2.f * N)) / (nn * den);
    }
    logsum = block_sum(logsum, scratch);
    if (threadIdx.x == 0) reg[0] = logsum / (float)N * T / 2.f;
    __syncthreads();
    // Gamma_ij = A_ij S_ij / T
    for (int pr = threadIdx.x; pr < N * N; pr += blockDim.x) {
        const int i = pr / N;
        Sm[pr] = (posS[pr] * rowc[i * 4] - negS[pr] * rowc[i * 4 + 1]) * Sm[pr] / T;
    }
    __syncthreads();
    // kappa_i = sum_j Lambda_ij cos_ij, Lambda = Gamma + Gamma^T
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        float kap = 0.f;
        for (int j = 0; j < N; ++j) kap = fmaf(Sm[i * N + j] + Sm[j * N + i], cosm[i * N + j], kap);
        rowc[i * 4 + 3] = kap;
    }
    __syncthreads();
    for (int pr = threadIdx.x; pr < N * N; pr += blockDim.x) {
        const int i = pr / N, j = pr % N;
        float k = (Sm[i * N + j] + Sm[j * N + i]) / (rho[i] * rho[j]);
        if (i == j) k -= rowc[i * 4 + 3] / (rho[i] * rho[i]);
        K[pr] = k;
    }
}

// g_img[i][e] (+)= g_reg * scale * (g1 * mask * g2) * sum_j K_ij xm_j[e]
__global__ void __launch_bounds__(128) simclr_bwd_kernel(const float* __restrict__ img, const float* __restrict__ mask, const float* __restrict__ K,
                                                         const float* __restrict__ g_reg, float scale, int N, int C, int HW, float c0, float g1,
                                                         float g2, float* __restrict__ g_img, int accumulate) {
    extern __shared__ float Ks[];  // [N][N]
    for (int i = threadIdx.x; i < N * N; i += blockDim.x) Ks[i] = K[i];
    __syncthreads();
    const int n = C * HW;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const float m = mask[e % HW];
    float xm[GMAXN];
    for (int j = 0; j < N; ++j) xm[j] = masked_value(img[(size_t)j * n + e], m, c0, g1, g2);
    const float up = g_reg[0] * scale * (g1 * m * g2);
    for (int i = 0; i < N; ++i) {
        float acc = 0.f;
        for (int j = 0; j < N; ++j) acc = fmaf(Ks[i * N + j], xm[j], acc);
        const float v = up * acc;
        if (accumulate) g_img[(size_t)i * n + e] += v;
        else g_img[(size_t)i * n + e] = v;
    }
}

// ===============================================================================================================
// Adam (torch.optim.Adam single-tensor update order; reference inv_temp_learn_match.py:162,335)
// ===============================================================================================================
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int n, float lr,
                            float b1, float b2, float eps, const int* __restrict__ step) {
    const int t = step[0] + 1;
    float step_size, bc2_sqrt;
    adam_step_consts(t, lr, b1, b2, &step_size, &bc2_sqrt);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) adam_element(g[i], m[i], v[i], p[i], b1, b2, eps, step_size, bc2_sqrt, &m[i], &v[i], &p[i]);
}
__global__ void counter_inc_kernel(int* c) { c[0] += 1; }

// one-launch variant: every CTA reads the step counter at its start; the LAST CTA to finish (ticket in step[1]) publishes the
// incremented counter and re-arms the ticket.  step: int32[2] = {steps taken, ticket (0 between launches)}.
__global__ void __launch_bounds__(256) adam_ticket_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                          int n, float lr, float b1, float b2, float eps, int* __restrict__ step) {
    const int t = *reinterpret_cast<volatile int*>(step) + 1;
    float step_size, bc2_sqrt;
    adam_step_consts(t, lr, b1, b2, &step_size, &bc2_sqrt);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) adam_element(g[i], m[i], v[i], p[i], b1, b2, eps, step_size, bc2_sqrt, &m[i], &v[i], &p[i]);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(step + 1, 1) == (int)gridDim.x - 1) {
            step[1] = 0;
            step[0] = t;
        }
    }
}

// several tensors in one launch: block b belongs to the group g with first[g] <= b < first[g + 1]; per-group step word and ticket
struct AdamGroups {
    lmr_adam_group g[LMR_ADAM_MAX_GROUPS];
    int first[LMR_ADAM_MAX_GROUPS + 1];
    int count;
};
__global__ void __launch_bounds__(256) adam_multi_kernel(AdamGroups a, float lr, float b1, float b2, float eps) {
    int gi = 0;
#pragma unroll
    for (int k = 1; k < LMR_ADAM_MAX_GROUPS; ++k)
        if (k < a.count && (int)blockIdx.x >= a.first[k]) gi = k;
    const lmr_adam_group& G = a.g[gi];
    const int t = *reinterpret_cast<volatile int*>(G.step) + 1;
    float step_size, bc2_sqrt;
    adam_step_consts(t, lr, b1, b2, &step_size, &bc2_sqrt);
    const int i = ((int)blockIdx.x - a.first[gi]) * blockDim.x + threadIdx.x;
    if (i < G.n) adam_element(G.g[i], G.exp_avg[i], G.exp_avg_sq[i], G.p[i], b1, b2, eps, step_size, bc2_sqrt, &G.exp_avg[i], &G.exp_avg_sq[i], &G.p[i]);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(G.step + 1, 1) == a.first[gi + 1] - a.first[gi] - 1) {
            G.step[1] = 0;
            G.step[0] = t;
        }
    }
}

// ===============================================================================================================
// match epoch statistics (laminr/inv_temp_learn_match.py:361-362,368-371): acts_d = mean_n act[d,n] / a*_d -> (sum_d, min_d, max_d).
// One CTA; fixed summation order (deterministic).  Lets the stop rule read 12 bytes instead of driving five reductions from the host.
// ===============================================================================================================
__global__ void __launch_bounds__(256) match_stats_kernel(const float* __restrict__ act, const float* __restrict__ mei_act, int D, int N,
                                                          float* __restrict__ out) {
    __shared__ float ssum[256], smin[256], smax[256];
    float s = 0.f, lo = INFINITY, hi = -INFINITY;
    const float inv = 1.f / (float)N;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float a = 0.f;
        for (int n = 0; n < N; ++n) a += act[(size_t)d * N + n];
        a = a * inv / mei_act[d];
        s += a, lo = fminf(lo, a), hi = fmaxf(hi, a);
    }
    ssum[threadIdx.x] = s, smin[threadIdx.x] = lo, smax[threadIdx.x] = hi;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) {
            ssum[threadIdx.x] += ssum[threadIdx.x + w];
            smin[threadIdx.x] = fminf(smin[threadIdx.x], smin[threadIdx.x + w]);
            smax[threadIdx.x] = fmaxf(smax[threadIdx.x], smax[threadIdx.x + w]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = ssum[0], out[1] = smin[0], out[2] = smax[0];
}

// ===============================================================================================================
// Batched MEI gradient ascent (reference featurevis/core.py:54-136 with SGD, laminr/utils/mei.py:33-59)
// One CTA per neuron; the image lives in shared memory for all iterations; filters stream from L2/HBM.
// ===============================================================================================================
constexpr int MEI_THREADS = 512;

__device__ inline void mei_post(float* xs, int n, int mode, float target, float mean, float pmin, float pmax, float* scratch) {
    // ChangeNorm / ChangeStats (eps 1e-9) then ClipRange
    if (mode == LMR_CONSTRAIN_NORM) {
        float acc = 0.f;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const float u = xs[i] - mean;
            acc = fmaf(u, u, acc);
        }
        acc = block_sum(acc, scratch);
        const float denom = sqrtf(acc) + 1e-9f;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const float y = (xs[i] - mean) / denom * target + mean;  // same op order as ops.py:479
            xs[i] = fminf(fmaxf(y, pmin), pmax);
        }
    } else {
        float sum = 0.f;
        for (int i = threadIdx.x; i < n; i += blockDim.x) sum += xs[i];
        sum = block_sum(sum, scratch);
        const float mu = sum / (float)n;
        float m2 = 0.f;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const float u = xs[i] - mu;
            m2 = fmaf(u, u, m2);
        }
        m2 = block_sum(m2, scratch);
        const float sd = sqrtf(m2 / (float)(n - 1));
        const float k = target / (sd + 1e-9f);
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const float y = (xs[i] - mu) * k + mean;
            xs[i] = fminf(fmaxf(y, pmin), pmax);
        }
    }
    __syncthreads();
}

// r[f] = sum_p (sum_c x[c][p]) w_f[p] -> rs[f] (smem).  All threads stream every filter together (coalesced 128-bit loads), FB filters at a
// time so that each image quad read from shared memory serves FB filters and FB x 2 independent global loads are in flight per thread
// (the bank streams from L2/HBM once per iteration: this loop is the bandwidth-bound part of the ascent).  Per-warp partial sums are
// combined in a fixed order.  `part`: [Fmax][MEI_THREADS/32] floats of shared memory.
constexpr int MEI_FB = 4;
__device__ inline void mei_dots(const float* xs, int C, int HW, const float* __restrict__ bank, int F, float* rs, float* part) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = MEI_THREADS / 32;
    if ((HW % 4) == 0) {
        const int HW4 = HW / 4;
        const float4* xs4 = reinterpret_cast<const float4*>(xs);
        for (int f0 = 0; f0 < F; f0 += MEI_FB) {
            float acc[MEI_FB];
            const float4* w4[MEI_FB];
#pragma unroll
            for (int b = 0; b < MEI_FB; ++b) {
                acc[b] = 0.f;
                w4[b] = reinterpret_cast<const float4*>(bank + (size_t)min(f0 + b, F - 1) * HW);  // ragged last block: recompute the last filter
            }
#pragma unroll 2
            for (int q = threadIdx.x; q < HW4; q += MEI_THREADS) {
                float4 ww[MEI_FB];
#pragma unroll
                for (int b = 0; b < MEI_FB; ++b) ww[b] = __ldg(w4[b] + q);
                float4 xv = xs4[q];
                for (int c = 1; c < C; ++c) {
                    const float4 t = xs4[(size_t)c * HW4 + q];
                    xv.x += t.x, xv.y += t.y, xv.z += t.z, xv.w += t.w;
                }
#pragma unroll
                for (int b = 0; b < MEI_FB; ++b)
                    acc[b] = fmaf(xv.x, ww[b].x, fmaf(xv.y, ww[b].y, fmaf(xv.z, ww[b].z, fmaf(xv.w, ww[b].w, acc[b]))));
            }
#pragma unroll
            for (int b = 0; b < MEI_FB; ++b) {
                const float v = warp_sum(acc[b]);
                if (lane == 0 && f0 + b < F) part[(f0 + b) * NW + warp] = v;
            }
        }
    } else {
        for (int f = 0; f < F; ++f) {
            const float* w = bank + (size_t)f * HW;
            float acc = 0.f;
            for (int p = threadIdx.x; p < HW; p += MEI_THREADS) {
                float xv = xs[p];
                for (int c = 1; c < C; ++c) xv += xs[(size_t)c * HW + p];
                acc = fmaf(xv, __ldg(w + p), acc);
            }
            acc = warp_sum(acc);
            if (lane == 0) part[f * NW + warp] = acc;
        }
    }
    __syncthreads();
    for (int f = threadIdx.x; f < F; f += MEI_THREADS) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) v += part[f * NW + w];
        rs[f] = v;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(MEI_THREADS) mei_kernel(float* __restrict__ x, int C, int HW, const float* __restrict__ filters, int Fmax,
                                                          const int* __restrict__ nfilt, const int* __restrict__ neuron_of_group, float step_size,
                                                          int iters, int mode, float target, float mean, float pmin, float pmax, float act_eps,
                                                          int apply_post_first, float* __restrict__ fevals) {
    extern __shared__ __align__(16) float xs[];  // [C*HW] + r[Fmax]
    __shared__ float scratch[40];
    __shared__ float sel[2];
    const int g = blockIdx.x, n = C * HW;
    const int neuron = neuron_of_group[g];
    const int F = nfilt[neuron];
    const float* bank = filters + (size_t)neuron * Fmax * HW;
    float* rs = xs + ((n + 3) / 4 * 4);
    float* part = rs + ((Fmax + 3) / 4 * 4);  // [Fmax][MEI_THREADS/32]
    float* xg = x + (size_t)g * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) xs[i] = xg[i];
    __syncthreads();
    if (apply_post_first) mei_post(xs, n, mode, target, mean, pmin, pmax, scratch);
    for (int it = 0; it <= iters; ++it) {
        mei_dots(xs, C, HW, bank, F, rs, part);
        if (threadIdx.x == 0) {
            float best = -INFINITY;
            int bi = 0;
            for (int f = 0; f < F; ++f)
                if (rs[f] > best) best = rs[f], bi = f;
            const float e = best > 0.f ? best : expm1f(best);
            fevals[(size_t)g * (iters + 1) + it] = (e + 1.f) / 2.f + act_eps;
            sel[0] = 0.5f * (best > 0.f ? 1.f : expf(best));
            sel[1] = __int_as_float(bi);
        }
        __syncthreads();
        if (it == iters) break;
        const float coef = sel[0];
        const float* w = bank + (size_t)__float_as_int(sel[1]) * HW;
        // SGD on -f: x <- x - lr * (-(coef * w)) for every channel
        for (int i = threadIdx.x; i < n; i += blockDim.x) xs[i] = xs[i] + step_size * (coef * __ldg(w + i % HW));
        __syncthreads();
        mei_post(xs, n, mode, target, mean, pmin, pmax, scratch);
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) xg[i] = xs[i];
}

}  // namespace post
}  // namespace lmr

// =================================================================================================================
// C ABI
// =================================================================================================================
using namespace lmr;
using namespace lmr::post;

extern "C" size_t lmr_constrain_workspace_bytes(int B, int n) {
    const int S = (n + CCH - 1) / CCH;
    return (size_t)B * S * 4 * sizeof(float);
}

extern "C" int lmr_constrain_fwd(int mode, const float* x, int B, int n, float target, float mean, int has_min, float pmin, int has_max,
                                 float pmax, float eps, float* z, float* stats, void* ws, size_t ws_bytes, void* stream) {
    LMR_CHECK_ARG(mode == LMR_CONSTRAIN_NORM || mode == LMR_CONSTRAIN_STD, "constrain: bad mode %d", mode);
    LMR_CHECK_ARG(B >= 1 && n >= 2, "constrain: bad sizes B=%d n=%d", B, n);
    if (ws_bytes < lmr_constrain_workspace_bytes(B, n)) {
        set_error("constrain_fwd: workspace too small");
        return LMR_ERR_WORKSPACE;
    }
    const int S = (n + CCH - 1) / CCH;
    LMR_CHECK_ARG(B <= 65535, "constrain: B=%d > 65535", B);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 grid(S, B);
    constrain_stats_kernel<<<grid, 256, 0, st>>>(mode, x, n, mean, static_cast<float*>(ws));
    LMR_LAUNCH_OK();
    constrain_apply_kernel<<<grid, 256, 0, st>>>(mode, x, n, target, mean, has_min, pmin, has_max, pmax, eps, static_cast<const float*>(ws), z, stats);
    LMR_LAUNCH_OK();
    return 0;
}

extern "C" int lmr_ascent_update(int mode, float* x, const float* g, float step_size, int B, int n, float target, float mean, int has_min,
                                 float pmin, int has_max, float pmax, float eps, void* ws, size_t ws_bytes, void* stream) {
    LMR_CHECK_ARG(mode == LMR_CONSTRAIN_NORM || mode == LMR_CONSTRAIN_STD, "ascent_update: bad mode %d", mode);
    LMR_CHECK_ARG(x && g && B >= 1 && B <= 65535 && n >= 2, "ascent_update: bad arguments B=%d n=%d", B, n);
    if (ws_bytes < lmr_constrain_workspace_bytes(B, n)) {
        set_error("ascent_update: workspace too small");
        return LMR_ERR_WORKSPACE;
    }
    const int S = (n + CCH - 1) / CCH;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // one CTA per image with the image in shared memory when it fits (bit-identical to the two launches below; LMR_ASCENT_FUSED=0 forces those)
    const size_t fused_smem = ((size_t)((n + 3) & ~3) + (size_t)S * 4) * sizeof(float);
    const char* env = getenv("LMR_ASCENT_FUSED");
    if (fused_smem <= 220 * 1024 && !(env && env[0] == '0')) {
        static size_t cache = 0;
        if (fused_smem > cache) {
            LMR_CUDA_OK(cudaFuncSetAttribute(ascent_update_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_smem));
            cache = fused_smem;
        }
        ascent_update_fused_kernel<<<B, AUF_NT, fused_smem, st>>>(mode, x, g, step_size, n, target, mean, has_min, pmin, has_max, pmax, eps);
        LMR_LAUNCH_OK();
        return 0;
    }
    dim3 grid(S, B);
    constrain_stats_kernel<<<grid, 256, 0, st>>>(mode, x, n, mean, static_cast<float*>(ws), g, step_size);
    LMR_LAUNCH_OK();
    constrain_apply_kernel<<<grid, 256, 0, st>>>(mode, x, n, target, mean, has_min, pmin, has_max, pmax, eps, static_cast<const float*>(ws), x, nullptr, g,
                                                 step_size);
    LMR_LAUNCH_OK();
    return 0;
}

extern "C" int lmr_constrain_bwd(int mode, const float* x, const float* g_z, const float* stats, int B, int n, float target, float mean,
                                 int has_min, float pmin, int has_max, float pmax, float eps, float* g_x, void* ws, size_t ws_bytes,
                                 void* stream) {
    LMR_CHECK_ARG(mode == LMR_CONSTRAIN_NORM || mode == LMR_CONSTRAIN_STD, "constrain: bad mode %d", mode);
    LMR_CHECK_ARG(B >= 1 && n >= 2 && B <= 65535, "constrain: bad sizes B=%d n=%d", B, n);
    if (ws_bytes < lmr_constrain_workspace_bytes(B, n)) {
        set_error("constrain_bwd: workspace too small");
        return LMR_ERR_WORKSPACE;
    }
    const int S = (n + CCH - 1) / CCH;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 grid(S, B);
    constrain_bwd_stats_kernel<<<grid, 256, 0, st>>>(mode, x, g_z, stats, n, target, mean, has_min, pmin, has_max, pmax, eps, static_cast<float*>(ws));
    LMR_LAUNCH_OK();
    constrain_bwd_apply_kernel<<<grid, 256, 0, st>>>(mode, x, g_z, stats, n, target, mean, has_min, pmin, has_max, pmax, eps,
                                                     static_cast<const float*>(ws), g_x);
    LMR_LAUNCH_OK();
    return 0;
}

// pixels per CTA of the response partials: 512 when the groups alone fill the machine, down to 128 for few groups (match with a handful of
// targets, the learn step's single neuron) so that about two CTAs per SM exist.  A pure function of (G, HW): the workspace size follows from it.
static inline int simresp_chunk(int G, int HW) {
    int rch = RCH;
    while (rch > RCH / 4 && (long long)G * ((HW + rch - 1) / rch) < 2 * 148) rch >>= 1;
    return rch;
}

extern "C" size_t lmr_simresp_workspace_bytes(int G, int NB, int Fmax, int HW) {
    const int rch = simresp_chunk(G, HW);
    const int S = (HW + rch - 1) / rch;
    return (size_t)G * S * NB * Fmax * sizeof(float);
}

extern "C" int lmr_simresp_fwd(const float* img, long long gstride, int NB, int G, int C, int HW, const float* filters, int Fmax, const int* nfilt,
                               const int* neuron_of_group, float eps, float* act, float* rmax, int* argmax, void* ws, size_t ws_bytes,
                               void* stream) {
    LMR_CHECK_ARG(NB >= 1 && G >= 1 && C >= 1 && HW >= 1 && Fmax >= 1, "simresp: bad sizes");
    LMR_CHECK_ARG(G <= 65535, "simresp: G=%d > 65535 groups per call", G);
    if (ws_bytes < lmr_simresp_workspace_bytes(G, NB, Fmax, HW)) {
        set_error("simresp_fwd: workspace too small");
        return LMR_ERR_WORKSPACE;
    }
    const int rch = simresp_chunk(G, HW);
    const int S = (HW + rch - 1) / rch;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    simresp_partial_kernel<<<dim3(S, G), 256, 0, st>>>(img, gstride, NB, C, HW, filters, Fmax, nfilt, neuron_of_group, static_cast<float*>(ws), rch);
    LMR_LAUNCH_OK();
    simresp_finalize_kernel<<<(G * NB + 3) / 4, 128, 0, st>>>(static_cast<const float*>(ws), S, NB, G, Fmax, nfilt, neuron_of_group, eps, act, rmax,
                                                                   argmax);
    LMR_LAUNCH_OK();
    return 0;
}

extern "C" int lmr_simresp_bwd(const float* g_act, const float* rmax, const int* argmax, long long gstride, int NB, int G, int C, int HW,
                               const float* filters, int Fmax, const int* neuron_of_group, float* g_img, int accumulate, void* stream) {
    LMR_CHECK_ARG(NB >= 1 && G >= 1 && C >= 1 && HW >= 1, "simresp_bwd: bad sizes");
    LMR_CHECK_ARG(NB <= 65535 && G <= 65535, "simresp_bwd: NB/G too large for one call");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 grid((HW + 255) / 256, NB, gstride ? G : 1);
    simresp_bwd_kernel<<<grid, 256, 0, st>>>(g_act, rmax, argmax, gstride, NB, G, C, HW, filters, Fmax, neuron_of_group, g_img, accumulate);
    LMR_LAUNCH_OK();
    return 0;
}

extern "C" size_t lmr_simclr_workspace_bytes(int N, int C, int HW) {
    const int S = (C * HW + GCH - 1) / GCH;
    return (size_t)S * N * N * sizeof(float);
}

extern "C" int lmr_simclr_fwd(const float* img, const float* mask, int N, int C, int HW, float pmin, float pmax, const float* pos,
                              const float* neg, float T, float eps, float* reg, float* K, void* ws, size_t ws_bytes, void* stream) {
    LMR_CHECK_ARG(N >= 2 && N <= GMAXN, "simclr: N=%d out of range [2,%d]", N, GMAXN);
    LMR_CHECK_ARG(pmax > pmin, "simclr: pixel bounds required (pmin < pmax)");
    if (ws_bytes < lmr_simclr_workspace_bytes(N, C, HW)) {
        set_error("simclr_fwd: workspace too small");
        return LMR_ERR_WORKSPACE;
    }
    const int S = (C * HW + GCH - 1) / GCH;
    const float c0 = (pmin + pmax) / 2, g1 = 2.f / (pmax - pmin), g2 = (pmax - pmin) / 2.f;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t smem1 = (size_t)N * (GCH + 1) * sizeof(float);
    static size_t smem_set = 0;
    if (smem1 > smem_set) {
        LMR_CUDA_OK(cudaFuncSetAttribute(simclr_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
        smem_set = smem1;
    }
    simclr_gram_kernel<<<S, 256, smem1, st>>>(img, mask, N, C, HW, c0, g1, g2, static_cast<float*>(ws));
    LMR_LAUNCH_OK();
    const size_t smem2 = ((size_t)4 * N * N + 5 * N) * sizeof(float);
    static size_t smem2_set = 0;
    if (smem2 > 40 * 1024 && smem2 > smem2_set) {
        LMR_CUDA_OK(cudaFuncSetAttribute(simclr_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        smem2_set = smem2;
    }
    simclr_finalize_kernel<<<1, 1024, smem2, st>>>(static_cast<const float*>(ws), S, N, pos, neg, T, eps, reg, K);
    LMR_LAUNCH_OK();
    return 0;
}

extern "C" int lmr_simclr_bwd(const float* img, const float* mask, const float* K, const float* g_reg, float scale, int N, int C, int HW,
                              float pmin, float pmax, float* g_img, int accumulate, void* stream) {
    LMR_CHECK_ARG(N >= 2 && N <= GMAXN, "simclr: N=%d out of range [2,%d]", N, GMAXN);
    const float c0 = (pmin + pmax) / 2, g1 = 2.f / (pmax - pmin), g2 = (pmax - pmin) / 2.f;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int n = C * HW;
    simclr_bwd_kernel<<<(n + 127) / 128, 128, (size_t)N * N * sizeof(float), st>>>(img, mask, K, g_reg, scale, N, C, HW, c0, g1, g2, g_img, accumulate);
    LMR_LAUNCH_OK();
    return 0;
}

extern "C" int lmr_match_epoch_stats(const float* act, const float* mei_act, int D, int N, float* out3, void* stream) {
    LMR_CHECK_ARG(act && mei_act && out3 && D >= 1 && N >= 1, "match_epoch_stats: bad arguments");
    match_stats_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(act, mei_act, D, N, out3);
    LMR_LAUNCH_OK();
    return 0;
}

extern "C" int lmr_adam_step(float* p, const float* g, float* m, float* v, int n, float lr, float b1, float b2, float eps, int* step,
                             void* stream) {
    LMR_CHECK_ARG(n >= 1 && step != nullptr, "adam: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n <= 256 * 1024) {  // one thread per element, one launch
        adam_ticket_kernel<<<(n + 255) / 256, 256, 0, st>>>(p, g, m, v, n, lr, b1, b2, eps, step);
        LMR_LAUNCH_OK();
        return 0;
    }
    const int blocks = 1024;
    adam_kernel<<<blocks, 256, 0, st>>>(p, g, m, v, n, lr, b1, b2, eps, step);
    LMR_LAUNCH_OK();
    counter_inc_kernel<<<1, 1, 0, st>>>(step);
    LMR_LAUNCH_OK();
    return 0;
}

extern "C" int lmr_adam_step_multi(const lmr_adam_group* groups, int count, float lr, float b1, float b2, float eps, void* stream) {
    LMR_CHECK_ARG(groups != nullptr && count >= 1 && count <= LMR_ADAM_MAX_GROUPS, "adam_multi: 1..%d groups", LMR_ADAM_MAX_GROUPS);
    post::AdamGroups a;
    a.count = count, a.first[0] = 0;
    for (int k = 0; k < count; ++k) {
        const lmr_adam_group& g = groups[k];
        LMR_CHECK_ARG(g.p && g.g && g.exp_avg && g.exp_avg_sq && g.step && g.n >= 1 && g.n <= 256 * 1024, "adam_multi: bad group %d", k);
        a.g[k] = g;
        a.first[k + 1] = a.first[k] + (g.n + 255) / 256;
    }
    for (int k = count; k < LMR_ADAM_MAX_GROUPS; ++k) a.g[k] = groups[0], a.first[k + 1] = a.first[count];
    post::adam_multi_kernel<<<a.first[count], 256, 0, static_cast<cudaStream_t>(stream)>>>(a, lr, b1, b2, eps);
    LMR_LAUNCH_OK();
    return 0;
}

extern "C" int lmr_mei_ascent(float* x, int G, int C, int HW, const float* filters, int Fmax, const int* nfilt, const int* neuron_of_group,
                              float step_size, int iters, int mode, float target, float mean, float pmin, float pmax, float act_eps,
                              int apply_post_first, float* fevals, void* stream) {
    LMR_CHECK_ARG(G >= 1 && C >= 1 && HW >= 2 && iters >= 0, "mei: bad sizes");
    LMR_CHECK_ARG(mode == LMR_CONSTRAIN_NORM || mode == LMR_CONSTRAIN_STD, "mei: bad mode %d", mode);
    const size_t smem = ((size_t)(C * HW + 3) / 4 * 4 + (size_t)(Fmax + 3) / 4 * 4 + (size_t)Fmax * (MEI_THREADS / 32)) * sizeof(float);
    LMR_CHECK_ARG(smem <= 200 * 1024, "mei: image of %d floats does not fit the shared-memory resident kernel", C * HW);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    static size_t smem_set = 0;
    if (smem > smem_set) {
        LMR_CUDA_OK(cudaFuncSetAttribute(mei_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    mei_kernel<<<G, MEI_THREADS, smem, st>>>(x, C, HW, filters, Fmax, nfilt, neuron_of_group, step_size, iters, mode, target, mean, pmin, pmax, act_eps,
                                            apply_post_first, fevals);
    LMR_LAUNCH_OK();
    return 0;
}
