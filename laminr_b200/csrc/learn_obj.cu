// Fused learn-stage objective: ONE persistent launch for everything between the INR forward and the INR backward of a learn step
// (reference laminr/inv_temp_learn_match.py:187-204 with FixEnergyNormClip / StandardizeClip (laminr/utils/img_transforms.py:5-65),
// SingleNeuronModel over a simulated neuron (laminr/utils/general.py:5-12, neuron_models/utils/simulated_model.py:8-17), apply_mask
// (laminr/utils/mask.py:7-22) and SimCLROnGrid.reg_term (laminr/contrastive_loss.py:139-154)), forward AND backward:
//
//     z = constrain(x);  act_n = (elu(max_f <z_n, w_f>) + 1)/2 + eps;  reg = simclr(mask(z));
//     loss = -mean_n(act_n / mei_act) + g_reg * reg           (g_reg = -reg_scale, a device scalar)
//     g_x  = d loss / d x
//
// Every stage is HBM/L2-bound scan/reduce work on 0.8-3.2 MB, so the cost of the unfused version is launch latency (12 launches).  Here
// each CTA owns one pixel slice of ALL N images (kept in shared memory across the phases); per-image and gram reductions go through
// per-CTA partials in global memory, a software grid barrier (all CTAs co-resident: cooperative launch, grid <= #SMs) and a fixed-order
// (deterministic) combine.  4 grid barriers (5 in STD mode).
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace lmr {
namespace lobj {

constexpr int NT = 512;  // 16 warps: the phases are latency-bound scans over a 68-pixel slice of 20 images, flattened over all threads

struct Params {
    int mode, N, C, HW, F;
    float target, mean, pmin, pmax, eps_con;
    int has_min, has_max;
    const float* x;        // [N, C, HW]
    const float* bank;     // [F, HW] filters of the template neuron
    float act_eps, inv_mei_act;
    const float* mask;     // [HW]
    const float* pos;      // [N, N]
    const float* neg;      // [N, N]
    float T, eps_clr, c0, g1, g2;
    const float* g_reg;    // device scalar: d loss / d reg
    float *z, *stats, *act, *rmax, *reg, *Kmat, *loss, *g_x;
    int* argmax;
    // workspace
    unsigned int* sync;    // [0] arrival counter (monotonic), [1] launch epoch
    float *part_a, *part_b, *fin, *part_c;
    int PS;                // pixels per slice
    FastDiv fdPS, fdF, fdN, fdC;  // the flattened loops divide by these
    long long* trace;      // debug (LMR_OBJ_TRACE=1): CTA 0's phase time stamps (globaltimer ns), else null
};

__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// all CTAs of the (co-resident) grid arrive, then continue.  `target` = arrivals expected since the counter was zero.
__device__ __forceinline__ void grid_sync(unsigned int* counter, unsigned int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        // release (cumulative over the CTA's writes ordered before it by the bar.sync above) / acquire at gpu scope: no separate fences
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        while ((int)(ld_acquire(counter) - target) < 0) {
        }
    }
    __syncthreads();
}

__device__ __forceinline__ float masked_value(float x, float m, float c0, float g1, float g2) {
    // rescale(x, pmin, pmax, -1, 1) * mask, then rescale back (same op order as mask.py:7-22)
    const float r = (x - c0) * g1 + 0.f;
    return (r * m - 0.f) * g2 + c0;
}
__device__ __forceinline__ float clamp_pass(float y, int has_min, float pmin, int has_max, float pmax) {
    return ((!has_min || y >= pmin) && (!has_max || y <= pmax)) ? 1.f : 0.f;  // torch.clamp backward: inclusive
}

// 8 consecutive lanes sum `count` values spaced `stride` apart in a fixed order (all loads of a lane in flight together, then a shuffle
// tree): deterministic.  The values were written by other CTAs of this launch: the loads bypass L1.
constexpr int GRP = 8, MAXLD = 20;  // up to GRP * MAXLD = 160 partials (grid <= 148 CTAs)
// (not inlined: five call sites x 20 predicated loads were a tenth of the kernel's code, and the kernel runs through its ~100 KB of code once per
// launch with a cold instruction cache -- `no_inst` was the top stall reason of its one-shot sections)
__device__ __noinline__ float group_strided_sum(const float* src, int count, size_t stride, int lane) {
    const int sub = lane & (GRP - 1);
    float v[MAXLD];
#pragma unroll
    for (int u = 0; u < MAXLD; ++u) {
        const int i = sub + GRP * u;
        v[u] = i < count ? __ldcg(src + (size_t)i * stride) : 0.f;
    }
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int u = 0; u < MAXLD; u += 2) s0 += v[u], s1 += v[u + 1];
    s0 += s1;
    const unsigned gmask = 0xffu << (lane & ~(GRP - 1));  // only this 8-lane group takes part (other groups of the warp may be idle)
    s0 += __shfl_xor_sync(gmask, s0, 1);
    s0 += __shfl_xor_sync(gmask, s0, 2);
    s0 += __shfl_xor_sync(gmask, s0, 4);
    return s0;
}
__device__ __forceinline__ void cp_async4(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}

__device__ __forceinline__ void stamp(const Params& p, int idx) {
    if (p.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
        long long tns;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tns));
        p.trace[idx] = tns;
    }
}

__global__ void __launch_bounds__(NT, 1) learn_objective_kernel(Params p) {
    extern __shared__ __align__(16) float sm[];
    const int N = p.N, C = p.C, HW = p.HW, F = p.F, NC = N * C;
    const int G = gridDim.x, b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5, NW = NT / 32;
    const int PS = p.PS, PSP = PS | 1;
    const int p0 = b * PS, np = max(0, min(PS, HW - p0));
    float* xs = sm;                    // [NC][PSP]  x
    float* xm = xs + NC * PSP;         // [NC][PSP]  masked z
    float* gy = xm + NC * PSP;         // [NC][PSP]  gradient wrt the pre-clamp image
    float* ws = gy + NC * PSP;         // [F][PSP]   filter slice
    float* zs = ws + F * PSP;          // [N][PSP]   sum_c z
    float* mk = zs + N * PSP;          // [PSP]
    float* sd = mk + PSP;              // [N]
    float* mu = sd + N;                // [N]
    float* coef = mu + N;              // [N]  g_act * 1/2 * elu'(rmax)
    float* dotu = coef + N;            // [NC][2] per (image, channel) row
    float* dotn = dotu + 2 * NC;       // [N][2]  per image, over all CTAs
    int* amax = reinterpret_cast<int*>(dotn + 2 * N);  // [N]
    float* fin = reinterpret_cast<float*>(amax + N);   // [N*F + N*N]
    float* cosm = fin + N * F + N * N; // [N][N]
    float* Sm = cosm + N * N;          // [N][N]
    float* rho = Sm + N * N;           // [N]
    float* rowc = rho + N;             // [N][4]
    float* Ks = rowc + 4 * N;          // [N][N]
    float* posS = Ks + N * N;          // [N][N]
    float* negS = posS + N * N;        // [N][N]
    float* actv = negS + N * N;        // [N]  activations (summed into the loss)
    float* scal = actv + N;            // [4]
    stamp(p, 0);
    const unsigned int epoch = ld_acquire(p.sync + 1);
    const float g_reg = __ldg(p.g_reg);  // issued up front: an L2 round trip that would otherwise sit on the critical path after barrier 3
    const unsigned int nbar = p.mode == LMR_CONSTRAIN_NORM ? 4u : 5u;
    unsigned int bar_no = 0;
    auto barrier = [&]() {
        ++bar_no;
        grid_sync(p.sync, (epoch * nbar + bar_no) * (unsigned int)G);
    };

    // ---- phase 0: load the slice (every global read in flight at once: cp.async, 4 bytes each), per-image statistics
#pragma unroll 1
    for (int r = warp; r < NC; r += NW) {
        const float* src = p.x + (size_t)r * HW + p0;
        for (int e = lane; e < PS; e += 32) {
            if (e < np) cp_async4(xs + r * PSP + e, src + e);
            else xs[r * PSP + e] = 0.f;
        }
    }
#pragma unroll 1
    for (int f = warp; f < F; f += NW) {
        const float* src = p.bank + (size_t)f * HW + p0;
        for (int e = lane; e < PS; e += 32) {
            if (e < np) cp_async4(ws + f * PSP + e, src + e);
            else ws[f * PSP + e] = 0.f;
        }
    }
#pragma unroll 1
    for (int e = t; e < PS; e += NT) {
        if (e < np) cp_async4(mk + e, p.mask + p0 + e);
        else mk[e] = 0.f;
    }
#pragma unroll 1
    for (int i = t; i < N * N; i += NT) cp_async4(posS + i, p.pos + i), cp_async4(negS + i, p.neg + i);
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    stamp(p, 1);
    if (p.mode == LMR_CONSTRAIN_NORM) {
#pragma unroll 1
        for (int n = warp; n < N; n += NW) {
            float acc = 0.f;
            for (int c = 0; c < C; ++c)
                for (int e = lane; e < np; e += 32) {
                    const float u = xs[(n * C + c) * PSP + e] - p.mean;
                    acc = fmaf(u, u, acc);
                }
            acc = warp_sum(acc);
            if (lane == 0) p.part_a[(size_t)b * N + n] = acc;
        }
        barrier();
#pragma unroll 1
        for (int n = t / GRP; n < N; n += NT / GRP) {
            const float tot = group_strided_sum(p.part_a + n, G, N, lane);
            if ((lane & (GRP - 1)) == 0) sd[n] = sqrtf(tot), mu[n] = 0.f;
        }
    } else {
#pragma unroll 1
        for (int n = warp; n < N; n += NW) {
            float acc = 0.f;
            for (int c = 0; c < C; ++c)
                for (int e = lane; e < np; e += 32) acc += xs[(n * C + c) * PSP + e];
            acc = warp_sum(acc);
            if (lane == 0) p.part_a[(size_t)b * N + n] = acc;
        }
        barrier();
#pragma unroll 1
        for (int n = t / GRP; n < N; n += NT / GRP) {
            const float tot = group_strided_sum(p.part_a + n, G, N, lane);
            if ((lane & (GRP - 1)) == 0) mu[n] = tot / (float)(C * HW);
        }
        __syncthreads();
#pragma unroll 1
        for (int n = warp; n < N; n += NW) {
            float acc = 0.f;
            for (int c = 0; c < C; ++c)
                for (int e = lane; e < np; e += 32) {
                    const float u = xs[(n * C + c) * PSP + e] - mu[n];
                    acc = fmaf(u, u, acc);
                }
            acc = warp_sum(acc);
            if (lane == 0) p.part_c[(size_t)b * 2 * N + n] = acc;
        }
        barrier();
#pragma unroll 1
        for (int n = t / GRP; n < N; n += NT / GRP) {
            const float tot = group_strided_sum(p.part_c + n, G, 2 * N, lane);
            if ((lane & (GRP - 1)) == 0) sd[n] = sqrtf(tot / (float)(C * HW - 1));  // unbiased (Tensor.std default)
        }
    }
    __syncthreads();
    if (b == 0 && p.stats != nullptr) {
#pragma unroll 1
        for (int n = t; n < N; n += NT) p.stats[2 * n] = sd[n], p.stats[2 * n + 1] = mu[n];
    }

    stamp(p, 2);
    // ---- phase 1: z (written out), sum_c z, masked z; partial responses and partial gram  (work items flattened over all threads)
#pragma unroll 1
    for (int i = t; i < NC * PS; i += NT) {
        const int r = fd_div(i, p.fdPS), e = i - r * PS, n = fd_div(r, p.fdC);
        const float sdn = sd[n] + p.eps_con;
        float y;
        if (p.mode == LMR_CONSTRAIN_NORM) y = (xs[r * PSP + e] - p.mean) / sdn * p.target + p.mean;
        else y = p.target * (xs[r * PSP + e] - mu[n]) / sdn + p.mean;
        if (p.has_min) y = fmaxf(y, p.pmin);
        if (p.has_max) y = fminf(y, p.pmax);
        if (e < np) {
            p.z[(size_t)r * HW + p0 + e] = y;
            xm[r * PSP + e] = masked_value(y, mk[e], p.c0, p.g1, p.g2);
            gy[r * PSP + e] = y;  // z, until phase 3 overwrites it
        } else {
            xm[r * PSP + e] = 0.f, gy[r * PSP + e] = 0.f;
        }
    }
    __syncthreads();
#pragma unroll 1
    for (int i = t; i < N * PS; i += NT) {
        const int n = fd_div(i, p.fdPS), e = i - n * PS;
        float v = 0.f;
        for (int c = 0; c < C; ++c) v += gy[(n * C + c) * PSP + e];
        zs[n * PSP + e] = v;
    }
    __syncthreads();
    {
        float* outR = p.part_b + (size_t)b * (N * F + N * N);
        float* outG = outR + N * F;
        // one thread per (image, filter) response and per (i <= j) gram entry; rows have the odd stride PSP: conflict-free across filters
#pragma unroll 1
        for (int it = t; it < N * F + N * N; it += NT) {
            if (it < N * F) {
                const int n = fd_div(it, p.fdF), f = it - n * F;
                const float* a = zs + n * PSP;
                const float* w = ws + f * PSP;
                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
                int e = 0;
#pragma unroll 2
                for (; e + 3 < PS; e += 4)
                    s0 = fmaf(a[e], w[e], s0), s1 = fmaf(a[e + 1], w[e + 1], s1), s2 = fmaf(a[e + 2], w[e + 2], s2), s3 = fmaf(a[e + 3], w[e + 3], s3);
                for (; e < PS; ++e) s0 = fmaf(a[e], w[e], s0);
                outR[it] = (s0 + s1) + (s2 + s3);
            } else {
                const int pr = it - N * F, i = fd_div(pr, p.fdN), j = pr - i * N;
                if (j < i) continue;
                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
                for (int c = 0; c < C; ++c) {
                    const float* a = xm + (i * C + c) * PSP;
                    const float* w = xm + (j * C + c) * PSP;
                    int e = 0;
#pragma unroll 2
                    for (; e + 3 < PS; e += 4)
                        s0 = fmaf(a[e], w[e], s0), s1 = fmaf(a[e + 1], w[e + 1], s1), s2 = fmaf(a[e + 2], w[e + 2], s2), s3 = fmaf(a[e + 3], w[e + 3], s3);
                    for (; e < PS; ++e) s0 = fmaf(a[e], w[e], s0);
                }
                const float g = (s0 + s1) + (s2 + s3);
                outG[i * N + j] = g;
                outG[j * N + i] = g;
            }
        }
    }
    stamp(p, 3);
    barrier();
    stamp(p, 4);
    // ---- phase 2: two-level combine of the partials (value v is summed by CTA v % G), then everybody reads the totals
    {
        const int nval = N * F + N * N;
#pragma unroll 1
        for (int v = b + G * (t / GRP); v < nval; v += G * (NT / GRP)) {
            const float tot = group_strided_sum(p.part_b + v, G, (size_t)nval, lane);
            if ((lane & (GRP - 1)) == 0) p.fin[v] = tot;
        }
    }
    stamp(p, 5);
    barrier();
    stamp(p, 6);
    {
        const int nval = N * F + N * N;
#pragma unroll 1
        for (int v = t; v < nval; v += NT) fin[v] = __ldcg(p.fin + v);
    }
    __syncthreads();
    // responses: first maximum wins (torch.max), act = (elu(m)+1)/2 + eps.  8 lanes per image scan the filters f = sub, sub + 8, ..; ties go to the
    // lower filter index in the lane-local scan and in the shuffle combine
    const int sub = lane & (GRP - 1);
    const unsigned gmask = 0xffu << (lane & ~(GRP - 1));
#pragma unroll 1
    for (int n = t / GRP; n < N; n += NT / GRP) {
        float best = -INFINITY;
        int bi = 0x7fffffff;
#pragma unroll 1
        for (int f = sub; f < F; f += GRP) {
            const float r = fin[n * F + f];
            if (r > best) best = r, bi = f;
        }
#pragma unroll
        for (int o = 1; o < GRP; o <<= 1) {
            const float ob = __shfl_xor_sync(gmask, best, o);
            const int oi = __shfl_xor_sync(gmask, bi, o);
            if (ob > best || (ob == best && oi < bi)) best = ob, bi = oi;
        }
        if (sub == 0) {
            if (bi == 0x7fffffff) bi = 0;  // every response is NaN / -inf: torch.max returns index 0 of the first maximum
            const float e = best > 0.f ? best : expm1f(best);
            const float act = (e + 1.f) / 2.f + p.act_eps;
            amax[n] = bi;
            coef[n] = (-p.inv_mei_act / (float)N) * 0.5f * (best > 0.f ? 1.f : expf(best));
            actv[n] = act;
            if (b == 0) {
                p.act[n] = act;
                if (p.rmax) p.rmax[n] = best;
                if (p.argmax) p.argmax[n] = bi;
            }
        }
    }
    // contrastive term and its backward coefficient matrix K (same maths as post.cu simclr_finalize_kernel)
    const float* gram = fin + N * F;
#pragma unroll 1
    for (int pr = t; pr < N * N; pr += NT) {
        const int i = fd_div(pr, p.fdN), j = pr - i * N;
        const float c = gram[pr] / (sqrtf(gram[i * N + i]) * sqrtf(gram[j * N + j]));
        cosm[pr] = c;
        Sm[pr] = expf(c / p.T);
        if (i == j) rho[i] = sqrtf(gram[pr]);
    }
    __syncthreads();
    // 8 lanes per row: masked row sums of S (fixed shuffle tree: deterministic); the last warp adds up the activations meanwhile
#pragma unroll 1
    for (int i = t / GRP; i < N; i += NT / GRP) {
        float ps = 0.f, np_ = 0.f, ns = 0.f, nn = 0.f;
#pragma unroll 1
        for (int j = sub; j < N; j += GRP) {
            ps = fmaf(Sm[i * N + j], posS[i * N + j], ps), np_ += posS[i * N + j];
            ns = fmaf(Sm[i * N + j], negS[i * N + j], ns), nn += negS[i * N + j];
        }
#pragma unroll
        for (int o = 1; o < GRP; o <<= 1) {
            ps += __shfl_xor_sync(gmask, ps, o), np_ += __shfl_xor_sync(gmask, np_, o);
            ns += __shfl_xor_sync(gmask, ns, o), nn += __shfl_xor_sync(gmask, nn, o);
        }
        if (sub == 0) {
            const float num = (ps == 0.f ? 0.f : ps / np_) + p.eps_clr;
            const float den = (ns == 0.f ? 0.f : ns / nn) + p.eps_clr;
            rowc[i * 4 + 0] = ps == 0.f ? 0.f : (p.T / (2.f * N)) / (np_ * num);
            rowc[i * 4 + 1] = ns == 0.f ? 0.f : (p.T / (2.f * N)) / (nn * den);
            rowc[i * 4 + 2] = logf(num / den);
        }
    }
    if (b == 0 && warp == NW - 1) {  // (N <= 64: two terms per lane)
        float a0 = (lane < N ? actv[lane] : 0.f) + (lane + 32 < N ? actv[lane + 32] : 0.f);
        a0 = warp_sum(a0);
        if (lane == 0) scal[0] = a0;
    }
    __syncthreads();
    if (b == 0 && warp == 0) {
        float ls = (lane < N ? rowc[lane * 4 + 2] : 0.f) + (lane + 32 < N ? rowc[(lane + 32) * 4 + 2] : 0.f);
        ls = warp_sum(ls);
        if (lane == 0) {
            const float reg = ls / (float)N * p.T / 2.f;
            p.reg[0] = reg;
            p.loss[0] = -(scal[0] * p.inv_mei_act) / (float)N + g_reg * reg;
        }
    }
#pragma unroll 1
    for (int pr = t; pr < N * N; pr += NT) {
        const int i = fd_div(pr, p.fdN);
        Sm[pr] = (posS[pr] * rowc[i * 4] - negS[pr] * rowc[i * 4 + 1]) * Sm[pr] / p.T;  // Gamma_ij = A_ij S_ij / T
    }
    __syncthreads();
#pragma unroll 1
    for (int i = t / GRP; i < N; i += NT / GRP) {
        float kap = 0.f;
#pragma unroll 1
        for (int j = sub; j < N; j += GRP) kap = fmaf(Sm[i * N + j] + Sm[j * N + i], cosm[i * N + j], kap);
#pragma unroll
        for (int o = 1; o < GRP; o <<= 1) kap += __shfl_xor_sync(gmask, kap, o);
        if (sub == 0) rowc[i * 4 + 3] = kap;
    }
    __syncthreads();
#pragma unroll 1
    for (int pr = t; pr < N * N; pr += NT) {
        const int i = fd_div(pr, p.fdN), j = pr - i * N;
        float k = (Sm[i * N + j] + Sm[j * N + i]) / (rho[i] * rho[j]);
        if (i == j) k -= rowc[i * 4 + 3] / (rho[i] * rho[i]);
        Ks[pr] = k;
        if (b == 0 && p.Kmat) p.Kmat[pr] = k;
    }
    __syncthreads();

    stamp(p, 7);
    // ---- phase 3: g_z = response backward + contrastive backward; gradient wrt the pre-clamp image; per-image sums for the projection
#pragma unroll 1
    for (int i = t; i < NC * PS; i += NT) {
        const int r = fd_div(i, p.fdPS), e = i - r * PS, n = fd_div(r, p.fdC), c = r - n * C;
        float v = 0.f;
        if (e < np) {
            const float sdn = sd[n] + p.eps_con, mun = p.mode == LMR_CONSTRAIN_NORM ? p.mean : mu[n];
            const float* kr = Ks + n * N;
            const float* xc = xm + c * PSP + e;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            int jn = 0;
#pragma unroll 1
            for (; jn + 3 < N; jn += 4) {
                a0 = fmaf(kr[jn], xc[(jn * C) * PSP], a0), a1 = fmaf(kr[jn + 1], xc[((jn + 1) * C) * PSP], a1);
                a2 = fmaf(kr[jn + 2], xc[((jn + 2) * C) * PSP], a2), a3 = fmaf(kr[jn + 3], xc[((jn + 3) * C) * PSP], a3);
            }
#pragma unroll 1
            for (; jn < N; ++jn) a0 = fmaf(kr[jn], xc[(jn * C) * PSP], a0);
            const float acc = (a0 + a1) + (a2 + a3);
            const float gz = coef[n] * ws[amax[n] * PSP + e] + g_reg * (p.g1 * mk[e] * p.g2) * acc;
            const float u = xs[r * PSP + e] - mun;
            float y;
            if (p.mode == LMR_CONSTRAIN_NORM) y = u / sdn * p.target + p.mean;
            else y = p.target * u / sdn + p.mean;
            v = gz * clamp_pass(y, p.has_min, p.pmin, p.has_max, p.pmax);
        }
        gy[r * PSP + e] = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int r = warp; r < NC; r += NW) {
        const float mun = p.mode == LMR_CONSTRAIN_NORM ? p.mean : mu[r / C];
        float a0 = 0.f, a1 = 0.f;
        for (int e = lane; e < np; e += 32) {
            const float v = gy[r * PSP + e];
            a0 += v;
            a1 = fmaf(v, xs[r * PSP + e] - mun, a1);
        }
        a0 = warp_sum(a0), a1 = warp_sum(a1);
        if (lane == 0) dotu[2 * r] = a0, dotu[2 * r + 1] = a1;  // per (image, channel) row; combined over channels below
    }
    __syncthreads();
#pragma unroll 1
    for (int n = t; n < N; n += NT) {
        float a0 = 0.f, a1 = 0.f;
        for (int c = 0; c < C; ++c) a0 += dotu[2 * (n * C + c)], a1 += dotu[2 * (n * C + c) + 1];
        p.part_c[(size_t)b * 2 * N + n] = a0, p.part_c[(size_t)b * 2 * N + N + n] = a1;
    }
    stamp(p, 8);
    barrier();
    stamp(p, 9);
#pragma unroll 1
    for (int n = t / GRP; n < 2 * N; n += NT / GRP) {  // part_c rows are [a0[0..N) | a1[0..N)]
        const float tot = group_strided_sum(p.part_c + n, G, 2 * N, lane);
        if ((lane & (GRP - 1)) == 0) dotn[n < N ? 2 * n : 2 * (n - N) + 1] = tot;
    }
    __syncthreads();
    // ---- phase 4: g_x   NORM: k gy - target dot / ((s+eps)^2 s) u ;  STD: k (gy - mean(gy)) - target dot / ((sd+eps)^2 (n-1) sd) (x-mu)
    // (the three per-image coefficients once per image -- three divisions per ELEMENT before -- in the finalize scratch rowc[n][0..2])
#pragma unroll 1
    for (int n = t; n < N; n += NT) {
        const float s_ = sd[n];
        const float denom = (s_ + p.eps_con) * (s_ + p.eps_con) * s_ * (p.mode == LMR_CONSTRAIN_NORM ? 1.f : (float)(C * HW - 1));
        rowc[4 * n + 0] = p.target / (s_ + p.eps_con);
        rowc[4 * n + 1] = p.target * dotn[2 * n + 1] / denom;
        rowc[4 * n + 2] = p.mode == LMR_CONSTRAIN_NORM ? 0.f : dotn[2 * n] / (float)(C * HW);
    }
    __syncthreads();
#pragma unroll 1
    for (int i = t; i < NC * PS; i += NT) {
        const int r = fd_div(i, p.fdPS), e = i - r * PS, n = fd_div(r, p.fdC);
        if (e >= np) continue;
        const float k = rowc[4 * n + 0], c2 = rowc[4 * n + 1], gmean = rowc[4 * n + 2];
        const float mun = p.mode == LMR_CONSTRAIN_NORM ? p.mean : mu[n];
        p.g_x[(size_t)r * HW + p0 + e] = k * (gy[r * PSP + e] - gmean) - c2 * (xs[r * PSP + e] - mun);
    }
    stamp(p, 10);
    // the epoch word is only written here, after every CTA has passed the last barrier (all of them read it at kernel start)
    if (b == 0 && t == 0) {
        __threadfence();
        p.sync[1] = epoch + 1;
    }
}

static size_t smem_floats(int N, int C, int F, int PS) {
    const int PSP = PS | 1, NC = N * C;
    return (size_t)3 * NC * PSP + (size_t)F * PSP + (size_t)N * PSP + PSP + 6 * (size_t)N + 2 * (size_t)NC + (size_t)N /*amax*/ + (size_t)N * F + N * N + 5 * (size_t)N * N + 6 * (size_t)N + 20;
}

struct Plan {
    int G, PS;
    size_t smem, off_a, off_b, off_fin, off_c, total;
};

static bool make_plan(int N, int C, int HW, int F, Plan* pl) {
    const int sms = num_sms();
    int G = sms;
    int PS = (HW + G - 1) / G;
    if (PS < 16) PS = 16;  // tiny images: fewer CTAs
    G = (HW + PS - 1) / PS;
    pl->G = G, pl->PS = PS;
    pl->smem = smem_floats(N, C, F, PS) * sizeof(float);
    size_t off = 256;  // sync words
    auto take = [&](size_t nfl) {
        size_t r = off;
        off += align_up(nfl * sizeof(float), 256);
        return r;
    };
    pl->off_a = take((size_t)G * N);
    pl->off_b = take((size_t)G * (N * F + N * N));
    pl->off_fin = take((size_t)N * F + N * N);
    pl->off_c = take((size_t)G * 2 * N);
    pl->total = off;
    return pl->smem <= 200 * 1024 && N <= 64 && F <= 128;
}

}  // namespace lobj
}  // namespace lmr

using namespace lmr;

extern "C" size_t lmr_learn_objective_workspace_bytes(int N, int C, int HW, int F) {
    lobj::Plan pl;
    if (!lobj::make_plan(N, C, HW, F, &pl)) return 0;
    return pl.total;
}

extern "C" int lmr_learn_objective(int mode, const float* x, int N, int C, int HW, float target, float mean, int has_min, float pmin, int has_max,
                                   float pmax, float eps_con, const float* bank, int F, float act_eps, float mei_act, const float* mask,
                                   const float* pos, const float* neg, float temperature, float eps_clr, const float* g_reg, float* z,
                                   float* stats, float* act, float* rmax, int* argmax, float* reg, float* Kmat, float* loss, float* g_x,
                                   void* workspace, size_t workspace_bytes, void* stream) {
    LMR_CHECK_ARG(mode == LMR_CONSTRAIN_NORM || mode == LMR_CONSTRAIN_STD, "learn_objective: bad mode %d", mode);
    LMR_CHECK_ARG(N >= 2 && C >= 1 && HW >= 1 && F >= 1, "learn_objective: bad sizes N=%d C=%d HW=%d F=%d", N, C, HW, F);
    LMR_CHECK_ARG(x && bank && mask && pos && neg && g_reg && z && act && reg && loss && g_x, "learn_objective: null pointer");
    lobj::Plan pl;
    LMR_CHECK_ARG(lobj::make_plan(N, C, HW, F, &pl), "learn_objective: configuration N=%d C=%d HW=%d F=%d does not fit the fused kernel (use the unfused ops)", N, C,
                  HW, F);
    if (workspace_bytes < pl.total) {
        set_error("learn_objective: workspace %zu < %zu bytes", workspace_bytes, pl.total);
        return LMR_ERR_WORKSPACE;
    }
    lobj::Params p;
    p.mode = mode, p.N = N, p.C = C, p.HW = HW, p.F = F;
    p.target = target, p.mean = mean, p.pmin = pmin, p.pmax = pmax, p.eps_con = eps_con, p.has_min = has_min, p.has_max = has_max;
    p.x = x, p.bank = bank, p.act_eps = act_eps, p.inv_mei_act = 1.f / mei_act, p.mask = mask, p.pos = pos, p.neg = neg;
    p.T = temperature, p.eps_clr = eps_clr;
    // apply_mask: rescale(x, pmin, pmax, -1, 1) then back (mask.py:7-22); without bounds the mask multiplies the raw image
    if (has_min && has_max) p.c0 = (pmin + pmax) / 2.f, p.g1 = 2.f / (pmax - pmin), p.g2 = (pmax - pmin) / 2.f;
    else p.c0 = 0.f, p.g1 = 1.f, p.g2 = 1.f;
    p.g_reg = g_reg;
    p.z = z, p.stats = stats, p.act = act, p.rmax = rmax, p.argmax = argmax, p.reg = reg, p.Kmat = Kmat, p.loss = loss, p.g_x = g_x;
    char* base = static_cast<char*>(workspace);
    p.sync = reinterpret_cast<unsigned int*>(base);
    p.part_a = reinterpret_cast<float*>(base + pl.off_a), p.part_b = reinterpret_cast<float*>(base + pl.off_b);
    p.fin = reinterpret_cast<float*>(base + pl.off_fin), p.part_c = reinterpret_cast<float*>(base + pl.off_c);
    p.PS = pl.PS;
    p.fdPS = make_fastdiv(pl.PS), p.fdF = make_fastdiv(F), p.fdN = make_fastdiv(N), p.fdC = make_fastdiv(C);
    static const bool tracing = getenv("LMR_OBJ_TRACE") != nullptr;
    static long long* trace_dev = nullptr;
    p.trace = nullptr;
    if (tracing) {
        if (trace_dev == nullptr) cudaMalloc(&trace_dev, 16 * sizeof(long long));
        p.trace = trace_dev;
    }
    static size_t cache = 0;
    if (pl.smem > cache) {
        LMR_CUDA_OK(cudaFuncSetAttribute(lobj::learn_objective_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
        cache = pl.smem;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(pl.G), cfg.blockDim = dim3(lobj::NT), cfg.dynamicSmemBytes = pl.smem, cfg.stream = static_cast<cudaStream_t>(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;  // guarantees that all CTAs are co-resident (the software grid barrier needs it)
    cfg.attrs = attr, cfg.numAttrs = 1;
    LMR_CUDA_OK(cudaLaunchKernelEx(&cfg, lobj::learn_objective_kernel, p));
    if (tracing) {  // debug only: synchronises and prints CTA 0's phase time line
        long long h[16];
        cudaStreamSynchronize(cfg.stream);
        cudaMemcpy(h, trace_dev, sizeof(h), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[obj trace] ns since entry: loaded %lld | stats+bar1 %lld | phase1 done %lld | bar2 %lld | combine %lld | bar3 %lld | finalize %lld | phase3 %lld | bar4 %lld | end %lld\n",
                h[1] - h[0], h[2] - h[0], h[3] - h[0], h[4] - h[0], h[5] - h[0], h[6] - h[0], h[7] - h[0], h[8] - h[0], h[9] - h[0], h[10] - h[0]);
    }
    return 0;
}
