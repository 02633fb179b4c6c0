// Shared pieces of the INR kernels (fp32 CUDA-core path in inr_simt.cu, tcgen05 path in inr_tc.cu): network / tiling
// descriptors, the prep kernels (latent encoding, separable layer-0 pieces), the affine coordinate transform, the layer-0
// weight-gradient kernel and the deterministic final reductions.  Global workspace arrays use a row stride of SHP = 52.
#pragma once
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace lmr {
namespace inr {

constexpr int TR_MAX = 128;  // rows (= threads) per tile
constexpr int SHP = 52;      // storage stride of 50-wide rows in global workspaces (16-byte aligned)

struct Net {
    int hid, nl, pe, ape, C, in_dim;
    __host__ __device__ size_t size0() const { return (size_t)hid * in_dim + hid; }
    __host__ __device__ size_t sizeh() const { return (size_t)hid * hid + hid; }
    __host__ __device__ size_t offW(int l) const { return l == 0 ? 0 : size0() + (size_t)(l - 1) * sizeh(); }
    __host__ __device__ size_t offb(int l) const { return offW(l) + (l == 0 ? (size_t)hid * in_dim : (size_t)hid * hid); }
    __host__ __device__ size_t offWout() const { return size0() + (size_t)(nl - 1) * sizeh(); }
    __host__ __device__ size_t offbout() const { return offWout() + (size_t)C * hid; }
    __host__ __device__ size_t count() const { return offbout() + C; }
};

struct Tiling {
    int N, P, D, NC, TP, tilesP, nChunks, numTiles;
};

// Kept-activation record of one 128-row tile (tensor-core path, LMR_PREC_KEEP_ACTIVATIONS): the 32 KB split-bf16 operand-tile images of
// h_1 .. h_{nl-1} as the backward's MMAs read them, then the tile's output values y[c][row] (fp32: the backward needs dy = g (1 - y^2)
// per row and would otherwise have to rebuild y from all column groups of the row).
constexpr size_t KEEP_TILE_BYTES = 32768, KEEP_Y_BYTES = 2048;  // 4 channels x 128 rows x 4 bytes
__host__ __device__ inline size_t keep_record_bytes(int nl) { return (size_t)(nl - 1) * KEEP_TILE_BYTES + KEEP_Y_BYTES; }

static inline Tiling make_tiling(int N, int P, int D, int tp_cap, int rows = TR_MAX) {
    Tiling t;
    t.N = N, t.P = P, t.D = D;
    t.NC = N < rows ? N : rows;
    t.TP = rows / t.NC;
    if (t.TP > tp_cap) t.TP = tp_cap;
    if (t.TP > P) t.TP = P;
    t.tilesP = (P + t.TP - 1) / t.TP;
    t.nChunks = (N + t.NC - 1) / t.NC;
    t.numTiles = D * t.tilesP * t.nChunks;
    return t;
}

struct Args {
    Net net;
    Tiling tl;
    const float* params;
    const float* B;
    const float* auxB;
    const float* grid;
    const float* coords;
    const float* coord_shift;
    lmr_affine aff;
    int has_aff;
    // workspace (prep kernel outputs)
    float* alpha;  // [N, 2*ape]
    float* Avec;   // [N, HP]
    float* W0cT;   // [2*pe, HP]
    float* Cpg;    // [D*P, HP]  layer-0 coordinate part Cp[d,p][o] = sum_k beta[d,p][k] W0c[o][k]   (tensor-core path)
    unsigned char* wtimg;  // (nl-1) x 16 KB: split-bf16 (hi | lo) K-major core-matrix images of the hidden weight tiles (tensor-core path)
    float* betaG;  // [D*P, 2*pe] sin/cos(coords.B) (tensor-core path, D == 1: consumed by the layer-0 weight-gradient kernel), else null
    unsigned char* hsave;  // [numTiles] kept-activation records (keep_record_bytes: operand-tile images of h_1..h_{nl-1}, then y) written by the forward for the backward (tensor-core path, opt-in), else null
    int want_cp;
    int reuse_beta;  // betaG already holds this call's table (LMR_PREC_REUSE_COORD_TABLE): the prep blocks load it instead of evaluating sin/cos
    int ppb;       // pixels per l0 / affine-gradient block
    int ppb_prep;  // pixels per prep block (two blocks per SM: the prep blocks are latency-bound)
    FastDiv fd_nb, fd_ppb_prep;  // divisions by 2*pe and by ppb_prep in the prep blocks' flattened copy loops
    // forward
    float* img;  // [D,N,C,P]
    // backward
    const float* g_img;
    float* G0;        // [D*P, HP]   (learn: sum_n delta0)
    float* partials;  // [grid, partial_size]
    float* affpart;   // [D, tilesP, 6]  (match)
    float* dAred;     // [N, HP]  sum over CTAs of the dA partials
    int want_weights, want_affine;
    long long* trace;  // debug: per-event clock64 stamps of CTA 0 (LMR_TC_TRACE=1), else null
    int stagger;       // two-stream backward: cycles by which stream 1 starts behind stream 0
};

// ---------------------------------------------------------------------------------------------------------------
// prep: latent encoding (inr.py:371-386,401-409), A[n] = b0 + W0[:,1:1+2a] alpha_n, packed transpose of W0's coord block
// ---------------------------------------------------------------------------------------------------------------
__host__ __device__ inline size_t partial_size(const Net& nt, const Tiling& tl, int HP) {
    return (size_t)(nt.nl - 1) * nt.hid * nt.hid + (size_t)(nt.nl - 1) * HP + (size_t)nt.C * HP + 4 + (size_t)tl.NC * HP;
}

// affine constants of target d -> c[0..3] = M, c[4..5] = t, c[6..7] = tc, c[8..9] = c2 (the no_grad shift)   
__device__ inline void affine_consts(const lmr_affine& af, int d, float* c) {
    float sn, cs;
    sincosf(af.angles[d], &sn, &cs);
    const float s0 = af.scalings[d * af.scal_cols], s1 = af.scalings[d * af.scal_cols + (af.scal_cols > 1 ? 1 : 0)];
    const float sh0 = af.shears[2 * d], sh1 = af.shears[2 * d + 1];
    const float rs00 = cs * s0, rs01 = -sn * s1, rs10 = sn * s0, rs11 = cs * s1;
    const float m00 = rs00 + rs01 * sh1, m01 = rs00 * sh0 + rs01, m10 = rs10 + rs11 * sh1, m11 = rs10 * sh0 + rs11;
    const float t0 = af.translation[2 * d], t1 = af.translation[2 * d + 1];
    const float tc0 = af.tc ? af.tc[0] : 0.f, tc1 = af.tc ? af.tc[1] : 0.f;
    const float a0 = (af.shift_from ? af.shift_from[2 * d] : 0.f) + (af.shift_to ? af.shift_to[0] : 0.f);
    const float a1 = (af.shift_from ? af.shift_from[2 * d + 1] : 0.f) + (af.shift_to ? af.shift_to[1] : 0.f);
    const float c20 = m00 * a0 + m01 * a1 + t0, c21 = m10 * a0 + m11 * a1 + t1;
    c[0] = m00, c[1] = m01, c[2] = m10, c[3] = m11;
    c[4] = t0, c[5] = t1, c[6] = tc0, c[7] = tc1;
    c[8] = c20, c[9] = c21;
}

__device__ inline float2 transformed_coord(const Args& a, const float* affc, int p) {
    float c0 = a.coords[2 * p], c1 = a.coords[2 * p + 1];
    if (!a.has_aff) {
        if (a.coord_shift) c0 += a.coord_shift[0], c1 += a.coord_shift[1];
        return make_float2(c0, c1);
    }
    const float r0 = c0 - affc[6], r1 = c1 - affc[7];
    float o0 = (affc[0] * r0 + affc[1] * r1 + affc[4]) + affc[6];
    float o1 = (affc[2] * r0 + affc[3] * r1 + affc[5]) + affc[7];
    o0 += affc[8], o1 += affc[9];
    if (a.aff.clamp) o0 = fminf(fmaxf(o0, -1.f), 1.f), o1 = fminf(fmaxf(o1, -1.f), 1.f);
    return make_float2(o0, o1);
}

// ---------------------------------------------------------------------------------------------------------------
// prep (ONE launch, heterogeneous blocks):
//   blocks [0, N)      latent n: alpha_n = sin/cos(a_n.auxB) (inr.py:371-386,401-409), A[n] = b0 + W0[:,1:1+2a] alpha_n, A[n][HID] = 20
//   block  N           packed transpose W0cT[k][o] of W0's coordinate block
//   blocks (N, ...)    (want_cp) 64 pixels each: beta = sin/cos(coords.B) (inr.py:390-397), Cp[d,p][o] = sum_k beta[k] W0c[o][k]
// ---------------------------------------------------------------------------------------------------------------
constexpr int PREP_PPB_MAX = 72;  // pixels per block (multiple of 4)

static inline int choose_ppb(int total_pixels, int blocks_per_sm = 1) {
    int ppb = (total_pixels + blocks_per_sm * num_sms() - 1) / (blocks_per_sm * num_sms());
    ppb = (ppb + 3) / 4 * 4;
    if (ppb < 16) ppb = 16;
    if (ppb > PREP_PPB_MAX) ppb = PREP_PPB_MAX;
    return ppb;
}

template <int HID>
__global__ void __launch_bounds__(256) prep_all_kernel(Args a) {
    constexpr int HP = (HID + 3) / 4 * 4;
    extern __shared__ __align__(16) float psm[];
    const Net& nt = a.net;
    const int N = a.tl.N, na = 2 * nt.ape, nb = 2 * nt.pe, pe = nt.pe;
    const float* W0 = a.params;
    const int bid = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (bid < N) {
        float* als = psm;  // [na]
        const float g = a.grid[bid];
        float sg, cg;
        sincosf(g, &sg, &cg);
        for (int m = t; m < nt.ape; m += blockDim.x) {
            float sn, cs;
            sincosf(sg * a.auxB[m] + cg * a.auxB[nt.ape + m], &sn, &cs);
            als[m] = sn, als[nt.ape + m] = cs;
            a.alpha[(size_t)bid * na + m] = sn, a.alpha[(size_t)bid * na + nt.ape + m] = cs;
        }
        __syncthreads();
        const float* b0 = a.params + nt.offb(0);
        // one warp per output (coalesced reads of W0's row); the rows of a warp's outputs are fetched together (latency-bound otherwise)
        constexpr int MAXO = (HP + 7) / 8, MAXK = 8;  // 8 warps; na <= 256 = 32 * MAXK
        float wv[MAXO][MAXK];
#pragma unroll
        for (int x = 0; x < MAXO; ++x) {
            const int o = warp + 8 * x;
#pragma unroll
            for (int y = 0; y < MAXK; ++y) {
                const int k = lane + 32 * y;
                wv[x][y] = (o < HID && k < na) ? W0[(size_t)o * nt.in_dim + 1 + k] : 0.f;
            }
        }
#pragma unroll
        for (int x = 0; x < MAXO; ++x) {
            const int o = warp + 8 * x;
            if (o >= HP) continue;
            float acc = 0.f;
#pragma unroll
            for (int y = 0; y < MAXK; ++y) {
                const int k = lane + 32 * y;
                if (k < na) acc = fmaf(wv[x][y], als[k], acc);
            }
            acc = warp_sum(acc);
            if (o < HID) acc += b0[o];
            else if (o == HID) acc = 20.f;  // tanh(20) == 1 in fp32: the tensor-core path keeps a constant-1 feature (bias column) in slot HID
            else acc = 0.f;
            if (lane == 0) a.Avec[(size_t)bid * HP + o] = acc;
        }
    } else if (bid == N) {
        if (!a.want_cp) {  // fp32 path: no pixel blocks; the tensor-core path's first pixel block writes W0cT from its shared-memory copy
            constexpr int U = 8;
            for (int i0 = t; i0 < HID * nb; i0 += U * 256) {
                float w[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int i = i0 + u * 256;
                    w[u] = i < HID * nb ? W0[(size_t)(i / nb) * nt.in_dim + 1 + na + i % nb] : 0.f;
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int i = i0 + u * 256;
                    if (i < HID * nb) a.W0cT[(size_t)(i % nb) * HP + i / nb] = w[u];
                }
            }
            for (int i = t; i < nb * (HP - HID); i += blockDim.x) a.W0cT[(size_t)(i / (HP - HID)) * HP + HID + i % (HP - HID)] = 0.f;
        }
    } else if (a.want_cp && bid >= N + 1 + (a.tl.D * a.tl.P + a.ppb_prep - 1) / a.ppb_prep) {
        // weight-tile images: W_l[o][k] for k < HID, column HID = bias b_l[o], entry [HID][HID] = 20 (regenerates the constant-1 feature
        // through tanh), rest 0; bf16 hi and lo = bf16(v - hi) halves in the no-swizzle K-major core-matrix layout of a [64 x 64] tile.
        constexpr int U = 4;
        const int wb = bid - (N + 1 + (a.tl.D * a.tl.P + a.ppb_prep - 1) / a.ppb_prep);
        float v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int e = wb * (U * 256) + u * 256 + t, l = e / 4096 + 1, i = e % 4096, o = i / 64, k = i % 64;
            v[u] = 0.f;
            if (l < nt.nl) {
                if (o < HID) v[u] = k < HID ? a.params[nt.offW(l) + o * HID + k] : (k == HID ? a.params[nt.offb(l) + o] : 0.f);
                else if (o == HID && k == HID) v[u] = 20.f;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int e = wb * (U * 256) + u * 256 + t, l = e / 4096 + 1, i = e % 4096, o = i / 64, k = i % 64;
            if (l < nt.nl) {
                const __nv_bfloat16 hi = __float2bfloat16(v[u]);
                const __nv_bfloat16 lo = __float2bfloat16(v[u] - __bfloat162float(hi));
                unsigned char* dst = a.wtimg + (size_t)(l - 1) * 16384 + ((o & 7) * 16 + (o >> 3) * 128 + (k >> 3) * 1024) + (k & 7) * 2;
                *reinterpret_cast<__nv_bfloat16*>(dst) = hi;
                *reinterpret_cast<__nv_bfloat16*>(dst + 8192) = lo;
            }
        }
    } else {
        // pixels [q0, q0 + nq): beta (kept k-major in shared memory: 4 pixels per float4), then a register-blocked 4 pixel x 4 output GEMM.
        // Every global read of the block is issued up front (the block is latency-bound otherwise).
        const int ppb = a.ppb_prep;
        float* Wc = psm;             // [nb][HP]
        float* bet = Wc + nb * HP;   // [nb][ppb]
        __shared__ float cst[2][10];
        __shared__ float2 cxy[PREP_PPB_MAX];
        __shared__ float Bsm[2 * 64];
        const int total = a.tl.D * a.tl.P;
        const int q0 = (bid - N - 1) * ppb, nq = min(ppb, total - q0);
        const int d0 = q0 / a.tl.P;
        const bool reuse = a.reuse_beta != 0;  // the table of this call is already in betaG (same pixel grid as the previous call): load it
        if (!reuse && a.has_aff && t < 2 && d0 + t < a.tl.D) affine_consts(a.aff, d0 + t, cst[t]);  // a block usually spans at most two targets
        constexpr int UB = 16;  // the first 4096 table entries of the slice (36 pixels x 100 features at 100x100) are fetched together with W0's block below
        float bv[UB];
        if (reuse) {
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const int i = t + u * 256;
                bv[u] = i < nq * nb ? a.betaG[(size_t)q0 * nb + i] : 0.f;
            }
        }
        if (t >= 64 && t < 64 + nb) Bsm[t - 64] = a.B[t - 64];
        {
            constexpr int U = 10;
            for (int i0 = t; i0 < HID * nb; i0 += U * 256) {
                float w[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int i = i0 + u * 256;
                    int o, k;
                    fd_divmod(i, a.fd_nb, &o, &k);
                    w[u] = i < HID * nb ? W0[(size_t)o * nt.in_dim + 1 + na + k] : 0.f;
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int i = i0 + u * 256;
                    int o, k;
                    fd_divmod(i, a.fd_nb, &o, &k);
                    if (i < HID * nb) Wc[k * HP + o] = w[u];
                }
            }
        }
        for (int i = t; i < nb * (HP - HID); i += blockDim.x) Wc[(i / (HP - HID)) * HP + HID + i % (HP - HID)] = 0.f;
        if (reuse) {
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const int i = t + u * 256;
                int jj, k;
                fd_divmod(i, a.fd_nb, &jj, &k);
                if (i < ppb * nb) bet[k * ppb + jj] = bv[u];  // pixel-major -> k-major (pixels past nq: zeros)
            }
            for (int i = t + UB * 256; i < ppb * nb; i += 256) {
                int jj, k;
                fd_divmod(i, a.fd_nb, &jj, &k);
                bet[k * ppb + jj] = i < nq * nb ? a.betaG[(size_t)q0 * nb + i] : 0.f;
            }
        }
        __syncthreads();
        if (bid == N + 1)  // packed transpose W0cT[k][o] of W0's coordinate block (consumed by the match backward)
            for (int i = t; i < nb * HP; i += blockDim.x) a.W0cT[i] = Wc[i];
        if (!reuse && t < nq) {  // transformed coordinates of the block's pixels, once per pixel
            const int q = q0 + t;
            float loc[10];
            const float* cc = cst[0];
            if (a.has_aff) {
                const int dq = q / a.tl.P - d0;
                if (dq < 2) cc = cst[dq];
                else affine_consts(a.aff, q / a.tl.P, loc), cc = loc;  // tiny images: a block may span more than two targets
            }
            cxy[t] = transformed_coord(a, cc, q % a.tl.P);
        }
        __syncthreads();
        if (!reuse)
            for (int i = t; i < pe * ppb; i += blockDim.x) {
                int m, jj;
                fd_divmod(i, a.fd_ppb_prep, &m, &jj);
                float sn = 0.f, cs = 0.f;
                if (jj < nq) {
                    const float2 c = cxy[jj];
                    sincosf(c.x * Bsm[m] + c.y * Bsm[pe + m], &sn, &cs);
                }
                bet[m * ppb + jj] = sn, bet[(pe + m) * ppb + jj] = cs;
            }
        __syncthreads();
        if (!reuse && a.betaG != nullptr) {  // pixel-major copy for the layer-0 weight-gradient kernel (coalesced over k)
            for (int i = t; i < nq * nb; i += blockDim.x) {
                int jj, k;
                fd_divmod(i, a.fd_nb, &jj, &k);
                a.betaG[(size_t)(q0 + jj) * nb + k] = bet[k * ppb + jj];
            }
        }
        const int npg = ppb / 4, nog = HP / 4;
        if (t < npg * nog) {
            const int pg = t / nog, og = t % nog;
            float4 acc[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) acc[x] = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4* bp = reinterpret_cast<const float4*>(bet) + pg;
            const float4* wp = reinterpret_cast<const float4*>(Wc) + og;
#pragma unroll 4
            for (int k = 0; k < nb; ++k) {
                const float4 bb = bp[k * npg];
                const float4 ww = wp[k * nog];
                acc[0].x = fmaf(bb.x, ww.x, acc[0].x), acc[0].y = fmaf(bb.x, ww.y, acc[0].y), acc[0].z = fmaf(bb.x, ww.z, acc[0].z), acc[0].w = fmaf(bb.x, ww.w, acc[0].w);
                acc[1].x = fmaf(bb.y, ww.x, acc[1].x), acc[1].y = fmaf(bb.y, ww.y, acc[1].y), acc[1].z = fmaf(bb.y, ww.z, acc[1].z), acc[1].w = fmaf(bb.y, ww.w, acc[1].w);
                acc[2].x = fmaf(bb.z, ww.x, acc[2].x), acc[2].y = fmaf(bb.z, ww.y, acc[2].y), acc[2].z = fmaf(bb.z, ww.z, acc[2].z), acc[2].w = fmaf(bb.z, ww.w, acc[2].w);
                acc[3].x = fmaf(bb.w, ww.x, acc[3].x), acc[3].y = fmaf(bb.w, ww.y, acc[3].y), acc[3].z = fmaf(bb.w, ww.z, acc[3].z), acc[3].w = fmaf(bb.w, ww.w, acc[3].w);
            }
#pragma unroll
            for (int x = 0; x < 4; ++x)
                if (4 * pg + x < nq) reinterpret_cast<float4*>(a.Cpg + (size_t)(q0 + 4 * pg + x) * HP)[og] = acc[x];
        }
    }
}

struct TileIdx {
    int d, p0, n0, np, nn;  // target, first pixel, first latent, #pixels, #latents
};

__device__ inline TileIdx decode_tile(const Tiling& tl, int tile) {
    TileIdx t;
    // the common shapes need no division: one latent chunk (N latents fit the tile's rows) and, in learn, one target
    int chunk = 0, rest = tile;
    if (tl.nChunks != 1) chunk = tile % tl.nChunks, rest = tile / tl.nChunks;
    int pb = rest;
    t.d = 0;
    if (rest >= tl.tilesP) t.d = rest / tl.tilesP, pb = rest - t.d * tl.tilesP;
    t.p0 = pb * tl.TP;
    t.n0 = chunk * tl.NC;
    t.np = min(tl.TP, tl.P - t.p0);
    t.nn = min(tl.NC, tl.N - t.n0);
    return t;
}

// ---------------------------------------------------------------------------------------------------------------
// learn: layer-0 coordinate block  dW0c[o][k] = sum_{d,p} G0[d,p][o] beta[d,p][k]   (split over pixel chunks)
// ---------------------------------------------------------------------------------------------------------------
constexpr int L0_THREADS = 352;

template <int HID>
__global__ void __launch_bounds__(L0_THREADS) l0_partial_kernel(Args a, float* l0part /*[grid, HID*2pe]*/) {
    constexpr int HP = (HID + 3) / 4 * 4;
    extern __shared__ __align__(16) float sm[];
    const Net& nt = a.net;
    const int pe = nt.pe, nb = 2 * pe, ppb = a.ppb;
    float* Gs = sm;                  // [ppb][HP]
    float* Bs = Gs + ppb * HP;       // [ppb][nb]   (pixel-major: 4 consecutive k per float4)
    const int total = a.tl.D * a.tl.P;   // learn: D == 1 and no affine, but keep it general (coords of target d)
    const int nchunks = (total + ppb - 1) / ppb;
    // register block: 4 outputs x 4 encoding features
    const int nOG = HP / 4, nKG = nb / 4;
    const int t = threadIdx.x;
    const bool owner = t < nOG * nKG;
    const int og = t % nOG, kg = t / nOG;
    float4 acc[4];
#pragma unroll
    for (int x = 0; x < 4; ++x) acc[x] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        const int q0 = ch * ppb, nq = min(ppb, total - q0);
        __syncthreads();
        for (int i = t; i < nq * HP / 4; i += blockDim.x) reinterpret_cast<float4*>(Gs)[i] = reinterpret_cast<const float4*>(a.G0 + (size_t)q0 * HP)[i];
        if (a.betaG != nullptr) {
            for (int i = t; i < nq * nb / 4; i += blockDim.x) reinterpret_cast<float4*>(Bs)[i] = reinterpret_cast<const float4*>(a.betaG + (size_t)q0 * nb)[i];
        } else {
            for (int i = t; i < nq * pe; i += blockDim.x) {
                const int jj = i / pe, m = i % pe;
                const int q = q0 + jj, p = q % a.tl.P;
                float cst[10];
                if (a.has_aff) affine_consts(a.aff, q / a.tl.P, cst);
                const float2 c = transformed_coord(a, cst, p);
                float sn, cs;
                sincosf(c.x * a.B[m] + c.y * a.B[pe + m], &sn, &cs);
                Bs[jj * nb + m] = sn, Bs[jj * nb + pe + m] = cs;
            }
        }
        __syncthreads();
        if (owner) {
            const float4* gp = reinterpret_cast<const float4*>(Gs) + og;
            const float4* bp = reinterpret_cast<const float4*>(Bs) + kg;
#pragma unroll 4
            for (int jj = 0; jj < nq; ++jj) {
                const float4 g = gp[jj * nOG], b = bp[jj * nKG];
                acc[0].x = fmaf(g.x, b.x, acc[0].x), acc[0].y = fmaf(g.x, b.y, acc[0].y), acc[0].z = fmaf(g.x, b.z, acc[0].z), acc[0].w = fmaf(g.x, b.w, acc[0].w);
                acc[1].x = fmaf(g.y, b.x, acc[1].x), acc[1].y = fmaf(g.y, b.y, acc[1].y), acc[1].z = fmaf(g.y, b.z, acc[1].z), acc[1].w = fmaf(g.y, b.w, acc[1].w);
                acc[2].x = fmaf(g.z, b.x, acc[2].x), acc[2].y = fmaf(g.z, b.y, acc[2].y), acc[2].z = fmaf(g.z, b.z, acc[2].z), acc[2].w = fmaf(g.z, b.w, acc[2].w);
                acc[3].x = fmaf(g.w, b.x, acc[3].x), acc[3].y = fmaf(g.w, b.y, acc[3].y), acc[3].z = fmaf(g.w, b.z, acc[3].z), acc[3].w = fmaf(g.w, b.w, acc[3].w);
            }
        }
    }
    if (owner) {
        float* dst = l0part + (size_t)blockIdx.x * HID * nb;
#pragma unroll
        for (int x = 0; x < 4; ++x)
            if (4 * og + x < HID) reinterpret_cast<float4*>(dst + (size_t)(4 * og + x) * nb)[kg] = acc[x];
    }
}

__device__ __forceinline__ void cp_async16_g(void* dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}

// sum of `count` values spaced `stride` apart: 8 independent chains (the loads are L2 round trips), combined in a fixed order
__device__ __forceinline__ float strided_sum8(const float* src, int count, size_t stride) {
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int b = 0;
    for (; b + 16 <= count; b += 16) {  // 16 round trips in flight, added in the order of the 8-wide loop below (same bits)
        float v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = src[(size_t)(b + u) * stride];
#pragma unroll
        for (int u = 0; u < 8; ++u) s[u] += v[u];
#pragma unroll
        for (int u = 0; u < 8; ++u) s[u] += v[8 + u];
    }
    for (; b + 8 <= count; b += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) s[u] += src[(size_t)(b + u) * stride];
    }
    for (int u = 0; b < count; ++b, ++u) s[u] += src[(size_t)b * stride];
    return ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
}

// the same for values written by OTHER blocks of the running launch (behind a grid barrier): the loads bypass L1
__device__ __forceinline__ float strided_sum8_cg(const float* src, int count, size_t stride) {
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int b = 0;
    for (; b + 16 <= count; b += 16) {  // 16 round trips in flight, added in the order of the 8-wide loop below (same bits)
        float v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = __ldcg(src + (size_t)(b + u) * stride);
#pragma unroll
        for (int u = 0; u < 8; ++u) s[u] += v[u];
#pragma unroll
        for (int u = 0; u < 8; ++u) s[u] += v[8 + u];
    }
    for (; b + 8 <= count; b += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) s[u] += __ldcg(src + (size_t)(b + u) * stride);
    }
    for (int u = 0; b < count; ++b, ++u) s[u] += __ldcg(src + (size_t)b * stride);
    return ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
}

__device__ __forceinline__ float quad_sum(float v) {  // lanes 4k..4k+3 -> (v0 + v1) + (v2 + v3) on every lane of the quad
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}

// learn: deterministic reductions of the per-CTA partials into g_params (torch layout).
// (1) one thread per output element sums it over the partial blocks (coalesced across threads): hidden / output layers, the
//     coordinate block of W0 (from l0part) and dAred[n][o]; (2) the latent block of W0 and b0 follow from dAred.
template <int HID>
__global__ void __launch_bounds__(128) reduce_partials_kernel(Args a, const float* l0part, int nPartMain, int nPartL0, float* g_params) {
    constexpr int HP = (HID + 3) / 4 * 4;
    const Net& nt = a.net;
    const Tiling& tl = a.tl;
    const size_t psz = partial_size(nt, tl, HP);
    const int L = nt.nl, na = 2 * nt.ape, nb = 2 * nt.pe;
    const size_t total = nt.count();
    const size_t offA = (size_t)(L - 1) * HID * HID + (size_t)(L - 1) * HP + (size_t)nt.C * HP + 4;
    // 4 consecutive lanes share one output element: lane `sub` sums the partials b = sub (mod 4), then a fixed-order quad combine
    const size_t gid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t i = gid >> 2;
    const int sub = (int)(gid & 3);
    if (i >= total + (size_t)tl.N * HP) return;  // whole quads leave together (the bound is per element)
    if (i >= total) {  // dAred
        const size_t e = i - total;
        const float part = strided_sum8(a.partials + offA + e + (size_t)sub * psz, (nPartMain - sub + 3) / 4, 4 * psz);
        const float tot = quad_sum(part);
        if (sub == 0) a.dAred[e] = tot;
        return;
    }
    const float* src = a.partials;
    size_t off = 0, stride = psz;
    int nPart = nPartMain;
    if (i < nt.size0()) {
        if (i >= (size_t)nt.hid * nt.in_dim) return;  // b0: finish_w0a
        const int o = i / nt.in_dim, col = i % nt.in_dim;
        if (col < 1 + na) return;                      // template / latent columns: finish_w0a   (quad-uniform exits)
        src = l0part, off = (size_t)o * nb + (col - 1 - na), stride = (size_t)HID * nb, nPart = nPartL0;
    } else if (i < nt.offWout()) {
        const size_t r = i - nt.size0();
        const int l = r / nt.sizeh();
        const size_t w = r % nt.sizeh();
        off = w < (size_t)HID * HID ? (size_t)l * HID * HID + w : (size_t)(L - 1) * HID * HID + (size_t)l * HP + (w - (size_t)HID * HID);
    } else {
        const size_t r = i - nt.offWout();
        const size_t base = (size_t)(L - 1) * HID * HID + (size_t)(L - 1) * HP;
        off = r < (size_t)nt.C * HID ? base + (r / HID) * HP + (r % HID) : base + (size_t)nt.C * HP + (r - (size_t)nt.C * HID);
    }
    const float part = strided_sum8(src + off + (size_t)sub * stride, (nPart - sub + 3) / 4, 4 * stride);
    const float tot = quad_sum(part);
    if (sub == 0) g_params[i] = tot;
}

template <int HID>
__global__ void __launch_bounds__(128) finish_w0a_kernel(Args a, float* g_params) {
    constexpr int HP = (HID + 3) / 4 * 4;
    const Net& nt = a.net;
    const int N = a.tl.N, na = 2 * nt.ape;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int nW = HID * (1 + na);
    if (i < nW) {
        const int o = i / (1 + na), col = i % (1 + na);
        float sum = 0.f;  // col 0: the template input is the constant 0 (inr.py:200-205)
        if (col > 0)
            for (int n = 0; n < N; ++n) sum = fmaf(a.dAred[(size_t)n * HP + o], a.alpha[(size_t)n * na + col - 1], sum);
        g_params[(size_t)o * nt.in_dim + col] = sum;
    } else if (i < nW + HID) {
        const int o = i - nW;
        float sum = 0.f;
        for (int n = 0; n < N; ++n) sum += a.dAred[(size_t)n * HP + o];
        g_params[nt.offb(0) + o] = sum;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// learn: everything behind the tile kernel of the backward AND the Adam step in ONE cooperative launch (round 2; the four launches
// l0_partial -> reduce_partials -> finish_w0a -> adam_ticket were 21 us of a 200 us step, each near the launch-latency floor):
//   phase A  (a) this block's pixel chunks of the layer-0 coordinate block  dW0c = G0^T beta  -> l0part[block]      (= l0_partial_kernel)
//            (b) this block's slice of the hidden / output layer gradients and of dA[n][o], summed over the tile kernel's per-CTA partials in
//                the fixed order of reduce_partials_kernel; Adam applied to the hidden / output layers right away
//   grid barrier (all blocks co-resident: cooperative launch)
//   phase B  (c) a slice of the coordinate block summed over the l0 partials, (d) a slice of the latent block of W0 and of b0 from dAred and alpha
//            (= finish_w0a_kernel); Adam on them
// Same summation orders and the same Adam arithmetic as the separate launches: bit-identical gradients and parameters.
// step: int32[2] = {steps taken, scratch (0 between launches)} as in lmr_adam_step; the scratch word is the arrival counter of the barrier and of
// the exit count (the last block out publishes the incremented step count and re-arms the word).
// ---------------------------------------------------------------------------------------------------------------
// (m0, v0, p0 = the element's state, loaded by the caller BEFORE the reduction that produces gi: three more round trips off the serial path)
__device__ __forceinline__ void adam_update(const AdamArgs& ad, size_t i, float gi, float m0, float v0, float p0, float step_size, float bc2_sqrt) {
    adam_element(gi, m0, v0, p0, ad.b1, ad.b2, ad.eps, step_size, bc2_sqrt, &ad.m[i], &ad.v[i], &ad.p[i]);
}

template <int HID>
__global__ void __launch_bounds__(L0_THREADS, 2) learn_tail_kernel(Args a, float* l0part, int nL0, int nPartMain, float* g_params, AdamArgs ad) {
    constexpr int HP = (HID + 3) / 4 * 4;
    extern __shared__ __align__(16) float sm[];
    const Net& nt = a.net;
    const Tiling& tl = a.tl;
    const int pe = nt.pe, nb = 2 * pe, na = 2 * nt.ape, ppb = a.ppb, L = nt.nl;
    const int t = threadIdx.x, G = (int)gridDim.x, blk = (int)blockIdx.x;
    // Two roles in phase A, side by side on every SM (the grid is 2 blocks per l0 block, all co-resident): blocks [0, nL0) compute the layer-0
    // partials, blocks [nL0, G) sum the tile kernel's partials -- both are latency-bound chains, so running them concurrently costs no more
    // than the longer one.  Phase B is spread over all G blocks.
    // Adam constants of this step (every block reads the counter before anybody can publish the new one: that happens after the last exit count)
    const int tstep = *reinterpret_cast<volatile int*>(ad.step) + 1;
    // (two powf: ~200 instructions behind an L2 round trip.  The reduction blocks need them in phase A and evaluate them here; the layer-0 blocks --
    //  the long pole of phase A, tools/trace_tail.py -- only in phase B: they start their loads first and evaluate them behind the product)
    float step_size = 0.f, bc2_sqrt = 1.f;
    if (blk >= nL0) adam_step_consts(tstep, ad.lr, ad.b1, ad.b2, &step_size, &bc2_sqrt);
    const size_t psz = partial_size(nt, tl, HP);
    const size_t nHid = nt.count() - nt.size0(), nA = nHid + (size_t)tl.N * HP;
    const size_t offA = (size_t)(L - 1) * HID * HID + (size_t)(L - 1) * HP + (size_t)nt.C * HP + 4;
    const int QB = blockDim.x / 4;  // element quads per block and pass
    // debug (LMR_TAIL_TRACE=1): globaltimer stamps of thread 0 of block 0 (layer-0 role, slots 0..7) and of block nL0 (reduction role, slots 8..15)
    auto stamp = [&](int idx) {
        if (a.trace != nullptr && t == 0 && (blk == 0 || blk == nL0)) {
            long long tns;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tns));
            a.trace[(blk == 0 ? 0 : 8) + idx] = tns;
        }
    };
    stamp(0);

    if (blk < nL0) {
        // ---- phase A (a): layer-0 coordinate block partial of this block's pixel chunks
        float* Gs = sm;             // [ppb][HP]
        float* Bs = Gs + ppb * HP;  // [ppb][nb]
        const int total = tl.D * tl.P;
        const int nchunks = (total + ppb - 1) / ppb;
        const int nOG = HP / 4, nKG = nb / 4;
        const bool owner = t < nOG * nKG;
        const int og = t % nOG, kg = t / nOG;
        float4 acc[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) acc[x] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int ch = blk; ch < nchunks; ch += nL0) {
            const int q0 = ch * ppb, nq = min(ppb, total - q0);
            __syncthreads();
            // every 16-byte piece of the two slices in flight at once (cp.async: no register staging, no dependence on the unrolling of these loops)
            for (int i = t; i < nq * HP / 4; i += blockDim.x) cp_async16_g(reinterpret_cast<float4*>(Gs) + i, reinterpret_cast<const float4*>(a.G0 + (size_t)q0 * HP) + i);
            for (int i = t; i < nq * nb / 4; i += blockDim.x) cp_async16_g(reinterpret_cast<float4*>(Bs) + i, reinterpret_cast<const float4*>(a.betaG + (size_t)q0 * nb) + i);
            asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
            __syncthreads();
            stamp(4);
            // (measured, round 2: this 4 x 4 register block runs at the shared-memory pipe's rate -- 2 broadcast 16-byte loads per 16 multiply-adds,
            //  ~185 cycles per pixel for the block's 11 warps: 6.4 of the kernel's 18 us.  8 x 8 blocks on packed fma.rn.f32x2 (4 loads per 64) need
            //  91 threads, whose 3 warps expose every load latency under the 88-register cap of two co-resident blocks: 8.2 us.  tools/trace_tail.py)
            if (owner) {
                const float4* gp = reinterpret_cast<const float4*>(Gs) + og;
                const float4* bp = reinterpret_cast<const float4*>(Bs) + kg;
#pragma unroll 4
                for (int jj = 0; jj < nq; ++jj) {
                    const float4 g = gp[jj * nOG], b = bp[jj * nKG];
                    acc[0].x = fmaf(g.x, b.x, acc[0].x), acc[0].y = fmaf(g.x, b.y, acc[0].y), acc[0].z = fmaf(g.x, b.z, acc[0].z), acc[0].w = fmaf(g.x, b.w, acc[0].w);
                    acc[1].x = fmaf(g.y, b.x, acc[1].x), acc[1].y = fmaf(g.y, b.y, acc[1].y), acc[1].z = fmaf(g.y, b.z, acc[1].z), acc[1].w = fmaf(g.y, b.w, acc[1].w);
                    acc[2].x = fmaf(g.z, b.x, acc[2].x), acc[2].y = fmaf(g.z, b.y, acc[2].y), acc[2].z = fmaf(g.z, b.z, acc[2].z), acc[2].w = fmaf(g.z, b.w, acc[2].w);
                    acc[3].x = fmaf(g.w, b.x, acc[3].x), acc[3].y = fmaf(g.w, b.y, acc[3].y), acc[3].z = fmaf(g.w, b.z, acc[3].z), acc[3].w = fmaf(g.w, b.w, acc[3].w);
                }
            }
        }
        stamp(5);
        adam_step_consts(tstep, ad.lr, ad.b1, ad.b2, &step_size, &bc2_sqrt);  // (used in phase B)
        if (owner) {
            float* dst = l0part + (size_t)blk * HID * nb;
#pragma unroll
            for (int x = 0; x < 4; ++x)
                if (4 * og + x < HID) reinterpret_cast<float4*>(dst + (size_t)(4 * og + x) * nb)[kg] = acc[x];
        }
    } else {
        // ---- phase A (b): hidden / output layers and dA[n][o] over the tile kernel's partials (4 lanes per element, fixed-order quad combine).
        //      Element list: the parameters from size0() on, then dAred.
        const int nR = G - nL0, rb = blk - nL0;
        const size_t per = (nA + nR - 1) / nR;
        const size_t e0 = (size_t)rb * per, e1 = e0 + per < nA ? e0 + per : nA;
        const int sub = t & 3;
        const size_t passes = (per + QB - 1) / QB;
        for (size_t ps = 0; ps < passes; ++ps) {
            const size_t e = e0 + ps * QB + (t >> 2);
            const bool live = e < e1;  // (whole warps run the shuffles together)
            size_t off = 0;
            if (live) {
                if (e >= nHid) {
                    off = offA + (e - nHid);
                } else {
                    const size_t i = nt.size0() + e;
                    if (i < nt.offWout()) {
                        const size_t r = i - nt.size0();
                        const int l = r / nt.sizeh();
                        const size_t w = r % nt.sizeh();
                        off = w < (size_t)HID * HID ? (size_t)l * HID * HID + w : (size_t)(L - 1) * HID * HID + (size_t)l * HP + (w - (size_t)HID * HID);
                    } else {
                        const size_t r = i - nt.offWout();
                        const size_t base = (size_t)(L - 1) * HID * HID + (size_t)(L - 1) * HP;
                        off = r < (size_t)nt.C * HID ? base + (r / HID) * HP + (r % HID) : base + (size_t)nt.C * HP + (r - (size_t)nt.C * HID);
                    }
                }
            }
            const bool upd = live && sub == 0 && e < nHid;
            const size_t pi = nt.size0() + e;
            const float m0 = upd ? ad.m[pi] : 0.f, v0 = upd ? ad.v[pi] : 0.f, p0 = upd ? ad.p[pi] : 0.f;
            const float part = live ? strided_sum8(a.partials + off + (size_t)sub * psz, (nPartMain - sub + 3) / 4, 4 * psz) : 0.f;
            const float tot = quad_sum(part);
            if (live && sub == 0) {
                if (e >= nHid) {
                    a.dAred[e - nHid] = tot;
                } else {
                    g_params[pi] = tot;
                    adam_update(ad, pi, tot, m0, v0, p0, step_size, bc2_sqrt);
                }
            }
        }
    }
    // ---- grid barrier
    __syncthreads();
    stamp(1);
    if (t == 0) {
        // release (cumulative over the block's writes ordered before it by the bar.sync above) / acquire at gpu scope: no separate fences
        asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(ad.step + 1) : "memory");
        int seen;
        do {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(ad.step + 1) : "memory");
        } while (seen < G);
    }
    __syncthreads();
    stamp(2);
    // ---- phase B, again two roles side by side: (c) on blocks [0, nL0), (d) on blocks [nL0, G)
    if (blk < nL0) {
        // (c) coordinate block of W0 over the nL0 layer-0 partials, 4 lanes per element
        const size_t nC = (size_t)HID * nb;
        const size_t per = (nC + nL0 - 1) / nL0;
        const size_t e0 = (size_t)blk * per, e1 = e0 + per < nC ? e0 + per : nC;
        const int sub = t & 3;
        const size_t passes = (per + QB - 1) / QB;
        for (size_t ps = 0; ps < passes; ++ps) {
            const size_t e = e0 + ps * QB + (t >> 2);
            const bool live = e < e1, upd = live && sub == 0;
            const int o = (int)(e / nb), k = (int)(e % nb);
            const size_t i = (size_t)o * nt.in_dim + 1 + na + k;
            const float m0 = upd ? ad.m[i] : 0.f, v0 = upd ? ad.v[i] : 0.f, p0 = upd ? ad.p[i] : 0.f;
            const float part = live ? strided_sum8_cg(l0part + e + (size_t)sub * nC, (nL0 - sub + 3) / 4, 4 * nC) : 0.f;
            const float tot = quad_sum(part);
            if (upd) {
                g_params[i] = tot;
                adam_update(ad, i, tot, m0, v0, p0, step_size, bc2_sqrt);
            }
        }
    } else {
        // (d) latent block of W0 (and the template column, whose input is the constant 0) and b0 from dAred and alpha; the N terms of an element
        //     are fetched together (a dependent chain of N L2 round trips otherwise), summed in the order of finish_w0a_kernel
        const int nR = G - nL0, rb = blk - nL0;
        const int nW = HID * (1 + na), nD = nW + HID;
        const int per = (nD + nR - 1) / nR;
        constexpr int NB8 = 8;  // terms fetched together
        for (int i = rb * per + t; i < min(nD, (rb + 1) * per); i += blockDim.x) {
            const int o = i < nW ? i / (1 + na) : i - nW, col = i < nW ? i % (1 + na) : -1;
            const size_t idx = i < nW ? (size_t)o * nt.in_dim + col : nt.offb(0) + o;
            const float m0 = ad.m[idx], v0 = ad.v[idx], p0 = ad.p[idx];
            float sum = 0.f;
            if (col != 0) {
                for (int n0 = 0; n0 < tl.N; n0 += NB8) {
                    float dv[NB8], av[NB8];
#pragma unroll
                    for (int u = 0; u < NB8; ++u) {
                        const int n = n0 + u;
                        dv[u] = n < tl.N ? __ldcg(a.dAred + (size_t)n * HP + o) : 0.f;
                        av[u] = (n < tl.N && col > 0) ? a.alpha[(size_t)n * na + col - 1] : 1.f;
                    }
#pragma unroll
                    for (int u = 0; u < NB8; ++u)
                        if (n0 + u < tl.N) sum = col > 0 ? fmaf(dv[u], av[u], sum) : sum + dv[u];
                }
            }
            g_params[idx] = sum;
            adam_update(ad, idx, sum, m0, v0, p0, step_size, bc2_sqrt);
        }
    }
    // ---- exit count: the last block out publishes the step count and re-arms the scratch word
    __syncthreads();
    stamp(3);
    if (t == 0) {
        __threadfence();
        if (atomicAdd(ad.step + 1, 1) == 2 * G - 1) {
            ad.step[1] = 0;
            ad.step[0] = tstep;
        }
    }
}

template <int HID>
static int launch_learn_tail(const Args& a, float* l0part, int grid_l0, int nPartMain, float* g_params, const AdamArgs& ad, cudaStream_t st) {
    constexpr int HP = (HID + 3) / 4 * 4;
    LMR_CHECK_ARG(a.betaG != nullptr, "learn tail: needs the kept sin/cos table of the forward workspace (D == 1)");
    LMR_CHECK_ARG((HP / 4) * (2 * a.net.pe / 4) <= L0_THREADS && a.net.pe % 2 == 0, "learn tail: pe_dim %d not supported", a.net.pe);
    LMR_CHECK_ARG(grid_l0 <= num_sms(), "learn tail: %d blocks cannot be co-resident", grid_l0);
    const size_t smem = (size_t)a.ppb * (HP + 2 * a.net.pe) * sizeof(float);
    static size_t cache = 0;
    if (smem > 40 * 1024 && smem > cache) {
        LMR_CUDA_OK(cudaFuncSetAttribute(learn_tail_kernel<HID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cache = smem;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * grid_l0), cfg.blockDim = dim3(L0_THREADS), cfg.dynamicSmemBytes = smem, cfg.stream = st;  // two roles per SM
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;  // all blocks co-resident: the software grid barrier needs it
    cfg.attrs = attr, cfg.numAttrs = 1;
    static const bool tracing = getenv("LMR_TAIL_TRACE") != nullptr;
    if (tracing) {  // debug only: synchronises and prints the phase time line of one block of each role
        static long long* trace_dev = nullptr;
        if (trace_dev == nullptr) cudaMalloc(&trace_dev, 16 * sizeof(long long));
        Args at = a;
        at.trace = trace_dev;
        LMR_CUDA_OK(cudaLaunchKernelEx(&cfg, learn_tail_kernel<HID>, at, l0part, grid_l0, nPartMain, g_params, ad));
        long long h[16];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, trace_dev, sizeof(h), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[tail trace] ns since block 0's entry: layer-0 role: slices loaded %lld | product %lld | phase A %lld | barrier %lld | phase B %lld ;  reduction role: entry %lld | phase A %lld | barrier %lld | phase B %lld\n",
                h[4] - h[0], h[5] - h[0], h[1] - h[0], h[2] - h[0], h[3] - h[0], h[8] - h[0], h[9] - h[0], h[10] - h[0], h[11] - h[0]);
        return 0;
    }
    Args an = a;
    an.trace = nullptr;  // (the tile kernel's LMR_TC_TRACE buffer is not this kernel's)
    LMR_CUDA_OK(cudaLaunchKernelEx(&cfg, learn_tail_kernel<HID>, an, l0part, grid_l0, nPartMain, g_params, ad));
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// match: gradient wrt the transformed coordinates and its per-target moments, pixel-parallel:
//   DS[p][k] = sum_o G0[p][o] W0c[o][k]  (k < pe: sine columns, k >= pe: cosine columns);  dphi[p][m] = DS[p][m] cos(phi) - DS[p][pe+m] sin(phi)
//   G[p] = sum_m dphi[p][m] B[:, m];  partial[d][blk] = sum_p (G0, G1, G0 r0, G0 r1, G1 r0, G1 r1),  r = coords_p - tc
// One block = `ppb` pixels of ONE target; register-blocked 4 pixel x 4 column GEMM like the prep kernel.
// ---------------------------------------------------------------------------------------------------------------
constexpr int AFF_THREADS = 448;

template <int HID>
__global__ void __launch_bounds__(AFF_THREADS) affine_grad_kernel(Args a) {
    constexpr int HP = (HID + 3) / 4 * 4;
    extern __shared__ __align__(16) float sm[];
    const Net& nt = a.net;
    const int pe = nt.pe, nb = 2 * pe, ppb = a.ppb, na = 2 * nt.ape;
    const int d = blockIdx.y, blk = blockIdx.x, t = threadIdx.x;
    const int p0 = blk * ppb, nq = min(ppb, a.tl.P - p0);
    float* Wt = sm;                 // [HP][nb]   W0c[o][k] (rows HID.. are zero)
    float* Gt = Wt + HP * nb;       // [HP][ppb]  G0 transposed (4 pixels per float4)
    float* DS = Gt + HP * ppb;      // [ppb][nb]
    __shared__ float cst[10];
    __shared__ float2 cxy[PREP_PPB_MAX];
    __shared__ float Bsm[2 * 64];
    __shared__ float red[6][AFF_THREADS / 32];
    if (t == 0) affine_consts(a.aff, d, cst);
    if (t >= 64 && t < 64 + nb) Bsm[t - 64] = a.B[t - 64];
    for (int i = t; i < HP * nb; i += AFF_THREADS) {
        int o, k;
        fd_divmod(i, a.fd_nb, &o, &k);
        Wt[i] = o < HID ? a.params[(size_t)o * nt.in_dim + 1 + na + k] : 0.f;
    }
    const float* gsrc = a.G0 + ((size_t)d * a.tl.P + p0) * HP;
    for (int i = t; i < ppb * HP; i += AFF_THREADS) {
        const int jj = i / HP, o = i % HP;
        Gt[o * ppb + jj] = jj < nq ? gsrc[i] : 0.f;
    }
    __syncthreads();
    if (t < nq) cxy[t] = transformed_coord(a, cst, p0 + t);
    // register-blocked items (4 pixels x 4 encoding features each).  A strided loop, not `if (t < items)`: 72 pixels x 100 features are 450 items
    // for 448 threads -- the two items left out fed uninitialised shared memory into the gradient of every block of 72 pixels (any match call
    // with D * P > 148 * 68 pixels, e.g. two targets at 100x100; found by tests/golden/match_step_100.npz)
    const int npg = ppb / 4, nkg = nb / 4;
    for (int item = t; item < npg * nkg; item += AFF_THREADS) {
        const int pg = item / nkg, kg = item % nkg;
        float4 acc[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) acc[x] = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4* gp = reinterpret_cast<const float4*>(Gt) + pg;
        const float4* wp = reinterpret_cast<const float4*>(Wt) + kg;
#pragma unroll 4
        for (int o = 0; o < HID; ++o) {
            const float4 g = gp[o * npg], w = wp[o * nkg];
            acc[0].x = fmaf(g.x, w.x, acc[0].x), acc[0].y = fmaf(g.x, w.y, acc[0].y), acc[0].z = fmaf(g.x, w.z, acc[0].z), acc[0].w = fmaf(g.x, w.w, acc[0].w);
            acc[1].x = fmaf(g.y, w.x, acc[1].x), acc[1].y = fmaf(g.y, w.y, acc[1].y), acc[1].z = fmaf(g.y, w.z, acc[1].z), acc[1].w = fmaf(g.y, w.w, acc[1].w);
            acc[2].x = fmaf(g.z, w.x, acc[2].x), acc[2].y = fmaf(g.z, w.y, acc[2].y), acc[2].z = fmaf(g.z, w.z, acc[2].z), acc[2].w = fmaf(g.z, w.w, acc[2].w);
            acc[3].x = fmaf(g.w, w.x, acc[3].x), acc[3].y = fmaf(g.w, w.y, acc[3].y), acc[3].z = fmaf(g.w, w.z, acc[3].z), acc[3].w = fmaf(g.w, w.w, acc[3].w);
        }
#pragma unroll
        for (int x = 0; x < 4; ++x) reinterpret_cast<float4*>(DS + (size_t)(4 * pg + x) * nb)[kg] = acc[x];
    }
    __syncthreads();
    // dphi and its projection on B, pixel by pixel: 8 lanes per pixel, fixed-order combine
    float v6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    {
        const int sub = t & 7;
        for (int jj = t >> 3; jj < nq; jj += AFF_THREADS / 8) {
            const float2 c = cxy[jj];
            float g0 = 0.f, g1 = 0.f;
            for (int m = sub; m < pe; m += 8) {
                float sn, cs;
                sincosf(c.x * Bsm[m] + c.y * Bsm[pe + m], &sn, &cs);
                const float dphi = DS[jj * nb + m] * cs - DS[jj * nb + pe + m] * sn;
                g0 = fmaf(dphi, Bsm[m], g0);
                g1 = fmaf(dphi, Bsm[pe + m], g1);
            }
            const unsigned gmask = 0xffu << ((t & 31) & ~7);
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) g0 += __shfl_xor_sync(gmask, g0, o), g1 += __shfl_xor_sync(gmask, g1, o);
            if (sub == 0) {
                const int p = p0 + jj;
                const float r0 = a.coords[2 * p] - cst[6], r1 = a.coords[2 * p + 1] - cst[7];
                v6[0] += g0, v6[1] += g1, v6[2] += g0 * r0, v6[3] += g0 * r1, v6[4] += g1 * r0, v6[5] += g1 * r1;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        const float w = warp_sum(v6[q]);
        if ((t & 31) == 0) red[q][t >> 5] = w;
    }
    __syncthreads();
    if (t < 6) {
        float sum = 0.f;
        for (int w = 0; w < AFF_THREADS / 32; ++w) sum += red[t][w];
        a.affpart[((size_t)d * gridDim.x + blk) * 6 + t] = sum;
    }
}

template <int HID>
static int launch_affine_grad(const Args& a, int* nblk_out, cudaStream_t st) {
    constexpr int HP = (HID + 3) / 4 * 4;
    const int nb = 2 * a.net.pe;
    const int nblk = (a.tl.P + a.ppb - 1) / a.ppb;
    const size_t smem = ((size_t)HP * nb + (size_t)HP * a.ppb + (size_t)a.ppb * nb) * sizeof(float);
    static size_t cache = 0;
    if (smem > 40 * 1024 && smem > cache) {
        LMR_CUDA_OK(cudaFuncSetAttribute(affine_grad_kernel<HID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cache = smem;
    }
    affine_grad_kernel<HID><<<dim3(nblk, a.tl.D), AFF_THREADS, smem, st>>>(a);
    LMR_LAUNCH_OK();
    *nblk_out = nblk;
    return 0;
}

// match: per-target reduction of the tile partials and chain rule to the 7 affine scalars
static __global__ void finish_affine_kernel(Args a, int npart, float* g_angles, float* g_scalings, float* g_shears, float* g_translation) {
    const int d = blockIdx.x;
    __shared__ float scratch[40];
    float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int i = threadIdx.x; i < npart; i += blockDim.x) {
        const float* src = a.affpart + ((size_t)d * npart + i) * 6;
#pragma unroll
        for (int q = 0; q < 6; ++q) v[q] += src[q];
    }
#pragma unroll
    for (int q = 0; q < 6; ++q) v[q] = block_sum(v[q], scratch);
    if (threadIdx.x == 0) {
        const lmr_affine& af = a.aff;
        float sn, cs;
        sincosf(af.angles[d], &sn, &cs);
        const float s0 = af.scalings[d * af.scal_cols], s1 = af.scalings[d * af.scal_cols + (af.scal_cols > 1 ? 1 : 0)];
        const float sh0 = af.shears[2 * d], sh1 = af.shears[2 * d + 1];
        // dM = [[v2, v3],[v4, v5]] (dL/dM_ab = sum_p G_a rel_b), dT = (v0, v1)
        const float dM00 = v[2], dM01 = v[3], dM10 = v[4], dM11 = v[5];
        // Hs = [[1,sh0],[sh1,1]];  S.Hs = [[s0, s0 sh0],[s1 sh1, s1]]
        const float sh00 = s0, sh01 = s0 * sh0, sh10 = s1 * sh1, sh11 = s1;
        // dL/dtheta = <dM, R'(theta) S Hs>, R' = [[-sn,-cs],[cs,-sn]]
        const float dth = dM00 * (-sn * sh00 - cs * sh10) + dM01 * (-sn * sh01 - cs * sh11) + dM10 * (cs * sh00 - sn * sh10) +
                          dM11 * (cs * sh01 - sn * sh11);
        // dL/ds0 = <dM, R E00 Hs> = <dM, [[cs, cs sh0],[sn, sn sh0]]>;  dL/ds1 = <dM, R E11 Hs> = <dM, [[-sn sh1, -sn],[cs sh1, cs]]>
        const float ds0 = dM00 * cs + dM01 * cs * sh0 + dM10 * sn + dM11 * sn * sh0;
        const float ds1 = dM00 * (-sn * sh1) + dM01 * (-sn) + dM10 * (cs * sh1) + dM11 * cs;
        // Q = (R S)^T dM;  dsh0 = Q[0][1], dsh1 = Q[1][0];  RS = [[cs s0, -sn s1],[sn s0, cs s1]]
        const float dsh0 = cs * s0 * dM01 + sn * s0 * dM11;
        const float dsh1 = -sn * s1 * dM00 + cs * s1 * dM10;
        g_angles[d] = dth;
        if (af.scal_cols > 1) {
            g_scalings[2 * d] = ds0, g_scalings[2 * d + 1] = ds1;
        } else {
            g_scalings[d] = ds0 + ds1;
        }
        g_shears[2 * d] = dsh0, g_shears[2 * d + 1] = dsh1;
        g_translation[2 * d] = v[0], g_translation[2 * d + 1] = v[1];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
struct WsLayout {
    size_t alpha, Avec, W0cT, wtimg, Cpg, betaG, hsave, G0, partials, l0part, affpart, dAred, total;
    int gridMain, gridL0, ppb;
};

static inline WsLayout ws_layout(const Net& nt, const Tiling& tl, bool bwd, bool want_cp = false, bool keep_act = false) {
    const int HP = (nt.hid + 3) / 4 * 4;
    WsLayout w;
    size_t off = 0;
    auto take = [&](size_t nfloats) {
        size_t r = off;
        off += align_up(nfloats * sizeof(float), 256);
        return r;
    };
    w.alpha = take((size_t)tl.N * 2 * nt.ape);
    w.Avec = take((size_t)tl.N * HP);
    w.W0cT = take((size_t)2 * nt.pe * HP);
    w.wtimg = want_cp ? take((size_t)(nt.nl - 1) * 16384 / 4) : 0;
    w.Cpg = want_cp ? take((size_t)tl.D * tl.P * HP) : 0;
    w.betaG = (want_cp && tl.D == 1) ? take((size_t)tl.P * 2 * nt.pe) : 0;  // identical prefix in the forward and backward layouts
    w.hsave = (want_cp && keep_act && !bwd) ? take((size_t)tl.numTiles * keep_record_bytes(nt.nl) / 4) : 0;  // forward layout only (a suffix of it)
    w.gridMain = tl.numTiles < num_sms() ? tl.numTiles : num_sms();
    const int total = tl.D * tl.P;
    static const int l0_bps = getenv("LMR_L0_BPS") && atoi(getenv("LMR_L0_BPS")) > 0 ? atoi(getenv("LMR_L0_BPS")) : 1;  // l0 / affine-gradient blocks per SM (tuning knob)
    w.ppb = choose_ppb(total, l0_bps);
    const int nchunks = (total + w.ppb - 1) / w.ppb;
    w.gridL0 = nchunks < l0_bps * num_sms() ? nchunks : l0_bps * num_sms();
    w.G0 = w.partials = w.l0part = w.affpart = w.dAred = 0;
    if (bwd) {
        w.G0 = take((size_t)tl.D * tl.P * HP);
        w.partials = take((size_t)2 * w.gridMain * partial_size(nt, tl, HP));  // the two-stream tensor-core backward writes one partial per stream
        w.l0part = take((size_t)w.gridL0 * nt.hid * 2 * nt.pe);
        w.affpart = take((size_t)tl.D * tl.tilesP * 6);
        w.dAred = take((size_t)tl.N * HP);
    }
    w.total = off;
    return w;
}

static inline int make_net(const lmr_inr_desc* d, Net* nt) {
    LMR_CHECK_ARG(d != nullptr, "inr: null descriptor");
    LMR_CHECK_ARG(d->hidden == 50, "inr: hidden width %d not supported by the fused kernels (supported: 50)", d->hidden);
    LMR_CHECK_ARG(d->n_hidden >= 2 && d->n_hidden <= 8, "inr: n_hidden %d out of range [2,8]", d->n_hidden);
    LMR_CHECK_ARG(d->channels >= 1 && d->channels <= 4, "inr: channels %d out of range [1,4]", d->channels);
    LMR_CHECK_ARG(d->pe_dim >= 1 && d->pe_dim % 2 == 0 && d->pe_dim <= 128, "inr: pe_dim %d must be even and <= 128", d->pe_dim);
    LMR_CHECK_ARG(d->aux_pe_dim >= 1 && d->aux_pe_dim <= 256, "inr: aux_pe_dim %d out of range", d->aux_pe_dim);
    nt->hid = d->hidden, nt->nl = d->n_hidden, nt->pe = d->pe_dim, nt->ape = d->aux_pe_dim, nt->C = d->channels;
    nt->in_dim = 1 + 2 * d->aux_pe_dim + 2 * d->pe_dim;
    return 0;
}

static inline int fill_args(Args& a, const Net& nt, const Tiling& tl, const float* params, const float* B, const float* auxB, const float* grid,
                     const float* coords, const float* coord_shift, const lmr_affine* affine, void* ws, const WsLayout& wl) {
    a.net = nt, a.tl = tl;
    a.params = params, a.B = B, a.auxB = auxB, a.grid = grid, a.coords = coords, a.coord_shift = coord_shift;
    a.has_aff = affine != nullptr && affine->angles != nullptr;
    if (a.has_aff) {
        a.aff = *affine;
        LMR_CHECK_ARG(affine->scalings && affine->shears && affine->translation, "inr: incomplete affine transform");
        LMR_CHECK_ARG(affine->scal_cols == 1 || affine->scal_cols == 2, "inr: scal_cols must be 1 or 2");
    } else {
        LMR_CHECK_ARG(tl.D == 1, "inr: D must be 1 without an affine transform");
        a.aff = lmr_affine{};
    }
    char* base = static_cast<char*>(ws);
    a.alpha = reinterpret_cast<float*>(base + wl.alpha);
    a.Avec = reinterpret_cast<float*>(base + wl.Avec);
    a.W0cT = reinterpret_cast<float*>(base + wl.W0cT);
    a.wtimg = reinterpret_cast<unsigned char*>(base + wl.wtimg);
    a.Cpg = reinterpret_cast<float*>(base + wl.Cpg);
    a.betaG = wl.betaG ? reinterpret_cast<float*>(base + wl.betaG) : nullptr;
    a.want_cp = 0;
    a.reuse_beta = 0;
    a.ppb = wl.ppb;
    static const int prep_bps = getenv("LMR_PREP_BPS") ? atoi(getenv("LMR_PREP_BPS")) : 2;  // prep blocks per SM (tuning knob; measured: 1 -> 22 us, 2 -> 16 us at 100x100)
    a.ppb_prep = choose_ppb(tl.D * tl.P, prep_bps > 0 ? prep_bps : 2);
    a.fd_nb = make_fastdiv(2 * nt.pe), a.fd_ppb_prep = make_fastdiv(a.ppb_prep);
    a.dAred = reinterpret_cast<float*>(base + wl.dAred);
    a.G0 = reinterpret_cast<float*>(base + wl.G0);
    a.partials = reinterpret_cast<float*>(base + wl.partials);
    a.affpart = reinterpret_cast<float*>(base + wl.affpart);
    a.img = nullptr, a.g_img = nullptr, a.want_weights = 0, a.want_affine = 0;
    a.trace = nullptr;
    a.stagger = 0;
    a.hsave = wl.hsave ? reinterpret_cast<unsigned char*>(base + wl.hsave) : nullptr;
    return 0;
}

template <int HID>
static int launch_prep(const Args& a, cudaStream_t st) {
    constexpr int HP = (HID + 3) / 4 * 4;
    const int nb = 2 * a.net.pe;
    const int cp_blocks = a.want_cp ? (a.tl.D * a.tl.P + a.ppb_prep - 1) / a.ppb_prep : 0;
    // the coordinate-part GEMM of a block runs one (4 pixel x 4 output) item per thread; the B table is staged by threads 64 .. 64 + 2 pe
    LMR_CHECK_ARG(!a.want_cp || ((a.ppb_prep / 4) * (HP / 4) <= 256 && 64 + nb <= 256 && nb <= 128), "inr: pe_dim %d not supported by the preparation kernel", a.net.pe);
    size_t smem = (size_t)2 * a.net.ape * sizeof(float);
    if (a.want_cp) smem = ((size_t)nb * HP + (size_t)nb * a.ppb_prep) * sizeof(float);
    static size_t cache = 0;
    if (smem > 40 * 1024 && smem > cache) {  // static shared memory counts against the 48 KB default limit too
        LMR_CUDA_OK(cudaFuncSetAttribute(prep_all_kernel<HID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cache = smem;
    }
    const int wt_blocks = a.want_cp ? ((a.net.nl - 1) * 4096 + 1023) / 1024 : 0;
    prep_all_kernel<HID><<<a.tl.N + 1 + cp_blocks + wt_blocks, 256, smem, st>>>(a);
    LMR_LAUNCH_OK();
    return 0;
}

template <int HID>
static int launch_l0_partial(const Args& a, float* l0part, int grid_l0, cudaStream_t st) {
    constexpr int HP = (HID + 3) / 4 * 4;
    const size_t smem = (size_t)a.ppb * (HP + 2 * a.net.pe) * sizeof(float);
    // one register-blocked item (4 outputs x 4 encoding features) per thread, accumulated across the block's pixel chunks
    LMR_CHECK_ARG((HP / 4) * (2 * a.net.pe / 4) <= L0_THREADS && a.net.pe % 2 == 0,
                  "inr_backward: pe_dim %d needs %d layer-0 weight-gradient items (> %d threads)", a.net.pe, (HP / 4) * (2 * a.net.pe / 4), L0_THREADS);
    static size_t cache = 0;
    if (smem > 40 * 1024 && smem > cache) {
        LMR_CUDA_OK(cudaFuncSetAttribute(l0_partial_kernel<HID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cache = smem;
    }
    l0_partial_kernel<HID><<<grid_l0, L0_THREADS, smem, st>>>(a, l0part);
    LMR_LAUNCH_OK();
    return 0;
}

// the two launches that turn the per-CTA partials into g_params
template <int HID>
static int launch_finish_weights(const Args& a, const float* l0part, int nPartMain, int nPartL0, float* g_params, cudaStream_t st) {
    constexpr int HP = (HID + 3) / 4 * 4;
    const size_t n1 = a.net.count() + (size_t)a.tl.N * HP;
    reduce_partials_kernel<HID><<<(int)((4 * n1 + 127) / 128), 128, 0, st>>>(a, l0part, nPartMain, nPartL0, g_params);
    LMR_LAUNCH_OK();
    const int n2 = HID * (1 + 2 * a.net.ape) + HID;
    finish_w0a_kernel<HID><<<(n2 + 127) / 128, 128, 0, st>>>(a, g_params);
    LMR_LAUNCH_OK();
    return 0;
}

}  // namespace inr
}  // namespace lmr
