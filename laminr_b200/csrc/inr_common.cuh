// Shared pieces of the INR kernels (fp32 CUDA-core path in inr_simt.cu, tcgen05 path in inr_tc.cu): network / tiling
// descriptors, the prep kernels (latent encoding, separable layer-0 pieces), the affine coordinate transform, the layer-0
// weight-gradient kernel and the deterministic final reductions.  Global workspace arrays use a row stride of SHP = 52.
#pragma once
#include <math.h>

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

static inline Tiling make_tiling(int N, int P, int D, int tp_cap) {
    Tiling t;
    t.N = N, t.P = P, t.D = D;
    t.NC = N < TR_MAX ? N : TR_MAX;
    t.TP = TR_MAX / t.NC;
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
    // forward
    float* img;  // [D,N,C,P]
    // backward
    const float* g_img;
    float* G0;        // [D*P, HP]   (learn: sum_n delta0)
    float* partials;  // [grid, partial_size]
    float* affpart;   // [D, tilesP, 6]  (match)
    int want_weights, want_affine;
};

// ---------------------------------------------------------------------------------------------------------------
// prep: latent encoding (inr.py:371-386,401-409), A[n] = b0 + W0[:,1:1+2a] alpha_n, packed transpose of W0's coord block
// ---------------------------------------------------------------------------------------------------------------
template <int HID>
__global__ void prep_kernel(Args a) {
    constexpr int HP = (HID + 3) / 4 * 4;
    const Net& nt = a.net;
    const int N = a.tl.N, na = 2 * nt.ape, nb = 2 * nt.pe;
    const float* W0 = a.params;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N * nt.ape; i += gridDim.x * blockDim.x) {
        const int n = i / nt.ape, m = i % nt.ape;
        const float g = a.grid[n];
        const float phi = sinf(g) * a.auxB[m] + cosf(g) * a.auxB[nt.ape + m];
        a.alpha[n * na + m] = sinf(phi);
        a.alpha[n * na + nt.ape + m] = cosf(phi);
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nb * HP; i += gridDim.x * blockDim.x) {
        const int k = i / HP, o = i % HP;
        a.W0cT[i] = o < HID ? W0[(size_t)o * nt.in_dim + 1 + na + k] : 0.f;
    }
}

template <int HID>
__global__ void prep_A_kernel(Args a) {
    constexpr int HP = (HID + 3) / 4 * 4;
    const Net& nt = a.net;
    const int N = a.tl.N, na = 2 * nt.ape;
    const float* W0 = a.params;
    const float* b0 = a.params + nt.offb(0);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N * HP; i += gridDim.x * blockDim.x) {
        const int n = i / HP, o = i % HP;
        float acc = 0.f;
        if (o < HID) {
            acc = b0[o];
            const float* w = W0 + (size_t)o * nt.in_dim + 1;
            const float* al = a.alpha + (size_t)n * na;
            for (int k = 0; k < na; ++k) acc = fmaf(w[k], al[k], acc);
        } else if (o == HID) {
            acc = 20.f;  // tanh(20) == 1 in fp32: the tensor-core path keeps a constant-1 feature (bias column) in slot HID
        }
        a.Avec[i] = acc;
    }
}

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

struct TileIdx {
    int d, p0, n0, np, nn;  // target, first pixel, first latent, #pixels, #latents
};

__device__ inline TileIdx decode_tile(const Tiling& tl, int tile) {
    TileIdx t;
    const int chunk = tile % tl.nChunks;
    const int rest = tile / tl.nChunks;
    const int pb = rest % tl.tilesP;
    t.d = rest / tl.tilesP;
    t.p0 = pb * tl.TP;
    t.n0 = chunk * tl.NC;
    t.np = min(tl.TP, tl.P - t.p0);
    t.nn = min(tl.NC, tl.N - t.n0);
    return t;
}

// ---------------------------------------------------------------------------------------------------------------
// learn: layer-0 coordinate block  dW0c[o][k] = sum_{d,p} G0[d,p][o] beta[d,p][k]   (split over pixel chunks)
// ---------------------------------------------------------------------------------------------------------------
constexpr int L0_CHUNK = 64;

template <int HID>
__global__ void __launch_bounds__(256) l0_partial_kernel(Args a, float* l0part /*[grid, HID*2pe]*/) {
    constexpr int HP = (HID + 3) / 4 * 4;
    extern __shared__ __align__(16) float sm[];
    const Net& nt = a.net;
    const int pe = nt.pe, nb = 2 * pe;
    float* Gs = sm;                      // [L0_CHUNK][HP]
    float* Bs = Gs + L0_CHUNK * HP;      // [L0_CHUNK][nb]
    const int total = a.tl.D * a.tl.P;   // learn: D == 1 and no affine, but keep it general (coords of target d)
    const int nchunks = (total + L0_CHUNK - 1) / L0_CHUNK;
    // entries: o-block of 5 x k-block of 4
    const int nOB = HID / 5, nKB = nb / 4;
    const int e = threadIdx.x;
    const bool owner = e < nOB * nKB;
    const int ob = e % nOB, kb = e / nOB;
    float acc[5][4];
#pragma unroll
    for (int x = 0; x < 5; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = 0.f;
    __shared__ float affc[10];
    for (int ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        const int q0 = ch * L0_CHUNK, nq = min(L0_CHUNK, total - q0);
        __syncthreads();
        for (int i = threadIdx.x; i < nq * HP; i += blockDim.x) Gs[i] = a.G0[(size_t)q0 * HP + i];
        for (int i = threadIdx.x; i < nq * pe; i += blockDim.x) {
            const int jj = i / pe, m = i % pe;
            const int q = q0 + jj, p = q % a.tl.P;
            float2 c;
            if (a.has_aff) {
                float cst[10];
                affine_consts(a.aff, q / a.tl.P, cst);
                c = transformed_coord(a, cst, p);
            } else {
                c = transformed_coord(a, affc, p);
            }
            float sn, cs;
            sincosf(c.x * a.B[m] + c.y * a.B[pe + m], &sn, &cs);
            Bs[jj * nb + m] = sn, Bs[jj * nb + pe + m] = cs;
        }
        __syncthreads();
        if (owner) {
            for (int jj = 0; jj < nq; ++jj) {
                float g[5];
#pragma unroll
                for (int x = 0; x < 5; ++x) g[x] = Gs[jj * HP + ob * 5 + x];
                const float4 b = *reinterpret_cast<const float4*>(Bs + jj * nb + kb * 4);
#pragma unroll
                for (int x = 0; x < 5; ++x) {
                    acc[x][0] = fmaf(g[x], b.x, acc[x][0]);
                    acc[x][1] = fmaf(g[x], b.y, acc[x][1]);
                    acc[x][2] = fmaf(g[x], b.z, acc[x][2]);
                    acc[x][3] = fmaf(g[x], b.w, acc[x][3]);
                }
            }
        }
    }
    if (owner) {
        float* dst = l0part + (size_t)blockIdx.x * HID * nb;
#pragma unroll
        for (int x = 0; x < 5; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) dst[(ob * 5 + x) * nb + kb * 4 + y] = acc[x][y];
    }
}

// learn: deterministic reduction of every partial into g_params (torch layout)
template <int HID>
__global__ void finish_weights_kernel(Args a, const float* l0part, int nPartMain, int nPartL0, float* g_params) {
    constexpr int HP = (HID + 3) / 4 * 4;
    const Net& nt = a.net;
    const Tiling& tl = a.tl;
    const size_t psz = partial_size(nt, tl, HP);
    const int L = nt.nl, na = 2 * nt.ape, nb = 2 * nt.pe;
    const size_t total = nt.count();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        float sum = 0.f;
        if (i < nt.size0()) {
            if (i < (size_t)nt.hid * nt.in_dim) {
                const int o = i / nt.in_dim, col = i % nt.in_dim;
                if (col == 0) {
                    sum = 0.f;  // template input is the constant 0 (inr.py:200-205)
                } else if (col < 1 + na) {
                    const int k = col - 1;
                    const size_t offA = (size_t)(L - 1) * HID * HID + (size_t)(L - 1) * HP + (size_t)nt.C * HP + 4;
                    for (int n = 0; n < tl.N; ++n) {
                        float dA = 0.f;
                        for (int b = 0; b < nPartMain; ++b) dA += a.partials[b * psz + offA + (size_t)n * HP + o];
                        sum = fmaf(dA, a.alpha[(size_t)n * na + k], sum);
                    }
                } else {
                    const int k = col - 1 - na;
                    for (int b = 0; b < nPartL0; ++b) sum += l0part[(size_t)b * HID * nb + o * nb + k];
                }
            } else {
                const int o = i - (size_t)nt.hid * nt.in_dim;
                const size_t offA = (size_t)(L - 1) * HID * HID + (size_t)(L - 1) * HP + (size_t)nt.C * HP + 4;
                for (int n = 0; n < tl.N; ++n)
                    for (int b = 0; b < nPartMain; ++b) sum += a.partials[b * psz + offA + (size_t)n * HP + o];
            }
        } else if (i < nt.offWout()) {
            const size_t r = i - nt.size0();
            const int l = r / nt.sizeh();  // hidden layer l+1
            const size_t w = r % nt.sizeh();
            size_t off;
            if (w < (size_t)HID * HID) off = (size_t)l * HID * HID + w;
            else off = (size_t)(L - 1) * HID * HID + (size_t)l * HP + (w - (size_t)HID * HID);
            for (int b = 0; b < nPartMain; ++b) sum += a.partials[b * psz + off];
        } else {
            const size_t r = i - nt.offWout();
            const size_t base = (size_t)(L - 1) * HID * HID + (size_t)(L - 1) * HP;
            size_t off;
            if (r < (size_t)nt.C * HID) off = base + (r / HID) * HP + (r % HID);
            else off = base + (size_t)nt.C * HP + (r - (size_t)nt.C * HID);
            for (int b = 0; b < nPartMain; ++b) sum += a.partials[b * psz + off];
        }
        g_params[i] = sum;
    }
}

// match: per-target reduction of the tile partials and chain rule to the 7 affine scalars
static __global__ void finish_affine_kernel(Args a, float* g_angles, float* g_scalings, float* g_shears, float* g_translation) {
    const int d = blockIdx.x;
    __shared__ float scratch[40];
    float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int i = threadIdx.x; i < a.tl.tilesP; i += blockDim.x) {
        const float* src = a.affpart + ((size_t)d * a.tl.tilesP + i) * 6;
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
    size_t alpha, Avec, W0cT, G0, partials, l0part, affpart, total;
    int gridMain, gridL0;
};

static inline WsLayout ws_layout(const Net& nt, const Tiling& tl, bool bwd) {
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
    w.gridMain = tl.numTiles < num_sms() ? tl.numTiles : num_sms();
    const int total = tl.D * tl.P;
    const int nchunks = (total + L0_CHUNK - 1) / L0_CHUNK;
    w.gridL0 = nchunks < num_sms() ? nchunks : num_sms();
    w.G0 = w.partials = w.l0part = w.affpart = 0;
    if (bwd) {
        w.G0 = take((size_t)tl.D * tl.P * HP);
        w.partials = take((size_t)w.gridMain * partial_size(nt, tl, HP));
        w.l0part = take((size_t)w.gridL0 * nt.hid * 2 * nt.pe);
        w.affpart = take((size_t)tl.D * tl.tilesP * 6);
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
    a.G0 = reinterpret_cast<float*>(base + wl.G0);
    a.partials = reinterpret_cast<float*>(base + wl.partials);
    a.affpart = reinterpret_cast<float*>(base + wl.affpart);
    a.img = nullptr, a.g_img = nullptr, a.want_weights = 0, a.want_affine = 0;
    return 0;
}

template <int HID>
static int launch_prep(const Args& a, cudaStream_t st) {
    prep_kernel<HID><<<4, 256, 0, st>>>(a);
    LMR_LAUNCH_OK();
    prep_A_kernel<HID><<<(a.tl.N * ((HID + 3) / 4 * 4) + 255) / 256, 256, 0, st>>>(a);
    LMR_LAUNCH_OK();
    return 0;
}


}  // namespace inr
}  // namespace lmr
