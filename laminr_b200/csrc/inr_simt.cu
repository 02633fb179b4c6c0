// INR generator, fp32 CUDA-core path (LMR_PREC_FP32): forward and recompute-backward.
//
// Replaces INRTemplates.forward / autograd through it (reference laminr/inr.py:53-100,317-413; the affine
// transform of laminr/coordinate_transforms.py:226-309).  See DESIGN.md "INR kernels" for the tiling.
//
// Row = one MLP evaluation (target d, latent n, pixel p).  A tile is TP pixels x NC latents of one target
// (<= 128 rows, thread t <-> row t, t = nl*TP + j).  Layer 0 is separable:
//     z0[n,d,p] = A[n] + Cp[d,p],   A[n] = W0a alpha_n + b0 (prep kernel),   Cp[d,p] = W0c beta_dp (per tile)
// so the [rows, 201] input of the reference (inr.py:436-452) never exists.  Activations live in shared memory
// as k-major panels Hs[l][k][t] (conflict-free for thread<->row access); nothing is saved between forward and
// backward: the backward recomputes the forward per tile and back-propagates in the tile.
#include <math.h>

#include "common.cuh"

namespace lmr {
namespace simt {

constexpr int TR_MAX = 128;  // rows (= threads) per tile
constexpr int STR = 132;     // row stride (floats) of the k-major activation panels
constexpr int TP_CAP = 32;   // pixels per tile cap

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

static Tiling make_tiling(int N, int P, int D) {
    Tiling t;
    t.N = N, t.P = P, t.D = D;
    t.NC = N < TR_MAX ? N : TR_MAX;
    t.TP = TR_MAX / t.NC;
    if (t.TP > TP_CAP) t.TP = TP_CAP;
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
        }
        a.Avec[i] = acc;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// shared-memory carve-up
// ---------------------------------------------------------------------------------------------------------------
template <int HID>
struct Smem {
    static constexpr int HP = (HID + 3) / 4 * 4;
    float *Wt, *bh, *Wout, *bout, *As, *Cp, *beta, *affc, *Hs, *Ds, *Dy, *dWs, *dbs, *dWouts, *dAs, *G0s, *red;
    size_t floats;
    __host__ __device__ Smem(float* base, const Net& nt, const Tiling& tl, bool bwd, bool want_w) {
        float* p = base;
        auto take = [&](size_t n) {
            float* r = p;
            p += (n + 3) / 4 * 4;
            return r;
        };
        Wt = take((size_t)(nt.nl - 1) * HID * HP);
        bh = take((size_t)(nt.nl - 1) * HP);
        Wout = take((size_t)nt.C * HP);
        bout = take(4);
        As = take((size_t)tl.NC * HP);
        Cp = take((size_t)tl.TP * HP);
        beta = take((size_t)tl.TP * 2 * nt.pe);
        affc = take(12);
        Hs = take((size_t)(bwd ? nt.nl : 1) * HID * STR);
        Ds = Dy = dWs = dbs = dWouts = dAs = G0s = red = nullptr;
        if (bwd) {
            Ds = take((size_t)HID * STR);
            Dy = take((size_t)4 * STR);
            G0s = take((size_t)tl.TP * HP);
            red = take(64);
            if (want_w) {
                dWs = take((size_t)(nt.nl - 1) * HID * HID);
                dbs = take((size_t)(nt.nl - 1) * HP);
                dWouts = take((size_t)nt.C * HP + 4);
                dAs = take((size_t)tl.NC * HP);
            }
        }
        floats = p - base;
    }
};

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

// Load the per-launch constants (hidden weights transposed, biases, output layer) into shared memory.
template <int HID>
__device__ inline void load_weights(const Args& a, Smem<HID>& s) {
    constexpr int HP = Smem<HID>::HP;
    const Net& nt = a.net;
    for (int l = 1; l < nt.nl; ++l) {
        const float* W = a.params + nt.offW(l);
        float* dst = s.Wt + (size_t)(l - 1) * HID * HP;
        for (int i = threadIdx.x; i < HID * HP; i += blockDim.x) {
            const int k = i / HP, o = i % HP;
            dst[i] = o < HID ? W[o * HID + k] : 0.f;
        }
        const float* b = a.params + nt.offb(l);
        for (int o = threadIdx.x; o < HP; o += blockDim.x) s.bh[(l - 1) * HP + o] = o < HID ? b[o] : 0.f;
    }
    const float* Wo = a.params + nt.offWout();
    for (int i = threadIdx.x; i < nt.C * HP; i += blockDim.x) {
        const int c = i / HP, k = i % HP;
        s.Wout[i] = k < HID ? Wo[c * HID + k] : 0.f;
    }
    if (threadIdx.x < 4) s.bout[threadIdx.x] = threadIdx.x < nt.C ? a.params[nt.offbout() + threadIdx.x] : 0.f;
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

// Phase 0 of a tile: positional encoding beta[j][2pe] and Cp[j][o] = sum_k beta[j][k] W0c[o][k].
template <int HID>
__device__ inline void tile_prologue(const Args& a, Smem<HID>& s, const TileIdx& ti, bool load_A) {
    constexpr int HP = Smem<HID>::HP;
    const Net& nt = a.net;
    const int pe = nt.pe;
    if (a.has_aff && threadIdx.x == 0) affine_consts(a.aff, ti.d, s.affc);
    __syncthreads();
    for (int i = threadIdx.x; i < ti.np * pe; i += blockDim.x) {
        const int j = i / pe, m = i % pe;
        const float2 c = transformed_coord(a, s.affc, ti.p0 + j);
        const float phi = c.x * a.B[m] + c.y * a.B[pe + m];
        float sn, cs;
        sincosf(phi, &sn, &cs);
        s.beta[j * 2 * pe + m] = sn;
        s.beta[j * 2 * pe + pe + m] = cs;
    }
    if (load_A) {
        for (int i = threadIdx.x; i < ti.nn * HP; i += blockDim.x) s.As[i] = a.Avec[(size_t)ti.n0 * HP + i];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ti.np * (HP / 4); i += blockDim.x) {
        const int j = i / (HP / 4), o4 = i % (HP / 4);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const float* bj = s.beta + j * 2 * pe;
        const float4* w = reinterpret_cast<const float4*>(a.W0cT) + o4;
        for (int k = 0; k < 2 * pe; ++k) {
            const float bk = bj[k];
            const float4 ww = __ldg(w + (size_t)k * (HP / 4));
            acc.x = fmaf(bk, ww.x, acc.x), acc.y = fmaf(bk, ww.y, acc.y), acc.z = fmaf(bk, ww.z, acc.z), acc.w = fmaf(bk, ww.w, acc.w);
        }
        reinterpret_cast<float4*>(s.Cp + j * HP)[o4] = acc;
    }
    __syncthreads();
}

// One hidden layer for the thread's row: out[o] = b[o] + sum_k Hin[k][t] Wt[k][o]
template <int HID>
__device__ __forceinline__ void dense_row(const float* __restrict__ Hin, const float* __restrict__ Wt, const float* __restrict__ b, int t,
                                          float (&acc)[Smem<HID>::HP]) {
    constexpr int HP = Smem<HID>::HP;
#pragma unroll
    for (int o = 0; o < HP; ++o) acc[o] = b[o];
#pragma unroll 2
    for (int k = 0; k < HID; ++k) {
        const float hk = Hin[k * STR + t];
        const float4* w = reinterpret_cast<const float4*>(Wt + k * HP);
#pragma unroll
        for (int o4 = 0; o4 < HP / 4; ++o4) {
            const float4 ww = w[o4];
            acc[4 * o4 + 0] = fmaf(hk, ww.x, acc[4 * o4 + 0]);
            acc[4 * o4 + 1] = fmaf(hk, ww.y, acc[4 * o4 + 1]);
            acc[4 * o4 + 2] = fmaf(hk, ww.z, acc[4 * o4 + 2]);
            acc[4 * o4 + 3] = fmaf(hk, ww.w, acc[4 * o4 + 3]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------
template <int HID>
__global__ void __launch_bounds__(TR_MAX) inr_fwd_kernel(Args a) {
    constexpr int HP = Smem<HID>::HP;
    extern __shared__ __align__(16) float smem_raw[];
    Smem<HID> s(smem_raw, a.net, a.tl, false, false);
    const Net& nt = a.net;
    const Tiling& tl = a.tl;
    const int t = threadIdx.x;
    load_weights<HID>(a, s);
    __syncthreads();
    for (int tile = blockIdx.x; tile < tl.numTiles; tile += gridDim.x) {
        const TileIdx ti = decode_tile(tl, tile);
        tile_prologue<HID>(a, s, ti, true);
        const int nl = t / tl.TP, j = t % tl.TP;
        if (nl < ti.nn && j < ti.np) {
            float acc[HP];
            float* H = s.Hs;
#pragma unroll
            for (int o = 0; o < HID; ++o) H[o * STR + t] = tanhf(s.As[nl * HP + o] + s.Cp[j * HP + o]);
            for (int l = 1; l < nt.nl; ++l) {
                dense_row<HID>(H, s.Wt + (size_t)(l - 1) * HID * HP, s.bh + (l - 1) * HP, t, acc);
#pragma unroll
                for (int o = 0; o < HID; ++o) H[o * STR + t] = tanhf(acc[o]);
            }
            const size_t base = (((size_t)ti.d * tl.N + (ti.n0 + nl)) * nt.C) * tl.P + ti.p0 + j;
            for (int c = 0; c < nt.C; ++c) {
                float y = s.bout[c];
                for (int k = 0; k < HID; ++k) y = fmaf(H[k * STR + t], s.Wout[c * HP + k], y);
                a.img[base + (size_t)c * tl.P] = tanhf(y);
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// backward (recompute + in-tile back-propagation)
// ---------------------------------------------------------------------------------------------------------------
template <int HID>
__global__ void __launch_bounds__(TR_MAX) inr_bwd_kernel(Args a) {
    constexpr int HP = Smem<HID>::HP;
    constexpr int OB = 5;               // dW register block: OB x OB entries per thread
    constexpr int NB_ = HID / OB;       // blocks per dim (HID % 5 == 0 for the instantiated widths)
    extern __shared__ __align__(16) float smem_raw[];
    const bool want_w = a.want_weights != 0;
    Smem<HID> s(smem_raw, a.net, a.tl, true, want_w);
    const Net& nt = a.net;
    const Tiling& tl = a.tl;
    const int t = threadIdx.x;
    const int L = nt.nl;
    load_weights<HID>(a, s);
    if (want_w) {
        // dWs, dbs, dWouts(+4), dAs are contiguous in the carve-up order -> zero them as one range
        for (size_t i = t; i < (size_t)(s.dAs + (size_t)tl.NC * HP - s.dWs); i += blockDim.x) s.dWs[i] = 0.f;
    }
    __syncthreads();
    const int e_ob = t % NB_, e_ib = t / NB_;  // dW entry block of this thread (valid for t < NB_*NB_)
    const bool dw_thread = t < NB_ * NB_;

    for (int tile = blockIdx.x; tile < tl.numTiles; tile += gridDim.x) {
        const TileIdx ti = decode_tile(tl, tile);
        tile_prologue<HID>(a, s, ti, true);
        const int nl = t / tl.TP, j = t % tl.TP;
        const bool active = nl < ti.nn && j < ti.np;
        const int TRu = ti.nn * tl.TP;  // rows in use are t < TRu with j < np; inactive rows hold zeros
        // ---- forward recompute, keep every activation panel
        {
            float acc[HP];
            if (active) {
#pragma unroll
                for (int o = 0; o < HID; ++o) s.Hs[o * STR + t] = tanhf(s.As[nl * HP + o] + s.Cp[j * HP + o]);
                for (int l = 1; l < L; ++l) {
                    dense_row<HID>(s.Hs + (size_t)(l - 1) * HID * STR, s.Wt + (size_t)(l - 1) * HID * HP, s.bh + (l - 1) * HP, t, acc);
                    float* Hn = s.Hs + (size_t)l * HID * STR;
#pragma unroll
                    for (int o = 0; o < HID; ++o) Hn[o * STR + t] = tanhf(acc[o]);
                }
                const float* Hl = s.Hs + (size_t)(L - 1) * HID * STR;
                const size_t base = (((size_t)ti.d * tl.N + (ti.n0 + nl)) * nt.C) * tl.P + ti.p0 + j;
                float dy[4];
                for (int c = 0; c < nt.C; ++c) {
                    float y = s.bout[c];
                    for (int k = 0; k < HID; ++k) y = fmaf(Hl[k * STR + t], s.Wout[c * HP + k], y);
                    y = tanhf(y);
                    dy[c] = a.g_img[base + (size_t)c * tl.P] * (1.f - y * y);
                    s.Dy[c * STR + t] = dy[c];
                }
                for (int k = 0; k < HID; ++k) {
                    float dh = 0.f;
                    for (int c = 0; c < nt.C; ++c) dh = fmaf(s.Wout[c * HP + k], dy[c], dh);
                    const float h = Hl[k * STR + t];
                    s.Ds[k * STR + t] = dh * (1.f - h * h);
                }
            } else {
                for (int l = 0; l < L; ++l)
                    for (int o = 0; o < HID; ++o) s.Hs[((size_t)l * HID + o) * STR + t] = 0.f;
                for (int o = 0; o < HID; ++o) s.Ds[o * STR + t] = 0.f;
                for (int c = 0; c < 4; ++c) s.Dy[c * STR + t] = 0.f;
            }
        }
        __syncthreads();
        // ---- layers L-1 .. 1: weight gradients (entry-parallel) and data gradient (row-parallel)
        for (int l = L - 1; l >= 1; --l) {
            const float* Hprev = s.Hs + (size_t)(l - 1) * HID * STR;
            if (want_w) {
                if (dw_thread) {
                    float w[OB][OB];
                    float* dst = s.dWs + (size_t)(l - 1) * HID * HID;
#pragma unroll
                    for (int x = 0; x < OB; ++x)
#pragma unroll
                        for (int y = 0; y < OB; ++y) w[x][y] = dst[(e_ob * OB + x) * HID + e_ib * OB + y];
                    for (int r = 0; r < TR_MAX; r += 4) {
                        float4 dv[OB], hv[OB];
#pragma unroll
                        for (int x = 0; x < OB; ++x) dv[x] = *reinterpret_cast<const float4*>(s.Ds + (e_ob * OB + x) * STR + r);
#pragma unroll
                        for (int y = 0; y < OB; ++y) hv[y] = *reinterpret_cast<const float4*>(Hprev + (e_ib * OB + y) * STR + r);
#pragma unroll
                        for (int x = 0; x < OB; ++x)
#pragma unroll
                            for (int y = 0; y < OB; ++y) {
                                w[x][y] = fmaf(dv[x].x, hv[y].x, w[x][y]);
                                w[x][y] = fmaf(dv[x].y, hv[y].y, w[x][y]);
                                w[x][y] = fmaf(dv[x].z, hv[y].z, w[x][y]);
                                w[x][y] = fmaf(dv[x].w, hv[y].w, w[x][y]);
                            }
                    }
#pragma unroll
                    for (int x = 0; x < OB; ++x)
#pragma unroll
                        for (int y = 0; y < OB; ++y) dst[(e_ob * OB + x) * HID + e_ib * OB + y] = w[x][y];
                } else {
                    // the spare threads reduce the bias gradients (and the output layer's at l == L-1)
                    const int q = t - NB_ * NB_, nq = TR_MAX - NB_ * NB_;
                    for (int o = q; o < HID; o += nq) {
                        float sum = 0.f;
                        for (int r = 0; r < TR_MAX; ++r) sum += s.Ds[o * STR + r];
                        s.dbs[(l - 1) * HP + o] += sum;
                    }
                    if (l == L - 1) {
                        for (int i = q; i < nt.C * HID; i += nq) {
                            const int c = i / HID, k = i % HID;
                            float sum = 0.f;
                            for (int r = 0; r < TR_MAX; ++r) sum = fmaf(s.Dy[c * STR + r], Hprev[(size_t)HID * STR + k * STR + r], sum);
                            s.dWouts[c * HP + k] += sum;
                        }
                        for (int c = q; c < nt.C; c += nq) {
                            float sum = 0.f;
                            for (int r = 0; r < TR_MAX; ++r) sum += s.Dy[c * STR + r];
                            s.dWouts[nt.C * HP + c] += sum;
                        }
                    }
                }
            }
            // data gradient: delta_{l-1}[k] = (sum_o W_l[o][k] delta_l[o]) * (1 - h_{l-1}[k]^2)
            float dz[HP];
#pragma unroll
            for (int o = 0; o < HP; ++o) dz[o] = o < HID ? s.Ds[o * STR + t] : 0.f;
            __syncthreads();  // every reader of the old Ds (dW phase, dz loads) is done
            if (active) {
                const float* Wt = s.Wt + (size_t)(l - 1) * HID * HP;
#pragma unroll 2
                for (int k = 0; k < HID; ++k) {
                    const float4* w = reinterpret_cast<const float4*>(Wt + k * HP);
                    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                    for (int o4 = 0; o4 < HP / 4; ++o4) {
                        const float4 ww = w[o4];
                        s0 = fmaf(ww.x, dz[4 * o4 + 0], s0);
                        s1 = fmaf(ww.y, dz[4 * o4 + 1], s1);
                        s2 = fmaf(ww.z, dz[4 * o4 + 2], s2);
                        s3 = fmaf(ww.w, dz[4 * o4 + 3], s3);
                    }
                    const float h = Hprev[k * STR + t];
                    s.Ds[k * STR + t] = ((s0 + s1) + (s2 + s3)) * (1.f - h * h);
                }
            }
            __syncthreads();
        }
        // ---- layer 0: Ds now holds delta0[o][t].  G0[j][o] = sum_n delta0, dA[n][o] += sum_j delta0
        for (int i = t; i < ti.np * HP; i += blockDim.x) {
            const int jj = i / HP, o = i % HP;
            float sum = 0.f;
            if (o < HID)
                for (int n = 0; n < ti.nn; ++n) sum += s.Ds[o * STR + n * tl.TP + jj];
            s.G0s[i] = sum;
        }
        if (want_w) {
            for (int i = t; i < ti.nn * HID; i += blockDim.x) {
                const int n = i / HID, o = i % HID;
                float sum = 0.f;
                for (int jj = 0; jj < ti.np; ++jj) sum += s.Ds[o * STR + n * tl.TP + jj];
                s.dAs[n * HP + o] += sum;
            }
        }
        __syncthreads();
        if (want_w) {
            // learn: hand sum_n delta0 to the layer-0 kernel (dW0 coord block = sum_p G0[p] (x) beta[p])
            float* dst = a.G0 + ((size_t)ti.d * tl.P + ti.p0) * HP;
            for (int i = t; i < ti.np * HP; i += blockDim.x) dst[i] = s.G0s[i];  // N <= 64 => one latent chunk per pixel
        }
        if (a.want_affine) {
            // match: dL/dcoords for the tile's pixels, reduced to (dT[2], dM[2][2]) partials of this tile
            const int pe = nt.pe;
            // dbeta[j][k] = sum_o G0[j][o] W0c[o][k]  -> reuse beta storage in place: dphi[j][m]
            for (int i = t; i < ti.np * pe; i += blockDim.x) {
                const int jj = i / pe, m = i % pe;
                const float4* ws_ = reinterpret_cast<const float4*>(a.W0cT + (size_t)m * HP);
                const float4* wc_ = reinterpret_cast<const float4*>(a.W0cT + (size_t)(pe + m) * HP);
                const float4* g4 = reinterpret_cast<const float4*>(s.G0s + jj * HP);
                float ds = 0.f, dc = 0.f;
                for (int o4 = 0; o4 < HP / 4; ++o4) {
                    const float4 g = g4[o4], a4 = __ldg(ws_ + o4), b4 = __ldg(wc_ + o4);
                    ds += g.x * a4.x + g.y * a4.y + g.z * a4.z + g.w * a4.w;
                    dc += g.x * b4.x + g.y * b4.y + g.z * b4.z + g.w * b4.w;
                }
                const float sn = s.beta[jj * 2 * pe + m], cs = s.beta[jj * 2 * pe + pe + m];
                // d/dphi of (sin, cos) = (cos, -sin)
                s.beta[jj * 2 * pe + m] = ds * cs - dc * sn;  // dphi, overwrites sin (each (jj,m) owned by one thread)
            }
            __syncthreads();
            float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (t < ti.np) {
                float g0 = 0.f, g1 = 0.f;
                for (int m = 0; m < pe; ++m) {
                    const float dphi = s.beta[t * 2 * pe + m];
                    g0 = fmaf(dphi, a.B[m], g0);
                    g1 = fmaf(dphi, a.B[pe + m], g1);
                }
                const int p = ti.p0 + t;
                const float r0 = a.coords[2 * p] - s.affc[6], r1 = a.coords[2 * p + 1] - s.affc[7];
                v[0] = g0, v[1] = g1, v[2] = g0 * r0, v[3] = g0 * r1, v[4] = g1 * r0, v[5] = g1 * r1;
            }
            // TP <= 32: the first warp holds every contribution
            if (t < 32) {
#pragma unroll
                for (int q = 0; q < 6; ++q) v[q] = warp_sum(v[q]);
                if (t == 0) {
                    float* dst = a.affpart + ((size_t)ti.d * tl.tilesP + ti.p0 / tl.TP) * 6;
                    for (int q = 0; q < 6; ++q) dst[q] = v[q];
                }
            }
        }
        __syncthreads();
    }
    if (want_w) {
        float* dst = a.partials + (size_t)blockIdx.x * partial_size(nt, tl, HP);
        const size_t n1 = (size_t)(L - 1) * HID * HID;
        for (size_t i = t; i < n1; i += blockDim.x) dst[i] = s.dWs[i];
        dst += n1;
        for (int i = t; i < (L - 1) * HP; i += blockDim.x) dst[i] = s.dbs[i];
        dst += (L - 1) * HP;
        for (int i = t; i < nt.C * HP + 4; i += blockDim.x) dst[i] = s.dWouts[i];
        dst += nt.C * HP + 4;
        for (int i = t; i < tl.NC * HP; i += blockDim.x) dst[i] = s.dAs[i];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// learn: layer-0 coordinate block  dW0c[o][k] = sum_{d,p} G0[d,p][o] beta[d,p][k]   (split over pixel chunks)
// ---------------------------------------------------------------------------------------------------------------
constexpr int L0_CHUNK = 64;

template <int HID>
__global__ void __launch_bounds__(256) l0_partial_kernel(Args a, float* l0part /*[grid, HID*2pe]*/) {
    constexpr int HP = Smem<HID>::HP;
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
    constexpr int HP = Smem<HID>::HP;
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
__global__ void finish_affine_kernel(Args a, float* g_angles, float* g_scalings, float* g_shears, float* g_translation) {
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

static WsLayout ws_layout(const Net& nt, const Tiling& tl, bool bwd) {
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

static int make_net(const lmr_inr_desc* d, Net* nt) {
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

static int fill_args(Args& a, const Net& nt, const Tiling& tl, const float* params, const float* B, const float* auxB, const float* grid,
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
    prep_A_kernel<HID><<<(a.tl.N * Smem<HID>::HP + 255) / 256, 256, 0, st>>>(a);
    LMR_LAUNCH_OK();
    return 0;
}

int forward(const lmr_inr_desc* desc, const float* params, const float* B, const float* auxB, const float* grid, int N, const float* coords,
            int P, const float* coord_shift, const lmr_affine* affine, int D, float* img, void* ws, size_t ws_bytes, cudaStream_t st) {
    Net nt;
    if (int rc = make_net(desc, &nt)) return rc;
    LMR_CHECK_ARG(N >= 1 && P >= 1 && D >= 1, "inr_forward: bad sizes N=%d P=%d D=%d", N, P, D);
    const Tiling tl = make_tiling(N, P, D);
    const WsLayout wl = ws_layout(nt, tl, false);
    if (ws_bytes < wl.total) {
        set_error("inr_forward: workspace %zu < %zu bytes", ws_bytes, wl.total);
        return LMR_ERR_WORKSPACE;
    }
    Args a;
    if (int rc = fill_args(a, nt, tl, params, B, auxB, grid, coords, coord_shift, affine, ws, wl)) return rc;
    a.img = img;
    if (int rc = launch_prep<50>(a, st)) return rc;
    float* dummy = nullptr;
    Smem<50> s(dummy, nt, tl, false, false);
    const size_t smem = s.floats * sizeof(float) + 16;
    static size_t smem_set_fwd = 0;  // per-process cache: avoid a driver call per launch (and inside graph capture)
    if (smem > smem_set_fwd) {
        LMR_CUDA_OK(cudaFuncSetAttribute(inr_fwd_kernel<50>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set_fwd = smem;
    }
    const int grid_x = tl.numTiles < 4 * num_sms() ? tl.numTiles : 4 * num_sms();
    inr_fwd_kernel<50><<<grid_x, TR_MAX, smem, st>>>(a);
    LMR_LAUNCH_OK();
    return 0;
}

int backward(const lmr_inr_desc* desc, const float* params, const float* B, const float* auxB, const float* grid, int N, const float* coords,
             int P, const float* coord_shift, const lmr_affine* affine, int D, const float* g_img, float* g_params, float* g_angles,
             float* g_scalings, float* g_shears, float* g_translation, void* ws, size_t ws_bytes, cudaStream_t st) {
    Net nt;
    if (int rc = make_net(desc, &nt)) return rc;
    LMR_CHECK_ARG(N >= 1 && P >= 1 && D >= 1, "inr_backward: bad sizes N=%d P=%d D=%d", N, P, D);
    LMR_CHECK_ARG(N <= 64, "inr_backward: N=%d latents per step not supported (max 64)", N);
    const bool want_aff = g_angles != nullptr;
    LMR_CHECK_ARG(g_params != nullptr || want_aff, "inr_backward: nothing to compute");
    if (want_aff) {
        LMR_CHECK_ARG(g_scalings && g_shears && g_translation, "inr_backward: incomplete affine gradient outputs");
        LMR_CHECK_ARG(affine && affine->angles, "inr_backward: affine gradients requested without an affine transform");
    }
    const Tiling tl = make_tiling(N, P, D);
    const WsLayout wl = ws_layout(nt, tl, true);
    if (ws_bytes < wl.total) {
        set_error("inr_backward: workspace %zu < %zu bytes", ws_bytes, wl.total);
        return LMR_ERR_WORKSPACE;
    }
    Args a;
    if (int rc = fill_args(a, nt, tl, params, B, auxB, grid, coords, coord_shift, affine, ws, wl)) return rc;
    a.g_img = g_img;
    a.want_weights = g_params != nullptr, a.want_affine = want_aff;
    if (int rc = launch_prep<50>(a, st)) return rc;
    float* dummy = nullptr;
    Smem<50> s(dummy, nt, tl, true, a.want_weights != 0);
    const size_t smem = s.floats * sizeof(float) + 16;
    LMR_CHECK_ARG(smem <= 227 * 1024, "inr_backward: configuration needs %zu bytes of shared memory (> 227 KB)", smem);
    static size_t smem_set_bwd = 0;
    if (smem > smem_set_bwd) {
        LMR_CUDA_OK(cudaFuncSetAttribute(inr_bwd_kernel<50>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set_bwd = smem;
    }
    inr_bwd_kernel<50><<<wl.gridMain, TR_MAX, smem, st>>>(a);
    LMR_LAUNCH_OK();
    if (a.want_weights) {
        float* l0part = reinterpret_cast<float*>(static_cast<char*>(ws) + wl.l0part);
        const size_t smem0 = (size_t)L0_CHUNK * (Smem<50>::HP + 2 * nt.pe) * sizeof(float);
        static size_t smem_set_l0 = 0;
        if (smem0 > smem_set_l0) {
            LMR_CUDA_OK(cudaFuncSetAttribute(l0_partial_kernel<50>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem0));
            smem_set_l0 = smem0;
        }
        l0_partial_kernel<50><<<wl.gridL0, 256, smem0, st>>>(a, l0part);
        LMR_LAUNCH_OK();
        const int blocks = (int)((nt.count() + 127) / 128);
        finish_weights_kernel<50><<<blocks, 128, 0, st>>>(a, l0part, wl.gridMain, wl.gridL0, g_params);
        LMR_LAUNCH_OK();
    }
    if (want_aff) {
        finish_affine_kernel<<<D, 128, 0, st>>>(a, g_angles, g_scalings, g_shears, g_translation);
        LMR_LAUNCH_OK();
    }
    return 0;
}

size_t workspace_bytes(const lmr_inr_desc* desc, int N, int P, int D, bool bwd) {
    Net nt;
    if (make_net(desc, &nt)) return 0;
    return ws_layout(nt, make_tiling(N, P, D), bwd).total;
}

size_t params_count(const lmr_inr_desc* desc) {
    Net nt;
    if (make_net(desc, &nt)) return 0;
    return nt.count();
}

}  // namespace simt
}  // namespace lmr
