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
#include "inr_common.cuh"

namespace lmr {
namespace simt {

using namespace lmr::inr;

constexpr int STR = 132;     // row stride (floats) of the k-major activation panels
constexpr int TP_CAP = 32;   // pixels per tile cap

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
// host side
// ---------------------------------------------------------------------------------------------------------------
int forward(const lmr_inr_desc* desc, const float* params, const float* B, const float* auxB, const float* grid, int N, const float* coords,
            int P, const float* coord_shift, const lmr_affine* affine, int D, float* img, void* ws, size_t ws_bytes, cudaStream_t st) {
    Net nt;
    if (int rc = make_net(desc, &nt)) return rc;
    LMR_CHECK_ARG(N >= 1 && P >= 1 && D >= 1, "inr_forward: bad sizes N=%d P=%d D=%d", N, P, D);
    const Tiling tl = make_tiling(N, P, D, TP_CAP);
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
    const Tiling tl = make_tiling(N, P, D, TP_CAP);
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
        if (int rc = launch_l0_partial<50>(a, l0part, wl.gridL0, st)) return rc;
        if (int rc = launch_finish_weights<50>(a, l0part, wl.gridMain, wl.gridL0, g_params, st)) return rc;
    }
    if (want_aff) {
        finish_affine_kernel<<<D, 128, 0, st>>>(a, tl.tilesP, g_angles, g_scalings, g_shears, g_translation);
        LMR_LAUNCH_OK();
    }
    return 0;
}

size_t workspace_bytes(const lmr_inr_desc* desc, int N, int P, int D, bool bwd) {
    Net nt;
    if (make_net(desc, &nt)) return 0;
    return ws_layout(nt, make_tiling(N, P, D, TP_CAP), bwd).total;
}

size_t params_count(const lmr_inr_desc* desc) {
    Net nt;
    if (make_net(desc, &nt)) return 0;
    return nt.count();
}

}  // namespace simt
}  // namespace lmr
