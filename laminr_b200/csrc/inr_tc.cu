// INR generator, tcgen05 tensor-core path (LMR_PREC_BF16X3 / LMR_PREC_BF16X3_FAST): forward and recompute-backward.
//
// Same tiling and layer-0 separability as inr_simt.cu (a tile = TP pixels x N latents <= 128 rows of one target; row r
// <-> TMEM lane r).  The 50x50 hidden layers run on the 5th-gen tensor cores:
//   * operands are error-compensated split-bf16: x = hi + lo (bf16 each), products hi*hi + lo*hi + hi*lo (3 tcgen05.mma,
//     kind::f16, fp32 accumulate in TMEM) -> ~2^-16 relative operand error, fp32-class results (precision probe);
//   * activation tiles [128 rows x 64 features] live in shared memory in the canonical no-swizzle core-matrix layout
//     (tc_common.cuh).  Feature 50 is the constant 1: the bias is column 50 of the weight tile and the bias gradient
//     column 50 of the weight-gradient accumulator.  The weight tile's entry [50][50] = 20 regenerates the constant through
//     the activation (tanh(20) == 1 in fp32) and every other padding entry is 0 (tanh(0) == 0), so all epilogues are
//     uniform over the 64 columns -- no per-column branches;
//   * forward      z_l   = h_l W_l^T        A = h tile (K-major),      B = W_l tile (K-major)         M=128 N=64 K=64
//     data grad    dh_l  = dz_l W_l         A = dz tile (K-major),     B = W_l tile read MN-major     M=128 N=64 K=64
//     weight grad  dW_l += dz_l^T h_l       A = dz tile read MN-major, B = h tile read MN-major       M=64  N=64 K=128
//     -- the transposed products reuse the SAME bytes through MN-major descriptors (no transposed copies);
//   * weight-gradient accumulators stay resident in TMEM across all tiles of the persistent CTA;
//   * nothing is stored between forward and backward: the backward recomputes the forward per tile (L+1 tiles of smem),
//     and the weight-gradient MMAs of a layer overlap the epilogue that produces the next dz tile (rotating buffers).
// 512 threads per CTA: thread = (row, column quarter); warp w reads TMEM lanes 32*(w%4).. and columns 16*(w/4)...
#include <cuda_bf16.h>

#include "inr_common.cuh"
#include "tc_common.cuh"

namespace lmr {
namespace tc {

using namespace lmr::inr;
using namespace lmr::tcx;

constexpr int HID = 50;               // hidden width served by this path
constexpr int HPAD = 64;              // padded feature count of a tile (K and N of the MMAs)
constexpr int ONE_COL = 50;           // feature slot holding the constant 1 (bias column)
constexpr float ONE_SEED = 20.f;      // tanh(20) == 1
constexpr int TP_CAP_TC = 8;          // pixels per tile cap (bounds the per-tile scratch)
constexpr int NT = 512;               // threads per CTA
constexpr int TILE_HALF = 128 * HPAD * 2;   // bytes of the hi (or lo) half of an activation tile
constexpr int TILE_BYTES = 2 * TILE_HALF;   // 32 KB
constexpr int WT_HALF = HPAD * HPAD * 2;    // 8 KB
constexpr int WT_BYTES = 2 * WT_HALF;       // 16 KB
constexpr int DY_HALF = 128 * 8 * 2;        // dy tile: 128 rows x 8 channels (padded), hi / lo

// TMEM columns
constexpr uint32_t TM_ACC = 0;        // [128 lanes x 64]  forward / data-gradient accumulator
constexpr uint32_t TM_DW = 64;        // (L-1) x [64 rows (M=64 lane layout) x 64] weight-gradient accumulators, then 16 columns for dWout
constexpr uint32_t TM_COLS = 512;

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// packed fp32x2 arithmetic (FFMA2 / FMUL2 / FADD2 on sm_100): halves the issue slots of the epilogue maths
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                       rc = *reinterpret_cast<unsigned long long*>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2*>(&rd);
}
template <bool FAST>
__device__ __forceinline__ float act_tanh(float x) {
    if (FAST) return tanh_mufu(x);
    // 1 - 2/(e^{2x} + 1): saturates cleanly (e -> inf gives 1, e -> 0 gives -1); ~1e-7 absolute error
    return fmaf(-2.f, rcp_approx(ex2_approx(x * 2.885390081777927f) + 1.f), 1.f);
}
template <bool FAST>
__device__ __forceinline__ float2 act_tanh2(float2 x) {
    if (FAST) return make_float2(tanh_mufu(x.x), tanh_mufu(x.y));
    const float2 t = mul2(x, make_float2(2.885390081777927f, 2.885390081777927f));
    const float2 d = add2(make_float2(ex2_approx(t.x), ex2_approx(t.y)), make_float2(1.f, 1.f));
    return fma2(make_float2(-2.f, -2.f), make_float2(rcp_approx(d.x), rcp_approx(d.y)), make_float2(1.f, 1.f));
}

// split 8 floats (4 pairs) into bf16 hi / lo and store the two 16-byte chunks of row r, feature chunk c
__device__ __forceinline__ void store_chunk2(uint8_t* tile, int r, int c, const float2* v) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[q].x, v[q].y);
        const uint32_t hw = *reinterpret_cast<const uint32_t*>(&h2);
        const float2 hf = make_float2(__uint_as_float(hw << 16), __uint_as_float(hw & 0xffff0000u));  // a bf16 is the upper half of an fp32
        const float2 d = fma2(hf, make_float2(-1.f, -1.f), v[q]);
        const __nv_bfloat162 l2 = __floats2bfloat162_rn(d.x, d.y);
        hi[q] = hw;
        lo[q] = *reinterpret_cast<const uint32_t*>(&l2);
    }
    const uint32_t off = tile_off(r, c, 128);
    *reinterpret_cast<uint4*>(tile + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(tile + TILE_HALF + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// read back 8 features (hi + lo) of row r, chunk c, as 4 pairs
__device__ __forceinline__ void load_chunk2(const uint8_t* tile, int r, int c, float2* v) {
    const uint32_t off = tile_off(r, c, 128);
    const uint4 h = *reinterpret_cast<const uint4*>(tile + off);
    const uint4 l = *reinterpret_cast<const uint4*>(tile + TILE_HALF + off);
    const uint32_t hh[4] = {h.x, h.y, h.z, h.w}, ll[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
    for (int q = 0; q < 4; ++q)
        v[q] = add2(make_float2(__uint_as_float(hh[q] << 16), __uint_as_float(hh[q] & 0xffff0000u)),
                    make_float2(__uint_as_float(ll[q] << 16), __uint_as_float(ll[q] & 0xffff0000u)));
}

struct TcSmem {
    uint8_t* tiles;   // ntiles x TILE_BYTES
    uint8_t* wt;      // (L-1) x WT_BYTES
    uint8_t* dy;      // 2 x DY_HALF
    float *Cp, *beta, *Wout, *bout, *affc, *scratch, *As;  // scratch: output-layer partial sums [128][4][C], later G0s [TP][SHP]
    uint64_t* bars;   // [0] accumulator ready, [1] all MMAs of the tile done
    uint32_t* tmem;
    size_t bytes;
    __host__ __device__ TcSmem(uint8_t* base, const Net& nt, const Tiling& tl, int ntiles) {
        uint8_t* p = base;
        auto take = [&](size_t n) {
            uint8_t* r = p;
            p += (n + 127) / 128 * 128;
            return r;
        };
        tiles = take((size_t)ntiles * TILE_BYTES);
        wt = take((size_t)(nt.nl - 1) * WT_BYTES);
        dy = take(2 * DY_HALF);
        Cp = reinterpret_cast<float*>(take((size_t)tl.TP * HPAD * 4));
        beta = reinterpret_cast<float*>(take((size_t)tl.TP * 2 * nt.pe * 4));
        Wout = reinterpret_cast<float*>(take((size_t)nt.C * HPAD * 4));
        bout = reinterpret_cast<float*>(take(16));
        affc = reinterpret_cast<float*>(take(48));
        const size_t sc1 = (size_t)128 * 4 * nt.C * 4, sc2 = (size_t)tl.TP * SHP * 4;
        scratch = reinterpret_cast<float*>(take(sc1 > sc2 ? sc1 : sc2));
        bars = reinterpret_cast<uint64_t*>(take(16));
        tmem = reinterpret_cast<uint32_t*>(take(16));
        // A[n] rows staged in shared memory when they fit (L1 is almost entirely carved out for shared memory here)
        As = tl.NC <= 24 ? reinterpret_cast<float*>(take((size_t)tl.NC * SHP * 4)) : nullptr;
        bytes = p - base;
    }
};

// weights of hidden layers 1..L-1 as split-bf16 K-major tiles: W_l[o][k] for k < 50, column 50 = bias b_l[o],
// entry [50][50] = ONE_SEED (regenerates the constant-1 feature), rest 0.  Also the fp32 output layer and zeroed Cp padding.
__device__ inline void load_weight_tiles(const Args& a, TcSmem& s) {
    const Net& nt = a.net;
    for (int l = 1; l < nt.nl; ++l) {
        const float* W = a.params + nt.offW(l);
        const float* b = a.params + nt.offb(l);
        uint8_t* dst = s.wt + (size_t)(l - 1) * WT_BYTES;
        for (int i = threadIdx.x; i < HPAD * HPAD; i += blockDim.x) {
            const int o = i / HPAD, k = i % HPAD;
            float v = 0.f;
            if (o < HID) v = k < HID ? W[o * HID + k] : (k == ONE_COL ? b[o] : 0.f);
            else if (o == ONE_COL && k == ONE_COL) v = ONE_SEED;
            const __nv_bfloat16 hi = __float2bfloat16(v);
            const __nv_bfloat16 lo = __float2bfloat16(v - __bfloat162float(hi));
            const uint32_t off = tile_off(o, k >> 3, HPAD) + (k & 7) * 2;
            *reinterpret_cast<__nv_bfloat16*>(dst + off) = hi;
            *reinterpret_cast<__nv_bfloat16*>(dst + WT_HALF + off) = lo;
        }
    }
    const float* Wo = a.params + nt.offWout();
    for (int i = threadIdx.x; i < nt.C * HPAD; i += blockDim.x) {
        const int c = i / HPAD, k = i % HPAD;
        s.Wout[i] = k < HID ? Wo[c * HID + k] : 0.f;
    }
    if (threadIdx.x < 4) s.bout[threadIdx.x] = threadIdx.x < nt.C ? a.params[nt.offbout() + threadIdx.x] : 0.f;
    for (int i = threadIdx.x; i < a.tl.TP * HPAD; i += blockDim.x) s.Cp[i] = 0.f;  // columns >= 52 stay 0 forever
}

// Phase 0 of a tile: fetch Cp[j][o] (precomputed for every pixel by the prep kernel) into shared memory; for the affine
// gradient (match backward) also beta[j][2pe] = sin/cos(coords.B)
__device__ inline void tile_prologue(const Args& a, TcSmem& s, const TileIdx& ti, bool need_beta) {
    const Net& nt = a.net;
    const int pe = nt.pe;
    if (need_beta) {
        if (a.has_aff && threadIdx.x == 0) affine_consts(a.aff, ti.d, s.affc);
        __syncthreads();
        for (int i = threadIdx.x; i < ti.np * pe; i += blockDim.x) {
            const int j = i / pe, m = i % pe;
            const float2 c = transformed_coord(a, s.affc, ti.p0 + j);
            float sn, cs;
            sincosf(c.x * a.B[m] + c.y * a.B[pe + m], &sn, &cs);
            s.beta[j * 2 * pe + m] = sn;
            s.beta[j * 2 * pe + pe + m] = cs;
        }
    }
    const float4* src = reinterpret_cast<const float4*>(a.Cpg + ((size_t)ti.d * a.tl.P + ti.p0) * SHP);
    for (int i = threadIdx.x; i < ti.np * (SHP / 4); i += blockDim.x)
        reinterpret_cast<float4*>(s.Cp + (i / (SHP / 4)) * HPAD)[i % (SHP / 4)] = __ldg(src + i);
    __syncthreads();
}

// h1 = tanh(A[n] + Cp[j]) for this thread's 16 columns (A[n][50] = ONE_SEED from the prep kernel -> constant-1 feature)
template <bool FAST>
__device__ __forceinline__ void layer0_to_tile(const Args& a, const TcSmem& s, uint8_t* tile, int row, int quarter, float actf, int n, int nl, int j) {
    const float2* Cj = reinterpret_cast<const float2*>(s.Cp + j * HPAD + quarter * 16);
    const float2 act2 = make_float2(actf, actf);
    float2 v[8];
    if (s.As != nullptr) {
        const float* An = s.As + nl * SHP;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int o = quarter * 16 + 2 * q;  // SHP is even: a pair never straddles the end of the row
            const float2 av = o < SHP ? *reinterpret_cast<const float2*>(An + o) : make_float2(0.f, 0.f);
            v[q] = mul2(act_tanh2<FAST>(add2(av, Cj[q])), act2);
        }
    } else {
        const float* An = a.Avec + (size_t)n * SHP;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int o = quarter * 16 + 2 * q;
            const float2 av = o < SHP ? __ldg(reinterpret_cast<const float2*>(An + o)) : make_float2(0.f, 0.f);
            v[q] = mul2(act_tanh2<FAST>(add2(av, Cj[q])), act2);
        }
    }
    store_chunk2(tile, row, 2 * quarter, v);
    store_chunk2(tile, row, 2 * quarter + 1, v + 4);
}

// D(tmem_d)[128 x 64] = A . B^T with split-bf16 operands: hi*hi + lo*hi + hi*lo.  A K-major [128 x 64]; B is a [64 x 64] weight tile,
// K-major (b_transposed = 0: forward) or read MN-major (b_transposed = 1: data gradient).  Issued by ONE thread.
__device__ __forceinline__ void issue_rows_x_weights(uint32_t tmem_d, const uint8_t* a_tile, const uint8_t* w_tile, int b_transposed) {
    const uint32_t idesc = instr_desc(FMT_BF16, 128, HPAD, 0, b_transposed);
    uint64_t ah = smem_desc(smem_u32(a_tile), 128 * 16, 128);
    uint64_t bh = b_transposed ? smem_desc(smem_u32(w_tile), 128, HPAD * 16) : smem_desc(smem_u32(w_tile), HPAD * 16, 128);
    const uint64_t a_step = (2 * 128 * 16) >> 4, b_step = (b_transposed ? 256 : 2 * HPAD * 16) >> 4;  // start-address field is in 16-byte units
    const uint64_t a_lo = TILE_HALF >> 4, b_lo = WT_HALF >> 4;
#pragma unroll 1
    for (int ks = 0; ks < 4; ++ks) {  // K = 16 per MMA
        mma_f16(tmem_d, ah, bh, idesc, ks > 0);
        mma_f16(tmem_d, ah + a_lo, bh, idesc, 1);
        mma_f16(tmem_d, ah, bh + b_lo, idesc, 1);
        ah += a_step, bh += b_step;
    }
}

// D(tmem_d)[64 (M=64 lane layout) x n_cols] (+)= X^T . Y over the 128 rows: X a [128 x 64] tile, Y a [128 x 64] tile or the [128 x 8] dy tile,
// both read MN-major
__device__ __forceinline__ void issue_transposed_product(uint32_t tmem_d, const uint8_t* x_tile, const uint8_t* y_tile, int n_cols, uint32_t y_half,
                                                         bool zero_first) {
    const uint32_t idesc = instr_desc(FMT_BF16, 64, n_cols, 1, 1);
    uint64_t xh = smem_desc(smem_u32(x_tile), 128, 128 * 16), yh = smem_desc(smem_u32(y_tile), 128, 128 * 16);
    const uint64_t x_lo = TILE_HALF >> 4, y_lo = y_half >> 4;
#pragma unroll 1
    for (int ks = 0; ks < 8; ++ks) {  // 16 rows per MMA
        mma_f16(tmem_d, xh, yh, idesc, (ks > 0 || !zero_first) ? 1u : 0u);
        mma_f16(tmem_d, xh + x_lo, yh, idesc, 1);
        mma_f16(tmem_d, xh, yh + y_lo, idesc, 1);
        xh += 256 >> 4, yh += 256 >> 4;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------
template <bool FAST, int C>
__global__ void __launch_bounds__(NT, 2) inr_fwd_tc_kernel(Args a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    TcSmem s(smem_raw, a.net, a.tl, 1);
    const Net& nt = a.net;
    const Tiling& tl = a.tl;
    const int t = threadIdx.x, warp = t >> 5;
    const int row = t & 127, quarter = warp >> 2;
    const uint32_t lane_addr = ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(quarter * 16);
    const int L = nt.nl;
    load_weight_tiles(a, s);
    if (s.As != nullptr && tl.nChunks == 1)
        for (int i = t; i < tl.NC * SHP; i += NT) s.As[i] = a.Avec[i];
    if (warp == 0) tmem_alloc(s.tmem, 64);
    if (t == 0) {
        mbar_init(&s.bars[0], 1);
        mbar_fence_init();
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm = *s.tmem;
    uint32_t phase = 0;
    uint8_t* T = s.tiles;
    for (int tile = blockIdx.x; tile < tl.numTiles; tile += gridDim.x) {
        const TileIdx ti = decode_tile(tl, tile);
        if (s.As != nullptr && tl.nChunks > 1)
            for (int i = t; i < ti.nn * SHP; i += NT) s.As[i] = a.Avec[(size_t)ti.n0 * SHP + i];
        tile_prologue(a, s, ti, false);
        const int nl = row / tl.TP, j = row % tl.TP;
        const bool active = nl < ti.nn && j < ti.np;
        const float actf = active ? 1.f : 0.f;
        layer0_to_tile<FAST>(a, s, T, row, quarter, actf, active ? ti.n0 + nl : 0, active ? nl : 0, active ? j : 0);
#pragma unroll 1
        for (int l = 1; l < L; ++l) {
            fence_proxy_async();
            fence_before_sync();
            __syncthreads();
            if (t == 0) {
                fence_after_sync();
                issue_rows_x_weights(tm + TM_ACC, T, s.wt + (size_t)(l - 1) * WT_BYTES, 0);
                mma_commit(&s.bars[0]);
            }
            mbar_wait(&s.bars[0], phase);
            phase ^= 1;
            fence_after_sync();
            float2 v[8];
            tmem_ld16(tm + TM_ACC + lane_addr, reinterpret_cast<float*>(v));
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = act_tanh2<FAST>(v[q]);  // inactive rows: h_1 == 0 (bias feature too) => z == 0 => tanh == 0
            if (l < L - 1) {
                store_chunk2(T, row, 2 * quarter, v);
                store_chunk2(T, row, 2 * quarter + 1, v + 4);
            } else {  // output layer: partial dot products of this thread's 16 features
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const float2* w = reinterpret_cast<const float2*>(s.Wout + c * HPAD + quarter * 16);
                    float2 y = make_float2(0.f, 0.f);
#pragma unroll
                    for (int q = 0; q < 8; ++q) y = fma2(v[q], w[q], y);
                    s.scratch[(row * 4 + quarter) * C + c] = y.x + y.y;
                }
            }
        }
        fence_before_sync();
        __syncthreads();
        if (quarter == 0 && active) {
            const size_t base = (((size_t)ti.d * tl.N + (ti.n0 + nl)) * C) * tl.P + ti.p0 + j;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float* yp = s.scratch + (row * 4) * C + c;
                const float y = s.bout[c] + ((yp[0] + yp[C]) + (yp[2 * C] + yp[3 * C]));
                a.img[base + (size_t)c * tl.P] = act_tanh<FAST>(y);
            }
        }
        __syncthreads();  // scratch / Cp / the tile buffer are rewritten by the next tile
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 64);
}

// ---------------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------------
constexpr int STG = 132;  // row stride (floats) of the fp32 staging panel stage[o][row] for dz_0

template <bool FAST, int C>
__global__ void __launch_bounds__(NT, 1) inr_bwd_tc_kernel(Args a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const Net& nt = a.net;
    const Tiling& tl = a.tl;
    const int L = nt.nl;  // hidden layers; activation tiles T[0..L-1] hold h_1..h_L, one extra rotating tile X = T[L]
    TcSmem s(smem_raw, nt, tl, L + 1);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int row = t & 127, quarter = warp >> 2;
    const uint32_t lane_addr = ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(quarter * 16);
    const bool want_w = a.want_weights != 0;
    load_weight_tiles(a, s);
    if (s.As != nullptr)
        for (int i = t; i < tl.NC * SHP; i += NT) s.As[i] = a.Avec[i];
    if (warp == 0) tmem_alloc(s.tmem, TM_COLS);
    if (t == 0) {
        mbar_init(&s.bars[0], 1);
        mbar_init(&s.bars[1], 1);
        mbar_fence_init();
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm = *s.tmem;
    const uint32_t tm_dwout = TM_DW + (uint32_t)(L - 1) * 64;
    uint32_t ph_acc = 0, ph_done = 0;
    bool first = true;
    auto Tile = [&](int i) { return s.tiles + (size_t)i * TILE_BYTES; };
    // dz_0 is staged in fp32 ([o][row], conflict-free) in the tile that held h_2: it is not an MMA operand.  L == 2 has no free
    // tile before the last data-gradient epilogue, so the staging panel then lives in X's successor: host guarantees L >= 3 here.
    float* stage = reinterpret_cast<float*>(Tile(1));
    // dA[n][o] accumulators of this thread: entries e = t + NT*q  ->  (n = e / HID, o = e % HID); N <= 40 (host check)
    constexpr int DA_PER_THREAD = 4;
    float dA[DA_PER_THREAD] = {0.f, 0.f, 0.f, 0.f};

    for (int tile = blockIdx.x; tile < tl.numTiles; tile += gridDim.x) {
        const TileIdx ti = decode_tile(tl, tile);
        tile_prologue(a, s, ti, a.want_affine != 0);
        const int nl = row / tl.TP, j = row % tl.TP;
        const bool active = nl < ti.nn && j < ti.np;
        const float actf = active ? 1.f : 0.f;
        if (!first) {  // every MMA of the previous tile (they read the tiles we are about to overwrite) has completed
            mbar_wait(&s.bars[1], ph_done);
            ph_done ^= 1;
            fence_after_sync();
        }
        // ---- forward recompute: h_1 (CUDA cores), h_2..h_L (tensor cores), all kept as tiles
        layer0_to_tile<FAST>(a, s, Tile(0), row, quarter, actf, active ? ti.n0 + nl : 0, active ? nl : 0, active ? j : 0);
#pragma unroll 1
        for (int l = 1; l < L; ++l) {
            fence_proxy_async();
            fence_before_sync();
            __syncthreads();
            if (t == 0) {
                fence_after_sync();
                issue_rows_x_weights(tm + TM_ACC, Tile(l - 1), s.wt + (size_t)(l - 1) * WT_BYTES, 0);
                mma_commit(&s.bars[0]);
            }
            mbar_wait(&s.bars[0], ph_acc);
            ph_acc ^= 1;
            fence_after_sync();
            float2 v[8];
            tmem_ld16(tm + TM_ACC + lane_addr, reinterpret_cast<float*>(v));
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = act_tanh2<FAST>(v[q]);  // inactive rows: h_1 == 0 (bias feature too) => z == 0 => tanh == 0
            store_chunk2(Tile(l), row, 2 * quarter, v);
            store_chunk2(Tile(l), row, 2 * quarter + 1, v + 4);
            if (l == L - 1) {  // output layer: partial dot products of this thread's 16 features of h_L
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const float2* w = reinterpret_cast<const float2*>(s.Wout + c * HPAD + quarter * 16);
                    float2 y = make_float2(0.f, 0.f);
#pragma unroll
                    for (int q = 0; q < 8; ++q) y = fma2(v[q], w[q], y);
                    s.scratch[(row * 4 + quarter) * C + c] = y.x + y.y;
                }
            }
        }
        __syncthreads();
        // ---- output layer (CUDA cores): y, dy = g (1 - y^2), dz_{L-1} = (Wout^T dy) (1 - h_L^2) -> tile X; dy -> dy tile
        {
            float dy[4] = {0.f, 0.f, 0.f, 0.f};
            const size_t base = (((size_t)ti.d * tl.N + (ti.n0 + (active ? nl : 0))) * C) * tl.P + ti.p0 + (active ? j : 0);
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float* yp = s.scratch + (row * 4) * C + c;
                const float y = act_tanh<FAST>(s.bout[c] + ((yp[0] + yp[C]) + (yp[2 * C] + yp[3 * C])));
                const float g = active ? a.g_img[base + (size_t)c * tl.P] : 0.f;
                dy[c] = g * (1.f - y * y);
            }
            float2 hv[8], v[8];
            load_chunk2(Tile(L - 1), row, 2 * quarter, hv);
            load_chunk2(Tile(L - 1), row, 2 * quarter + 1, hv + 4);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                float2 dh = make_float2(0.f, 0.f);
#pragma unroll
                for (int c = 0; c < C; ++c)
                    dh = fma2(reinterpret_cast<const float2*>(s.Wout + c * HPAD + quarter * 16)[q], make_float2(dy[c], dy[c]), dh);
                // dh (1 - h^2); padding columns: Wout == 0; constant-1 column: 1 - 1 == 0
                v[q] = fma2(mul2(dh, mul2(hv[q], hv[q])), make_float2(-1.f, -1.f), dh);
            }
            store_chunk2(Tile(L), row, 2 * quarter, v);
            store_chunk2(Tile(L), row, 2 * quarter + 1, v + 4);
            if (want_w && quarter == 0) {  // dy as a [128 x 8] split-bf16 tile (one 16-byte chunk per row)
                uint32_t hi[2], lo[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const __nv_bfloat162 h2 = __floats2bfloat162_rn(dy[2 * q], dy[2 * q + 1]);
                    const float2 hf = __bfloat1622float2(h2);
                    const __nv_bfloat162 l2 = __floats2bfloat162_rn(dy[2 * q] - hf.x, dy[2 * q + 1] - hf.y);
                    hi[q] = *reinterpret_cast<const uint32_t*>(&h2), lo[q] = *reinterpret_cast<const uint32_t*>(&l2);
                }
                *reinterpret_cast<uint4*>(s.dy + row * 16) = make_uint4(hi[0], hi[1], 0u, 0u);
                *reinterpret_cast<uint4*>(s.dy + DY_HALF + row * 16) = make_uint4(lo[0], lo[1], 0u, 0u);
            }
        }
        // ---- layers L-1 .. 1.  dz_l sits in tile `src`; h_l in Tile(l-1); dz_{l-1} goes to `dst` (a tile no pending MMA reads)
        int src = L, dst = L - 1;
#pragma unroll 1
        for (int l = L - 1; l >= 1; --l) {
            fence_proxy_async();
            fence_before_sync();
            __syncthreads();
            if (t == 0) {
                fence_after_sync();
                if (want_w && l == L - 1)  // output layer: dWout[c][i] (dbout in column ONE_COL) = sum_r h_L[r][i] dy[r][c]
                    issue_transposed_product(tm + tm_dwout, Tile(L - 1), s.dy, 8, DY_HALF, first);
                issue_rows_x_weights(tm + TM_ACC, Tile(src), s.wt + (size_t)(l - 1) * WT_BYTES, 1);
                mma_commit(&s.bars[0]);
                if (want_w)  // dW_l[o][i] (bias gradient in column ONE_COL) += sum_r dz_l[r][o] h_l[r][i]; overlaps the epilogue below
                    issue_transposed_product(tm + TM_DW + (uint32_t)(l - 1) * 64, Tile(src), Tile(l - 1), HPAD, TILE_HALF, first);
                if (l == 1) mma_commit(&s.bars[1]);
            }
            mbar_wait(&s.bars[0], ph_acc);
            ph_acc ^= 1;
            fence_after_sync();
            // dz_{l-1}[k] = dh_l[k] (1 - h_l[k]^2)   (h_l[50] == 1 zeroes the bias column; padding columns of W are 0)
            float2 v[8], hv[8];
            tmem_ld16(tm + TM_ACC + lane_addr, reinterpret_cast<float*>(v));
            load_chunk2(Tile(l - 1), row, 2 * quarter, hv);
            load_chunk2(Tile(l - 1), row, 2 * quarter + 1, hv + 4);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = fma2(mul2(v[q], mul2(hv[q], hv[q])), make_float2(-1.f, -1.f), v[q]);
            if (l > 1) {
                store_chunk2(Tile(dst), row, 2 * quarter, v);
                store_chunk2(Tile(dst), row, 2 * quarter + 1, v + 4);
            } else {  // dz_0 feeds only CUDA-core reductions: stage it in fp32 (the tile of h_2 is free: dW_2's MMAs completed before bar 0 fired)
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    stage[(quarter * 16 + 2 * q) * STG + row] = v[q].x;
                    stage[(quarter * 16 + 2 * q + 1) * STG + row] = v[q].y;
                }
            }
            // rotate: the old src is free once the NEXT step's data-gradient MMAs (issued after this step's dW MMAs) have completed
            src = dst;
            dst = (src == L) ? L - 1 : L;
        }
        __syncthreads();  // dz_0 staged
        // ---- layer 0: G0[j][o] = sum_n dz0, dA[n][o] += sum_j dz0   (CUDA cores, fp32 staging panel)
        float* G0s = s.scratch;
        for (int i = t; i < ti.np * SHP; i += NT) {
            const int o = i / ti.np, jj = i % ti.np;  // consecutive lanes -> consecutive pixels of one feature row
            float sum = 0.f;
            if (o < HID) {
                const float* zr = stage + o * STG + jj;
                for (int n = 0; n < ti.nn; ++n) sum += zr[n * tl.TP];
            }
            G0s[jj * SHP + o] = sum;
        }
        if (want_w) {
#pragma unroll
            for (int q = 0; q < DA_PER_THREAD; ++q) {
                const int e = t + NT * q;
                if (e < ti.nn * HID) {
                    const int o = e / ti.nn, n = e % ti.nn;  // NOTE: entry order (o-major) is this kernel's private convention, see the read-out
                    const float* zr = stage + o * STG + n * tl.TP;
                    float sum = 0.f;
                    for (int jj = 0; jj < ti.np; ++jj) sum += zr[jj];
                    dA[q] += sum;
                }
            }
        }
        __syncthreads();
        if (want_w) {
            float* gdst = a.G0 + ((size_t)ti.d * tl.P + ti.p0) * SHP;
            for (int i = t; i < ti.np * SHP; i += NT) gdst[i] = G0s[i];
        }
        if (a.want_affine) {
            const int pe = nt.pe;
            for (int i = t; i < ti.np * pe; i += NT) {
                const int jj = i / pe, m = i % pe;
                const float4* ws_ = reinterpret_cast<const float4*>(a.W0cT + (size_t)m * SHP);
                const float4* wc_ = reinterpret_cast<const float4*>(a.W0cT + (size_t)(pe + m) * SHP);
                const float4* g4 = reinterpret_cast<const float4*>(G0s + jj * SHP);
                float ds = 0.f, dc = 0.f;
                for (int o4 = 0; o4 < SHP / 4; ++o4) {
                    const float4 g = g4[o4], a4 = __ldg(ws_ + o4), b4 = __ldg(wc_ + o4);
                    ds += g.x * a4.x + g.y * a4.y + g.z * a4.z + g.w * a4.w;
                    dc += g.x * b4.x + g.y * b4.y + g.z * b4.z + g.w * b4.w;
                }
                const float sn = s.beta[jj * 2 * pe + m], cs = s.beta[jj * 2 * pe + pe + m];
                s.beta[jj * 2 * pe + m] = ds * cs - dc * sn;  // d/dphi
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
            if (t < 32) {  // TP <= 8: the first warp holds every contribution
#pragma unroll
                for (int q = 0; q < 6; ++q) v[q] = warp_sum(v[q]);
                if (t == 0) {
                    float* pdst = a.affpart + ((size_t)ti.d * tl.tilesP + ti.p0 / tl.TP) * 6;
                    for (int q = 0; q < 6; ++q) pdst[q] = v[q];
                }
            }
        }
        first = false;
        __syncthreads();
    }
    // ---- drain: wait for the last tile's MMAs, then read the weight-gradient accumulators out of TMEM
    if (!first) {
        mbar_wait(&s.bars[1], ph_done);
        fence_after_sync();
    }
    if (want_w) {
        float* part = a.partials + (size_t)blockIdx.x * partial_size(nt, tl, SHP);
        float* p_dW = part;
        float* p_db = part + (size_t)(L - 1) * HID * HID;
        float* p_dWout = p_db + (size_t)(L - 1) * SHP;
        float* p_dA = p_dWout + (size_t)C * SHP + 4;
        if (warp < 4) {  // M=64 accumulator layout: row o -> lane (o/16)*32 + o%16
            const int o = 16 * warp + lane;
            const uint32_t lb = (uint32_t)(warp * 32) << 16;
#pragma unroll 1
            for (int l = 1; l < L; ++l) {
#pragma unroll 1
                for (int g = 0; g < 4; ++g) {
                    float v[16];
                    tmem_ld16(tm + TM_DW + (uint32_t)(l - 1) * 64 + lb + g * 16, v);
                    tmem_ld_wait();
                    if (lane < 16 && o < HID) {
#pragma unroll
                        for (int q = 0; q < 16; ++q) {
                            const int i = g * 16 + q;
                            if (i < HID) p_dW[(size_t)(l - 1) * HID * HID + o * HID + i] = v[q];
                            else if (i == ONE_COL) p_db[(l - 1) * SHP + o] = v[q];
                        }
                    }
                }
            }
            float v[16];
            tmem_ld16(tm + tm_dwout + lb, v);
            tmem_ld_wait();
            if (lane < 16) {  // row i = o of D[i][c]
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    if (o < HID) p_dWout[c * SHP + o] = v[c];
                    else if (o == ONE_COL) p_dWout[C * SHP + c] = v[c];
                }
            }
        }
#pragma unroll
        for (int q = 0; q < DA_PER_THREAD; ++q) {
            const int e = t + NT * q;
            if (e < tl.NC * HID) p_dA[(e % tl.NC) * SHP + e / tl.NC] = dA[q];  // e = o * N + n (single latent chunk: nn == NC == N)
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, TM_COLS);
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static int check_tc(const Net& nt) {
    LMR_CHECK_ARG(nt.hid == HID, "tensor-core path: hidden width %d not supported (50)", nt.hid);
    LMR_CHECK_ARG(nt.nl >= 3 && nt.nl <= 4, "tensor-core path: n_hidden %d out of range [3,4]", nt.nl);
    LMR_CHECK_ARG(nt.pe <= 64, "tensor-core path: pe_dim %d > 64", nt.pe);
    return 0;
}

template <typename K>
static int set_smem(K kernel, size_t bytes, size_t* cache) {
    if (bytes > *cache) {
        LMR_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        *cache = bytes;
    }
    return 0;
}

int forward(const lmr_inr_desc* desc, const float* params, const float* B, const float* auxB, const float* grid, int N, const float* coords,
            int P, const float* coord_shift, const lmr_affine* affine, int D, float* img, void* ws, size_t ws_bytes, int precision,
            cudaStream_t st) {
    Net nt;
    if (int rc = make_net(desc, &nt)) return rc;
    if (int rc = check_tc(nt)) return rc;
    LMR_CHECK_ARG(N >= 1 && P >= 1 && D >= 1, "inr_forward: bad sizes N=%d P=%d D=%d", N, P, D);
    const Tiling tl = make_tiling(N, P, D, TP_CAP_TC);
    const WsLayout wl = ws_layout(nt, tl, false, true);
    if (ws_bytes < wl.total) {
        set_error("inr_forward: workspace %zu < %zu bytes", ws_bytes, wl.total);
        return LMR_ERR_WORKSPACE;
    }
    Args a;
    if (int rc = fill_args(a, nt, tl, params, B, auxB, grid, coords, coord_shift, affine, ws, wl)) return rc;
    a.img = img;
    a.want_cp = 1;
    if (int rc = launch_prep<50>(a, st)) return rc;
    TcSmem s(nullptr, nt, tl, 1);
    const size_t smem = s.bytes + 1024;
    const int grid_x = tl.numTiles < 2 * num_sms() ? tl.numTiles : 2 * num_sms();
    static size_t cache[2][5] = {};
    const bool fast = precision == LMR_PREC_BF16X3_FAST;
#define LMR_FWD_CASE(FASTV, CV)                                                                          \
    if (fast == FASTV && nt.C == CV) {                                                                   \
        if (int rc = set_smem(inr_fwd_tc_kernel<FASTV, CV>, smem, &cache[FASTV][CV])) return rc;        \
        inr_fwd_tc_kernel<FASTV, CV><<<grid_x, NT, smem, st>>>(a);                                       \
    }
    LMR_FWD_CASE(false, 1) LMR_FWD_CASE(false, 2) LMR_FWD_CASE(false, 3) LMR_FWD_CASE(false, 4)
    LMR_FWD_CASE(true, 1) LMR_FWD_CASE(true, 2) LMR_FWD_CASE(true, 3) LMR_FWD_CASE(true, 4)
#undef LMR_FWD_CASE
    LMR_LAUNCH_OK();
    return 0;
}

int backward(const lmr_inr_desc* desc, const float* params, const float* B, const float* auxB, const float* grid, int N, const float* coords,
             int P, const float* coord_shift, const lmr_affine* affine, int D, const float* g_img, float* g_params, float* g_angles,
             float* g_scalings, float* g_shears, float* g_translation, void* ws, size_t ws_bytes, int precision, cudaStream_t st) {
    Net nt;
    if (int rc = make_net(desc, &nt)) return rc;
    if (int rc = check_tc(nt)) return rc;
    LMR_CHECK_ARG(N >= 1 && P >= 1 && D >= 1, "inr_backward: bad sizes N=%d P=%d D=%d", N, P, D);
    LMR_CHECK_ARG(N <= 40, "inr_backward (tensor-core path): N=%d latents per step not supported (max 40); use the fp32 path", N);
    const bool want_aff = g_angles != nullptr;
    LMR_CHECK_ARG(g_params != nullptr || want_aff, "inr_backward: nothing to compute");
    if (want_aff) {
        LMR_CHECK_ARG(g_scalings && g_shears && g_translation, "inr_backward: incomplete affine gradient outputs");
        LMR_CHECK_ARG(affine && affine->angles, "inr_backward: affine gradients requested without an affine transform");
    }
    const Tiling tl = make_tiling(N, P, D, TP_CAP_TC);
    const WsLayout wl = ws_layout(nt, tl, true, true);
    if (ws_bytes < wl.total) {
        set_error("inr_backward: workspace %zu < %zu bytes", ws_bytes, wl.total);
        return LMR_ERR_WORKSPACE;
    }
    Args a;
    if (int rc = fill_args(a, nt, tl, params, B, auxB, grid, coords, coord_shift, affine, ws, wl)) return rc;
    a.g_img = g_img;
    a.want_weights = g_params != nullptr, a.want_affine = want_aff;
    a.want_cp = 1;
    if (int rc = launch_prep<50>(a, st)) return rc;
    TcSmem s(nullptr, nt, tl, nt.nl + 1);
    const size_t smem = s.bytes + 1024;
    LMR_CHECK_ARG(smem <= 227 * 1024, "inr_backward (tensor-core path): needs %zu bytes of shared memory (> 227 KB)", smem);
    static size_t cache[2][5] = {};
    const bool fast = precision == LMR_PREC_BF16X3_FAST;
#define LMR_BWD_CASE(FASTV, CV)                                                                          \
    if (fast == FASTV && nt.C == CV) {                                                                   \
        if (int rc = set_smem(inr_bwd_tc_kernel<FASTV, CV>, smem, &cache[FASTV][CV])) return rc;        \
        inr_bwd_tc_kernel<FASTV, CV><<<wl.gridMain, NT, smem, st>>>(a);                                  \
    }
    LMR_BWD_CASE(false, 1) LMR_BWD_CASE(false, 2) LMR_BWD_CASE(false, 3) LMR_BWD_CASE(false, 4)
    LMR_BWD_CASE(true, 1) LMR_BWD_CASE(true, 2) LMR_BWD_CASE(true, 3) LMR_BWD_CASE(true, 4)
#undef LMR_BWD_CASE
    LMR_LAUNCH_OK();
    if (a.want_weights) {
        float* l0part = reinterpret_cast<float*>(static_cast<char*>(ws) + wl.l0part);
        const size_t smem0 = (size_t)L0_CHUNK * (SHP + 2 * nt.pe) * sizeof(float);
        static size_t c2 = 0;
        if (int rc = set_smem(l0_partial_kernel<50>, smem0, &c2)) return rc;
        l0_partial_kernel<50><<<wl.gridL0, 256, smem0, st>>>(a, l0part);
        LMR_LAUNCH_OK();
        if (int rc = launch_finish_weights<50>(a, l0part, wl.gridMain, wl.gridL0, g_params, st)) return rc;
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
    return ws_layout(nt, make_tiling(N, P, D, TP_CAP_TC), bwd, true).total;
}

}  // namespace tc
}  // namespace lmr
