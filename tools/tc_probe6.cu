// Development probe 6: (a) where does a cta_group::1 M=64 tcgen05.mma put row m of D in TMEM, and can a second M=64 tile share the columns at
// lane offset 16;  (b) the register fragment of tcgen05.ld.16x256b.x1 (also with the lane base at +16).
// D[r][n] = (r + 1) + 128 n  (A[r][0] = r + 1, A[r][1] = 1; B[n][0] = 1, B[n][1] = 128 n: all exact in bf16).
#include <cstdio>
#include <cuda_bf16.h>
#include "../laminr_b200/csrc/tc_common.cuh"
using namespace lmr::tcx;

__global__ void __launch_bounds__(128) probe(int laneoff, float* raw /*[128][64]*/, float* frag /*[4][32][4]*/) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    uint8_t* A = sm;           // [64 rows][64 k] bf16, K-major core-matrix layout, R = 64
    uint8_t* B = sm + 8192;    // [64 n][64 k]
    for (int i = t; i < 16384 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0u;
    __syncthreads();
    if (t < 64) {
        const int r = t;
        __nv_bfloat16* a = reinterpret_cast<__nv_bfloat16*>(A + tile_off(r, 0, 64));
        a[0] = __float2bfloat16((float)(r + 1)), a[1] = __float2bfloat16(1.f);
        __nv_bfloat16* b = reinterpret_cast<__nv_bfloat16*>(B + tile_off(r, 0, 64));
        b[0] = __float2bfloat16(1.f), b[1] = __float2bfloat16(128.f * r);
    }
    if (warp == 0) tmem_alloc(&tmem_base, 64);
    if (t == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm = tmem_base;
    {  // sentinel -1 everywhere
        uint32_t v[8];
        for (int i = 0; i < 8; ++i) v[i] = __float_as_uint(-1.f);
        for (int c = 0; c < 64; c += 8) tmem_st8(tm + c + ((uint32_t)(32 * warp) << 16), v);
        tmem_st_wait();
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (t == 0) {
        const uint32_t idesc = instr_desc(FMT_BF16, 64, 64, 0, 0);
        const uint64_t da = smem_desc(smem_u32(A), 64 * 16, 128), db = smem_desc(smem_u32(B), 64 * 16, 128);
        for (int kc = 0; kc < 4; ++kc)
            mma_f16(tm + ((uint32_t)laneoff << 16), da + (uint64_t)kc * ((2 * 64 * 16) >> 4), db + (uint64_t)kc * ((2 * 64 * 16) >> 4), idesc, kc > 0);
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    fence_after_sync();
    {
        float v[16];
        for (int c = 0; c < 64; c += 16) {
            tmem_ld16(tm + c + ((uint32_t)(32 * warp) << 16), v);
            tmem_ld_wait();
            for (int i = 0; i < 16; ++i) raw[(32 * warp + lane) * 64 + c + i] = v[i];
        }
    }
    {
        uint32_t r[4];
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                     : "r"(tm + 8 + ((uint32_t)(32 * warp + laneoff) << 16)));   // columns 8..15
        tmem_ld_wait();
        for (int i = 0; i < 4; ++i) frag[(warp * 32 + lane) * 4 + i] = __uint_as_float(r[i]);
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 64);
}

int main() {
    float *raw, *frag;
    cudaMalloc(&raw, 128 * 64 * 4), cudaMalloc(&frag, 4 * 32 * 4 * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 1024);
    static float h[128 * 64], f[512];
    for (int laneoff : {0, 16}) {
        probe<<<1, 128, 16384 + 1024>>>(laneoff, raw, frag);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("laneoff %d: %s\n", laneoff, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, raw, sizeof(h), cudaMemcpyDeviceToHost), cudaMemcpy(f, frag, sizeof(f), cudaMemcpyDeviceToHost);
        printf("== D lane offset %d: TMEM lane -> row of D (from column 0; -1 = untouched), and whether columns decode as n\n", laneoff);
        for (int L = 0; L < 128; ++L) {
            const float v = h[L * 64];
            int row = v < 0 ? -1 : (int)v - 1;
            bool ok = true;
            if (row >= 0) for (int n = 0; n < 64; ++n) ok = ok && h[L * 64 + n] == (float)(row + 1) + 128.f * n;
            printf("%d%s ", row, row >= 0 && !ok ? "!" : "");
            if (L % 32 == 31) printf("\n");
        }
        printf("   16x256b.x1 at columns 8..15, lane base 32w + %d: thread -> (row, col) of its 4 registers\n", laneoff);
        for (int w = 0; w < 4; ++w) {
            printf("   warp %d:", w);
            for (int l = 0; l < 32; ++l) {
                if (l == 0 || l == 1 || l == 3 || l == 4 || l == 31) {
                    printf("  t%d:", l);
                    for (int i = 0; i < 4; ++i) { const int v = (int)f[(w * 32 + l) * 4 + i]; printf(" (%d,%d)", v % 128 - 1, v / 128); }
                }
            }
            printf("\n");
        }
    }
    return 0;
}
