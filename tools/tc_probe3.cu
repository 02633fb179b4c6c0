// Development probe 3 (bf16): MN-major (transposed) operand semantics with the no-swizzle core-matrix layout.
// Part 1: A (K-major, M=128, bf16) one-hot A[m][o] = (o == m%64); B stored K-major [o][k] and read MN-major (N = k, K = o):
//         D[m][n] = B_read(K = m%64, N = n).
// Part 2: dW-style product, both operands MN-major views of [128 rows x 64] tiles, M=64: D[f][g] = sum_r X[r][f] Y[r][g];
//         X one-hot rows X[r][f] = (f == r%64) and Y[r][g] = r/2 (pass 0) or g (pass 1)  =>  D[f][g] = sum over r with r%64==f.
#include <cstdio>
#include <vector>
#include <cuda_bf16.h>
#include "../laminr_b200/csrc/tc_common.cuh"
using namespace lmr::tcx;

__device__ void put(uint8_t* tile, int r, int k, int R, float v) {
    *reinterpret_cast<__nv_bfloat16*>(tile + tile_off(r, k / 8, R) + (k % 8) * 2) = __float2bfloat16(v);
}

__global__ void __launch_bounds__(128) probe(int part, int pass, uint32_t lbo, uint32_t sbo, uint32_t kstep, float* D) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t* sA = sm;                    // 128 x 64 bf16 = 16 KB
    uint8_t* sB = sm + 16384;            // part 1: 64 x 64 (8 KB); part 2: 128 x 64 (16 KB)
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int t = threadIdx.x, warp = t >> 5;
    for (int i = t; i < 128 * 64; i += 128) put(sA, i / 64, i % 64, 128, (i % 64 == (i / 64) % 64) ? 1.f : 0.f);
    if (part == 1) {
        for (int i = t; i < 64 * 64; i += 128) put(sB, i / 64, i % 64, 64, pass == 0 ? (float)(i / 64) : (float)(i % 64));
    } else {
        for (int i = t; i < 128 * 64; i += 128) put(sB, i / 64, i % 64, 128, pass == 0 ? (float)((i / 64) / 2) : (float)(i % 64));
    }
    if (warp == 0) tmem_alloc(&tmem_base, 64);
    if (t == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm = tmem_base;
    // zero the accumulator region first via a K-major MMA of zeros? simpler: first MMA uses accumulate = 0
    if (t == 0) {
        if (part == 1) {
            const uint32_t idesc = instr_desc(FMT_BF16, 128, 64, 0, 1);
            for (int ks = 0; ks < 4; ++ks) {  // K = 16 per MMA
                const uint64_t da = smem_desc(smem_u32(sA) + ks * 2 * (128 * 16), 128 * 16, 128);
                const uint64_t db = smem_desc(smem_u32(sB) + ks * kstep, lbo, sbo);
                mma_f16(tm, da, db, idesc, ks > 0);
            }
        } else {
            const uint32_t idesc = instr_desc(FMT_BF16, 64, 64, 1, 1);
            for (int ks = 0; ks < 8; ++ks) {  // 16 rows per MMA
                const uint64_t da = smem_desc(smem_u32(sA) + ks * kstep, lbo, sbo);
                const uint64_t db = smem_desc(smem_u32(sB) + ks * kstep, lbo, sbo);
                mma_f16(tm, da, db, idesc, ks > 0);
            }
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    fence_after_sync();
    float v[16];
    for (int c = 0; c < 64; c += 16) {
        tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + c, v);
        tmem_ld_wait();
        for (int q = 0; q < 16; ++q) D[t * 64 + c + q] = v[q];
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 64);
}

int main() {
    float* dD;
    cudaMalloc(&dD, 128 * 64 * 4);
    const int smem = 16384 * 2 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    {
        const uint32_t variants[][3] = {{128, 1024, 256}, {1024, 128, 256}};
        for (auto& v : variants) {
            std::vector<float> Do(128 * 64), Dk(128 * 64);
            for (int pass = 0; pass < 2; ++pass) {
                cudaMemset(dD, 0xff, 128 * 64 * 4);
                probe<<<1, 128, smem>>>(1, pass, v[0], v[1], v[2], dD);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("kernel: %s\n", cudaGetErrorString(e)); return 1; }
                cudaMemcpy(pass == 0 ? Do.data() : Dk.data(), dD, 128 * 64 * 4, cudaMemcpyDeviceToHost);
            }
            int good = 0;
            for (int kk = 0; kk < 64; ++kk)
                for (int n = 0; n < 64; ++n) good += (Do[kk * 64 + n] == (float)kk && Dk[kk * 64 + n] == (float)n);
            printf("PART1 LBO=%u SBO=%u kstep=%u : %d / 4096 read as B[o=K][k=N]\n", v[0], v[1], v[2], good);
            const int ks[] = {0, 1, 7, 8, 9, 16, 17}, ns[] = {0, 1, 7, 8, 9, 16, 32};
            for (int kk : ks) {
                printf("   K=%2d:", kk);
                for (int n : ns) printf("  N=%2d->(o=%2.0f,k=%2.0f)", n, Do[kk * 64 + n], Dk[kk * 64 + n]);
                printf("\n");
            }
        }
    }
    {
        const uint32_t variants[][3] = {{128, 2048, 256}, {2048, 128, 256}};
        for (auto& v : variants) {
            std::vector<float> D0(128 * 64), D1(128 * 64);
            for (int pass = 0; pass < 2; ++pass) {
                cudaMemset(dD, 0xff, 128 * 64 * 4);
                probe<<<1, 128, smem>>>(2, pass, v[0], v[1], v[2], dD);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("kernel: %s\n", cudaGetErrorString(e)); return 1; }
                cudaMemcpy(pass == 0 ? D0.data() : D1.data(), dD, 128 * 64 * 4, cudaMemcpyDeviceToHost);
            }
            // expected: D[f][g] = sum_{r: r%64==f} Y[r][g]; pass0: (f/2) + ((f+64)/2) = f/2 + f/2 + 32 (integer division); pass1: 2*g
            printf("PART2 LBO=%u SBO=%u kstep=%u : lane map of result rows (f -> TMEM lane), checking pass1 == 2g and pass0 == 2*(f/2)+32\n", v[0], v[1], v[2]);
            int located = 0;
            for (int f = 0; f < 64; ++f) {
                int hit = -1;
                for (int lane = 0; lane < 128 && hit < 0; ++lane) {
                    bool ok = true;
                    for (int g = 0; g < 64 && ok; ++g) ok = (D1[lane * 64 + g] == 2.f * g) && (D0[lane * 64 + g] == (float)(2 * (f / 2) + 32));
                    if (ok) hit = lane;
                }
                located += hit >= 0;
                if (f % 16 == 0) printf("\n   ");
                printf(" %2d->%3d", f, hit);
            }
            printf("\n   located %d / 64;  sample lane0: pass0 %g %g pass1 %g %g %g | lane 16: %g %g | lane 32: %g %g\n", located, D0[0], D0[1], D1[0], D1[1], D1[2],
                   D0[16 * 64], D1[16 * 64 + 1], D0[32 * 64], D1[32 * 64 + 1]);
        }
    }
    return 0;
}
