// Development probe (not part of the product): validates the tcgen05 descriptors / TMEM layouts used by inr_tc.cu on a B200.
//   T1  D[128x64] = A[128x64] . B[64x64]^T          (A, B K-major core-matrix tiles, kind::tf32, M=128)
//   T2  D[64x64]  = A^T . A2  over the 128 rows      (both operands MN-major views of K-major tiles, M=64): prints the TMEM lane map
//   T3  D[128x64] = A[128x64] . W  with W stored K-major [o][k] and read as an MN-major B operand
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../laminr_b200/csrc/tc_common.cuh"
using namespace lmr::tcx;

__global__ void __launch_bounds__(128) probe(const float* A, const float* A2, const float* B, float* D1, float* D2raw, float* D3) {
    extern __shared__ __align__(1024) uint8_t sm[];
    float* sA = reinterpret_cast<float*>(sm);            // 128 x 64 tf32, 32 KB
    float* sA2 = sA + 128 * 64;                          // 32 KB
    float* sB = sA2 + 128 * 64;                          // 64 x 64, 16 KB
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int t = threadIdx.x, warp = t >> 5;
    for (int i = t; i < 128 * 64; i += 128) {
        const int r = i / 64, k = i % 64;
        *reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sA) + tile_off(r, k / 4, 128) + (k % 4) * 4) = A[i];
        *reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sA2) + tile_off(r, k / 4, 128) + (k % 4) * 4) = A2[i];
    }
    for (int i = t; i < 64 * 64; i += 128) {
        const int r = i / 64, k = i % 64;
        *reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sB) + tile_off(r, k / 4, 64) + (k % 4) * 4) = B[i];
    }
    if (warp == 0) tmem_alloc(&tmem_base, 256);
    if (t == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm = tmem_base;
    uint32_t phase = 0;
    // ---- T1
    if (t == 0) {
        const uint32_t idesc = instr_desc(FMT_TF32, 128, 64, 0, 0);
        for (int ks = 0; ks < 8; ++ks) {
            const uint64_t da = smem_desc(smem_u32(sA) + ks * 2 * (128 * 16), 128 * 16, 128);
            const uint64_t db = smem_desc(smem_u32(sB) + ks * 2 * (64 * 16), 64 * 16, 128);
            mma_tf32(tm + 0, da, db, idesc, ks > 0);
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, phase); phase ^= 1;
    fence_after_sync();
    {
        float v[16];
        for (int c = 0; c < 64; c += 16) {
            tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + c, v);
            tmem_ld_wait();
            for (int q = 0; q < 16; ++q) D1[t * 64 + c + q] = v[q];
        }
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    // ---- T2 (M=64, both MN-major): D[m][n] = sum_r A[r][m] A2[r][n]
    if (t == 0) {
        const uint32_t idesc = instr_desc(FMT_TF32, 64, 64, 1, 1);
        for (int ks = 0; ks < 16; ++ks) {  // 8 rows per MMA
            const uint64_t da = smem_desc(smem_u32(sA) + ks * 128, 128, 128 * 16);
            const uint64_t db = smem_desc(smem_u32(sA2) + ks * 128, 128, 128 * 16);
            mma_tf32(tm + 64, da, db, idesc, ks > 0);
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, phase); phase ^= 1;
    fence_after_sync();
    {
        float v[16];
        for (int c = 0; c < 64; c += 16) {
            tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + 64 + c, v);
            tmem_ld_wait();
            for (int q = 0; q < 16; ++q) D2raw[t * 64 + c + q] = v[q];
        }
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    // ---- T3: D[m][k] = sum_o A[m][o] B[o][k]   (B read MN-major: N = k, K = o)
    if (t == 0) {
        const uint32_t idesc = instr_desc(FMT_TF32, 128, 64, 0, 1);
        for (int ks = 0; ks < 8; ++ks) {
            const uint64_t da = smem_desc(smem_u32(sA) + ks * 2 * (128 * 16), 128 * 16, 128);
            const uint64_t db = smem_desc(smem_u32(sB) + ks * 128, 128, 64 * 16);  // 8 o-rows per MMA
            mma_tf32(tm + 128, da, db, idesc, ks > 0);
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, phase); phase ^= 1;
    fence_after_sync();
    {
        float v[16];
        for (int c = 0; c < 64; c += 16) {
            tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + 128 + c, v);
            tmem_ld_wait();
            for (int q = 0; q < 16; ++q) D3[t * 64 + c + q] = v[q];
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 256);
}

int main() {
    std::vector<float> A(128 * 64), A2(128 * 64), B(64 * 64);
    srand(1);
    auto rnd = [] { return (float)((rand() % 17) - 8) / 8.0f; };  // exactly representable in tf32; sums exact in fp32
    for (auto& x : A) x = rnd();
    for (auto& x : A2) x = rnd();
    for (auto& x : B) x = rnd();
    float *dA, *dA2, *dB, *dD1, *dD2, *dD3;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dA2, A2.size() * 4); cudaMalloc(&dB, B.size() * 4);
    cudaMalloc(&dD1, 128 * 64 * 4); cudaMalloc(&dD2, 128 * 64 * 4); cudaMalloc(&dD3, 128 * 64 * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dA2, A2.data(), A2.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD2, 0xff, 128 * 64 * 4);
    const int smem = (128 * 64 * 2 + 64 * 64) * 4 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<<<1, 128, smem>>>(dA, dA2, dB, dD1, dD2, dD3);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> D1(128 * 64), D2(128 * 64), D3(128 * 64);
    cudaMemcpy(D1.data(), dD1, D1.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(D2.data(), dD2, D2.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(D3.data(), dD3, D3.size() * 4, cudaMemcpyDeviceToHost);
    double e1 = 0, e3 = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
            double r1 = 0, r3 = 0;
            for (int k = 0; k < 64; ++k) r1 += (double)A[m * 64 + k] * B[n * 64 + k], r3 += (double)A[m * 64 + k] * B[k * 64 + n];
            e1 = fmax(e1, fabs(r1 - D1[m * 64 + n]));
            e3 = fmax(e3, fabs(r3 - D3[m * 64 + n]));
        }
    printf("T1 (K-major A,B; M=128) max abs err %g\n", e1);
    printf("T3 (B as MN-major view)  max abs err %g\n", e3);
    // T2: find which TMEM lane holds each of the 64 result rows
    std::vector<double> R(64 * 64);
    for (int m = 0; m < 64; ++m)
        for (int n = 0; n < 64; ++n) {
            double r = 0;
            for (int k = 0; k < 128; ++k) r += (double)A[k * 64 + m] * A2[k * 64 + n];
            R[m * 64 + n] = r;
        }
    int found = 0;
    printf("T2 (M=64, MN-major A,B) row -> lane:");
    for (int m = 0; m < 64; ++m) {
        int hit = -1;
        for (int lane = 0; lane < 128; ++lane) {
            bool ok = true;
            for (int n = 0; n < 64 && ok; ++n) ok = fabs(R[m * 64 + n] - D2[lane * 64 + n]) < 1e-3;
            if (ok) { hit = lane; break; }
        }
        if (hit >= 0) ++found;
        if (m % 8 == 0) printf("\n   ");
        printf(" %2d->%3d", m, hit);
    }
    printf("\nT2 rows located: %d / 64\n", found);
    return 0;
}
