// Development probe 2: which smem element does an MN-major (transposed) tf32 B operand read for (K, N)?
// A (K-major, M=128) is one-hot: A[m][o] = (o == m % 64)  =>  D[m][n] = B_read(K = m%64, N = n).
// B is filled with B[o][k] = o (pass 0) or k (pass 1) in the K-major core-matrix layout of tc_common.cuh.
#include <cstdio>
#include <vector>
#include "../laminr_b200/csrc/tc_common.cuh"
using namespace lmr::tcx;

__global__ void __launch_bounds__(128) probe(int pass, uint32_t lbo, uint32_t sbo, uint32_t kstep, float* D) {
    extern __shared__ __align__(1024) uint8_t sm[];
    float* sA = reinterpret_cast<float*>(sm);
    float* sB = sA + 128 * 64;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int t = threadIdx.x, warp = t >> 5;
    for (int i = t; i < 128 * 64; i += 128) {
        const int r = i / 64, k = i % 64;
        *reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sA) + tile_off(r, k / 4, 128) + (k % 4) * 4) = (k == r % 64) ? 1.f : 0.f;
    }
    for (int i = t; i < 64 * 64; i += 128) {
        const int o = i / 64, k = i % 64;
        *reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sB) + tile_off(o, k / 4, 64) + (k % 4) * 4) = pass == 0 ? (float)o : (float)k;
    }
    if (warp == 0) tmem_alloc(&tmem_base, 64);
    if (t == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm = tmem_base;
    if (t == 0) {
        const uint32_t idesc = instr_desc(FMT_TF32, 128, 64, 0, 1);
        for (int ks = 0; ks < 8; ++ks) {
            const uint64_t da = smem_desc(smem_u32(sA) + ks * 2 * (128 * 16), 128 * 16, 128);
            const uint64_t db = smem_desc(smem_u32(sB) + ks * kstep, lbo, sbo);
            mma_tf32(tm, da, db, idesc, ks > 0);
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
    const int smem = (128 * 64 + 64 * 64) * 4 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const uint32_t variants[][3] = {{128, 1024, 128}, {1024, 128, 128}, {128, 1024, 256}, {1024, 128, 2048}, {128, 1024, 2048}, {1024, 128, 256}};
    for (auto& v : variants) {
        std::vector<float> Do(128 * 64), Dk(128 * 64);
        for (int pass = 0; pass < 2; ++pass) {
            probe<<<1, 128, smem>>>(pass, v[0], v[1], v[2], dD);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("kernel: %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(pass == 0 ? Do.data() : Dk.data(), dD, 128 * 64 * 4, cudaMemcpyDeviceToHost);
        }
        int good = 0;
        for (int kk = 0; kk < 64; ++kk)
            for (int n = 0; n < 64; ++n) good += (Do[kk * 64 + n] == (float)kk && Dk[kk * 64 + n] == (float)n);
        printf("LBO=%u SBO=%u kstep=%u : %d / 4096 elements read as B[o=K][k=N]\n", v[0], v[1], v[2], good);
        const int ks[] = {0, 1, 7, 8, 9, 16}, ns[] = {0, 1, 3, 4, 5, 8, 32};
        for (int kk : ks) {
            printf("   K=%2d:", kk);
            for (int n : ns) printf("  N=%2d->(o=%2.0f,k=%2.0f)", n, Do[kk * 64 + n], Dk[kk * 64 + n]);
            printf("\n");
        }
    }
    return 0;
}
