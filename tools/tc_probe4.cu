// Development probe 4: tcgen05.mma throughput (cycles per MMA) for the shared-memory operand layouts used / considered by inr_tc.cu.
#include <cstdio>
#include <vector>
#include "../laminr_b200/csrc/tc_common.cuh"
using namespace lmr::tcx;

__device__ __forceinline__ uint64_t desc_sw(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return smem_desc(saddr, lbo, sbo) | ((uint64_t)layout << 61);
}

// mode 0: A K-major noswz [128x16], B K-major noswz [64x16]      (forward)
// mode 1: A K-major noswz, B MN-major view                       (data gradient)
// mode 2: A, B MN-major views, M=64                              (weight gradient)
// mode 3: A, B K-major SWIZZLE_128B (M=128,N=64)                 (what a swizzled layout would cost)
// mode 4: like 0 with N=32
// mode 5: like 2 with N=8
__global__ void __launch_bounds__(128) probe(int mode, int reps, long long* out) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int t = threadIdx.x, warp = t >> 5;
    for (int i = t; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;
    if (warp == 0) tmem_alloc(&tmem_base, 128);
    if (t == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm = tmem_base;
    const uint32_t a0 = smem_u32(sm), b0 = smem_u32(sm + 32768);
    long long t0 = 0, t1 = 0;
    if (t == 0) {
        uint32_t idesc; uint64_t da, db;
        if (mode == 0) { idesc = instr_desc(FMT_BF16, 128, 64, 0, 0); da = smem_desc(a0, 2048, 128); db = smem_desc(b0, 1024, 128); }
        else if (mode == 1) { idesc = instr_desc(FMT_BF16, 128, 64, 0, 1); da = smem_desc(a0, 2048, 128); db = smem_desc(b0, 128, 1024); }
        else if (mode == 2) { idesc = instr_desc(FMT_BF16, 64, 64, 1, 1); da = smem_desc(a0, 128, 2048); db = smem_desc(b0, 128, 2048); }
        else if (mode == 3) { idesc = instr_desc(FMT_BF16, 128, 64, 0, 0); da = desc_sw(a0, 16, 1024, 2); db = desc_sw(b0, 16, 1024, 2); }
        else if (mode == 4) { idesc = instr_desc(FMT_BF16, 128, 32, 0, 0); da = smem_desc(a0, 2048, 128); db = smem_desc(b0, 1024, 128); }
        else { idesc = instr_desc(FMT_BF16, 64, 8, 1, 1); da = smem_desc(a0, 128, 2048); db = smem_desc(b0, 128, 2048); }
        t0 = clock64();
        for (int i = 0; i < reps; ++i) mma_f16(tm, da, db, idesc, i > 0);
        mma_commit(&bar);
        mbar_wait(&bar, 0);
        t1 = clock64();
        out[mode] = t1 - t0;
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 128);
}

int main() {
    long long* d; cudaMalloc(&d, 64);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024);
    const char* names[] = {"fwd  A,B K-major noswz  M128 N64", "dgrad B MN-major view     M128 N64", "wgrad A,B MN-major views  M64  N64", "K-major SWIZZLE_128B      M128 N64",
                           "fwd  noswz                M128 N32", "wgrad MN-major            M64  N8 "};
    for (int reps : {12, 96, 768}) {
        for (int mode = 0; mode < 6; ++mode) {
            probe<<<1, 128, 65536 + 1024>>>(mode, reps, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
        }
        long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
        for (int mode = 0; mode < 6; ++mode) printf("reps %4d  %s : %8lld cycles total, %7.1f per MMA\n", reps, names[mode], h[mode], (double)h[mode] / reps);
    }
    return 0;
}
