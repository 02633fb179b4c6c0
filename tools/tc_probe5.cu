// Development probe 5: is the ~48 cycles/MMA an issue-rate limit of ONE issuing thread?  W warps issue reps/W MMAs each (own accumulator).
#include <cstdio>
#include "../laminr_b200/csrc/tc_common.cuh"
using namespace lmr::tcx;

__global__ void __launch_bounds__(256) probe(int nwarps, int reps, int ncols, int mrows, long long* out) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bars[8];
    __shared__ uint32_t tmem_base;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    for (int i = t; i < 131072 / 4; i += 256) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;
    if (warp == 0) tmem_alloc(&tmem_base, 512);
    if (t == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1); mbar_fence_init(); }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm = tmem_base;
    long long t0 = clock64();
    if (warp < nwarps && lane == 0) {
        const uint32_t idesc = instr_desc(FMT_BF16, mrows, ncols, 0, 0);
        const uint64_t da = smem_desc(smem_u32(sm) + warp * 4096, 2048, 128), db = smem_desc(smem_u32(sm + 65536) + warp * 2048, 1024, 128);
        for (int i = 0; i < reps / nwarps; ++i) mma_f16(tm + warp * 64, da, db, idesc, i > 0);
        mma_commit(&bars[warp]);
    }
    if (warp < nwarps) mbar_wait(&bars[warp], 0);
    __syncthreads();
    long long t1 = clock64();
    if (t == 0) out[0] = t1 - t0;
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
    long long* d; cudaMalloc(&d, 64);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072 + 1024);
    for (int m : {128, 64}) for (int ncols : {64, 128, 256}) for (int nw : {1, 2, 4}) {
        const int reps = 768;
        probe<<<1, 256, 131072 + 1024>>>(nw, reps, ncols, m, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); return 1; }
        long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("M=%3d N=%3d issuing warps %d : %7.1f cycles per MMA (aggregate)\n", m, ncols, nw, (double)h / reps);
    }
    return 0;
}
