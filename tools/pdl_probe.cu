// Probe (B200): does programmatic dependent launch (PDL) combine with cooperative launches, inside stream capture, and what does a kernel boundary
// cost with and without it?  Chain of 5 single-wave kernels (148 CTAs, ~10 us each, the last 2 us of every kernel on ONE straggler CTA), replayed
// as a CUDA graph.       nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pdl_probe tools/pdl_probe.cu && tools/pdl_probe
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(512, 1) work(float* buf, int iters, int pdl, unsigned* sync, int use_barrier) {
    if (pdl) asm volatile("griddepcontrol.launch_dependents;");
    __shared__ float s[1024];
    s[threadIdx.x] = threadIdx.x;  // "prologue"
    __syncthreads();
    if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    float v = buf[blockIdx.x * 512 + threadIdx.x];
    const int n = blockIdx.x == 0 ? iters + iters / 5 : iters;
    for (int i = 0; i < n; ++i) v = fmaf(v, 1.0001f, 0.001f);
    if (use_barrier) {  // software grid barrier (needs co-residency)
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(sync, 1u);
            while (*(volatile unsigned*)sync % gridDim.x != 0) __nanosleep(32);
        }
        __syncthreads();
    }
    buf[blockIdx.x * 512 + threadIdx.x] = v + s[(threadIdx.x + 1) & 511];
}

static int launch(cudaStream_t st, float* buf, int iters, int pdl, int coop, unsigned* sync) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148), cfg.blockDim = dim3(512), cfg.dynamicSmemBytes = 0, cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (coop) attr[na].id = cudaLaunchAttributeCooperative, attr[na].val.cooperative = 1, ++na;
    if (pdl) attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization, attr[na].val.programmaticStreamSerializationAllowed = 1, ++na;
    cfg.attrs = attr, cfg.numAttrs = na;
    cudaError_t e = cudaLaunchKernelEx(&cfg, work, buf, iters, pdl, sync, coop);
    if (e != cudaSuccess) { printf("  launch error (pdl=%d coop=%d): %s\n", pdl, coop, cudaGetErrorString(e)); cudaGetLastError(); return 1; }
    return 0;
}

static float run(int pdl, int coop, int graph) {
    cudaStream_t st;
    cudaStreamCreate(&st);
    float* buf;
    unsigned* sync;
    cudaMalloc(&buf, 148 * 512 * 4), cudaMemset(buf, 0, 148 * 512 * 4);
    cudaMalloc(&sync, 4), cudaMemset(sync, 0, 4);
    const int iters = 4000, chain = 5, reps = 200;
    cudaGraphExec_t ex = nullptr;
    int bad = 0;
    if (graph) {
        cudaGraph_t g;
        cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
        for (int k = 0; k < chain; ++k) bad |= launch(st, buf, iters, pdl, coop && (k & 1), sync);
        cudaError_t e = cudaStreamEndCapture(st, &g);
        if (e != cudaSuccess || bad) { printf("  capture failed: %s\n", cudaGetErrorString(e)); cudaGetLastError(); return -1; }
        e = cudaGraphInstantiate(&ex, g, 0);
        if (e != cudaSuccess) { printf("  instantiate failed: %s\n", cudaGetErrorString(e)); cudaGetLastError(); return -1; }
    }
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    for (int w = 0; w < 2; ++w) {
        cudaEventRecord(a, st);
        for (int r = 0; r < reps; ++r) {
            if (graph) cudaGraphLaunch(ex, st);
            else for (int k = 0; k < chain; ++k) bad |= launch(st, buf, iters, pdl, coop && (k & 1), sync);
        }
        cudaEventRecord(b, st);
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess || bad) { printf("  run failed: %s\n", cudaGetErrorString(e)); return -1; }
    }
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    cudaFree(buf), cudaFree(sync);
    return ms * 1000.f / (reps * chain);
}

int main() {
    for (int graph = 0; graph < 2; ++graph)
        for (int coop = 0; coop < 2; ++coop)
            for (int pdl = 0; pdl < 2; ++pdl) {
                printf("graph=%d cooperative(odd kernels)=%d pdl=%d:\n", graph, coop, pdl);
                const float us = run(pdl, coop, graph);
                printf("  %.2f us per kernel\n", us);
            }
    return 0;
}
