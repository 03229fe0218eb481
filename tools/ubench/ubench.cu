// micro-benchmarks that decide the B=96 design: launch cadence (stream / graph / PDL) vs grid-barrier cost
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void k_empty(float* p) { if (p && threadIdx.x == 9999) p[0] = 1.f; }
__global__ void k_pdl(float* p) {
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (p && threadIdx.x == 9999) p[0] = 1.f;
}
// one dependent global round trip: read a[i], write b[i]
__global__ void k_rt(const float* a, float* b, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = a[i] + 1.f;
}
__global__ void k_rt_pdl(const float* a, float* b, int n) {
    asm volatile("griddepcontrol.launch_dependents;");
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (i < n) b[i] = a[i] + 1.f;
}

__device__ __forceinline__ void grid_barrier(unsigned int* ctr, unsigned int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        while (*((volatile unsigned int*)ctr) < target) {}
        __threadfence();
    }
    __syncthreads();
}
__global__ void k_persist(unsigned int* ctr, int iters, float* a, float* b, int n, int work) {
    unsigned int target = 0;
    for (int it = 0; it < iters; ++it) {
        if (work) {
            const float* src = (it & 1) ? b : a;
            float* dst = (it & 1) ? a : b;
            for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = __ldcg(src + i) + 1.f;
        }
        target += gridDim.x;
        grid_barrier(ctr, target);
    }
}

int main() {
    float *a, *b; unsigned int* ctr;
    const int n = 96 * 1024;
    CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4)); CK(cudaMalloc(&ctr, 4));
    CK(cudaMemset(a, 0, n * 4)); CK(cudaMemset(ctr, 0, 4));
    cudaStream_t s; CK(cudaStreamCreate(&s));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    const int R = 2000;
    for (int grid : {1, 96, 148, 592}) {
        for (int w = 0; w < 2; ++w) {
            cudaEventRecord(e0, s);
            for (int i = 0; i < R; ++i) k_empty<<<grid, 256, 0, s>>>(nullptr);
            cudaEventRecord(e1, s); CK(cudaStreamSynchronize(s));
        }
        cudaEventElapsedTime(&ms, e0, e1);
        printf("stream empty grid=%d: %.3f us/launch\n", grid, ms * 1e3 / R);
    }
    for (int w = 0; w < 2; ++w) {
        cudaEventRecord(e0, s);
        for (int i = 0; i < R; ++i) k_rt<<<n / 256, 256, 0, s>>>((i & 1) ? b : a, (i & 1) ? a : b, n);
        cudaEventRecord(e1, s); CK(cudaStreamSynchronize(s));
    }
    cudaEventElapsedTime(&ms, e0, e1);
    printf("stream roundtrip(384 CTAs, 393KB): %.3f us/launch\n", ms * 1e3 / R);
    // graphs: chain of 200 nodes
    for (int mode = 0; mode < 4; ++mode) {
        cudaGraph_t g; cudaGraphExec_t ge;
        CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal));
        for (int i = 0; i < 200; ++i) {
            cudaLaunchConfig_t cfg = {};
            cfg.blockDim = 256; cfg.stream = s;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at;
            if (mode == 0) { cfg.gridDim = 148; cfg.numAttrs = 0; CK(cudaLaunchKernelEx(&cfg, k_empty, (float*)nullptr)); }
            if (mode == 1) { cfg.gridDim = 148; cfg.numAttrs = 1; CK(cudaLaunchKernelEx(&cfg, k_pdl, (float*)nullptr)); }
            if (mode == 2) { cfg.gridDim = n / 256; cfg.numAttrs = 0; CK(cudaLaunchKernelEx(&cfg, k_rt, (const float*)((i & 1) ? b : a), (i & 1) ? a : b, n)); }
            if (mode == 3) { cfg.gridDim = n / 256; cfg.numAttrs = 1; CK(cudaLaunchKernelEx(&cfg, k_rt_pdl, (const float*)((i & 1) ? b : a), (i & 1) ? a : b, n)); }
        }
        CK(cudaStreamEndCapture(s, &g));
        CK(cudaGraphInstantiate(&ge, g, 0));
        for (int w = 0; w < 3; ++w) {
            cudaEventRecord(e0, s);
            for (int r = 0; r < 10; ++r) CK(cudaGraphLaunch(ge, s));
            cudaEventRecord(e1, s); CK(cudaStreamSynchronize(s));
        }
        cudaEventElapsedTime(&ms, e0, e1);
        const char* names[] = {"graph empty", "graph empty+PDL", "graph roundtrip", "graph roundtrip+PDL"};
        printf("%s: %.3f us/node\n", names[mode], ms * 1e3 / 2000);
    }
    // persistent kernel with grid barriers
    for (int grid : {96, 148}) for (int work = 0; work < 2; ++work) {
        const int iters = 2000;
        for (int w = 0; w < 2; ++w) {
            CK(cudaMemsetAsync(ctr, 0, 4, s));
            cudaEventRecord(e0, s);
            k_persist<<<grid, 256, 0, s>>>(ctr, iters, a, b, n, work);
            cudaEventRecord(e1, s); CK(cudaStreamSynchronize(s));
        }
        cudaEventElapsedTime(&ms, e0, e1);
        printf("persistent grid=%d work=%d: %.3f us/phase\n", grid, work, ms * 1e3 / iters);
    }
    return 0;
}
