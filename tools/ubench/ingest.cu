// per-SM L2->SM ingest rate: cp.async.bulk vs LDG.128 vs cp.async, hot (all CTAs same addresses) vs distinct
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long gt() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// method 0: bulk TMA in `piece`-byte copies; 1: LDG.128 -> STS; 2: cp.async 16B
__global__ void __launch_bounds__(256, 1) k_ingest(const uint8_t* src, long long cta_stride, int bytes, int piece, int method,
                                                    unsigned long long* tout, float* sink) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar;
    const uint8_t* my = src + (long long)blockIdx.x * cta_stride;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned long long t0 = gt();
    if (method == 0) {
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(s32(&bar)) : "memory");
            for (int o = 0; o < bytes; o += piece)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(sm + o)),
                             "l"(my + o), "r"(piece), "r"(s32(&bar)) : "memory");
        }
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                         : "=r"(done) : "r"(s32(&bar)), "r"(0u) : "memory");
        }
    } else if (method == 1) {
        const int n16 = bytes / 16;
        for (int base = 0; base < n16; base += 256 * 16) {
            float4 v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                int i = base + j * 256 + threadIdx.x;
                v[j] = i < n16 ? __ldcg(reinterpret_cast<const float4*>(my) + i) : make_float4(0, 0, 0, 0);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                int i = base + j * 256 + threadIdx.x;
                if (i < n16) reinterpret_cast<float4*>(sm)[i] = v[j];
            }
        }
        __syncthreads();
    } else {
        const int n16 = bytes / 16;
        for (int i = threadIdx.x; i < n16; i += 256)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(sm + i * 16)), "l"(my + (long long)i * 16) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
    }
    unsigned long long t1 = gt();
    if (threadIdx.x == 0) { tout[blockIdx.x * 2] = t0; tout[blockIdx.x * 2 + 1] = t1; }
    if (sink && threadIdx.x == 999) sink[0] = reinterpret_cast<float*>(sm)[5];
}

int main() {
    const int maxcta = 148, maxbytes = 192 * 1024;
    uint8_t* src; unsigned long long* tout;
    CK(cudaMalloc(&src, (size_t)maxcta * maxbytes)); CK(cudaMemset(src, 1, (size_t)maxcta * maxbytes));
    CK(cudaMalloc(&tout, maxcta * 16));
    CK(cudaFuncSetAttribute(k_ingest, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    unsigned long long h[maxcta * 2];
    const char* mn[] = {"bulk", "ldg128", "cp.async"};
    for (int ctas : {1, 24, 72, 148}) for (int hot = 1; hot >= 0; --hot) for (int bytes : {32768, 131072}) for (int method = 0; method < 3; ++method)
        for (int piece : {32768, 4096}) {
            if (method != 0 && piece != 32768) continue;
            double best = 1e30;
            for (int rep = 0; rep < 5; ++rep) {
                k_ingest<<<ctas, 256, 196 * 1024>>>(src, hot ? 0 : maxbytes, bytes, piece, method, tout, nullptr);
                CK(cudaDeviceSynchronize());
                CK(cudaMemcpy(h, tout, ctas * 16, cudaMemcpyDeviceToHost));
                unsigned long long a = ~0ull, b = 0; double avg = 0;
                for (int i = 0; i < ctas; ++i) { if (h[2 * i] < a) a = h[2 * i]; if (h[2 * i + 1] > b) b = h[2 * i + 1]; avg += (double)(h[2 * i + 1] - h[2 * i]); }
                avg /= ctas;
                if (rep > 0 && avg < best) best = avg;
            }
            printf("ctas=%3d %s bytes=%6d %-8s piece=%5d : avg per-CTA %.2f us -> %.1f GB/s per SM, %.2f TB/s total\n", ctas, hot ? "hot " : "dist", bytes,
                   mn[method], piece, best / 1e3, bytes / best, bytes / best * ctas / 1e3);
        }
    return 0;
}
