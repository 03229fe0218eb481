// micro-benchmark: tcgen05.mma SS issue rate (M=128) for the tile shapes the fused MixedOp kernel can use.
//   same accumulator vs 4 rotating accumulators, tf32 (K=8) and bf16 (K=16), N in {64, 128, 256};
//   plus the cost of a 148-CTA grid barrier (sense-reversal on two L2 words).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I bm-nas_b200/csrc -o tools/ubench/mma_rate tools/ubench/mma_rate.cu
#include <cstdio>
#include "tc_ptx.cuh"
using namespace bmnas::tc;

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// nmma MMAs issued back to back by one thread, accumulators rotate over `nacc` TMEM regions of N columns
template <int N, bool BF>
__global__ void __launch_bounds__(128, 1) k_rate(int nmma, int nacc, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tbase;
    for (int i = threadIdx.x; i < (16384 + N * 128) / 4; i += 128) reinterpret_cast<float*>(sm)[i] = 0.f;
    if (threadIdx.x < 32) tmem_alloc(&tbase, 512);
    if (threadIdx.x == 32) { mbar_init(&bar, 1); fence_barrier_init(); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x < 32) {          // the whole warp runs the issue loop (uniform control flow), one elected lane issues
        const uint32_t a = s32(sm), b = s32(sm + 16384);
        const uint32_t id = BF ? idesc_bf16(128, N) : idesc_tf32(128, N);
        const uint32_t tb = tbase;
        const bool leader = elect_one();
        long long t0 = clock64();
#pragma unroll 4
        for (int i = 0; i < nmma; ++i) {
            const uint32_t d = tb + (uint32_t)((i & (nacc - 1)) * N);
            const uint32_t ko = (uint32_t)(i & 3) * 32u;
            if (leader) {
                if (BF) umma_bf16(d, kdesc(a + ko), kdesc(b + ko), id, 1u);
                else umma_tf32(d, kdesc(a + ko), kdesc(b + ko), id, 1u);
            }
        }
        if (leader) umma_commit(&bar);
        long long t1 = clock64();
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        if (leader) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tbase, 512);
}

__global__ void k_gridbar(unsigned int* w, int iters, long long* out) {
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        __syncthreads();
        if (threadIdx.x == 0) {
            volatile unsigned int* gen = w + 1;
            const unsigned int g = *gen;
            __threadfence();
            if (atomicAdd(w, 1u) == gridDim.x - 1) { w[0] = 0; __threadfence(); atomicAdd(w + 1, 1u); }
            else while (*gen == g) {}
            __threadfence();
        }
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = (clock64() - t0) / iters;
}

template <int N, bool BF>
void run(long long* d_out) {
    const int smem = 16384 + N * 128 + 2048;
    cudaFuncSetAttribute(k_rate<N, BF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int nacc : {1, 2, 4}) {
        if (nacc * N > 512) continue;
        long long h[2];
        for (int rep = 0; rep < 2; ++rep) {
            k_rate<N, BF><<<1, 128, smem>>>(512, nacc, d_out);
            cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
        }
        const double floor_ = 128.0 * N / 256.0;
        printf("%s M=128 N=%3d nacc=%d: issue %.1f cyc/mma, complete %.1f cyc/mma (floor %.0f) -> %.0f%% of floor rate\n", BF ? "bf16" : "tf32",
               N, nacc, h[0] / 512.0, h[1] / 512.0, floor_, 100.0 * floor_ / (h[1] / 512.0));
    }
}

int main() {
    long long* d_out; cudaMalloc(&d_out, 64);
    run<64, false>(d_out); run<128, false>(d_out); run<256, false>(d_out);
    run<64, true>(d_out); run<128, true>(d_out); run<256, true>(d_out);
    unsigned int* w; cudaMalloc(&w, 8); cudaMemset(w, 0, 8);
    for (int grid : {12, 48, 148}) {
        long long h;
        for (int rep = 0; rep < 2; ++rep) { k_gridbar<<<grid, 256>>>(w, 1000, d_out); cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost); }
        printf("grid barrier %3d CTAs: %lld cycles\n", grid, h);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
