// Shared device helpers for the BM-NAS B200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/bmnas_b200.h"

#define BMNAS_LAUNCH_CHECK()                                   \
    do {                                                       \
        cudaError_t e__ = cudaGetLastError();                  \
        if (e__ != cudaSuccess) return BMNAS_ELAUNCH;          \
    } while (0)

extern int bmnas_validate_only_flag;
#define BMNAS_DRY_RETURN()                         \
    do {                                           \
        if (bmnas_validate_only_flag) return BMNAS_OK; \
    } while (0)

extern int bmnas_pdl_flag;

namespace bmnas {

// Programmatic dependent launch.  Every kernel is launched with the stream-serialization attribute (a programmatic
// edge of the CUDA graph under capture).  Protocol, in every kernel:
//     [early section]  ->  pdl_wait()  ->  pdl_trigger()  ->  main body
//   pdl_wait()    griddepcontrol.wait: blocks until the preceding grid has COMPLETED and its writes are visible;
//   pdl_trigger() griddepcontrol.launch_dependents: lets the next kernel of the stream start (its CTAs are
//                 scheduled as soon as every CTA of this grid has triggered).
// Because a kernel only triggers after its own wait, the kernel that starts early knows that everything up to
// the grid BEFORE its predecessor is complete.  Its early section may therefore read data produced two or more
// kernels ago (weights, weight images, the x tile of a node op whose conv sits in between, forward tensors
// during the backward pass) and overlaps the predecessor's whole body; anything the predecessor writes is only
// touched after pdl_wait().  Kernels with no early section call pdl_prologue() first thing.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_prologue() {
    pdl_wait();
    pdl_trigger();
}

template <class... KArgs, class... Args>
static inline void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                            Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = bmnas_pdl_flag ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs
constexpr float kBnEps = 1e-5f;
constexpr float kLnEps = 1e-5f;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum of up to NV values per thread. `red` is shared scratch of at
// least NV * 32 floats.  Every thread receives the totals.  Deterministic.
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    __syncthreads();  // protect `red` from a previous use
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) red[i * 32 + warp] = v[i];
    }
    __syncthreads();
    // every warp re-reduces the per-warp partials with the same butterfly: identical bits in all threads
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(lane < nw ? red[i * 32 + lane] : 0.f);
}

// "last block done" election. Returns true in every thread of the block that
// arrives last among `expected` blocks; that block may then read what all the
// others wrote before their call. The counter is reset to 0 by the winner.
__device__ __forceinline__ bool last_block(unsigned int* counter, unsigned int expected) {
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(counter, 1u);
        s_last = (t == expected - 1);
        if (s_last) *counter = 0u;
    }
    __syncthreads();
    bool r = s_last != 0;
    if (r) __threadfence();
    return r;
}

__device__ __forceinline__ float ld_cg(const float* p) { return __ldcg(p); }

// ---------------------------------------------------------------- Philox4x32-7
// Seven rounds is the smallest Philox4x32 variant that passes BigCrush (Salmon et al., SC'11, table 2); dropout
// masks need no more, and the generator is a third of the fused node kernels' instruction stream at large batch.
constexpr int kPhiloxRounds = 7;
__device__ __forceinline__ uint4 philox4x32_inl(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < kPhiloxRounds; ++r) {
        uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}
// out-of-line copy for the latency-bound kernels (keeps their code small)
static __device__ __noinline__ uint4 philox4x32(uint4 ctr, uint2 key) { return philox4x32_inl(ctr, key); }

// keep-decision of dropout for one element. idx = global element index (sample
// index already offset by the rank's first sample, so masks do not depend on the
// world size).  Same function in forward and backward.
__device__ __forceinline__ bool philox_keep(const unsigned long long* rng_state, uint32_t uid,
                                            unsigned long long idx, float p) {
    unsigned long long seed = rng_state[0], step = rng_state[1];
    uint2 key = make_uint2((uint32_t)seed ^ (uid * 0x9E3779B1u), (uint32_t)(seed >> 32) + uid);
    uint4 ctr = make_uint4((uint32_t)(idx >> 2), (uint32_t)(idx >> 34), (uint32_t)step, (uint32_t)(step >> 32));
    uint4 r = philox4x32(ctr, key);
    uint32_t sel = (uint32_t)idx & 3u;
    uint32_t bits = sel == 0 ? r.x : sel == 1 ? r.y : sel == 2 ? r.z : r.w;
    float u = (float)(bits >> 8) * (1.0f / 16777216.0f);  // [0,1)
    return u >= p;
}

// dropout scale for one element: 0 (dropped) or 1/(1-p) (kept); 1 if inactive.
__device__ __forceinline__ float drop_scale(bool active, const unsigned char* mask, const unsigned long long* rng,
                                            uint32_t uid, long long local_idx, unsigned long long global_idx,
                                            float p) {
    if (!active) return 1.f;
    bool keep = mask ? (mask[local_idx] != 0) : philox_keep(rng, uid, global_idx, p);
    return keep ? 1.f / (1.f - p) : 0.f;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float softplusf_(float x) { return x > 20.f ? x : log1pf(expf(x)); }
static __device__ __noinline__ float mishf_(float x) { return x * tanhf(softplusf_(x)); }
static __device__ __noinline__ float mish_grad(float x) {
    float sp = softplusf_(x), t = tanhf(sp);
    return t + x * sigmoidf_(x) * (1.f - t * t);
}

}  // namespace bmnas
