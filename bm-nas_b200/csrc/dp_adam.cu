// Data-parallel optimiser step as ONE kernel over NVLink peer memory -- bmnas_dp_adam_step (include/bmnas_b200.h).
//
// replaces, for one gradient bucket of the batch-sharded search step,
//     ncclAllReduce(sum) of the bucket  ->  Adam on every rank (torch.optim.Adam x2: ntu_darts_searchable.py:42, 46-47;
//     nn.DataParallel's gradient reduce-add: ntu_darts_searchable.py:50-51)
// by   reduce-scatter (P2P loads)  ->  Adam on this rank's 1/world shard  ->  all-gather of the UPDATED parameters
// (P2P stores), so a gradient element crosses NVLink once in each direction, the optimiser does 1/world of the
// arithmetic per GPU and no gradient is ever written back.  Every rank runs the same launch inside its own captured
// step graph; ranks meet at two flag barriers in each other's signal pads (system-scope release / acquire):
//     A  "my gradients are complete and visible"   -> peers may read them
//     B  "my parameter pushes have landed"          -> the next forward may read its parameters
// Replicas stay bit-identical: every parameter element is computed exactly once, by its owner, and the same bits are
// stored into every replica.  The shard sum is taken in rank order 0..world-1 on every owner (deterministic).
#include "common.cuh"

namespace bmnas {
namespace dp {

constexpr int TH = 256;

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_peer4(const float* p) {          // peer data changes every step: no read-only cache
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

// WT: the world size as a compile-time constant (2 / 4 / 8; 0 = generic), U: float4 groups in flight per thread
template <int WT, int U>
__global__ void __launch_bounds__(TH) k_dp_adam(const bmnas_dp_adam_params p) {
    pdl_prologue();
    __shared__ float s_c[2];
    __shared__ unsigned int s_epoch;
    const int W = p.world, R = p.rank;
    // ---- barrier A: announce my gradients (written by earlier kernels of this stream; make them visible system-wide),
    //      then wait for every peer's announcement of this epoch
    if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile unsigned int*>(p.epoch) + 1u;
    __syncthreads();
    const unsigned int epoch = s_epoch;
    if (blockIdx.x == 0 && threadIdx.x < W) {
        __threadfence_system();
        st_release_sys(p.signal_ptrs[threadIdx.x] + p.signal_base + R, epoch);
    }
    if (threadIdx.x < W) {
        const unsigned int* flag = p.signal_ptrs[R] + p.signal_base + threadIdx.x;
        while ((int)(ld_acquire_sys(flag) - epoch) < 0) {}
    }
    if (threadIdx.x == 0) {
        const long long t0 = p.step[0];
        const double t = (double)(t0 + 1);
        const double bc1 = 1.0 - pow((double)p.beta1, t), bc2 = 1.0 - pow((double)p.beta2, t);
        const float lr = p.lr[p.lr_ring > 0 ? (int)(t0 % p.lr_ring) : 0];
        s_c[0] = (float)((double)lr / bc1);
        s_c[1] = (float)sqrt(bc2);
    }
    __syncthreads();
    const float step_size = s_c[0], bc2s = s_c[1], b1 = p.beta1, b2 = p.beta2;
    // ---- my shard [lo, hi) of the bucket, in float4 units (n % 4 == 0, shard boundaries on float4)
    const long long n4 = p.n / 4;
    const long long per = (n4 + W - 1) / W;
    const long long lo = (long long)R * per, hi = min(n4, lo + per);
    // every remote load is a ~2 us NVLink round trip: a thread keeps U float4 groups x W ranks in flight at once
    const long long stride = (long long)gridDim.x * TH;
    for (long long q0 = lo + (long long)blockIdx.x * TH + threadIdx.x; q0 < hi; q0 += stride * U) {
        float4 g[U];
#pragma unroll
        for (int u = 0; u < U; ++u) g[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (WT > 0) {
            float4 a[U][WT > 0 ? WT : 1];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long q = q0 + u * stride;
#pragma unroll
                for (int r = 0; r < WT; ++r) a[u][r] = q < hi ? ld_peer4(p.grad_ptrs[r] + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
#pragma unroll
                for (int r = 0; r < WT; ++r) {               // rank order: the same sum on whichever rank owns the element
                    g[u].x += a[u][r].x; g[u].y += a[u][r].y; g[u].z += a[u][r].z; g[u].w += a[u][r].w;
                }
            }
        } else {
            for (int u = 0; u < U; ++u) {
                const long long q = q0 + u * stride;
                if (q >= hi) break;
                for (int r = 0; r < W; ++r) {
                    const float4 a = ld_peer4(p.grad_ptrs[r] + 4 * q);
                    g[u].x += a.x; g[u].y += a.y; g[u].z += a.z; g[u].w += a.w;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long q = q0 + u * stride;
            if (q >= hi) break;
            float* pw = p.param_ptrs[R] + 4 * q;
            const float4 w = *reinterpret_cast<const float4*>(pw);
            const float4 m = *reinterpret_cast<const float4*>(p.m + 4 * (q - lo));
            const float4 v = *reinterpret_cast<const float4*>(p.v + 4 * (q - lo));
            float ww[4] = {w.x, w.y, w.z, w.w}, gg[4] = {g[u].x, g[u].y, g[u].z, g[u].w}, mm[4] = {m.x, m.y, m.z, m.w}, vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {                    // identical arithmetic to k_adam (loss_optim.cu)
                float gj = gg[j] * p.grad_scale;
                if (p.weight_decay != 0.f) gj = fmaf(p.weight_decay, ww[j], gj);
                mm[j] = mm[j] + (gj - mm[j]) * (1.f - b1);
                vv[j] = vv[j] * b2 + (1.f - b2) * gj * gj;
                const float denom = sqrtf(vv[j]) / bc2s + p.eps;
                ww[j] = ww[j] - step_size * (mm[j] / denom);
            }
            *reinterpret_cast<float4*>(p.m + 4 * (q - lo)) = make_float4(mm[0], mm[1], mm[2], mm[3]);
            *reinterpret_cast<float4*>(p.v + 4 * (q - lo)) = make_float4(vv[0], vv[1], vv[2], vv[3]);
            const float4 nw = make_float4(ww[0], ww[1], ww[2], ww[3]);
            for (int r = 0; r < W; ++r) *reinterpret_cast<float4*>(p.param_ptrs[r] + 4 * q) = nw;    // all-gather by push
        }
    }
    // ---- barrier B: the last CTA of this rank announces that all its pushes are out; everyone waits for all ranks
    __threadfence_system();
    __syncthreads();
    __shared__ int s_last;
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(p.done_counter, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        if (threadIdx.x < W) st_release_sys(p.signal_ptrs[threadIdx.x] + p.signal_base + W + R, epoch);
        if (threadIdx.x == 0) {
            *p.done_counter = 0u;
            *p.epoch = epoch;
            p.step[0] += 1;
        }
    }
    if (threadIdx.x < W) {
        const unsigned int* flag = p.signal_ptrs[R] + p.signal_base + W + threadIdx.x;
        while ((int)(ld_acquire_sys(flag) - epoch) < 0) {}
    }
}

}  // namespace dp
}  // namespace bmnas

using namespace bmnas;

extern "C" int bmnas_dp_adam_step(const bmnas_dp_adam_params* p, void* stream) {
    if (!p || p->world < 1 || p->world > 16 || p->rank < 0 || p->rank >= p->world || p->n < 4 || (p->n & 3)) return BMNAS_EINVAL;
    if (!p->grad_ptrs || !p->param_ptrs || !p->signal_ptrs || !p->m || !p->v || !p->lr || !p->step || !p->epoch || !p->done_counter)
        return BMNAS_EINVAL;
    BMNAS_DRY_RETURN();
    const long long n4 = p->n / 4, per = (n4 + p->world - 1) / p->world;
    // the shard is a few hundred KB and every pass over it costs an NVLink round trip: spread it over the whole machine
    // (nothing else runs at this point of the step) so that one pass of U groups per thread covers it
    const int U = p->world <= 2 ? 4 : (p->world <= 4 ? 2 : 1);
    long long blocks = (per + (long long)dp::TH * U - 1) / ((long long)dp::TH * U);
    if (blocks < 1) blocks = 1;
    if (blocks > kNumSMs) blocks = kNumSMs;
    const dim3 grid((unsigned)blocks);
    cudaStream_t st = (cudaStream_t)stream;
    if (p->world == 2) launch_k(dp::k_dp_adam<2, 4>, grid, dp::TH, 0, st, *p);
    else if (p->world == 4) launch_k(dp::k_dp_adam<4, 2>, grid, dp::TH, 0, st, *p);
    else if (p->world == 8) launch_k(dp::k_dp_adam<8, 1>, grid, dp::TH, 0, st, *p);
    else launch_k(dp::k_dp_adam<0, 1>, grid, dp::TH, 0, st, *p);
    BMNAS_LAUNCH_CHECK();
    return BMNAS_OK;
}
