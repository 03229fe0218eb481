// Edge mix (softmax-weighted sum of candidate input tensors) forward/backward.
// Pure streaming kernels: float4 loads, one pass, grid sized in multiples of the
// SM count; d(alpha) dot products reduced warp-shuffle -> block -> fixed-order
// last-block sum (deterministic).
#include "common.cuh"

namespace bmnas {

constexpr int kMixThreads = 256;
constexpr int kMixMaxBlocks = kNumSMs * 4;

__device__ __forceinline__ void mix_weights(const bmnas_mix_params& p, float* ws, float* wn) {
    if (threadIdx.x < p.n) {
        float a = p.w[2 * threadIdx.x], b = p.w[2 * threadIdx.x + 1];
        if (p.w_is_logits) {
            float s = 1.f / (1.f + expf(a - b));  // softmax over (none, skip) -> skip weight
            ws[threadIdx.x] = s;
            wn[threadIdx.x] = 1.f - s;
        } else {
            ws[threadIdx.x] = b;
            wn[threadIdx.x] = a;
        }
    }
    __syncthreads();
}

// s2 = sum of the skip weights of the chained mix (0 when there is none); uniform, a handful of scalar loads
__device__ __forceinline__ float mix_chain_scale(const bmnas_mix_params& p) {
    float s2 = 0.f;
    if (p.w2) {
#pragma unroll 1
        for (int j = 0; j < p.n2; ++j) {
            const float a = __ldg(p.w2 + 2 * j), b = __ldg(p.w2 + 2 * j + 1);
            s2 += p.w_is_logits ? 1.f / (1.f + expf(a - b)) : b;
        }
    }
    return s2;
}

template <int VEC>
__global__ void __launch_bounds__(kMixThreads) k_mix_fwd(const bmnas_mix_params p) {
    pdl_prologue();
    __shared__ float ws[BMNAS_MAX_MIX], wn[BMNAS_MAX_MIX];
    mix_weights(p, ws, wn);
    const float s2 = mix_chain_scale(p);
    const long long nvec = p.numel / VEC;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        if (VEC == 4) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int j = 0; j < p.n; ++j) {
                float4 v = __ldg(reinterpret_cast<const float4*>(p.x[j]) + i);
                float w = ws[j];
                acc.x = fmaf(w, v.x, acc.x);
                acc.y = fmaf(w, v.y, acc.y);
                acc.z = fmaf(w, v.z, acc.z);
                acc.w = fmaf(w, v.w, acc.w);
            }
            reinterpret_cast<float4*>(p.out)[i] = acc;
            if (p.out2) reinterpret_cast<float4*>(p.out2)[i] = make_float4(s2 * acc.x, s2 * acc.y, s2 * acc.z, s2 * acc.w);
        } else {
            float acc = 0.f;
            for (int j = 0; j < p.n; ++j) acc = fmaf(ws[j], __ldg(p.x[j] + i), acc);
            p.out[i] = acc;
            if (p.out2) p.out2[i] = s2 * acc;
        }
    }
}

template <int VEC>
__global__ void __launch_bounds__(kMixThreads) k_mix_bwd(const bmnas_mix_params p) {
    pdl_prologue();
    __shared__ float ws[BMNAS_MAX_MIX], wn[BMNAS_MAX_MIX];
    __shared__ float red[BMNAS_MAX_MIX * 32];
    mix_weights(p, ws, wn);
    const float s2 = mix_chain_scale(p);
    float dot[BMNAS_MAX_MIX];
#pragma unroll
    for (int j = 0; j < BMNAS_MAX_MIX; ++j) dot[j] = 0.f;
    const long long nvec = p.numel / VEC;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        if (VEC == 4) {
            float4 g = p.gout ? __ldg(reinterpret_cast<const float4*>(p.gout) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.gout2) {
                const float4 h = __ldg(reinterpret_cast<const float4*>(p.gout2) + i);
                g.x = fmaf(s2, h.x, g.x); g.y = fmaf(s2, h.y, g.y); g.z = fmaf(s2, h.z, g.z); g.w = fmaf(s2, h.w, g.w);
            }
#pragma unroll
            for (int j = 0; j < BMNAS_MAX_MIX; ++j) {
                if (j < p.n) {
                    if (p.gw) {
                        float4 v = __ldg(reinterpret_cast<const float4*>(p.x[j]) + i);
                        dot[j] += g.x * v.x + g.y * v.y + g.z * v.z + g.w * v.w;
                    }
                    if (p.gx[j]) {
                        float w = ws[j];
                        float4* d = reinterpret_cast<float4*>(p.gx[j]) + i;
                        float4 o = make_float4(w * g.x, w * g.y, w * g.z, w * g.w);
                        if (p.gx_accum[j]) {
                            float4 c = *d;
                            o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w;
                        }
                        *d = o;
                    }
                }
            }
        } else {
            float g = p.gout ? __ldg(p.gout + i) : 0.f;
            if (p.gout2) g = fmaf(s2, __ldg(p.gout2 + i), g);
#pragma unroll
            for (int j = 0; j < BMNAS_MAX_MIX; ++j) {
                if (j < p.n) {
                    if (p.gw) dot[j] += g * __ldg(p.x[j] + i);
                    if (p.gx[j]) {
                        float o = ws[j] * g;
                        if (p.gx_accum[j]) o += p.gx[j][i];
                        p.gx[j][i] = o;
                    }
                }
            }
        }
    }
    if (!p.gw) return;
    block_sum<BMNAS_MAX_MIX>(dot, red);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int j = 0; j < BMNAS_MAX_MIX; ++j)
            if (j < p.n) p.partials[(long long)blockIdx.x * p.n + j] = dot[j];
    }
    if (last_block(p.counter, gridDim.x)) {
        // fixed-order parallel reduction over the blocks: one warp per edge, lanes stride over blocks
        const int lane = threadIdx.x & 31;
        for (int j = threadIdx.x >> 5; j < p.n; j += kMixThreads / 32) {
            float d = 0.f;
            for (unsigned b = lane; b < gridDim.x; b += 32) d += ld_cg(p.partials + (long long)b * p.n + j);
            d = warp_sum(d);
            if (lane == 0) red[j] = d;
        }
        __syncthreads();
        if (threadIdx.x < p.n) {
            const int j = threadIdx.x;
            const float d = red[j];
            if (p.w_is_logits) {
                // 2-way softmax backward; dL/dw_none == 0 for finite inputs (Zero op: x*0)
                float gs = ws[j] * wn[j] * d;
                p.gw[2 * j] = -gs;
                p.gw[2 * j + 1] = gs;
            } else {
                p.gw[2 * j] = 0.f;
                p.gw[2 * j + 1] = d;
            }
        }
    }
}

static bool aligned16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }

static int mix_blocks(long long nvec) {
    long long b = (nvec + kMixThreads - 1) / kMixThreads;
    if (b < 1) b = 1;
    if (b > kMixMaxBlocks) b = kMixMaxBlocks;
    return (int)b;
}

static int mix_vec(const bmnas_mix_params* p, bool bwd) {
    if (p->numel % 4) return 1;
    for (int j = 0; j < p->n; ++j) {
        if (!aligned16(p->x[j])) return 1;
        if (bwd && p->gx[j] && !aligned16(p->gx[j])) return 1;
    }
    if (!bwd && (!aligned16(p->out) || !aligned16(p->out2))) return 1;
    if (bwd && (!aligned16(p->gout) || !aligned16(p->gout2))) return 1;
    return 4;
}

}  // namespace bmnas

using namespace bmnas;

extern "C" long long bmnas_mix_partials_size(const bmnas_mix_params* p) {
    (void)p;
    return (long long)kMixMaxBlocks * BMNAS_MAX_MIX;
}

extern "C" int bmnas_mix_fwd(const bmnas_mix_params* p, void* stream) {
    if (!p || p->n < 1 || p->n > BMNAS_MAX_MIX || p->numel <= 0 || !p->w || !p->out) return BMNAS_EINVAL;
    if (p->out2 && (!p->w2 || p->n2 < 1)) return BMNAS_EINVAL;
    for (int j = 0; j < p->n; ++j)
        if (!p->x[j]) return BMNAS_EINVAL;
    BMNAS_DRY_RETURN();
    cudaStream_t s = (cudaStream_t)stream;
    const int vec = mix_vec(p, false);
    const int blocks = mix_blocks(p->numel / vec);
    if (vec == 4)
        launch_k(k_mix_fwd<4>, blocks, kMixThreads, 0, s, *p);
    else
        launch_k(k_mix_fwd<1>, blocks, kMixThreads, 0, s, *p);
    BMNAS_LAUNCH_CHECK();
    return BMNAS_OK;
}

extern "C" int bmnas_mix_bwd(const bmnas_mix_params* p, void* stream) {
    if (!p || p->n < 1 || p->n > BMNAS_MAX_MIX || p->numel <= 0 || !p->w || (!p->gout && !p->gout2)) return BMNAS_EINVAL;
    if (p->gw && (!p->partials || !p->counter)) return BMNAS_EINVAL;
    if (p->gout2 && (!p->w2 || p->n2 < 1)) return BMNAS_EINVAL;
    for (int j = 0; j < p->n; ++j)
        if (!p->x[j]) return BMNAS_EINVAL;
    BMNAS_DRY_RETURN();
    cudaStream_t s = (cudaStream_t)stream;
    const int vec = mix_vec(p, true);
    const int blocks = mix_blocks(p->numel / vec);
    if (vec == 4)
        launch_k(k_mix_bwd<4>, blocks, kMixThreads, 0, s, *p);
    else
        launch_k(k_mix_bwd<1>, blocks, kMixThreads, 0, s, *p);
    BMNAS_LAUNCH_CHECK();
    return BMNAS_OK;
}
