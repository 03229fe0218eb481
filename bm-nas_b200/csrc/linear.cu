// Classifier head nn.Linear (central_classifier, ntu_darts_searchable.py:100-101): out = x W^T + bias with
// x (B, K) = the flattened fusion-cell output (K = C*L*multiplier, 2048 on NTU), W (N, K), N = classes (60).
// 23.6 MFLOP at the reference batch: three latency-bound skinny GEMMs, each a single launch in which every
// operand element is read exactly once per CTA with coalesced 128-bit loads and nothing is staged twice:
//   fwd  out[b][c] = bias[c] + sum_k x[b][k] W[c][k]     CTA = 12 samples x 5 classes, threads split k, block reduce
//   dX   gx[b][k]  = sum_c g[b][c] W[c][k]                CTA = 8 samples x 512 columns, thread = one float4 column
//   dW   gW[c][k]  = sum_b g[b][c] x[b][k], gb[c] = sum_b g[b][c]
//                                                         CTA = 128 columns x 15 classes, warp = 2 classes, lane = float4 column,
//                                                         gout staged 512 samples at a time
// No atomics, deterministic, results overwrite their destination (the gradient arena views).
#include "common.cuh"

namespace bmnas {
namespace lin {

constexpr int FB = 12, FC = 5, FT = 256;      // fwd: samples x classes per CTA, threads
constexpr int XB = 8, XT = 128;               // dX: samples per CTA, threads (one float4 column each)
constexpr int WC = 15, WT = 256, WCOLS = 128; // dW: classes per CTA, threads, columns per CTA

__global__ void __launch_bounds__(FT) k_lin_fwd(const bmnas_linear_params p) {
    pdl_prologue();
    __shared__ float red[FT / 32][FB * FC];
    const int b0 = blockIdx.x * FB, c0 = blockIdx.y * FC;
    const int K4 = p.K >> 2, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float acc[FB][FC];
#pragma unroll
    for (int i = 0; i < FB; ++i)
#pragma unroll
        for (int j = 0; j < FC; ++j) acc[i][j] = 0.f;
    for (int k4 = tid; k4 < K4; k4 += FT) {
        float4 xv[FB], wv[FC];
#pragma unroll
        for (int i = 0; i < FB; ++i)
            xv[i] = (b0 + i < p.B) ? __ldg(reinterpret_cast<const float4*>(p.x + (long long)(b0 + i) * p.K) + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < FC; ++j)
            wv[j] = (c0 + j < p.N) ? __ldg(reinterpret_cast<const float4*>(p.W + (long long)(c0 + j) * p.K) + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < FB; ++i)
#pragma unroll
            for (int j = 0; j < FC; ++j)
                acc[i][j] += (xv[i].x * wv[j].x + xv[i].y * wv[j].y) + (xv[i].z * wv[j].z + xv[i].w * wv[j].w);
    }
#pragma unroll
    for (int i = 0; i < FB; ++i)
#pragma unroll
        for (int j = 0; j < FC; ++j) {
            const float v = warp_sum(acc[i][j]);
            if (lane == 0) red[warp][i * FC + j] = v;
        }
    __syncthreads();
    if (tid < FB * FC) {
        const int i = tid / FC, j = tid - i * FC;
        if (b0 + i < p.B && c0 + j < p.N) {
            float v = p.bias ? __ldg(p.bias + c0 + j) : 0.f;
#pragma unroll
            for (int w = 0; w < FT / 32; ++w) v += red[w][tid];
            p.out[(long long)(b0 + i) * p.N + c0 + j] = v;
        }
    }
}

// gx[b][k4] for XB samples; gout tile (XB x N) in shared memory
__global__ void __launch_bounds__(XT) k_lin_dx(const bmnas_linear_params p) {
    pdl_prologue();
    extern __shared__ float gs[];   // [XB][N]
    const int b0 = blockIdx.y * XB, tid = threadIdx.x;
    const int K4 = p.K >> 2, k4 = blockIdx.x * XT + tid;
    for (int u = tid; u < XB * p.N; u += XT) {
        const int i = u / p.N, c = u - i * p.N;
        gs[u] = (b0 + i < p.B) ? __ldg(p.gout + (long long)(b0 + i) * p.N + c) : 0.f;
    }
    __syncthreads();
    if (k4 >= K4) return;
    float4 acc[XB];
#pragma unroll
    for (int i = 0; i < XB; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* wp = reinterpret_cast<const float4*>(p.W) + k4;
#pragma unroll 4
    for (int c = 0; c < p.N; ++c) {
        const float4 w = __ldg(wp + (long long)c * K4);
#pragma unroll
        for (int i = 0; i < XB; ++i) {
            const float g = gs[i * p.N + c];
            acc[i].x = fmaf(g, w.x, acc[i].x); acc[i].y = fmaf(g, w.y, acc[i].y);
            acc[i].z = fmaf(g, w.z, acc[i].z); acc[i].w = fmaf(g, w.w, acc[i].w);
        }
    }
#pragma unroll
    for (int i = 0; i < XB; ++i)
        if (b0 + i < p.B) reinterpret_cast<float4*>(p.gx + (long long)(b0 + i) * p.K)[k4] = acc[i];
}

// gW[c][k4] for WC classes x 128 columns; gout (B x WC slice) staged in shared memory WB samples at a time;
// gbias by the first column tile
constexpr int WB = 512;
__global__ void __launch_bounds__(WT) k_lin_dw(const bmnas_linear_params p) {
    pdl_prologue();
    __shared__ float gs[WB * WC];   // [sample][WC]
    const int c0 = blockIdx.y * WC, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K4 = p.K >> 2, k4 = blockIdx.x * (WCOLS / 4) + lane;
    // warp w owns classes c0 + 2w, c0 + 2w + 1 (8 warps x 2 >= 15)
    const int j0 = 2 * warp, j1 = 2 * warp + 1;
    const bool has0 = j0 < WC, has1 = j1 < WC;
    const bool work = p.gW && k4 < K4 && has0;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    float bsum = 0.f;
    const float4* xp = reinterpret_cast<const float4*>(p.x) + k4;
    for (int bb = 0; bb < p.B; bb += WB) {
        const int nb = min(WB, p.B - bb);
        __syncthreads();
        for (int u = tid; u < nb * WC; u += WT) {
            const int b = u / WC, j = u - b * WC;
            gs[u] = (c0 + j < p.N) ? __ldg(p.gout + (long long)(bb + b) * p.N + c0 + j) : 0.f;
        }
        __syncthreads();
        if (work) {
#pragma unroll 4
            for (int b = 0; b < nb; ++b) {
                const float4 x = __ldg(xp + (long long)(bb + b) * K4);
                const float g0 = gs[b * WC + j0], g1 = has1 ? gs[b * WC + j1] : 0.f;
                a0.x = fmaf(g0, x.x, a0.x); a0.y = fmaf(g0, x.y, a0.y); a0.z = fmaf(g0, x.z, a0.z); a0.w = fmaf(g0, x.w, a0.w);
                a1.x = fmaf(g1, x.x, a1.x); a1.y = fmaf(g1, x.y, a1.y); a1.z = fmaf(g1, x.z, a1.z); a1.w = fmaf(g1, x.w, a1.w);
            }
        }
        if (p.gbias && blockIdx.x == 0 && tid < WC)
            for (int b = 0; b < nb; ++b) bsum += gs[b * WC + tid];
    }
    if (work) {
        if (c0 + j0 < p.N) reinterpret_cast<float4*>(p.gW + (long long)(c0 + j0) * p.K)[k4] = a0;
        if (has1 && c0 + j1 < p.N) reinterpret_cast<float4*>(p.gW + (long long)(c0 + j1) * p.K)[k4] = a1;
    }
    if (p.gbias && blockIdx.x == 0 && tid < WC && c0 + tid < p.N) p.gbias[c0 + tid] = bsum;
}

static bool al16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }

}  // namespace lin
}  // namespace bmnas

using namespace bmnas;

extern "C" int bmnas_linear_fwd(const bmnas_linear_params* p, void* stream) {
    using namespace lin;
    if (!p || p->B < 1 || p->K < 4 || (p->K & 3) || p->N < 1 || !p->x || !p->W || !p->out) return BMNAS_EINVAL;
    if (!al16(p->x) || !al16(p->W)) return BMNAS_EINVAL;
    BMNAS_DRY_RETURN();
    dim3 grid((p->B + FB - 1) / FB, (p->N + FC - 1) / FC);
    launch_k(k_lin_fwd, grid, FT, 0, (cudaStream_t)stream, *p);
    BMNAS_LAUNCH_CHECK();
    return BMNAS_OK;
}

extern "C" int bmnas_linear_bwd(const bmnas_linear_params* p, void* stream) {
    using namespace lin;
    if (!p || p->B < 1 || p->K < 4 || (p->K & 3) || p->N < 1 || !p->gout) return BMNAS_EINVAL;
    if (p->gx && (!p->W || !al16(p->W) || !al16(p->gx))) return BMNAS_EINVAL;
    if ((p->gW || p->gbias) && (!p->x || !al16(p->x) || (p->gW && !al16(p->gW)))) return BMNAS_EINVAL;
    if ((size_t)XB * p->N * 4 > 48 * 1024) return BMNAS_EINVAL;
    BMNAS_DRY_RETURN();
    const int K4 = p->K >> 2;
    if (p->gx) {
        dim3 grid((K4 + XT - 1) / XT, (p->B + XB - 1) / XB);
        launch_k(k_lin_dx, grid, XT, (size_t)XB * p->N * 4, (cudaStream_t)stream, *p);
        BMNAS_LAUNCH_CHECK();
    }
    if (p->gW || p->gbias) {
        dim3 grid((p->K + WCOLS - 1) / WCOLS, (p->N + WC - 1) / WC);
        launch_k(k_lin_dw, grid, WT, 0, (cudaStream_t)stream, *p);
        BMNAS_LAUNCH_CHECK();
    }
    return BMNAS_OK;
}
