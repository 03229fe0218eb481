// Helpers shared by the FFMA (gemm.cu) and tcgen05 (gemm_tc.cu) GEMM kernels: stacked-weight /
// virtual-concat addressing, the BatchNorm-backward operand fold, Welford/Chan statistics merge and the
// "last CTA of the row tile" BatchNorm finalize.
#pragma once
#include "common.cuh"

namespace bmnas {

// row m of the stacked weight: pointer to W_seg[m_local][0]; also returns segment/local index
__device__ __forceinline__ const float* w_row(const bmnas_conv_params& p, int m, int ldw, int* seg, int* ml) {
    int s = 0;
    while (s + 1 < p.n_seg && m >= p.seg_M[s]) {
        m -= p.seg_M[s];
        ++s;
    }
    if (seg) *seg = s;
    if (ml) *ml = m;
    return p.W[s] + (long long)m * ldw;
}

// channel k of the virtual concat -> (source, local channel)
__device__ __forceinline__ void src_of(const bmnas_conv_params& p, int k, int* s, int* kl) {
    int i = 0;
    while (i + 1 < p.n_src && k >= p.src_C[i]) {
        k -= p.src_C[i];
        ++i;
    }
    *s = i;
    *kl = k;
}

struct Wf {  // Welford triple
    float n, mean, m2;
};
__device__ __forceinline__ Wf wf_merge(Wf a, Wf b) {
    Wf r;
    r.n = a.n + b.n;
    if (r.n <= 0.f) {
        r.mean = 0.f;
        r.m2 = 0.f;
        return r;
    }
    const float d = b.mean - a.mean, inv = 1.f / r.n;
    r.mean = a.mean + d * (b.n * inv);
    r.m2 = a.m2 + b.m2 + d * d * (a.n * b.n * inv);
    return r;
}

// upstream-gradient operand with BatchNorm backward folded in (scalar / float4)
__device__ __forceinline__ float dz1(const bmnas_conv_params& p, long long idx, int m) {
    float g = __ldg(p.GV + idx);
    if (p.coef_a) g = fmaf(__ldg(p.coef_a + m), g, fmaf(__ldg(p.coef_b + m), __ldg(p.Z + idx), __ldg(p.coef_c + m)));
    return g;
}
__device__ __forceinline__ float4 dz4(const bmnas_conv_params& p, long long idx, int m) {
    float4 g = __ldg(reinterpret_cast<const float4*>(p.GV + idx));
    if (p.coef_a) {
        const float a = __ldg(p.coef_a + m), b = __ldg(p.coef_b + m), c = __ldg(p.coef_c + m);
        const float4 z = __ldg(reinterpret_cast<const float4*>(p.Z + idx));
        g.x = fmaf(a, g.x, fmaf(b, z.x, c));
        g.y = fmaf(a, g.y, fmaf(b, z.y, c));
        g.z = fmaf(a, g.z, fmaf(b, z.z, c));
        g.w = fmaf(a, g.w, fmaf(b, z.w, c));
    }
    return g;
}

// eval-mode BatchNorm: statistics come from the running buffers (row m of the stacked output)
__device__ __forceinline__ void bn_eval_stats(const bmnas_conv_params& p, int m, int ldw) {
    if (m >= p.M) return;
    int s, ml;
    w_row(p, m, ldw, &s, &ml);
    p.mean[m] = p.running_mean[s][ml];
    p.rstd[m] = 1.f / sqrtf(p.running_var[s][ml] + p.eps);
}

// Train-mode BatchNorm finalize for rows [m0, m0+rows): merge the per-CTA (mean, M2) partials in stat_part
// ([part][M][2]; part t covers cnt_of(t) columns) with Chan's parallel update in a fixed order, write
// mean / rstd, update the running statistics (momentum, unbiased variance) and num_batches_tracked.
// Called by every thread of ONE CTA.  All rows are handled in ONE pass: blockDim.x / rows lanes per row walk
// interleaved partials (12 independent loads in flight per lane) and finish with a lane-symmetric butterfly,
// and the running statistics are fetched while the partials are still in flight.
template <class CntFn>
__device__ __forceinline__ void bn_finalize_rows(const bmnas_conv_params& p, int N, int n_parts, CntFn cnt_of, int m0,
                                                 int rows, int ldw, int nthr = 0) {
    const int tid = threadIdx.x;
    const int M = p.M;
    // nthr: threads that take part (whole warps; default the CTA) -- CTAs whose size / rows is not a power of two pass one
    int tpr = (nthr > 0 ? nthr : (int)blockDim.x) / rows;   // lanes per row: power of two, <= 32 (256 / 128 = 2, 256 / 32 = 8)
    if (tpr > 32) tpr = 32;
    const int r = tid / tpr, q = tid - r * tpr;
    const int m = m0 + r;
    const bool ok = r < rows && m < M;
    int s = 0, ml = 0;
    float rm_old = 0.f, rv_old = 0.f;
    bool has_run = false;
    if (ok && q == 0) {
        w_row(p, m, ldw, &s, &ml);
        has_run = p.running_mean[s] != nullptr;
        if (has_run) {
            rm_old = __ldcg(p.running_mean[s] + ml);
            rv_old = __ldcg(p.running_var[s] + ml);
        }
    }
    Wf w = {0.f, 0.f, 0.f};
    if (ok) {
        constexpr int FB = 12;                   // independent loads in flight per lane
        for (int t0 = q; t0 < n_parts; t0 += tpr * FB) {
            float2 v[FB];
#pragma unroll
            for (int j = 0; j < FB; ++j) {
                const int tix = t0 + tpr * j;
                v[j] = tix < n_parts ? __ldcg(reinterpret_cast<const float2*>(p.stat_part + ((long long)tix * M + m) * 2))
                                     : make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int j = 0; j < FB; ++j) {
                const int tix = t0 + tpr * j;
                if (tix < n_parts) {
                    Wf b = {(float)cnt_of(tix), v[j].x, v[j].y};
                    w = wf_merge(w, b);
                }
            }
        }
    }
    for (int o = 1; o < tpr; o <<= 1) {
        Wf b;
        b.n = __shfl_xor_sync(0xffffffffu, w.n, o);
        b.mean = __shfl_xor_sync(0xffffffffu, w.mean, o);
        b.m2 = __shfl_xor_sync(0xffffffffu, w.m2, o);
        w = ((q & o) == 0) ? wf_merge(w, b) : wf_merge(b, w);  // same operand order in both lanes
    }
    if (ok && q == 0) {
        const float var = w.m2 / (float)N;
        p.mean[m] = w.mean;
        p.rstd[m] = 1.f / sqrtf(var + p.eps);
        if (has_run) {
            const float unb = w.m2 / (float)max(N - 1, 1);
            p.running_mean[s][ml] = (1.f - p.momentum) * rm_old + p.momentum * w.mean;
            p.running_var[s][ml] = (1.f - p.momentum) * rv_old + p.momentum * unb;
            if (ml == 0 && p.num_batches_tracked[s]) *p.num_batches_tracked[s] += 1;
        }
    }
}

}  // namespace bmnas
