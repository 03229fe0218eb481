// Adaptive max pooling of a raw backbone feature map onto the (C_in, L) grid the fusion cells work on: the first
// stage of ReshapeInputLayer / ReshapeInputLayer_MMIMDB (models/auxiliary/aux_models.py:61-69, 102-110).
//   x (B, C_in, H, W) contiguous  ->  out (B, C_in, OH*OW),   bin (i, j) = rows [floor(i*H/OH), ceil((i+1)*H/OH))
//                                                                       x cols [floor(j*W/OW), ceil((j+1)*W/OW))
// (the ATen adaptive-pooling rule: bins overlap when H % OH != 0 and repeat when H < OH).  NaN propagates and ties
// keep the first index, as torch does.  The F.interpolate(size=L) that follows in the reference is the identity
// (nearest neighbour onto the same length) and is not materialised.
// HBM bound: one read of x, OH*OW/(H*W) of it written.  Feature maps (window of tens to a thousand elements per
// bin, contiguous when OW == 1): one warp per bin, lanes stride the window with coalesced loads, warp arg-max by
// shuffle.  Vectors and tiny maps (window <= 8 elements, e.g. the pooled (B, C_in) features that are merely
// replicated over L): one thread per bin.
#include "common.cuh"

namespace bmnas {

constexpr int kPoolThreads = 256;

__device__ __forceinline__ void argmax_merge(float& v, int& i, float ov, int oi) {
    // NaN wins; otherwise larger value; ties -> smaller index (first occurrence in row-major order)
    const bool take = (ov != ov && !(v != v)) || (!(v != v) && (ov > v || (ov == v && oi < i))) || ((ov != ov) && (v != v) && oi < i);
    if (take) {
        v = ov;
        i = oi;
    }
}

__global__ void __launch_bounds__(kPoolThreads) k_pool_fwd(const bmnas_pool_params p) {
    pdl_prologue();
    const int lane = threadIdx.x & 31;
    const long long n_bins = (long long)p.B * p.C * p.OH * p.OW;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int HW = p.H * p.W;
    for (long long bin = warp0; bin < n_bins; bin += nwarps) {
        const int j = (int)(bin % p.OW), i = (int)((bin / p.OW) % p.OH);
        const long long plane = bin / ((long long)p.OW * p.OH);
        const int h0 = (i * p.H) / p.OH, h1 = ((i + 1) * p.H + p.OH - 1) / p.OH;
        const int w0 = (j * p.W) / p.OW, w1 = ((j + 1) * p.W + p.OW - 1) / p.OW;
        const int ww = w1 - w0, cnt = (h1 - h0) * ww;
        const float* src = p.x + plane * HW;
        float best = -INFINITY;
        int bi = h0 * p.W + w0;
        bool first = true;
        for (int t = lane; t < cnt; t += 32) {
            const int hh = h0 + t / ww, wc = w0 + t % ww, idx = hh * p.W + wc;
            const float v = __ldg(src + idx);
            if (first) {
                best = v;
                bi = idx;
                first = false;
            } else {
                argmax_merge(best, bi, v, idx);
            }
        }
        if (first) bi = 0x7fffffff;          // lanes without an element never win a tie
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            const bool ofirst = __shfl_xor_sync(0xffffffffu, (int)first, o) != 0;
            if (!ofirst) {
                if (first) {
                    best = ov;
                    bi = oi;
                    first = false;
                } else {
                    argmax_merge(best, bi, ov, oi);
                }
            }
        }
        if (lane == 0) {
            p.out[bin] = best;
            if (p.argmax) p.argmax[bin] = bi;
        }
    }
}

__global__ void __launch_bounds__(kPoolThreads) k_pool_fwd_small(const bmnas_pool_params p) {
    pdl_prologue();
    const long long n_bins = (long long)p.B * p.C * p.OH * p.OW;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const int HW = p.H * p.W;
    for (long long bin = (long long)blockIdx.x * blockDim.x + threadIdx.x; bin < n_bins; bin += stride) {
        const int j = (int)(bin % p.OW), i = (int)((bin / p.OW) % p.OH);
        const long long plane = bin / ((long long)p.OW * p.OH);
        const int h0 = (i * p.H) / p.OH, h1 = ((i + 1) * p.H + p.OH - 1) / p.OH;
        const int w0 = (j * p.W) / p.OW, w1 = ((j + 1) * p.W + p.OW - 1) / p.OW;
        const float* src = p.x + plane * HW;
        float best = __ldg(src + h0 * p.W + w0);
        int bi = h0 * p.W + w0;
        for (int hh = h0; hh < h1; ++hh)
            for (int wc = w0; wc < w1; ++wc) argmax_merge(best, bi, __ldg(src + hh * p.W + wc), hh * p.W + wc);
        p.out[bin] = best;
        if (p.argmax) p.argmax[bin] = bi;
    }
}

// gx[b, c, h, w] = sum over the bins whose arg-max is (h, w) of gout[b, c, bin]: a gather, so it needs neither a
// zero fill nor atomics and is deterministic (bins overlap, several may elect the same element)
__global__ void __launch_bounds__(kPoolThreads) k_pool_bwd(const bmnas_pool_params p) {
    pdl_prologue();
    const int HW = p.H * p.W, NB = p.OH * p.OW;
    const long long total = (long long)p.B * p.C * HW;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const long long plane = e / HW;
        const int idx = (int)(e - plane * HW), h = idx / p.W, w = idx - h * p.W;
        // bins that contain (h, w): rows i with floor(i*H/OH) <= h < ceil((i+1)*H/OH)
        int i_lo = (int)(((long long)h * p.OH) / p.H), j_lo = (int)(((long long)w * p.OW) / p.W);
        while (i_lo > 0 && ((i_lo) * p.H + p.OH - 1) / p.OH > h) --i_lo;      // previous bin still covers h
        while (j_lo > 0 && ((j_lo) * p.W + p.OW - 1) / p.OW > w) --j_lo;
        float acc = 0.f;
        for (int i = i_lo; i < p.OH && (i * p.H) / p.OH <= h; ++i)
            for (int j = j_lo; j < p.OW && (j * p.W) / p.OW <= w; ++j) {
                const long long bin = plane * NB + (long long)i * p.OW + j;
                if (p.argmax[bin] == idx) acc += __ldg(p.gout + bin);
            }
        if (p.gx_accum) acc += p.gx[e];
        p.gx[e] = acc;
    }
}

static int pool_check(const bmnas_pool_params* p) {
    if (!p || p->B < 1 || p->C < 1 || p->H < 1 || p->W < 1 || p->OH < 1 || p->OW < 1) return BMNAS_EINVAL;
    if ((long long)p->H * p->W > 0x7fffffffLL) return BMNAS_EINVAL;
    return BMNAS_OK;
}

}  // namespace bmnas

using namespace bmnas;

extern "C" int bmnas_pool_fwd(const bmnas_pool_params* p, void* stream) {
    int e = pool_check(p);
    if (e) return e;
    if (!p->x || !p->out) return BMNAS_EINVAL;
    BMNAS_DRY_RETURN();
    const long long n_bins = (long long)p->B * p->C * p->OH * p->OW;
    const int win = ((p->H + p->OH - 1) / p->OH + 1) * ((p->W + p->OW - 1) / p->OW + 1);   // upper bound of a bin's window
    const bool small = win <= 8;
    long long blocks = (n_bins * (small ? 1 : 32) + kPoolThreads - 1) / kPoolThreads;
    if (blocks > (long long)kNumSMs * 16) blocks = (long long)kNumSMs * 16;
    if (small)
        launch_k(k_pool_fwd_small, (int)blocks, kPoolThreads, 0, (cudaStream_t)stream, *p);
    else
        launch_k(k_pool_fwd, (int)blocks, kPoolThreads, 0, (cudaStream_t)stream, *p);
    BMNAS_LAUNCH_CHECK();
    return BMNAS_OK;
}

extern "C" int bmnas_pool_bwd(const bmnas_pool_params* p, void* stream) {
    int e = pool_check(p);
    if (e) return e;
    if (!p->gout || !p->gx || !p->argmax) return BMNAS_EINVAL;
    BMNAS_DRY_RETURN();
    const long long total = (long long)p->B * p->C * p->H * p->W;
    long long blocks = (total + kPoolThreads - 1) / kPoolThreads;
    if (blocks > (long long)kNumSMs * 16) blocks = (long long)kNumSMs * 16;
    launch_k(k_pool_bwd, (int)blocks, kPoolThreads, 0, (cudaStream_t)stream, *p);
    BMNAS_LAUNCH_CHECK();
    return BMNAS_OK;
}
