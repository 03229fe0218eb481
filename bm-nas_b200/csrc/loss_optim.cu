// Loss head (mean CE / BCE-with-logits, fused softmax + gradient), classifier bias
// helpers, multi-tensor Adam with device-resident lr/step, dropout RNG step.
#include "common.cuh"

namespace bmnas {

constexpr int OTH = 256;
constexpr int kLossMaxBlocks = kNumSMs * 2;

// one warp per row (CE) -- row = sample
__global__ void __launch_bounds__(OTH) k_loss_fwd(const bmnas_loss_params p) {
    pdl_prologue();
    __shared__ float red[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = OTH / 32;
    const int n = p.n_classes;
    float lsum[1] = {0.f};
    if (p.kind == 0) {
        const float invB = 1.f / (float)p.B;
        for (int b = blockIdx.x * nw + warp; b < p.B; b += gridDim.x * nw) {
            const float* x = p.logits + (long long)b * n;
            float mx = -INFINITY;
            for (int j = lane; j < n; j += 32) mx = fmaxf(mx, x[j]);
            mx = warp_max(mx);
            float s = 0.f;
            for (int j = lane; j < n; j += 32) s += expf(x[j] - mx);
            s = warp_sum(s);
            const long long lab64 = p.labels[b];
            // a label outside [0, n_classes) (ignore_index = -100, corrupt data) must not become an out-of-bounds read and a
            // silently wrong loss: the loss turns NaN (torch asserts on the device; a captured graph cannot) and the row's
            // gradient is NaN too, so the first optimiser step makes the failure visible
            const bool bad = lab64 < 0 || lab64 >= (long long)n;
            const int lab = bad ? 0 : (int)lab64;
            const float lse = mx + logf(s);
            if (lane == 0) lsum[0] += bad ? __int_as_float(0x7fc00000) : lse - x[lab];
            const float inv = 1.f / s;
            for (int j = lane; j < n; j += 32)
                p.glogits[(long long)b * n + j] = bad ? __int_as_float(0x7fc00000) : (expf(x[j] - mx) * inv - (j == lab ? 1.f : 0.f)) * invB;
        }
    } else {
        const long long tot = (long long)p.B * n;
        const float inv = 1.f / (float)tot;
        for (long long i = (long long)blockIdx.x * OTH + threadIdx.x; i < tot; i += (long long)gridDim.x * OTH) {
            const float x = p.logits[i], t = p.targets[i];
            lsum[0] += fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));
            p.glogits[i] = (sigmoidf_(x) - t) * inv;
        }
    }
    block_sum<1>(lsum, red);
    if (threadIdx.x == 0) p.partials[blockIdx.x] = lsum[0];
    if (last_block(p.counter, gridDim.x)) {
        if (threadIdx.x == 0) {
            float s = 0.f;
            for (unsigned b = 0; b < gridDim.x; ++b) s += ld_cg(p.partials + b);
            p.loss[0] = s / (p.kind == 0 ? (float)p.B : (float)p.B * (float)n);
        }
    }
}

__global__ void __launch_bounds__(OTH) k_loss_bwd(const bmnas_loss_params p) {
    pdl_prologue();
    const long long tot = (long long)p.B * p.n_classes;
    const float sc = p.gscale ? p.gscale[0] : 1.f;
    for (long long i = (long long)blockIdx.x * OTH + threadIdx.x; i < tot; i += (long long)gridDim.x * OTH)
        p.gout_logits[i] = p.glogits[i] * sc;
}

__global__ void __launch_bounds__(OTH) k_bias_rows(float* out, const float* bias, long long tot, int n) {
    pdl_prologue();
    for (long long i = (long long)blockIdx.x * OTH + threadIdx.x; i < tot; i += (long long)gridDim.x * OTH)
        out[i] = bias ? bias[i % n] : 0.f;
}

// out[j] = sum_r in[r][j]; one block per column, fixed-order reduction
__global__ void __launch_bounds__(OTH) k_colsum(float* out, const float* in, int rows, int n) {
    pdl_prologue();
    __shared__ float red[32];
    const int j = blockIdx.x;
    float s[1] = {0.f};
    for (int r = threadIdx.x; r < rows; r += OTH) s[0] += in[(long long)r * n + j];
    block_sum<1>(s, red);
    if (threadIdx.x == 0) out[j] = s[0];
}

__global__ void __launch_bounds__(OTH) k_adam(const bmnas_adam_params p) {
    pdl_prologue();
    __shared__ float s_c[4];
    // locate this block's tensor (block_start is ascending)
    int lo = 0, hi = p.n_tensors - 1;
    const long long blk = blockIdx.x;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (p.tensors[mid].block_start <= blk) lo = mid;
        else hi = mid - 1;
    }
    const bmnas_adam_tensor T = p.tensors[lo];
    if (threadIdx.x == 0) {
        const long long t0 = p.step[0];
        const double t = (double)(t0 + 1);
        const double bc1 = 1.0 - pow((double)p.beta1, t), bc2 = 1.0 - pow((double)p.beta2, t);
        const float lr = p.lr[p.lr_ring > 0 ? (int)(t0 % p.lr_ring) : 0];
        s_c[0] = (float)((double)lr / bc1);  // step size
        s_c[1] = (float)sqrt(bc2);
    }
    __syncthreads();
    const float step_size = s_c[0], bc2s = s_c[1];
    const float b1 = p.beta1, b2 = p.beta2;
    const long long base = (blk - T.block_start) * p.block_elems;
    for (int i = threadIdx.x; i < p.block_elems; i += OTH) {
        const long long e = base + i;
        if (e >= T.n) break;
        float w = T.p[e];
        float g = T.g[e] * p.grad_scale;
        if (p.weight_decay != 0.f) g = fmaf(p.weight_decay, w, g);
        float m = T.m[e], v = T.v[e];
        m = m + (g - m) * (1.f - b1);                 // exp_avg.lerp_(grad, 1-beta1)
        v = v * b2 + (1.f - b2) * g * g;              // mul_(beta2).addcmul_(g, g, 1-beta2)
        const float denom = sqrtf(v) / bc2s + p.eps;
        w = w - step_size * (m / denom);
        T.p[e] = w;
        T.m[e] = m;
        T.v[e] = v;
    }
    if (last_block(p.counter, gridDim.x)) {
        if (threadIdx.x == 0) p.step[0] += 1;
    }
}

// zero fill as a KERNEL node: inside a captured graph a memset node sits on another engine and every edge into or out of
// it costs 5-7 us of dependency latency (tools/timeline.py: the first backward kernel waited 6.2 us behind a 1.1 us memset)
__global__ void __launch_bounds__(OTH) k_zero(uint4* p16, long long n16, unsigned char* tail, int ntail) {
    pdl_prologue();
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (long long i = (long long)blockIdx.x * OTH + threadIdx.x; i < n16; i += (long long)gridDim.x * OTH) p16[i] = z;
    if (blockIdx.x == 0 && (int)threadIdx.x < ntail) tail[threadIdx.x] = 0;
}

// device-to-device copy on the SMs (a copy-engine cudaMemcpyAsync moved the 6.3 MB of a step's inputs in 7.4 us; this takes ~2.5)
__global__ void __launch_bounds__(OTH) k_copy(uint4* dst, const uint4* src, long long n16) {
    pdl_prologue();
    const long long stride = (long long)gridDim.x * OTH;
    long long i = (long long)blockIdx.x * OTH + threadIdx.x;
    for (; i + 3 * stride < n16; i += 4 * stride) {          // four independent 16-byte loads in flight per thread
        const uint4 a = __ldcs(src + i), b = __ldcs(src + i + stride), c = __ldcs(src + i + 2 * stride), d = __ldcs(src + i + 3 * stride);
        dst[i] = a; dst[i + stride] = b; dst[i + 2 * stride] = c; dst[i + 3 * stride] = d;
    }
    for (; i < n16; i += stride) dst[i] = __ldcs(src + i);
}

__global__ void k_rng_advance(unsigned long long* st) {
    pdl_prologue();
    st[1] += 1ull;
}

// the keep decision the fused kernels draw for dropout site `uid` in this (seed, step): one byte per element
__global__ void k_philox_keep_mask(const unsigned long long* rng, uint32_t uid, float p, long long sample_offset,
                                   long long per_sample, long long total, unsigned char* out) {
    pdl_prologue();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
        out[i] = philox_keep(rng, uid, (unsigned long long)(sample_offset * per_sample + i), p) ? 1 : 0;
}

}  // namespace bmnas

using namespace bmnas;

static int grid_for(long long tot) {
    long long b = (tot + OTH - 1) / OTH;
    if (b < 1) b = 1;
    if (b > kNumSMs * 8) b = kNumSMs * 8;
    return (int)b;
}

extern "C" long long bmnas_loss_partials_size(const bmnas_loss_params* p) {
    (void)p;
    return kLossMaxBlocks;
}

extern "C" int bmnas_loss_fwd(const bmnas_loss_params* p, void* stream) {
    if (!p || p->B < 1 || p->n_classes < 1 || !p->logits || !p->loss || !p->glogits || !p->partials || !p->counter)
        return BMNAS_EINVAL;
    if (p->kind == 0 ? !p->labels : !p->targets) return BMNAS_EINVAL;
    if (p->kind != 0 && p->kind != 1) return BMNAS_EINVAL;
    BMNAS_DRY_RETURN();
    long long work = p->kind == 0 ? ((long long)p->B + 7) / 8 : ((long long)p->B * p->n_classes + OTH - 1) / OTH;
    int blocks = (int)(work < 1 ? 1 : (work > kLossMaxBlocks ? kLossMaxBlocks : work));
    launch_k(k_loss_fwd, blocks, OTH, 0, (cudaStream_t)stream, *p);
    BMNAS_LAUNCH_CHECK();
    return BMNAS_OK;
}

extern "C" int bmnas_loss_bwd(const bmnas_loss_params* p, void* stream) {
    if (!p || p->B < 1 || p->n_classes < 1 || !p->glogits || !p->gout_logits) return BMNAS_EINVAL;
    BMNAS_DRY_RETURN();
    launch_k(k_loss_bwd, grid_for((long long)p->B * p->n_classes), OTH, 0, (cudaStream_t)stream, *p);
    BMNAS_LAUNCH_CHECK();
    return BMNAS_OK;
}

extern "C" int bmnas_bias_rows(float* out, const float* bias, int rows, int n, void* stream) {
    if (!out || rows < 1 || n < 1) return BMNAS_EINVAL;
    BMNAS_DRY_RETURN();
    const long long tot = (long long)rows * n;
    launch_k(k_bias_rows, grid_for(tot), OTH, 0, (cudaStream_t)stream, out, bias, tot, n);
    BMNAS_LAUNCH_CHECK();
    return BMNAS_OK;
}

extern "C" int bmnas_colsum(float* out, const float* in, int rows, int n, void* stream) {
    if (!out || !in || rows < 1 || n < 1) return BMNAS_EINVAL;
    BMNAS_DRY_RETURN();
    launch_k(k_colsum, n, OTH, 0, (cudaStream_t)stream, out, in, rows, n);
    BMNAS_LAUNCH_CHECK();
    return BMNAS_OK;
}

extern "C" int bmnas_adam_step(const bmnas_adam_params* p, void* stream) {
    if (!p || p->n_tensors < 1 || p->total_blocks < 1 || p->block_elems < 1 || !p->tensors || !p->lr || !p->step ||
        !p->counter || p->lr_ring < 0)
        return BMNAS_EINVAL;
    if (p->total_blocks > 0x7fffffffLL) return BMNAS_EINVAL;
    BMNAS_DRY_RETURN();
    launch_k(k_adam, (unsigned)p->total_blocks, OTH, 0, (cudaStream_t)stream, *p);
    BMNAS_LAUNCH_CHECK();
    return BMNAS_OK;
}

extern "C" int bmnas_rng_advance(unsigned long long* rng_state, void* stream) {
    if (!rng_state) return BMNAS_EINVAL;
    BMNAS_DRY_RETURN();
    launch_k(k_rng_advance, 1, 1, 0, (cudaStream_t)stream, rng_state);
    BMNAS_LAUNCH_CHECK();
    return BMNAS_OK;
}

extern "C" int bmnas_philox_keep_mask(const unsigned long long* rng_state, unsigned int uid, float p, long long sample_offset,
                                      long long per_sample, long long B, unsigned char* out, void* stream) {
    if (!rng_state || !out || per_sample <= 0 || B <= 0 || !(p >= 0.f && p < 1.f)) return BMNAS_EINVAL;
    BMNAS_DRY_RETURN();
    const long long total = per_sample * B;
    launch_k(k_philox_keep_mask, grid_for(total), OTH, 0, (cudaStream_t)stream, rng_state, (uint32_t)uid, p, sample_offset,
             per_sample, total, out);
    BMNAS_LAUNCH_CHECK();
    return BMNAS_OK;
}

extern "C" const char* bmnas_strerror(int code) {
    switch (code) {
        case BMNAS_OK: return "ok";
        case BMNAS_EINVAL: return "invalid argument (bad shape, unsupported size or null pointer)";
        case BMNAS_ELAUNCH: return "CUDA launch failure";
        default: return "unknown error";
    }
}
extern "C" int bmnas_abi_version(void) { return 1; }
int bmnas_validate_only_flag = 0;
int bmnas_pdl_flag = 1;
extern "C" int bmnas_set_pdl(int on) {
    bmnas_pdl_flag = on ? 1 : 0;
    return BMNAS_OK;
}
extern "C" int bmnas_set_validate_only(int on) {
    bmnas_validate_only_flag = on ? 1 : 0;
    return BMNAS_OK;
}

extern "C" int bmnas_zero(void* ptr, long long nbytes, void* stream) {
    if (!ptr || nbytes < 0) return BMNAS_EINVAL;
    if (nbytes == 0) return BMNAS_OK;
    BMNAS_DRY_RETURN();
    if ((reinterpret_cast<uintptr_t>(ptr) & 15u) || nbytes > (1ll << 28))       // unaligned or huge: the driver's memset
        return cudaMemsetAsync(ptr, 0, (size_t)nbytes, (cudaStream_t)stream) == cudaSuccess ? BMNAS_OK : BMNAS_ELAUNCH;
    const long long n16 = nbytes >> 4;
    long long blocks = (n16 + OTH * 4 - 1) / (OTH * 4);
    if (blocks < 1) blocks = 1;
    if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
    launch_k(k_zero, (unsigned)blocks, OTH, 0, (cudaStream_t)stream, reinterpret_cast<uint4*>(ptr), n16,
             reinterpret_cast<unsigned char*>(ptr) + (n16 << 4), (int)(nbytes & 15));
    BMNAS_LAUNCH_CHECK();
    return BMNAS_OK;
}
extern "C" int bmnas_copy(void* dst, const void* src, long long nbytes, void* stream) {
    if (!dst || !src || nbytes < 0 || (nbytes & 15) || ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15u))
        return BMNAS_EINVAL;
    if (nbytes == 0) return BMNAS_OK;
    BMNAS_DRY_RETURN();
    const long long n16 = nbytes >> 4;
    long long blocks = (n16 + OTH * 4 - 1) / (OTH * 4);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    launch_k(k_copy, (unsigned)blocks, OTH, 0, (cudaStream_t)stream, reinterpret_cast<uint4*>(dst), reinterpret_cast<const uint4*>(src), n16);
    BMNAS_LAUNCH_CHECK();
    return BMNAS_OK;
}
extern "C" int bmnas_sizeof_params(int which) {
    switch (which) {
        case 0: return (int)sizeof(bmnas_mix_params);
        case 1: return (int)sizeof(bmnas_conv_params);
        case 2: return (int)sizeof(bmnas_node_params);
        case 3: return (int)sizeof(bmnas_ln_params);
        case 4: return (int)sizeof(bmnas_loss_params);
        case 5: return (int)sizeof(bmnas_adam_tensor);
        case 6: return (int)sizeof(bmnas_adam_params);
        default: return -1;
    }
}
