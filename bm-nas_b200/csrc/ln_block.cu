// LayerNorm blocks of the fusion cell, one CTA per sample (grid-stride):
//   mode 0 (CAT):  out = [ReLU] LN_{[Ctot,L]}( cat(src...) [+ residual] )
//   mode 1 (TAIL): out = LN_{[C,L]}( dropout(ReLU(BN(src0))) + residual )
// The channel concat is virtual (read straight from the source tensors), the
// whole sample lives in shared memory between the statistics passes, and the
// backward fuses LN-backward, residual split, dropout/ReLU backward, the
// BatchNorm affine-grad reductions and the conv-backward coefficient finalise.
#include "common.cuh"

namespace bmnas {

constexpr int LTH = 256;
constexpr int kLnMaxBlocksFwd = kNumSMs * 8;
constexpr int kLnMaxBlocksBwd = kNumSMs * 2;

__host__ __device__ inline size_t lrnd4(size_t n) { return (n + 3) & ~(size_t)3; }
__host__ __device__ inline size_t ln_smem_floats(int Ctot, int L, bool bwd) {
    const size_t E = lrnd4((size_t)Ctot * L);
    return E + 4 * 32 + (bwd ? 2 * E + 2 * lrnd4((size_t)Ctot) : 0) + 16;
}

// pre-LayerNorm value of element e=(c,l) of sample b; also returns what the backward needs
struct PreLN {
    float v;     // value entering LayerNorm
    float zh;    // TAIL: normalised conv output (BN x-hat)
    float dmul;  // TAIL: d v / d(BN output) = dropout scale * [BN output > 0]
};

__device__ __forceinline__ PreLN ln_pre(const bmnas_ln_params& p, int b, int e) {
    const int L = p.L, E = p.Ctot * L;
    PreLN r;
    r.zh = 0.f;
    r.dmul = 0.f;
    const long long li = (long long)b * E + e;
    if (p.mode == 0) {
        int c = e / L;
        const int l = e - c * L;
        int s = 0;
        while (s + 1 < p.n_src && c >= p.src_C[s]) {
            c -= p.src_C[s];
            ++s;
        }
        r.v = __ldg(p.src[s] + ((long long)b * p.src_C[s] + c) * L + l);
    } else {
        const int c = e / L;
        r.zh = (__ldg(p.src[0] + li) - __ldg(p.mean + c)) * __ldg(p.rstd + c);
        const float bn = r.zh * __ldg(p.bn_w + c) + __ldg(p.bn_b + c);
        const bool drop = p.training && p.p_drop > 0.f;
        const float ds = drop_scale(drop, p.mask, p.rng_state, p.op_uid, li,
                                    (unsigned long long)(p.sample_offset + b) * E + e, p.p_drop);
        r.v = fmaxf(bn, 0.f) * ds;
        r.dmul = bn > 0.f ? ds : 0.f;
    }
    if (p.residual) r.v += __ldg(p.residual + li);
    return r;
}

__device__ __forceinline__ void ln_stats(const float* vs, int E, float* red, float* mean, float* rstd) {
    float s0[1] = {0.f}, s1[1] = {0.f};
    for (int e = threadIdx.x; e < E; e += LTH) s0[0] += vs[e];
    block_sum<1>(s0, red);
    const float m = s0[0] / (float)E;
    for (int e = threadIdx.x; e < E; e += LTH) {
        const float d = vs[e] - m;
        s1[0] += d * d;
    }
    block_sum<1>(s1, red);
    *mean = m;
    *rstd = 1.f / sqrtf(s1[0] / (float)E + kLnEps);
}

__global__ void __launch_bounds__(LTH) k_ln_fwd(const bmnas_ln_params p) {
    extern __shared__ __align__(16) float smem[];
    const int E = p.Ctot * p.L;
    float* vs = smem;
    float* red = smem + lrnd4((size_t)E);
    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        __syncthreads();
        for (int e = threadIdx.x; e < E; e += LTH) vs[e] = ln_pre(p, b, e).v;
        __syncthreads();
        float mean, rstd;
        ln_stats(vs, E, red, &mean, &rstd);
        for (int e = threadIdx.x; e < E; e += LTH) {
            float o = (vs[e] - mean) * rstd * __ldg(p.ln_w + e) + __ldg(p.ln_b + e);
            if (p.relu_out) o = fmaxf(o, 0.f);
            p.out[(long long)b * E + e] = o;
        }
    }
}

template <bool SEG>
__device__ __forceinline__ void ln_chan_add(float* acc, int m, float v, int L, bool active) {
    if (SEG) {
        for (int o = L >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (active && ((threadIdx.x & (L - 1)) == 0)) acc[m] += v;
    } else {
        if (active) atomicAdd(acc + m, v);
    }
}

template <bool SEG>
__global__ void __launch_bounds__(LTH) k_ln_bwd(const bmnas_ln_params p) {
    extern __shared__ __align__(16) float smem[];
    const int L = p.L, Ctot = p.Ctot, E = Ctot * L;
    const size_t Er = lrnd4((size_t)E);
    float* vs = smem;
    float* red = vs + Er;
    float* lnG = red + 4 * 32;
    float* lnH = lnG + Er;
    float* S1s = lnH + Er;
    float* S2s = S1s + lrnd4((size_t)Ctot);
    for (int e = threadIdx.x; e < E; e += LTH) {
        lnG[e] = 0.f;
        lnH[e] = 0.f;
    }
    for (int c = threadIdx.x; c < Ctot; c += LTH) {
        S1s[c] = 0.f;
        S2s[c] = 0.f;
    }
    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        __syncthreads();
        for (int e = threadIdx.x; e < E; e += LTH) vs[e] = ln_pre(p, b, e).v;
        __syncthreads();
        float mean, rstd;
        ln_stats(vs, E, red, &mean, &rstd);
        const float* gb = p.gout + (long long)b * E;
        float qs[2] = {0.f, 0.f};
        for (int e = threadIdx.x; e < E; e += LTH) {
            const float vh = (vs[e] - mean) * rstd;
            const float G = __ldg(p.ln_w + e);
            float g = __ldg(gb + e);
            if (p.relu_out && !(vh * G + __ldg(p.ln_b + e) > 0.f)) g = 0.f;
            lnG[e] += g * vh;
            lnH[e] += g;
            const float q = g * G;
            qs[0] += q;
            qs[1] += q * vh;
        }
        block_sum<2>(qs, red);
        const float mq = qs[0] / (float)E, mqo = qs[1] / (float)E;
        for (int e0 = 0; e0 < E; e0 += LTH) {
            const int e = e0 + threadIdx.x;
            const bool act = e < E;
            const int ee = act ? e : 0;
            const float vh = (vs[ee] - mean) * rstd;
            const float G = __ldg(p.ln_w + ee);
            float g = act ? __ldg(gb + ee) : 0.f;
            if (p.relu_out && !(vh * G + __ldg(p.ln_b + ee) > 0.f)) g = 0.f;
            const float dv = act ? rstd * (g * G - mq - vh * mqo) : 0.f;
            const long long li = (long long)b * E + ee;
            if (act && p.gresidual) p.gresidual[li] = p.gres_accum ? p.gresidual[li] + dv : dv;
            if (p.mode == 0) {
                if (act) {
                    int c = ee / L;
                    const int l = ee - c * L;
                    int s = 0;
                    while (s + 1 < p.n_src && c >= p.src_C[s]) {
                        c -= p.src_C[s];
                        ++s;
                    }
                    if (p.gsrc[s]) {
                        float* d = p.gsrc[s] + ((long long)b * p.src_C[s] + c) * L + l;
                        *d = p.gsrc_accum[s] ? (*d + dv) : dv;
                    }
                }
            } else {
                const PreLN pr = ln_pre(p, b, ee);
                const float gv = dv * pr.dmul;
                if (act) p.gsrc[0][li] = gv;
                const int c = ee / L;
                ln_chan_add<SEG>(S1s, c, gv, L, act);
                ln_chan_add<SEG>(S2s, c, gv * pr.zh, L, act);
            }
        }
    }
    __syncthreads();
    if (p.g_ln_w) {
        for (int e = threadIdx.x; e < E; e += LTH) {
            atomicAdd(p.g_ln_w + e, lnG[e]);
            atomicAdd(p.g_ln_b + e, lnH[e]);
        }
    }
    if (p.mode == 0) return;

    float* part = p.partials + (long long)blockIdx.x * 2 * Ctot;
    for (int c = threadIdx.x; c < Ctot; c += LTH) {
        part[c] = S1s[c];
        part[Ctot + c] = S2s[c];
    }
    if (!last_block(p.counter, gridDim.x)) return;
    const float n = (float)p.B * (float)L;
    for (int c = threadIdx.x; c < Ctot; c += LTH) {
        float s1 = 0.f, s2 = 0.f;
        for (unsigned cta = 0; cta < gridDim.x; ++cta) {
            s1 += ld_cg(p.partials + (long long)cta * 2 * Ctot + c);
            s2 += ld_cg(p.partials + (long long)cta * 2 * Ctot + Ctot + c);
        }
        if (p.g_bn_w) {
            p.g_bn_w[c] = s2;
            p.g_bn_b[c] = s1;
        }
        const float rs = p.rstd[c], mu = p.mean[c];
        if (p.training) {
            const float a = p.bn_w[c] * rs, m1 = s1 / n, m2 = s2 / n;
            p.coef_a[c] = a;
            p.coef_b[c] = -a * rs * m2;
            p.coef_c[c] = a * (mu * rs * m2 - m1);
        } else {
            p.coef_a[c] = p.bn_w[c] * rs;
            p.coef_b[c] = 0.f;
            p.coef_c[c] = 0.f;
        }
    }
}

static int ln_check(const bmnas_ln_params* p, bool bwd) {
    if (!p || p->B < 1 || p->L < 1 || p->Ctot < 1 || p->n_src < 1 || p->n_src > BMNAS_MAX_SRC) return BMNAS_EINVAL;
    if (p->mode != 0 && p->mode != 1) return BMNAS_EINVAL;
    if (!p->ln_w || !p->ln_b) return BMNAS_EINVAL;
    int c = 0;
    for (int i = 0; i < p->n_src; ++i) {
        if (!p->src[i]) return BMNAS_EINVAL;
        c += p->src_C[i];
    }
    if (c != p->Ctot) return BMNAS_EINVAL;
    if (p->mode == 1) {
        if (p->n_src != 1 || !p->mean || !p->rstd || !p->bn_w || !p->bn_b) return BMNAS_EINVAL;
        if (p->p_drop < 0.f || p->p_drop >= 1.f) return BMNAS_EINVAL;
        if (p->training && p->p_drop > 0.f && !p->mask && !p->rng_state) return BMNAS_EINVAL;
        if (bwd && (!p->gsrc[0] || !p->coef_a || !p->coef_b || !p->coef_c || !p->partials || !p->counter))
            return BMNAS_EINVAL;
    }
    if (bwd && !p->gout) return BMNAS_EINVAL;
    if (!bwd && !p->out) return BMNAS_EINVAL;
    return BMNAS_OK;
}

}  // namespace bmnas

using namespace bmnas;

extern "C" long long bmnas_ln_partials_size(const bmnas_ln_params* p) {
    return (long long)kLnMaxBlocksBwd * 2 * p->Ctot;
}

extern "C" int bmnas_ln_fwd(const bmnas_ln_params* p, void* stream) {
    int e = ln_check(p, false);
    if (e) return e;
    const size_t smem = ln_smem_floats(p->Ctot, p->L, false) * sizeof(float);
    if (smem > 227 * 1024) return BMNAS_EINVAL;
    BMNAS_DRY_RETURN();
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        if (cudaFuncSetAttribute(k_ln_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return BMNAS_ELAUNCH;
        configured = smem;
    }
    const int blocks = p->B < kLnMaxBlocksFwd ? p->B : kLnMaxBlocksFwd;
    k_ln_fwd<<<blocks, LTH, smem, (cudaStream_t)stream>>>(*p);
    BMNAS_LAUNCH_CHECK();
    return BMNAS_OK;
}

extern "C" int bmnas_ln_bwd(const bmnas_ln_params* p, void* stream) {
    int e = ln_check(p, true);
    if (e) return e;
    const size_t smem = ln_smem_floats(p->Ctot, p->L, true) * sizeof(float);
    if (smem > 227 * 1024) return BMNAS_EINVAL;
    const bool seg = (p->L & (p->L - 1)) == 0 && p->L <= 32;
    BMNAS_DRY_RETURN();
    static size_t configured[2] = {0, 0};
    if (smem > 48 * 1024 && smem > configured[seg]) {
        cudaError_t ce = seg ? cudaFuncSetAttribute(k_ln_bwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                             : cudaFuncSetAttribute(k_ln_bwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (ce != cudaSuccess) return BMNAS_ELAUNCH;
        configured[seg] = smem;
    }
    const int blocks = p->B < kLnMaxBlocksBwd ? p->B : kLnMaxBlocksBwd;
    if (seg)
        k_ln_bwd<true><<<blocks, LTH, smem, (cudaStream_t)stream>>>(*p);
    else
        k_ln_bwd<false><<<blocks, LTH, smem, (cudaStream_t)stream>>>(*p);
    BMNAS_LAUNCH_CHECK();
    return BMNAS_OK;
}
