// LayerNorm blocks of the fusion cell, one CTA per sample (grid-stride):
//   mode 0 (CAT):  out = [ReLU] LN_{[Ctot,L]}( cat(src...) [+ residual] )
//   mode 1 (TAIL): out = LN_{[C,L]}( dropout(ReLU(BN(src0))) + residual )
// The channel concat is virtual (read straight from the source tensors), the
// whole sample lives in shared memory between the statistics passes, and the
// backward fuses LN-backward, residual split, dropout/ReLU backward, the
// BatchNorm affine-grad reductions and the conv-backward coefficient finalise.
// Threads own groups of G=4 consecutive elements (128-bit traffic, one Philox call
// per group); per-channel BatchNorm constants are folded once per CTA.
#include "common.cuh"

namespace bmnas {

// threads per CTA (one sample): 256 up to 256 four-element groups per sample, 512 / 1024 for larger samples (Ego-large's
// cell tail is 16 384 elements) -- see node_apply.cu.  Device code reads the size from blockDim.
constexpr int LTH0 = 256, LTH_MAX = 1024;
static inline int ln_threads(int Ctot, int L, int B) {
    const int groups = (Ctot * L + 3) / 4;
    int n = LTH0;
    if (B > 2 * kNumSMs) return n;     // enough samples to fill the machine with 256-thread CTAs (wider ones cost occupancy:
                                       // B = 8192, E = 2048: forward 41 -> 67 us, backward 105 -> 140 us with 512 threads)
    while (n < groups && n < LTH_MAX) n *= 2;
    return n;
}
constexpr int kLnMaxBlocksFwd = kNumSMs * 8;
// backward: 60 registers x 256 threads -> four CTAs fit an SM.  Two per SM (round 1) left the kernel at 25 % warp occupancy
// with every phase of a sample waiting on a global round trip (ncu: long_sb 30 %, DRAM 13 %); the price of more CTAs is the
// LayerNorm-affine gradient flush (2 E atomics per CTA), ~1 M more L2 atomics per launch at B = 8 192
constexpr int kLnMaxBlocksBwd = kNumSMs * 4;

__host__ __device__ inline size_t lrnd4(size_t n) { return (n + 3) & ~(size_t)3; }
__host__ __device__ inline size_t ln_smem_floats(int Ctot, int L, bool bwd) {
    const size_t E = lrnd4((size_t)Ctot * L), Cr = lrnd4((size_t)Ctot);
    return E + 4 * 32 + 4 * Cr + (bwd ? 2 * E + 2 * Cr : 0) + 16;
}

template <int G>
__device__ __forceinline__ void lg_v(const float* p, float (&v)[G]) {
    if (G == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1 % G] = t.y; v[2 % G] = t.z; v[3 % G] = t.w;
    } else {
        v[0] = __ldg(p);
    }
}
template <int G>
__device__ __forceinline__ void ll_v(const float* p, float (&v)[G]) {
    if (G == 4) {
        const float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1 % G] = t.y; v[2 % G] = t.z; v[3 % G] = t.w;
    } else {
        v[0] = *p;
    }
}
template <int G>
__device__ __forceinline__ void ls_v(float* p, const float (&v)[G]) {
    if (G == 4) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1 % G], v[2 % G], v[3 % G]);
    else *p = v[0];
}

template <int G>
__device__ __forceinline__ void ln_drop(const bmnas_ln_params& p, long long li, unsigned long long gi, float (&ds)[G]) {
    const bool active = p.training && p.p_drop > 0.f;
    if (!active) {
#pragma unroll
        for (int j = 0; j < G; ++j) ds[j] = 1.f;
        return;
    }
    const float keep = 1.f / (1.f - p.p_drop);
    if (p.mask) {
        if (G == 4) {
            const uchar4 m = *reinterpret_cast<const uchar4*>(p.mask + li);
            ds[0] = m.x ? keep : 0.f; ds[1 % G] = m.y ? keep : 0.f; ds[2 % G] = m.z ? keep : 0.f; ds[3 % G] = m.w ? keep : 0.f;
        } else {
            ds[0] = p.mask[li] ? keep : 0.f;
        }
    } else if (G == 4) {
        const unsigned long long seed = p.rng_state[0], step = p.rng_state[1];
        const uint32_t uid = p.op_uid;
        const uint2 key = make_uint2((uint32_t)seed ^ (uid * 0x9E3779B1u), (uint32_t)(seed >> 32) + uid);
        const uint4 r = philox4x32(make_uint4((uint32_t)(gi >> 2), (uint32_t)(gi >> 34), (uint32_t)step,
                                              (uint32_t)(step >> 32)), key);
        const float sc = 1.0f / 16777216.0f;
        ds[0] = ((float)(r.x >> 8) * sc >= p.p_drop) ? keep : 0.f;
        ds[1 % G] = ((float)(r.y >> 8) * sc >= p.p_drop) ? keep : 0.f;
        ds[2 % G] = ((float)(r.z >> 8) * sc >= p.p_drop) ? keep : 0.f;
        ds[3 % G] = ((float)(r.w >> 8) * sc >= p.p_drop) ? keep : 0.f;
    } else {
        ds[0] = philox_keep(p.rng_state, p.op_uid, gi, p.p_drop) ? keep : 0.f;
    }
}

// value entering LayerNorm for one group; TAIL mode also returns x-hat of BN and d v / d(BN out)
template <int G>
__device__ __forceinline__ void ln_pre(const bmnas_ln_params& p, const float* cst, int b, int e0, float (&v)[G],
                                       float (&zh)[G], float (&dmul)[G]) {
    const int L = p.L, E = p.Ctot * L, Cr = (int)lrnd4((size_t)p.Ctot);
    const long long li = (long long)b * E + e0;
    int c = e0 / L;
    if (p.mode == 0) {
        const int l = e0 - c * L;
        int s = 0;
        while (s + 1 < p.n_src && c >= p.src_C[s]) {
            c -= p.src_C[s];
            ++s;
        }
        lg_v<G>(p.src[s] + ((long long)b * p.src_C[s] + c) * L + l, v);
#pragma unroll
        for (int q = 0; q < G; ++q) zh[q] = dmul[q] = 0.f;
    } else {
        float z[G], ds[G];
        lg_v<G>(p.src[0] + li, z);
        ln_drop<G>(p, li, (unsigned long long)(p.sample_offset + b) * E + e0, ds);
        const float r = cst[c], mr = cst[Cr + c], w = cst[2 * Cr + c], bb = cst[3 * Cr + c];
#pragma unroll
        for (int q = 0; q < G; ++q) {
            zh[q] = fmaf(z[q], r, -mr);
            const float bn = fmaf(zh[q], w, bb);
            v[q] = fmaxf(bn, 0.f) * ds[q];
            dmul[q] = bn > 0.f ? ds[q] : 0.f;
        }
    }
    if (p.residual) {
        float rr[G];
        lg_v<G>(p.residual + li, rr);
#pragma unroll
        for (int q = 0; q < G; ++q) v[q] += rr[q];
    }
}

template <int LTH>
__device__ __forceinline__ void ln_consts(const bmnas_ln_params& p, float* cst) {
    if (p.mode == 1) {
        const int Cr = (int)lrnd4((size_t)p.Ctot);
        for (int c = threadIdx.x; c < p.Ctot; c += LTH) {
            const float r = __ldg(p.rstd + c);
            cst[c] = r;
            cst[Cr + c] = __ldg(p.mean + c) * r;
            cst[2 * Cr + c] = __ldg(p.bn_w + c);
            cst[3 * Cr + c] = __ldg(p.bn_b + c);
        }
    }
    __syncthreads();
}

template <int G, int LTH>
__device__ __forceinline__ void ln_stats(const float* vs, int E, float* red, float* mean, float* rstd) {
    float s0[1] = {0.f}, s1[1] = {0.f};
    for (int g = threadIdx.x; g < E / G; g += LTH) {
        float v[G];
        ll_v<G>(vs + g * G, v);
#pragma unroll
        for (int q = 0; q < G; ++q) s0[0] += v[q];
    }
    block_sum<1>(s0, red);
    const float m = s0[0] / (float)E;
    for (int g = threadIdx.x; g < E / G; g += LTH) {
        float v[G];
        ll_v<G>(vs + g * G, v);
#pragma unroll
        for (int q = 0; q < G; ++q) {
            const float d = v[q] - m;
            s1[0] += d * d;
        }
    }
    block_sum<1>(s1, red);
    *mean = m;
    *rstd = 1.f / sqrtf(s1[0] / (float)E + kLnEps);
}

template <int G, int LTH>
__global__ void __launch_bounds__(LTH) k_ln_fwd(const bmnas_ln_params p) {
    pdl_prologue();
    extern __shared__ __align__(16) float smem[];
    const int E = p.Ctot * p.L;
    float* vs = smem;
    float* red = smem + lrnd4((size_t)E);
    float* cst = red + 4 * 32;
    ln_consts<LTH>(p, cst);
    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        __syncthreads();
        for (int g = threadIdx.x; g < E / G; g += LTH) {
            float v[G], zh[G], dm[G];
            ln_pre<G>(p, cst, b, g * G, v, zh, dm);
            ls_v<G>(vs + g * G, v);
        }
        __syncthreads();
        float mean, rstd;
        ln_stats<G, LTH>(vs, E, red, &mean, &rstd);
        for (int g = threadIdx.x; g < E / G; g += LTH) {
            const int e0 = g * G;
            float v[G], w[G], bb[G], o[G];
            ll_v<G>(vs + e0, v);
            lg_v<G>(p.ln_w + e0, w);
            lg_v<G>(p.ln_b + e0, bb);
#pragma unroll
            for (int q = 0; q < G; ++q) {
                o[q] = (v[q] - mean) * rstd * w[q] + bb[q];
                if (p.relu_out) o[q] = fmaxf(o[q], 0.f);
            }
            ls_v<G>(p.out + (long long)b * E + e0, o);
        }
    }
}

template <bool SEG>
__device__ __forceinline__ void ln_chan_add(float* acc, int m, float v, int lanes, bool active) {
    if (SEG) {
        for (int o = lanes >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (active && ((threadIdx.x & (lanes - 1)) == 0)) acc[m] += v;
    } else {
        if (active) atomicAdd(acc + m, v);
    }
}

template <int G, bool SEG, int LTH>
__global__ void __launch_bounds__(LTH) k_ln_bwd(const bmnas_ln_params p) {
    pdl_prologue();
    extern __shared__ __align__(16) float smem[];
    const int L = p.L, Ctot = p.Ctot, E = Ctot * L, NG = E / G;
    const size_t Er = lrnd4((size_t)E), Cr = lrnd4((size_t)Ctot);
    float* vs = smem;
    float* red = vs + Er;
    float* cst = red + 4 * 32;
    float* lnG = cst + 4 * Cr;
    float* lnH = lnG + Er;
    float* S1s = lnH + Er;
    float* S2s = S1s + Cr;
    const int lanes = SEG ? L / G : 1;
    for (int e = threadIdx.x; e < E; e += LTH) {
        lnG[e] = 0.f;
        lnH[e] = 0.f;
    }
    for (int c = threadIdx.x; c < Ctot; c += LTH) {
        S1s[c] = 0.f;
        S2s[c] = 0.f;
    }
    ln_consts<LTH>(p, cst);
    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        __syncthreads();
        for (int g = threadIdx.x; g < NG; g += LTH) {
            float v[G], zh[G], dm[G];
            ln_pre<G>(p, cst, b, g * G, v, zh, dm);
            ls_v<G>(vs + g * G, v);
        }
        __syncthreads();
        float mean, rstd;
        ln_stats<G, LTH>(vs, E, red, &mean, &rstd);
        const float* gb = p.gout + (long long)b * E;
        float qs[2] = {0.f, 0.f};
        for (int g = threadIdx.x; g < NG; g += LTH) {
            const int e0 = g * G;
            float v[G], Gw[G], Gb[G], gg[G], lg[G], lh[G];
            ll_v<G>(vs + e0, v);
            lg_v<G>(p.ln_w + e0, Gw);
            lg_v<G>(p.ln_b + e0, Gb);
            lg_v<G>(gb + e0, gg);
            ll_v<G>(lnG + e0, lg);
            ll_v<G>(lnH + e0, lh);
#pragma unroll
            for (int q = 0; q < G; ++q) {
                const float vh = (v[q] - mean) * rstd;
                float g_ = gg[q];
                if (p.relu_out && !(vh * Gw[q] + Gb[q] > 0.f)) g_ = 0.f;
                lg[q] += g_ * vh;
                lh[q] += g_;
                const float qq = g_ * Gw[q];
                qs[0] += qq;
                qs[1] += qq * vh;
            }
            ls_v<G>(lnG + e0, lg);
            ls_v<G>(lnH + e0, lh);
        }
        block_sum<2>(qs, red);
        const float mq = qs[0] / (float)E, mqo = qs[1] / (float)E;
        for (int g0 = 0; g0 < NG; g0 += LTH) {
            const int g = g0 + threadIdx.x;
            const bool act = g < NG;
            const int e0 = act ? g * G : 0;
            const long long li = (long long)b * E + e0;
            float v[G], Gw[G], Gb[G], gg[G], dv[G];
            ll_v<G>(vs + e0, v);
            lg_v<G>(p.ln_w + e0, Gw);
            lg_v<G>(p.ln_b + e0, Gb);
            lg_v<G>(gb + e0, gg);
#pragma unroll
            for (int q = 0; q < G; ++q) {
                const float vh = (v[q] - mean) * rstd;
                float g_ = act ? gg[q] : 0.f;
                if (p.relu_out && !(vh * Gw[q] + Gb[q] > 0.f)) g_ = 0.f;
                dv[q] = act ? rstd * (g_ * Gw[q] - mq - vh * mqo) : 0.f;
            }
            if (act && p.gresidual) {
                float o[G];
#pragma unroll
                for (int q = 0; q < G; ++q) o[q] = dv[q];
                if (p.gres_accum) {
                    float c_[G];
                    ll_v<G>(p.gresidual + li, c_);
#pragma unroll
                    for (int q = 0; q < G; ++q) o[q] += c_[q];
                }
                ls_v<G>(p.gresidual + li, o);
            }
            if (p.mode == 0) {
                if (act) {
                    int c = e0 / L;
                    const int l = e0 - c * L;
                    int s = 0;
                    while (s + 1 < p.n_src && c >= p.src_C[s]) {
                        c -= p.src_C[s];
                        ++s;
                    }
                    if (p.gsrc[s]) {
                        float* d = p.gsrc[s] + ((long long)b * p.src_C[s] + c) * L + l;
                        float o[G];
#pragma unroll
                        for (int q = 0; q < G; ++q) o[q] = dv[q];
                        if (p.gsrc_accum[s]) {
                            float c_[G];
                            ll_v<G>(d, c_);
#pragma unroll
                            for (int q = 0; q < G; ++q) o[q] += c_[q];
                        }
                        ls_v<G>(d, o);
                    }
                }
            } else {
                float pv[G], zh[G], dm[G], gv[G];
                ln_pre<G>(p, cst, b, e0, pv, zh, dm);
                float s1 = 0.f, s2 = 0.f;
#pragma unroll
                for (int q = 0; q < G; ++q) {
                    gv[q] = dv[q] * dm[q];
                    s1 += gv[q];
                    s2 += gv[q] * zh[q];
                }
                if (act) ls_v<G>(p.gsrc[0] + li, gv);
                const int c = e0 / L;
                ln_chan_add<SEG>(S1s, c, s1, lanes, act);
                ln_chan_add<SEG>(S2s, c, s2, lanes, act);
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

    float* gacc = p.partials;            // self-cleaning global accumulator [2*Ctot]
    for (int c = threadIdx.x; c < Ctot; c += LTH) {
        atomicAdd(gacc + c, S1s[c]);
        atomicAdd(gacc + Ctot + c, S2s[c]);
    }
    if (!last_block(p.counter, gridDim.x)) return;
    const float n = (float)p.B * (float)L;
    for (int c = threadIdx.x; c < Ctot; c += LTH) {
        const float s1 = ld_cg(gacc + c), s2 = ld_cg(gacc + Ctot + c);
        gacc[c] = 0.f;
        gacc[Ctot + c] = 0.f;
        if (p.g_bn_w) {
            p.g_bn_w[c] = s2;
            p.g_bn_b[c] = s1;
        }
        const float rs = cst[c], mur = cst[Cr + c], w = cst[2 * Cr + c];
        if (p.training) {
            const float a = w * rs, m1 = s1 / n, m2 = s2 / n;
            p.coef_a[c] = a;
            p.coef_b[c] = -a * rs * m2;
            p.coef_c[c] = a * (mur * m2 - m1);
        } else {
            p.coef_a[c] = w * rs;
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

static bool lal16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }
static bool ln_vec_ok(const bmnas_ln_params* p) {
    if (p->L % 4) return false;
    if (!lal16(p->residual) || !lal16(p->ln_w) || !lal16(p->ln_b) || !lal16(p->out) || !lal16(p->gout) ||
        !lal16(p->gresidual))
        return false;
    for (int i = 0; i < p->n_src; ++i)
        if (!lal16(p->src[i]) || !lal16(p->gsrc[i])) return false;
    if (p->mask && (reinterpret_cast<uintptr_t>(p->mask) & 3u)) return false;
    return true;
}

template <class Kern>
static int ln_smem_attr(Kern kern, size_t smem, size_t* configured) {
    if (smem > 40 * 1024 && smem > *configured) {   // the 48 KB default counts static + dynamic: opt in with a margin
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return BMNAS_ELAUNCH;
        *configured = smem;
    }
    return BMNAS_OK;
}

}  // namespace bmnas

using namespace bmnas;

extern "C" long long bmnas_ln_partials_size(const bmnas_ln_params* p) { return 2LL * p->Ctot; }

template <class Kern>
static int ln_launch(Kern kern, size_t* configured, int blocks, int threads, size_t smem, const bmnas_ln_params* p, cudaStream_t stream) {
    if (int e = ln_smem_attr(kern, smem, configured)) return e;
    launch_k(kern, blocks, threads, smem, stream, *p);
    BMNAS_LAUNCH_CHECK();
    return BMNAS_OK;
}

template <int LTH>
static int ln_fwd_t(const bmnas_ln_params* p, bool vec, int blocks, size_t smem, cudaStream_t stream) {
    static size_t configured[2] = {0, 0};
    return vec ? ln_launch(k_ln_fwd<4, LTH>, &configured[1], blocks, LTH, smem, p, stream)
               : ln_launch(k_ln_fwd<1, LTH>, &configured[0], blocks, LTH, smem, p, stream);
}

template <int LTH>
static int ln_bwd_t(const bmnas_ln_params* p, bool vec, bool seg, int blocks, size_t smem, cudaStream_t stream) {
    static size_t configured[3] = {0, 0, 0};
    if (vec && seg) return ln_launch(k_ln_bwd<4, true, LTH>, &configured[0], blocks, LTH, smem, p, stream);
    if (vec) return ln_launch(k_ln_bwd<4, false, LTH>, &configured[1], blocks, LTH, smem, p, stream);
    return ln_launch(k_ln_bwd<1, false, LTH>, &configured[2], blocks, LTH, smem, p, stream);
}

extern "C" int bmnas_ln_fwd(const bmnas_ln_params* p, void* stream) {
    int e = ln_check(p, false);
    if (e) return e;
    const size_t smem = ln_smem_floats(p->Ctot, p->L, false) * sizeof(float);
    if (smem > 227 * 1024) return BMNAS_EINVAL;
    BMNAS_DRY_RETURN();
    const bool vec = ln_vec_ok(p);
    const int blocks = p->B < kLnMaxBlocksFwd ? p->B : kLnMaxBlocksFwd;
    const int th = ln_threads(p->Ctot, p->L, p->B);
    if (th >= 1024) return ln_fwd_t<1024>(p, vec, blocks, smem, (cudaStream_t)stream);
    if (th >= 512) return ln_fwd_t<512>(p, vec, blocks, smem, (cudaStream_t)stream);
    return ln_fwd_t<256>(p, vec, blocks, smem, (cudaStream_t)stream);
}

extern "C" int bmnas_ln_bwd(const bmnas_ln_params* p, void* stream) {
    int e = ln_check(p, true);
    if (e) return e;
    const size_t smem = ln_smem_floats(p->Ctot, p->L, true) * sizeof(float);
    if (smem > 227 * 1024) return BMNAS_EINVAL;
    BMNAS_DRY_RETURN();
    const bool vec = ln_vec_ok(p);
    const int lanes = p->L / 4;
    const bool seg = vec && lanes >= 1 && lanes <= 32 && (lanes & (lanes - 1)) == 0;
    const int blocks = p->B < kLnMaxBlocksBwd ? p->B : kLnMaxBlocksBwd;
    const int th = ln_threads(p->Ctot, p->L, p->B);
    if (th >= 1024) return ln_bwd_t<1024>(p, vec, seg, blocks, smem, (cudaStream_t)stream);
    if (th >= 512) return ln_bwd_t<512>(p, vec, seg, blocks, smem, (cudaStream_t)stream);
    return ln_bwd_t<256>(p, vec, seg, blocks, smem, (cudaStream_t)stream);
}
