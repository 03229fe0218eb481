// Step-node mixed op: out = sum_k gamma~_k * op_k(x, y) evaluated per sample from
// the x / y tiles staged once in shared memory (Sum, ScaledDotAttn + LayerNorm,
// LinearGLU and ConcatFC / CatConvMish epilogues on the pre-BN conv output Z),
// softmax(gamma) weighted sum in the epilogue; per-op outputs never reach HBM.
// The backward recomputes the primitives, emits GV = dL/d(BN output) for the conv
// GEMMs, reduces dL/dgamma (warp shuffle -> block -> fixed-order last-block sum),
// the BatchNorm affine grads and the coefficients that fold BatchNorm-backward
// into the conv backward operand loads.
// One CTA per sample (grid-stride over samples), 256 threads.
#include "common.cuh"

namespace bmnas {

constexpr int NTH = 256;
constexpr int kNodeMaxBlocksFwd = kNumSMs * 8;
constexpr int kNodeMaxBlocksBwd = kNumSMs * 2;

struct NodeSmem {
    float *xs, *ys, *gs, *as, *dxs, *dys, *S, *S2, *Sp, *S1s, *S2s, *lnG, *lnH, *red, *gw;
};

__host__ __device__ inline size_t rnd4(size_t n) { return (n + 3) & ~(size_t)3; }

__host__ __device__ inline size_t node_smem_floats(int C, int L, int M, bool bwd) {
    const size_t CL = rnd4((size_t)C * L), LL = rnd4((size_t)L * L);
    size_t n = 0;
    n += 3 * CL;                            // xs, ys, as
    n += 2 * LL + (LL > NTH ? LL : NTH);    // S, S2, Sp
    n += 8 * 32 + 8;                        // red, gw
    if (bwd) n += 5 * CL + 2 * rnd4((size_t)M);  // gs, dxs, dys, lnG, lnH, S1s, S2s
    return n + 16;
}

__device__ __forceinline__ NodeSmem node_carve(float* base, int C, int L, int M, bool bwd) {
    const size_t CL = rnd4((size_t)C * L), LL = rnd4((size_t)L * L);
    NodeSmem s;
    float* q = base;
    s.xs = q; q += CL;
    s.ys = q; q += CL;
    s.as = q; q += CL;
    s.S = q; q += LL;
    s.S2 = q; q += LL;
    s.Sp = q; q += (LL > NTH ? LL : NTH);
    s.red = q; q += 8 * 32;
    s.gw = q; q += 8;
    s.gs = s.dxs = s.dys = s.S1s = s.S2s = s.lnG = s.lnH = nullptr;
    if (bwd) {
        s.gs = q; q += CL;
        s.dxs = q; q += CL;
        s.dys = q; q += CL;
        s.lnG = q; q += CL;
        s.lnH = q; q += CL;
        s.S1s = q; q += rnd4((size_t)M);
        s.S2s = q; q += rnd4((size_t)M);
    }
    return s;
}

// out[i*L+j] = scale * sum_c A[c*L+i] * Bm[c*L+j]   (L x L contraction over channels)
__device__ __forceinline__ void lxl_contract(const float* A, const float* Bm, float* Sp, float* out, int C, int L,
                                             float scale) {
    const int pairs = L * L;
    const int nslice = pairs <= NTH ? NTH / pairs : 1;
    for (int w = threadIdx.x; w < nslice * pairs; w += NTH) {
        const int s = w / pairs, pr = w - s * pairs, i = pr / L, j = pr - i * L;
        float acc = 0.f;
        for (int c = s; c < C; c += nslice) acc = fmaf(A[c * L + i], Bm[c * L + j], acc);
        Sp[w] = acc;
    }
    __syncthreads();
    for (int pr = threadIdx.x; pr < pairs; pr += NTH) {
        float a = 0.f;
        for (int s = 0; s < nslice; ++s) a += Sp[s * pairs + pr];
        out[pr] = a * scale;
    }
    __syncthreads();
}

__device__ __forceinline__ void load_tile(float* dst, const float* src, int CL) {
    if ((CL & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0)) {
        for (int i = threadIdx.x; i < CL / 4; i += NTH)
            reinterpret_cast<float4*>(dst)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
    } else {
        for (int i = threadIdx.x; i < CL; i += NTH) dst[i] = __ldg(src + i);
    }
}

__device__ __forceinline__ void node_weights(const bmnas_node_params& p, float* gw) {
    if (threadIdx.x == 0) {
        if (!p.gamma) {
            for (int k = 0; k < p.n_ops; ++k) gw[k] = 1.f;
        } else if (p.gamma_is_logits) {
            float mx = -INFINITY;
            for (int k = 0; k < p.n_ops; ++k) mx = fmaxf(mx, p.gamma[k]);
            float s = 0.f;
            for (int k = 0; k < p.n_ops; ++k) {
                gw[k] = expf(p.gamma[k] - mx);
                s += gw[k];
            }
            for (int k = 0; k < p.n_ops; ++k) gw[k] /= s;
        } else {
            for (int k = 0; k < p.n_ops; ++k) gw[k] = p.gamma[k];
        }
    }
    __syncthreads();
}

// attention forward for one sample: P (L x L) in sm.S, dropped output a[c,i] in sm.as,
// returns LayerNorm statistics of a.  (ScaledDotAttn.forward node_operations.py:92-108)
__device__ __forceinline__ void attn_forward(const bmnas_node_params& p, const NodeSmem& sm, int k, int b,
                                             float* mean_out, float* rstd_out) {
    const int C = p.C, L = p.L, CL = C * L;
    lxl_contract(sm.xs, sm.ys, sm.Sp, sm.S, C, L, 1.f / sqrtf((float)C));  // S[i][j] = q_i . k_j / sqrt(C)
    for (int i = threadIdx.x; i < L; i += NTH) {  // softmax over key positions j
        float mx = -INFINITY;
        for (int j = 0; j < L; ++j) mx = fmaxf(mx, sm.S[i * L + j]);
        float s = 0.f;
        for (int j = 0; j < L; ++j) {
            const float e = expf(sm.S[i * L + j] - mx);
            sm.S[i * L + j] = e;
            s += e;
        }
        const float inv = 1.f / s;
        for (int j = 0; j < L; ++j) sm.S[i * L + j] *= inv;
    }
    __syncthreads();
    const bool drop = p.training && p.p_drop[k] > 0.f;
    float s0[1] = {0.f}, s1[1] = {0.f};
    for (int e = threadIdx.x; e < CL; e += NTH) {
        const int c = e / L, i = e - c * L;
        float o = 0.f;
        for (int j = 0; j < L; ++j) o = fmaf(sm.S[i * L + j], sm.ys[c * L + j], o);
        const long long li = (long long)b * CL + e;
        o *= drop_scale(drop, p.mask[k], p.rng_state, p.op_uid[k], li,
                        (unsigned long long)(p.sample_offset + b) * CL + e, p.p_drop[k]);
        sm.as[e] = o;
        s0[0] += o;
    }
    block_sum<1>(s0, sm.red);
    const float mean = s0[0] / (float)CL;
    for (int e = threadIdx.x; e < CL; e += NTH) {
        const float d = sm.as[e] - mean;
        s1[0] += d * d;
    }
    block_sum<1>(s1, sm.red);
    *mean_out = mean;
    *rstd_out = 1.f / sqrtf(s1[0] / (float)CL + kLnEps);
}

__global__ void __launch_bounds__(NTH) k_node_fwd(const bmnas_node_params p) {
    extern __shared__ __align__(16) float smem[];
    const int C = p.C, L = p.L, CL = C * L, M = p.M;
    NodeSmem sm = node_carve(smem, C, L, M, false);
    if (p.alias_xy) sm.ys = sm.xs;
    node_weights(p, sm.gw);
    int k_attn = -1;
    for (int k = 0; k < p.n_ops; ++k)
        if (p.op_type[k] == BMNAS_OP_ATTN) k_attn = k;

    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        __syncthreads();
        load_tile(sm.xs, p.x + (long long)b * CL, CL);
        if (!p.alias_xy) load_tile(sm.ys, p.y + (long long)b * CL, CL);
        __syncthreads();
        float a_mean = 0.f, a_rstd = 0.f;
        if (k_attn >= 0) attn_forward(p, sm, k_attn, b, &a_mean, &a_rstd);
        const float* Zb = p.Z ? p.Z + (long long)b * M * L : nullptr;
        for (int e = threadIdx.x; e < CL; e += NTH) {
            const int c = e / L;
            const float xv = sm.xs[e], yv = sm.ys[e];
            const long long li = (long long)b * CL + e;
            const unsigned long long gi = (unsigned long long)(p.sample_offset + b) * CL + e;
            float acc = 0.f;
            for (int k = 0; k < p.n_ops; ++k) {
                float o;
                const int ty = p.op_type[k];
                if (ty == BMNAS_OP_SUM) {
                    o = xv + yv;
                } else if (ty == BMNAS_OP_ATTN) {
                    o = (sm.as[e] - a_mean) * a_rstd * __ldg(p.ln_w[k] + e) + __ldg(p.ln_b[k] + e);
                } else {
                    const int zo = p.z_off[k];
                    const bool drop = p.training && p.p_drop[k] > 0.f;
                    const float ds = drop_scale(drop, p.mask[k], p.rng_state, p.op_uid[k], li, gi, p.p_drop[k]);
                    const float va = (__ldg(Zb + (long long)zo * L + e) - __ldg(p.mean + zo + c)) *
                                         __ldg(p.rstd + zo + c) * __ldg(p.bn_w[k] + c) + __ldg(p.bn_b[k] + c);
                    if (ty == BMNAS_OP_GLU) {
                        const float vg = (__ldg(Zb + (long long)(zo + C) * L + e) - __ldg(p.mean + zo + C + c)) *
                                             __ldg(p.rstd + zo + C + c) * __ldg(p.bn_w[k] + C + c) +
                                         __ldg(p.bn_b[k] + C + c);
                        o = va * sigmoidf_(vg) * ds;
                    } else if (ty == BMNAS_OP_FC_RELU) {
                        o = fmaxf(va, 0.f) * ds;
                    } else {
                        o = mishf_(va) * ds;
                    }
                }
                acc = fmaf(sm.gw[k], o, acc);
            }
            p.out[li] = acc;
        }
    }
}

// add v into acc[m] for the channel m shared by the L consecutive lanes of a segment
template <bool SEG>
__device__ __forceinline__ void chan_add(float* acc, int m, float v, int L, bool active) {
    if (SEG) {
        for (int o = L >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (active && ((threadIdx.x & (L - 1)) == 0)) acc[m] += v;
    } else {
        if (active) atomicAdd(acc + m, v);
    }
}

template <bool SEG>
__global__ void __launch_bounds__(NTH) k_node_bwd(const bmnas_node_params p) {
    extern __shared__ __align__(16) float smem[];
    const int C = p.C, L = p.L, CL = C * L, M = p.M;
    NodeSmem sm = node_carve(smem, C, L, M, true);
    if (p.alias_xy) sm.ys = sm.xs;
    node_weights(p, sm.gw);
    int k_attn = -1;
    for (int k = 0; k < p.n_ops; ++k)
        if (p.op_type[k] == BMNAS_OP_ATTN) k_attn = k;
    for (int i = threadIdx.x; i < M; i += NTH) {
        sm.S1s[i] = 0.f;
        sm.S2s[i] = 0.f;
    }
    for (int i = threadIdx.x; i < CL; i += NTH) {
        sm.lnG[i] = 0.f;
        sm.lnH[i] = 0.f;
    }
    float dg[BMNAS_MAX_OPS];
#pragma unroll
    for (int k = 0; k < BMNAS_MAX_OPS; ++k) dg[k] = 0.f;
    const float inv_sqrt_c = 1.f / sqrtf((float)C);

    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        __syncthreads();
        load_tile(sm.xs, p.x + (long long)b * CL, CL);
        if (!p.alias_xy) load_tile(sm.ys, p.y + (long long)b * CL, CL);
        load_tile(sm.gs, p.gout + (long long)b * CL, CL);
        __syncthreads();
        float a_mean = 0.f, a_rstd = 0.f;
        if (k_attn >= 0) attn_forward(p, sm, k_attn, b, &a_mean, &a_rstd);
        const float* Zb = p.Z ? p.Z + (long long)b * M * L : nullptr;
        float* GVb = p.GV ? p.GV + (long long)b * M * L : nullptr;
        float lnsum[2] = {0.f, 0.f};  // sum q, sum q*ohat for the attention LayerNorm backward

        for (int e0 = 0; e0 < CL; e0 += NTH) {
            const int e = e0 + threadIdx.x;
            const bool act = e < CL;
            const int ee = act ? e : 0;
            const int c = ee / L;
            const float xv = sm.xs[ee], yv = sm.ys[ee], g = act ? sm.gs[ee] : 0.f;
            const long long li = (long long)b * CL + ee;
            const unsigned long long gi = (unsigned long long)(p.sample_offset + b) * CL + ee;
            float gxe = 0.f, gye = 0.f;
#pragma unroll
            for (int k = 0; k < BMNAS_MAX_OPS; ++k) {
                if (k >= p.n_ops) break;
                const int ty = p.op_type[k];
                const float wk = sm.gw[k];
                if (ty == BMNAS_OP_SUM) {
                    dg[k] += g * (xv + yv);
                    gxe += wk * g;
                    gye += wk * g;
                } else if (ty == BMNAS_OP_ATTN) {
                    const float oh = (sm.as[ee] - a_mean) * a_rstd;
                    const float G = __ldg(p.ln_w[k] + ee);
                    dg[k] += g * (oh * G + __ldg(p.ln_b[k] + ee));
                    const float go = wk * g;
                    if (act) {
                        sm.lnG[ee] += go * oh;
                        sm.lnH[ee] += go;
                    }
                    const float q = go * G;
                    lnsum[0] += q;
                    lnsum[1] += q * oh;
                } else {
                    const int zo = p.z_off[k];
                    const bool drop = p.training && p.p_drop[k] > 0.f;
                    const float ds = drop_scale(drop, p.mask[k], p.rng_state, p.op_uid[k], li, gi, p.p_drop[k]);
                    const float zha = (__ldg(Zb + (long long)zo * L + ee) - __ldg(p.mean + zo + c)) * __ldg(p.rstd + zo + c);
                    const float va = zha * __ldg(p.bn_w[k] + c) + __ldg(p.bn_b[k] + c);
                    const float go = wk * g * ds;
                    if (ty == BMNAS_OP_GLU) {
                        const float zhg = (__ldg(Zb + (long long)(zo + C) * L + ee) - __ldg(p.mean + zo + C + c)) *
                                          __ldg(p.rstd + zo + C + c);
                        const float vg = zhg * __ldg(p.bn_w[k] + C + c) + __ldg(p.bn_b[k] + C + c);
                        const float s = sigmoidf_(vg);
                        dg[k] += g * (va * s * ds);
                        const float gva = go * s, gvg = go * va * s * (1.f - s);
                        if (act) {
                            GVb[(long long)zo * L + ee] = gva;
                            GVb[(long long)(zo + C) * L + ee] = gvg;
                        }
                        chan_add<SEG>(sm.S1s, zo + c, gva, L, act);
                        chan_add<SEG>(sm.S2s, zo + c, gva * zha, L, act);
                        chan_add<SEG>(sm.S1s, zo + C + c, gvg, L, act);
                        chan_add<SEG>(sm.S2s, zo + C + c, gvg * zhg, L, act);
                    } else {
                        float o, d;
                        if (ty == BMNAS_OP_FC_RELU) {
                            o = fmaxf(va, 0.f);
                            d = va > 0.f ? 1.f : 0.f;
                        } else {
                            o = mishf_(va);
                            d = mish_grad(va);
                        }
                        dg[k] += g * (o * ds);
                        const float gv = go * d;
                        if (act) GVb[(long long)zo * L + ee] = gv;
                        chan_add<SEG>(sm.S1s, zo + c, gv, L, act);
                        chan_add<SEG>(sm.S2s, zo + c, gv * zha, L, act);
                    }
                }
            }
            if (act) {
                sm.dxs[ee] = gxe;
                sm.dys[ee] = gye;
            }
        }

        if (k_attn >= 0) {
            const int k = k_attn;
            block_sum<2>(lnsum, sm.red);
            const float mq = lnsum[0] / (float)CL, mqo = lnsum[1] / (float)CL;
            const float wk = sm.gw[k];
            const bool drop = p.training && p.p_drop[k] > 0.f;
            for (int e = threadIdx.x; e < CL; e += NTH) {
                const float oh = (sm.as[e] - a_mean) * a_rstd;
                const float q = wk * sm.gs[e] * __ldg(p.ln_w[k] + e);
                float dd = a_rstd * (q - mq - oh * mqo);
                const long long li = (long long)b * CL + e;
                dd *= drop_scale(drop, p.mask[k], p.rng_state, p.op_uid[k], li,
                                 (unsigned long long)(p.sample_offset + b) * CL + e, p.p_drop[k]);
                sm.as[e] = dd;  // dO[c,i]
            }
            __syncthreads();
            lxl_contract(sm.as, sm.ys, sm.Sp, sm.S2, C, L, 1.f);  // dP[i][j] = sum_c dO[c,i] y[c,j]
            for (int i = threadIdx.x; i < L; i += NTH) {
                float rd = 0.f;
                for (int j = 0; j < L; ++j) rd = fmaf(sm.S2[i * L + j], sm.S[i * L + j], rd);
                for (int j = 0; j < L; ++j)
                    sm.S2[i * L + j] = sm.S[i * L + j] * (sm.S2[i * L + j] - rd) * inv_sqrt_c;  // dS / sqrt(C)
            }
            __syncthreads();
            for (int e = threadIdx.x; e < CL; e += NTH) {
                const int c = e / L, i = e - c * L;
                float dx = 0.f, dy = 0.f;
                for (int j = 0; j < L; ++j) dx = fmaf(sm.S2[i * L + j], sm.ys[c * L + j], dx);
                // dy[c, j=i] = sum_i' dO[c,i'] P[i'][j] + x[c,i'] dS[i'][j]
                for (int ii = 0; ii < L; ++ii) {
                    dy = fmaf(sm.as[c * L + ii], sm.S[ii * L + i], dy);
                    dy = fmaf(sm.xs[c * L + ii], sm.S2[ii * L + i], dy);
                }
                sm.dxs[e] += dx;
                sm.dys[e] += dy;
            }
        }
        __syncthreads();
        for (int e = threadIdx.x; e < CL; e += NTH) {
            const long long li = (long long)b * CL + e;
            if (p.alias_xy) {
                if (p.gx) {
                    const float v = sm.dxs[e] + sm.dys[e];
                    p.gx[li] = p.gx_accum ? p.gx[li] + v : v;
                }
            } else {
                if (p.gx) p.gx[li] = p.gx_accum ? p.gx[li] + sm.dxs[e] : sm.dxs[e];
                if (p.gy) p.gy[li] = p.gy_accum ? p.gy[li] + sm.dys[e] : sm.dys[e];
            }
        }
    }

    // ---- per-CTA partials -> global, then last CTA finalises
    block_sum<BMNAS_MAX_OPS>(dg, sm.red);
    __syncthreads();
    const int PW = 2 * M + BMNAS_MAX_OPS;
    float* part = p.partials + (long long)blockIdx.x * PW;
    for (int i = threadIdx.x; i < M; i += NTH) {
        part[i] = sm.S1s[i];
        part[M + i] = sm.S2s[i];
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < BMNAS_MAX_OPS; ++k) part[2 * M + k] = dg[k];
    }
    if (k_attn >= 0 && p.g_ln_w[k_attn]) {
        for (int e = threadIdx.x; e < CL; e += NTH) {
            atomicAdd(p.g_ln_w[k_attn] + e, sm.lnG[e]);
            atomicAdd(p.g_ln_b[k_attn] + e, sm.lnH[e]);
        }
    }
    if (!last_block(p.counter, gridDim.x)) return;

    const float n = (float)p.B * (float)L;
    for (int k = 0; k < p.n_ops; ++k) {
        const int ty = p.op_type[k];
        if (ty == BMNAS_OP_SUM || ty == BMNAS_OP_ATTN) continue;
        const int rows = ty == BMNAS_OP_GLU ? 2 * C : C, zo = p.z_off[k];
        for (int ml = threadIdx.x; ml < rows; ml += NTH) {
            const int m = zo + ml;
            float s1 = 0.f, s2 = 0.f;
            for (unsigned cta = 0; cta < gridDim.x; ++cta) {
                s1 += ld_cg(p.partials + (long long)cta * PW + m);
                s2 += ld_cg(p.partials + (long long)cta * PW + M + m);
            }
            if (p.g_bn_w[k]) {
                p.g_bn_w[k][ml] = s2;
                p.g_bn_b[k][ml] = s1;
            }
            const float rs = p.rstd[m], mu = p.mean[m];
            if (p.training) {
                const float a = p.bn_w[k][ml] * rs, m1 = s1 / n, m2 = s2 / n;
                p.coef_a[m] = a;
                p.coef_b[m] = -a * rs * m2;
                p.coef_c[m] = a * (mu * rs * m2 - m1);
            } else {  // eval-mode BN is a fixed affine map
                p.coef_a[m] = p.bn_w[k][ml] * rs;
                p.coef_b[m] = 0.f;
                p.coef_c[m] = 0.f;
            }
        }
    }
    if (p.g_gamma && threadIdx.x == 0) {
        float d[BMNAS_MAX_OPS];
        float dot = 0.f;
        for (int k = 0; k < p.n_ops; ++k) {
            float s = 0.f;
            for (unsigned cta = 0; cta < gridDim.x; ++cta) s += ld_cg(p.partials + (long long)cta * PW + 2 * M + k);
            d[k] = s;
            dot += sm.gw[k] * s;
        }
        for (int k = 0; k < p.n_ops; ++k) p.g_gamma[k] = p.gamma_is_logits ? sm.gw[k] * (d[k] - dot) : d[k];
    }
}

static int node_check(const bmnas_node_params* p, bool bwd) {
    if (!p || p->B < 1 || p->C < 1 || p->L < 1 || p->L > 64 || p->n_ops < 1 || p->n_ops > BMNAS_MAX_OPS)
        return BMNAS_EINVAL;
    if (!p->x || !p->y) return BMNAS_EINVAL;
    int n_attn = 0;
    for (int k = 0; k < p->n_ops; ++k) {
        const int ty = p->op_type[k];
        if (ty == BMNAS_OP_ATTN) {
            ++n_attn;
            if (!p->ln_w[k] || !p->ln_b[k]) return BMNAS_EINVAL;
        } else if (ty != BMNAS_OP_SUM) {
            if (ty < 0 || ty > BMNAS_OP_FC_MISH) return BMNAS_EINVAL;
            if (!p->Z || !p->mean || !p->rstd || !p->bn_w[k] || !p->bn_b[k]) return BMNAS_EINVAL;
            const int rows = ty == BMNAS_OP_GLU ? 2 * p->C : p->C;
            if (p->z_off[k] < 0 || p->z_off[k] + rows > p->M) return BMNAS_EINVAL;
            if (bwd && (!p->GV || !p->coef_a || !p->coef_b || !p->coef_c)) return BMNAS_EINVAL;
        }
        if (p->training && p->p_drop[k] > 0.f && !p->mask[k] && !p->rng_state && ty != BMNAS_OP_SUM)
            return BMNAS_EINVAL;
        if (p->p_drop[k] < 0.f || p->p_drop[k] >= 1.f) return BMNAS_EINVAL;
    }
    if (n_attn > 1) return BMNAS_EINVAL;
    if (bwd && (!p->gout || !p->partials || !p->counter)) return BMNAS_EINVAL;
    if (!bwd && !p->out) return BMNAS_EINVAL;
    return BMNAS_OK;
}

}  // namespace bmnas

using namespace bmnas;

extern "C" long long bmnas_node_partials_size(const bmnas_node_params* p) {
    return (long long)kNodeMaxBlocksBwd * (2LL * p->M + BMNAS_MAX_OPS);
}

extern "C" int bmnas_node_fwd(const bmnas_node_params* p, void* stream) {
    int e = node_check(p, false);
    if (e) return e;
    const size_t smem = node_smem_floats(p->C, p->L, p->M, false) * sizeof(float);
    if (smem > 227 * 1024) return BMNAS_EINVAL;
    BMNAS_DRY_RETURN();
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        if (cudaFuncSetAttribute(k_node_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return BMNAS_ELAUNCH;
        configured = smem;
    }
    const int blocks = p->B < kNodeMaxBlocksFwd ? p->B : kNodeMaxBlocksFwd;
    k_node_fwd<<<blocks, NTH, smem, (cudaStream_t)stream>>>(*p);
    BMNAS_LAUNCH_CHECK();
    return BMNAS_OK;
}

extern "C" int bmnas_node_bwd(const bmnas_node_params* p, void* stream) {
    int e = node_check(p, true);
    if (e) return e;
    const size_t smem = node_smem_floats(p->C, p->L, p->M, true) * sizeof(float);
    if (smem > 227 * 1024) return BMNAS_EINVAL;
    const bool seg = (p->L & (p->L - 1)) == 0 && p->L <= 32;
    BMNAS_DRY_RETURN();
    static size_t configured[2] = {0, 0};
    if (smem > 48 * 1024 && smem > configured[seg]) {
        cudaError_t ce = seg ? cudaFuncSetAttribute(k_node_bwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                             : cudaFuncSetAttribute(k_node_bwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (ce != cudaSuccess) return BMNAS_ELAUNCH;
        configured[seg] = smem;
    }
    const int blocks = p->B < kNodeMaxBlocksBwd ? p->B : kNodeMaxBlocksBwd;
    if (seg)
        k_node_bwd<true><<<blocks, NTH, smem, (cudaStream_t)stream>>>(*p);
    else
        k_node_bwd<false><<<blocks, NTH, smem, (cudaStream_t)stream>>>(*p);
    BMNAS_LAUNCH_CHECK();
    return BMNAS_OK;
}
