// Step-node mixed op: out = sum_k gamma~_k * op_k(x, y) evaluated per sample from
// the x / y tiles staged once in shared memory (Sum, ScaledDotAttn + LayerNorm,
// LinearGLU and ConcatFC / CatConvMish epilogues on the pre-BN conv output Z),
// softmax(gamma) weighted sum in the epilogue; per-op outputs never reach HBM.
// The backward recomputes the primitives, emits GV = dL/d(BN output) for the conv
// GEMMs, reduces dL/dgamma (warp shuffle -> block -> fixed-order last-block tree),
// the BatchNorm affine grads and the coefficients that fold BatchNorm-backward
// into the conv backward operand loads.
// One CTA per sample (grid-stride over samples), 256 threads, every thread owns
// groups of G=4 consecutive elements (128-bit global traffic, one Philox call per
// group and dropout site); per-channel BatchNorm constants are folded once per CTA
// into shared memory.
#include "common.cuh"

namespace bmnas {

// threads of the CTA-per-sample kernels: 256 while a sample is at most 256 four-element groups (NTU: C*L = 1024), more for
// larger samples -- a CTA owns a whole sample and every phase ends in a block barrier, so at Ego-large (C*L = 4096, 96
// samples on 148 SMs) 256 threads left each SM with 8 warps working through 16 elements per thread and phase: 36 us forward /
// 60 us backward, latency bound (ncu: 12 % warps active, long_sb + short_sb 47 %).  Device code reads the size from blockDim.
constexpr int NTH0 = 256, NTH_FWD_MAX = 1024, NTH_BWD_MAX = 512;    // backward: 80 registers per thread
#define NTH ((int)blockDim.x)
static inline int node_threads(int C, int L, bool bwd, int B) {
    const int groups = (C * L + 3) / 4;
    int n = NTH0;
    if (B > 2 * kNumSMs) return n;     // the machine is full of 256-thread CTAs anyway: wider ones would only cost occupancy
    while (n < groups && n < (bwd ? NTH_BWD_MAX : NTH_FWD_MAX)) n *= 2;
    return n;
}
constexpr int kNodeMaxBlocksFwd = kNumSMs * 8;
constexpr int kNodeMaxBlocksBwd = kNumSMs * 2;

struct NodeSmem {
    float *xs, *ys, *gs, *as, *dxs, *dys, *S, *S2, *Sp, *S1s, *S2s, *lnG, *lnH, *red, *gw;
    float *rs, *mr, *bw, *bb, *tot, *dgs;
};

__host__ __device__ inline size_t rnd4(size_t n) { return (n + 3) & ~(size_t)3; }

__host__ __device__ inline size_t node_smem_floats(int C, int L, int M, bool bwd, int nth) {
    const size_t CL = rnd4((size_t)C * L), LL = rnd4((size_t)L * L), Mr = rnd4((size_t)M);
    size_t n = 0;
    n += 3 * CL;                            // xs, ys, as
    n += 2 * LL + (LL > (size_t)nth ? LL : (size_t)nth);    // S, S2, Sp
    n += 8 * 32 + 8;                        // red, gw
    n += 4 * Mr;                            // rs, mr, bw, bb
    if (bwd) n += 5 * CL + 2 * Mr + (2 * Mr + 8) + BMNAS_MAX_OPS * (size_t)nth;  // gs, dxs, dys, lnG, lnH, S1s, S2s, tot, dgs
    return n + 16;
}

__device__ __forceinline__ NodeSmem node_carve(float* base, int C, int L, int M, bool bwd) {
    const size_t CL = rnd4((size_t)C * L), LL = rnd4((size_t)L * L), Mr = rnd4((size_t)M);
    NodeSmem s;
    float* q = base;
    s.xs = q; q += CL;
    s.ys = q; q += CL;
    s.as = q; q += CL;
    s.S = q; q += LL;
    s.S2 = q; q += LL;
    s.Sp = q; q += (LL > (size_t)NTH ? LL : (size_t)NTH);
    s.red = q; q += 8 * 32;
    s.gw = q; q += 8;
    s.rs = q; q += Mr;
    s.mr = q; q += Mr;
    s.bw = q; q += Mr;
    s.bb = q; q += Mr;
    s.gs = s.dxs = s.dys = s.S1s = s.S2s = s.lnG = s.lnH = s.tot = s.dgs = nullptr;
    if (bwd) {
        s.gs = q; q += CL;
        s.dxs = q; q += CL;
        s.dys = q; q += CL;
        s.lnG = q; q += CL;
        s.lnH = q; q += CL;
        s.S1s = q; q += Mr;
        s.S2s = q; q += Mr;
        s.tot = q; q += 2 * Mr + 8;
        s.dgs = q; q += BMNAS_MAX_OPS * NTH;
    }
    return s;
}

// ---- G-wide element groups -------------------------------------------------------------
template <int G>
__device__ __forceinline__ void ldg_v(const float* p, float (&v)[G]) {
    if (G == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1 % G] = t.y; v[2 % G] = t.z; v[3 % G] = t.w;
    } else {
        v[0] = __ldg(p);
    }
}
template <int G>
__device__ __forceinline__ void lds_v(const float* p, float (&v)[G]) {
    if (G == 4) {
        const float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1 % G] = t.y; v[2 % G] = t.z; v[3 % G] = t.w;
    } else {
        v[0] = *p;
    }
}
template <int G>
__device__ __forceinline__ void st_v(float* p, const float (&v)[G]) {
    if (G == 4) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1 % G], v[2 % G], v[3 % G]);
    else *p = v[0];
}

// dropout scales of one group (0 or 1/(1-p); 1 when inactive).  li/gi = local/global index of its first element.
template <int G>
__device__ __forceinline__ void drop_v(bool active, const unsigned char* mask, const unsigned long long* rng,
                                       uint32_t uid, long long li, unsigned long long gi, float p, float (&ds)[G]) {
    if (!active) {
#pragma unroll
        for (int j = 0; j < G; ++j) ds[j] = 1.f;
        return;
    }
    const float keep = 1.f / (1.f - p);
    if (mask) {
        if (G == 4) {
            const uchar4 m = *reinterpret_cast<const uchar4*>(mask + li);
            ds[0] = m.x ? keep : 0.f; ds[1 % G] = m.y ? keep : 0.f; ds[2 % G] = m.z ? keep : 0.f; ds[3 % G] = m.w ? keep : 0.f;
        } else {
            ds[0] = mask[li] ? keep : 0.f;
        }
    } else if (G == 4) {   // one Philox call covers the group (same stream as the per-element philox_keep)
        const unsigned long long seed = rng[0], step = rng[1];
        const uint2 key = make_uint2((uint32_t)seed ^ (uid * 0x9E3779B1u), (uint32_t)(seed >> 32) + uid);
        const uint4 r = philox4x32(make_uint4((uint32_t)(gi >> 2), (uint32_t)(gi >> 34), (uint32_t)step,
                                              (uint32_t)(step >> 32)), key);
        const float sc = 1.0f / 16777216.0f;
        ds[0] = ((float)(r.x >> 8) * sc >= p) ? keep : 0.f;
        ds[1 % G] = ((float)(r.y >> 8) * sc >= p) ? keep : 0.f;
        ds[2 % G] = ((float)(r.z >> 8) * sc >= p) ? keep : 0.f;
        ds[3 % G] = ((float)(r.w >> 8) * sc >= p) ? keep : 0.f;
    } else {
        ds[0] = philox_keep(rng, uid, gi, p) ? keep : 0.f;
    }
}

// out[i*L+j] = scale * sum_c A[c*L+i] * Bm[c*L+j]   (L x L contraction over channels)
__device__ __forceinline__ void lxl_contract(const float* A, const float* Bm, float* Sp, float* out, int C, int L,
                                             float scale) {
    const int pairs = L * L;
    const int nslice = pairs <= NTH ? NTH / pairs : 1;
    for (int w = threadIdx.x; w < nslice * pairs; w += NTH) {
        const int s = w / pairs, pr = w - s * pairs, i = pr / L, j = pr - i * L;
        float a0 = 0.f, a1 = 0.f;
        int c = s;
        for (; c + nslice < C; c += 2 * nslice) {
            a0 = fmaf(A[c * L + i], Bm[c * L + j], a0);
            a1 = fmaf(A[(c + nslice) * L + i], Bm[(c + nslice) * L + j], a1);
        }
        if (c < C) a0 = fmaf(A[c * L + i], Bm[c * L + j], a0);
        Sp[w] = a0 + a1;
    }
    __syncthreads();
    for (int pr = threadIdx.x; pr < pairs; pr += NTH) {
        float a = 0.f;
        for (int s = 0; s < nslice; ++s) a += Sp[s * pairs + pr];
        out[pr] = a * scale;
    }
    __syncthreads();
}

// G == 4 kernels are only launched on 16-byte aligned tensors with L % 4 == 0 (node_vec_ok): no scalar path in them
template <int G>
__device__ __forceinline__ void load_tile(float* dst, const float* src, int CL) {
    if (G == 4) {
#pragma unroll 2
        for (int i = threadIdx.x; i < CL / 4; i += NTH)
            reinterpret_cast<float4*>(dst)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
    } else {
        for (int i = threadIdx.x; i < CL; i += NTH) dst[i] = __ldg(src + i);
    }
}

// softmax(gamma) (or given weights / ones): architecture parameters are never written inside a forward/backward
// pass, so this is legal before pdl_wait()
__device__ __forceinline__ void node_setup_gamma(const bmnas_node_params& p, const NodeSmem& sm) {
    // rolled loops on purpose: this runs once, in one thread, and every instruction of an execute-once kernel is
    // an instruction-cache miss on the launch's critical path (the unrolled form was 450 SASS instructions)
    if (threadIdx.x == 0) {
        float* gw = sm.gw;
        if (!p.gamma) {
#pragma unroll 1
            for (int k = 0; k < p.n_ops; ++k) gw[k] = 1.f;
        } else if (p.gamma_is_logits) {
            float mx = -INFINITY;
#pragma unroll 1
            for (int k = 0; k < p.n_ops; ++k) mx = fmaxf(mx, p.gamma[k]);
            float s = 0.f;
#pragma unroll 1
            for (int k = 0; k < p.n_ops; ++k) {
                gw[k] = expf(p.gamma[k] - mx);
                s += gw[k];
            }
#pragma unroll 1
            for (int k = 0; k < p.n_ops; ++k) gw[k] /= s;
        } else {
#pragma unroll 1
            for (int k = 0; k < p.n_ops; ++k) gw[k] = p.gamma[k];
        }
    }
}
// skip weights of the chained edge mix (n_chain priors + the op's own output); architecture parameters again
__device__ __forceinline__ void node_setup_chain(const bmnas_node_params& p, float* cw) {
    if (threadIdx.x == 0 && p.chain_w) {
#pragma unroll 1
        for (int j = 0; j <= p.n_chain; ++j) {
            const float a = p.chain_w[2 * j], b = p.chain_w[2 * j + 1];
            cw[j] = p.chain_is_logits ? 1.f / (1.f + expf(a - b)) : b;
        }
    }
}
// the folded per-channel BatchNorm constants (mean / rstd come from the conv kernel right before this one:
// only after pdl_wait())
__device__ __forceinline__ void node_setup_bn(const bmnas_node_params& p, const NodeSmem& sm) {
    for (int k = 0; k < p.n_ops; ++k) {
        const int ty = p.op_type[k];
        if (ty == BMNAS_OP_SUM || ty == BMNAS_OP_ATTN) continue;
        const int rows = ty == BMNAS_OP_GLU ? 2 * p.C : p.C, zo = p.z_off[k];
        for (int ml = threadIdx.x; ml < rows; ml += NTH) {
            const int m = zo + ml;
            const float r = __ldg(p.rstd + m);
            sm.rs[m] = r;
            sm.mr[m] = __ldg(p.mean + m) * r;
            sm.bw[m] = __ldg(p.bn_w[k] + ml);
            sm.bb[m] = __ldg(p.bn_b[k] + ml);
        }
    }
    __syncthreads();
}

// attention forward for one sample: P (L x L) in sm.S, dropped output a[c,i] in sm.as,
// returns LayerNorm statistics of a.  (ScaledDotAttn.forward node_operations.py:92-108)
template <int G>
__device__ __forceinline__ void attn_forward(const bmnas_node_params& p, const NodeSmem& sm, int k, int b,
                                             float* mean_out, float* rstd_out) {
    const int C = p.C, L = p.L, CL = C * L, LL = L * L;
    lxl_contract(sm.xs, sm.ys, sm.Sp, sm.S, C, L, 1.f / sqrtf((float)C));  // S[i][j] = q_i . k_j / sqrt(C)
    // softmax over key positions j, one thread per (i, j)
    for (int pr = threadIdx.x; pr < LL; pr += NTH) {
        const int i = pr / L;
        float mx = -INFINITY;
        for (int j = 0; j < L; ++j) mx = fmaxf(mx, sm.S[i * L + j]);
        sm.S2[pr] = expf(sm.S[pr] - mx);
    }
    __syncthreads();
    for (int pr = threadIdx.x; pr < LL; pr += NTH) {
        const int i = pr / L;
        float s = 0.f;
        for (int j = 0; j < L; ++j) s += sm.S2[i * L + j];
        sm.S[pr] = sm.S2[pr] / s;
    }
    __syncthreads();
    const bool drop = p.training && p.p_drop[k] > 0.f;
    float s0[1] = {0.f}, s1[1] = {0.f};
    for (int g = threadIdx.x; g < CL / G; g += NTH) {
        const int e0 = g * G, c = e0 / L, i0 = e0 - c * L;
        float o[G], ds[G];
#pragma unroll
        for (int q = 0; q < G; ++q) o[q] = 0.f;
        for (int j = 0; j < L; ++j) {
            const float yv = sm.ys[c * L + j];
#pragma unroll
            for (int q = 0; q < G; ++q) o[q] = fmaf(sm.S[(i0 + q) * L + j], yv, o[q]);
        }
        drop_v<G>(drop, p.mask[k], p.rng_state, p.op_uid[k], (long long)b * CL + e0,
                  (unsigned long long)(p.sample_offset + b) * CL + e0, p.p_drop[k], ds);
#pragma unroll
        for (int q = 0; q < G; ++q) {
            o[q] *= ds[q];
            s0[0] += o[q];
        }
        st_v<G>(sm.as + e0, o);
    }
    block_sum<1>(s0, sm.red);
    const float mean = s0[0] / (float)CL;
    for (int g = threadIdx.x; g < CL / G; g += NTH) {
        float o[G];
        lds_v<G>(sm.as + g * G, o);
#pragma unroll
        for (int q = 0; q < G; ++q) {
            const float d = o[q] - mean;
            s1[0] += d * d;
        }
    }
    block_sum<1>(s1, sm.red);
    *mean_out = mean;
    *rstd_out = 1.f / sqrtf(s1[0] / (float)CL + kLnEps);
}

template <int G>
__global__ void __launch_bounds__(NTH_FWD_MAX) k_node_fwd(const bmnas_node_params p) {
    // early section (p.early_ok: x / y were produced at least two kernels ago, i.e. the conv GEMM sits between
    // their producer and this kernel): tile loads and the whole attention primitive of the CTA's first sample run
    // BEFORE pdl_wait(), overlapping the conv kernel; Z / mean / rstd are only touched after it
    bool waited = !p.early_ok;
    if (waited) pdl_prologue();
    extern __shared__ __align__(16) float smem[];
    const int C = p.C, L = p.L, CL = C * L, M = p.M;
    NodeSmem sm = node_carve(smem, C, L, M, false);
    if (p.alias_xy) sm.ys = sm.xs;
    __shared__ float s_cw[BMNAS_MAX_SRC + 1];
    node_setup_gamma(p, sm);
    node_setup_chain(p, s_cw);
    if (waited) node_setup_bn(p, sm);
    int k_attn = -1;
#pragma unroll 1
    for (int k = 0; k < p.n_ops; ++k)
        if (p.op_type[k] == BMNAS_OP_ATTN) k_attn = k;

    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        __syncthreads();
        load_tile<G>(sm.xs, p.x + (long long)b * CL, CL);
        if (!p.alias_xy) load_tile<G>(sm.ys, p.y + (long long)b * CL, CL);
        __syncthreads();
        float a_mean = 0.f, a_rstd = 0.f;
        if (k_attn >= 0) attn_forward<G>(p, sm, k_attn, b, &a_mean, &a_rstd);
        if (!waited) {
            pdl_prologue();
            waited = true;
            node_setup_bn(p, sm);
        }
        const float* Zb = p.Z ? p.Z + (long long)b * M * L : nullptr;
        for (int g = threadIdx.x; g < CL / G; g += NTH) {
            const int e0 = g * G, c = e0 / L;
            const long long li = (long long)b * CL + e0;
            const unsigned long long gi = (unsigned long long)(p.sample_offset + b) * CL + e0;
            float xv[G], yv[G], acc[G];
            lds_v<G>(sm.xs + e0, xv);
            lds_v<G>(sm.ys + e0, yv);
#pragma unroll
            for (int q = 0; q < G; ++q) acc[q] = 0.f;
            for (int k = 0; k < p.n_ops; ++k) {
                const int ty = p.op_type[k];
                const float wk = sm.gw[k];
                float o[G];
                if (ty == BMNAS_OP_SUM) {
#pragma unroll
                    for (int q = 0; q < G; ++q) o[q] = xv[q] + yv[q];
                } else if (ty == BMNAS_OP_ATTN) {
                    float a[G], gw_[G], gb_[G];
                    lds_v<G>(sm.as + e0, a);
                    ldg_v<G>(p.ln_w[k] + e0, gw_);
                    ldg_v<G>(p.ln_b[k] + e0, gb_);
#pragma unroll
                    for (int q = 0; q < G; ++q) o[q] = (a[q] - a_mean) * a_rstd * gw_[q] + gb_[q];
                } else {
                    const int m = p.z_off[k] + c;
                    const bool drop = p.training && p.p_drop[k] > 0.f;
                    float ds[G], z[G];
                    drop_v<G>(drop, p.mask[k], p.rng_state, p.op_uid[k], li, gi, p.p_drop[k], ds);
                    ldg_v<G>(Zb + (long long)p.z_off[k] * L + e0, z);
                    const float r = sm.rs[m], mr = sm.mr[m], w = sm.bw[m], bb = sm.bb[m];
                    if (ty == BMNAS_OP_GLU) {
                        float zg[G];
                        ldg_v<G>(Zb + (long long)(p.z_off[k] + C) * L + e0, zg);
                        const float r2 = sm.rs[m + C], mr2 = sm.mr[m + C], w2 = sm.bw[m + C], bb2 = sm.bb[m + C];
#pragma unroll
                        for (int q = 0; q < G; ++q) {
                            const float va = fmaf(fmaf(z[q], r, -mr), w, bb);
                            const float vg = fmaf(fmaf(zg[q], r2, -mr2), w2, bb2);
                            o[q] = va * sigmoidf_(vg) * ds[q];
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < G; ++q) {
                            const float va = fmaf(fmaf(z[q], r, -mr), w, bb);
                            o[q] = (ty == BMNAS_OP_FC_RELU ? fmaxf(va, 0.f) : mishf_(va)) * ds[q];
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < G; ++q) acc[q] = fmaf(wk, o[q], acc[q]);
            }
            st_v<G>(p.out + li, acc);
            if (p.out2) {                    // the next inner edge mix, from the same registers
                float o2[G];
                const float cl = s_cw[p.n_chain];
#pragma unroll
                for (int q = 0; q < G; ++q) o2[q] = cl * acc[q];
#pragma unroll 1
                for (int j = 0; j < p.n_chain; ++j) {
                    float v[G];
                    ldg_v<G>(p.chain_x[j] + li, v);
                    const float cj = s_cw[j];
#pragma unroll
                    for (int q = 0; q < G; ++q) o2[q] = fmaf(cj, v[q], o2[q]);
                }
                st_v<G>(p.out2 + li, o2);
            }
        }
    }
    if (!waited) pdl_prologue();
}

// ----------------------------------------------------------------------------------------------------------
// Warp-per-sample forward (large batch).  The CTA-per-sample kernel above is latency optimal when the batch is
// smaller than the machine (B=96: one sample per SM, 256 threads cooperate on it, ~10 block barriers per sample);
// once there are more samples than warp slots those barriers and the short dependent phases between them cap it
// near 18 % of the HBM roofline (profiles/r01_v6_bench.json, B=8192).  Here ONE WARP owns a sample from load to
// store: lane `lane` holds the channels c = t*32 + lane (t < T) with all L positions in registers (a channel row
// is L*4 contiguous bytes, so a warp's loads and stores are contiguous kilobyte runs), the x / y tiles and the
// L x L attention matrix live in a private shared-memory slab, and every reduction is a warp shuffle -- no block
// barrier inside the sample loop.  Needs L in {4, 8, 16}, T*L <= 32 register slots per tensor, 128-bit alignment.
// Dropout uses exactly the group-of-4 Philox stream of the kernels above, so a forward through this kernel and a
// backward through k_node_bwd see the same masks.
constexpr int WPC = 8;            // warps (= samples in flight) per CTA
constexpr int WMAXZ = 3;          // conv-backed primitives per mixed op the warp kernel takes (at most one of them a GLU)

// one dropout site, everything that does not depend on the element folded once per kernel
struct DropSite {
    uint2 key;                   // Philox key of (seed, op uid)
    uint32_t step_lo, step_hi;   // rng_state[1]
    uint32_t thr;                // keep iff (bits >> 8) >= thr, thr = ceil(p * 2^24): the same decision as
                                 // (float)(bits >> 8) * 2^-24 >= p in drop_v / philox_keep (both sides exact)
    float keep;                  // 1 / (1 - p)
    int mode;                    // 0 inactive, 1 Philox, 2 injected uint8 mask
    const unsigned char* mask;
};

struct WarpOps {
    float wsum, wattn;
    int has_sum, k_attn, nz, glu_zo;
    int type[WMAXZ], zo[WMAXZ];
    float w[WMAXZ];
    DropSite attn_drop, zdrop[WMAXZ];
};

__device__ __forceinline__ DropSite make_drop_site(const bmnas_node_params& p, int k) {
    DropSite d;
    const float pd = p.p_drop[k];
    d.mode = (p.training && pd > 0.f) ? (p.mask[k] ? 2 : 1) : 0;
    d.mask = p.mask[k];
    d.keep = 1.f / (1.f - pd);
    d.thr = (uint32_t)ceilf(pd * 16777216.0f);
    d.key = make_uint2(0u, 0u);
    d.step_lo = d.step_hi = 0u;
    if (d.mode == 1) {
        const unsigned long long seed = p.rng_state[0], step = p.rng_state[1];
        const uint32_t uid = p.op_uid[k];
        d.key = make_uint2((uint32_t)seed ^ (uid * 0x9E3779B1u), (uint32_t)(seed >> 32) + uid);
        d.step_lo = (uint32_t)step;
        d.step_hi = (uint32_t)(step >> 32);
    }
    return d;
}

// dropout scales of the 4 consecutive elements starting at local index li / global index gi (gi % 4 == 0)
__device__ __forceinline__ void drop4(const DropSite& d, long long li, unsigned long long gi, float (&ds)[4]) {
    if (d.mode == 0) {
        ds[0] = ds[1] = ds[2] = ds[3] = 1.f;
    } else if (d.mode == 1) {
        // inlined: a CALL would first wait for every load in flight (the Z-row prefetch) to land
        const uint4 r = philox4x32_inl(make_uint4((uint32_t)(gi >> 2), (uint32_t)(gi >> 34), d.step_lo, d.step_hi), d.key);
        ds[0] = (r.x >> 8) >= d.thr ? d.keep : 0.f;
        ds[1] = (r.y >> 8) >= d.thr ? d.keep : 0.f;
        ds[2] = (r.z >> 8) >= d.thr ? d.keep : 0.f;
        ds[3] = (r.w >> 8) >= d.thr ? d.keep : 0.f;
    } else {
        const uchar4 m = *reinterpret_cast<const uchar4*>(d.mask + li);
        ds[0] = m.x ? d.keep : 0.f; ds[1] = m.y ? d.keep : 0.f; ds[2] = m.z ? d.keep : 0.f; ds[3] = m.w ? d.keep : 0.f;
    }
}

// sigmoid on the fast paths (ex2.approx / rcp.approx: ~2 ulp, three orders below the 1e-5 parity gate)
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

__host__ __device__ inline size_t node_warp_slab_floats(int C, int L, bool alias) {
    // per warp: x (and y) tile, double buffered (cp.async prefetch of the warp's next sample), the attention output
    // tile, the L x L probability matrix
    return (alias ? 2 : 4) * rnd4((size_t)C * L) + rnd4((size_t)C * L) + rnd4((size_t)L * L);
}
__host__ __device__ inline size_t node_warp_smem_floats(int C, int L, int M, bool alias) {
    // folded BatchNorm constants (2 per conv row), the attention LayerNorm affine (2 x C*L), the per-warp slabs
    return 2 * rnd4((size_t)M) + 2 * rnd4((size_t)C * L) + WPC * node_warp_slab_floats(C, L, alias) + 16;
}

__device__ __forceinline__ float4 lds4(const float* q) { return *reinterpret_cast<const float4*>(q); }
__device__ __forceinline__ float4 ldg4(const float* q) { return __ldg(reinterpret_cast<const float4*>(q)); }
__device__ __forceinline__ void cp_async16(float* dst_smem, const float* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Z rows of one channel for the conv-backed primitives (register prefetch buffer)
template <int Q>
struct ZRows {
    float4 a[WMAXZ][Q];
    float4 g[Q];
};

template <int L, int T>
__global__ void __launch_bounds__(WPC * 32, 2) k_node_fwd_warp(const bmnas_node_params p) {
    pdl_prologue();
    extern __shared__ __align__(16) float smem[];
    __shared__ float s_gw[BMNAS_MAX_OPS];
    __shared__ WarpOps s_ops;
    constexpr int Q = L / 4, KS = 32 / L;     // float4 per channel row; lanes sharing one query row in the score pass
    const int C = p.C, CL = C * L, M = p.M;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool alias = p.alias_xy != 0;
    const size_t CLr = rnd4((size_t)CL), Mr = rnd4((size_t)M);
    float2* bnc = reinterpret_cast<float2*>(smem);      // BatchNorm + affine folded per channel: v = z * A + Bc
    float* lnw = smem + 2 * Mr;                         // LayerNorm affine of the attention primitive, staged once
    float* lnb = lnw + CLr;
    float* slab = lnb + CLr + (size_t)warp * node_warp_slab_floats(C, L, alias);
    const size_t xstride = (alias ? 1 : 2) * CLr;                   // [x | y] per buffer, two buffers
    float* os = slab + (alias ? 2 : 4) * CLr;
    float* Ps = os + CLr;
    {
        NodeSmem sm;
        sm.gw = s_gw;
        node_setup_gamma(p, sm);
    }
    if (threadIdx.x == 0) {                  // same thread that wrote gw: fold the op list into one descriptor
        WarpOps& o = s_ops;
        o.wsum = o.wattn = 0.f;
        o.has_sum = 0; o.k_attn = -1; o.nz = 0; o.glu_zo = -1;
        o.attn_drop.mode = 0;
        for (int k = 0; k < p.n_ops; ++k) {
            const int ty = p.op_type[k];
            if (ty == BMNAS_OP_SUM) { o.wsum += s_gw[k]; o.has_sum = 1; }
            else if (ty == BMNAS_OP_ATTN) { o.wattn = s_gw[k]; o.k_attn = k; o.attn_drop = make_drop_site(p, k); }
            else if (o.nz < WMAXZ) {
                const int z = o.nz++;
                o.type[z] = ty; o.zo[z] = p.z_off[k]; o.w[z] = s_gw[k]; o.zdrop[z] = make_drop_site(p, k);
                if (ty == BMNAS_OP_GLU) o.glu_zo = p.z_off[k];
            }
        }
    }
    for (int k = 0; k < p.n_ops; ++k) {
        const int ty = p.op_type[k];
        if (ty == BMNAS_OP_SUM || ty == BMNAS_OP_ATTN) continue;
        const int rows = ty == BMNAS_OP_GLU ? 2 * C : C, zo = p.z_off[k];
        for (int ml = threadIdx.x; ml < rows; ml += WPC * 32) {
            const float a = __ldg(p.rstd + zo + ml) * __ldg(p.bn_w[k] + ml);
            bnc[zo + ml] = make_float2(a, fmaf(-__ldg(p.mean + zo + ml), a, __ldg(p.bn_b[k] + ml)));
        }
    }
    for (int k = 0; k < p.n_ops; ++k) {
        if (p.op_type[k] != BMNAS_OP_ATTN) continue;
        for (int e = threadIdx.x * 4; e < CL; e += WPC * 32 * 4) {
            *reinterpret_cast<float4*>(lnw + e) = ldg4(p.ln_w[k] + e);
            *reinterpret_cast<float4*>(lnb + e) = ldg4(p.ln_b[k] + e);
        }
    }
    __syncthreads();
    const int k_attn = s_ops.k_attn, nz = s_ops.nz, glu_zo = s_ops.glu_zo;
    const bool has_sum = s_ops.has_sum != 0;
    const float wsum = s_ops.wsum, wattn = s_ops.wattn;
    const float inv_sqrt_c = 1.f / sqrtf((float)C);
    const int gwarp = blockIdx.x * WPC + warp, gstride = gridDim.x * WPC;

    auto stage = [&](int b, float* dst) {    // the warp's next sample: global -> shared, no registers, no wait
        const long long base = (long long)b * CL;
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const int c = t * 32 + lane;
            if (c < C) {
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    cp_async16(dst + c * L + 4 * q, p.x + base + c * L + 4 * q);
                    if (!alias) cp_async16(dst + CLr + c * L + 4 * q, p.y + base + c * L + 4 * q);
                }
            }
        }
    };
    auto load_z = [&](const float* Zb, int c, ZRows<Q>& z) {
#pragma unroll
        for (int zi = 0; zi < WMAXZ; ++zi) {
            if (zi < nz) {
                const int zo = s_ops.zo[zi];
#pragma unroll
                for (int q = 0; q < Q; ++q) z.a[zi][q] = ldg4(Zb + (long long)(zo + c) * L + 4 * q);
            }
        }
        if (glu_zo >= 0) {                   // the gate rows of the (single) LinearGLU
#pragma unroll
            for (int q = 0; q < Q; ++q) z.g[q] = ldg4(Zb + (long long)(glu_zo + C + c) * L + 4 * q);
        }
    };

    int buf = 0;
    if (gwarp < p.B) stage(gwarp, slab);
    for (int b = gwarp; b < p.B; b += gstride, buf ^= 1) {
        const long long base = (long long)b * CL;
        const unsigned long long gbase = (unsigned long long)(p.sample_offset + b) * CL;
        const float* xs = slab + buf * xstride;
        const float* ys = alias ? xs : xs + CLr;
        const float* Zb = p.Z ? p.Z + (long long)b * M * L : nullptr;
        cp_async_wait_all();
        __syncwarp();                        // this sample's tiles are visible to every lane; the other buffer is free
        if (b + gstride < p.B) stage(b + gstride, slab + (buf ^ 1) * xstride);

        ZRows<Q> zn;                          // Z rows of the lane's first channel: in flight across the attention tail
        float a_mean = 0.f, a_rstd = 0.f;
        if (k_attn >= 0) {
            // ---- scores: lane = (query position i, channel slice qs); S[i][:] over the slice, then over the KS lanes
            const int i = lane / KS, qs = lane % KS;
            float sc[L];
#pragma unroll
            for (int j = 0; j < L; ++j) sc[j] = 0.f;
#pragma unroll 4
            for (int c = qs; c < C; c += KS) {
                const float a = xs[c * L + i];
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    const float4 v = lds4(ys + c * L + 4 * q);
                    sc[4 * q] = fmaf(a, v.x, sc[4 * q]);
                    sc[4 * q + 1] = fmaf(a, v.y, sc[4 * q + 1]);
                    sc[4 * q + 2] = fmaf(a, v.z, sc[4 * q + 2]);
                    sc[4 * q + 3] = fmaf(a, v.w, sc[4 * q + 3]);
                }
            }
#pragma unroll
            for (int o = KS / 2; o > 0; o >>= 1)
#pragma unroll
                for (int j = 0; j < L; ++j) sc[j] += __shfl_xor_sync(0xffffffffu, sc[j], o);
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < L; ++j) {
                sc[j] *= inv_sqrt_c;
                mx = fmaxf(mx, sc[j]);
            }
            float den = 0.f;
#pragma unroll
            for (int j = 0; j < L; ++j) {
                sc[j] = expf(sc[j] - mx);
                den += sc[j];
            }
            if (qs == 0) {
                const float inv = 1.f / den;
#pragma unroll
                for (int q = 0; q < Q; ++q)
                    *reinterpret_cast<float4*>(Ps + i * L + 4 * q) =
                        make_float4(sc[4 * q] * inv, sc[4 * q + 1] * inv, sc[4 * q + 2] * inv, sc[4 * q + 3] * inv);
            }
            __syncwarp();
            // ---- O[c][i] = sum_j P[i][j] y[c][j] for the lane's own channels
            float O[T][L];
            {
                float yr[T][L];
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    const int c = t * 32 + lane;
#pragma unroll
                    for (int q = 0; q < Q; ++q) {
                        const float4 v = c < C ? lds4(ys + c * L + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
                        yr[t][4 * q] = v.x; yr[t][4 * q + 1] = v.y; yr[t][4 * q + 2] = v.z; yr[t][4 * q + 3] = v.w;
                    }
                }
#pragma unroll
                for (int i2 = 0; i2 < L; ++i2) {
                    float pr[L];
#pragma unroll
                    for (int q = 0; q < Q; ++q) {
                        const float4 v = lds4(Ps + i2 * L + 4 * q);
                        pr[4 * q] = v.x; pr[4 * q + 1] = v.y; pr[4 * q + 2] = v.z; pr[4 * q + 3] = v.w;
                    }
#pragma unroll
                    for (int t = 0; t < T; ++t) {
                        float a = 0.f;
#pragma unroll
                        for (int j = 0; j < L; ++j) a = fmaf(pr[j], yr[t][j], a);
                        O[t][i2] = a;
                    }
                }
            }
            if (nz > 0 && lane < C) load_z(Zb, lane, zn);
            // ---- dropout(0.1) on the attention output, LayerNorm statistics over the whole (C, L) sample
            const DropSite dsite = s_ops.attn_drop;
            float s0 = 0.f;
#pragma unroll
            for (int t = 0; t < T; ++t) {
                const int c = t * 32 + lane;
                if (c < C) {
#pragma unroll
                    for (int q = 0; q < Q; ++q) {
                        float ds[4];
                        const int e0 = c * L + 4 * q;
                        drop4(dsite, base + e0, gbase + e0, ds);
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            O[t][4 * q + r] *= ds[r];
                            s0 += O[t][4 * q + r];
                        }
                        *reinterpret_cast<float4*>(os + e0) =
                            make_float4(O[t][4 * q], O[t][4 * q + 1], O[t][4 * q + 2], O[t][4 * q + 3]);
                    }
                }
            }
            a_mean = warp_sum(s0) / (float)CL;
            float s1 = 0.f;
#pragma unroll
            for (int t = 0; t < T; ++t) {
                if (t * 32 + lane < C) {
#pragma unroll
                    for (int j = 0; j < L; ++j) {
                        const float d = O[t][j] - a_mean;
                        s1 = fmaf(d, d, s1);
                    }
                }
            }
            a_rstd = 1.f / sqrtf(warp_sum(s1) / (float)CL + kLnEps);
        } else if (nz > 0 && lane < C) {
            load_z(Zb, lane, zn);
        }

        // ---- epilogue: every primitive at the lane's own elements (the lane wrote os[] itself: no sync needed),
        //      softmax(gamma)-weighted sum, one store.  Rolled over the lane's channels; the Z rows of the next
        //      channel are requested before the current one is evaluated.
#pragma unroll 1
        for (int t = 0; t < T; ++t) {
            const int c = t * 32 + lane;
            if (c >= C) break;
            ZRows<Q> z = zn;
            if (nz > 0 && t + 1 < T && c + 32 < C) load_z(Zb, c + 32, zn);
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const int e0 = c * L + 4 * q;
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                if (has_sum) {
                    const float4 xv = lds4(xs + e0), yv = lds4(ys + e0);
                    acc[0] = wsum * (xv.x + yv.x); acc[1] = wsum * (xv.y + yv.y);
                    acc[2] = wsum * (xv.z + yv.z); acc[3] = wsum * (xv.w + yv.w);
                }
                if (k_attn >= 0) {
                    const float4 lw = lds4(lnw + e0), lb = lds4(lnb + e0), ov = lds4(os + e0);
                    acc[0] = fmaf(wattn, fmaf((ov.x - a_mean) * a_rstd, lw.x, lb.x), acc[0]);
                    acc[1] = fmaf(wattn, fmaf((ov.y - a_mean) * a_rstd, lw.y, lb.y), acc[1]);
                    acc[2] = fmaf(wattn, fmaf((ov.z - a_mean) * a_rstd, lw.z, lb.z), acc[2]);
                    acc[3] = fmaf(wattn, fmaf((ov.w - a_mean) * a_rstd, lw.w, lb.w), acc[3]);
                }
#pragma unroll
                for (int zi = 0; zi < WMAXZ; ++zi) {
                    if (zi < nz) {
                        const int ty = s_ops.type[zi], m = s_ops.zo[zi] + c;
                        const float wk = s_ops.w[zi];
                        float ds[4];
                        drop4(s_ops.zdrop[zi], base + e0, gbase + e0, ds);
                        const float zv[4] = {z.a[zi][q].x, z.a[zi][q].y, z.a[zi][q].z, z.a[zi][q].w};
                        const float2 ab = bnc[m];
                        if (ty == BMNAS_OP_GLU) {
                            const float g[4] = {z.g[q].x, z.g[q].y, z.g[q].z, z.g[q].w};
                            const float2 ab2 = bnc[m + C];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float va = fmaf(zv[e], ab.x, ab.y);
                                const float vg = fmaf(g[e], ab2.x, ab2.y);
                                acc[e] = fmaf(wk * ds[e], va * sigmoid_fast(vg), acc[e]);
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float va = fmaf(zv[e], ab.x, ab.y);
                                acc[e] = fmaf(wk * ds[e], ty == BMNAS_OP_FC_RELU ? fmaxf(va, 0.f) : mishf_(va), acc[e]);
                            }
                        }
                    }
                }
                *reinterpret_cast<float4*>(p.out + base + e0) = make_float4(acc[0], acc[1], acc[2], acc[3]);
            }
        }
    }
    cp_async_wait_all();
}

// ----------------------------------------------------------------------------------------------------------
// Warp-per-sample backward (large batch; x is y in the searchable cell, x != y in the found cell).  Same ownership
// as the forward: lane holds the
// channels c = t*32 + lane.  Per sample the warp (A) stages x and gout in its slab, (B) recomputes the attention
// primitive exactly as the forward did (P, dropped output, LayerNorm statistics; the keep bits of its 32 own
// elements stay in one register), (C) walks its own elements once: d(gamma) partials, GV rows for the conv
// backward, per-channel BatchNorm sums and the LayerNorm affine gradients (shared-memory atomics on CTA-wide
// accumulators: the lane owns the address within its warp, warps collide rarely), (D) LayerNorm backward ->
// dO, dP = dO^T y, dS, (E) dx + dy for its channel rows and the single gx store.  CTA-level sums go to the global
// accumulator with one atomic per value, the last CTA finalises exactly like k_node_bwd.
constexpr int WPCB = 8;           // warps per CTA of the backward (one CTA per SM; 204 registers per thread for the 64 LN-affine accumulators)

__host__ __device__ inline size_t node_bwarp_slab_floats(int C, int L, int M) {
    // x and gout tiles (double buffered: cp.async prefetch of the warp's next sample), a/dO tile; P, dS; the warp's
    // private per-channel BatchNorm sums S1 | S2
    return 5 * rnd4((size_t)C * L) + 2 * rnd4((size_t)L * L) + 2 * rnd4((size_t)M);
}
__host__ __device__ inline size_t node_bwarp_smem_floats(int C, int L, int M) {
    // (rstd, mean*rstd) and (w, b) per conv row; S1 | S2 totals (last CTA); LN weight; slabs
    return 6 * rnd4((size_t)M) + rnd4((size_t)C * L) + WPCB * node_bwarp_slab_floats(C, L, M) + 16;
}

template <int L, int T>
__global__ void __launch_bounds__(WPCB * 32, 1) k_node_bwd_warp(const bmnas_node_params p) {
    pdl_prologue();
    extern __shared__ __align__(16) float smem[];
    __shared__ float s_gw[BMNAS_MAX_OPS];
    __shared__ float s_dg[BMNAS_MAX_OPS];
    __shared__ WarpOps s_ops;
    constexpr int Q = L / 4, KS = 32 / L;
    const int C = p.C, CL = C * L, M = p.M;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool alias = p.alias_xy != 0;
    const size_t CLr = rnd4((size_t)CL), Mr = rnd4((size_t)M), LLr = rnd4((size_t)L * L);
    float2* bnr = reinterpret_cast<float2*>(smem);            // (rstd, mean * rstd)
    float2* bnw = bnr + Mr;                                   // (weight, bias)
    float* S1s = smem + 4 * Mr;
    float* S2s = S1s + Mr;
    float* lnw = S2s + Mr;
    float* slab0 = lnw + CLr;
    const size_t slab_floats = node_bwarp_slab_floats(C, L, M);
    // x is y (searchable cell): [x0 | g0 | x1 | g1 | a/dO | P | dS | S1 | S2], tiles double buffered;
    // x is not y (found cell):  [x  | y  | g  | -- | a/dO | P | dS | S1 | S2], single buffered
    float* slab = slab0 + (size_t)warp * slab_floats;
    float* os = slab + 4 * CLr;
    float* Ps = os + CLr;
    float* dSs = Ps + LLr;
    float* wS1 = dSs + LLr;                                   // this warp's S1 | S2 (lane-owned rows: no atomics)
    float* wS2 = wS1 + Mr;
    {
        NodeSmem sm;
        sm.gw = s_gw;
        node_setup_gamma(p, sm);
    }
    if (threadIdx.x == 0) {
        WarpOps& o = s_ops;
        o.wsum = o.wattn = 0.f;
        o.has_sum = 0; o.k_attn = -1; o.nz = 0; o.glu_zo = -1;
        o.attn_drop.mode = 0;
        for (int k = 0; k < p.n_ops; ++k) {
            const int ty = p.op_type[k];
            if (ty == BMNAS_OP_SUM) { o.wsum += s_gw[k]; o.has_sum = 1; }
            else if (ty == BMNAS_OP_ATTN) { o.wattn = s_gw[k]; o.k_attn = k; o.attn_drop = make_drop_site(p, k); }
            else if (o.nz < WMAXZ) {
                const int z = o.nz++;
                o.type[z] = ty; o.zo[z] = p.z_off[k]; o.w[z] = s_gw[k]; o.zdrop[z] = make_drop_site(p, k);
                if (ty == BMNAS_OP_GLU) o.glu_zo = p.z_off[k];
            }
        }
    }
    if (threadIdx.x < BMNAS_MAX_OPS) s_dg[threadIdx.x] = 0.f;
    for (int k = 0; k < p.n_ops; ++k) {
        const int ty = p.op_type[k];
        if (ty == BMNAS_OP_ATTN) {
            for (int e = threadIdx.x * 4; e < CL; e += WPCB * 32 * 4)
                *reinterpret_cast<float4*>(lnw + e) = ldg4(p.ln_w[k] + e);
            continue;
        }
        if (ty == BMNAS_OP_SUM) continue;
        const int rows = ty == BMNAS_OP_GLU ? 2 * C : C, zo = p.z_off[k];
        for (int ml = threadIdx.x; ml < rows; ml += WPCB * 32) {
            const float r = __ldg(p.rstd + zo + ml);
            bnr[zo + ml] = make_float2(r, __ldg(p.mean + zo + ml) * r);
            bnw[zo + ml] = make_float2(__ldg(p.bn_w[k] + ml), __ldg(p.bn_b[k] + ml));
        }
    }
    for (int i = threadIdx.x; i < M; i += WPCB * 32) {
        S1s[i] = 0.f;
        S2s[i] = 0.f;
    }
    for (int i = lane; i < 2 * (int)Mr; i += 32) wS1[i] = 0.f;
    __syncthreads();
    const int k_attn = s_ops.k_attn, nz = s_ops.nz, glu_zo = s_ops.glu_zo;
    const bool has_sum = s_ops.has_sum != 0;
    const float wsum = s_ops.wsum, wattn = s_ops.wattn;
    const float inv_sqrt_c = 1.f / sqrtf((float)C);
    const int gwarp = blockIdx.x * WPCB + warp, gstride = gridDim.x * WPCB;
    const float* lnb_g = k_attn >= 0 ? p.ln_b[k_attn] : nullptr;
    float dg_sum = 0.f, dg_attn = 0.f, dg_z[WMAXZ];
#pragma unroll
    for (int zi = 0; zi < WMAXZ; ++zi) dg_z[zi] = 0.f;
    // LayerNorm affine gradients of the lane's own elements, summed over the warp's samples in registers
    float aG[T][L], aH[T][L];
#pragma unroll
    for (int t = 0; t < T; ++t)
#pragma unroll
        for (int j = 0; j < L; ++j) aG[t][j] = aH[t][j] = 0.f;

    auto load_z = [&](const float* Zb, int c, ZRows<Q>& z) {
#pragma unroll
        for (int zi = 0; zi < WMAXZ; ++zi) {
            if (zi < nz) {
                const int zo = s_ops.zo[zi];
#pragma unroll
                for (int q = 0; q < Q; ++q) z.a[zi][q] = ldg4(Zb + (long long)(zo + c) * L + 4 * q);
            }
        }
        if (glu_zo >= 0) {
#pragma unroll
            for (int q = 0; q < Q; ++q) z.g[q] = ldg4(Zb + (long long)(glu_zo + C + c) * L + 4 * q);
        }
    };

    auto stage = [&](int b, float* dst) {    // x (, y) and gout of one sample: global -> shared, asynchronously
        const long long sb = (long long)b * CL;
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const int c = t * 32 + lane;
            if (c < C) {
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    cp_async16(dst + c * L + 4 * q, p.x + sb + c * L + 4 * q);
                    if (alias) {
                        cp_async16(dst + CLr + c * L + 4 * q, p.gout + sb + c * L + 4 * q);
                    } else {
                        cp_async16(dst + CLr + c * L + 4 * q, p.y + sb + c * L + 4 * q);
                        cp_async16(dst + 2 * CLr + c * L + 4 * q, p.gout + sb + c * L + 4 * q);
                    }
                }
            }
        }
    };
    int buf = 0;
    if (alias && gwarp < p.B) stage(gwarp, slab);
    for (int b = gwarp; b < p.B; b += gstride, buf ^= 1) {
        const long long base = (long long)b * CL;
        const unsigned long long gbase = (unsigned long long)(p.sample_offset + b) * CL;
        const float* Zb = p.Z ? p.Z + (long long)b * M * L : nullptr;
        float* GVb = p.GV ? p.GV + (long long)b * M * L : nullptr;
        const float* xs = alias ? slab + buf * 2 * CLr : slab;
        const float* ys = alias ? xs : slab + CLr;
        const float* gs = alias ? xs + CLr : slab + 2 * CLr;
        if (!alias) {
            __syncwarp();                    // the previous sample's readers of the single-buffered tiles are done
            stage(b, slab);
        }
        cp_async_wait_all();
        __syncwarp();                        // this sample's tiles are visible; the other buffer and os / P / dS are free
        if (alias && b + gstride < p.B) stage(b + gstride, slab + (buf ^ 1) * 2 * CLr);

        // ---- (B) attention primitive, recomputed as in the forward
        float a_mean = 0.f, a_rstd = 0.f;
        uint32_t keepbits = 0xffffffffu;     // bit (t*L + l): element kept by the attention dropout
        if (k_attn >= 0) {
            const int i = lane / KS, qs = lane % KS;
            float sc[L];
#pragma unroll
            for (int j = 0; j < L; ++j) sc[j] = 0.f;
#pragma unroll 4
            for (int c = qs; c < C; c += KS) {
                const float a = xs[c * L + i];
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    const float4 v = lds4(ys + c * L + 4 * q);
                    sc[4 * q] = fmaf(a, v.x, sc[4 * q]);
                    sc[4 * q + 1] = fmaf(a, v.y, sc[4 * q + 1]);
                    sc[4 * q + 2] = fmaf(a, v.z, sc[4 * q + 2]);
                    sc[4 * q + 3] = fmaf(a, v.w, sc[4 * q + 3]);
                }
            }
#pragma unroll
            for (int o = KS / 2; o > 0; o >>= 1)
#pragma unroll
                for (int j = 0; j < L; ++j) sc[j] += __shfl_xor_sync(0xffffffffu, sc[j], o);
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < L; ++j) {
                sc[j] *= inv_sqrt_c;
                mx = fmaxf(mx, sc[j]);
            }
            float den = 0.f;
#pragma unroll
            for (int j = 0; j < L; ++j) {
                sc[j] = expf(sc[j] - mx);
                den += sc[j];
            }
            if (qs == 0) {
                const float inv = 1.f / den;
#pragma unroll
                for (int q = 0; q < Q; ++q)
                    *reinterpret_cast<float4*>(Ps + i * L + 4 * q) =
                        make_float4(sc[4 * q] * inv, sc[4 * q + 1] * inv, sc[4 * q + 2] * inv, sc[4 * q + 3] * inv);
            }
            __syncwarp();
            float O[T][L];
            {
                float yr[T][L];
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    const int c = t * 32 + lane;
#pragma unroll
                    for (int q = 0; q < Q; ++q) {
                        const float4 v = c < C ? lds4(ys + c * L + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
                        yr[t][4 * q] = v.x; yr[t][4 * q + 1] = v.y; yr[t][4 * q + 2] = v.z; yr[t][4 * q + 3] = v.w;
                    }
                }
#pragma unroll
                for (int i2 = 0; i2 < L; ++i2) {
                    float pr[L];
#pragma unroll
                    for (int q = 0; q < Q; ++q) {
                        const float4 v = lds4(Ps + i2 * L + 4 * q);
                        pr[4 * q] = v.x; pr[4 * q + 1] = v.y; pr[4 * q + 2] = v.z; pr[4 * q + 3] = v.w;
                    }
#pragma unroll
                    for (int t = 0; t < T; ++t) {
                        float a = 0.f;
#pragma unroll
                        for (int j = 0; j < L; ++j) a = fmaf(pr[j], yr[t][j], a);
                        O[t][i2] = a;
                    }
                }
            }
            const DropSite dsite = s_ops.attn_drop;
            float s0 = 0.f;
            keepbits = 0u;
#pragma unroll
            for (int t = 0; t < T; ++t) {
                const int c = t * 32 + lane;
                if (c < C) {
#pragma unroll
                    for (int q = 0; q < Q; ++q) {
                        float ds[4];
                        const int e0 = c * L + 4 * q;
                        drop4(dsite, base + e0, gbase + e0, ds);
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            O[t][4 * q + r] *= ds[r];
                            s0 += O[t][4 * q + r];
                            if (ds[r] != 0.f) keepbits |= 1u << (t * L + 4 * q + r);
                        }
                        *reinterpret_cast<float4*>(os + e0) =
                            make_float4(O[t][4 * q], O[t][4 * q + 1], O[t][4 * q + 2], O[t][4 * q + 3]);
                    }
                }
            }
            a_mean = warp_sum(s0) / (float)CL;
            float s1 = 0.f;
#pragma unroll
            for (int t = 0; t < T; ++t) {
                if (t * 32 + lane < C) {
#pragma unroll
                    for (int j = 0; j < L; ++j) {
                        const float d = O[t][j] - a_mean;
                        s1 = fmaf(d, d, s1);
                    }
                }
            }
            a_rstd = 1.f / sqrtf(warp_sum(s1) / (float)CL + kLnEps);
        }

        // ---- (C1) attention LayerNorm: d(gamma) partial, affine gradients (registers), the two LayerNorm-backward sums
        ZRows<Q> zn;                          // Z rows of the lane's first channel travel while C1 runs
        if (nz > 0 && lane < C) load_z(Zb, lane, zn);
        float lnsum0 = 0.f, lnsum1 = 0.f;
        if (k_attn >= 0) {
#pragma unroll
            for (int t = 0; t < T; ++t) {
                const int c = t * 32 + lane;
                if (c < C) {
#pragma unroll
                    for (int q = 0; q < Q; ++q) {
                        const int e0 = c * L + 4 * q;
                        const float4 g4 = lds4(gs + e0), a4 = lds4(os + e0), w4 = lds4(lnw + e0), b4 = ldg4(lnb_g + e0);
                        const float gv[4] = {g4.x, g4.y, g4.z, g4.w}, av[4] = {a4.x, a4.y, a4.z, a4.w};
                        const float wv[4] = {w4.x, w4.y, w4.z, w4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float oh = (av[e] - a_mean) * a_rstd;
                            dg_attn = fmaf(gv[e], fmaf(oh, wv[e], bv[e]), dg_attn);
                            const float go = wattn * gv[e];
                            aG[t][4 * q + e] = fmaf(go, oh, aG[t][4 * q + e]);
                            aH[t][4 * q + e] += go;
                            const float qq = go * wv[e];
                            lnsum0 += qq;
                            lnsum1 = fmaf(qq, oh, lnsum1);
                        }
                    }
                }
            }
        }
        // ---- (C2) conv-backed primitives and Sum: d(gamma) partials, GV rows, per-channel BatchNorm sums
#pragma unroll 1
        for (int t = 0; t < T; ++t) {
            const int c = t * 32 + lane;
            if (c >= C) break;
            ZRows<Q> z = zn;
            if (nz > 0 && t + 1 < T && c + 32 < C) load_z(Zb, c + 32, zn);
            float s1r[WMAXZ], s2r[WMAXZ], s1g = 0.f, s2g = 0.f;
#pragma unroll
            for (int zi = 0; zi < WMAXZ; ++zi) s1r[zi] = s2r[zi] = 0.f;
            float4 g4q[Q];
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                g4q[q] = lds4(gs + c * L + 4 * q);
                if (has_sum) {
                    const float4 xv = lds4(xs + c * L + 4 * q), yv = lds4(ys + c * L + 4 * q);
                    dg_sum += g4q[q].x * (xv.x + yv.x) + g4q[q].y * (xv.y + yv.y) + g4q[q].z * (xv.z + yv.z) + g4q[q].w * (xv.w + yv.w);
                }
            }
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const int e0 = c * L + 4 * q;
                const float gv[4] = {g4q[q].x, g4q[q].y, g4q[q].z, g4q[q].w};
#pragma unroll
                for (int zi = 0; zi < WMAXZ; ++zi) {
                    if (zi < nz) {
                        const int ty = s_ops.type[zi], zo = s_ops.zo[zi], m = zo + c;
                        const float wk = s_ops.w[zi];
                        float ds[4];
                        drop4(s_ops.zdrop[zi], base + e0, gbase + e0, ds);
                        const float zv[4] = {z.a[zi][q].x, z.a[zi][q].y, z.a[zi][q].z, z.a[zi][q].w};
                        const float2 rm = bnr[m], wb = bnw[m];
                        float gva[4];
                        if (ty == BMNAS_OP_GLU) {
                            const float g[4] = {z.g[q].x, z.g[q].y, z.g[q].z, z.g[q].w};
                            const float2 rm2 = bnr[m + C], wb2 = bnw[m + C];
                            float gvg[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float zha = fmaf(zv[e], rm.x, -rm.y), zhg = fmaf(g[e], rm2.x, -rm2.y);
                                const float va = fmaf(zha, wb.x, wb.y), vg = fmaf(zhg, wb2.x, wb2.y);
                                const float sg = sigmoid_fast(vg);
                                dg_z[zi] += gv[e] * (va * sg * ds[e]);
                                const float go = wk * gv[e] * ds[e];
                                gva[e] = go * sg;
                                gvg[e] = go * va * sg * (1.f - sg);
                                s1r[zi] += gva[e]; s2r[zi] = fmaf(gva[e], zha, s2r[zi]);
                                s1g += gvg[e]; s2g = fmaf(gvg[e], zhg, s2g);
                            }
                            *reinterpret_cast<float4*>(GVb + (long long)(zo + C + c) * L + 4 * q) = make_float4(gvg[0], gvg[1], gvg[2], gvg[3]);
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float zha = fmaf(zv[e], rm.x, -rm.y);
                                const float va = fmaf(zha, wb.x, wb.y);
                                float o, d;
                                if (ty == BMNAS_OP_FC_RELU) {
                                    o = fmaxf(va, 0.f);
                                    d = va > 0.f ? 1.f : 0.f;
                                } else {
                                    o = mishf_(va);
                                    d = mish_grad(va);
                                }
                                dg_z[zi] += gv[e] * (o * ds[e]);
                                gva[e] = wk * gv[e] * ds[e] * d;
                                s1r[zi] += gva[e]; s2r[zi] = fmaf(gva[e], zha, s2r[zi]);
                            }
                        }
                        *reinterpret_cast<float4*>(GVb + (long long)m * L + 4 * q) = make_float4(gva[0], gva[1], gva[2], gva[3]);
                    }
                }
            }
#pragma unroll
            for (int zi = 0; zi < WMAXZ; ++zi) {
                if (zi < nz) {                       // lane-owned rows of the warp's private sums
                    wS1[s_ops.zo[zi] + c] += s1r[zi];
                    wS2[s_ops.zo[zi] + c] += s2r[zi];
                }
            }
            if (glu_zo >= 0) {
                wS1[glu_zo + C + c] += s1g;
                wS2[glu_zo + C + c] += s2g;
            }
        }

        // ---- (D) LayerNorm backward -> dO (into os), dP, dS
        if (k_attn >= 0) {
            const float mq = warp_sum(lnsum0) / (float)CL, mqo = warp_sum(lnsum1) / (float)CL;
            const float keep = s_ops.attn_drop.mode ? s_ops.attn_drop.keep : 1.f;
#pragma unroll
            for (int t = 0; t < T; ++t) {
                const int c = t * 32 + lane;
                if (c < C) {
#pragma unroll
                    for (int q = 0; q < Q; ++q) {
                        const int e0 = c * L + 4 * q;
                        const float4 a4 = lds4(os + e0), w4 = lds4(lnw + e0), g4 = lds4(gs + e0);
                        const float av[4] = {a4.x, a4.y, a4.z, a4.w}, wv[4] = {w4.x, w4.y, w4.z, w4.w}, gv[4] = {g4.x, g4.y, g4.z, g4.w};
                        float dd[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float oh = (av[e] - a_mean) * a_rstd;
                            const float sc_ = ((keepbits >> (t * L + 4 * q + e)) & 1u) ? keep : 0.f;
                            dd[e] = a_rstd * (wattn * gv[e] * wv[e] - mq - oh * mqo) * sc_;
                        }
                        *reinterpret_cast<float4*>(os + e0) = make_float4(dd[0], dd[1], dd[2], dd[3]);
                    }
                }
            }
            __syncwarp();
            const int i = lane / KS, qs = lane % KS;
            float dp[L];
#pragma unroll
            for (int j = 0; j < L; ++j) dp[j] = 0.f;
#pragma unroll 4
            for (int c = qs; c < C; c += KS) {        // dP[i][j] = sum_c dO[c,i] y[c,j]
                const float a = os[c * L + i];
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    const float4 v = lds4(ys + c * L + 4 * q);
                    dp[4 * q] = fmaf(a, v.x, dp[4 * q]);
                    dp[4 * q + 1] = fmaf(a, v.y, dp[4 * q + 1]);
                    dp[4 * q + 2] = fmaf(a, v.z, dp[4 * q + 2]);
                    dp[4 * q + 3] = fmaf(a, v.w, dp[4 * q + 3]);
                }
            }
#pragma unroll
            for (int o = KS / 2; o > 0; o >>= 1)
#pragma unroll
                for (int j = 0; j < L; ++j) dp[j] += __shfl_xor_sync(0xffffffffu, dp[j], o);
            if (qs == 0) {                            // dS = P o (dP - <dP, P>_row) / sqrt(C)
                float pr[L];
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    const float4 v = lds4(Ps + i * L + 4 * q);
                    pr[4 * q] = v.x; pr[4 * q + 1] = v.y; pr[4 * q + 2] = v.z; pr[4 * q + 3] = v.w;
                }
                float rd = 0.f;
#pragma unroll
                for (int j = 0; j < L; ++j) rd = fmaf(dp[j], pr[j], rd);
#pragma unroll
                for (int q = 0; q < Q; ++q)
                    *reinterpret_cast<float4*>(dSs + i * L + 4 * q) =
                        make_float4(pr[4 * q] * (dp[4 * q] - rd) * inv_sqrt_c, pr[4 * q + 1] * (dp[4 * q + 1] - rd) * inv_sqrt_c,
                                    pr[4 * q + 2] * (dp[4 * q + 2] - rd) * inv_sqrt_c, pr[4 * q + 3] * (dp[4 * q + 3] - rd) * inv_sqrt_c);
            }
            __syncwarp();
        }

        // ---- (E) d/dx and d/dy for the lane's channel rows (one tensor gx = dx + dy when x is y)
        if (p.gx || p.gy) {
#pragma unroll 1
            for (int t = 0; t < T; ++t) {
                const int c = t * 32 + lane;
                if (c >= C) break;
                float dxo[L], dyo[L];
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    const float4 g4 = lds4(gs + c * L + 4 * q);
                    const float w1 = has_sum ? wsum : 0.f;
                    dxo[4 * q] = w1 * g4.x; dxo[4 * q + 1] = w1 * g4.y; dxo[4 * q + 2] = w1 * g4.z; dxo[4 * q + 3] = w1 * g4.w;
#pragma unroll
                    for (int e = 0; e < 4; ++e) dyo[4 * q + e] = dxo[4 * q + e];
                }
                if (k_attn >= 0) {
                    float xr[L], yr2[L], dor[L];
#pragma unroll
                    for (int q = 0; q < Q; ++q) {
                        const float4 v = lds4(xs + c * L + 4 * q), u = lds4(ys + c * L + 4 * q), d = lds4(os + c * L + 4 * q);
                        xr[4 * q] = v.x; xr[4 * q + 1] = v.y; xr[4 * q + 2] = v.z; xr[4 * q + 3] = v.w;
                        yr2[4 * q] = u.x; yr2[4 * q + 1] = u.y; yr2[4 * q + 2] = u.z; yr2[4 * q + 3] = u.w;
                        dor[4 * q] = d.x; dor[4 * q + 1] = d.y; dor[4 * q + 2] = d.z; dor[4 * q + 3] = d.w;
                    }
#pragma unroll
                    for (int r = 0; r < L; ++r) {
                        float pr[L], dsr[L];
#pragma unroll
                        for (int q = 0; q < Q; ++q) {
                            const float4 v = lds4(Ps + r * L + 4 * q), d = lds4(dSs + r * L + 4 * q);
                            pr[4 * q] = v.x; pr[4 * q + 1] = v.y; pr[4 * q + 2] = v.z; pr[4 * q + 3] = v.w;
                            dsr[4 * q] = d.x; dsr[4 * q + 1] = d.y; dsr[4 * q + 2] = d.z; dsr[4 * q + 3] = d.w;
                        }
                        float dxr = 0.f;
#pragma unroll
                        for (int j = 0; j < L; ++j) {
                            dxr = fmaf(dsr[j], yr2[j], dxr);                      // dx[c,r] = sum_j dS[r][j] y[c,j]
                            dyo[j] = fmaf(dor[r], pr[j], dyo[j]);                 // dy[c,j] += dO[c,r] P[r][j]
                            dyo[j] = fmaf(xr[r], dsr[j], dyo[j]);                 //          + x[c,r] dS[r][j]
                        }
                        dxo[r] += dxr;
                    }
                }
                if (alias) {
#pragma unroll
                    for (int j = 0; j < L; ++j) dxo[j] += dyo[j];
                }
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    if (p.gx) {
                        float* dst = p.gx + base + c * L + 4 * q;
                        float4 o4 = make_float4(dxo[4 * q], dxo[4 * q + 1], dxo[4 * q + 2], dxo[4 * q + 3]);
                        if (p.gx_accum) {
                            const float4 cur = *reinterpret_cast<const float4*>(dst);
                            o4.x += cur.x; o4.y += cur.y; o4.z += cur.z; o4.w += cur.w;
                        }
                        *reinterpret_cast<float4*>(dst) = o4;
                    }
                    if (!alias && p.gy) {
                        float* dst = p.gy + base + c * L + 4 * q;
                        float4 o4 = make_float4(dyo[4 * q], dyo[4 * q + 1], dyo[4 * q + 2], dyo[4 * q + 3]);
                        if (p.gy_accum) {
                            const float4 cur = *reinterpret_cast<const float4*>(dst);
                            o4.x += cur.x; o4.y += cur.y; o4.z += cur.z; o4.w += cur.w;
                        }
                        *reinterpret_cast<float4*>(dst) = o4;
                    }
                }
            }
        }
    }

    // ---- CTA sums -> global accumulator (one atomic per value), last CTA finalises (as k_node_bwd)
    dg_sum = warp_sum(dg_sum);
    dg_attn = warp_sum(dg_attn);
#pragma unroll
    for (int zi = 0; zi < WMAXZ; ++zi) dg_z[zi] = warp_sum(dg_z[zi]);
    if (lane == 0) {
        int zi = 0;
        for (int k = 0; k < p.n_ops; ++k) {
            const int ty = p.op_type[k];
            float v;
            if (ty == BMNAS_OP_SUM) v = dg_sum;
            else if (ty == BMNAS_OP_ATTN) v = dg_attn;
            else { v = zi == 0 ? dg_z[0] : (zi == 1 ? dg_z[1] : dg_z[2]); ++zi; }
            atomicAdd(s_dg + k, v);
        }
    }
    __syncthreads();
    // warp-private sums -> global accumulators: every warp parks its LayerNorm-affine registers in its (now idle)
    // slab, then all threads sum the WPCB slabs element-wise and issue ONE global atomic per value
    cp_async_wait_all();
    if (k_attn >= 0) {
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const int c = t * 32 + lane;
            if (c < C) {
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    *reinterpret_cast<float4*>(slab + c * L + 4 * q) = make_float4(aG[t][4 * q], aG[t][4 * q + 1], aG[t][4 * q + 2], aG[t][4 * q + 3]);
                    *reinterpret_cast<float4*>(slab + CLr + c * L + 4 * q) = make_float4(aH[t][4 * q], aH[t][4 * q + 1], aH[t][4 * q + 2], aH[t][4 * q + 3]);
                }
            }
        }
    }
    __syncthreads();
    float* gacc = p.partials;
    const size_t s_off = 5 * CLr + 2 * LLr;            // S1 | S2 inside a slab
    for (int i = threadIdx.x; i < M; i += WPCB * 32) {
        float a1 = 0.f, a2 = 0.f;
#pragma unroll
        for (int w = 0; w < WPCB; ++w) {
            a1 += slab0[w * slab_floats + s_off + i];
            a2 += slab0[w * slab_floats + s_off + Mr + i];
        }
        atomicAdd(gacc + i, a1);
        atomicAdd(gacc + M + i, a2);
    }
    if (threadIdx.x < p.n_ops) atomicAdd(gacc + 2 * M + threadIdx.x, s_dg[threadIdx.x]);
    if (k_attn >= 0 && p.g_ln_w[k_attn]) {
        for (int e = threadIdx.x; e < CL; e += WPCB * 32) {
            float a1 = 0.f, a2 = 0.f;
#pragma unroll
            for (int w = 0; w < WPCB; ++w) {
                a1 += slab0[w * slab_floats + e];
                a2 += slab0[w * slab_floats + CLr + e];
            }
            atomicAdd(p.g_ln_w[k_attn] + e, a1);
            atomicAdd(p.g_ln_b[k_attn] + e, a2);
        }
    }
    if (!last_block(p.counter, gridDim.x)) return;
    const int PW = 2 * M + BMNAS_MAX_OPS;
    float* tot = S1s;                                  // S1s | S2s are contiguous (2 * Mr); totals for gamma in s_dg
    for (int v = threadIdx.x; v < PW; v += WPCB * 32) {
        const float x = ld_cg(gacc + v);
        gacc[v] = 0.f;
        if (v < M) S1s[v] = x;
        else if (v < 2 * M) S2s[v - M] = x;
        else s_dg[v - 2 * M] = x;
    }
    (void)tot;
    __syncthreads();
    const float n = (float)p.B * (float)L;
    for (int k = 0; k < p.n_ops; ++k) {
        const int ty = p.op_type[k];
        if (ty == BMNAS_OP_SUM || ty == BMNAS_OP_ATTN) continue;
        const int rows = ty == BMNAS_OP_GLU ? 2 * C : C, zo = p.z_off[k];
        for (int ml = threadIdx.x; ml < rows; ml += WPCB * 32) {
            const int m = zo + ml;
            const float s1 = S1s[m], s2 = S2s[m];
            if (p.g_bn_w[k]) {
                p.g_bn_w[k][ml] = s2;
                p.g_bn_b[k][ml] = s1;
            }
            const float rs = bnr[m].x, mur = bnr[m].y, w = bnw[m].x;
            if (p.training) {
                const float a = w * rs, m1 = s1 / n, m2 = s2 / n;
                p.coef_a[m] = a;
                p.coef_b[m] = -a * rs * m2;
                p.coef_c[m] = a * (mur * m2 - m1);
            } else {
                p.coef_a[m] = w * rs;
                p.coef_b[m] = 0.f;
                p.coef_c[m] = 0.f;
            }
        }
    }
    if (p.g_gamma && threadIdx.x == 0) {
        float dot = 0.f;
#pragma unroll 1
        for (int k = 0; k < p.n_ops; ++k) dot += s_gw[k] * s_dg[k];
#pragma unroll 1
        for (int k = 0; k < p.n_ops; ++k)
            p.g_gamma[k] = p.gamma_is_logits ? s_gw[k] * (s_dg[k] - dot) : s_dg[k];
    }
}

// 0 = by batch size (default), 1 = always the CTA-per-sample kernels, 2 = the warp-per-sample kernels whenever eligible
int node_variant_flag = 0;

// add v (already summed over the thread's group) into acc[m]; the L/G lanes that share channel m are
// adjacent and aligned, so a segmented shuffle + one plain store per channel is race-free and deterministic
template <bool SEG>
__device__ __forceinline__ void chan_add(float* acc, int m, float v, int lanes, bool active) {
    if (SEG) {
        for (int o = lanes >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (active && ((threadIdx.x & (lanes - 1)) == 0)) acc[m] += v;
    } else {
        if (active) atomicAdd(acc + m, v);
    }
}

template <int G, bool SEG>
__global__ void __launch_bounds__(NTH_BWD_MAX) k_node_bwd(const bmnas_node_params p) {
    // early section (p.early_ok): x, y, Z, mean, rstd are forward tensors, complete long before any backward kernel;
    // the tile loads, the BatchNorm constants and the recomputation of the attention primitive for the CTA's first
    // sample run BEFORE pdl_wait() and overlap the kernel that produces gout
    bool waited = !p.early_ok;
    if (waited) pdl_prologue();
    extern __shared__ __align__(16) float smem[];
    const int C = p.C, L = p.L, CL = C * L, M = p.M;
    NodeSmem sm = node_carve(smem, C, L, M, true);
    if (p.alias_xy) sm.ys = sm.xs;
    __shared__ float s_cw[BMNAS_MAX_SRC + 1];
    node_setup_gamma(p, sm);
    node_setup_chain(p, s_cw);
    node_setup_bn(p, sm);
    int k_attn = -1;
#pragma unroll 1
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
    // per-thread dL/dgamma partials live in shared memory slots so the loop over primitives stays a real
    // loop (an unrolled 8-way switch made this kernel 240 KB of SASS and instruction-fetch bound)
#pragma unroll
    for (int k = 0; k < BMNAS_MAX_OPS; ++k) sm.dgs[k * NTH + threadIdx.x] = 0.f;
    const float inv_sqrt_c = 1.f / sqrtf((float)C);
    const int lanes = SEG ? L / G : 1;   // lanes sharing one channel
    const int NG = CL / G;

    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        __syncthreads();
        load_tile<G>(sm.xs, p.x + (long long)b * CL, CL);
        if (!p.alias_xy) load_tile<G>(sm.ys, p.y + (long long)b * CL, CL);
        __syncthreads();
        float a_mean = 0.f, a_rstd = 0.f;
        if (k_attn >= 0) attn_forward<G>(p, sm, k_attn, b, &a_mean, &a_rstd);
        if (!waited) {
            pdl_prologue();
            waited = true;
        }
        if (!p.gout2) {
            load_tile<G>(sm.gs, p.gout + (long long)b * CL, CL);
        } else {                             // upstream gradient = gout (if any) + cw_last * gout2 (chained edge mix)
            const float cl = s_cw[p.n_chain];
#pragma unroll 1
            for (int g = threadIdx.x; g < CL / G; g += NTH) {
                float a[G], h[G];
                ldg_v<G>(p.gout2 + (long long)b * CL + g * G, h);
                if (p.gout) {
                    ldg_v<G>(p.gout + (long long)b * CL + g * G, a);
                } else {
#pragma unroll
                    for (int q = 0; q < G; ++q) a[q] = 0.f;
                }
#pragma unroll
                for (int q = 0; q < G; ++q) a[q] = fmaf(cl, h[q], a[q]);
                st_v<G>(sm.gs + g * G, a);
            }
        }
        __syncthreads();
        const float* Zb = p.Z ? p.Z + (long long)b * M * L : nullptr;
        float* GVb = p.GV ? p.GV + (long long)b * M * L : nullptr;
        float lnsum[2] = {0.f, 0.f};  // sum q, sum q*ohat for the attention LayerNorm backward

        for (int g0 = 0; g0 < NG; g0 += NTH) {
            const int g = g0 + threadIdx.x;
            const bool act = g < NG;
            const int e0 = act ? g * G : 0, c = e0 / L;
            const long long li = (long long)b * CL + e0;
            const unsigned long long gi = (unsigned long long)(p.sample_offset + b) * CL + e0;
            float xv[G], yv[G], gv_[G], gxe[G], gye[G];
            lds_v<G>(sm.xs + e0, xv);
            lds_v<G>(sm.ys + e0, yv);
            lds_v<G>(sm.gs + e0, gv_);
#pragma unroll
            for (int q = 0; q < G; ++q) {
                if (!act) gv_[q] = 0.f;
                gxe[q] = 0.f;
                gye[q] = 0.f;
            }
#pragma unroll 1
            for (int k = 0; k < p.n_ops; ++k) {
                const int ty = p.op_type[k];
                const float wk = sm.gw[k];
                float dgk = 0.f;
                if (ty == BMNAS_OP_SUM) {
#pragma unroll
                    for (int q = 0; q < G; ++q) {
                        dgk += gv_[q] * (xv[q] + yv[q]);
                        gxe[q] += wk * gv_[q];
                        gye[q] += wk * gv_[q];
                    }
                } else if (ty == BMNAS_OP_ATTN) {
                    float a[G], Gw[G], Gb[G], lg[G], lh[G];
                    lds_v<G>(sm.as + e0, a);
                    ldg_v<G>(p.ln_w[k] + e0, Gw);
                    ldg_v<G>(p.ln_b[k] + e0, Gb);
                    lds_v<G>(sm.lnG + e0, lg);
                    lds_v<G>(sm.lnH + e0, lh);
#pragma unroll
                    for (int q = 0; q < G; ++q) {
                        const float oh = (a[q] - a_mean) * a_rstd;
                        dgk += gv_[q] * (oh * Gw[q] + Gb[q]);
                        const float go = wk * gv_[q];
                        lg[q] += go * oh;
                        lh[q] += go;
                        const float qq = go * Gw[q];
                        lnsum[0] += qq;
                        lnsum[1] += qq * oh;
                    }
                    if (act) {
                        st_v<G>(sm.lnG + e0, lg);
                        st_v<G>(sm.lnH + e0, lh);
                    }
                } else {
                    const int zo = p.z_off[k], m = zo + c;
                    const bool drop = p.training && p.p_drop[k] > 0.f;
                    float ds[G], z[G];
                    drop_v<G>(drop, p.mask[k], p.rng_state, p.op_uid[k], li, gi, p.p_drop[k], ds);
                    ldg_v<G>(Zb + (long long)zo * L + e0, z);
                    const float r = sm.rs[m], mr = sm.mr[m], w = sm.bw[m], bb = sm.bb[m];
                    if (ty == BMNAS_OP_GLU) {
                        float zg[G], gva[G], gvg[G];
                        ldg_v<G>(Zb + (long long)(zo + C) * L + e0, zg);
                        const float r2 = sm.rs[m + C], mr2 = sm.mr[m + C], w2 = sm.bw[m + C], bb2 = sm.bb[m + C];
                        float s1a = 0.f, s2a = 0.f, s1g = 0.f, s2g = 0.f;
#pragma unroll
                        for (int q = 0; q < G; ++q) {
                            const float zha = fmaf(z[q], r, -mr), zhg = fmaf(zg[q], r2, -mr2);
                            const float va = fmaf(zha, w, bb), vg = fmaf(zhg, w2, bb2);
                            const float s = sigmoidf_(vg);
                            dgk += gv_[q] * (va * s * ds[q]);
                            const float go = wk * gv_[q] * ds[q];
                            gva[q] = go * s;
                            gvg[q] = go * va * s * (1.f - s);
                            s1a += gva[q]; s2a += gva[q] * zha;
                            s1g += gvg[q]; s2g += gvg[q] * zhg;
                        }
                        if (act) {
                            st_v<G>(GVb + (long long)zo * L + e0, gva);
                            st_v<G>(GVb + (long long)(zo + C) * L + e0, gvg);
                        }
                        chan_add<SEG>(sm.S1s, m, s1a, lanes, act);
                        chan_add<SEG>(sm.S2s, m, s2a, lanes, act);
                        chan_add<SEG>(sm.S1s, m + C, s1g, lanes, act);
                        chan_add<SEG>(sm.S2s, m + C, s2g, lanes, act);
                    } else {
                        float gvv[G];
                        float s1 = 0.f, s2 = 0.f;
#pragma unroll
                        for (int q = 0; q < G; ++q) {
                            const float zha = fmaf(z[q], r, -mr);
                            const float va = fmaf(zha, w, bb);
                            float o, d;
                            if (ty == BMNAS_OP_FC_RELU) {
                                o = fmaxf(va, 0.f);
                                d = va > 0.f ? 1.f : 0.f;
                            } else {
                                o = mishf_(va);
                                d = mish_grad(va);
                            }
                            dgk += gv_[q] * (o * ds[q]);
                            gvv[q] = wk * gv_[q] * ds[q] * d;
                            s1 += gvv[q];
                            s2 += gvv[q] * zha;
                        }
                        if (act) st_v<G>(GVb + (long long)zo * L + e0, gvv);
                        chan_add<SEG>(sm.S1s, m, s1, lanes, act);
                        chan_add<SEG>(sm.S2s, m, s2, lanes, act);
                    }
                }
                sm.dgs[k * NTH + threadIdx.x] += dgk;
            }
            if (act) {
                st_v<G>(sm.dxs + e0, gxe);
                st_v<G>(sm.dys + e0, gye);
            }
        }

        if (k_attn >= 0) {
            const int k = k_attn;
            block_sum<2>(lnsum, sm.red);
            const float mq = lnsum[0] / (float)CL, mqo = lnsum[1] / (float)CL;
            const float wk = sm.gw[k];
            const bool drop = p.training && p.p_drop[k] > 0.f;
            for (int g = threadIdx.x; g < NG; g += NTH) {
                const int e0 = g * G;
                float a[G], Gw[G], gg[G], ds[G], dd[G];
                lds_v<G>(sm.as + e0, a);
                ldg_v<G>(p.ln_w[k] + e0, Gw);
                lds_v<G>(sm.gs + e0, gg);
                drop_v<G>(drop, p.mask[k], p.rng_state, p.op_uid[k], (long long)b * CL + e0,
                          (unsigned long long)(p.sample_offset + b) * CL + e0, p.p_drop[k], ds);
#pragma unroll
                for (int q = 0; q < G; ++q) {
                    const float oh = (a[q] - a_mean) * a_rstd;
                    dd[q] = a_rstd * (wk * gg[q] * Gw[q] - mq - oh * mqo) * ds[q];
                }
                st_v<G>(sm.as + e0, dd);  // dO[c,i]
            }
            __syncthreads();
            lxl_contract(sm.as, sm.ys, sm.Sp, sm.S2, C, L, 1.f);  // dP[i][j] = sum_c dO[c,i] y[c,j]
            for (int pr = threadIdx.x; pr < L * L; pr += NTH) {    // dS = P o (dP - rowdot) / sqrt(C)
                const int i = pr / L;
                float rd = 0.f;
                for (int j = 0; j < L; ++j) rd = fmaf(sm.S2[i * L + j], sm.S[i * L + j], rd);
                sm.Sp[pr] = sm.S[pr] * (sm.S2[pr] - rd) * inv_sqrt_c;
            }
            __syncthreads();
            for (int g = threadIdx.x; g < NG; g += NTH) {
                const int e0 = g * G, c = e0 / L, i0 = e0 - c * L;
                float dx[G], dy[G];
                lds_v<G>(sm.dxs + e0, dx);
                lds_v<G>(sm.dys + e0, dy);
                for (int j = 0; j < L; ++j) {
                    const float yj = sm.ys[c * L + j], dOj = sm.as[c * L + j], xj = sm.xs[c * L + j];
#pragma unroll
                    for (int q = 0; q < G; ++q) {
                        dx[q] = fmaf(sm.Sp[(i0 + q) * L + j], yj, dx[q]);        // dx[c,i] += dS[i][j] y[c,j]
                        dy[q] = fmaf(dOj, sm.S[j * L + i0 + q], dy[q]);          // dy[c,i] += dO[c,j] P[j][i]
                        dy[q] = fmaf(xj, sm.Sp[j * L + i0 + q], dy[q]);          //          + x[c,j] dS[j][i]
                    }
                }
                st_v<G>(sm.dxs + e0, dx);
                st_v<G>(sm.dys + e0, dy);
            }
        }
        __syncthreads();
        for (int g = threadIdx.x; g < NG; g += NTH) {
            const int e0 = g * G;
            const long long li = (long long)b * CL + e0;
            if (p.gout2) {                   // gradients of the chained mix's earlier states: cw_j * gout2
                float h[G];
                ldg_v<G>(p.gout2 + li, h);
#pragma unroll 1
                for (int j = 0; j < p.n_chain; ++j) {
                    if (!p.chain_gx[j]) continue;
                    float o[G];
                    const float cj = s_cw[j];
#pragma unroll
                    for (int q = 0; q < G; ++q) o[q] = cj * h[q];
                    if (p.chain_gx_accum[j]) {
                        float c_[G];
                        lds_v<G>(p.chain_gx[j] + li, c_);
#pragma unroll
                        for (int q = 0; q < G; ++q) o[q] += c_[q];
                    }
                    st_v<G>(p.chain_gx[j] + li, o);
                }
            }
            float dx[G], dy[G];
            lds_v<G>(sm.dxs + e0, dx);
            lds_v<G>(sm.dys + e0, dy);
            if (p.alias_xy) {
                if (p.gx) {
                    float o[G];
#pragma unroll
                    for (int q = 0; q < G; ++q) o[q] = dx[q] + dy[q];
                    if (p.gx_accum) {
                        float c_[G];
                        lds_v<G>(p.gx + li, c_);
#pragma unroll
                        for (int q = 0; q < G; ++q) o[q] += c_[q];
                    }
                    st_v<G>(p.gx + li, o);
                }
            } else {
                if (p.gx) {
                    if (p.gx_accum) {
                        float c_[G];
                        lds_v<G>(p.gx + li, c_);
#pragma unroll
                        for (int q = 0; q < G; ++q) dx[q] += c_[q];
                    }
                    st_v<G>(p.gx + li, dx);
                }
                if (p.gy) {
                    if (p.gy_accum) {
                        float c_[G];
                        lds_v<G>(p.gy + li, c_);
#pragma unroll
                        for (int q = 0; q < G; ++q) dy[q] += c_[q];
                    }
                    st_v<G>(p.gy + li, dy);
                }
            }
        }
    }

    if (!waited) pdl_prologue();
    // ---- per-CTA sums -> one global accumulator (red.add), then the last CTA finalises.
    //      (a fixed-order reduction of per-CTA partials by a single CTA costs ~100 dependent L2 round
    //      trips; the accumulator is self-cleaning like the counter)
    float dg[BMNAS_MAX_OPS];
#pragma unroll
    for (int k = 0; k < BMNAS_MAX_OPS; ++k) dg[k] = sm.dgs[k * NTH + threadIdx.x];
    block_sum<BMNAS_MAX_OPS>(dg, sm.red);
    __syncthreads();
    const int PW = 2 * M + BMNAS_MAX_OPS;
    float* gacc = p.partials;
    for (int i = threadIdx.x; i < M; i += NTH) {
        atomicAdd(gacc + i, sm.S1s[i]);
        atomicAdd(gacc + M + i, sm.S2s[i]);
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < BMNAS_MAX_OPS; ++k)
            if (k < p.n_ops) atomicAdd(gacc + 2 * M + k, dg[k]);
    }
    if (k_attn >= 0 && p.g_ln_w[k_attn]) {
        for (int e = threadIdx.x; e < CL; e += NTH) {
            atomicAdd(p.g_ln_w[k_attn] + e, sm.lnG[e]);
            atomicAdd(p.g_ln_b[k_attn] + e, sm.lnH[e]);
        }
    }
    if (!last_block(p.counter, gridDim.x)) return;
    for (int v = threadIdx.x; v < PW; v += NTH) {
        sm.tot[v] = ld_cg(gacc + v);
        gacc[v] = 0.f;
    }
    __syncthreads();
    const float n = (float)p.B * (float)L;
    for (int k = 0; k < p.n_ops; ++k) {
        const int ty = p.op_type[k];
        if (ty == BMNAS_OP_SUM || ty == BMNAS_OP_ATTN) continue;
        const int rows = ty == BMNAS_OP_GLU ? 2 * C : C, zo = p.z_off[k];
        for (int ml = threadIdx.x; ml < rows; ml += NTH) {
            const int m = zo + ml;
            const float s1 = sm.tot[m], s2 = sm.tot[M + m];
            if (p.g_bn_w[k]) {
                p.g_bn_w[k][ml] = s2;
                p.g_bn_b[k][ml] = s1;
            }
            const float rs = sm.rs[m], mur = sm.mr[m];   // mean * rstd
            if (p.training) {
                const float a = sm.bw[m] * rs, m1 = s1 / n, m2 = s2 / n;
                p.coef_a[m] = a;
                p.coef_b[m] = -a * rs * m2;
                p.coef_c[m] = a * (mur * m2 - m1);
            } else {  // eval-mode BN is a fixed affine map
                p.coef_a[m] = sm.bw[m] * rs;
                p.coef_b[m] = 0.f;
                p.coef_c[m] = 0.f;
            }
        }
    }
    if (p.g_gamma && threadIdx.x == 0) {
        float dot = 0.f;
#pragma unroll 1
        for (int k = 0; k < p.n_ops; ++k) dot += sm.gw[k] * sm.tot[2 * M + k];
#pragma unroll 1
        for (int k = 0; k < p.n_ops; ++k)
            p.g_gamma[k] = p.gamma_is_logits ? sm.gw[k] * (sm.tot[2 * M + k] - dot) : sm.tot[2 * M + k];
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
    if (bwd && ((!p->gout && !p->gout2) || !p->partials || !p->counter)) return BMNAS_EINVAL;
    if (!bwd && !p->out) return BMNAS_EINVAL;
    if (p->n_chain < 0 || p->n_chain > BMNAS_MAX_SRC) return BMNAS_EINVAL;
    if ((p->out2 || p->gout2) && !p->chain_w) return BMNAS_EINVAL;
    if (!bwd && p->out2)
        for (int j = 0; j < p->n_chain; ++j)
            if (!p->chain_x[j]) return BMNAS_EINVAL;
    return BMNAS_OK;
}

static bool al16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }

// 128-bit groups need L % 4 == 0 (a group never straddles a channel) and 16-byte aligned tensors
static bool node_vec_ok(const bmnas_node_params* p, bool bwd) {
    if (p->L % 4) return false;
    if (!al16(p->x) || !al16(p->y) || !al16(p->Z) || !al16(p->out) || !al16(p->gout) || !al16(p->gx) || !al16(p->gy) ||
        !al16(p->GV))
        return false;
    for (int k = 0; k < p->n_ops; ++k) {
        if (!al16(p->ln_w[k]) || !al16(p->ln_b[k]) || !al16(p->g_ln_w[k]) || !al16(p->g_ln_b[k])) return false;
        if (p->mask[k] && (reinterpret_cast<uintptr_t>(p->mask[k]) & 3u)) return false;
    }
    if (!al16(p->out2) || !al16(p->gout2)) return false;
    for (int j = 0; j < BMNAS_MAX_SRC; ++j)
        if (!al16(p->chain_x[j]) || !al16(p->chain_gx[j])) return false;
    (void)bwd;
    return true;
}

template <class Kern>
static int node_smem_attr(Kern kern, size_t smem, size_t* configured) {
    // the 48 KB default limit counts static + dynamic shared memory: opt in with a margin for the static part
    if (smem > 40 * 1024 && smem > *configured) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return BMNAS_ELAUNCH;
        *configured = smem;
    }
    return BMNAS_OK;
}

}  // namespace bmnas

using namespace bmnas;

extern "C" long long bmnas_node_partials_size(const bmnas_node_params* p) {
    return (long long)kNodeMaxBlocksBwd * (2LL * p->M + BMNAS_MAX_OPS);
}

// warp-per-sample kernels: shapes they take (see k_node_fwd_warp)
static bool node_warp_ok(const bmnas_node_params* p, bool bwd) {
    const int L = p->L, T = (p->C + 31) / 32;
    if (p->out2 || p->gout2) return false;          // the chained edge mix lives in the CTA-per-sample kernels
    if (!(L == 4 || L == 8 || L == 16) || T * L > 32 || !node_vec_ok(p, bwd)) return false;
    int nz = 0, nglu = 0;
    for (int k = 0; k < p->n_ops; ++k) {
        const int ty = p->op_type[k];
        if (ty == BMNAS_OP_GLU) ++nglu;
        if (ty != BMNAS_OP_SUM && ty != BMNAS_OP_ATTN) ++nz;
    }
    if (nz > WMAXZ || nglu > 1) return false;
    return node_warp_smem_floats(p->C, L, p->M, p->alias_xy != 0) * sizeof(float) <= 200 * 1024;
}

// the CTA-per-sample kernels win while the batch is smaller than the machine's warp slots (latency bound).
// Measured crossover at NTU shapes (profiles/r01_v7_node_variants.txt): forward 10.8 vs 15.9 us at B=512 and
// 19.9 vs 16.2 us at B=1024; backward 32.1 vs 32.8 us at B=512 and 61.1 vs 33.1 us at B=1024.
constexpr int kWarpFwdMinB = 768, kWarpBwdMinB = 640;
static bool node_use_warp(const bmnas_node_params* p, bool bwd) {
    if (node_variant_flag == 1) return false;
    if (!node_warp_ok(p, bwd)) return false;
    return node_variant_flag == 2 || p->B >= kWarpFwdMinB;
}

template <int L, int T>
static int launch_node_fwd_warp(const bmnas_node_params* p, size_t smem, cudaStream_t stream) {
    static size_t configured = 0;
    int e = node_smem_attr(k_node_fwd_warp<L, T>, smem, &configured);
    if (e) return e;
    const int want = (p->B + WPC - 1) / WPC, cap = kNumSMs * 2;
    launch_k(k_node_fwd_warp<L, T>, want < cap ? want : cap, WPC * 32, smem, stream, *p);
    BMNAS_LAUNCH_CHECK();
    return BMNAS_OK;
}

static int node_fwd_warp_dispatch(const bmnas_node_params* p, cudaStream_t stream) {
    const int L = p->L, T = (p->C + 31) / 32;
    const size_t smem = node_warp_smem_floats(p->C, L, p->M, p->alias_xy != 0) * sizeof(float);
    // the smallest instantiated T' >= T (channels beyond C are masked off inside the kernel)
#define BMNAS_WCASE(l, t) if (L == l && T <= t) return launch_node_fwd_warp<l, t>(p, smem, stream)
    BMNAS_WCASE(4, 1); BMNAS_WCASE(4, 2); BMNAS_WCASE(4, 4); BMNAS_WCASE(4, 8);
    BMNAS_WCASE(8, 1); BMNAS_WCASE(8, 2); BMNAS_WCASE(8, 4);
    BMNAS_WCASE(16, 1); BMNAS_WCASE(16, 2);
#undef BMNAS_WCASE
    return BMNAS_EINVAL;
}

template <int L, int T>
static int launch_node_bwd_warp(const bmnas_node_params* p, size_t smem, cudaStream_t stream) {
    static size_t configured = 0;
    int e = node_smem_attr(k_node_bwd_warp<L, T>, smem, &configured);
    if (e) return e;
    const int want = (p->B + WPCB - 1) / WPCB;
    launch_k(k_node_bwd_warp<L, T>, want < kNumSMs ? want : kNumSMs, WPCB * 32, smem, stream, *p);
    BMNAS_LAUNCH_CHECK();
    return BMNAS_OK;
}

static bool node_bwd_warp_ok(const bmnas_node_params* p) {
    if (!node_warp_ok(p, true)) return false;
    return node_bwarp_smem_floats(p->C, p->L, p->M) * sizeof(float) <= 220 * 1024;
}

static int node_bwd_warp_dispatch(const bmnas_node_params* p, cudaStream_t stream) {
    const int L = p->L, T = (p->C + 31) / 32;
    const size_t smem = node_bwarp_smem_floats(p->C, L, p->M) * sizeof(float);
#define BMNAS_WCASE(l, t) if (L == l && T <= t) return launch_node_bwd_warp<l, t>(p, smem, stream)
    BMNAS_WCASE(4, 1); BMNAS_WCASE(4, 2); BMNAS_WCASE(4, 4); BMNAS_WCASE(4, 8);
    BMNAS_WCASE(8, 1); BMNAS_WCASE(8, 2); BMNAS_WCASE(8, 4);
    BMNAS_WCASE(16, 1); BMNAS_WCASE(16, 2);
#undef BMNAS_WCASE
    return BMNAS_EINVAL;
}

extern "C" int bmnas_set_node_variant(int v) {
    if (v < 0 || v > 2) return BMNAS_EINVAL;
    node_variant_flag = v;
    return BMNAS_OK;
}
extern "C" int bmnas_get_node_variant(void) { return node_variant_flag; }

extern "C" int bmnas_node_fwd(const bmnas_node_params* p, void* stream) {
    int e = node_check(p, false);
    if (e) return e;
    const int nth = node_threads(p->C, p->L, false, p->B);
    const size_t smem = node_smem_floats(p->C, p->L, p->M, false, nth) * sizeof(float);
    if (smem > 227 * 1024) return BMNAS_EINVAL;
    BMNAS_DRY_RETURN();
    if (node_use_warp(p, false)) return node_fwd_warp_dispatch(p, (cudaStream_t)stream);
    const bool vec = node_vec_ok(p, false);
    static size_t configured[2] = {0, 0};
    e = vec ? node_smem_attr(k_node_fwd<4>, smem, &configured[1]) : node_smem_attr(k_node_fwd<1>, smem, &configured[0]);
    if (e) return e;
    const int blocks = p->B < kNodeMaxBlocksFwd ? p->B : kNodeMaxBlocksFwd;
    if (vec)
        launch_k(k_node_fwd<4>, blocks, nth, smem, (cudaStream_t)stream, *p);
    else
        launch_k(k_node_fwd<1>, blocks, nth, smem, (cudaStream_t)stream, *p);
    BMNAS_LAUNCH_CHECK();
    return BMNAS_OK;
}

extern "C" int bmnas_node_bwd(const bmnas_node_params* p, void* stream) {
    int e = node_check(p, true);
    if (e) return e;
    const int nth = node_threads(p->C, p->L, true, p->B);
    const size_t smem = node_smem_floats(p->C, p->L, p->M, true, nth) * sizeof(float);
    if (smem > 227 * 1024) return BMNAS_EINVAL;
    BMNAS_DRY_RETURN();
    if (node_variant_flag != 1 && node_bwd_warp_ok(p) && (node_variant_flag == 2 || p->B >= kWarpBwdMinB))
        return node_bwd_warp_dispatch(p, (cudaStream_t)stream);
    const bool vec = node_vec_ok(p, true);
    const int lanes = p->L / 4;
    const bool seg = vec && lanes >= 1 && lanes <= 32 && (lanes & (lanes - 1)) == 0;
    static size_t configured[3] = {0, 0, 0};
    const int blocks = p->B < kNodeMaxBlocksBwd ? p->B : kNodeMaxBlocksBwd;
    if (vec && seg) {
        if ((e = node_smem_attr(k_node_bwd<4, true>, smem, &configured[0]))) return e;
        launch_k(k_node_bwd<4, true>, blocks, nth, smem, (cudaStream_t)stream, *p);
    } else if (vec) {
        if ((e = node_smem_attr(k_node_bwd<4, false>, smem, &configured[1]))) return e;
        launch_k(k_node_bwd<4, false>, blocks, nth, smem, (cudaStream_t)stream, *p);
    } else {
        if ((e = node_smem_attr(k_node_bwd<1, false>, smem, &configured[2]))) return e;
        launch_k(k_node_bwd<1, false>, blocks, nth, smem, (cudaStream_t)stream, *p);
    }
    BMNAS_LAUNCH_CHECK();
    return BMNAS_OK;
}
