// Fused step-node mixed op for batches SMALLER than the machine -- bmnas_mixed_small_fwd (include/bmnas_b200.h).
//
//   out[b] = sum_k softmax(gamma)_k * op_k(t_b, t_b)      ops in {Sum, ScaleDotAttn, LinearGLU, ConcatFC | CatConvMish}
//
// NodeMixedOp.forward and everything under it (node_operations.py:19-20, 30-39, 49-56, 75-82, 92-108, 118-120) for the
// searchable cell (both inputs are the same tensor, node_search.py:55) in ONE launch, where the reference batch (NTU:
// 96 samples x 8 positions = 768 columns) used bmnas_conv_fwd + bmnas_node_fwd = two dependent kernels of 8.9 + 6.8 us
// (tools/timeline.py: 16 us from conv start to the next conv start).  The tcgen05 kernel of mixed_tc.cu is built for
// thousands of columns (64-column tiles, a 393 KB hi/lo weight stream per CTA: 28 us at this size); here the problem is
// 75 MFLOP and pure latency, so the tile is small and the machine is filled instead:
//
//   CTA = (32 columns = 32/L whole samples) x (32 channels c, i.e. the 96 pre-BatchNorm rows c, C+c, 2C+c: GLU value,
//         GLU gate, FC), 256 threads; grid = (B*L/32) x (C/32) = 96 CTAs at NTU, every CTA resident at once.
//   1  weight tile (3 blocks of the tile-major fp32 image bmnas_wprep writes) by TMA bulk copy BEFORE griddepcontrol.wait
//      (overlaps the preceding kernel), activation tile (all C channels of the CTA's samples) by TMA bulk copy after it;
//   2  FFMA GEMM from shared memory: a thread owns one channel x 4 consecutive positions x the 3 rows = 12 accumulators,
//      i.e. exactly the 4-element group the epilogue (and one Philox call per dropout site) works on;
//   3  Z (kept for the backward) and per-tile BatchNorm (mean, M2) partials leave the CTA;
//      attention runs meanwhile from the same activation tile: S = t^T t / sqrt(C) per sample, softmax, O = P t for the
//      CTA's own channels, dropout, and the LayerNorm (mean, M2) partial of (sample, channel group);
//   4  ONE grid barrier (atomic ticket + generation flag; the grid is co-resident by construction);
//   5  every CTA merges the tile partials of ITS 96 rows (Chan, fixed order: deterministic) -> mean / rstd (the CTAs of
//      column tile 0 also store them for the backward and update running_mean / running_var / num_batches_tracked) and
//      the channel-group partials of ITS samples' attention LayerNorm;
//   6  epilogue from registers: BatchNorm + GLU / ReLU / Mish + dropout, LayerNorm affine, softmax(gamma)-weighted sum,
//      `out` (and the next inner edge mix `out2`, node_search.py:52-55) -- the pre-BatchNorm activations never come back
//      from HBM.
#include "common.cuh"
#include "gemm_shared.cuh"

namespace bmnas {
namespace ms {

constexpr int TH = 256;
constexpr int GT = 128;            // threads of the GEMM / epilogue role (warps 0-3); warps 4-7 run the attention primitive meanwhile
constexpr int TNC = 32;            // columns per CTA
constexpr int TCH = 32;            // channels per CTA
constexpr int SPAD = 8;            // floats between two sample blocks of the activation tile (bank spread)
constexpr int MAXSPT = 8;          // samples per tile (L = 4)
constexpr int MAXPARTS = 160;      // column tiles (grid.x): the whole grid must be resident

struct Ws {                        // zeroed once by the caller
    unsigned int bar_count;
    unsigned int bar_gen;
    unsigned int timeline;         // test hook: != 0 -> CTA (0, 0) records %globaltimer stamps into tl[]
    unsigned int pad;
    unsigned long long tl[16];
    // float part[n_col_tiles][M][2]; float ln_part[B][C/32][2] follow
};
#define MS_TL(i)                                                        \
    do {                                                                \
        if (tl_on) {                                                    \
            unsigned long long t__;                                     \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__));     \
            ws->tl[i] = t__;                                            \
        }                                                               \
    } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
            : "=r"(done)
            : "r"(s32(bar)), "r"(phase)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst_smem)),
                 "l"(src), "r"(bytes), "r"(s32(bar))
                 : "memory");
}

// grid-wide barrier among the co-resident CTAs of this launch; called by ONE thread per CTA
__device__ __forceinline__ void grid_barrier(Ws* ws, unsigned int n_ctas) {
    unsigned int gen;
    asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(gen) : "l"(&ws->bar_gen) : "memory");
    unsigned int prev;
    asm volatile("atom.acq_rel.gpu.add.u32 %0, [%1], 1;" : "=r"(prev) : "l"(&ws->bar_count) : "memory");
    if (prev == n_ctas - 1) {
        asm volatile("st.relaxed.gpu.u32 [%0], %1;" ::"l"(&ws->bar_count), "r"(0u) : "memory");
        asm volatile("st.release.gpu.u32 [%0], %1;" ::"l"(&ws->bar_gen), "r"(gen + 1u) : "memory");
    } else {
        unsigned int g;
        do {
            asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(g) : "l"(&ws->bar_gen) : "memory");
        } while (g == gen);
    }
}

struct Ops {                         // host-resolved op list (canonical order Sum < Attn < GLU < FC)
    int k_sum, k_attn, k_glu, k_fc, fc_mish;
};

// dropout scales of one 4-element group: the stream of drop_v<4> (node_apply.cu), so the backward (bmnas_node_bwd) redraws it
__device__ __forceinline__ void drop4(bool active, const unsigned char* mask, const unsigned long long* rng, uint32_t uid,
                                      long long li, unsigned long long gi, float p, float (&ds)[4]) {
    if (!active) {
        ds[0] = ds[1] = ds[2] = ds[3] = 1.f;
        return;
    }
    const float keep = 1.f / (1.f - p);
    if (mask) {
        const uchar4 m = *reinterpret_cast<const uchar4*>(mask + li);
        ds[0] = m.x ? keep : 0.f; ds[1] = m.y ? keep : 0.f; ds[2] = m.z ? keep : 0.f; ds[3] = m.w ? keep : 0.f;
    } else {
        const unsigned long long seed = rng[0], step = rng[1];
        const uint2 key = make_uint2((uint32_t)seed ^ (uid * 0x9E3779B1u), (uint32_t)(seed >> 32) + uid);
        const uint4 r = philox4x32(make_uint4((uint32_t)(gi >> 2), (uint32_t)(gi >> 34), (uint32_t)step, (uint32_t)(step >> 32)), key);
        const float sc = 1.0f / 16777216.0f;
        ds[0] = ((float)(r.x >> 8) * sc >= p) ? keep : 0.f;
        ds[1] = ((float)(r.y >> 8) * sc >= p) ? keep : 0.f;
        ds[2] = ((float)(r.z >> 8) * sc >= p) ? keep : 0.f;
        ds[3] = ((float)(r.w >> 8) * sc >= p) ? keep : 0.f;
    }
}

template <int L>
__global__ void __launch_bounds__(TH) k_mixed_small(const bmnas_conv_params cv, const bmnas_node_params nd, Ws* ws, const Ops ops,
                                                    const int n_ct) {
    constexpr int SPT = TNC / L;                 // samples per column tile
    constexpr int LL = L * L;
    constexpr int PR = L + 4;                    // padded row of P (floats): conflict-free 128-bit reads of 4 rows x 8 lanes
    constexpr int PS = L * PR + 4;               // sample stride of P
    constexpr int OS = TNC + 4;                  // row stride of the attention output tile [channel][column]
    extern __shared__ __align__(128) float smem[];
    const int C = nd.C, K = C, M = 3 * C, CL = C * L;
    const int ncg = C / TCH;
    float* As = smem;                            // [3][K][32]
    float* Bs = As + 3 * K * TCH;                // [SPT][CL + SPAD]
    const int sstr = CL + SPAD;
    __shared__ uint64_t bar_w, bar_x;
    __shared__ float s_gw[BMNAS_MAX_OPS], s_cw[BMNAS_MAX_SRC + 1];
    __shared__ __align__(16) float s_P[SPT * PS];           // softmax(QK^T / sqrt C) rows of the tile's samples
    __shared__ __align__(16) float s_O[TCH * OS];           // dropped attention output of the CTA's (channel, column) block
    __shared__ float s_red[4 * MAXSPT];
    __shared__ float s_lnm[SPT], s_lnr[SPT];     // attention LayerNorm mean / rstd per sample
    __shared__ float s_rs[3 * TCH], s_mr[3 * TCH];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool gemm_role = tid < GT;             // warps 0-3: GEMM + epilogue; warps 4-7: attention
    const int ct = blockIdx.x, cg = blockIdx.y;
    const int n0 = ct * TNC;                     // first column of the tile
    const int b0 = n0 / L;
    const int nsv = min(SPT, cv.B - b0);         // valid samples of this tile
    const bool has_attn = ops.k_attn >= 0;
    const bool tl_on = ws->timeline != 0 && blockIdx.x == 0 && blockIdx.y == 0 && (tid == 0 || tid == GT);
    if (tid == 0) MS_TL(0);
    float* part = reinterpret_cast<float*>(ws + 1);               // [n_ct][M][2]
    float* ln_part = part + (size_t)n_ct * M * 2;                 // [B][ncg][2]
    // (channel pair | channel, column quad) decomposition shared by both roles: 8 adjacent lanes = the 8 column quads
    const int rt = tid & (GT - 1);
    const int cq = rt & 7;
    const int s_my = (cq * 4) / L, l0 = (cq * 4) % L;
    const int b = b0 + s_my;
    const bool ok = s_my < nsv;

    // ---- early section: nothing here was written by the preceding kernel
    if (tid == 0) {
        mbar_init(&bar_w, 1);
        mbar_init(&bar_x, 1);
    }
    __syncthreads();
    const float* img = cv.wimg_fwd;
    const bool early = cv.early_ok != 0;
    auto load_w = [&]() {
        if (tid == 0) {
            mbar_expect_tx(&bar_w, 3u * (uint32_t)K * TCH * 4u);
#pragma unroll
            for (int sg = 0; sg < 3; ++sg)
                bulk_g2s(As + sg * K * TCH, img + (size_t)(sg * ncg + cg) * K * TCH, (uint32_t)K * TCH * 4u, &bar_w);
        }
    };
    if (early) load_w();
    if (tid == GT) {
        if (!nd.gamma) {
            for (int k = 0; k < nd.n_ops; ++k) s_gw[k] = 1.f;
        } else if (nd.gamma_is_logits) {
            float mxv = -INFINITY, s = 0.f;
            for (int k = 0; k < nd.n_ops; ++k) mxv = fmaxf(mxv, nd.gamma[k]);
            for (int k = 0; k < nd.n_ops; ++k) {
                s_gw[k] = expf(nd.gamma[k] - mxv);
                s += s_gw[k];
            }
            for (int k = 0; k < nd.n_ops; ++k) s_gw[k] /= s;
        } else {
            for (int k = 0; k < nd.n_ops; ++k) s_gw[k] = nd.gamma[k];
        }
        if (nd.chain_w) {
            for (int j = 0; j <= nd.n_chain; ++j) {
                const float a = nd.chain_w[2 * j], bq = nd.chain_w[2 * j + 1];
                s_cw[j] = nd.chain_is_logits ? 1.f / (1.f + expf(a - bq)) : bq;
            }
        }
    }
    // GEMM role: conv bias and BatchNorm affine of this thread's 2 channels x 3 rows (parameters: written long ago)
    const int cl2 = rt >> 3;                     // channel pair within the group (GEMM role): channels 2*cl2, 2*cl2 + 1
    float bias[2][3], bnw[2][3], bnb[2][3];
    if (gemm_role) {
        const float* bg = cv.bias[0];
        const float* bf = cv.bias[1];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = cg * TCH + 2 * cl2 + h;
            bias[h][0] = bg ? __ldg(bg + c) : 0.f;
            bias[h][1] = bg ? __ldg(bg + C + c) : 0.f;
            bias[h][2] = bf ? __ldg(bf + c) : 0.f;
            bnw[h][0] = __ldg(nd.bn_w[ops.k_glu] + c); bnb[h][0] = __ldg(nd.bn_b[ops.k_glu] + c);
            bnw[h][1] = __ldg(nd.bn_w[ops.k_glu] + C + c); bnb[h][1] = __ldg(nd.bn_b[ops.k_glu] + C + c);
            bnw[h][2] = __ldg(nd.bn_w[ops.k_fc] + c); bnb[h][2] = __ldg(nd.bn_b[ops.k_fc] + c);
        }
    }
    pdl_wait();
    pdl_trigger();
    if (tid == 0) MS_TL(1);
    if (!early) load_w();
    // ---- activation tile: all C channels of the tile's samples, one contiguous block per sample
    if (tid == 0) {
        mbar_expect_tx(&bar_x, (uint32_t)nsv * (uint32_t)CL * 4u);
        for (int s = 0; s < nsv; ++s) bulk_g2s(Bs + (size_t)s * sstr, nd.x + (long long)(b0 + s) * CL, (uint32_t)CL * 4u, &bar_x);
    }
    for (int s = nsv; s < SPT; ++s)
        for (int u = tid; u < CL / 4; u += TH) *reinterpret_cast<float4*>(Bs + (size_t)s * sstr + u * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();                             // the zero fill of missing samples
    mbar_wait(&bar_x, 0);
    if (tid == 0) MS_TL(2);

    float acc[2][3][4];
    float pre_dg[2][4], pre_df[2][4];            // dropout scales of the GLU / FC sites (GEMM role)
    float4 pre_lw[2], pre_lb[2], pre_sum[2], pre_c2[2];
    if (gemm_role) {
        // =========================================================== GEMM: rows (c, C + c, 2C + c) of 2 channels x 4 columns
        mbar_wait(&bar_w, 0);
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[h][i][j] = 0.f;
        const float* ap = As + 2 * cl2;
        const float* bp = Bs + (size_t)s_my * sstr + l0;
        constexpr int KB = 4;
        float2 a0[KB][3], a1[KB][3];
        float4 b0v[KB], b1v[KB];
        auto load_blk = [&](float2 (&a)[KB][3], float4 (&bv)[KB], int kk) {
#pragma unroll
            for (int j = 0; j < KB; ++j) {
                a[j][0] = *reinterpret_cast<const float2*>(ap + (kk + j) * TCH);
                a[j][1] = *reinterpret_cast<const float2*>(ap + (K + kk + j) * TCH);
                a[j][2] = *reinterpret_cast<const float2*>(ap + (2 * K + kk + j) * TCH);
                bv[j] = *reinterpret_cast<const float4*>(bp + (kk + j) * L);
            }
        };
        auto fma_blk = [&](const float2 (&a)[KB][3], const float4 (&bv)[KB]) {
#pragma unroll
            for (int j = 0; j < KB; ++j)
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    acc[0][i][0] = fmaf(a[j][i].x, bv[j].x, acc[0][i][0]);
                    acc[0][i][1] = fmaf(a[j][i].x, bv[j].y, acc[0][i][1]);
                    acc[0][i][2] = fmaf(a[j][i].x, bv[j].z, acc[0][i][2]);
                    acc[0][i][3] = fmaf(a[j][i].x, bv[j].w, acc[0][i][3]);
                    acc[1][i][0] = fmaf(a[j][i].y, bv[j].x, acc[1][i][0]);
                    acc[1][i][1] = fmaf(a[j][i].y, bv[j].y, acc[1][i][1]);
                    acc[1][i][2] = fmaf(a[j][i].y, bv[j].z, acc[1][i][2]);
                    acc[1][i][3] = fmaf(a[j][i].y, bv[j].w, acc[1][i][3]);
                }
        };
        const int nblk = K / KB;
        load_blk(a0, b0v, 0);
        int blk = 0;
        for (; blk + 2 <= nblk; blk += 2) {
            load_blk(a1, b1v, (blk + 1) * KB);
            fma_blk(a0, b0v);
            if (blk + 2 < nblk) load_blk(a0, b0v, (blk + 2) * KB);
            fma_blk(a1, b1v);
        }
        if (blk < nblk) fma_blk(a0, b0v);
        if (tid == 0) MS_TL(3);
        // ---- Z = acc + bias: kept for the backward; per-tile BatchNorm partials (the 8 lanes cq of a channel pair are adjacent)
        const int cnt = nsv * L;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = cg * TCH + 2 * cl2 + h;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[h][i][j] += bias[h][i];
                if (cv.bn_mode == 1) {
                    float s = ok ? (acc[h][i][0] + acc[h][i][1]) + (acc[h][i][2] + acc[h][i][3]) : 0.f;
#pragma unroll
                    for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                    const float mean = s / (float)cnt;
                    const float d0 = acc[h][i][0] - mean, d1 = acc[h][i][1] - mean, d2 = acc[h][i][2] - mean, d3 = acc[h][i][3] - mean;
                    float q = ok ? (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3) : 0.f;
#pragma unroll
                    for (int o = 4; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
                    if (cq == 0) *reinterpret_cast<float2*>(part + ((size_t)ct * M + i * C + c) * 2) = make_float2(mean, q);
                }
            }
        }
        // ---- everything of the epilogue that does not depend on the statistics happens BEFORE the grid barrier, in the
        //      time this CTA would otherwise wait for the slowest one: the four Philox draws, the LayerNorm affine, the
        //      Sum primitive and the prior states of the chained edge mix
        if (ok) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = cg * TCH + 2 * cl2 + h;
                const long long li = (long long)b * CL + (long long)c * L + l0;
                const unsigned long long gi = (unsigned long long)(nd.sample_offset + b) * CL + (unsigned long long)c * L + l0;
                drop4(nd.training && nd.p_drop[ops.k_glu] > 0.f, nd.mask[ops.k_glu], nd.rng_state, nd.op_uid[ops.k_glu], li, gi,
                      nd.p_drop[ops.k_glu], pre_dg[h]);
                drop4(nd.training && nd.p_drop[ops.k_fc] > 0.f, nd.mask[ops.k_fc], nd.rng_state, nd.op_uid[ops.k_fc], li, gi,
                      nd.p_drop[ops.k_fc], pre_df[h]);
                if (has_attn) {
                    pre_lw[h] = __ldg(reinterpret_cast<const float4*>(nd.ln_w[ops.k_attn] + (long long)c * L + l0));
                    pre_lb[h] = __ldg(reinterpret_cast<const float4*>(nd.ln_b[ops.k_attn] + (long long)c * L + l0));
                }
                float4 base = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ops.k_sum >= 0) {
                    const float4 xv = *reinterpret_cast<const float4*>(Bs + (size_t)s_my * sstr + (size_t)c * L + l0);
                    const float w = s_gw[ops.k_sum];
                    base = make_float4(w * (xv.x + xv.x), w * (xv.y + xv.y), w * (xv.z + xv.z), w * (xv.w + xv.w));
                }
                pre_sum[h] = base;
                float4 c2 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (nd.out2) {
#pragma unroll 1
                    for (int j = 0; j < nd.n_chain; ++j) {
                        const float4 v = __ldg(reinterpret_cast<const float4*>(nd.chain_x[j] + li));
                        const float cj = s_cw[j];
                        c2.x = fmaf(cj, v.x, c2.x); c2.y = fmaf(cj, v.y, c2.y); c2.z = fmaf(cj, v.z, c2.z); c2.w = fmaf(cj, v.w, c2.w);
                    }
                }
                pre_c2[h] = c2;
            }
        }
    } else if (has_attn) {
        // =========================================================== attention, everything that does not need the batch
        // statistics (ScaledDotAttn.forward node_operations.py:92-108), on the other four warps while the GEMM runs
        const int at = tid - GT;
        const float inv_sqrt_c = 1.f / sqrtf((float)C);
        constexpr int NIT = (SPT * LL + GT - 1) / GT;
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            // S[s][i][j] = q_i . k_j / sqrt(C): one entry per thread; the L entries of a row sit in L adjacent lanes, so the
            // row softmax is a segmented shuffle
            const int w = at + it * GT;
            const bool valid = w < SPT * LL;
            const int s = valid ? w / LL : 0, pr = valid ? w - s * LL : 0, i = pr / L, j = pr - i * L;
            const float* xs = Bs + (size_t)s * sstr;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            for (int ch = 0; ch < C; ch += 4) {
                a0 = fmaf(xs[ch * L + i], xs[ch * L + j], a0);
                a1 = fmaf(xs[(ch + 1) * L + i], xs[(ch + 1) * L + j], a1);
                a2 = fmaf(xs[(ch + 2) * L + i], xs[(ch + 2) * L + j], a2);
                a3 = fmaf(xs[(ch + 3) * L + i], xs[(ch + 3) * L + j], a3);
            }
            const float sv = ((a0 + a1) + (a2 + a3)) * inv_sqrt_c;
            float mx = sv;
#pragma unroll
            for (int o = L / 2; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            const float e = expf(sv - mx);
            float sum = e;
#pragma unroll
            for (int o = L / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if (valid) s_P[s * PS + i * PR + j] = e / sum;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(GT) : "memory");
        if (tid == GT) MS_TL(5);
        // O[c][l0 + q] = sum_j P[l0 + q][j] t[c][j] for the CTA's 32 channels: two (channel, quad) groups per thread; dropout
        const int k = ops.k_attn;
        const bool drop = nd.training && nd.p_drop[k] > 0.f && ok;
        float o_at[2][4];
        float tsum = 0.f;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int cl = (at >> 3) + 16 * h, c = cg * TCH + cl;
            const float* xr = Bs + (size_t)s_my * sstr + (size_t)c * L;
            const float* Pm = s_P + s_my * PS + l0 * PR;
            float xv[L];
#pragma unroll
            for (int j4 = 0; j4 < L / 4; ++j4) {
                const float4 t = *reinterpret_cast<const float4*>(xr + j4 * 4);
                xv[j4 * 4] = t.x; xv[j4 * 4 + 1] = t.y; xv[j4 * 4 + 2] = t.z; xv[j4 * 4 + 3] = t.w;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float a = 0.f;
#pragma unroll
                for (int j4 = 0; j4 < L / 4; ++j4) {
                    const float4 pq = *reinterpret_cast<const float4*>(Pm + q * PR + j4 * 4);
                    a = fmaf(pq.x, xv[j4 * 4], a); a = fmaf(pq.y, xv[j4 * 4 + 1], a);
                    a = fmaf(pq.z, xv[j4 * 4 + 2], a); a = fmaf(pq.w, xv[j4 * 4 + 3], a);
                }
                o_at[h][q] = a;
            }
            float ds[4];
            drop4(drop, nd.mask[k], nd.rng_state, nd.op_uid[k], (long long)b * CL + (long long)c * L + l0,
                  (unsigned long long)(nd.sample_offset + b) * CL + (unsigned long long)c * L + l0, nd.p_drop[k], ds);
#pragma unroll
            for (int q = 0; q < 4; ++q) o_at[h][q] *= ds[q];
            *reinterpret_cast<float4*>(s_O + cl * OS + cq * 4) = make_float4(o_at[h][0], o_at[h][1], o_at[h][2], o_at[h][3]);
            tsum += (o_at[h][0] + o_at[h][1]) + (o_at[h][2] + o_at[h][3]);
        }
        if (tid == GT) MS_TL(6);
        // LayerNorm (mean, M2) of this (sample, channel group): two passes over the CTA's 32 x L elements per sample.
        // Lanes of a warp that share a sample: the L/4 column quads of the sample (adjacent lanes) x the warp's 4 channel rows
        auto sample_sum = [&](float v) {
            if (L >= 8) v += __shfl_xor_sync(0xffffffffu, v, 1);
            if (L == 16) v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            return v;
        };
        const bool writer = (lane & 24) == 0 && (cq % (L / 4)) == 0;      // one lane per (warp, sample)
        const int wa = warp - GT / 32;
        {
            const float v = sample_sum(tsum);
            if (writer) s_red[wa * MAXSPT + s_my] = v;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(GT) : "memory");
        if (at < SPT) s_lnm[at] = ((s_red[at] + s_red[MAXSPT + at]) + (s_red[2 * MAXSPT + at] + s_red[3 * MAXSPT + at])) / (float)(TCH * L);
        asm volatile("bar.sync 1, %0;" ::"n"(GT) : "memory");
        {
            const float m = s_lnm[s_my];
            float q = 0.f;
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float d = o_at[h][u] - m;
                    q = fmaf(d, d, q);
                }
            const float v = sample_sum(q);
            if (writer) s_red[wa * MAXSPT + s_my] = v;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(GT) : "memory");
        if (at < nsv) {
            const float q = (s_red[at] + s_red[MAXSPT + at]) + (s_red[2 * MAXSPT + at] + s_red[3 * MAXSPT + at]);
            *reinterpret_cast<float2*>(ln_part + ((size_t)(b0 + at) * ncg + cg) * 2) = make_float2(s_lnm[at], q);
        }
    }

    // the running statistics of this CTA's rows (only this launch's column-tile-0 CTAs write them): fetched before the barrier
    float rm_old = 0.f, rv_old = 0.f;
    if (cv.bn_mode == 1 && ct == 0 && tid < 3 * TCH) {
        const int i = tid / TCH, r = tid - i * TCH;
        const int seg = i < 2 ? 0 : 1, ml = i < 2 ? i * C + cg * TCH + r : cg * TCH + r;
        if (cv.running_mean[seg]) {
            rm_old = __ldcg(cv.running_mean[seg] + ml);
            rv_old = __ldcg(cv.running_var[seg] + ml);
        }
    }
    // ---- the one grid-wide dependency: batch statistics of every row, LayerNorm statistics of every sample
    if (tid == 0) MS_TL(4);
    if (tid == GT) MS_TL(7);
    __threadfence();
    __syncthreads();
    if (tid == 0) grid_barrier(ws, gridDim.x * gridDim.y);
    if (tid == 0) MS_TL(8);
    // Z (kept for the backward; nobody reads it inside this launch) leaves the CTA while its thread 0 waits at the grid
    // barrier instead of in front of the fence above, where the stores' visibility was on the critical path
    if (gemm_role && cv.Z && ok) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = cg * TCH + 2 * cl2 + h;
#pragma unroll
            for (int i = 0; i < 3; ++i)
                *reinterpret_cast<float4*>(cv.Z + ((long long)b * M + i * C + c) * L + l0) =
                    make_float4(acc[h][i][0], acc[h][i][1], acc[h][i][2], acc[h][i][3]);
        }
    }
    __syncthreads();

    // ---- mean / rstd of this CTA's 96 rows (threads 0..95) and of its samples' attention LayerNorm (threads GT..)
    if (tid < 3 * TCH) {
        const int i = tid / TCH, r = tid - i * TCH;
        const int m = i * C + cg * TCH + r;
        const int seg = i < 2 ? 0 : 1, ml = i < 2 ? i * C + cg * TCH + r : cg * TCH + r;
        float mean, rstd;
        if (cv.bn_mode == 1) {
            // Chan's merge of the tile partials in two fixed-order passes (no per-merge division):
            //   mean = sum_t n_t mean_t / N,   M2 = sum_t [M2_t + n_t (mean_t - mean)^2]
            // up to 32 tiles (the reference batch: 24) every partial is fetched ONCE, all loads in flight together
            const int N = cv.B * L;
            constexpr int FB = 32;
            float m2 = 0.f;
            if (n_ct <= FB) {
                float2 v[FB];
#pragma unroll
                for (int j = 0; j < FB; ++j)
                    v[j] = j < n_ct ? __ldcg(reinterpret_cast<const float2*>(part + ((size_t)j * M + m) * 2)) : make_float2(0.f, 0.f);
                float sm = 0.f;
#pragma unroll
                for (int j = 0; j < FB; ++j)
                    if (j < n_ct) sm = fmaf((float)min(TNC, N - j * TNC), v[j].x, sm);
                mean = sm / (float)N;
#pragma unroll
                for (int j = 0; j < FB; ++j)
                    if (j < n_ct) {
                        const float d = v[j].x - mean;
                        m2 += fmaf((float)min(TNC, N - j * TNC) * d, d, v[j].y);
                    }
            } else {
                float sm = 0.f;
                for (int t0 = 0; t0 < n_ct; t0 += FB) {
                    float v[FB];
#pragma unroll
                    for (int j = 0; j < FB; ++j) v[j] = t0 + j < n_ct ? __ldcg(part + ((size_t)(t0 + j) * M + m) * 2) : 0.f;
#pragma unroll
                    for (int j = 0; j < FB; ++j)
                        if (t0 + j < n_ct) sm = fmaf((float)min(TNC, N - (t0 + j) * TNC), v[j], sm);
                }
                mean = sm / (float)N;
                for (int t0 = 0; t0 < n_ct; t0 += FB) {
                    float2 v[FB];
#pragma unroll
                    for (int j = 0; j < FB; ++j)
                        v[j] = t0 + j < n_ct ? __ldcg(reinterpret_cast<const float2*>(part + ((size_t)(t0 + j) * M + m) * 2)) : make_float2(0.f, 0.f);
#pragma unroll
                    for (int j = 0; j < FB; ++j)
                        if (t0 + j < n_ct) {
                            const float d = v[j].x - mean;
                            m2 += fmaf((float)min(TNC, N - (t0 + j) * TNC) * d, d, v[j].y);
                        }
                }
            }
            rstd = 1.f / sqrtf(m2 / (float)N + cv.eps);
            if (ct == 0) {
                cv.mean[m] = mean;
                cv.rstd[m] = rstd;
                if (cv.running_mean[seg]) {
                    const float unb = m2 / (float)max(N - 1, 1);
                    cv.running_mean[seg][ml] = (1.f - cv.momentum) * rm_old + cv.momentum * mean;
                    cv.running_var[seg][ml] = (1.f - cv.momentum) * rv_old + cv.momentum * unb;
                    if (ml == 0 && cv.num_batches_tracked[seg]) *cv.num_batches_tracked[seg] += 1;
                }
            }
        } else {                                                  // eval: running statistics
            mean = cv.running_mean[seg][ml];
            rstd = 1.f / sqrtf(cv.running_var[seg][ml] + cv.eps);
            if (ct == 0) {
                cv.mean[m] = mean;
                cv.rstd[m] = rstd;
            }
        }
        s_rs[tid] = rstd;
        s_mr[tid] = mean * rstd;
    } else if (has_attn && tid >= GT && tid < GT + nsv) {
        const int s = tid - GT;
        Wf w = {0.f, 0.f, 0.f};
        for (int g = 0; g < ncg; ++g) {
            const float2 v = __ldcg(reinterpret_cast<const float2*>(ln_part + ((size_t)(b0 + s) * ncg + g) * 2));
            const Wf bq = {(float)(TCH * L), v.x, v.y};
            w = wf_merge(w, bq);
        }
        s_lnm[s] = w.mean;
        s_lnr[s] = 1.f / sqrtf(w.m2 / (float)CL + kLnEps);
    }
    __syncthreads();

    if (tid == 0) MS_TL(9);
    // ---- epilogue (GEMM role): the mixed op for (channels 2*cl2, 2*cl2 + 1; sample b; positions l0 .. l0 + 3)
    if (!gemm_role || !ok) return;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int cl = 2 * cl2 + h, c = cg * TCH + cl;
        const long long li = (long long)b * CL + (long long)c * L + l0;
        // same accumulation order as k_node_fwd: Sum, ScaleDotAttn, LinearGLU, ConcatFC
        float out[4] = {pre_sum[h].x, pre_sum[h].y, pre_sum[h].z, pre_sum[h].w};
        if (has_attn) {
            const float w = s_gw[ops.k_attn], am = s_lnm[s_my], ar = s_lnr[s_my];
            const float4 o = *reinterpret_cast<const float4*>(s_O + cl * OS + cq * 4);
            out[0] = fmaf(w, (o.x - am) * ar * pre_lw[h].x + pre_lb[h].x, out[0]);
            out[1] = fmaf(w, (o.y - am) * ar * pre_lw[h].y + pre_lb[h].y, out[1]);
            out[2] = fmaf(w, (o.z - am) * ar * pre_lw[h].z + pre_lb[h].z, out[2]);
            out[3] = fmaf(w, (o.w - am) * ar * pre_lw[h].w + pre_lb[h].w, out[3]);
        }
        {
            const float r0 = s_rs[cl], m0 = s_mr[cl], r1 = s_rs[TCH + cl], m1 = s_mr[TCH + cl];
            const float w = s_gw[ops.k_glu];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float va = fmaf(fmaf(acc[h][0][q], r0, -m0), bnw[h][0], bnb[h][0]);
                const float vg = fmaf(fmaf(acc[h][1][q], r1, -m1), bnw[h][1], bnb[h][1]);
                out[q] = fmaf(w, va * sigmoidf_(vg) * pre_dg[h][q], out[q]);
            }
        }
        {
            const float r2 = s_rs[2 * TCH + cl], m2 = s_mr[2 * TCH + cl];
            const float w = s_gw[ops.k_fc];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float va = fmaf(fmaf(acc[h][2][q], r2, -m2), bnw[h][2], bnb[h][2]);
                out[q] = fmaf(w, (ops.fc_mish ? mishf_(va) : fmaxf(va, 0.f)) * pre_df[h][q], out[q]);
            }
        }
        *reinterpret_cast<float4*>(nd.out + li) = make_float4(out[0], out[1], out[2], out[3]);
        if (nd.out2) {                               // the next inner edge mix, from the same registers
            const float cw_l = s_cw[nd.n_chain];
            *reinterpret_cast<float4*>(nd.out2 + li) = make_float4(fmaf(cw_l, out[0], pre_c2[h].x), fmaf(cw_l, out[1], pre_c2[h].y),
                                                                   fmaf(cw_l, out[2], pre_c2[h].z), fmaf(cw_l, out[3], pre_c2[h].w));
        }
    }
    if (tid == 0) MS_TL(10);
}

static bool al16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }

static bool resolve_ops(const bmnas_conv_params* cv, const bmnas_node_params* nd, Ops* o) {
    o->k_sum = o->k_attn = o->k_glu = o->k_fc = -1;
    o->fc_mish = 0;
    int last = -1;
    for (int k = 0; k < nd->n_ops; ++k) {
        const int ty = nd->op_type[k];
        const int rank = ty == BMNAS_OP_SUM ? 0 : ty == BMNAS_OP_ATTN ? 1 : ty == BMNAS_OP_GLU ? 2 : 3;
        if (rank <= last) return false;                 // canonical order, every kind at most once
        last = rank;
        if (ty == BMNAS_OP_SUM) o->k_sum = k;
        else if (ty == BMNAS_OP_ATTN) o->k_attn = k;
        else if (ty == BMNAS_OP_GLU) o->k_glu = k;
        else if (ty == BMNAS_OP_FC_RELU || ty == BMNAS_OP_FC_MISH) {
            o->k_fc = k;
            o->fc_mish = ty == BMNAS_OP_FC_MISH;
        } else return false;
    }
    if (o->k_glu < 0 || o->k_fc < 0) return false;
    const int C = nd->C;
    if (nd->z_off[o->k_glu] != 0 || nd->z_off[o->k_fc] != 2 * C) return false;
    if (cv->n_seg != 2 || cv->seg_M[0] != 2 * C || cv->seg_M[1] != C) return false;
    return true;
}

static size_t smem_bytes(int C, int L) {
    return ((size_t)3 * C * TCH + (size_t)(TNC / L) * ((size_t)C * L + SPAD)) * sizeof(float);
}

}  // namespace ms
}  // namespace bmnas

using namespace bmnas;

extern "C" int bmnas_mixed_small_supported(const bmnas_conv_params* cv, const bmnas_node_params* nd) {
    using namespace ms;
    if (!cv || !nd) return 0;
    const int C = nd->C, L = nd->L;
    if (C < TCH || (C % TCH) || C > 256 || !(L == 4 || L == 8 || L == 16)) return 0;
    if (cv->K != C || cv->M != 3 * C || cv->n_src != 1 || cv->src_C[0] != C || cv->L != L || cv->B != nd->B || cv->B < 1) return 0;
    if (!nd->alias_xy || cv->src[0] != nd->x || nd->M != cv->M) return 0;
    if (cv->bn_mode != 1 && cv->bn_mode != 2) return 0;
    if (cv->bn_mode == 2 && (!cv->running_mean[0] || !cv->running_mean[1] || !cv->running_var[0] || !cv->running_var[1])) return 0;
    if (!cv->wimg_fwd || cv->wimg_fmt != 1) return 0;
    const long long n_ct = ((long long)cv->B * L + TNC - 1) / TNC;
    if (n_ct > MAXPARTS || n_ct * (C / TCH) > kNumSMs) return 0;      // the grid barrier needs every CTA resident
    if (smem_bytes(C, L) > 200 * 1024) return 0;
    Ops o;
    if (!resolve_ops(cv, nd, &o)) return 0;
    if (!al16(nd->x) || !al16(nd->out) || (cv->Z && !al16(cv->Z)) || !al16(cv->wimg_fwd) || (nd->out2 && !al16(nd->out2))) return 0;
    if (o.k_attn >= 0 && (!al16(nd->ln_w[o.k_attn]) || !al16(nd->ln_b[o.k_attn]))) return 0;
    for (int k = 0; k < nd->n_ops; ++k)
        if (nd->mask[k] && (reinterpret_cast<uintptr_t>(nd->mask[k]) & 3u)) return 0;
    if (nd->n_chain < 0 || nd->n_chain > BMNAS_MAX_SRC || (nd->out2 && !nd->chain_w)) return 0;
    for (int j = 0; j < nd->n_chain; ++j)
        if (nd->out2 && !al16(nd->chain_x[j])) return 0;
    if (!cv->mean || !cv->rstd) return 0;
    return 1;
}

extern "C" long long bmnas_mixed_small_workspace_bytes(const bmnas_conv_params* cv, const bmnas_node_params* nd) {
    using namespace ms;
    if (!cv || !nd || nd->L < 1 || nd->C < TCH) return 0;
    const long long n_ct = ((long long)cv->B * nd->L + TNC - 1) / TNC;
    return (long long)sizeof(Ws) + (n_ct * 3 * nd->C * 2 + (long long)cv->B * (nd->C / TCH) * 2) * (long long)sizeof(float);
}

extern "C" int bmnas_mixed_small_fwd(const bmnas_conv_params* cv, const bmnas_node_params* nd, void* workspace, void* stream) {
    using namespace ms;
    if (!bmnas_mixed_small_supported(cv, nd) || !workspace) return BMNAS_EINVAL;
    if (nd->training && !nd->rng_state) {
        for (int k = 0; k < nd->n_ops; ++k)
            if (nd->op_type[k] != BMNAS_OP_SUM && nd->p_drop[k] > 0.f && !nd->mask[k]) return BMNAS_EINVAL;
    }
    BMNAS_DRY_RETURN();
    Ops o;
    resolve_ops(cv, nd, &o);
    const int L = nd->L, C = nd->C;
    const int n_ct = (cv->B * L + TNC - 1) / TNC;
    const size_t smem = smem_bytes(C, L);
    using KFn = void (*)(const bmnas_conv_params, const bmnas_node_params, Ws*, const Ops, const int);
    const int li = L == 4 ? 0 : L == 8 ? 1 : 2;
    static const KFn table[3] = {k_mixed_small<4>, k_mixed_small<8>, k_mixed_small<16>};
    const KFn kern = table[li];
    static size_t configured[3] = {0, 0, 0};
    if (smem > configured[li]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return BMNAS_ELAUNCH;
        configured[li] = smem;
    }
    launch_k(kern, dim3(n_ct, C / TCH), dim3(TH), smem, (cudaStream_t)stream, *cv, *nd, reinterpret_cast<Ws*>(workspace), o, n_ct);
    BMNAS_LAUNCH_CHECK();
    return BMNAS_OK;
}
