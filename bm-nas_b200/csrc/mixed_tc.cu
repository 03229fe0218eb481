// Fused step-node mixed op on the tcgen05 tensor cores -- bmnas_mixed_fwd (include/bmnas_b200.h).
//
//   out[b] = sum_k softmax(gamma)_k * op_k(t_b, t_b)      ops in {Sum, ScaleDotAttn, LinearGLU, ConcatFC | CatConvMish}
//
// replaces NodeMixedOp.forward and everything under it (node_operations.py:19-20, 30-39, 49-56, 75-82, 92-108, 118-120)
// for the searchable cell, where both inputs are the same tensor (node_search.py:55), in ONE persistent, cooperative,
// warp-specialised kernel.  The pre-BatchNorm activations Z = Weff t + bias (3C rows) never leave the SM unless the
// caller wants them for the backward pass.
//
// Tile = 64 columns n = (b, l) (64 / L samples).  Per tile the tensor core produces four accumulators of 128 lanes x
// 64 fp32 columns in tensor memory:
//     Z0 = W[0:C)   t   (GLU value rows)      Z1 = W[C:2C)  t   (GLU gate rows)      Z2 = W[2C:3C) t   (FC rows)
//     G  = t^T t        (Gram matrix of the tile: its diagonal L x L blocks are the attention scores q_i . k_j)
// so every contraction of the mixed op -- the two 1x1 convolutions and QK^T -- runs on tcgen05 from ONE staged copy of
// the activation tile; PV (an L x L by L x C product per sample) is 64 FMAs per output row and stays in the epilogue.
// Two accumulator sets (2 x 256 of the 512 TMEM columns) let the epilogue of tile i overlap the MMAs of tile i + 1.
//
// Warp roles (448 threads, one CTA per SM):
//   warps 0-3   producers: global fp32 t -> registers (4x4 / 8x4 transposes) -> hi/lo tf32 or bf16 -> K-major
//               SWIZZLE_128B shared-memory stages (one stage = one 128-byte reduction row per column)
//   warps 4-11  epilogue: TMEM -> registers; thread = (output channel c = TMEM lane, half of the tile's columns)
//   warp 12     MMA issue (warp-uniform loop, one elected lane issues; see tc_ptx.cuh elect_one)
//   warp 13     weight slabs by TMA bulk copy from the bmnas_wprep image (3xTF32: streamed through a 3-slab ring;
//               bf16: the whole 3C x C weight, 96 KB, stays resident) + TMEM allocation
//
// Train-mode BatchNorm needs the batch statistics of every Z row before the first output can be written:
//   pass 1  every tile: MMAs (Z only) -> per-row Welford statistics in the epilogue warps (per CTA, deterministic)
//           -> one fp64 atomicAdd pair per row and CTA -> ONE grid barrier -> mean / rstd (CTA 0 also writes them for
//           the backward, updates running_mean / running_var / num_batches_tracked)
//   pass 2  the CTA's last two tiles still sit in tensor memory and are finished directly; earlier tiles (batches
//           beyond one resident wave of 2 x gridDim tiles) recompute their MMAs, now with the Gram accumulator.
// Eval-mode BatchNorm (running statistics) is a single pass.
#include "common.cuh"
#include "gemm_shared.cuh"
#include "tc_ptx.cuh"

namespace bmnas {
namespace mx {
using namespace tc;

constexpr int CT = 128;             // channels: K of the folded conv and rows of one output tile
constexpr int NT = 64;              // columns per tile
constexpr int MT = 3;               // Z row tiles
constexpr int NPROD = 128;          // producer threads
constexpr int NEPI = 256;           // epilogue threads
constexpr int W_MMA = 12, W_TMA = 13;
constexpr int THREADS = 14 * 32;
constexpr uint32_t TBUF = 256;      // TMEM columns per accumulator set: Z0 | Z1 | Z2 | Gram
constexpr int MAXL = 16;

template <bool BF>
struct Cfg {
    static constexpr int KS = BF ? 64 : 32;                        // reduction elements per stage (one 128-byte row)
    static constexpr int NKC = CT / KS;                            // stages per tile
    static constexpr int KSTEPS = 4;                               // UMMA k-steps per stage (32 bytes each)
    static constexpr uint32_t B_HALF = NT * 128;                   // 8 KB
    static constexpr uint32_t B_ST = BF ? B_HALF : 2 * B_HALF;     // [hi | lo]
    static constexpr uint32_t A_HALF = TCM * 128;                  // 16 KB
    static constexpr uint32_t A_ST = BF ? A_HALF : 2 * A_HALF;
    static constexpr int NB = BF ? 8 : 6;                          // activation ring stages
    static constexpr int NA = BF ? MT * NKC : 3;                   // weight slots (bf16: everything resident)
    static constexpr uint32_t DYN = NB * B_ST + NA * A_ST + 1024;
};

struct Ws {                          // per-(conv, node) workspace, zeroed once by the caller
    unsigned int bar_count;
    unsigned int bar_gen;
    unsigned int epoch;
    unsigned int pad;
    double acc[2][MT * CT][2];       // [epoch parity][row][sum, sum of squares]
};

struct Ops {                         // host-resolved op list (canonical order Sum < Attn < GLU < FC)
    int k_sum, k_attn, k_glu, k_fc, fc_mish;
};

struct Shared {
    uint64_t b_full[8], b_empty[8], a_full[6], a_empty[6], t_full[2], t_empty[2];
    uint32_t tmem_base;
    uint32_t epoch;
    float gw[BMNAS_MAX_OPS];
    float rs[MT * CT], mr[MT * CT], bw[MT * CT], bb[MT * CT], bias[MT * CT];
    float P[NT * MAXL];              // softmax(QK^T / sqrt C) rows of the tile's samples
    float G[NT * 33];                // Gram window staging (row n, 32 columns of its warp's window)
    float red[8][16];                // LayerNorm partial sums: [epilogue warp][sample slot]
    float4 hst[MT][CT];              // column-half exchange of the Welford triples
};

// ---- work list shared by all roles: item i of this CTA (T tiles, two_pass = train-mode BatchNorm)
struct Item {
    int t;          // local tile index (global tile = blockIdx.x + t * gridDim.x); accumulator set = t & 1
    bool mma;       // the tensor core (re)computes the tile for this item
    bool gram;      // ... including the Gram accumulator
    bool pass1;     // statistics item
    bool keep;      // pass-1 item whose accumulators stay in TMEM for pass 2
};
__device__ __forceinline__ Item item_at(int i, int T, bool two_pass, bool has_attn) {
    Item w;
    if (!two_pass) {
        w.t = i; w.mma = true; w.gram = has_attn; w.pass1 = false; w.keep = false;
    } else if (i < T) {
        w.t = i; w.mma = true; w.pass1 = true; w.keep = i >= T - 2; w.gram = has_attn && w.keep;
    } else {
        const int j = i - T;
        w.t = T - 1 - j; w.mma = j >= 2; w.gram = has_attn; w.pass1 = false; w.keep = false;
    }
    return w;
}

__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    float a[16], b[16];
    tmem_ld16(taddr, a);
    tmem_ld16(taddr + 16, b);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        v[i] = a[i];
        v[16 + i] = b[i];
    }
}

struct Drop {                        // one dropout site (same decision as drop_v / philox_keep in the node kernels)
    int mode;                        // 0 inactive, 1 Philox, 2 injected mask
    uint2 key;
    uint32_t step_lo, step_hi, thr;
    float keep;
    const unsigned char* mask;
};
__device__ __forceinline__ Drop make_drop(const bmnas_node_params& p, int k) {
    Drop d;
    const float pd = k >= 0 ? p.p_drop[k] : 0.f;
    d.mode = (k >= 0 && p.training && pd > 0.f) ? (p.mask[k] ? 2 : 1) : 0;
    d.mask = k >= 0 ? p.mask[k] : nullptr;
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
// scales of the 4 consecutive elements at local index li / global index gi (gi % 4 == 0)
__device__ __forceinline__ void drop4(const Drop& d, long long li, unsigned long long gi, float (&ds)[4]) {
    if (d.mode == 0) {
        ds[0] = ds[1] = ds[2] = ds[3] = 1.f;
    } else if (d.mode == 1) {
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

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

// grid-wide barrier among the co-resident CTAs of this (cooperative) launch; called by ONE thread per CTA
__device__ __forceinline__ void grid_barrier(Ws* ws) {
    unsigned int gen;
    asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(gen) : "l"(&ws->bar_gen) : "memory");
    unsigned int prev;
    asm volatile("atom.acq_rel.gpu.add.u32 %0, [%1], 1;" : "=r"(prev) : "l"(&ws->bar_count) : "memory");
    if (prev == gridDim.x - 1) {
        asm volatile("st.relaxed.gpu.u32 [%0], %1;" ::"l"(&ws->bar_count), "r"(0u) : "memory");
        asm volatile("st.release.gpu.u32 [%0], %1;" ::"l"(&ws->bar_gen), "r"(gen + 1u) : "memory");
    } else {
        unsigned int g;
        do {
            asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(g) : "l"(&ws->bar_gen) : "memory");
        } while (g == gen);
    }
}

template <bool BF, int L>
__global__ void __launch_bounds__(THREADS, 1) k_mixed_fwd(const bmnas_conv_params cv, const bmnas_node_params nd, Ws* ws,
                                                          const Ops ops, const int N, const int n_tiles) {
    using CF = Cfg<BF>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smB = smem;                                   // activation ring first: the Gram A-descriptor of the last
    uint8_t* smA = smem + (size_t)CF::NB * CF::B_ST;        // stage reads 64 rows past it, into the weight ring
    __shared__ Shared sh;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int B = cv.B;
    const int G_ = (int)gridDim.x;
    const int T = (n_tiles - (int)blockIdx.x + G_ - 1) / G_;           // tiles of this CTA (>= 1: grid <= n_tiles)
    const bool two_pass = cv.bn_mode == 1;
    const bool has_attn = ops.k_attn >= 0;
    const int n_items = two_pass ? 2 * T : T;

    if (tid == 0) {
        for (int i = 0; i < CF::NB; ++i) {
            mbar_init(&sh.b_full[i], NPROD);
            mbar_init(&sh.b_empty[i], 1);
        }
        for (int i = 0; i < CF::NA; ++i) {
            mbar_init(&sh.a_full[i], 1);
            mbar_init(&sh.a_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&sh.t_full[i], 1);
            mbar_init(&sh.t_empty[i], NEPI);
        }
        fence_barrier_init();
        unsigned int e;
        asm volatile("ld.relaxed.gpu.u32 %0, [%1];" : "=r"(e) : "l"(&ws->epoch) : "memory");
        sh.epoch = e;
        // softmax(gamma) (architecture tensors are never written inside a forward pass)
        if (!nd.gamma) {
            for (int k = 0; k < nd.n_ops; ++k) sh.gw[k] = 1.f;
        } else if (nd.gamma_is_logits) {
            float mxv = -INFINITY, s = 0.f;
            for (int k = 0; k < nd.n_ops; ++k) mxv = fmaxf(mxv, nd.gamma[k]);
            for (int k = 0; k < nd.n_ops; ++k) {
                sh.gw[k] = expf(nd.gamma[k] - mxv);
                s += sh.gw[k];
            }
            for (int k = 0; k < nd.n_ops; ++k) sh.gw[k] /= s;
        } else {
            for (int k = 0; k < nd.n_ops; ++k) sh.gw[k] = nd.gamma[k];
        }
    }
    if (warp == W_TMA) tmem_alloc(&sh.tmem_base, 512);
    // conv bias and BatchNorm affine per Z row (row m: segment / local row through the conv block)
    for (int m = tid; m < MT * CT; m += THREADS) {
        int seg, ml;
        w_row(cv, m, cv.w_fold * cv.K, &seg, &ml);
        sh.bias[m] = cv.bias[seg] ? __ldg(cv.bias[seg] + ml) : 0.f;
        const int k = m < 2 * CT ? ops.k_glu : ops.k_fc;
        const int lr = m < 2 * CT ? m : m - 2 * CT;
        sh.bw[m] = __ldg(nd.bn_w[k] + lr);
        sh.bb[m] = __ldg(nd.bn_b[k] + lr);
        if (cv.bn_mode == 2) {                                  // eval: running statistics
            const float r = 1.f / sqrtf(cv.running_var[seg][ml] + cv.eps);
            const float mu = cv.running_mean[seg][ml];
            sh.rs[m] = r;
            sh.mr[m] = mu * r;
            if (blockIdx.x == 0) {
                cv.mean[m] = mu;
                cv.rstd[m] = r;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sh.tmem_base;

    // number of MMA items and the local tile of MMA item m (pass 1: tiles 0..T-1, pass 2: tiles T-3 .. 0)
    const int n_mma = two_pass ? T + max(T - 2, 0) : T;
    auto tile_of_mma = [&](int m) { return (two_pass && m >= T) ? (T - 3 - (m - T)) : m; };
    auto gram_of_mma = [&](int m) { return has_attn && (!two_pass || m >= T - 2); };

    if (warp < 4) {
        // =============================================================== producers
        // one 4-column x (4 | 8)-row register block per thread and stage: kb = 16-byte chunk of the 128-byte row,
        // cg = column group; lanes 0-7 of a quarter warp write the 8 chunks of ONE row: conflict-free 128-bit stores
        const int kb = tid & 7, cg = tid >> 3;
        constexpr int RK = BF ? 8 : 4;                            // reduction rows per block
        const float* x = nd.x;
        float4 cur[RK], nxt[RK];
        auto load_blk = [&](int m, int kc, float4 (&r)[RK]) {
            const int tl = tile_of_mma(m);
            const int n = ((int)blockIdx.x + tl * G_) * NT + cg * 4;
            if (n < N) {
                const int b = n / L, l0 = n - b * L;
                const float* src = x + ((long long)b * CT + kc * CF::KS + kb * RK) * L + l0;
#pragma unroll
                for (int j = 0; j < RK; ++j) r[j] = __ldg(reinterpret_cast<const float4*>(src + (long long)j * L));
            } else {
#pragma unroll
                for (int j = 0; j < RK; ++j) r[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        const int total = n_mma * CF::NKC;
        if (total > 0) load_blk(0, 0, cur);
        for (int it = 0; it < total; ++it) {
            const int stage = it % CF::NB, round = it / CF::NB;
            if (it + 1 < total) load_blk((it + 1) / CF::NKC, (it + 1) % CF::NKC, nxt);
            if (round > 0) mbar_wait(&sh.b_empty[stage], (uint32_t)(round - 1) & 1u);
            uint8_t* hi = smB + (size_t)stage * CF::B_ST;
            uint8_t* lo = hi + CF::B_HALF;
#pragma unroll
            for (int i = 0; i < 4; ++i) {                         // column cg*4 + i of the block
                const uint32_t off = sw_off(cg * 4 + i, kb);
                float e[RK];
#pragma unroll
                for (int j = 0; j < RK; ++j) e[j] = i == 0 ? cur[j].x : i == 1 ? cur[j].y : i == 2 ? cur[j].z : cur[j].w;
                if (BF) {
                    *reinterpret_cast<uint4*>(hi + off) = make_uint4(pack_bf16(e[0], e[1]), pack_bf16(e[2], e[3]),
                                                                      pack_bf16(e[4 % RK], e[5 % RK]), pack_bf16(e[6 % RK], e[7 % RK]));
                } else {
                    put_chunk<true>(hi, lo, off, make_float4(e[0], e[1], e[2], e[3]));
                }
            }
            fence_proxy_async();                                  // generic-proxy writes -> visible to the tensor core
            mbar_arrive(&sh.b_full[stage]);
#pragma unroll
            for (int j = 0; j < RK; ++j) cur[j] = nxt[j];
        }
    } else if (warp == W_TMA) {
        // =============================================================== weight slabs (TMA bulk copies)
        const uint8_t* img = reinterpret_cast<const uint8_t*>(cv.wimg_fwd);
        const bool leader = elect_one();
        if (BF) {
            if (leader) {
                for (int s = 0; s < CF::NA; ++s) {                // image order [row tile][k slab] = slot order
                    mbar_expect_tx(&sh.a_full[s], CF::A_ST);
                    tma_bulk_g2s(smA + (size_t)s * CF::A_ST, img + (size_t)s * CF::A_ST, CF::A_ST, &sh.a_full[s]);
                }
            }
        } else {
            uint32_t ia = 0;
            for (int m = 0; m < n_mma; ++m) {
                for (int kc = 0; kc < CF::NKC; ++kc) {
                    for (int mt = 0; mt < MT; ++mt, ++ia) {
                        const uint32_t slot = ia % CF::NA, round = ia / CF::NA;
                        if (round > 0) mbar_wait(&sh.a_empty[slot], (round - 1) & 1u);
                        if (leader) {
                            mbar_expect_tx(&sh.a_full[slot], CF::A_ST);
                            tma_bulk_g2s(smA + (size_t)slot * CF::A_ST, img + (size_t)(mt * CF::NKC + kc) * CF::A_ST, CF::A_ST,
                                         &sh.a_full[slot]);
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == W_MMA) {
        // =============================================================== MMA issue (warp-uniform, elected lane issues)
        const bool leader = elect_one();
        constexpr uint32_t IDESC = BF ? idesc_bf16(TCM, NT) : idesc_tf32(TCM, NT);
        uint32_t itb = 0, ia = 0, use0 = 0, use1 = 0;
        for (int m = 0; m < n_mma; ++m) {
            const int tl = tile_of_mma(m);
            const int buf = tl & 1;
            const uint32_t used = buf ? use1 : use0;
            if (used > 0) mbar_wait(&sh.t_empty[buf], (used - 1) & 1u);      // the epilogue has drained this set
            if (buf) ++use1; else ++use0;
            tc_fence_after();
            const uint32_t d0 = tmem_base + (uint32_t)buf * TBUF;
            const bool gram = gram_of_mma(m);
            for (int kc = 0; kc < CF::NKC; ++kc, ++itb) {
                const uint32_t stage = itb % CF::NB;
                mbar_wait(&sh.b_full[stage], (itb / CF::NB) & 1u);
                tc_fence_after();
                const uint32_t b_hi = s32(smB + (size_t)stage * CF::B_ST), b_lo = b_hi + CF::B_HALF;
                if (gram && leader) {                            // G += t^T t: the activation stage is both operands
#pragma unroll
                    for (int ks = 0; ks < CF::KSTEPS; ++ks) {
                        const uint32_t ko = (uint32_t)ks * 32u;
                        const uint32_t acc = (kc > 0 || ks > 0) ? 1u : 0u;
                        if (BF) {
                            umma_bf16(d0 + 3 * NT, kdesc(b_hi + ko), kdesc(b_hi + ko), IDESC, acc);
                        } else {
                            umma_tf32(d0 + 3 * NT, kdesc(b_lo + ko), kdesc(b_hi + ko), IDESC, acc);
                            umma_tf32(d0 + 3 * NT, kdesc(b_hi + ko), kdesc(b_lo + ko), IDESC, 1u);
                            umma_tf32(d0 + 3 * NT, kdesc(b_hi + ko), kdesc(b_hi + ko), IDESC, 1u);
                        }
                    }
                }
                for (int mt = 0; mt < MT; ++mt, ++ia) {
                    const uint32_t slot = BF ? (uint32_t)(mt * CF::NKC + kc) : ia % CF::NA;
                    if (BF) {
                        if (m == 0) mbar_wait(&sh.a_full[slot], 0u);
                    } else {
                        mbar_wait(&sh.a_full[slot], (ia / CF::NA) & 1u);
                    }
                    tc_fence_after();
                    const uint32_t a_hi = s32(smA + (size_t)slot * CF::A_ST), a_lo = a_hi + CF::A_HALF;
                    if (leader) {
#pragma unroll
                        for (int ks = 0; ks < CF::KSTEPS; ++ks) {
                            const uint32_t ko = (uint32_t)ks * 32u;
                            const uint32_t acc = (kc > 0 || ks > 0) ? 1u : 0u;
                            if (BF) {
                                umma_bf16(d0 + (uint32_t)mt * NT, kdesc(a_hi + ko), kdesc(b_hi + ko), IDESC, acc);
                            } else {
                                umma_tf32(d0 + (uint32_t)mt * NT, kdesc(a_lo + ko), kdesc(b_hi + ko), IDESC, acc);
                                umma_tf32(d0 + (uint32_t)mt * NT, kdesc(a_hi + ko), kdesc(b_lo + ko), IDESC, 1u);
                                umma_tf32(d0 + (uint32_t)mt * NT, kdesc(a_hi + ko), kdesc(b_hi + ko), IDESC, 1u);
                            }
                        }
                        if (!BF) umma_commit(&sh.a_empty[slot]);
                    }
                }
                if (leader) umma_commit(&sh.b_empty[stage]);
            }
            if (leader) umma_commit(&sh.t_full[buf]);
            __syncwarp();
        }
    } else {
        // =============================================================== epilogue warps
        const int e = tid - NPROD;                  // 0..255
        const int we = e >> 5;                      // epilogue warp 0..7
        const int lq = warp & 3;                    // TMEM lane quarter this warp may read
        const int h = we >> 2;                      // which 32 of the tile's 64 columns
        const int c = lq * 32 + lane;               // output channel = TMEM lane
        const uint32_t lane_addr = (uint32_t)(lq * 32) << 16;
        constexpr int SH = 32 / L;                  // samples per column half
        constexpr int CL = CT * L;
        uint32_t fc0 = 0, fc1 = 0;                  // t_full completions consumed per accumulator set
        Wf run[MT] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
        const Drop d_attn = make_drop(nd, ops.k_attn), d_glu = make_drop(nd, ops.k_glu), d_fc = make_drop(nd, ops.k_fc);
        const float w_sum = ops.k_sum >= 0 ? sh.gw[ops.k_sum] : 0.f, w_attn = has_attn ? sh.gw[ops.k_attn] : 0.f;
        const float w_glu = sh.gw[ops.k_glu], w_fc = sh.gw[ops.k_fc];
        const float inv_sqrt_c = 1.f / sqrtf((float)CT);
        const float bias0 = sh.bias[c], bias1 = sh.bias[CT + c], bias2 = sh.bias[2 * CT + c];

        for (int i = 0; i < n_items; ++i) {
            const Item w = item_at(i, T, two_pass, has_attn);
            const int buf = w.t & 1;
            const int tile = (int)blockIdx.x + w.t * G_;
            const int col0 = tile * NT + h * 32;                   // first global column of this thread's half
            if (w.mma) {
                const uint32_t f = buf ? fc1 : fc0;
                mbar_wait(&sh.t_full[buf], f & 1u);
                if (buf) ++fc1; else ++fc0;
                tc_fence_after();
            }
            const uint32_t tz = tmem_base + (uint32_t)buf * TBUF + lane_addr;

            if (w.pass1) {
                // ---- BatchNorm statistics of rows c, C + c, 2C + c over this thread's 32 columns
                const int nv = max(0, min(32, N - col0));           // valid columns (N % 4 == 0)
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    float v[32];
                    tmem_ld32(tz + (uint32_t)(mt * NT + h * 32), v);
                    const float bs = mt == 0 ? bias0 : mt == 1 ? bias1 : bias2;
                    float s = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        v[j] += bs;
                        if (j < nv) s += v[j];
                    }
                    if (nv > 0) {
                        const float mean = s / (float)nv;
                        float m2 = 0.f;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float d = v[j] - mean;
                            if (j < nv) m2 = fmaf(d, d, m2);
                        }
                        const Wf t = {(float)nv, mean, m2};
                        run[mt] = wf_merge(run[mt], t);
                    }
                }
                if (!w.keep) {
                    tc_fence_before();
                    mbar_arrive(&sh.t_empty[buf]);
                }
                if (i == T - 1) {
                    // ---- end of pass 1: CTA totals -> fp64 atomics -> grid barrier -> mean / rstd
                    if (h == 1) {
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) sh.hst[mt][c] = make_float4(run[mt].n, run[mt].mean, run[mt].m2, 0.f);
                    }
                    epi_sync();
                    double* acc = &ws->acc[sh.epoch & 1u][0][0];
                    if (h == 0) {
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) {
                            const float4 o = sh.hst[mt][c];
                            const Wf b = {o.x, o.y, o.z};
                            const Wf t = wf_merge(run[mt], b);
                            const double mu = (double)t.mean, n = (double)t.n;
                            atomicAdd(acc + 2 * (mt * CT + c), n * mu);
                            atomicAdd(acc + 2 * (mt * CT + c) + 1, (double)t.m2 + n * mu * mu);
                        }
                        __threadfence();
                    }
                    epi_sync();
                    if (e == 0) grid_barrier(ws);
                    epi_sync();
                    if (h == 0) {
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) {
                            const int m = mt * CT + c;
                            const double s1 = __ldcg(acc + 2 * m), s2 = __ldcg(acc + 2 * m + 1);
                            const double mu = s1 / (double)N;
                            double var = s2 / (double)N - mu * mu;
                            if (var < 0.0) var = 0.0;
                            const float mean = (float)mu, r = 1.f / sqrtf((float)var + cv.eps);
                            sh.rs[m] = r;
                            sh.mr[m] = mean * r;
                            if (blockIdx.x == 0) {
                                cv.mean[m] = mean;
                                cv.rstd[m] = r;
                                int seg, ml;
                                w_row(cv, m, cv.w_fold * cv.K, &seg, &ml);
                                if (cv.running_mean[seg]) {
                                    const float unb = (float)(var * (double)N / (double)max(N - 1, 1));
                                    cv.running_mean[seg][ml] = (1.f - cv.momentum) * cv.running_mean[seg][ml] + cv.momentum * mean;
                                    cv.running_var[seg][ml] = (1.f - cv.momentum) * cv.running_var[seg][ml] + cv.momentum * unb;
                                    if (ml == 0 && cv.num_batches_tracked[seg]) *cv.num_batches_tracked[seg] += 1;
                                }
                            }
                        }
                    }
                    epi_sync();
                    // the other parity's accumulators were last used by the previous launch: clear them for the next one
                    if (blockIdx.x == 0) {
                        double* other = &ws->acc[(sh.epoch & 1u) ^ 1u][0][0];
                        for (int q = e; q < MT * CT * 2; q += NEPI) other[q] = 0.0;
                        if (e == 0) ws->epoch = sh.epoch + 1u;
                    }
                }
                continue;
            }

            // ---- pass 2: the mixed op for this thread's channel over its 32 columns (SH samples x L positions)
            const long long b0 = (long long)(col0 / L);            // first sample of this half
            float ov[32];                                           // dropped attention output O[c, (s, i)]
            float a_mean[SH], a_rstd[SH];
            if (has_attn) {
                // softmax rows: Gram rows n < 64 live in lanes 0..63; the 4 warps of column half 0 with lane quarter 0 / 1
                // stage their 32-column window, every thread then picks the L columns of its own sample
                if (h == 0 && lq < 2) {
                    float g[32];
                    tmem_ld32(tz + (uint32_t)(3 * NT + lq * 32), g);
#pragma unroll
                    for (int j = 0; j < 32; ++j) sh.G[c * 33 + j] = g[j];
                    __syncwarp();
                    const int j0 = (lane / L) * L;
                    float sc[L], mxv = -INFINITY, sum = 0.f;
#pragma unroll
                    for (int j = 0; j < L; ++j) {
                        sc[j] = sh.G[c * 33 + j0 + j] * inv_sqrt_c;
                        mxv = fmaxf(mxv, sc[j]);
                    }
#pragma unroll
                    for (int j = 0; j < L; ++j) {
                        sc[j] = expf(sc[j] - mxv);
                        sum += sc[j];
                    }
#pragma unroll
                    for (int j = 0; j < L; ++j) sh.P[c * L + j] = sc[j] / sum;
                }
                epi_sync();
                // O[c, i] = sum_j P[i][j] t[c, j] per sample, dropout, LayerNorm statistics over (C, L)
                float s1[SH];
#pragma unroll
                for (int s = 0; s < SH; ++s) {
                    s1[s] = 0.f;
                    const long long b = b0 + s;
                    const bool ok = b < B;
                    float xv[L];
                    const float* xp = nd.x + (b * CT + c) * L;
#pragma unroll
                    for (int j4 = 0; j4 < L / 4; ++j4) {
                        const float4 q = ok ? __ldg(reinterpret_cast<const float4*>(xp) + j4) : make_float4(0.f, 0.f, 0.f, 0.f);
                        xv[j4 * 4] = q.x; xv[j4 * 4 + 1] = q.y; xv[j4 * 4 + 2] = q.z; xv[j4 * 4 + 3] = q.w;
                    }
                    const float* Pb = sh.P + (h * SH + s) * L * L;
#pragma unroll
                    for (int i4 = 0; i4 < L / 4; ++i4) {
                        float o[4], ds[4] = {1.f, 1.f, 1.f, 1.f};
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            o[q] = 0.f;
#pragma unroll
                            for (int j = 0; j < L; ++j) o[q] = fmaf(Pb[(i4 * 4 + q) * L + j], xv[j], o[q]);
                        }
                        const long long e0 = (long long)c * L + i4 * 4;
                        if (ok) drop4(d_attn, b * CL + e0, (unsigned long long)(nd.sample_offset + b) * CL + e0, ds);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            o[q] *= ds[q];
                            s1[s] += o[q];
                            ov[s * L + i4 * 4 + q] = o[q];
                        }
                    }
                }
                // two-pass LayerNorm statistics: sum over the 128 channels (4 warps of this half) x L positions
#pragma unroll
                for (int s = 0; s < SH; ++s) s1[s] = warp_sum(s1[s]);
                if (lane == 0) {
#pragma unroll
                    for (int s = 0; s < SH; ++s) sh.red[we][s] = s1[s];
                }
                epi_sync();
#pragma unroll
                for (int s = 0; s < SH; ++s) {
                    const float tot = (sh.red[h * 4][s] + sh.red[h * 4 + 1][s]) + (sh.red[h * 4 + 2][s] + sh.red[h * 4 + 3][s]);
                    a_mean[s] = tot / (float)CL;
                }
                float s2[SH];
#pragma unroll
                for (int s = 0; s < SH; ++s) {
                    s2[s] = 0.f;
#pragma unroll
                    for (int i = 0; i < L; ++i) {
                        const float d = ov[s * L + i] - a_mean[s];
                        s2[s] = fmaf(d, d, s2[s]);
                    }
                    s2[s] = warp_sum(s2[s]);
                }
                if (lane == 0) {
#pragma unroll
                    for (int s = 0; s < SH; ++s) sh.red[we][8 + s] = s2[s];
                }
                epi_sync();
#pragma unroll
                for (int s = 0; s < SH; ++s) {
                    const float tot = (sh.red[h * 4][8 + s] + sh.red[h * 4 + 1][8 + s]) + (sh.red[h * 4 + 2][8 + s] + sh.red[h * 4 + 3][8 + s]);
                    a_rstd[s] = 1.f / sqrtf(tot / (float)CL + kLnEps);
                }
            }

            // ---- BatchNorm + GLU / FC, LayerNorm affine, gamma-weighted sum; 16 columns at a time
            const float r0 = sh.rs[c], m0 = sh.mr[c], g0 = sh.bw[c], h0 = sh.bb[c];
            const float r1 = sh.rs[CT + c], m1 = sh.mr[CT + c], g1 = sh.bw[CT + c], h1 = sh.bb[CT + c];
            const float r2 = sh.rs[2 * CT + c], m2 = sh.mr[2 * CT + c], g2 = sh.bw[2 * CT + c], h2 = sh.bb[2 * CT + c];
#pragma unroll
            for (int half16 = 0; half16 < 2; ++half16) {
                float za[16], zb[16], zf[16];
                tmem_ld16(tz + (uint32_t)(0 * NT + h * 32 + half16 * 16), za);
                tmem_ld16(tz + (uint32_t)(1 * NT + h * 32 + half16 * 16), zb);
                tmem_ld16(tz + (uint32_t)(2 * NT + h * 32 + half16 * 16), zf);
                if (half16 == 1) {                               // last TMEM read of this item: release the accumulator set
                    tc_fence_before();
                    mbar_arrive(&sh.t_empty[buf]);
                }
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const int jc = half16 * 16 + q4 * 4;          // column of this thread's half
                    const int n = col0 + jc;
                    if (n < N) {
                        const long long b = n / L;
                        const int l0 = n - (int)b * L;
                        constexpr int dummy_ = 0; (void)dummy_;
                        const int s = jc / L;                      // sample slot in this half (compile time after unrolling)
                        const long long e0 = (long long)c * L + l0, li = b * CL + e0;
                        const unsigned long long gi = (unsigned long long)(nd.sample_offset + b) * CL + e0;
                        const float4 xq = __ldg(reinterpret_cast<const float4*>(nd.x + li));
                        const float xv[4] = {xq.x, xq.y, xq.z, xq.w};
                        float va[4], vb[4], vf[4], dg[4], df[4], out[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            va[q] = za[q4 * 4 + q] + bias0;
                            vb[q] = zb[q4 * 4 + q] + bias1;
                            vf[q] = zf[q4 * 4 + q] + bias2;
                        }
                        if (cv.Z) {                                // keep the pre-BatchNorm activations for the backward pass
                            float* zp = cv.Z + (b * (MT * CT) + c) * L + l0;
                            *reinterpret_cast<float4*>(zp) = make_float4(va[0], va[1], va[2], va[3]);
                            *reinterpret_cast<float4*>(zp + (long long)CT * L) = make_float4(vb[0], vb[1], vb[2], vb[3]);
                            *reinterpret_cast<float4*>(zp + (long long)2 * CT * L) = make_float4(vf[0], vf[1], vf[2], vf[3]);
                        }
                        drop4(d_glu, li, gi, dg);
                        drop4(d_fc, li, gi, df);
                        float lw[4] = {0.f, 0.f, 0.f, 0.f}, lb[4] = {0.f, 0.f, 0.f, 0.f};
                        if (has_attn) {
                            const float4 a = __ldg(reinterpret_cast<const float4*>(nd.ln_w[ops.k_attn] + e0));
                            const float4 bq = __ldg(reinterpret_cast<const float4*>(nd.ln_b[ops.k_attn] + e0));
                            lw[0] = a.x; lw[1] = a.y; lw[2] = a.z; lw[3] = a.w;
                            lb[0] = bq.x; lb[1] = bq.y; lb[2] = bq.z; lb[3] = bq.w;
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            float acc = 0.f;
                            if (ops.k_sum >= 0) acc = fmaf(w_sum, xv[q] + xv[q], acc);
                            if (has_attn) {
                                const float o = (ov[jc + q] - a_mean[s]) * a_rstd[s] * lw[q] + lb[q];
                                acc = fmaf(w_attn, o, acc);
                            }
                            const float ya = fmaf(fmaf(va[q], r0, -m0), g0, h0);
                            const float yg = fmaf(fmaf(vb[q], r1, -m1), g1, h1);
                            acc = fmaf(w_glu, ya * sigmoidf_(yg) * dg[q], acc);
                            const float yf = fmaf(fmaf(vf[q], r2, -m2), g2, h2);
                            acc = fmaf(w_fc, (ops.fc_mish ? mishf_(yf) : fmaxf(yf, 0.f)) * df[q], acc);
                            out[q] = acc;
                        }
                        *reinterpret_cast<float4*>(nd.out + li) = make_float4(out[0], out[1], out[2], out[3]);
                    }
                }
            }
        }
    }

    // ---- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == W_TMA) tmem_dealloc(tmem_base, 512);
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
    if (nd->z_off[o->k_glu] != 0 || nd->z_off[o->k_fc] != 2 * CT) return false;
    if (cv->n_seg != 2 || cv->seg_M[0] != 2 * CT || cv->seg_M[1] != CT) return false;
    return true;
}

}  // namespace mx
}  // namespace bmnas

using namespace bmnas;

extern "C" long long bmnas_mixed_workspace_bytes(void) { return (long long)sizeof(mx::Ws); }

extern "C" int bmnas_mixed_supported(const bmnas_conv_params* cv, const bmnas_node_params* nd) {
    using namespace mx;
    if (!cv || !nd) return 0;
    if (nd->C != CT || cv->K != CT || cv->M != MT * CT || cv->n_src != 1 || cv->src_C[0] != CT) return 0;
    if (!(cv->L == 4 || cv->L == 8 || cv->L == 16) || nd->L != cv->L || nd->B != cv->B || cv->B < 1) return 0;
    if (!nd->alias_xy || cv->src[0] != nd->x || nd->M != cv->M) return 0;
    if (cv->bn_mode != 1 && cv->bn_mode != 2) return 0;
    if (!cv->wimg_fwd || (cv->wimg_fmt != 0 && cv->wimg_fmt != 2)) return 0;
    if (nd->out2 || nd->n_chain) return 0;
    Ops o;
    if (!resolve_ops(cv, nd, &o)) return 0;
    if (!al16(nd->x) || !al16(nd->out) || (cv->Z && !al16(cv->Z)) || !al16(cv->wimg_fwd)) return 0;
    if (o.k_attn >= 0 && (!al16(nd->ln_w[o.k_attn]) || !al16(nd->ln_b[o.k_attn]))) return 0;
    for (int k = 0; k < nd->n_ops; ++k)
        if (nd->mask[k] && (reinterpret_cast<uintptr_t>(nd->mask[k]) & 3u)) return 0;
    if (!cv->mean || !cv->rstd) return 0;
    return 1;
}

extern "C" int bmnas_mixed_fwd(const bmnas_conv_params* cv, const bmnas_node_params* nd, void* workspace, void* stream) {
    using namespace mx;
    if (!bmnas_mixed_supported(cv, nd) || !workspace) return BMNAS_EINVAL;
    if (nd->training && !nd->rng_state) {
        for (int k = 0; k < nd->n_ops; ++k)
            if (nd->op_type[k] != BMNAS_OP_SUM && nd->p_drop[k] > 0.f && !nd->mask[k]) return BMNAS_EINVAL;
    }
    BMNAS_DRY_RETURN();
    Ops o;
    resolve_ops(cv, nd, &o);
    const int N = cv->B * cv->L;
    const int n_tiles = (N + NT - 1) / NT;
    const int grid = n_tiles < kNumSMs ? n_tiles : kNumSMs;
    const bool bf = cv->wimg_fmt == 2;
    const size_t smem = bf ? Cfg<true>::DYN : Cfg<false>::DYN;
    using KFn = void (*)(const bmnas_conv_params, const bmnas_node_params, Ws*, const Ops, const int, const int);
    const int li = cv->L == 4 ? 0 : cv->L == 8 ? 1 : 2;
    static const KFn table[2][3] = {{k_mixed_fwd<false, 4>, k_mixed_fwd<false, 8>, k_mixed_fwd<false, 16>},
                                    {k_mixed_fwd<true, 4>, k_mixed_fwd<true, 8>, k_mixed_fwd<true, 16>}};
    const KFn kern = table[bf ? 1 : 0][li];
    static bool configured[2][3] = {{false, false, false}, {false, false, false}};
    if (!configured[bf ? 1 : 0][li]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return BMNAS_ELAUNCH;
        configured[bf ? 1 : 0][li] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;       // the grid barrier needs every CTA resident (grid <= #SMs, 1 CTA / SM)
    at[0].val.cooperative = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    Ws* ws = reinterpret_cast<Ws*>(workspace);
    cudaLaunchKernelEx(&cfg, kern, *cv, *nd, ws, o, N, n_tiles);
    BMNAS_LAUNCH_CHECK();
    return BMNAS_OK;
}
