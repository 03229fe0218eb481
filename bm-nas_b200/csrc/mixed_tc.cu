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

#ifndef MX_UNROLL_A
#define MX_UNROLL_A 1
#endif
#ifndef MX_UNROLL_B
#define MX_UNROLL_B 1
#endif

namespace bmnas {
namespace mx {
using namespace tc;
constexpr int UNROLL_A = MX_UNROLL_A, UNROLL_B = MX_UNROLL_B;   // samples per iteration of the rolled epilogue loops (ILP vs registers)

constexpr int CT = 128;             // channels: K of the folded conv and rows of one output tile
constexpr int NT = 64;              // columns per tile
constexpr int MT = 3;               // Z row tiles
constexpr int NPROD = 128;          // producer threads
// column groups of a tile, one per epilogue warp quartet.  Measured (B=8192, 3xTF32): NQ = 2 (8 epilogue warps, 128
// registers) 176 us; NQ = 4 (16 warps, capped at 80 registers, 328 B of spills in the epilogue) 219 us -- kept at 2
constexpr int NQ = 2;
constexpr int CW = NT / NQ;         // columns per epilogue thread
constexpr int NEPI = 128 * NQ;      // epilogue threads
constexpr int W_MMA = 4 + 4 * NQ, W_TMA = W_MMA + 1;
constexpr int THREADS = (W_TMA + 1) * 32;
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
    unsigned int timeline;           // test hook: != 0 -> CTA 0 records %globaltimer stamps into tl[]
    double acc[2][MT * CT][2];       // [epoch parity][row][sum, sum of squares]
    unsigned long long tl[16];
};
#define MX_TL(i)                                                        \
    do {                                                                \
        if (tl_on) {                                                    \
            unsigned long long t__;                                     \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__));     \
            ws->tl[i] = t__;                                            \
        }                                                               \
    } while (0)

struct Ops {                         // host-resolved op list (canonical order Sum < Attn < GLU < FC)
    int k_sum, k_attn, k_glu, k_fc, fc_mish;
};

struct Shared {
    uint64_t b_full[8], b_empty[8], a_full[6], a_empty[6], t_full[2], t_empty[2];
    uint32_t tmem_base;
    uint32_t epoch;
    float gw[BMNAS_MAX_OPS];
    float rs[MT * CT], mr[MT * CT], bw[MT * CT], bb[MT * CT], bias[MT * CT];
    float red[4 * NQ][16];           // LayerNorm partial sums: [epilogue warp][sample slot | 8 + sample slot]
    union {                          // pass 2 | end of pass 1 (separated by epilogue-wide barriers)
        struct {
            alignas(16) float P[NT * MAXL];      // softmax(QK^T / sqrt C) rows of the tile's samples
            float G[NT * (CW + 1)];              // Gram window staging (row n, the CW columns of its own group)
        } a;
        float4 hst[NQ - 1][MT][CT];  // column-group exchange of the Welford triples
    } u;
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

__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NEPI) : "memory"); }
// the four epilogue warps that share a column group (128 threads): named barriers 2 .. 2 + NQ - 1
__device__ __forceinline__ void half_sync(int h) { asm volatile("bar.sync %0, 128;" ::"r"(2 + h) : "memory"); }
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
// tcgen05.ld without the wait: issue several, then one tmem_ld_wait()
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// NW consecutive fp32 columns of this thread's TMEM lane, NW in {4, 8, 16}
template <int NW>
__device__ __forceinline__ void tmem_ld_nowait(uint32_t taddr, float (&v)[NW]) {
    uint32_t r[NW];
    if (NW == 4) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
    } else if (NW == 8) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4 % NW]), "=r"(r[5 % NW]), "=r"(r[6 % NW]), "=r"(r[7 % NW])
                     : "r"(taddr));
    } else {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4 % NW]), "=r"(r[5 % NW]), "=r"(r[6 % NW]), "=r"(r[7 % NW]),
              "=r"(r[8 % NW]), "=r"(r[9 % NW]), "=r"(r[10 % NW]), "=r"(r[11 % NW]), "=r"(r[12 % NW]), "=r"(r[13 % NW]), "=r"(r[14 % NW]),
              "=r"(r[15 % NW])
            : "r"(taddr));
    }
#pragma unroll
    for (int i = 0; i < NW; ++i) v[i] = __uint_as_float(r[i]);
}
template <int NW>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[NW]) {
    if constexpr (NW == 32) {
        float a[16], b[16];
        tmem_ld_nowait<16>(taddr, a);
        tmem_ld_nowait<16>(taddr + 16, b);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            v[i] = a[i];
            v[16 + i] = b[i];
        }
    } else {
        tmem_ld_nowait<NW>(taddr, v);
        tmem_ld_wait();
    }
}
template <int NW>
__device__ __forceinline__ void tmem_st(uint32_t taddr, const float (&v)[NW]) {
    uint32_t r[NW];
#pragma unroll
    for (int i = 0; i < NW; ++i) r[i] = __float_as_uint(v[i]);
    if (NW == 4) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3])
                     : "memory");
    } else if (NW == 8) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                     "r"(r[2]), "r"(r[3]), "r"(r[4 % NW]), "r"(r[5 % NW]), "r"(r[6 % NW]), "r"(r[7 % NW])
                     : "memory");
    } else {
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
            "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4 % NW]), "r"(r[5 % NW]), "r"(r[6 % NW]), "r"(r[7 % NW]), "r"(r[8 % NW]),
            "r"(r[9 % NW]), "r"(r[10 % NW]), "r"(r[11 % NW]), "r"(r[12 % NW]), "r"(r[13 % NW]), "r"(r[14 % NW]), "r"(r[15 % NW])
            : "memory");
    }
}

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

// t[b, c, 0..L) of one sample for this thread's channel (zeros past the batch end)
template <int L>
__device__ __forceinline__ void load_xrow(const float* x, long long b, int B, int c, float4 (&r)[L / 4]) {
    const float4* xp = reinterpret_cast<const float4*>(x + (b * CT + c) * L);
#pragma unroll
    for (int j4 = 0; j4 < L / 4; ++j4) r[j4] = b < B ? __ldg(xp + j4) : make_float4(0.f, 0.f, 0.f, 0.f);
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
    const bool tl_on = ws->timeline != 0 && blockIdx.x == 0 && lane == 0;
    if (tid == 0) MX_TL(0);
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
    if (tid == 0) MX_TL(1);

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
            if (tid == 0 && it == 0) MX_TL(2);
            if (tid == 0 && it == total - 1) MX_TL(3);
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
                if (itb == 0) MX_TL(4);
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
            if (m == 0) MX_TL(5);
        }
    } else {
        // =============================================================== epilogue warps
        const int e = tid - NPROD;                  // 0..NEPI-1
        const int we = e >> 5;                      // epilogue warp 0..4*NQ-1
        const int lq = warp & 3;                    // TMEM lane quarter this warp may read
        const int h = we >> 2;                      // column group: columns [h * CW, (h + 1) * CW) of the tile
        const int c = lq * 32 + lane;               // output channel = TMEM lane
        const uint32_t lane_addr = (uint32_t)(lq * 32) << 16;
        constexpr int SH = CW / L;                  // samples per column group
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
            const int col0 = tile * NT + h * CW;                   // first global column of this thread's group
            if (w.mma) {
                const uint32_t f = buf ? fc1 : fc0;
                mbar_wait(&sh.t_full[buf], f & 1u);
                if (buf) ++fc1; else ++fc0;
                tc_fence_after();
                if (e == 0 && i == 0) MX_TL(6);
            }
            const uint32_t tz = tmem_base + (uint32_t)buf * TBUF + lane_addr;

            if (w.pass1) {
                // ---- BatchNorm statistics of rows c, C + c, 2C + c over this thread's 32 columns
                const int nv = max(0, min(CW, N - col0));           // valid columns (N % 4 == 0)
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    float v[CW];
                    tmem_ld<CW>(tz + (uint32_t)(mt * NT + h * CW), v);
                    const float bs = mt == 0 ? bias0 : mt == 1 ? bias1 : bias2;
                    float s = 0.f;
#pragma unroll
                    for (int j = 0; j < CW; ++j) {
                        v[j] += bs;
                        if (j < nv) s += v[j];
                    }
                    if (nv > 0) {
                        const float mean = s / (float)nv;
                        float m2 = 0.f;
#pragma unroll
                        for (int j = 0; j < CW; ++j) {
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
                    if (h > 0) {
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) sh.u.hst[h - 1][mt][c] = make_float4(run[mt].n, run[mt].mean, run[mt].m2, 0.f);
                    }
                    epi_sync();
                    double* acc = &ws->acc[sh.epoch & 1u][0][0];
                    if (h == 0) {
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) {
                            Wf t = run[mt];
#pragma unroll
                            for (int q = 0; q < NQ - 1; ++q) {      // fixed merge order: deterministic
                                const float4 o = sh.u.hst[q][mt][c];
                                const Wf b = {o.x, o.y, o.z};
                                t = wf_merge(t, b);
                            }
                            const double mu = (double)t.mean, n = (double)t.n;
                            atomicAdd(acc + 2 * (mt * CT + c), n * mu);
                            atomicAdd(acc + 2 * (mt * CT + c) + 1, (double)t.m2 + n * mu * mu);
                        }
                        __threadfence();
                    }
                    epi_sync();
                    if (e == 0) MX_TL(7);
                    if (e == 0) grid_barrier(ws);
                    if (e == 0) MX_TL(8);
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

            // ---- pass 2: the mixed op for this thread's channel over its 32 columns (SH samples x L positions).
            // Every loop over samples is ROLLED and per-sample state lives in tensor memory / shared memory, not in
            // registers: the unrolled form of this block was 12 k instructions with 555 spill accesses, and at one tile
            // per CTA every instruction is an instruction-cache miss (profiles/: 16 us for 32 columns).
            const long long b0 = (long long)(col0 / L);            // first sample of this half
            const uint32_t t_ov = tz + (uint32_t)(3 * NT + h * CW); // the Gram accumulator, reused as scratch for O
            float4 xn[L / 4];                                       // software-prefetched activations of the next sample
            load_xrow<L>(nd.x, b0, B, c, xn);
            if (has_attn) {
                // softmax rows of this half's samples: Gram rows h*32 .. h*32+31 live in the TMEM lanes of quarter lq == h,
                // so warp (lq == h) of each half computes them; barriers below are local to the half (128 threads)
                if (lq == (h * CW) / 32) {                       // the warp whose TMEM lanes hold the Gram rows of this group
                    float g[CW];
                    tmem_ld<CW>(t_ov, g);
#pragma unroll
                    for (int j = 0; j < CW; ++j) sh.u.a.G[c * (CW + 1) + j] = g[j];
                    __syncwarp();
                    if (c >= h * CW && c < (h + 1) * CW) {       // row c = column (sample, position) c of the tile
                        const int j0 = ((c - h * CW) / L) * L;
                        float sc[L], mxv = -INFINITY, sum = 0.f;
#pragma unroll
                        for (int j = 0; j < L; ++j) {
                            sc[j] = sh.u.a.G[c * (CW + 1) + j0 + j] * inv_sqrt_c;
                            mxv = fmaxf(mxv, sc[j]);
                        }
#pragma unroll
                        for (int j = 0; j < L; ++j) {
                            sc[j] = __expf(sc[j] - mxv);
                            sum += sc[j];
                        }
                        const float inv = 1.f / sum;
#pragma unroll
                        for (int j = 0; j < L; ++j) sh.u.a.P[c * L + j] = sc[j] * inv;
                    }
                    tc_fence_before();       // this warp's Gram reads are done before any thread overwrites the region
                }
                half_sync(h);
                tc_fence_after();
                // O[c, i] = sum_j P[i][j] t[c, j] per sample, dropout; parked in the Gram region; LayerNorm sums
#pragma unroll UNROLL_A
                for (int s = 0; s < SH; ++s) {
                    const long long b = b0 + s;
                    const bool ok = b < B;
                    float xs[L], o[L];
#pragma unroll
                    for (int j4 = 0; j4 < L / 4; ++j4) {
                        xs[j4 * 4] = xn[j4].x; xs[j4 * 4 + 1] = xn[j4].y; xs[j4 * 4 + 2] = xn[j4].z; xs[j4 * 4 + 3] = xn[j4].w;
                    }
                    load_xrow<L>(nd.x, s + 1 < SH ? b + 1 : b0, B, c, xn);   // next sample; after the last one: sample 0 for the loop below
                    const float* Pb = sh.u.a.P + (h * SH + s) * L * L;
                    float s1 = 0.f;
#pragma unroll
                    for (int i4 = 0; i4 < L / 4; ++i4) {
                        float ds[4] = {1.f, 1.f, 1.f, 1.f};
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            float acc = 0.f;
#pragma unroll
                            for (int j4 = 0; j4 < L / 4; ++j4) {
                                const float4 pq = *reinterpret_cast<const float4*>(Pb + (i4 * 4 + q) * L + j4 * 4);
                                acc = fmaf(pq.x, xs[j4 * 4], acc);
                                acc = fmaf(pq.y, xs[j4 * 4 + 1], acc);
                                acc = fmaf(pq.z, xs[j4 * 4 + 2], acc);
                                acc = fmaf(pq.w, xs[j4 * 4 + 3], acc);
                            }
                            o[i4 * 4 + q] = acc;
                        }
                        const long long e0 = (long long)c * L + i4 * 4;
                        if (ok) drop4(d_attn, b * CL + e0, (unsigned long long)(nd.sample_offset + b) * CL + e0, ds);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            o[i4 * 4 + q] *= ds[q];
                            s1 += o[i4 * 4 + q];
                        }
                    }
                    tmem_st<L>(t_ov + (uint32_t)(s * L), o);
                    s1 = warp_sum(s1);
                    if (lane == 0) sh.red[we][s] = s1;
                }
                tmem_st_wait();
                half_sync(h);
                // second LayerNorm pass: sum of squared deviations (two-pass, as the node kernels do)
#pragma unroll 1
                for (int s = 0; s < SH; ++s) {
                    const float tot = (sh.red[h * 4][s] + sh.red[h * 4 + 1][s]) + (sh.red[h * 4 + 2][s] + sh.red[h * 4 + 3][s]);
                    const float mean = tot / (float)CL;
                    float o[L];
                    tmem_ld<L>(t_ov + (uint32_t)(s * L), o);
                    float s2 = 0.f;
#pragma unroll
                    for (int i = 0; i < L; ++i) {
                        const float d = o[i] - mean;
                        s2 = fmaf(d, d, s2);
                    }
                    s2 = warp_sum(s2);
                    if (lane == 0) sh.red[we][8 + s] = s2;
                }
                half_sync(h);
            }
            if (e == 0 && i == n_items - 1) MX_TL(10);

            // ---- BatchNorm + GLU / FC, LayerNorm affine, gamma-weighted sum: one sample (L columns) per iteration
#pragma unroll UNROLL_B
            for (int s = 0; s < SH; ++s) {
                const long long b = b0 + s;
                float za[L], zb[L], zf[L], ovs[L];
                const uint32_t tcol = (uint32_t)(h * CW + s * L);
                tmem_ld_nowait<L>(tz + tcol, za);
                tmem_ld_nowait<L>(tz + (uint32_t)NT + tcol, zb);
                tmem_ld_nowait<L>(tz + (uint32_t)(2 * NT) + tcol, zf);
                if (has_attn) tmem_ld_nowait<L>(t_ov + (uint32_t)(s * L), ovs);
                float a_mean = 0.f, a_rstd = 0.f;
                if (has_attn) {
                    const float t1 = (sh.red[h * 4][s] + sh.red[h * 4 + 1][s]) + (sh.red[h * 4 + 2][s] + sh.red[h * 4 + 3][s]);
                    const float t2 = (sh.red[h * 4][8 + s] + sh.red[h * 4 + 1][8 + s]) + (sh.red[h * 4 + 2][8 + s] + sh.red[h * 4 + 3][8 + s]);
                    a_mean = t1 / (float)CL;
                    a_rstd = 1.f / sqrtf(t2 / (float)CL + kLnEps);
                }
                const bool ok = b < B;
                float xs[L];
#pragma unroll
                for (int j4 = 0; j4 < L / 4; ++j4) {
                    xs[j4 * 4] = xn[j4].x; xs[j4 * 4 + 1] = xn[j4].y; xs[j4 * 4 + 2] = xn[j4].z; xs[j4 * 4 + 3] = xn[j4].w;
                }
                if (s + 1 < SH) load_xrow<L>(nd.x, b + 1, B, c, xn);
                // folded BatchNorm constants of rows c, C + c, 2C + c and the attention LayerNorm affine of channel c: re-read
                // per sample (shared memory / L1) instead of living in 28 registers across the whole item loop
                const float r0 = sh.rs[c], m0 = sh.mr[c], g0 = sh.bw[c], h0 = sh.bb[c];
                const float r1 = sh.rs[CT + c], m1 = sh.mr[CT + c], g1 = sh.bw[CT + c], h1 = sh.bb[CT + c];
                const float r2 = sh.rs[2 * CT + c], m2 = sh.mr[2 * CT + c], g2 = sh.bw[2 * CT + c], h2 = sh.bb[2 * CT + c];
                float lnw[L], lnb[L];
#pragma unroll
                for (int j4 = 0; j4 < L / 4; ++j4) {
                    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), bq = a;
                    if (has_attn) {
                        a = __ldg(reinterpret_cast<const float4*>(nd.ln_w[ops.k_attn] + (long long)c * L) + j4);
                        bq = __ldg(reinterpret_cast<const float4*>(nd.ln_b[ops.k_attn] + (long long)c * L) + j4);
                    }
                    lnw[j4 * 4] = a.x; lnw[j4 * 4 + 1] = a.y; lnw[j4 * 4 + 2] = a.z; lnw[j4 * 4 + 3] = a.w;
                    lnb[j4 * 4] = bq.x; lnb[j4 * 4 + 1] = bq.y; lnb[j4 * 4 + 2] = bq.z; lnb[j4 * 4 + 3] = bq.w;
                }
                tmem_ld_wait();
                if (s == SH - 1) {                               // last TMEM read of this item: release the accumulator set
                    tc_fence_before();
                    mbar_arrive(&sh.t_empty[buf]);
                }
                if (ok) {
#pragma unroll
                    for (int q4 = 0; q4 < L / 4; ++q4) {
                        const int l0 = q4 * 4;
                        const long long e0 = (long long)c * L + l0, li = b * CL + e0;
                        const unsigned long long gi = (unsigned long long)(nd.sample_offset + b) * CL + e0;
                        float va[4], vb[4], vf[4], dg[4], df[4], out[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            va[q] = za[l0 + q] + bias0;
                            vb[q] = zb[l0 + q] + bias1;
                            vf[q] = zf[l0 + q] + bias2;
                        }
                        if (cv.Z) {                                // keep the pre-BatchNorm activations for the backward pass
                            float* zp = cv.Z + (b * (MT * CT) + c) * L + l0;
                            *reinterpret_cast<float4*>(zp) = make_float4(va[0], va[1], va[2], va[3]);
                            *reinterpret_cast<float4*>(zp + (long long)CT * L) = make_float4(vb[0], vb[1], vb[2], vb[3]);
                            *reinterpret_cast<float4*>(zp + (long long)2 * CT * L) = make_float4(vf[0], vf[1], vf[2], vf[3]);
                        }
                        drop4(d_glu, li, gi, dg);
                        drop4(d_fc, li, gi, df);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            float acc = 0.f;
                            if (ops.k_sum >= 0) acc = fmaf(w_sum, xs[l0 + q] + xs[l0 + q], acc);
                            if (has_attn) {
                                const float o = (ovs[l0 + q] - a_mean) * a_rstd * lnw[l0 + q] + lnb[l0 + q];
                                acc = fmaf(w_attn, o, acc);
                            }
                            const float ya = fmaf(fmaf(va[q], r0, -m0), g0, h0);
                            const float yg = fmaf(fmaf(vb[q], r1, -m1), g1, h1);
                            acc = fmaf(w_glu, ya * sigmoid_fast(yg) * dg[q], acc);
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
    if (tid == NPROD) MX_TL(11);
    tc_fence_before();
    __syncthreads();
    if (tid == 0) MX_TL(12);
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
