// Warp-specialised tcgen05 weight-gradient GEMM -- the large-batch engine behind bmnas_conv_wgrad:
//
//   dW[m, f K + k] += sum_{b,l} dz[b,m,l] U[b,k,l]   (both halves f of a folded cat([t,t]) weight receive the same sum)
//   dbias[m]       += sum_{b,l} dz[b,m,l]            dz = a[m] GV + b[m] Z + c[m]   (BatchNorm backward folded in)
//
// i.e. the weight / bias gradients of Conv1d(k=1)+BatchNorm (node_operations.py:30-34, 49-53; node_search.py:59-62) that
// autograd forms in the reference.  The reduction runs over the columns n = (b, l): long (65 536 at B = 8 192) against a
// small 128 x 128 output tile, so the launch is split-K: CTA (x, y, z) owns output tile (row tile y of m, column tile x of
// k) and the reduction range z, accumulates in tensor memory and adds its partial tile with red.global.add.v4.f32.
//
// Both operands are activations in (B, rows, L) layout: a 16-byte chunk (4 consecutive l of one sample and row) is
// exactly one 16-byte K-chunk of the K-major SWIZZLE_128B operand row, so staging is load -> (fold) -> hi/lo split ->
// store with no transposition.  The kernel it replaces (k_gemm_tc<WGRAD>) staged, synchronised the CTA, issued and
// waited in turn; here
//   warps 0-7   producers: 4 units (6 x 16-byte cp.async copies each) in flight per thread = 96 KB per SM, staged through a
//               thread-private shared-memory ring; the 8 lanes of
//               a quarter warp own the 8 chunks of ONE operand row (conflict-free shared stores; 4 rows x 4 samples =
//               16 fully used sectors per load instruction); row sums of dz (the bias gradient) accumulate in registers
//   warp 8      MMA issue: waits on the stage's mbarrier, issues the 4 k-steps (x3 in 3xTF32), commits the stage back
//   epilogue    warps 0-3 after the reduction: tcgen05.ld, red.add of the partial tile (both fold halves), bias gradient
#include "common.cuh"
#include "gemm_shared.cuh"
#include "tc_ptx.cuh"

namespace bmnas {
namespace wg {
using namespace tc;

constexpr int NPW = 8, NPROD = NPW * 32;
constexpr int W_MMA = NPW;
constexpr int THREADS = (W_MMA + 1) * 32;
constexpr int BNK = 128;                          // output-tile columns (input channels k)

template <bool X3>
struct Cfg {
    static constexpr uint32_t HALF = TCM * 128;                    // 16 KB: 128 operand rows x one 128-byte reduction row
    static constexpr uint32_t OPND = X3 ? 2 * HALF : HALF;        // one operand of a stage: [hi | lo]
    static constexpr uint32_t STAGE = 2 * OPND;                    // [A (dz rows) | B (U rows)]
    static constexpr int NS = 2;
    static constexpr int D = 4;                                    // half-stage units in flight per thread (even)
    static constexpr uint32_t STG = D * 6 * NPROD * 16;           // staging ring: [D][GV0 GV1 Z0 Z1 X0 X1][thread] 16-byte slots = 96 KB
    static constexpr uint32_t DYN = NS * STAGE + STG + 1024;
};

__device__ __forceinline__ void cp_async16(uint32_t dst_s, const void* src, bool valid) {
    // src-size 0 zero-fills the 16 bytes (rows / reduction columns past the end stay exactly zero)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_s), "l"(src), "r"(valid ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

template <bool X3>
__global__ void __launch_bounds__(THREADS, 1) k_wgrad_ws(const bmnas_conv_params p, const int N, const int chunkN) {
    using CF = Cfg<X3>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t s_full[6], s_empty[6], s_done;
    __shared__ uint32_t tmem_base_s;
    __shared__ float rowsum[TCM];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = p.K, M = p.M, L = p.L, ldw = p.w_fold * p.K;
    const int row0 = blockIdx.y * TCM, col0 = blockIdx.x * BNK;
    const int r_beg = blockIdx.z * chunkN, r_end = min(N, r_beg + chunkN);
    const int n_st = (r_end - r_beg + KC - 1) / KC;             // stages (32 reduction columns each); > 0 by construction

    if (tid == 0) {
        for (int i = 0; i < CF::NS; ++i) {
            mbar_init(&s_full[i], NPW);                          // one arrival per producer warp
            mbar_init(&s_empty[i], 1);
        }
        mbar_init(&s_done, 1);
        fence_barrier_init();
    }
    if (tid < TCM) rowsum[tid] = 0.f;
    if (warp == W_MMA) tmem_alloc(&tmem_base_s, 2 * BNK);        // [big | small] accumulators (see gemm_ws.cu)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_base_s;
    pdl_wait();
    pdl_trigger();

    if (warp < NPW) {
        // =============================================================== producers
        // chunk c = tid & 7 of operand row (it * 32 + (tid >> 3)), it = 0..3: unit u of a stage = rows of it = 2u, 2u + 1
        const int c = tid & 7, rl = tid >> 3;
        const bool has_coef = p.coef_a != nullptr;
        const bool want_bias = blockIdx.x == 0;
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float ka[4], kb[4], kc_[4], rs[4];
        const float* asrc[4];                                  // U row base (sample 0) of this thread's 4 B rows
        int a_ok = 0, b_ok = 0, b_C[4];
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int m = row0 + it * 32 + rl, k = col0 + it * 32 + rl;
            ka[it] = 1.f; kb[it] = 0.f; kc_[it] = 0.f; rs[it] = 0.f;
            asrc[it] = nullptr; b_C[it] = 0;
            if (m < M) {
                a_ok |= 1 << it;
                if (has_coef) {
                    ka[it] = __ldg(p.coef_a + m);
                    kb[it] = __ldg(p.coef_b + m);
                    kc_[it] = __ldg(p.coef_c + m);
                }
            }
            if (k < K) {
                int s, kl;
                src_of(p, k, &s, &kl);
                b_ok |= 1 << it;
                asrc[it] = p.src[s] + (long long)kl * L;
                b_C[it] = p.src_C[s];
            }
        }
        // Global loads are cp.async copies into a thread-private staging ring (D half-stage units deep, tracked by cp.async
        // groups), not register loads: with register prefetch the four prefetch slots all waited on the warp's scoreboards
        // (ncu: 43 % of the samples on the first use of a loaded value) -- see gemm_ws.cu.
        constexpr int D = CF::D;
        const uint32_t stg_s = s32(smem + (size_t)CF::NS * CF::STAGE) + (uint32_t)tid * 16u;
        auto slot_s = [&](int d, int j) { return stg_s + (uint32_t)((d * 6 + j) * NPROD) * 16u; };
        const int total_q = n_st * 2;
        auto issue = [&](int q, const int d) {
            const int u = d & 1;
            if (q < total_q) {
                const int st = q >> 1;
                const int n = r_beg + st * KC + c * 4;
                const bool in = n < r_end;
                const int b = in ? n / L : 0, l0 = in ? n - b * L : 0;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int it = 2 * u + j;
                    const bool va = in && ((a_ok >> it) & 1);
                    const long long idx = va ? ((long long)b * M + row0 + it * 32 + rl) * L + l0 : 0;
                    cp_async16(slot_s(d, j), p.GV + idx, va);
                    if (has_coef) cp_async16(slot_s(d, 2 + j), p.Z + idx, va);
                    const bool vb = in && ((b_ok >> it) & 1);
                    cp_async16(slot_s(d, 4 + j), vb ? asrc[it] + (long long)b * b_C[it] * L + l0 : p.GV, vb);
                }
            }
            cp_async_commit();
        };
        auto consume = [&](int q, const int d) {
            const int u = d & 1;
            cp_async_wait<D - 1>();                              // this thread's copies of unit q have landed
            const int st = q >> 1;
            const int stage = st % CF::NS, round = st / CF::NS;
            float4 v[2], xv[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int it = 2 * u + j;
                v[j] = lds128(slot_s(d, j));
                xv[j] = lds128(slot_s(d, 4 + j));
                const bool in = r_beg + st * KC + c * 4 < r_end;  // reduction columns past the range stay exactly zero
                if (in && ((a_ok >> it) & 1)) {
                    if (has_coef) {
                        const float4 z = lds128(slot_s(d, 2 + j));
                        v[j].x = fmaf(ka[it], v[j].x, fmaf(kb[it], z.x, kc_[it]));
                        v[j].y = fmaf(ka[it], v[j].y, fmaf(kb[it], z.y, kc_[it]));
                        v[j].z = fmaf(ka[it], v[j].z, fmaf(kb[it], z.z, kc_[it]));
                        v[j].w = fmaf(ka[it], v[j].w, fmaf(kb[it], z.w, kc_[it]));
                    }
                    rs[it] += (v[j].x + v[j].y) + (v[j].z + v[j].w);
                }
            }
            if (u == 0 && round > 0) mbar_wait(&s_empty[stage], (uint32_t)(round - 1) & 1u);
            const uint32_t a_hi = s32(smem) + (uint32_t)stage * CF::STAGE, a_lo = a_hi + CF::HALF;
            const uint32_t b_hi = a_hi + CF::OPND, b_lo = b_hi + CF::HALF;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int row = (2 * u + j) * 32 + rl;
                put_chunk_fast<X3>(a_hi, a_lo, sw_off(row, c), v[j]);
                put_chunk_fast<X3>(b_hi, b_lo, sw_off(row, c), xv[j]);
            }
            if (u == 1) {
                fence_proxy_async();                             // generic-proxy writes -> visible to the tensor core
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_full[stage]);
            }
        };
#pragma unroll
        for (int d = 0; d < D; ++d) issue(d, d);
        for (int q0 = 0; q0 < total_q; q0 += D) {               // total_q and D are even: slot d always holds half u = d & 1
#pragma unroll
            for (int d = 0; d < D; ++d) {
                if (q0 + d < total_q) {
                    consume(q0 + d, d);
                    issue(q0 + d + D, d);                        // refills the slot just read
                }
            }
        }
        cp_async_wait<0>();
        // bias gradient: the 8 lanes that share a row (consecutive lanes) fold their partial row sums
        if (want_bias) {
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                float s = rs[it];
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                s += __shfl_xor_sync(0xffffffffu, s, 4);
                if (c == 0) rowsum[it * 32 + rl] = s;           // one writer per row
            }
        }
    } else {
        // =============================================================== MMA issue (warp-uniform, elected lane issues)
        const bool leader = elect_one();
        constexpr uint32_t IDESC = idesc_tf32(TCM, BNK);
        for (int st = 0; st < n_st; ++st) {
            const int stage = st % CF::NS;
            mbar_wait(&s_full[stage], (uint32_t)(st / CF::NS) & 1u);
            tc_fence_after();
            const uint32_t a_hi = s32(smem + (size_t)stage * CF::STAGE), a_lo = a_hi + CF::HALF;
            const uint32_t b_hi = a_hi + CF::OPND, b_lo = b_hi + CF::HALF;
            if (leader) {
#pragma unroll
                for (int ks = 0; ks < KC / 8; ++ks) {
                    const uint32_t ko = (uint32_t)ks * 32u;
                    const uint32_t acc = (st > 0 || ks > 0) ? 1u : 0u;
                    if (X3) {      // the two correction products accumulate apart from hi*hi (truncating tensor-core accumulate)
                        umma_tf32(tmem_d + BNK, kdesc(a_lo + ko), kdesc(b_hi + ko), IDESC, acc);
                        umma_tf32(tmem_d + BNK, kdesc(a_hi + ko), kdesc(b_lo + ko), IDESC, 1u);
                    }
                    umma_tf32(tmem_d, kdesc(a_hi + ko), kdesc(b_hi + ko), IDESC, acc);
                }
                umma_commit(&s_empty[stage]);
                if (st + 1 == n_st) umma_commit(&s_done);
            }
            __syncwarp();
        }
    }

    // ---- epilogue (warps 0-3: TMEM lane quarter = warp): partial tile -> red.add into the weight gradient
    __syncthreads();                                             // rowsum complete; every role is past its loop
    if (warp < 4) {
        mbar_wait(&s_done, 0u);
        tc_fence_after();
        const int row = warp * 32 + lane, gr = row0 + row;
        const bool row_ok = gr < M;
        int seg = 0, ml = 0;
        if (row_ok) w_row(p, gr, ldw, &seg, &ml);
        float* grow = (row_ok && p.gW[seg]) ? p.gW[seg] + (long long)ml * ldw : nullptr;
        const uint32_t t_row = tmem_d + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
        for (int g16 = 0; g16 < BNK / 16; ++g16) {
            float v[16];
            if (X3) tmem_ld16_sum<2, BNK>(t_row + (uint32_t)(g16 * 16), v, 2);
            else tmem_ld16(t_row + (uint32_t)(g16 * 16), v);
            if (grow) {
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    const int k = col0 + g16 * 16 + j4 * 4;
                    if (k < K) {
                        const float4 o = make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
                        red_add_v4(grow + k, o);
                        if (p.w_fold == 2) red_add_v4(grow + K + k, o);
                    }
                }
            }
        }
        if (blockIdx.x == 0 && row_ok && p.gbias[seg]) atomicAdd(p.gbias[seg] + ml, rowsum[row]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) tmem_dealloc(tmem_d, 2 * BNK);
}

template <bool X3>
static int launch_wg(const bmnas_conv_params* p, cudaStream_t stream) {
    using CF = Cfg<X3>;
    const int N = p->B * p->L;
    const int row_tiles = (p->M + TCM - 1) / TCM, col_tiles = (p->K + BNK - 1) / BNK;
    const int tiles = row_tiles * col_tiles;
    int splits = p->splits;
    if (splits <= 0) {
        splits = kNumSMs / tiles;
        const int maxs = (N + 4 * KC - 1) / (4 * KC);     // at least 4 stages of reduction per split
        if (splits > maxs) splits = maxs;
        if (splits < 1) splits = 1;
    }
    int chunkN = ((N + splits - 1) / splits + KC - 1) / KC * KC;
    splits = (N + chunkN - 1) / chunkN;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(k_wgrad_ws<X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CF::DYN) != cudaSuccess)
            return BMNAS_ELAUNCH;
        configured = true;
    }
    dim3 grid(col_tiles, row_tiles, splits);
    launch_k(k_wgrad_ws<X3>, grid, THREADS, CF::DYN, stream, *p, N, chunkN);
    BMNAS_LAUNCH_CHECK();
    return BMNAS_OK;
}

}  // namespace wg

bool ws_enabled();

bool wgrad_ws_eligible(const bmnas_conv_params* p) { return ws_enabled(); }

int ws_conv_wgrad(const bmnas_conv_params* p, int x3, cudaStream_t stream) {
    return x3 ? wg::launch_wg<true>(p, stream) : wg::launch_wg<false>(p, stream);
}

}  // namespace bmnas
