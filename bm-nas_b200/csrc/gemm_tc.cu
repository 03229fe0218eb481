// tcgen05 (5th-gen tensor core) GEMM family for the 1x1 convolutions of the fusion cell -- the sm_100a
// contraction path behind bmnas_conv_fwd / bmnas_conv_dgrad / bmnas_conv_wgrad:
//   FWD    Z[b,m,l]  = sum_k Weff[m,k] U[b,k,l] + bias[m]      (+ BN batch statistics per output row)
//   DGRAD  dU[b,k,l] = sum_m Weff[m,k] dz[b,m,l]
//   WGRAD  dW[m,k]  += sum_{b,l} dz[b,m,l] U[b,k,l],  dbias[m] += sum dz
// One CTA owns a 128-row x BN-column accumulator tile that lives in TENSOR MEMORY (TMEM); a single
// elected thread issues tcgen05.mma.cta_group::1.kind::tf32 (UMMA M=128, N=BN, K=8) on shared-memory
// operand descriptors; tcgen05.commit arrives on mbarriers that recycle the operand ring and release the
// epilogue, which reads the accumulators back with tcgen05.ld (one TMEM lane = one output row per thread,
// so bias, BatchNorm row statistics and the row-wise stores need no cross-thread traffic).
//
// Precision.  The reference computes in fp32 and the parity gate is 1e-5, which a plain TF32 product
// (10-bit mantissa) cannot meet.  Mode 3xTF32 splits every operand element into hi = tf32(v) and
// lo = v - hi while it is staged and issues three MMAs per k-step (lo*hi + hi*lo + hi*hi) into the same
// fp32 TMEM accumulator: the dropped lo*lo term is O(2^-22).  Mode 1xTF32 stages hi only (one MMA per
// k-step, half the shared memory); it is the reduced-precision mode (north_star's 2e-2 class).
//
// Operand staging.  Both operands are written in the canonical K-major SWIZZLE_NONE ("interleave")
// UMMA layout: core matrix = 8 rows x 16 bytes stored as 128 contiguous bytes, LBO (next 16-byte chunk
// along K) = 128 B, SBO (next 8-row group) = 1 KB for a 32-element K slab.  The virtual channel concat,
// the cat([t,t]) weight fold, the (B,C,L) -> (column, k) transposition and BatchNorm-backward
// (dz = a*GV + b*Z + c) are applied to the ACTIVATION operand in a global->register->shared staging pass,
// which is also where its hi/lo split happens.  The WEIGHT operand does not change inside a half step, so
// bmnas_wprep (k_wprep below) folds, transposes, splits and lays it out ONCE per forward as ready-made
// shared-memory images (one contiguous [hi 16 KB | lo 16 KB] block per 128-row x 32-k slab); the GEMM CTAs
// then fetch a slab with a single TMA bulk copy (cp.async.bulk ... mbarrier::complete_tx) two slabs ahead
// of the MMAs.  A 4-stage ring lets the next slab's loads fly while the current slab's MMAs run.
#include "common.cuh"
#include "gemm_shared.cuh"
#include "tc_ptx.cuh"

#ifdef BMNAS_TIMELINE
__device__ unsigned long long g_tl[64];
#define TL(i)                                                                                       \
    do {                                                                                            \
        if (blockIdx.x == gridDim.x / 2 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) { \
            unsigned long long t__;                                                                 \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__));                                 \
            g_tl[MODE * 20 + (i)] = t__;                                                            \
        }                                                                                           \
    } while (0)
extern "C" int bmnas_debug_timeline(unsigned long long* out) {
    return (int)cudaMemcpyFromSymbol(out, g_tl, sizeof(g_tl));
}
#else
#define TL(i)
#endif

namespace bmnas {
namespace tc {

constexpr int TCT = 256;  // threads per CTA (8 warps: all stage; all read TMEM in the epilogue)
constexpr int NST_MAX = 4;  // ring stages (3 when four would not fit in 227 KB)
constexpr int FWD = 0, DGRAD = 1, WGRAD = 2;

template <int BN, bool X3>
struct Smem {
    static constexpr uint32_t A_BYTES = TCM * KC * 4;   // 16 KB
    static constexpr uint32_t B_BYTES = BN * KC * 4;
    static constexpr uint32_t STAGE = (X3 ? 2u : 1u) * (A_BYTES + B_BYTES);
    static constexpr int NST = (NST_MAX * STAGE + 1024 <= 227u * 1024u) ? NST_MAX : 3;
    static constexpr uint32_t TOTAL = NST * STAGE + 1024;  // + alignment slack
    static constexpr uint32_t SBO = (KC / 4) * 128;     // bytes between 8-row groups
    static constexpr uint32_t LBO = 128;                // bytes between 16-byte K chunks
};

// --------------------------------------------------------------------------------------------------
// MODE FWD  : rows = output channels m (tile blockIdx.y), cols = n=(b,l) (tile blockIdx.x), red = k
// MODE DGRAD: rows = input channels k  (tile blockIdx.y), cols = n=(b,l) (tile blockIdx.x), red = m
// MODE WGRAD: rows = output channels m (tile blockIdx.y), cols = input channels k (tile blockIdx.x),
//             red = n in [blockIdx.z*chunkN, ...)
// --------------------------------------------------------------------------------------------------
template <int MODE, int BN, bool X3>
__global__ void __launch_bounds__(TCT, 1) k_gemm_tc(const bmnas_conv_params p, const int N, const int aux) {
    TL(0);
    pdl_prologue();
    using S = Smem<BN, X3>;
    constexpr int NST = S::NST;
    constexpr int LA = NST - 2;                    // TMA runs this many slabs ahead of the MMAs
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bars[2 * NST_MAX + 1];     // [0,NST) slot free, [NST_MAX, NST_MAX+NST) weight slab landed, last: done
    __shared__ uint32_t tmem_base_s;
    __shared__ float rowsum[TCM];
    __shared__ float2 halfstat[TCM];
    uint64_t* bar_free = bars;
    uint64_t* bar_full = bars + NST_MAX;
    uint64_t* bar_done = bars + 2 * NST_MAX;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = p.K, M = p.M, L = p.L, ldw = p.w_fold * p.K;
    constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;

    if (warp == 0) tmem_alloc(&tmem_base_s, TMEM_COLS);
    if (tid == 32) {
        for (int i = 0; i <= 2 * NST_MAX; ++i) mbar_init(&bars[i], 1);
        fence_barrier_init();
    }
    if (MODE == WGRAD && tid < TCM) rowsum[tid] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_base_s;
    TL(1);

    const int row0 = blockIdx.y * TCM;
    const int col0 = blockIdx.x * BN;
    const int n_rows = MODE == DGRAD ? K : M;                   // valid accumulator rows overall
    int r_beg = 0, r_end = MODE == FWD ? K : (MODE == DGRAD ? M : N);
    if (MODE == WGRAD) {
        r_beg = blockIdx.z * aux;
        r_end = min(N, r_beg + aux);
    }
    const int n_chunks = (r_end - r_beg + KC - 1) / KC;

    // prepared weight images (bmnas_wprep): slab (row tile rt, reduction slab c) = [hi 16 KB | lo 16 KB]
    const float* img = MODE == FWD ? p.wimg_fwd : (MODE == DGRAD ? p.wimg_dgrad : nullptr);
    const bool use_img = img != nullptr;
    constexpr uint32_t IMG_SLAB = 2u * S::A_BYTES;
    constexpr uint32_t IMG_COPY = (X3 ? 2u : 1u) * S::A_BYTES;
    const uint8_t* img_rt = reinterpret_cast<const uint8_t*>(img) + (size_t)blockIdx.y * (size_t)n_chunks * IMG_SLAB;
    auto tma_slab = [&](int c) {   // thread 0 only
        const int st = c % NST;
        mbar_expect_tx(&bar_full[st], IMG_COPY);
        tma_bulk_g2s(smem + (size_t)st * S::STAGE, img_rt + (size_t)c * IMG_SLAB, IMG_COPY, &bar_full[st]);
    };
    if (use_img && tid == 0) {
        for (int c = 0; c < LA && c < n_chunks; ++c) tma_slab(c);
    }

    constexpr int QA = TCM * (KC / 4) / TCT;            // 16-byte chunks of A per thread per stage (4)
    constexpr int QB = (BN * (KC / 4) + TCT - 1) / TCT; // ... of B (BN=32: 1, 64: 2, 128: 4)
    // raw staging registers: loads only in load_stage (no dependent math, so the loads stay in flight across
    // the barrier and the MMA issue); fold / BatchNorm-backward / hi-lo split happen in store_stage
    float4 ra[QA], ra2[QA], rb[QB], rb2[QB];
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool has_coef = MODE != FWD && p.coef_a != nullptr;

    auto load_stage = [&](int r0) {
        if (!use_img) {
#pragma unroll
            for (int it = 0; it < QA; ++it) {
                const int q = it * TCT + tid;
                float4 v = z4, v2 = z4;
                if (MODE == FWD) {
                    const int r8 = q & 7, kc = (q >> 3) & 7, rg = q >> 6;
                    const int m = row0 + rg * 8 + r8, k = r0 + kc * 4;
                    if (m < M && k < K) {
                        const float* w = w_row(p, m, ldw, nullptr, nullptr) + k;
                        v = __ldg(reinterpret_cast<const float4*>(w));
                        if (p.w_fold == 2) v2 = __ldg(reinterpret_cast<const float4*>(w + K));
                    }
                } else if (MODE == DGRAD) {
                    const int row = q & (TCM - 1), mc = q >> 7;
                    const int k = row0 + row, m = r0 + mc * 4;
                    if (k < K) {
                        float e[4], f[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            e[j] = f[j] = 0.f;
                            if (m + j < M) {
                                const float* w = w_row(p, m + j, ldw, nullptr, nullptr) + k;
                                e[j] = __ldg(w);
                                if (p.w_fold == 2) f[j] = __ldg(w + K);
                            }
                        }
                        v = make_float4(e[0], e[1], e[2], e[3]);
                        v2 = make_float4(f[0], f[1], f[2], f[3]);
                    }
                } else {
                    const int r8 = q & 7, kc = (q >> 3) & 7, rg = q >> 6;
                    const int m = row0 + rg * 8 + r8, n = r0 + kc * 4;
                    if (m < M && n < r_end) {
                        const long long idx = ((long long)(n / L) * M + m) * L + (n % L);
                        v = __ldg(reinterpret_cast<const float4*>(p.GV + idx));
                        if (has_coef) v2 = __ldg(reinterpret_cast<const float4*>(p.Z + idx));
                    }
                }
                ra[it] = v;
                ra2[it] = v2;
            }
        }
#pragma unroll
        for (int it = 0; it < QB; ++it) {
            const int q = it * TCT + tid;
            float4 v = z4, v2 = z4;
            if (q < BN * (KC / 4)) {
                if (MODE == WGRAD) {
                    const int r8 = q & 7, kc = (q >> 3) & 7, rg = q >> 6;
                    const int k = col0 + rg * 8 + r8, n = r0 + kc * 4;
                    if (k < K && n < r_end) {
                        int s, kl;
                        src_of(p, k, &s, &kl);
                        v = __ldg(reinterpret_cast<const float4*>(p.src[s] + ((long long)(n / L) * p.src_C[s] + kl) * L + (n % L)));
                    }
                } else {
                    const int nl = q % BN, kc = q / BN;
                    const int n = col0 + nl, r = r0 + kc * 4;
                    if (n < N) {
                        const int b = n / L, l = n - b * L;
                        float e[4], f[4];
                        if (MODE == FWD) {
                            int s = 0, kl = 0;
                            if (r < K) src_of(p, r, &s, &kl);   // a 4-chunk never straddles sources (C % 4 == 0)
                            const float* u = p.src[s] + ((long long)b * p.src_C[s] + kl) * L + l;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                e[j] = (r + j < K) ? __ldg(u + (long long)j * L) : 0.f;
                                f[j] = 0.f;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const long long idx = ((long long)b * M + r + j) * L + l;
                                e[j] = (r + j < M) ? __ldg(p.GV + idx) : 0.f;
                                f[j] = (has_coef && r + j < M) ? __ldg(p.Z + idx) : 0.f;
                            }
                        }
                        v = make_float4(e[0], e[1], e[2], e[3]);
                        v2 = make_float4(f[0], f[1], f[2], f[3]);
                    }
                }
            }
            rb[it] = v;
            rb2[it] = v2;
        }
    };

    // dz = a[m]*GV + b[m]*Z + c[m] for 4 reduction rows m..m+3 (DGRAD B) or one row m (WGRAD A)
    auto bn_fold4 = [&](float4 g, float4 z, int m, bool per_elem) {
        if (!has_coef) return g;
        if (per_elem) {
            float gg[4] = {g.x, g.y, g.z, g.w}, zz[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
            for (int j = 0; j < 4; ++j)
                gg[j] = (m + j < M) ? fmaf(__ldg(p.coef_a + m + j), gg[j], fmaf(__ldg(p.coef_b + m + j), zz[j], __ldg(p.coef_c + m + j))) : 0.f;
            return make_float4(gg[0], gg[1], gg[2], gg[3]);
        }
        const float a = __ldg(p.coef_a + m), b = __ldg(p.coef_b + m), c = __ldg(p.coef_c + m);
        return make_float4(fmaf(a, g.x, fmaf(b, z.x, c)), fmaf(a, g.y, fmaf(b, z.y, c)), fmaf(a, g.z, fmaf(b, z.z, c)),
                           fmaf(a, g.w, fmaf(b, z.w, c)));
    };

    // ---- register -> shared (fold / BN-backward, hi/lo split, canonical K-major core-matrix layout)
    auto store_stage = [&](int stage, int r0) {
        uint8_t* base = smem + (size_t)stage * S::STAGE;
        uint8_t* a_hi = base;
        uint8_t* a_lo = base + S::A_BYTES;                       // only used when X3
        uint8_t* b_hi = base + (X3 ? 2u : 1u) * S::A_BYTES;
        uint8_t* b_lo = b_hi + S::B_BYTES;
        if (!use_img) {
#pragma unroll
            for (int it = 0; it < QA; ++it) {
                const int q = it * TCT + tid;
                uint32_t off;
                float4 v = ra[it];
                if (MODE == DGRAD) {
                    const int row = q & (TCM - 1), mc = q >> 7;
                    off = sw_off(row, mc);
                    v.x += ra2[it].x; v.y += ra2[it].y; v.z += ra2[it].z; v.w += ra2[it].w;
                } else {
                    const int r8 = q & 7, kc = (q >> 3) & 7, rg = q >> 6;
                    off = sw_off(rg * 8 + r8, kc);
                    if (MODE == FWD) {
                        v.x += ra2[it].x; v.y += ra2[it].y; v.z += ra2[it].z; v.w += ra2[it].w;
                    } else {
                        const int m = row0 + rg * 8 + r8, n = r0 + kc * 4;
                        if (m < M && n < r_end) v = bn_fold4(v, ra2[it], m, false);
                        if (blockIdx.x == 0) {                      // bias gradient = row sums of dz
                            const float s = (v.x + v.y) + (v.z + v.w);
                            if (s != 0.f) atomicAdd(&rowsum[rg * 8 + r8], s);
                        }
                    }
                }
                put_chunk<X3>(a_hi, a_lo, off, v);
            }
        }
#pragma unroll
        for (int it = 0; it < QB; ++it) {
            const int q = it * TCT + tid;
            if (q < BN * (KC / 4)) {
                uint32_t off;
                float4 v = rb[it];
                if (MODE == WGRAD) {
                    const int r8 = q & 7, kc = (q >> 3) & 7, rg = q >> 6;
                    off = sw_off(rg * 8 + r8, kc);
                } else {
                    const int nl = q % BN, kc = q / BN;
                    off = sw_off(nl, kc);
                    if (MODE == DGRAD && col0 + nl < N) v = bn_fold4(v, rb2[it], r0 + kc * 4, true);
                }
                put_chunk<X3>(b_hi, b_lo, off, v);
            }
        }
    };

    // ---- main loop: stage slab c, then one thread issues its MMAs; slab c+1's loads are already in flight
    constexpr uint32_t IDESC = idesc_tf32(TCM, BN);
    if (n_chunks > 0) load_stage(r_beg);
    TL(2);
    for (int c = 0; c < n_chunks; ++c) {
        const int stage = c % NST;
        if (c >= NST) mbar_wait(&bar_free[stage], (uint32_t)((c / NST) - 1) & 1u);   // MMAs that read this slot are done
        if (use_img && tid == 0) {
            const int t = c + LA;                     // keep the weight TMA LA slabs ahead
            if (t < n_chunks) {
                if (t >= NST) mbar_wait(&bar_free[t % NST], (uint32_t)((t / NST) - 1) & 1u);
                tma_slab(t);
            }
        }
        store_stage(stage, r_beg + c * KC);
        if (c == 0) TL(3);
        if (c + 1 < n_chunks) load_stage(r_beg + (c + 1) * KC);
        fence_proxy_async();       // generic-proxy smem writes -> visible to the tensor-core (async) proxy
        __syncthreads();
        if (c == 0) TL(4);
        if (warp == 0) {             // warp-uniform issue loop, one elected lane issues (see elect_one())
            if (use_img) mbar_wait(&bar_full[stage], (uint32_t)(c / NST) & 1u);      // weight slab has landed
            if (c == 0) TL(5);
            tc_fence_after();
            const uint32_t base = s32(smem + (size_t)stage * S::STAGE);
            const uint32_t a_hi = base, a_lo = base + S::A_BYTES;
            const uint32_t b_hi = base + (X3 ? 2u : 1u) * S::A_BYTES, b_lo = b_hi + S::B_BYTES;
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < KC / 8; ++ks) {
                    const uint32_t ko = (uint32_t)ks * 32u;              // 8 tf32 = 32 bytes inside the 128-byte swizzled row
                    const uint32_t first = (c == 0 && ks == 0) ? 0u : 1u;
                    if (X3) {
                        umma_tf32(tmem_d, kdesc(a_lo + ko), kdesc(b_hi + ko), IDESC, first);
                        umma_tf32(tmem_d, kdesc(a_hi + ko), kdesc(b_lo + ko), IDESC, 1u);
                        umma_tf32(tmem_d, kdesc(a_hi + ko), kdesc(b_hi + ko), IDESC, 1u);
                    } else {
                        umma_tf32(tmem_d, kdesc(a_hi + ko), kdesc(b_hi + ko), IDESC, first);
                    }
                }
                umma_commit(&bar_free[stage]);                   // slot reusable once these MMAs have read it
                if (c + 1 == n_chunks) umma_commit(bar_done);    // accumulator complete
            }
            __syncwarp();
        }
    }

    // ---- epilogue: TMEM -> registers; thread t of warp w owns accumulator row (w & 3) * 32 + t and the
    //      column half (w >> 2)
    TL(6);
    if (n_chunks > 0) mbar_wait(bar_done, 0u);
    tc_fence_after();
    TL(7);
    const int row = (warp & 3) * 32 + lane;
    const int half = warp >> 2;
    constexpr int HC = BN / 2;                           // columns per thread
    const uint32_t t_row = tmem_d + ((uint32_t)((warp & 3) * 32) << 16);
    const int gr = row0 + row;                           // global row (m for FWD/WGRAD, k for DGRAD)
    const bool row_ok = gr < n_rows;

    if (MODE == FWD) {
        float bias = 0.f;
        int seg = 0, ml = 0;
        if (row_ok) {
            w_row(p, gr, ldw, &seg, &ml);
            if (p.bias[seg]) bias = __ldg(p.bias[seg] + ml);
        }
        float sum = 0.f, m2 = 0.f;
        int cnt = 0;
        float vals[HC];
#pragma unroll
        for (int g = 0; g < HC / 16; ++g) {
            float v[16];
            tmem_ld16(t_row + (uint32_t)(half * HC + g * 16), v);
#pragma unroll
            for (int j = 0; j < 16; ++j) vals[g * 16 + j] = (n_chunks > 0 ? v[j] : 0.f) + bias;
        }
#pragma unroll
        for (int j4 = 0; j4 < HC / 4; ++j4) {
            const int n = col0 + half * HC + j4 * 4;
            if (n < N) {                                  // N % 4 == 0 and L % 4 == 0: a 4-group is whole and in one sample
                if (row_ok) {
                    *reinterpret_cast<float4*>(p.Z + ((long long)(n / L) * M + gr) * L + (n % L)) =
                        make_float4(vals[j4 * 4], vals[j4 * 4 + 1], vals[j4 * 4 + 2], vals[j4 * 4 + 3]);
                }
                sum += (vals[j4 * 4] + vals[j4 * 4 + 1]) + (vals[j4 * 4 + 2] + vals[j4 * 4 + 3]);
                cnt += 4;
            }
        }
        if (p.bn_mode == 1) {
            const float mean = cnt ? sum / (float)cnt : 0.f;
#pragma unroll
            for (int j4 = 0; j4 < HC / 4; ++j4) {
                if (col0 + half * HC + j4 * 4 < N) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float d = vals[j4 * 4 + j] - mean;
                        m2 = fmaf(d, d, m2);
                    }
                }
            }
            // merge the two column halves of this row (Chan), then one (mean, M2) per (column tile, row)
            if (half == 1) halfstat[row] = make_float2(mean, m2);
            __syncthreads();
            if (half == 0 && row_ok) {
                const int cnt1 = max(0, min(HC, N - (col0 + HC)));
                Wf a = {(float)cnt, mean, m2};
                Wf b = {(float)cnt1, halfstat[row].x, halfstat[row].y};
                const Wf w = wf_merge(a, b);
                float* q = p.stat_part + ((long long)blockIdx.x * M + gr) * 2;
                q[0] = w.mean;
                q[1] = w.m2;
            }
        }
    } else if (MODE == DGRAD) {
        int s = 0, kl = 0;
        if (row_ok) src_of(p, gr, &s, &kl);
        float* dst = row_ok ? p.gsrc[s] : nullptr;
        const bool accum = row_ok && p.gsrc_accum[s] != 0;
#pragma unroll
        for (int g = 0; g < HC / 16; ++g) {
            float v[16];
            tmem_ld16(t_row + (uint32_t)(half * HC + g * 16), v);
            if (dst) {
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    const int n = col0 + half * HC + g * 16 + j4 * 4;
                    if (n < N) {
                        float4* d = reinterpret_cast<float4*>(dst + ((long long)(n / L) * p.src_C[s] + kl) * L + (n % L));
                        float4 o = make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
                        if (n_chunks == 0) o = z4;
                        if (accum) {
                            const float4 c_ = *d;
                            o.x += c_.x; o.y += c_.y; o.z += c_.z; o.w += c_.w;
                        }
                        *d = o;
                    }
                }
            }
        }
    } else {
        int seg = 0, ml = 0;
        if (row_ok) w_row(p, gr, ldw, &seg, &ml);
        float* grow = (row_ok && p.gW[seg]) ? p.gW[seg] + (long long)ml * ldw : nullptr;
#pragma unroll
        for (int g = 0; g < HC / 16; ++g) {
            float v[16];
            tmem_ld16(t_row + (uint32_t)(half * HC + g * 16), v);
            if (grow && n_chunks > 0) {
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    const int k = col0 + half * HC + g * 16 + j4 * 4;
                    if (k < K) {
                        const float4 o = make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
                        red_add_v4(grow + k, o);
                        if (p.w_fold == 2) red_add_v4(grow + K + k, o);
                    }
                }
            }
        }
        if (blockIdx.x == 0 && half == 0 && row_ok && p.gbias[seg] && n_chunks > 0) atomicAdd(p.gbias[seg] + ml, rowsum[row]);
    }

    // ---- teardown (+ FWD: BatchNorm statistics finalize by the last CTA of this row tile)
    TL(8);
    tc_fence_before();
    __syncthreads();
    TL(9);
    if (warp == 0) tmem_dealloc(tmem_d, TMEM_COLS);
    if (MODE == FWD) {
        if (p.bn_mode == 2) {
            if (blockIdx.x == 0 && tid < TCM) bn_eval_stats(p, row0 + tid, ldw);
            return;
        }
        if (p.bn_mode != 1) return;
        TL(10);
        const bool lastb = last_block(p.counter + blockIdx.y, gridDim.x);
        TL(11);
        if (!lastb) return;
        bn_finalize_rows(p, N, aux /* n_col_tiles */, [=](int t) { return min(BN, N - t * BN); }, row0, TCM, ldw);
        if (threadIdx.x == 0) {
#ifdef BMNAS_TIMELINE
            unsigned long long t__;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__));
            g_tl[MODE * 20 + 12] = t__;
#endif
        }
    }
}

// --------------------------------------------------------------------------------------------------
// Panel kernel (FWD / DGRAD when bmnas_wprep images exist): the design for short reductions.
//   * persistent over column tiles: CTA (x, y) owns row tile y and column tiles x, x+gridDim.x, ...; the TMEM
//     allocation, barriers and (FWD) the running BatchNorm statistics of its rows live across tiles, so the
//     number of statistics partials is gridDim.x, not the number of column tiles;
//   * the ACTIVATION operand of a whole tile (all reduction chunks = one "panel") is staged in ONE burst:
//     every thread issues all its 128-bit loads (4 reduction rows x 4 columns per block) before the first
//     dependent instruction, transposes 4x4 in registers, applies BatchNorm-backward (DGRAD), splits hi/lo
//     and writes K-major core matrices with a lane rotation that keeps the 128-bit shared stores conflict
//     free -- one global round trip per tile instead of one per 32-element chunk;
//   * the WEIGHT operand arrives by TMA bulk copies issued by the same elected thread that issues the MMAs,
//     NSTA slabs ahead; when the whole reduction fits in the ring (K_red <= 32*NSTA) the weights are loaded
//     once per CTA and stay resident across column tiles;
//   * no block-wide barrier inside the reduction: one __syncthreads after staging, tcgen05.commit -> mbarrier
//     for everything else.
// --------------------------------------------------------------------------------------------------
template <int BN, bool X3>
struct PSmem {
    static constexpr uint32_t A_ST = (X3 ? 2u : 1u) * TCM * KC * 4;   // one weight slab in the ring: [hi | lo]
    static constexpr uint32_t B_HALF = BN * KC * 4;                    // one activation chunk, hi or lo
    static constexpr uint32_t B_CH = (X3 ? 2u : 1u) * B_HALF;
    static constexpr uint32_t SBO = (KC / 4) * 128, LBO = 128;
};

template <int MODE, int BN, bool X3>
__global__ void __launch_bounds__(TCT, 1) k_gemm_panel(const bmnas_conv_params p, const int N, const int n_col_tiles,
                                                        const int nsta, const int pcap) {
    TL(0);
    pdl_prologue();
    using S = PSmem<BN, X3>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smA = smem;
    uint8_t* smB = smem + (size_t)nsta * S::A_ST;
    __shared__ uint64_t bars[2 * NST_MAX + 2];     // [0,4) slot free, [4,8) slab landed, 8: accumulator done, 9: panel consumed
    __shared__ uint32_t tmem_base_s;
    float4* halfstat = reinterpret_cast<float4*>(smB);   // the panel buffer is free once the last tile's MMAs are done
    uint64_t* bar_free = bars;
    uint64_t* bar_full = bars + NST_MAX;
    uint64_t* bar_done = bars + 2 * NST_MAX;
    uint64_t* bar_panel = bars + 2 * NST_MAX + 1;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = p.K, M = p.M, L = p.L, ldw = p.w_fold * p.K;
    // NACC independent accumulators, used round-robin by consecutive MMAs: back-to-back tcgen05.mma into the SAME
    // TMEM accumulator serialise on the accumulate latency (~130 cycles per 128x32x8 MMA measured, vs 16 cycles of
    // math); rotating over 4 accumulators overlaps them, the epilogue adds the 4 partial sums
    constexpr int NACC = 4;
    constexpr uint32_t TMEM_COLS = NACC * BN;
    if (warp == 0) tmem_alloc(&tmem_base_s, TMEM_COLS);
    if (tid == 32) {
        for (int i = 0; i < 2 * NST_MAX + 2; ++i) mbar_init(&bars[i], 1);
        fence_barrier_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_base_s;
    TL(1);

    const int row0 = blockIdx.y * TCM;
    const int n_rows = MODE == DGRAD ? K : M;                  // valid accumulator rows overall
    const int r_end = MODE == FWD ? K : M;                     // reduction extent
    const int n_chunks = (r_end + KC - 1) / KC;
    const int my_tiles = ((int)blockIdx.x < n_col_tiles) ? (n_col_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const bool resident = n_chunks <= nsta;
    const long long total_slabs = resident ? n_chunks : (long long)my_tiles * n_chunks;

    constexpr uint32_t IMG_SLAB = 2u * TCM * KC * 4;           // image slab: [hi 16 KB | lo 16 KB]
    const uint8_t* img_rt = reinterpret_cast<const uint8_t*>(MODE == FWD ? p.wimg_fwd : p.wimg_dgrad) +
                            (size_t)blockIdx.y * (size_t)n_chunks * IMG_SLAB;
    // warp 0 runs the TMA / MMA issue code warp-uniformly; `leader` (one elected lane) executes the instructions
    // themselves -- in a thread-divergent branch every tcgen05.mma costs ~157 cycles of issue (see elect_one())
    long long a_issued = 0, a_used = 0;                        // warp 0 (uniform)
    const bool leader = warp == 0 && elect_one();
    auto produce = [&]() {                                      // warp 0: keep up to nsta weight slabs in flight
        while (a_issued < total_slabs && a_issued < a_used + nsta) {
            const int st = (int)(a_issued % nsta);
            if (a_issued >= nsta) mbar_wait(&bar_free[st], (uint32_t)((a_issued / nsta) - 1) & 1u);
            if (leader) {
                mbar_expect_tx(&bar_full[st], S::A_ST);
                tma_bulk_g2s(smA + (size_t)st * S::A_ST, img_rt + (size_t)(a_issued % n_chunks) * IMG_SLAB, S::A_ST, &bar_full[st]);
            }
            ++a_issued;
        }
    };
    if (warp == 0 && my_tiles > 0) produce();

    const bool has_coef = MODE == DGRAD && p.coef_a != nullptr;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    // lane roles inside a staging unit (8 column groups x 4 k-blocks)
    const int cgw = lane & 7, kbl = lane >> 3;                 // column group inside the unit, k-block inside the unit
    constexpr int UPC = 2 * (BN / 32);                          // units per chunk
    constexpr int UB = 3;                                       // units in flight per warp
    constexpr uint32_t IDESC = idesc_tf32(TCM, BN);

    // running BatchNorm statistics of this thread's (row, column half) across the CTA's tiles
    Wf run = {0.f, 0.f, 0.f};
    const int erow = (warp & 3) * 32 + lane;                    // accumulator row of this thread in the epilogue
    const int half = warp >> 2;
    constexpr int HC = BN / 2;
    const uint32_t t_row = tmem_d + ((uint32_t)((warp & 3) * 32) << 16);
    const int gr = row0 + erow;
    const bool row_ok = gr < n_rows;
    float bias = 0.f;
    if (MODE == FWD && row_ok) {
        int seg, ml;
        w_row(p, gr, ldw, &seg, &ml);
        if (p.bias[seg]) bias = __ldg(p.bias[seg] + ml);
    }
    uint32_t ph_panel = 0;

    for (int ti = 0; ti < my_tiles; ++ti) {
        const int col0 = ((int)blockIdx.x + ti * (int)gridDim.x) * BN;
        uint32_t n_mma = 0;                                    // warp 0: MMAs issued for this tile
        for (int pc0 = 0; pc0 < n_chunks; pc0 += pcap) {
            const int npc = min(pcap, n_chunks - pc0);
            if (pc0 > 0) {                                       // the MMAs that read the previous panel must be done
                mbar_wait(bar_panel, ph_panel);
                ph_panel ^= 1u;
            }
            // ---- stage the activation panel: units of (8 column groups x 4 k-blocks) per warp
            const int units = npc * UPC;
            for (int u0 = warp; u0 < units; u0 += 8 * UB) {
                float4 g[UB][4], zz[UB][4];
#pragma unroll
                for (int ui = 0; ui < UB; ++ui) {
                    const int u = u0 + ui * 8;
#pragma unroll
                    for (int j = 0; j < 4; ++j) g[ui][j] = zz[ui][j] = z4;
                    if (u < units) {
                        const int ch = u / UPC, rem = u - ch * UPC;
                        const int kb = (rem & 1) * 4 + kbl, cg = (rem >> 1) * 8 + cgw;
                        const int n = col0 + cg * 4, r = (pc0 + ch) * KC + kb * 4;
                        if (n < N && r < r_end) {
                            const int b = n / L, l0 = n - b * L;
                            if (MODE == FWD) {
                                int s, kl;
                                src_of(p, r, &s, &kl);
                                const float* u_ = p.src[s] + ((long long)b * p.src_C[s] + kl) * L + l0;
#pragma unroll
                                for (int j = 0; j < 4; ++j) g[ui][j] = __ldg(reinterpret_cast<const float4*>(u_ + (long long)j * L));
                            } else {
                                const long long idx = ((long long)b * M + r) * L + l0;
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    if (r + j < M) {
                                        g[ui][j] = __ldg(reinterpret_cast<const float4*>(p.GV + idx + (long long)j * L));
                                        if (has_coef) zz[ui][j] = __ldg(reinterpret_cast<const float4*>(p.Z + idx + (long long)j * L));
                                    }
                                }
                            }
                        }
                    }
                }
#pragma unroll
                for (int ui = 0; ui < UB; ++ui) {
                    const int u = u0 + ui * 8;
                    if (u < units) {
                        const int ch = u / UPC, rem = u - ch * UPC;
                        const int kb = (rem & 1) * 4 + kbl, cg = (rem >> 1) * 8 + cgw;
                        if (has_coef) {
                            const int r = (pc0 + ch) * KC + kb * 4;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                if (r + j < M) {
                                    const float a = __ldg(p.coef_a + r + j), bb = __ldg(p.coef_b + r + j), c = __ldg(p.coef_c + r + j);
                                    g[ui][j].x = fmaf(a, g[ui][j].x, fmaf(bb, zz[ui][j].x, c));
                                    g[ui][j].y = fmaf(a, g[ui][j].y, fmaf(bb, zz[ui][j].y, c));
                                    g[ui][j].z = fmaf(a, g[ui][j].z, fmaf(bb, zz[ui][j].z, c));
                                    g[ui][j].w = fmaf(a, g[ui][j].w, fmaf(bb, zz[ui][j].w, c));
                                }
                            }
                            if (col0 + cg * 4 >= N) {            // padding columns stay exactly zero
#pragma unroll
                                for (int j = 0; j < 4; ++j) g[ui][j] = z4;
                            }
                        }
                        uint8_t* b_hi = smB + (size_t)ch * S::B_CH;
                        uint8_t* b_lo = b_hi + S::B_HALF;
                        // 4x4 transpose: column i of the block = (row0[i], row1[i], row2[i], row3[i]); store step s
                        // handles column (s + cgw/2) & 3, so the 8 lanes of a quarter warp (same k-block, 8 column
                        // groups) write 8 distinct rows mod 8 = 8 distinct swizzled 16-byte slots: conflict free
#pragma unroll
                        for (int s_ = 0; s_ < 4; ++s_) {
                            const int i = (s_ + (cgw >> 1)) & 3;
                            float4 v;
                            v.x = i == 0 ? g[ui][0].x : i == 1 ? g[ui][0].y : i == 2 ? g[ui][0].z : g[ui][0].w;
                            v.y = i == 0 ? g[ui][1].x : i == 1 ? g[ui][1].y : i == 2 ? g[ui][1].z : g[ui][1].w;
                            v.z = i == 0 ? g[ui][2].x : i == 1 ? g[ui][2].y : i == 2 ? g[ui][2].z : g[ui][2].w;
                            v.w = i == 0 ? g[ui][3].x : i == 1 ? g[ui][3].y : i == 2 ? g[ui][3].z : g[ui][3].w;
                            const int nl = cg * 4 + i;
                            put_chunk<X3>(b_hi, b_lo, sw_off(nl, kb), v);
                        }
                    }
                }
            }
            fence_proxy_async();       // generic-proxy smem writes -> visible to the tensor-core (async) proxy
            tc_fence_before();         // this thread's TMEM reads of the previous tile are ordered before the sync
            __syncthreads();
            if (ti == 0 && pc0 == 0) TL(4);
            if (warp == 0) {
                tc_fence_after();
                for (int c = 0; c < npc; ++c) {
                    produce();
                    const long long sidx = resident ? (pc0 + c) : a_used;
                    const int st = (int)(sidx % nsta);
                    if (!resident || ti == 0) mbar_wait(&bar_full[st], (uint32_t)(sidx / nsta) & 1u);
                    if (ti == 0 && pc0 == 0 && c == 0) TL(5);
                    if (ti == 0 && pc0 == 0 && c > 0 && c < 7) TL(12 + c);
                    tc_fence_after();
                    const uint32_t a_hi = s32(smA + (size_t)st * S::A_ST), a_lo = a_hi + TCM * KC * 4;
                    const uint32_t b_hi = s32(smB + (size_t)c * S::B_CH), b_lo = b_hi + S::B_HALF;
                    if (leader) {
#pragma unroll
                        for (int ks = 0; ks < KC / 8; ++ks) {
                            const uint32_t ko = (uint32_t)ks * 32u;          // 8 tf32 = 32 bytes inside the swizzled row
                            const uint32_t m0 = n_mma + (uint32_t)ks * (X3 ? 3u : 1u);
                            if (X3) {
                                umma_tf32(tmem_d + ((m0 + 0) % NACC) * BN, kdesc(a_lo + ko), kdesc(b_hi + ko), IDESC, m0 + 0 >= NACC);
                                umma_tf32(tmem_d + ((m0 + 1) % NACC) * BN, kdesc(a_hi + ko), kdesc(b_lo + ko), IDESC, m0 + 1 >= NACC);
                                umma_tf32(tmem_d + ((m0 + 2) % NACC) * BN, kdesc(a_hi + ko), kdesc(b_hi + ko), IDESC, m0 + 2 >= NACC);
                            } else {
                                umma_tf32(tmem_d + (m0 % NACC) * BN, kdesc(a_hi + ko), kdesc(b_hi + ko), IDESC, m0 >= NACC);
                            }
                        }
                        if (!resident) umma_commit(&bar_free[st]);           // slot reusable once these MMAs have read it
                    }
                    n_mma += (KC / 8) * (X3 ? 3u : 1u);
                    if (!resident) ++a_used;
                }
                if (leader) {
                    if (pc0 + npc >= n_chunks) umma_commit(bar_done);    // accumulator complete
                    else umma_commit(bar_panel);                         // panel buffer reusable
                }
                __syncwarp();
                if (!resident) produce();                            // next tile's first slabs fly during the epilogue
            }
        }
        if (ti == 0) TL(6);
        mbar_wait(bar_done, (uint32_t)ti & 1u);
        tc_fence_after();
        if (ti == 0) TL(7);

        // ---- epilogue: thread = (accumulator row erow, column half); accumulators that received an MMA are summed
        const int used_acc = min(NACC, n_chunks * (KC / 8) * (X3 ? 3 : 1));
        if (MODE == FWD) {
            float vals[HC];
#pragma unroll
            for (int g_ = 0; g_ < HC / 16; ++g_) {
                float v[16];
                tmem_ld16_sum<NACC, BN>(t_row + (uint32_t)(half * HC + g_ * 16), v, used_acc);
#pragma unroll
                for (int j = 0; j < 16; ++j) vals[g_ * 16 + j] = v[j] + bias;
            }
            float sum = 0.f;
            int cnt = 0;
#pragma unroll
            for (int j4 = 0; j4 < HC / 4; ++j4) {
                const int n = col0 + half * HC + j4 * 4;
                if (n < N) {                                  // N % 4 == 0 and L % 4 == 0: a 4-group is whole and in one sample
                    if (row_ok) {
                        *reinterpret_cast<float4*>(p.Z + ((long long)(n / L) * M + gr) * L + (n % L)) =
                            make_float4(vals[j4 * 4], vals[j4 * 4 + 1], vals[j4 * 4 + 2], vals[j4 * 4 + 3]);
                    }
                    sum += (vals[j4 * 4] + vals[j4 * 4 + 1]) + (vals[j4 * 4 + 2] + vals[j4 * 4 + 3]);
                    cnt += 4;
                }
            }
            if (p.bn_mode == 1 && cnt > 0) {
                const float mean = sum / (float)cnt;
                float m2 = 0.f;
#pragma unroll
                for (int j4 = 0; j4 < HC / 4; ++j4) {
                    if (col0 + half * HC + j4 * 4 < N) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float d = vals[j4 * 4 + j] - mean;
                            m2 = fmaf(d, d, m2);
                        }
                    }
                }
                const Wf t = {(float)cnt, mean, m2};
                run = wf_merge(run, t);
            }
        } else {
            int s = 0, kl = 0;
            if (row_ok) src_of(p, gr, &s, &kl);
            float* dst = row_ok ? p.gsrc[s] : nullptr;
            const bool accum = row_ok && p.gsrc_accum[s] != 0;
#pragma unroll
            for (int g_ = 0; g_ < HC / 16; ++g_) {
                float v[16];
                tmem_ld16_sum<NACC, BN>(t_row + (uint32_t)(half * HC + g_ * 16), v, used_acc);
                if (dst) {
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        const int n = col0 + half * HC + g_ * 16 + j4 * 4;
                        if (n < N) {
                            float4* d = reinterpret_cast<float4*>(dst + ((long long)(n / L) * p.src_C[s] + kl) * L + (n % L));
                            float4 o = make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
                            if (accum) {
                                const float4 c_ = *d;
                                o.x += c_.x; o.y += c_.y; o.z += c_.z; o.w += c_.w;
                            }
                            *d = o;
                        }
                    }
                }
            }
        }
        if (ti == 0) TL(8);
    }

    // ---- teardown (+ FWD: one statistics partial per CTA and row, finalize by the last CTA of the row tile)
    if (MODE == FWD && p.bn_mode == 1) {
        if (half == 1) halfstat[erow] = make_float4(run.n, run.mean, run.m2, 0.f);
    }
    tc_fence_before();
    __syncthreads();
    TL(9);
    if (warp == 0) tmem_dealloc(tmem_d, TMEM_COLS);
    if (MODE == FWD) {
        if (p.bn_mode == 2) {
            if (blockIdx.x == 0 && tid < TCM) bn_eval_stats(p, row0 + tid, ldw);
            return;
        }
        if (p.bn_mode != 1) return;
        if (half == 0 && row_ok) {
            const Wf b = {halfstat[erow].x, halfstat[erow].y, halfstat[erow].z};
            const Wf w = wf_merge(run, b);
            float* qd = p.stat_part + ((long long)blockIdx.x * M + gr) * 2;
            qd[0] = w.mean;
            qd[1] = w.m2;
        }
        TL(10);
        const bool lastb = last_block(p.counter + blockIdx.y, gridDim.x);
        TL(11);
        if (!lastb) return;
        const int GX = (int)gridDim.x, last_w = N - (n_col_tiles - 1) * BN;
        bn_finalize_rows(p, N, min(GX, n_col_tiles),
                         [=](int x) {
                             const int nt = (n_col_tiles - x + GX - 1) / GX;
                             const bool owns_last = ((n_col_tiles - 1) % GX) == x;
                             return nt * BN - (owns_last ? BN - last_w : 0);
                         },
                         row0, TCM, ldw);
#ifdef BMNAS_TIMELINE
        if (threadIdx.x == 0) {
            unsigned long long t__;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__));
            g_tl[MODE * 20 + 12] = t__;
        }
#endif
    }
}

template <int MODE, int BN, bool X3>
static int launch_panel(const bmnas_conv_params* p, int N, cudaStream_t stream) {
    using S = PSmem<BN, X3>;
    const int row_tiles = ((MODE == DGRAD ? p->K : p->M) + TCM - 1) / TCM;
    const int n_col_tiles = (N + BN - 1) / BN;
    const int n_chunks = ((MODE == FWD ? p->K : p->M) + KC - 1) / KC;
    // shared-memory plan: weight ring (whole reduction resident when it fits in 4 slabs) + activation panel
    const uint32_t budget = 225u * 1024u;
    int nsta = n_chunks < NST_MAX ? n_chunks : NST_MAX;
    int pcap = (int)((budget - (uint32_t)nsta * S::A_ST) / S::B_CH);
    if (pcap >= n_chunks) pcap = n_chunks;
    else if (n_chunks > nsta && nsta > 2) {          // streaming weights anyway: trade a ring slot for a wider panel
        const int alt = (int)((budget - (uint32_t)(nsta - 1) * S::A_ST) / S::B_CH);
        if ((n_chunks + alt - 1) / alt < (n_chunks + pcap - 1) / pcap) { --nsta; pcap = alt < n_chunks ? alt : n_chunks; }
    }
    if (pcap < 1) return BMNAS_EINVAL;
    const size_t smem = (size_t)nsta * S::A_ST + (size_t)pcap * S::B_CH + 1024;
    static size_t configured = 0;
    if (smem > configured) {
        if (cudaFuncSetAttribute(k_gemm_panel<MODE, BN, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return BMNAS_ELAUNCH;
        configured = smem;
    }
    int gx = kNumSMs / row_tiles;
    if (gx < 1) gx = 1;
    if (gx > n_col_tiles) gx = n_col_tiles;
    dim3 grid(gx, row_tiles);
    launch_k(k_gemm_panel<MODE, BN, X3>, grid, TCT, smem, stream, *p, N, n_col_tiles, nsta, pcap);
    BMNAS_LAUNCH_CHECK();
    return BMNAS_OK;
}

template <int MODE>
static int panel_dispatch(const bmnas_conv_params* p, int x3, cudaStream_t stream) {
    const int N = p->B * p->L;
    const int row_tiles = ((MODE == DGRAD ? p->K : p->M) + TCM - 1) / TCM;
    const bool wide = (long long)row_tiles * ((N + 31) / 32) > 2LL * kNumSMs;
    if (wide) return x3 ? launch_panel<MODE, 64, true>(p, N, stream) : launch_panel<MODE, 64, false>(p, N, stream);
    return x3 ? launch_panel<MODE, 32, true>(p, N, stream) : launch_panel<MODE, 32, false>(p, N, stream);
}

// --------------------------------------------------------------------------------------------------
// Weight images.  For every conv: forward image (rows m, reduction k) and dgrad image (rows k, reduction
// m) of Weff[m,k] = sum_f W[m, f*K + k], zero padded to 128-row tiles x 32-element slabs, each slab stored
// as the exact shared-memory picture the MMA descriptors expect: [hi: 16 row-groups x 8 chunks x 128 B |
// lo: same].  One thread produces one 16-byte chunk (hi and lo).
// --------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_wprep(const bmnas_wprep_params p) {
    pdl_prologue();
    const long long q = (long long)blockIdx.x * 256 + threadIdx.x;
    // the dropout step counter of this forward: every reader is a later kernel of the same stream
    if (q == 0 && p.rng_state) p.rng_state[1] += 1ull;
    if (q >= p.q_start[p.n]) return;
    int i = 0;
    while (i + 1 < p.n && q >= p.q_start[i + 1]) ++i;
    long long ql = q - p.q_start[i];
    const int M = p.M[i], K = p.K[i], fold = p.w_fold[i], nseg = p.n_seg[i], ldw = fold * K;
    auto wrow = [&](int m) -> const float* {
        int s = 0;
        while (s + 1 < nseg && m >= p.seg_M[i * BMNAS_MAX_SEG + s]) {
            m -= p.seg_M[i * BMNAS_MAX_SEG + s];
            ++s;
        }
        return p.W[i * BMNAS_MAX_SEG + s] + (long long)m * ldw;
    };
    if (p.fmt[i] == 1) {
        // plain fp32 tile-major images for the small-N FFMA GEMMs (gemm_sg.cu): FWD [ceil(M/32)][K][32] then
        // DGRAD [ceil(K/32)][M][32]; one work item = one float4 of output
        const int MT32 = (M + 31) / 32, KT32 = (K + 31) / 32;
        const long long nf = (long long)MT32 * K * 8;
        if (ql < nf) {
            if (!p.img_fwd[i]) return;
            const int c4 = (int)(ql & 7), k = (int)((ql >> 3) % K), t = (int)((ql >> 3) / K);
            float e[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int m = t * 32 + c4 * 4 + j;
                e[j] = 0.f;
                if (m < M) {
                    const float* r = wrow(m) + k;
                    e[j] = __ldg(r);
                    if (fold == 2) e[j] += __ldg(r + K);
                }
            }
            *reinterpret_cast<float4*>(p.img_fwd[i] + ql * 4) = make_float4(e[0], e[1], e[2], e[3]);
        } else {
            ql -= nf;
            if (!p.img_dgrad[i]) return;
            const int c4 = (int)(ql & 7), m = (int)((ql >> 3) % M), t = (int)((ql >> 3) / M);
            const int k4 = t * 32 + c4 * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k4 < K) {
                const float* r = wrow(m) + k4;
                v = __ldg(reinterpret_cast<const float4*>(r));
                if (fold == 2) {
                    const float4 u = __ldg(reinterpret_cast<const float4*>(r + K));
                    v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
                }
            }
            *reinterpret_cast<float4*>(p.img_dgrad[i] + ql * 4) = v;
        }
        (void)KT32;
        return;
    }
    if (p.fmt[i] == 2) {
        // bf16 forward image for the fused mixed-op kernel: slab (row tile, 64-element k slab) = 128 rows x 128 bytes,
        // SWIZZLE_128B K-major; one work item = one 16-byte chunk (8 bf16 along k).  No dgrad image in this format.
        const int KS64 = (K + 63) / 64, RT = (M + TCM - 1) / TCM;
        if (ql >= (long long)RT * KS64 * 1024 || !p.img_fwd[i]) return;
        const long long slab = ql >> 10;
        const int w = (int)(ql & 1023), rt = (int)(slab / KS64), ks = (int)(slab % KS64);
        const int r8 = w & 7, kc = (w >> 3) & 7, rg = w >> 6;
        const int m = rt * TCM + rg * 8 + r8, k = ks * 64 + kc * 8;
        float e[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            e[j] = 0.f;
            if (m < M && k + j < K) {
                const float* r = wrow(m) + k + j;
                e[j] = __ldg(r);
                if (fold == 2) e[j] += __ldg(r + K);
            }
        }
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk[j]) : "f"(e[2 * j + 1]), "f"(e[2 * j]));
        uint8_t* dst8 = reinterpret_cast<uint8_t*>(p.img_fwd[i]) + slab * (TCM * 128) + sw_off(rg * 8 + r8, kc);
        *reinterpret_cast<uint4*>(dst8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        return;
    }
    const int KSf = (K + KC - 1) / KC, RTf = (M + TCM - 1) / TCM;
    const long long nf = (long long)RTf * KSf * 1024;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    float* dst;
    if (ql < nf) {
        if (!p.img_fwd[i]) return;
        const long long slab = ql >> 10;
        const int w = (int)(ql & 1023), rt = (int)(slab / KSf), ks = (int)(slab % KSf);
        const int r8 = w & 7, kc = (w >> 3) & 7, rg = w >> 6;
        const int m = rt * TCM + rg * 8 + r8, k = ks * KC + kc * 4;
        if (m < M && k < K) {
            const float* r = wrow(m) + k;
            v = __ldg(reinterpret_cast<const float4*>(r));
            if (fold == 2) {
                const float4 u = __ldg(reinterpret_cast<const float4*>(r + K));
                v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
            }
        }
        dst = p.img_fwd[i] + slab * (2 * TCM * KC) + sw_off(rg * 8 + r8, kc) / 4;
    } else {
        ql -= nf;
        if (!p.img_dgrad[i]) return;
        const int MSd = (M + KC - 1) / KC;
        const long long slab = ql >> 10;
        const int w = (int)(ql & 1023), rt = (int)(slab / MSd), ms = (int)(slab % MSd);
        const int row = w & (TCM - 1), mc = w >> 7;
        const int k = rt * TCM + row, m = ms * KC + mc * 4;
        if (k < K) {
            float e[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                e[j] = 0.f;
                if (m + j < M) {
                    const float* r = wrow(m + j) + k;
                    e[j] = __ldg(r);
                    if (fold == 2) e[j] += __ldg(r + K);
                }
            }
            v = make_float4(e[0], e[1], e[2], e[3]);
        }
        dst = p.img_dgrad[i] + slab * (2 * TCM * KC) + sw_off(row, mc) / 4;
    }
    const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
    *reinterpret_cast<float4*>(dst) = h;
    *reinterpret_cast<float4*>(dst + TCM * KC) =
        make_float4(tf32_hi(v.x - h.x), tf32_hi(v.y - h.y), tf32_hi(v.z - h.z), tf32_hi(v.w - h.w));   // see put_chunk
}

template <int MODE, int BN, bool X3>
static int launch_tc(const bmnas_conv_params* p, dim3 grid, int N, int aux, cudaStream_t stream) {
    using S = Smem<BN, X3>;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(k_gemm_tc<MODE, BN, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::TOTAL) !=
            cudaSuccess)
            return BMNAS_ELAUNCH;
        configured = true;
    }
    launch_k(k_gemm_tc<MODE, BN, X3>, grid, TCT, S::TOTAL, stream, *p, N, aux);
    BMNAS_LAUNCH_CHECK();
    return BMNAS_OK;
}

static bool al16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }

}  // namespace tc

// The tensor-core path needs whole 16-byte groups everywhere: L % 4 == 0, K % 4 == 0, every channel
// count of the virtual concat % 4 == 0, 16-byte aligned tensors.  Otherwise the caller keeps the FFMA path.
bool tc_eligible(const bmnas_conv_params* p, int mode) {
    using namespace tc;
    if ((p->L & 3) || (p->K & 3)) return false;
    for (int i = 0; i < p->n_src; ++i)
        if ((p->src_C[i] & 3) || (mode != DGRAD && !al16(p->src[i])) || (mode == DGRAD && p->gsrc[i] && !al16(p->gsrc[i])))
            return false;
    for (int i = 0; i < p->n_seg; ++i) {
        if (mode != WGRAD && !al16(p->W[i])) return false;
        if (mode == WGRAD && p->gW[i] && !al16(p->gW[i])) return false;
    }
    if (mode == FWD && !al16(p->Z)) return false;
    if (mode != FWD && (!al16(p->GV) || (p->coef_a && !al16(p->Z)))) return false;
    return true;
}

bool ws_eligible(const bmnas_conv_params* p, int mode);
int ws_conv_fwd(const bmnas_conv_params* p, int x3, cudaStream_t stream);
int ws_conv_dgrad(const bmnas_conv_params* p, int x3, cudaStream_t stream);
bool wgrad_ws_eligible(const bmnas_conv_params* p);
int ws_conv_wgrad(const bmnas_conv_params* p, int x3, cudaStream_t stream);

int tc_conv_fwd(const bmnas_conv_params* p, int x3, cudaStream_t stream) {
    using namespace tc;
    if (p->wimg_fwd && p->wimg_fmt == 0 && ws_eligible(p, 0)) return ws_conv_fwd(p, x3, stream);   // gemm_ws.cu
    if (p->wimg_fwd) return panel_dispatch<FWD>(p, x3, stream);
    const int N = p->B * p->L;
    const int row_tiles = (p->M + TCM - 1) / TCM;
    // enough CTAs to spread the operand staging over the machine, wide tiles once the batch is large
    const int bn = (long long)row_tiles * ((N + 127) / 128) >= 2 * kNumSMs ? 128 : ((long long)row_tiles * ((N + 63) / 64) >= kNumSMs ? 64 : 32);
    dim3 grid((N + bn - 1) / bn, row_tiles);
    const int aux = (int)grid.x;
    if (bn == 128) return x3 ? launch_tc<FWD, 128, true>(p, grid, N, aux, stream) : launch_tc<FWD, 128, false>(p, grid, N, aux, stream);
    if (bn == 64) return x3 ? launch_tc<FWD, 64, true>(p, grid, N, aux, stream) : launch_tc<FWD, 64, false>(p, grid, N, aux, stream);
    return x3 ? launch_tc<FWD, 32, true>(p, grid, N, aux, stream) : launch_tc<FWD, 32, false>(p, grid, N, aux, stream);
}

int tc_conv_dgrad(const bmnas_conv_params* p, int x3, cudaStream_t stream) {
    using namespace tc;
    if (p->wimg_dgrad && p->wimg_fmt == 0 && ws_eligible(p, 1)) return ws_conv_dgrad(p, x3, stream);   // gemm_ws.cu
    if (p->wimg_dgrad) return panel_dispatch<DGRAD>(p, x3, stream);
    const int N = p->B * p->L;
    const int row_tiles = (p->K + TCM - 1) / TCM;
    const int bn = (long long)row_tiles * ((N + 127) / 128) >= 2 * kNumSMs ? 128 : ((long long)row_tiles * ((N + 63) / 64) >= kNumSMs ? 64 : 32);
    dim3 grid((N + bn - 1) / bn, row_tiles);
    if (bn == 128) return x3 ? launch_tc<DGRAD, 128, true>(p, grid, N, 0, stream) : launch_tc<DGRAD, 128, false>(p, grid, N, 0, stream);
    if (bn == 64) return x3 ? launch_tc<DGRAD, 64, true>(p, grid, N, 0, stream) : launch_tc<DGRAD, 64, false>(p, grid, N, 0, stream);
    return x3 ? launch_tc<DGRAD, 32, true>(p, grid, N, 0, stream) : launch_tc<DGRAD, 32, false>(p, grid, N, 0, stream);
}

int tc_conv_wgrad(const bmnas_conv_params* p, int x3, cudaStream_t stream) {
    using namespace tc;
    if (wgrad_ws_eligible(p)) return ws_conv_wgrad(p, x3, stream);     // wgrad_ws.cu
    const int N = p->B * p->L;
    const int row_tiles = (p->M + TCM - 1) / TCM;
    const int bn = p->K >= 128 ? 128 : (p->K > 32 ? 64 : 32);
    const int col_tiles = (p->K + bn - 1) / bn;
    const int tiles = row_tiles * col_tiles;
    int splits = p->splits;
    if (splits <= 0) {
        splits = (kNumSMs + tiles - 1) / tiles;
        const int maxs = (N + 4 * KC - 1) / (4 * KC);     // at least 4 ring stages of reduction per split
        if (splits > maxs) splits = maxs;
        if (splits < 1) splits = 1;
    }
    int chunkN = ((N + splits - 1) / splits + KC - 1) / KC * KC;
    splits = (N + chunkN - 1) / chunkN;
    dim3 grid(col_tiles, row_tiles, splits);
    if (bn == 128) return x3 ? launch_tc<WGRAD, 128, true>(p, grid, N, chunkN, stream) : launch_tc<WGRAD, 128, false>(p, grid, N, chunkN, stream);
    if (bn == 64) return x3 ? launch_tc<WGRAD, 64, true>(p, grid, N, chunkN, stream) : launch_tc<WGRAD, 64, false>(p, grid, N, chunkN, stream);
    return x3 ? launch_tc<WGRAD, 32, true>(p, grid, N, chunkN, stream) : launch_tc<WGRAD, 32, false>(p, grid, N, chunkN, stream);
}

}  // namespace bmnas

using namespace bmnas;

extern "C" long long bmnas_wimg_floats(int M, int K, int which);
extern "C" long long bmnas_wimg_floats_fmt(int M, int K, int which, int fmt) {
    if (fmt == 2) return which == 0 ? (long long)((M + tc::TCM - 1) / tc::TCM) * ((K + 63) / 64) * (tc::TCM * 128 / 4) : 0;
    if (fmt == 1) return which == 0 ? (long long)((M + 31) / 32) * K * 32 : (long long)((K + 31) / 32) * M * 32;
    return bmnas_wimg_floats(M, K, which);
}

extern "C" long long bmnas_wprep_items(int M, int K, int fmt) {
    const long long f = bmnas_wimg_floats_fmt(M, K, 0, fmt) + bmnas_wimg_floats_fmt(M, K, 1, fmt);
    return fmt == 0 ? f / 8 : f / 4;      // one work item = one 16-byte chunk of output (fmt 0: hi and lo)
}

extern "C" long long bmnas_wimg_floats(int M, int K, int which) {
    const long long slab = 2LL * tc::TCM * tc::KC;
    if (which == 0) return (long long)((M + tc::TCM - 1) / tc::TCM) * ((K + tc::KC - 1) / tc::KC) * slab;
    return (long long)((K + tc::TCM - 1) / tc::TCM) * ((M + tc::KC - 1) / tc::KC) * slab;
}

extern "C" int bmnas_wprep(const bmnas_wprep_params* p, void* stream) {
    if (!p || p->n < 1 || p->n > BMNAS_MAX_PREP) return BMNAS_EINVAL;
    long long q = 0;
    for (int i = 0; i < p->n; ++i) {
        if (p->M[i] < 1 || p->K[i] < 4 || (p->K[i] & 3) || p->n_seg[i] < 1 || p->n_seg[i] > BMNAS_MAX_SEG) return BMNAS_EINVAL;
        if (p->w_fold[i] != 1 && p->w_fold[i] != 2) return BMNAS_EINVAL;
        if (!p->img_fwd[i] && !p->img_dgrad[i]) return BMNAS_EINVAL;      // either image may be skipped (NULL)
        int m = 0;
        for (int s = 0; s < p->n_seg[i]; ++s) {
            if (!p->W[i * BMNAS_MAX_SEG + s] || (reinterpret_cast<uintptr_t>(p->W[i * BMNAS_MAX_SEG + s]) & 15u)) return BMNAS_EINVAL;
            m += p->seg_M[i * BMNAS_MAX_SEG + s];
        }
        if (m != p->M[i]) return BMNAS_EINVAL;
        if (p->q_start[i] != q) return BMNAS_EINVAL;
        if (p->fmt[i] < 0 || p->fmt[i] > 2 || (p->fmt[i] == 2 && !p->img_fwd[i])) return BMNAS_EINVAL;
        q += bmnas_wprep_items(p->M[i], p->K[i], p->fmt[i]);
    }
    if (p->q_start[p->n] != q) return BMNAS_EINVAL;
    BMNAS_DRY_RETURN();
    launch_k(tc::k_wprep, (unsigned)((q + 255) / 256), 256, 0, (cudaStream_t)stream, *p);
    BMNAS_LAUNCH_CHECK();
    return BMNAS_OK;
}
