// Small-N fp32 GEMMs for the 1x1 convolutions at the reference batch (B*L <= a few thousand columns):
//   FWD    Z[b,m,l]  = sum_k Weff[m,k] U[b,k,l] + bias[m]      (+ BN batch statistics per output row)
//   DGRAD  dU[b,k,l] = sum_m Weff[m,k] dz[b,m,l],   dz = coef_a*GV + coef_b*Z + coef_c  (BatchNorm backward folded in)
// Why not the tensor cores here: with fp32 parity (3xTF32) a 128-row UMMA costs ~130 cycles per 128x32x8
// instruction whatever N is (the A operand is re-fetched from shared memory for every instruction), so a
// K = 128..384 reduction is 48..144 dependent-issue MMAs = 3..9 us per CTA (measured, profiles/), while the
// whole problem is 75 MFLOP = ~1.3 us of FFMA spread over 144 CTAs.  The tcgen05 panel kernel (gemm_tc.cu)
// takes over when the column count is large enough to amortise that.
//
// Design (latency first):
//   * operands arrive with cp.async (16-byte LDGSTS, no register staging, everything in flight at once):
//     the weight tile comes from the plain-fp32 tile-major images bmnas_wprep writes once per forward (fmt 1:
//     [row tile of 32][reduction][32], cat([t,t]) fold already applied) as one contiguous block per 32 rows,
//     the activation tile straight from the (B,C,L) tensors -- a 32-column tile is 32/L whole samples whose
//     (row, l) blocks are contiguous, 4 consecutive l are one 16-byte chunk and land exactly where the FFMA
//     loop wants them (k-major rows, columns contiguous), so no transposition;
//   * the WHOLE reduction extent (<= 384 rows) is staged in one burst: one global round trip per CTA;
//   * DGRAD applies the BatchNorm-backward fold in place in shared memory (each thread transforms the chunks
//     it copied itself: no extra barrier);
//   * 64x32 (or 32x32) output tiles, 128 threads, 4x4 (2x4) register micro-tiles, register double buffering in
//     the FFMA loop:
//     144 CTAs for the NTU node conv = one CTA per SM, one wave;
//   * FWD epilogue: bias, 128-bit stores, per-tile (mean, M2) -> last CTA of the row tile finalises BatchNorm.
#include "common.cuh"
#include "gemm_shared.cuh"

#ifdef BMNAS_TIMELINE
__device__ unsigned long long g_tl_sg[64];
#define TLS(i)                                                                                  \
    do {                                                                                        \
        if (blockIdx.x == gridDim.x / 2 && blockIdx.y == 0 && threadIdx.x == 0) {               \
            unsigned long long t__;                                                             \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__));                             \
            g_tl_sg[MODE * 20 + (i)] = t__;                                                     \
        }                                                                                       \
    } while (0)
extern "C" int bmnas_debug_timeline_sg(unsigned long long* out) {
    return (int)cudaMemcpyFromSymbol(out, g_tl_sg, sizeof(g_tl_sg));
}
#else
#define TLS(i)
#endif

namespace bmnas {
namespace sg {

constexpr int ST = 128;       // threads.  The FFMA loop is shared-memory bound at this tile size (a 128-bit LDS costs 4
                              // quarter-warp wavefronts even when the quarters read the same addresses): 4x4 micro-tiles
                              // on 128 threads measured 2.65 us for the 64x32x128 tile, 2x4 on 256 threads 3.07 us
constexpr int TN = 32;        // columns per CTA
constexpr int KCS = 384;      // reduction rows staged per pass
constexpr int FWD = 0, DGRAD = 1;
constexpr int SPAD = 8;       // floats between two sample blocks of the activation tile (bank spread)

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
// TMA bulk copy global -> shared (size multiple of 16 bytes), completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst_smem)),
                 "l"(src), "r"(bytes), "r"(s32(bar))
                 : "memory");
}

// shared-memory plan: As [TM/32][KC][32] | Bs [32/L samples][KC*L + SPAD] | Bz (DGRAD with coef) same as Bs
template <int MODE, int TM>
__host__ __device__ inline size_t smem_floats(int kc, int L, bool coef) {
    const size_t bsz = (size_t)(TN / L) * ((size_t)kc * L + SPAD);
    return (size_t)kc * TM + bsz * ((MODE == DGRAD && coef) ? 2 : 1) + ((MODE == DGRAD && coef) ? 3 * (size_t)kc : 0);
}

// MODE FWD  : rows = output channels m, cols = n=(b,l), reduction k (channels of the virtual concat)
// MODE DGRAD: rows = input channels k,  cols = n=(b,l), reduction m (stacked output channels)
template <int MODE, int TM>
__global__ void __launch_bounds__(ST) k_sg(const bmnas_conv_params p, const int N, const int n_col_tiles, const int KC,
                                           const int lshift) {
    TLS(0);
    // early section (p.early_ok: the weight image was written at least two kernels ago): barrier setup, bias fetch and
    // the TMA copies of the weight tile are issued BEFORE pdl_wait() and overlap the preceding kernel
    bool waited = !p.early_ok;
    if (waited) pdl_prologue();
    TLS(8);
    constexpr int MT = TM / 16;                   // rows per thread
    constexpr int NH = TM / 32;                   // 32-row halves of the weight tile
    extern __shared__ __align__(128) float smem[];
    __shared__ uint64_t bar;
    const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
    const int r0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    const int K = p.K, M = p.M, L = p.L, ldw = p.w_fold * p.K;
    const int R = MODE == FWD ? K : M;            // reduction extent
    const int n_rows = MODE == FWD ? M : K;       // valid output rows
    const bool coef = MODE == DGRAD && p.coef_a != nullptr;
    const int spt = TN >> lshift;                 // samples per tile
    const int sstr = KC * L + SPAD;               // floats between sample blocks
    float* As = smem;                             // [NH][KC][32]
    float* Bs = smem + (size_t)KC * TM;           // [spt][sstr]
    float* Bz = Bs + (size_t)spt * sstr;          // DGRAD with coef only
    float* cf = Bz + (size_t)spt * sstr;          // DGRAD with coef only: coef_a | coef_b | coef_c of the staged rows
    const float* img = MODE == FWD ? p.wimg_fwd : p.wimg_dgrad;   // [row tile of 32][R][32]
    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();

    float acc[MT][4];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const int b0 = n0 >> lshift;                  // first sample of the tile
    const int nB = p.B;
    const int nsv = min(spt, nB - b0);            // valid samples of this tile

    // FWD epilogue constants fetched while the copies fly
    float bias[MT];
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        bias[i] = 0.f;
        if (MODE == FWD) {
            const int m = r0 + ty * MT + i;
            if (m < n_rows) {
                int s, ml;
                w_row(p, m, ldw, &s, &ml);
                if (p.bias[s]) bias[i] = __ldg(p.bias[s] + ml);
            }
        }
    }
    TLS(9);
    uint32_t phase = 0;
    for (int k0 = 0; k0 < R; k0 += KC) {
        const int kc = min(KC, R - k0);
        if (k0) __syncthreads();
        // ---- one elected thread issues every copy of the pass as TMA bulk transfers:
        //      weight tile = NH contiguous blocks of the tile-major image; activation tile = for every sample of
        //      the tile and every source tensor one contiguous (rows x L) block (a 32-column tile is 32/L whole samples)
        if (tid < 32) {
            if (tid == 0) {
                uint32_t bytes = 0;
#pragma unroll
                for (int h = 0; h < NH; ++h)
                    if ((int)(blockIdx.y * NH + h) * 32 < n_rows) bytes += (uint32_t)kc * 128u;
                bytes += (uint32_t)nsv * (uint32_t)kc * (uint32_t)L * 4u * ((MODE == DGRAD && coef) ? 2u : 1u);
                mbar_expect_tx(&bar, bytes);
            }
            __syncwarp();
            // lanes issue in parallel (a bulk-copy issue costs ~100 ns): lane si < nsv -> the blocks of sample si,
            // lanes 16.. -> the weight halves
            if (tid >= 16 && tid < 16 + NH) {
                const int h = tid - 16;
                if ((int)(blockIdx.y * NH + h) * 32 < n_rows)
                    bulk_g2s(As + (size_t)h * KC * 32, img + ((long long)(blockIdx.y * NH + h) * R + k0) * 32, (uint32_t)kc * 128u, &bar);
            }
        }
        if (!waited) {       // everything below reads what the preceding kernel wrote
            pdl_prologue();
            waited = true;
        }
        if (tid < 16) {
            for (int si = tid; si < nsv; si += 16) {
                const int b = b0 + si;
                if (MODE == FWD) {
                    int kk = 0;                      // rows [k0, k0+kc) of the virtual concat, source by source
                    while (kk < kc) {
                        int s, kl;
                        src_of(p, k0 + kk, &s, &kl);
                        const int rows = min(kc - kk, p.src_C[s] - kl);
                        bulk_g2s(Bs + (size_t)si * sstr + kk * L, p.src[s] + ((long long)b * p.src_C[s] + kl) * L,
                                 (uint32_t)rows * (uint32_t)L * 4u, &bar);
                        kk += rows;
                    }
                } else {
                    const long long idx = ((long long)b * M + k0) * L;
                    bulk_g2s(Bs + (size_t)si * sstr, p.GV + idx, (uint32_t)kc * (uint32_t)L * 4u, &bar);
                    if (coef) bulk_g2s(Bz + (size_t)si * sstr, p.Z + idx, (uint32_t)kc * (uint32_t)L * 4u, &bar);
                }
            }
        }
        // tiles hanging over the edge: zero the parts no copy writes
#pragma unroll
        for (int h = 0; h < NH; ++h)
            if ((int)(blockIdx.y * NH + h) * 32 >= n_rows)
                for (int u = tid; u < kc * 8; u += ST) *reinterpret_cast<float4*>(As + (size_t)h * KC * 32 + u * 4) = z4;
        for (int si = nsv; si < spt; ++si)
            for (int u = tid; u < (kc * L) >> 2; u += ST) *reinterpret_cast<float4*>(Bs + (size_t)si * sstr + u * 4) = z4;
        if (coef) {
            for (int u = tid; u < kc; u += ST) {
                cf[u] = __ldg(p.coef_a + k0 + u);
                cf[KC + u] = __ldg(p.coef_b + k0 + u);
                cf[2 * KC + u] = __ldg(p.coef_c + k0 + u);
            }
        }
        if (k0 == 0) TLS(1);
        mbar_wait(&bar, phase);
        phase ^= 1u;
        if (k0 == 0) TLS(2);
        if (coef) __syncthreads();                // cf[] staged by all threads
        if (coef) {
            // dz = a*GV + b*Z + c in place (valid samples only: padding columns stay zero); the per-row coefficients
            // were staged in shared memory while the copies were in flight
            const int per = (kc * L) >> 2;        // float4 per sample
            for (int si = 0; si < nsv; ++si) {
                float* bs = Bs + (size_t)si * sstr;
                const float* bz = Bz + (size_t)si * sstr;
                for (int q = tid; q < per; q += ST) {
                    const int kk = (q << 2) >> lshift;
                    const float ca = cf[kk], cb = cf[KC + kk], cc = cf[2 * KC + kk];
                    float4* d = reinterpret_cast<float4*>(bs + q * 4);
                    const float4 g = *d, z = *reinterpret_cast<const float4*>(bz + q * 4);
                    *d = make_float4(fmaf(ca, g.x, fmaf(cb, z.x, cc)), fmaf(ca, g.y, fmaf(cb, z.y, cc)), fmaf(ca, g.z, fmaf(cb, z.z, cc)),
                                     fmaf(ca, g.w, fmaf(cb, z.w, cc)));
                }
            }
        }
        __syncthreads();
        if (k0 == 0) TLS(3);
        // ---- FFMA: MT x 4 micro-tile; register double buffering over blocks of KB reduction steps (the loads of
        //      block i+1 are issued before the FMAs of block i)
        const float* ap = As + (size_t)((ty * MT) >> 5) * KC * 32 + ((ty * MT) & 31);
        const float* bp = Bs + (size_t)((tx * 4) >> lshift) * sstr + ((tx * 4) & (L - 1));
        constexpr int KB = 4;
        float a0[KB][MT], a1[KB][MT];
        float4 b0v[KB], b1v[KB];
        auto load_blk = [&](float (&a)[KB][MT], float4 (&b)[KB], int kk) {
#pragma unroll
            for (int j = 0; j < KB; ++j) {
                if (MT == 4) {
                    const float4 t = *reinterpret_cast<const float4*>(ap + (kk + j) * 32);
                    a[j][0] = t.x; a[j][1 % MT] = t.y; a[j][2 % MT] = t.z; a[j][3 % MT] = t.w;
                } else {
                    const float2 t = *reinterpret_cast<const float2*>(ap + (kk + j) * 32);
                    a[j][0] = t.x; a[j][1 % MT] = t.y;
                }
                b[j] = *reinterpret_cast<const float4*>(bp + (kk + j) * L);
            }
        };
        auto fma_blk = [&](const float (&a)[KB][MT], const float4 (&b)[KB]) {
#pragma unroll
            for (int j = 0; j < KB; ++j)
#pragma unroll
                for (int i = 0; i < MT; ++i) {
                    acc[i][0] = fmaf(a[j][i], b[j].x, acc[i][0]);
                    acc[i][1] = fmaf(a[j][i], b[j].y, acc[i][1]);
                    acc[i][2] = fmaf(a[j][i], b[j].z, acc[i][2]);
                    acc[i][3] = fmaf(a[j][i], b[j].w, acc[i][3]);
                }
        };
        const int nblk = kc / KB;                 // kc % 4 == 0 (K % 4 == 0, M % 4 == 0)
        if (nblk > 0) load_blk(a0, b0v, 0);
        int blk = 0;
        for (; blk + 2 <= nblk; blk += 2) {
            load_blk(a1, b1v, (blk + 1) * KB);
            fma_blk(a0, b0v);
            if (blk + 2 < nblk) load_blk(a0, b0v, (blk + 2) * KB);
            fma_blk(a1, b1v);
        }
        if (blk < nblk) fma_blk(a0, b0v);
    }

    TLS(4);
    // ---- epilogue: thread owns rows r0 + ty*MT + i and the 4 columns n0 + tx*4 .. +3 (one sample: L % 4 == 0)
    const int n = n0 + tx * 4;
    const bool col_ok = n < N;
    const int b = col_ok ? n / L : 0, l0 = col_ok ? n - b * L : 0;
    if (MODE == FWD) {
        const int cnt = min(TN, N - n0);
#pragma unroll
        for (int i = 0; i < MT; ++i) {
            const int m = r0 + ty * MT + i;
            const bool row_ok = m < n_rows;
            const float4 z = make_float4(acc[i][0] + bias[i], acc[i][1] + bias[i], acc[i][2] + bias[i], acc[i][3] + bias[i]);
            if (row_ok && col_ok) *reinterpret_cast<float4*>(p.Z + ((long long)b * M + m) * L + l0) = z;
            if (p.bn_mode == 1) {
                float s = col_ok ? (z.x + z.y) + (z.z + z.w) : 0.f;
#pragma unroll
                for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                const float mean = s / (float)cnt;
                const float d0 = z.x - mean, d1 = z.y - mean, d2 = z.z - mean, d3 = z.w - mean;
                float q = col_ok ? (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3) : 0.f;
#pragma unroll
                for (int o = 4; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
                if (tx == 0 && row_ok) {
                    float* sp = p.stat_part + ((long long)blockIdx.x * M + m) * 2;
                    sp[0] = mean;
                    sp[1] = q;
                }
            }
        }
        if (p.bn_mode == 2) {  // eval: statistics come from the running buffers
            if (blockIdx.x == 0 && tid < TM) bn_eval_stats(p, r0 + tid, ldw);
            return;
        }
        if (p.bn_mode != 1) return;
        TLS(5);
        const bool lastb = last_block(p.counter + blockIdx.y, gridDim.x);
        TLS(6);
        if (!lastb) return;
        bn_finalize_rows(p, N, n_col_tiles, [=](int t) { return min(TN, N - t * TN); }, r0, TM, ldw);
        TLS(7);
    } else {
#pragma unroll
        for (int i = 0; i < MT; ++i) {
            const int k = r0 + ty * MT + i;
            if (k >= n_rows || !col_ok) continue;
            int s, kl;
            src_of(p, k, &s, &kl);
            float* dst = p.gsrc[s];
            if (!dst) continue;
            float4* d = reinterpret_cast<float4*>(dst + ((long long)b * p.src_C[s] + kl) * L + l0);
            float4 o = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
            if (p.gsrc_accum[s]) {
                const float4 c = *d;
                o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w;
            }
            *d = o;
        }
        TLS(5);
    }
}

// plain-fp32 weight images (fmt 1), tile major: [output-row tile of 32][reduction][32]; which 0 = FWD (rows m,
// reduction k), 1 = DGRAD (rows k, reduction m); the last tile is zero padded

// --------------------------------------------------------------------------------------------------
// WGRAD  dW[m,k] += sum_{b,l} dz[b,m,l] U[b,k,l],  dbias[m] += sum_{b,l} dz[b,m,l]       (red.add onto gW / gbias)
// 64 (m) x 32 (k) output tile per CTA; the reduction runs over chunks of 128/L samples, blockIdx.z strides over
// the chunks (split-K across the machine: 144 CTAs at the NTU batch) and the partial tiles are accumulated with
// vector red.add.  Both operands are (sample, row, l) in memory with l fastest, but the FFMA loop wants
// reduction-major rows, so the staging pass transposes: coalesced 128-bit loads (a row tile of one sample is one
// contiguous block), BatchNorm-backward fold on the fly, conflict-free scalar stores to [r = (sample, l)][row].
// --------------------------------------------------------------------------------------------------
constexpr int WTM = 64, WTK = 32, WR = 128;          // tile rows (m), tile cols (k), reduction elements per chunk
constexpr int WLDA = WTM + 4, WLDB = WTK + 4;

__global__ void __launch_bounds__(ST) k_sgw(const bmnas_conv_params p, const int lshift) {
    pdl_prologue();
    extern __shared__ __align__(16) float smem[];
    float* A2 = smem;                       // [WR][WLDA]
    float* B2 = A2 + WR * WLDA;             // [WR][WLDB]
    float* cf = B2 + WR * WLDB;             // coef_a | coef_b | coef_c of the 64 rows
    const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
    const int m0 = blockIdx.y * WTM, kt0 = blockIdx.x * WTK;
    const int K = p.K, M = p.M, L = p.L, ldw = p.w_fold * p.K;
    const int SB = WR >> lshift;            // samples per chunk
    const int q4 = L >> 2, q4s = lshift - 2;   // float4 per row
    const bool coef = p.coef_a != nullptr;
    if (coef) {
        for (int u = tid; u < WTM; u += ST) {
            const bool ok = m0 + u < M;
            cf[u] = ok ? __ldg(p.coef_a + m0 + u) : 0.f;
            cf[WTM + u] = ok ? __ldg(p.coef_b + m0 + u) : 0.f;
            cf[2 * WTM + u] = ok ? __ldg(p.coef_c + m0 + u) : 0.f;
        }
    }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float rowsum = 0.f;                      // bias gradient of row m0 + tid (threads < 64 of the k-tile-0 CTAs)
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const int n_chunks = (p.B + SB - 1) / SB;

    for (int ch = blockIdx.z; ch < n_chunks; ch += gridDim.z) {
        const int b0 = ch * SB, nsv = min(SB, p.B - b0);
        __syncthreads();                     // previous chunk's FFMA loop (and cf[]) done
        // ---- dz tile, transposed: unit u = (sample si, row m, l-group lq), lq fastest
        constexpr int UB = 8;
        const int unitsA = SB * WTM * q4;
        for (int u0 = tid; u0 < unitsA; u0 += ST * UB) {
            float4 g[UB], z[UB];
#pragma unroll
            for (int i = 0; i < UB; ++i) {
                const int u = u0 + i * ST;
                g[i] = z[i] = z4;
                if (u < unitsA) {
                    const int lq = u & (q4 - 1), m = (u >> q4s) & (WTM - 1), si = u >> (q4s + 6);
                    if (si < nsv && m0 + m < M) {
                        const long long idx = ((long long)(b0 + si) * M + m0 + m) * L + lq * 4;
                        g[i] = __ldg(reinterpret_cast<const float4*>(p.GV + idx));
                        if (coef) z[i] = __ldg(reinterpret_cast<const float4*>(p.Z + idx));
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < UB; ++i) {
                const int u = u0 + i * ST;
                if (u < unitsA) {
                    const int lq = u & (q4 - 1), m = (u >> q4s) & (WTM - 1), si = u >> (q4s + 6);
                    float4 v = g[i];
                    if (coef && si < nsv && m0 + m < M) {
                        const float a = cf[m], b = cf[WTM + m], c = cf[2 * WTM + m];
                        v = make_float4(fmaf(a, v.x, fmaf(b, z[i].x, c)), fmaf(a, v.y, fmaf(b, z[i].y, c)), fmaf(a, v.z, fmaf(b, z[i].z, c)),
                                        fmaf(a, v.w, fmaf(b, z[i].w, c)));
                    }
                    float* d = A2 + ((si << lshift) + lq * 4) * WLDA + m;
                    d[0] = v.x; d[WLDA] = v.y; d[2 * WLDA] = v.z; d[3 * WLDA] = v.w;
                }
            }
        }
        // ---- activation tile, transposed: unit u = (sample si, channel kk, l-group lq)
        const int unitsB = SB * WTK * q4;
        for (int u0 = tid; u0 < unitsB; u0 += ST * UB) {
            float4 g[UB];
#pragma unroll
            for (int i = 0; i < UB; ++i) {
                const int u = u0 + i * ST;
                g[i] = z4;
                if (u < unitsB) {
                    const int lq = u & (q4 - 1), kk = (u >> q4s) & (WTK - 1), si = u >> (q4s + 5);
                    if (si < nsv && kt0 + kk < K) {
                        int s, kl;
                        src_of(p, kt0 + kk, &s, &kl);
                        g[i] = __ldg(reinterpret_cast<const float4*>(p.src[s] + ((long long)(b0 + si) * p.src_C[s] + kl) * L + lq * 4));
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < UB; ++i) {
                const int u = u0 + i * ST;
                if (u < unitsB) {
                    const int lq = u & (q4 - 1), kk = (u >> q4s) & (WTK - 1), si = u >> (q4s + 5);
                    float* d = B2 + ((si << lshift) + lq * 4) * WLDB + kk;
                    d[0] = g[i].x; d[WLDB] = g[i].y; d[2 * WLDB] = g[i].z; d[3 * WLDB] = g[i].w;
                }
            }
        }
        __syncthreads();
        // ---- FFMA: 4 x 4 micro-tile, register double buffering (see k_sg)
        const float* ap = A2 + ty * 4;
        const float* bp = B2 + tx * 4;
        constexpr int KB = 4;
        float4 a0[KB], a1[KB], b0v[KB], b1v[KB];
        auto load_blk = [&](float4 (&a)[KB], float4 (&b)[KB], int r) {
#pragma unroll
            for (int j = 0; j < KB; ++j) {
                a[j] = *reinterpret_cast<const float4*>(ap + (r + j) * WLDA);
                b[j] = *reinterpret_cast<const float4*>(bp + (r + j) * WLDB);
            }
        };
        auto fma_blk = [&](const float4 (&a)[KB], const float4 (&b)[KB]) {
#pragma unroll
            for (int j = 0; j < KB; ++j) {
                const float av[4] = {a[j].x, a[j].y, a[j].z, a[j].w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    acc[i][0] = fmaf(av[i], b[j].x, acc[i][0]);
                    acc[i][1] = fmaf(av[i], b[j].y, acc[i][1]);
                    acc[i][2] = fmaf(av[i], b[j].z, acc[i][2]);
                    acc[i][3] = fmaf(av[i], b[j].w, acc[i][3]);
                }
            }
        };
        constexpr int NBLK = WR / KB;
        load_blk(a0, b0v, 0);
#pragma unroll 1
        for (int blk = 0; blk + 2 <= NBLK; blk += 2) {
            load_blk(a1, b1v, (blk + 1) * KB);
            fma_blk(a0, b0v);
            if (blk + 2 < NBLK) load_blk(a0, b0v, (blk + 2) * KB);
            fma_blk(a1, b1v);
        }
        if (blockIdx.x == 0 && tid < WTM) {   // bias gradient = row sums of dz
            float s_ = 0.f;
            for (int r = 0; r < WR; ++r) s_ += A2[r * WLDA + tid];
            rowsum += s_;
        }
    }

    // ---- epilogue: vector red.add of the partial tile (both halves of a folded weight)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i, k = kt0 + tx * 4;
        if (m >= M || k >= K) continue;
        int sg_, ml;
        w_row(p, m, ldw, &sg_, &ml);
        if (!p.gW[sg_]) continue;
        float* row = p.gW[sg_] + (long long)ml * ldw + k;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(row), "f"(acc[i][0]), "f"(acc[i][1]), "f"(acc[i][2]), "f"(acc[i][3]) : "memory");
        if (p.w_fold == 2)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(row + K), "f"(acc[i][0]), "f"(acc[i][1]), "f"(acc[i][2]), "f"(acc[i][3]) : "memory");
    }
    if (blockIdx.x == 0 && tid < WTM && m0 + tid < M) {
        int sg_, ml;
        w_row(p, m0 + tid, ldw, &sg_, &ml);
        if (p.gbias[sg_]) atomicAdd(p.gbias[sg_] + ml, rowsum);
    }
}

template <int MODE, int TM>
static int launch_sg(const bmnas_conv_params* p, cudaStream_t stream) {
    const int N = p->B * p->L;
    const int R = MODE == FWD ? p->K : p->M;
    const int rows = MODE == FWD ? p->M : p->K;
    const int KC = R < KCS ? R : KCS;
    const bool coef = MODE == DGRAD && p->coef_a != nullptr;
    const size_t smem = smem_floats<MODE, TM>(KC, p->L, coef) * sizeof(float);
    static size_t configured = 0;
    if (smem > 40 * 1024 && smem > configured) {   // the 48 KB default counts static + dynamic: opt in with a margin
        if (cudaFuncSetAttribute(k_sg<MODE, TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return BMNAS_ELAUNCH;
        configured = smem;
    }
    dim3 grid((N + TN - 1) / TN, (rows + TM - 1) / TM);
    int lshift = 0;
    while ((1 << lshift) < p->L) ++lshift;
    launch_k(k_sg<MODE, TM>, grid, ST, smem, stream, *p, N, (int)grid.x, KC, lshift);
    BMNAS_LAUNCH_CHECK();
    return BMNAS_OK;
}

static bool al16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }

}  // namespace sg

// shapes the cp.async kernels can take: whole 16-byte groups everywhere
bool sg_eligible(const bmnas_conv_params* p, int mode) {
    using namespace sg;
    if ((p->L & 3) || (p->K & 3) || (p->M & 3)) return false;
    if ((p->L & (p->L - 1)) || p->L > TN) return false;     // a column tile is TN / L whole samples
    for (int i = 0; i < p->n_src; ++i)
        if ((p->src_C[i] & 3) || (mode == FWD && !al16(p->src[i])) || (mode == DGRAD && p->gsrc[i] && !al16(p->gsrc[i]))) return false;
    if (mode == FWD && (!al16(p->Z) || !al16(p->wimg_fwd))) return false;
    if (mode == DGRAD && (!al16(p->GV) || (p->coef_a && !al16(p->Z)) || !al16(p->wimg_dgrad))) return false;
    return true;
}

int sg_conv_fwd(const bmnas_conv_params* p, cudaStream_t stream) {
    using namespace sg;
    const int N = p->B * p->L;
    const long long t64 = (long long)((p->M + 63) / 64) * ((N + TN - 1) / TN);
    return t64 >= 120 ? launch_sg<FWD, 64>(p, stream) : launch_sg<FWD, 32>(p, stream);
}

// wgrad: vector red.add needs K % 4 == 0 and 16-byte aligned weight-gradient rows
bool sgw_eligible(const bmnas_conv_params* p) {
    using namespace sg;
    if ((p->L & 3) || (p->K & 3) || (p->L & (p->L - 1)) || p->L > 32) return false;
    for (int i = 0; i < p->n_src; ++i)
        if ((p->src_C[i] & 3) || !al16(p->src[i])) return false;
    for (int i = 0; i < p->n_seg; ++i)
        if (p->gW[i] && !al16(p->gW[i])) return false;
    return al16(p->GV) && (!p->coef_a || al16(p->Z));
}

int sg_conv_wgrad(const bmnas_conv_params* p, cudaStream_t stream) {
    using namespace sg;
    int lshift = 0;
    while ((1 << lshift) < p->L) ++lshift;
    const int SB = WR >> lshift, n_chunks = (p->B + SB - 1) / SB;
    const int tiles = ((p->K + WTK - 1) / WTK) * ((p->M + WTM - 1) / WTM);
    int zs = (kNumSMs + tiles - 1) / tiles;
    if (p->splits > 0) zs = p->splits;
    if (zs > n_chunks) zs = n_chunks;
    if (zs < 1) zs = 1;
    const size_t smem = (size_t)(WR * (WLDA + WLDB) + 3 * WTM) * sizeof(float);
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(k_sgw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return BMNAS_ELAUNCH;
        configured = true;
    }
    dim3 grid((p->K + WTK - 1) / WTK, (p->M + WTM - 1) / WTM, zs);
    launch_k(k_sgw, grid, ST, smem, stream, *p, lshift);
    BMNAS_LAUNCH_CHECK();
    return BMNAS_OK;
}

int sg_conv_dgrad(const bmnas_conv_params* p, cudaStream_t stream) {
    using namespace sg;
    const int N = p->B * p->L;
    const long long t64 = (long long)((p->K + 63) / 64) * ((N + TN - 1) / TN);
    return t64 >= 120 ? launch_sg<DGRAD, 64>(p, stream) : launch_sg<DGRAD, 32>(p, stream);
}

}  // namespace bmnas
