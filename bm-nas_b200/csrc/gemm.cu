// fp32 (FFMA) GEMM family for the 1x1 convolutions of the fusion cell:
//   conv_fwd   Z[b,m,l]  = sum_k Weff[m,k] U[b,k,l] + bias[m]      (+ BN batch statistics)
//   conv_dgrad dU[b,k,l] = sum_m Weff[m,k] dz[b,m,l]
//   conv_wgrad dW[m,k]  += sum_{b,l} dz[b,m,l] U[b,k,l],  dbias[m] += sum dz
// U is a virtual channel concat (never materialised), the output rows are stacked
// weight segments, w_fold folds cat([t,t]); dz is produced on the fly from
// (GV, Z, coef) = BatchNorm backward fused into the operand load.
// This is the fp32-exact path (parity 1e-5 vs the reference on CPU).
//
// Shape of the problem on this path: reductions are short (C..3C = 128..768) and at the
// reference batch (B*L = 768 columns) the whole GEMM is ~0.1 GFLOP, i.e. latency bound.
// So: 32x32 output tiles (hundreds of CTAs -> every SM busy), the ENTIRE reduction
// extent of both operands staged in shared memory with one burst of independent
// 128-bit loads (one DRAM/L2 round trip, up to 384 reduction rows per pass), then a
// dependency-free FFMA loop (4x2 register micro-tile, 128 threads).
#include <cstdlib>
#include "common.cuh"
#include "gemm_shared.cuh"

namespace bmnas {

constexpr int TM = 32, TN = 32, GT = 256;
constexpr int LDA = TM + 4, LDB = TN + 4;      // padded rows keep 16-byte alignment and spread banks
constexpr int KC_MAX = 384;                    // reduction rows staged per pass (2 CTAs/SM at 110 KB)

__host__ __device__ inline int round_up(int a, int b) { return (a + b - 1) / b * b; }
__host__ __device__ inline size_t gemm_smem_bytes(int kc) { return (size_t)kc * (LDA + LDB) * sizeof(float); }

// acc[2][2] += A[kk][ty*2 .. +1] (x) B[kk][tx*2 .. +1] over kk < kc   (16 x 16 threads cover 32 x 32)
constexpr int MR = 2;
__device__ __forceinline__ void mma_tile(const float* As, const float* Bs, int kc, float (&acc)[MR][2], int ty, int tx) {
#pragma unroll 8
    for (int kk = 0; kk < kc; ++kk) {
        const float2 a = *reinterpret_cast<const float2*>(As + kk * LDA + ty * 2);
        const float2 b = *reinterpret_cast<const float2*>(Bs + kk * LDB + tx * 2);
        acc[0][0] = fmaf(a.x, b.x, acc[0][0]); acc[0][1] = fmaf(a.x, b.y, acc[0][1]);
        acc[1][0] = fmaf(a.y, b.x, acc[1][0]); acc[1][1] = fmaf(a.y, b.y, acc[1][1]);
    }
}

// issue U independent loads per thread before the first dependent store (memory-level parallelism:
// the staging phase is ONE latency round trip per U*GT elements instead of one per element)
template <int U, class T, class Load, class Store>
__device__ __forceinline__ void batched(int total, T zero, Load ld, Store st) {
    for (int base = 0; base < total; base += GT * U) {
        T v[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const int u = base + j * GT + threadIdx.x;
            v[j] = u < total ? ld(u) : zero;
        }
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const int u = base + j * GT + threadIdx.x;
            if (u < total) st(u, v[j]);
        }
    }
}

// both operands of a tile staged in ONE pass: U loads of A and U loads of B are in flight per thread
// before the first dependent shared-memory store
template <int U, class LA, class SA, class LB, class SB>
__device__ __forceinline__ void stage2(int totalA, LA ldA, SA stA, int totalB, LB ldB, SB stB) {
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int base = 0; base < totalA || base < totalB; base += GT * U) {
        float4 va[U], vb[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const int u = base + j * GT + threadIdx.x;
            va[j] = u < totalA ? ldA(u) : z4;
        }
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const int u = base + j * GT + threadIdx.x;
            vb[j] = u < totalB ? ldB(u) : z4;
        }
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const int u = base + j * GT + threadIdx.x;
            if (u < totalA) stA(u, va[j]);
        }
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const int u = base + j * GT + threadIdx.x;
            if (u < totalB) stB(u, vb[j]);
        }
    }
}

// ------------------------------------------------------------------ forward
template <bool VEC>
__global__ void __launch_bounds__(GT, 1) k_conv_fwd(const bmnas_conv_params p, const int N, const int n_col_tiles,
                                                     const int KC) {
    pdl_prologue();
    extern __shared__ __align__(16) float smem[];
    float* As = smem;
    float* Bs = smem + (size_t)KC * LDA;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    const int K = p.K, M = p.M, L = p.L, ldw = p.w_fold * p.K;
    float acc[MR][2];
#pragma unroll
    for (int i = 0; i < MR; ++i) acc[i][0] = acc[i][1] = 0.f;

    for (int k0 = 0; k0 < K; k0 += KC) {
        const int kc = min(KC, round_up(K - k0, 4));
        if (k0) __syncthreads();
        // A: As[k][m] = Weff[m0+m][k0+k]; lanes walk m (conflict-free transposed stores), float4 along k
        // B: Bs[k][col] = U[k0+k][n0+col]: 8 float4 per staged row
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (VEC) {
            stage2<6>((kc >> 2) * TM,
                      [&](int u) {
                          const int m = u & 31, k4 = (u >> 5) << 2;
                          float4 v = z4;
                          if (m0 + m < M && k0 + k4 < K) {
                              const float* r = w_row(p, m0 + m, ldw, nullptr, nullptr) + k0 + k4;
                              v = __ldg(reinterpret_cast<const float4*>(r));
                              if (p.w_fold == 2) {
                                  const float4 w = __ldg(reinterpret_cast<const float4*>(r + K));
                                  v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
                              }
                          }
                          return v;
                      },
                      [&](int u, float4 v) {
                          const int m = u & 31, k4 = (u >> 5) << 2;
                          As[(k4 + 0) * LDA + m] = v.x; As[(k4 + 1) * LDA + m] = v.y;
                          As[(k4 + 2) * LDA + m] = v.z; As[(k4 + 3) * LDA + m] = v.w;
                      },
                      kc * 8,
                      [&](int u) {
                          const int n = n0 + (u & 7) * 4, k = k0 + (u >> 3);
                          if (n >= N || k >= K) return z4;
                          int s, kl;
                          src_of(p, k, &s, &kl);
                          return __ldg(reinterpret_cast<const float4*>(p.src[s] + ((long long)(n / L) * p.src_C[s] + kl) * L + (n % L)));
                      },
                      [&](int u, float4 v) { *reinterpret_cast<float4*>(Bs + (u >> 3) * LDB + (u & 7) * 4) = v; });
        } else {
            batched<8>(kc * TM, 0.f,
                       [&](int u) {
                           const int m = u & 31, k = u >> 5;
                           float v = 0.f;
                           if (m0 + m < M && k0 + k < K) {
                               const float* r = w_row(p, m0 + m, ldw, nullptr, nullptr) + k0 + k;
                               v = __ldg(r);
                               if (p.w_fold == 2) v += __ldg(r + K);
                           }
                           return v;
                       },
                       [&](int u, float v) { As[(u >> 5) * LDA + (u & 31)] = v; });
            batched<8>(kc * TN, 0.f,
                       [&](int u) {
                           const int n = n0 + (u & 31), k = k0 + (u >> 5);
                           if (n >= N || k >= K) return 0.f;
                           int s, kl;
                           src_of(p, k, &s, &kl);
                           return __ldg(p.src[s] + ((long long)(n / L) * p.src_C[s] + kl) * L + (n % L));
                       },
                       [&](int u, float v) { Bs[(u >> 5) * LDB + (u & 31)] = v; });
        }
        __syncthreads();
        mma_tile(As, Bs, kc, acc, ty, tx);
    }

    // ---- epilogue: bias, store, per-tile BN statistics
    const int cnt = min(TN, N - n0);
    const int nb = n0 + tx * 2;
    const bool vec = ((L & 1) == 0) && (nb + 1 < N);
#pragma unroll
    for (int i = 0; i < MR; ++i) {
        const int m = m0 + ty * MR + i;
        float bias = 0.f;
        if (m < M) {
            int s, ml;
            w_row(p, m, ldw, &s, &ml);
            if (p.bias[s]) bias = __ldg(p.bias[s] + ml);
        }
        const float z0 = acc[i][0] + bias, z1 = acc[i][1] + bias;
        if (m < M) {
            if (vec) {
                *reinterpret_cast<float2*>(p.Z + ((long long)(nb / L) * M + m) * L + (nb % L)) = make_float2(z0, z1);
            } else {
                if (nb < N) p.Z[((long long)(nb / L) * M + m) * L + (nb % L)] = z0;
                if (nb + 1 < N) p.Z[((long long)((nb + 1) / L) * M + m) * L + ((nb + 1) % L)] = z1;
            }
        }
        if (p.bn_mode == 1) {
            float s = (nb < N ? z0 : 0.f) + (nb + 1 < N ? z1 : 0.f);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            const float mean = s / (float)cnt;
            const float d0 = z0 - mean, d1 = z1 - mean;
            float d2 = (nb < N ? d0 * d0 : 0.f) + (nb + 1 < N ? d1 * d1 : 0.f);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) d2 += __shfl_xor_sync(0xffffffffu, d2, o);
            if (tx == 0 && m < M) {
                float* q = p.stat_part + ((long long)blockIdx.x * M + m) * 2;
                q[0] = mean;
                q[1] = d2;
            }
        }
    }

    if (p.bn_mode == 2) {  // eval: statistics come from the running buffers
        if (blockIdx.x == 0 && tid < TM) bn_eval_stats(p, m0 + tid, ldw);
        return;
    }
    if (p.bn_mode != 1) return;
    // ---- last CTA of this row-tile merges the per-column-tile statistics (Chan), fixed order
    if (!last_block(p.counter + blockIdx.y, gridDim.x)) return;
    bn_finalize_rows(p, N, n_col_tiles, [=](int t) { return min(TN, N - t * TN); }, m0, TM, ldw);
}

// ------------------------------------------------------------------ dgrad
template <bool VEC>
__global__ void __launch_bounds__(GT, 1) k_conv_dgrad(const bmnas_conv_params p, const int N, const int KC) {
    pdl_prologue();
    extern __shared__ __align__(16) float smem[];
    float* As = smem;
    float* Bs = smem + (size_t)KC * LDA;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int kt0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    const int K = p.K, M = p.M, L = p.L, ldw = p.w_fold * p.K;

    float acc[MR][2];
#pragma unroll
    for (int i = 0; i < MR; ++i) acc[i][0] = acc[i][1] = 0.f;

    for (int mk0 = 0; mk0 < M; mk0 += KC) {
        const int kc = min(KC, round_up(M - mk0, 4));
        if (mk0) __syncthreads();
        // A: As[mm][k] = Weff[mk0+mm][kt0+k]: straight (coalesced) copies of weight-row slices
        // B: Bs[mm][col] = dz[mk0+mm][n0+col] (BatchNorm backward folded into the load)
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (VEC) {
            stage2<6>(kc * 8,
                      [&](int u) {
                          const int k4 = (u & 7) << 2, mm = u >> 3;
                          float4 v = z4;
                          if (mk0 + mm < M && kt0 + k4 < K) {
                              const float* r = w_row(p, mk0 + mm, ldw, nullptr, nullptr) + kt0 + k4;
                              v = __ldg(reinterpret_cast<const float4*>(r));
                              if (p.w_fold == 2) {
                                  const float4 w = __ldg(reinterpret_cast<const float4*>(r + K));
                                  v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
                              }
                          }
                          return v;
                      },
                      [&](int u, float4 v) { *reinterpret_cast<float4*>(As + (u >> 3) * LDA + ((u & 7) << 2)) = v; },
                      kc * 8,
                      [&](int u) {
                          const int n = n0 + (u & 7) * 4, m = mk0 + (u >> 3);
                          if (n >= N || m >= M) return z4;
                          return dz4(p, ((long long)(n / L) * M + m) * L + (n % L), m);
                      },
                      [&](int u, float4 v) { *reinterpret_cast<float4*>(Bs + (u >> 3) * LDB + (u & 7) * 4) = v; });
        } else {
            batched<8>(kc * TM, 0.f,
                       [&](int u) {
                           const int k = u & 31, mm = u >> 5;
                           float v = 0.f;
                           if (mk0 + mm < M && kt0 + k < K) {
                               const float* r = w_row(p, mk0 + mm, ldw, nullptr, nullptr) + kt0 + k;
                               v = __ldg(r);
                               if (p.w_fold == 2) v += __ldg(r + K);
                           }
                           return v;
                       },
                       [&](int u, float v) { As[(u >> 5) * LDA + (u & 31)] = v; });
            batched<8>(kc * TN, 0.f,
                       [&](int u) {
                           const int n = n0 + (u & 31), m = mk0 + (u >> 5);
                           if (n >= N || m >= M) return 0.f;
                           return dz1(p, ((long long)(n / L) * M + m) * L + (n % L), m);
                       },
                       [&](int u, float v) { Bs[(u >> 5) * LDB + (u & 31)] = v; });
        }
        __syncthreads();
        mma_tile(As, Bs, kc, acc, ty, tx);
    }

    const int nb = n0 + tx * 2;
    const bool vec = ((L & 1) == 0) && (nb + 1 < N);
#pragma unroll
    for (int i = 0; i < MR; ++i) {
        const int k = kt0 + ty * MR + i;
        if (k >= K) continue;
        int s, kl;
        src_of(p, k, &s, &kl);
        float* dst = p.gsrc[s];
        if (!dst) continue;
        const bool accum = p.gsrc_accum[s] != 0;
        if (vec) {
            float2* d = reinterpret_cast<float2*>(dst + ((long long)(nb / L) * p.src_C[s] + kl) * L + (nb % L));
            float2 o = make_float2(acc[i][0], acc[i][1]);
            if (accum) {
                const float2 c = *d;
                o.x += c.x; o.y += c.y;
            }
            *d = o;
        } else {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int n = nb + j;
                if (n < N) {
                    float* d = dst + ((long long)(n / L) * p.src_C[s] + kl) * L + (n % L);
                    *d = accum ? (*d + acc[i][j]) : acc[i][j];
                }
            }
        }
    }
}

// ------------------------------------------------------------------ wgrad
template <bool VEC>
__global__ void __launch_bounds__(GT, 1) k_conv_wgrad(const bmnas_conv_params p, const int N, const int chunkN,
                                                       const int KC) {
    pdl_prologue();
    extern __shared__ __align__(16) float smem[];
    float* As = smem;
    float* Bs = smem + (size_t)KC * LDA;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int kt0 = blockIdx.x * TN, m0 = blockIdx.y * TM;
    const int K = p.K, M = p.M, L = p.L, ldw = p.w_fold * p.K;
    const int nbeg = blockIdx.z * chunkN, nend = min(N, nbeg + chunkN);
    // vector staging needs whole samples per chunk: the host makes chunkN and KC multiples of L when VEC
    const bool vec = VEC;

    float acc[MR][2], rs[MR] = {0.f, 0.f};
#pragma unroll
    for (int i = 0; i < MR; ++i) acc[i][0] = acc[i][1] = 0.f;

    for (int nk0 = nbeg; nk0 < nend; nk0 += KC) {
        const int kc = min(KC, round_up(nend - nk0, vec ? L : 4));
        if (nk0 != nbeg) __syncthreads();
        if (VEC) {
            const int q = L >> 2;   // float4 per (sample,row); lanes: l4 fastest, then row
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
            auto st_t = [&](float* S, int ld, int u, float4 v) {
                const int l4 = u % q, r = (u / q) & 31, sm = u / (q * TM);
                const int nn = sm * L + l4 * 4;
                S[(nn + 0) * ld + r] = v.x; S[(nn + 1) * ld + r] = v.y;
                S[(nn + 2) * ld + r] = v.z; S[(nn + 3) * ld + r] = v.w;
            };
            stage2<6>((kc >> 2) * TM,
                      [&](int u) {
                          const int l4 = u % q, r = (u / q) & 31, sm = u / (q * TM);
                          const int n = nk0 + sm * L + l4 * 4, m = m0 + r;
                          if (n >= nend || m >= M) return z4;
                          return dz4(p, ((long long)(n / L) * M + m) * L + (n % L), m);
                      },
                      [&](int u, float4 v) { st_t(As, LDA, u, v); },
                      (kc >> 2) * TM,
                      [&](int u) {
                          const int l4 = u % q, r = (u / q) & 31, sm = u / (q * TM);
                          const int n = nk0 + sm * L + l4 * 4, k = kt0 + r;
                          if (n >= nend || k >= K) return z4;
                          int s, kl;
                          src_of(p, k, &s, &kl);
                          return __ldg(reinterpret_cast<const float4*>(p.src[s] + ((long long)(n / L) * p.src_C[s] + kl) * L + (n % L)));
                      },
                      [&](int u, float4 v) { st_t(Bs, LDB, u, v); });
        } else {
            batched<8>(kc * TM, 0.f,
                       [&](int u) {
                           const int nn = u % kc, r = u / kc;
                           const int n = nk0 + nn, m = m0 + r;
                           if (n >= nend || m >= M) return 0.f;
                           return dz1(p, ((long long)(n / L) * M + m) * L + (n % L), m);
                       },
                       [&](int u, float v) { As[(u % kc) * LDA + (u / kc)] = v; });
            batched<8>(kc * TM, 0.f,
                       [&](int u) {
                           const int nn = u % kc, r = u / kc;
                           const int n = nk0 + nn, k = kt0 + r;
                           if (n >= nend || k >= K) return 0.f;
                           int s, kl;
                           src_of(p, k, &s, &kl);
                           return __ldg(p.src[s] + ((long long)(n / L) * p.src_C[s] + kl) * L + (n % L));
                       },
                       [&](int u, float v) { Bs[(u % kc) * LDB + (u / kc)] = v; });
        }
        __syncthreads();
        mma_tile(As, Bs, kc, acc, ty, tx);
        if (blockIdx.x == 0 && tx == 0) {  // bias gradient: row sums of dz
            for (int kk = 0; kk < kc; ++kk) {
                const float2 a = *reinterpret_cast<const float2*>(As + kk * LDA + ty * MR);
                rs[0] += a.x; rs[1] += a.y;
            }
        }
    }

#pragma unroll
    for (int i = 0; i < MR; ++i) {
        const int m = m0 + ty * MR + i;
        if (m >= M) continue;
        int s, ml;
        w_row(p, m, ldw, &s, &ml);
        if (p.gW[s]) {
            float* row = p.gW[s] + (long long)ml * ldw;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int k = kt0 + tx * 2 + j;
                if (k < K) {
                    for (int f = 0; f < p.w_fold; ++f) atomicAdd(row + f * K + k, acc[i][j]);
                }
            }
        }
        if (blockIdx.x == 0 && tx == 0 && p.gbias[s]) atomicAdd(p.gbias[s] + ml, rs[i]);
    }
}

static int conv_check(const bmnas_conv_params* p) {
    if (!p || p->B < 1 || p->L < 1 || p->K < 1 || p->M < 1) return BMNAS_EINVAL;
    if (p->n_src < 1 || p->n_src > BMNAS_MAX_SRC || p->n_seg < 1 || p->n_seg > BMNAS_MAX_SEG) return BMNAS_EINVAL;
    if (p->w_fold != 1 && p->w_fold != 2) return BMNAS_EINVAL;
    int k = 0, m = 0;
    for (int i = 0; i < p->n_src; ++i) k += p->src_C[i];
    for (int i = 0; i < p->n_seg; ++i) m += p->seg_M[i];
    if (k != p->K || m != p->M) return BMNAS_EINVAL;
    if ((long long)p->B * p->L > 0x7fffffffLL) return BMNAS_EINVAL;
    return BMNAS_OK;
}

template <class Kern>
static int set_smem(Kern kern, size_t bytes, size_t* configured) {
    if (bytes > 40 * 1024 && bytes > *configured) {   // the 48 KB default counts static + dynamic: opt in with a margin
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess)
            return BMNAS_ELAUNCH;
        *configured = bytes;
    }
    return BMNAS_OK;
}

}  // namespace bmnas

using namespace bmnas;

namespace bmnas {
bool tc_eligible(const bmnas_conv_params* p, int mode);
int tc_conv_fwd(const bmnas_conv_params* p, int x3, cudaStream_t stream);
int tc_conv_dgrad(const bmnas_conv_params* p, int x3, cudaStream_t stream);
int tc_conv_wgrad(const bmnas_conv_params* p, int x3, cudaStream_t stream);
bool sg_eligible(const bmnas_conv_params* p, int mode);
int sg_conv_fwd(const bmnas_conv_params* p, cudaStream_t stream);
int sg_conv_dgrad(const bmnas_conv_params* p, cudaStream_t stream);
bool sgw_eligible(const bmnas_conv_params* p);
int sg_conv_wgrad(const bmnas_conv_params* p, cudaStream_t stream);
}  // namespace bmnas

// GEMM engine: 0 = fp32 FFMA tiles, 1 = tcgen05 3xTF32 (fp32-class accuracy), 2 = tcgen05 1xTF32 (reduced precision),
// 3 = as 1, but the fused NodeMixedOp forward (mixed_tc.cu) stages bf16 operands (kind::f16, fp32 accumulation)
int bmnas_gemm_mode_flag = 1;
static inline bool gemm_x3() { return bmnas_gemm_mode_flag == 1 || bmnas_gemm_mode_flag == 3; }
extern "C" int bmnas_set_gemm_mode(int mode) {
    if (mode < 0 || mode > 3) return BMNAS_EINVAL;
    bmnas_gemm_mode_flag = mode;
    return BMNAS_OK;
}
extern "C" int bmnas_get_gemm_mode(void) { return bmnas_gemm_mode_flag; }

// Which weight-image format (= which GEMM engine) serves a conv over B*L columns best.  Up to a few thousand
// columns the problem is latency bound and the cp.async fp32 kernels win (see gemm_sg.cu); beyond that the
// tcgen05 panel kernel amortises its per-instruction cost.  BMNAS_GEMM_MODE=0 keeps everything on fp32 FFMA.
extern "C" int bmnas_conv_image_fmt(int B, int L, int K, int M) {
    if ((L & 3) || (K & 3) || (M & 3)) return -1;
    const long long N = (long long)B * L;
    const bool sg_ok = (L & (L - 1)) == 0 && L <= 32;      // gemm_sg.cu: a 32-column tile is 32 / L whole samples
    if (bmnas_gemm_mode_flag == 0) return sg_ok ? 1 : -1;
    // by the GEMM's work, not by its column count alone: the small-N cp.async FFMA engine wins while the problem is
    // latency bound (NTU B=96: 768 x 384 x 128 = 38 M MACs), the tensor cores once there is arithmetic to amortise
    // their pipeline -- Ego-large (C=256, L=16, B=96: 1536 columns but 768 x 256 weights, 302 M MACs) belongs there.
    // BMNAS_TC_MIN_MACS overrides the crossover.  Measured with the warp-specialised kernels (profiles/r02_engine_crossover.txt,
    // NTU node conv): forward 20.3 (FFMA) vs 19.3 us (tcgen05) at B = 256 = 100 M MACs, 36.4 vs 21.0 at B = 512.
    static long long min_macs = -1;
    if (min_macs < 0) {
        const char* e = getenv("BMNAS_TC_MIN_MACS");
        min_macs = e ? atoll(e) : 100000000LL;
    }
    return (sg_ok && N * (long long)M * K <= min_macs) ? 1 : 0;
}

// ... and the same question for the dgrad GEMM of that conv (rows K, reduction M): its crossover is lower than the forward's
// -- no BatchNorm finalize at the end, and the BatchNorm-backward fold rides in the staging pass.  Measured, NTU node conv
// (M = 384), us FFMA vs tcgen05: 12.4 vs 12.7 at B = 192 (75 M MACs), 18.6 vs 12.9 at B = 384
// (profiles/r02_engine_crossover.txt).  A short reduction (out_conv's dgrad, M = 128) is faster on the tensor-core kernel even
// at B = 96 stand-alone (5.3 vs 7.4 us), but inside the captured step the FFMA kernel's weight prefetch ahead of
// griddepcontrol.wait and its small footprint (it shares SMs with its predecessor) win the difference back: the step time did
// not move (0.554 vs 0.558 ms), so small problems stay on the FFMA engine.  BMNAS_TC_MIN_MACS_D overrides the threshold.
extern "C" int bmnas_conv_image_fmt_dgrad(int B, int L, int K, int M) {
    const int f = bmnas_conv_image_fmt(B, L, K, M);
    if (f != 1) return f;                      // not eligible for the FFMA engine (or already on the tensor cores)
    if (bmnas_gemm_mode_flag == 0) return f;
    static long long min_macs_d = -1;
    if (min_macs_d < 0) {
        const char* e = getenv("BMNAS_TC_MIN_MACS_D");
        min_macs_d = e ? atoll(e) : 75000000LL;
    }
    return (long long)B * L * M * K > min_macs_d ? 0 : 1;
}

extern "C" long long bmnas_conv_stat_part_size(const bmnas_conv_params* p) {
    const long long N = (long long)p->B * p->L;
    return ((N + TN - 1) / TN) * p->M * 2;
}
extern "C" int bmnas_conv_num_counters(const bmnas_conv_params* p) { return (p->M + TM - 1) / TM; }

static bool gal16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }

// 128-bit staging: L % 4 == 0 and K % 4 == 0 (rows of every operand are whole float4s) and 16-byte alignment
static bool conv_vec_ok(const bmnas_conv_params* p, bool need_gv) {
    if ((p->L & 3) || (p->K & 3)) return false;
    for (int i = 0; i < p->n_src; ++i)
        if (!gal16(p->src[i]) || (p->src_C[i] & 3 && p->n_src > 1 && (p->L & 3))) return false;
    for (int i = 0; i < p->n_seg; ++i)
        if (!gal16(p->W[i])) return false;
    if (need_gv && (!gal16(p->GV) || !gal16(p->Z))) return false;
    return true;
}

extern "C" int bmnas_conv_fwd(const bmnas_conv_params* p, void* stream) {
    int e = conv_check(p);
    if (e) return e;
    if (!p->Z) return BMNAS_EINVAL;
    for (int i = 0; i < p->n_src; ++i)
        if (!p->src[i]) return BMNAS_EINVAL;
    for (int i = 0; i < p->n_seg; ++i)
        if (!p->W[i]) return BMNAS_EINVAL;
    if (p->bn_mode == 1 && (!p->stat_part || !p->counter || !p->mean || !p->rstd)) return BMNAS_EINVAL;
    if (p->bn_mode == 2) {
        if (!p->mean || !p->rstd) return BMNAS_EINVAL;
        for (int i = 0; i < p->n_seg; ++i)
            if (!p->running_mean[i] || !p->running_var[i]) return BMNAS_EINVAL;
    }
    BMNAS_DRY_RETURN();
    bmnas_conv_params q_;
    if (p->wimg_fwd && p->wimg_fmt == 1) {
        if (sg_eligible(p, 0)) return sg_conv_fwd(p, (cudaStream_t)stream);
        q_ = *p;                       // unaligned call-time tensor: the generic kernels below stage W themselves
        q_.wimg_fwd = q_.wimg_dgrad = nullptr;
        p = &q_;
    }
    if (bmnas_gemm_mode_flag && tc_eligible(p, 0)) return tc_conv_fwd(p, gemm_x3(), (cudaStream_t)stream);
    const int N = p->B * p->L;
    const int KC = min(KC_MAX, round_up(p->K, 4));
    const bool vec = conv_vec_ok(p, false);
    static size_t configured[2] = {0, 0};
    e = vec ? set_smem(k_conv_fwd<true>, gemm_smem_bytes(KC), &configured[1])
            : set_smem(k_conv_fwd<false>, gemm_smem_bytes(KC), &configured[0]);
    if (e) return e;
    dim3 grid((N + TN - 1) / TN, (p->M + TM - 1) / TM);
    if (vec)
        launch_k(k_conv_fwd<true>, grid, GT, gemm_smem_bytes(KC), (cudaStream_t)stream, *p, N, (int)grid.x, KC);
    else
        launch_k(k_conv_fwd<false>, grid, GT, gemm_smem_bytes(KC), (cudaStream_t)stream, *p, N, (int)grid.x, KC);
    BMNAS_LAUNCH_CHECK();
    return BMNAS_OK;
}

extern "C" int bmnas_conv_dgrad(const bmnas_conv_params* p, void* stream) {
    int e = conv_check(p);
    if (e) return e;
    if (!p->GV) return BMNAS_EINVAL;
    if (p->coef_a && (!p->coef_b || !p->coef_c || !p->Z)) return BMNAS_EINVAL;
    for (int i = 0; i < p->n_seg; ++i)
        if (!p->W[i]) return BMNAS_EINVAL;
    BMNAS_DRY_RETURN();
    bmnas_conv_params q_;
    if (p->wimg_dgrad && p->wimg_fmt == 1) {
        if (sg_eligible(p, 1)) return sg_conv_dgrad(p, (cudaStream_t)stream);
        q_ = *p;
        q_.wimg_fwd = q_.wimg_dgrad = nullptr;
        p = &q_;
    }
    if (bmnas_gemm_mode_flag && tc_eligible(p, 1)) return tc_conv_dgrad(p, gemm_x3(), (cudaStream_t)stream);
    const int N = p->B * p->L;
    const int KC = min(KC_MAX, round_up(p->M, 4));
    bool vec = (p->L & 3) == 0 && (p->K & 3) == 0 && gal16(p->GV) && (!p->coef_a || gal16(p->Z));
    for (int i = 0; i < p->n_seg; ++i) vec = vec && gal16(p->W[i]);
    static size_t configured[2] = {0, 0};
    e = vec ? set_smem(k_conv_dgrad<true>, gemm_smem_bytes(KC), &configured[1])
            : set_smem(k_conv_dgrad<false>, gemm_smem_bytes(KC), &configured[0]);
    if (e) return e;
    dim3 grid((N + TN - 1) / TN, (p->K + TM - 1) / TM);
    if (vec)
        launch_k(k_conv_dgrad<true>, grid, GT, gemm_smem_bytes(KC), (cudaStream_t)stream, *p, N, KC);
    else
        launch_k(k_conv_dgrad<false>, grid, GT, gemm_smem_bytes(KC), (cudaStream_t)stream, *p, N, KC);
    BMNAS_LAUNCH_CHECK();
    return BMNAS_OK;
}

extern "C" int bmnas_conv_wgrad(const bmnas_conv_params* p, void* stream) {
    int e = conv_check(p);
    if (e) return e;
    if (!p->GV) return BMNAS_EINVAL;
    if (p->coef_a && (!p->coef_b || !p->coef_c || !p->Z)) return BMNAS_EINVAL;
    for (int i = 0; i < p->n_src; ++i)
        if (!p->src[i]) return BMNAS_EINVAL;
    BMNAS_DRY_RETURN();
    // fp32 parity engines: the transposing FFMA kernel (gemm_sg.cu) beats the 3xTF32 UMMA ring while the problem is
    // latency bound (B=96: 9 us vs 21 us); the UMMA ring wins from a few thousand reduction columns on (measured,
    // profiles/r01_v6_kernel_times_B1024.txt: 65 us vs 33 us at B*L = 8192).  FFMA mode (0) has no tensor-core
    // engine and the reduced-precision mode (2) keeps the single-pass TF32 kernel.
    // The crossover is on the GEMM's work, like bmnas_conv_image_fmt: NTU (M*K = 384 x 128) changes engine at 2560 columns as
    // measured in round 1; Ego-large (768 x 256 weights, 1536 columns: 302 M MACs, 56 us on the FFMA kernel against
    // ~15 us on wgrad_ws.cu) now lands on the tensor cores too.  BMNAS_TC_MIN_MACS_W overrides.
    const long long macs = (long long)p->B * p->L * p->M * p->K;
    static long long min_macs_w = -1;
    if (min_macs_w < 0) {
        const char* e_ = getenv("BMNAS_TC_MIN_MACS_W");
        min_macs_w = e_ ? atoll(e_) : 60000000LL;
    }
    const bool tc_ok = bmnas_gemm_mode_flag && tc_eligible(p, 2);
    if (bmnas_gemm_mode_flag != 2 && sgw_eligible(p) && (macs <= min_macs_w || !tc_ok)) return sg_conv_wgrad(p, (cudaStream_t)stream);
    if (bmnas_gemm_mode_flag && tc_eligible(p, 2)) return tc_conv_wgrad(p, gemm_x3(), (cudaStream_t)stream);
    const int N = p->B * p->L, L = p->L;
    const int tiles = ((p->K + TN - 1) / TN) * ((p->M + TM - 1) / TM);
    bool vec = (L & 3) == 0 && L <= KC_MAX && gal16(p->GV) && (!p->coef_a || gal16(p->Z));
    for (int i = 0; i < p->n_src; ++i) vec = vec && gal16(p->src[i]);
    const int unit = vec ? L : 4;   // chunk granularity (whole samples when vectorised)
    int splits = p->splits;
    if (splits <= 0) {
        splits = (2 * kNumSMs + tiles - 1) / tiles;          // ~2 CTAs per SM
        const int maxs = (N + 63) / 64;                      // at least 64 reduction columns per split
        if (splits > maxs) splits = maxs;
        if (splits < 1) splits = 1;
    }
    int chunkN = round_up((N + splits - 1) / splits, unit);
    splits = (N + chunkN - 1) / chunkN;
    const int KC = min(round_up(chunkN, unit), (KC_MAX / unit) * unit);
    static size_t configured[2] = {0, 0};
    e = vec ? set_smem(k_conv_wgrad<true>, gemm_smem_bytes(KC), &configured[1])
            : set_smem(k_conv_wgrad<false>, gemm_smem_bytes(KC), &configured[0]);
    if (e) return e;
    dim3 grid((p->K + TN - 1) / TN, (p->M + TM - 1) / TM, splits);
    if (vec)
        launch_k(k_conv_wgrad<true>, grid, GT, gemm_smem_bytes(KC), (cudaStream_t)stream, *p, N, chunkN, KC);
    else
        launch_k(k_conv_wgrad<false>, grid, GT, gemm_smem_bytes(KC), (cudaStream_t)stream, *p, N, chunkN, KC);
    BMNAS_LAUNCH_CHECK();
    return BMNAS_OK;
}
