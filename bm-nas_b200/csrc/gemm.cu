// fp32 (FFMA) GEMM family for the 1x1 convolutions of the fusion cell:
//   conv_fwd   Z[b,m,l]  = sum_k Weff[m,k] U[b,k,l] + bias[m]      (+ BN batch statistics)
//   conv_dgrad dU[b,k,l] = sum_m Weff[m,k] dz[b,m,l]
//   conv_wgrad dW[m,k]   = sum_{b,l} dz[b,m,l] U[b,k,l],  dbias[m] = sum dz
// U is a virtual channel concat (never materialised), the output rows are stacked
// weight segments, w_fold folds cat([t,t]); dz is produced on the fly from
// (GV, Z, coef) = BatchNorm backward fused into the operand load.
// This is the fp32-exact path (parity 1e-5 vs the reference on CPU); the bf16
// tcgen05 path lives in gemm_tc.cu.
// 64x64x16 tiles, 256 threads, 4x4 register micro-tiles, register prefetch of
// the next K chunk.  All shapes are bounds-checked (M, K, N arbitrary).
#include "common.cuh"

namespace bmnas {

constexpr int TM = 64, TN = 64, TK = 16, GT = 256;
constexpr int PAD = 4;

struct Tiles {
    float A[TK][TM + PAD];
    float B[TK][TN + PAD];
};

__device__ __forceinline__ void mma_chunk(const Tiles& t, float (&acc)[4][4], int ty, int tx) {
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
        const float4 a = *reinterpret_cast<const float4*>(&t.A[kk][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&t.B[kk][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w};
        const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
}

// row m of the stacked weight: pointer to W_seg[m_local][0]; also returns segment/local index
__device__ __forceinline__ const float* w_row(const bmnas_conv_params& p, int m, int ldw, int* seg, int* ml) {
    int s = 0;
    while (s + 1 < p.n_seg && m >= p.seg_M[s]) {
        m -= p.seg_M[s];
        ++s;
    }
    if (seg) *seg = s;
    if (ml) *ml = m;
    return p.W[s] + (long long)m * ldw;
}

// channel k of the virtual concat -> (source, local channel)
__device__ __forceinline__ void src_of(const bmnas_conv_params& p, int k, int* s, int* kl) {
    int i = 0;
    while (i + 1 < p.n_src && k >= p.src_C[i]) {
        k -= p.src_C[i];
        ++i;
    }
    *s = i;
    *kl = k;
}

struct Wf {  // Welford triple
    float n, mean, m2;
};
__device__ __forceinline__ Wf wf_merge(Wf a, Wf b) {
    Wf r;
    r.n = a.n + b.n;
    if (r.n <= 0.f) {
        r.mean = 0.f;
        r.m2 = 0.f;
        return r;
    }
    const float d = b.mean - a.mean;
    r.mean = a.mean + d * (b.n / r.n);
    r.m2 = a.m2 + b.m2 + d * d * (a.n * b.n / r.n);
    return r;
}

// ------------------------------------------------------------------ forward
__global__ void __launch_bounds__(GT, 2) k_conv_fwd(const bmnas_conv_params p, const int N, const int n_col_tiles) {
    __shared__ __align__(16) Tiles t;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    const int K = p.K, M = p.M, L = p.L, ldw = p.w_fold * p.K;

    // A loader: kk = tid&15, rows mi + 16 i
    const int a_kk = tid & 15, a_mi = tid >> 4;
    const float* a_row[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + a_mi + 16 * i;
        a_row[i] = m < M ? w_row(p, m, ldw, nullptr, nullptr) : nullptr;
    }
    // B loader: col = tid&63, kk = (tid>>6) + 4 i
    const int b_col = tid & 63, b_kq = tid >> 6;
    const int nB = n0 + b_col;
    const bool vB = nB < N;
    const int bB = vB ? nB / L : 0, lB = vB ? nB % L : 0;

    float ra[4], rb[4];
    auto load = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = k0 + a_kk;
            float v = 0.f;
            if (a_row[i] && k < K) {
                v = __ldg(a_row[i] + k);
                if (p.w_fold == 2) v += __ldg(a_row[i] + K + k);
            }
            ra[i] = v;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = k0 + b_kq + 4 * i;
            float v = 0.f;
            if (vB && k < K) {
                int s, kl;
                src_of(p, k, &s, &kl);
                v = __ldg(p.src[s] + ((long long)bB * p.src_C[s] + kl) * L + lB);
            }
            rb[i] = v;
        }
    };

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    load(0);
    for (int k0 = 0; k0 < K; k0 += TK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            t.A[a_kk][a_mi + 16 * i] = ra[i];
            t.B[b_kq + 4 * i][b_col] = rb[i];
        }
        __syncthreads();
        if (k0 + TK < K) load(k0 + TK);
        mma_chunk(t, acc, ty, tx);
        __syncthreads();
    }

    // ---- epilogue: bias, store, per-tile BN statistics
    const int cnt = min(TN, N - n0);
    const int nb = n0 + tx * 4;
    const bool vec = (L % 4 == 0) && (nb + 3 < N);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        float bias = 0.f;
        if (m < M) {
            int s, ml;
            w_row(p, m, ldw, &s, &ml);
            if (p.bias[s]) bias = __ldg(p.bias[s] + ml);
        }
        float z[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) z[j] = acc[i][j] + bias;
        if (m < M) {
            if (vec) {
                const int b = nb / L, l = nb % L;
                *reinterpret_cast<float4*>(p.Z + ((long long)b * M + m) * L + l) = make_float4(z[0], z[1], z[2], z[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int n = nb + j;
                    if (n < N) p.Z[((long long)(n / L) * M + m) * L + (n % L)] = z[j];
                }
            }
        }
        if (p.bn_mode == 1) {
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) s += (nb + j < N) ? z[j] : 0.f;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            const float mean = s / (float)cnt;
            float d2 = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float d = z[j] - mean;
                d2 += (nb + j < N) ? d * d : 0.f;
            }
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
        if (blockIdx.x == 0 && tid < TM) {
            const int m = m0 + tid;
            if (m < M) {
                int s, ml;
                w_row(p, m, ldw, &s, &ml);
                p.mean[m] = p.running_mean[s][ml];
                p.rstd[m] = 1.f / sqrtf(p.running_var[s][ml] + p.eps);
            }
        }
        return;
    }
    if (p.bn_mode != 1) return;

    // ---- last CTA of this row-tile merges the per-column-tile statistics (Chan), fixed order
    if (!last_block(p.counter + blockIdx.y, gridDim.x)) return;
    const int r = tid >> 2, q = tid & 3;
    const int m = m0 + r;
    Wf w = {0.f, 0.f, 0.f};
    if (m < M) {
        for (int tix = q; tix < n_col_tiles; tix += 4) {
            const float* pp = p.stat_part + ((long long)tix * M + m) * 2;
            Wf b = {(float)min(TN, N - tix * TN), ld_cg(pp), ld_cg(pp + 1)};
            w = wf_merge(w, b);
        }
    }
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
        Wf b;
        b.n = __shfl_xor_sync(0xffffffffu, w.n, o);
        b.mean = __shfl_xor_sync(0xffffffffu, w.mean, o);
        b.m2 = __shfl_xor_sync(0xffffffffu, w.m2, o);
        // merge in a lane-independent order so all 4 lanes agree bit-for-bit
        w = ((q & o) == 0) ? wf_merge(w, b) : wf_merge(b, w);
    }
    if (q == 0 && m < M) {
        const float var = w.m2 / (float)N;
        p.mean[m] = w.mean;
        p.rstd[m] = 1.f / sqrtf(var + p.eps);
        int s, ml;
        w_row(p, m, ldw, &s, &ml);
        if (p.running_mean[s]) {
            const float unb = w.m2 / (float)max(N - 1, 1);
            p.running_mean[s][ml] = (1.f - p.momentum) * p.running_mean[s][ml] + p.momentum * w.mean;
            p.running_var[s][ml] = (1.f - p.momentum) * p.running_var[s][ml] + p.momentum * unb;
            if (ml == 0 && p.num_batches_tracked[s]) *p.num_batches_tracked[s] += 1;
        }
    }
}

// upstream-gradient operand with BatchNorm backward folded in
__device__ __forceinline__ float load_dz(const bmnas_conv_params& p, long long idx, int m) {
    float g = __ldg(p.GV + idx);
    if (p.coef_a) g = fmaf(__ldg(p.coef_a + m), g, fmaf(__ldg(p.coef_b + m), __ldg(p.Z + idx), __ldg(p.coef_c + m)));
    return g;
}

// ------------------------------------------------------------------ dgrad
__global__ void __launch_bounds__(GT, 2) k_conv_dgrad(const bmnas_conv_params p, const int N) {
    __shared__ __align__(16) Tiles t;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int kt0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    const int K = p.K, M = p.M, L = p.L, ldw = p.w_fold * p.K;

    const int a_kl = tid & 63, a_mq = tid >> 6;  // A[mm][k]: k fastest (coalesced along W rows)
    const int b_col = tid & 63, b_mq = tid >> 6;
    const int nB = n0 + b_col;
    const bool vB = nB < N;
    const int bB = vB ? nB / L : 0, lB = vB ? nB % L : 0;

    float ra[4], rb[4];
    auto load = [&](int mk0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = mk0 + a_mq + 4 * i, k = kt0 + a_kl;
            float v = 0.f;
            if (m < M && k < K) {
                const float* r = w_row(p, m, ldw, nullptr, nullptr);
                v = __ldg(r + k);
                if (p.w_fold == 2) v += __ldg(r + K + k);
            }
            ra[i] = v;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = mk0 + b_mq + 4 * i;
            float v = 0.f;
            if (vB && m < M) v = load_dz(p, ((long long)bB * M + m) * L + lB, m);
            rb[i] = v;
        }
    };

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    load(0);
    for (int mk0 = 0; mk0 < M; mk0 += TK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            t.A[a_mq + 4 * i][a_kl] = ra[i];
            t.B[b_mq + 4 * i][b_col] = rb[i];
        }
        __syncthreads();
        if (mk0 + TK < M) load(mk0 + TK);
        mma_chunk(t, acc, ty, tx);
        __syncthreads();
    }

    const int nb = n0 + tx * 4;
    const bool vec = (L % 4 == 0) && (nb + 3 < N);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k = kt0 + ty * 4 + i;
        if (k >= K) continue;
        int s, kl;
        src_of(p, k, &s, &kl);
        float* dst = p.gsrc[s];
        if (!dst) continue;
        const bool accum = p.gsrc_accum[s] != 0;
        if (vec) {
            const int b = nb / L, l = nb % L;
            float4* d = reinterpret_cast<float4*>(dst + ((long long)b * p.src_C[s] + kl) * L + l);
            float4 o = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
            if (accum) {
                const float4 c = *d;
                o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w;
            }
            *d = o;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
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
__global__ void __launch_bounds__(GT, 2) k_conv_wgrad(const bmnas_conv_params p, const int N, const int chunkN) {
    __shared__ __align__(16) Tiles t;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int kt0 = blockIdx.x * TN, m0 = blockIdx.y * TM;
    const int K = p.K, M = p.M, L = p.L, ldw = p.w_fold * p.K;
    const int nbeg = blockIdx.z * chunkN, nend = min(N, nbeg + chunkN);

    const int l_nn = tid & 15, l_ri = tid >> 4;  // both loaders: reduction index fastest
    float ca[4], cb[4], cc[4];
    const float* ub[4];
    int ucs[4], ukl[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + l_ri + 16 * i;
        ca[i] = 1.f; cb[i] = 0.f; cc[i] = 0.f;
        if (p.coef_a && m < M) {
            ca[i] = __ldg(p.coef_a + m); cb[i] = __ldg(p.coef_b + m); cc[i] = __ldg(p.coef_c + m);
        }
        const int k = kt0 + l_ri + 16 * i;
        ub[i] = nullptr; ucs[i] = 0; ukl[i] = 0;
        if (k < K) {
            int s, kl;
            src_of(p, k, &s, &kl);
            ub[i] = p.src[s]; ucs[i] = p.src_C[s]; ukl[i] = kl;
        }
    }

    float ra[4], rb[4];
    auto load = [&](int nk0) {
        const int n = nk0 + l_nn;
        const bool v = n < nend;
        const int b = v ? n / L : 0, l = v ? n % L : 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + l_ri + 16 * i;
            float x = 0.f;
            if (v && m < M) {
                const long long idx = ((long long)b * M + m) * L + l;
                x = __ldg(p.GV + idx);
                if (p.coef_a) x = fmaf(ca[i], x, fmaf(cb[i], __ldg(p.Z + idx), cc[i]));
            }
            ra[i] = x;
            float u = 0.f;
            if (v && ub[i]) u = __ldg(ub[i] + ((long long)b * ucs[i] + ukl[i]) * L + l);
            rb[i] = u;
        }
    };

    float acc[4][4], rs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    if (nbeg < nend) load(nbeg);
    for (int nk0 = nbeg; nk0 < nend; nk0 += TK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            t.A[l_nn][l_ri + 16 * i] = ra[i];
            t.B[l_nn][l_ri + 16 * i] = rb[i];
        }
        __syncthreads();
        if (nk0 + TK < nend) load(nk0 + TK);
        mma_chunk(t, acc, ty, tx);
        if (blockIdx.x == 0 && tx == 0) {
#pragma unroll
            for (int kk = 0; kk < TK; ++kk)
#pragma unroll
                for (int i = 0; i < 4; ++i) rs[i] += t.A[kk][ty * 4 + i];
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
        int s, ml;
        w_row(p, m, ldw, &s, &ml);
        if (p.gW[s]) {
            float* row = p.gW[s] + (long long)ml * ldw;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = kt0 + tx * 4 + j;
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

}  // namespace bmnas

using namespace bmnas;

extern "C" long long bmnas_conv_stat_part_size(const bmnas_conv_params* p) {
    const long long N = (long long)p->B * p->L;
    return ((N + TN - 1) / TN) * p->M * 2;
}
extern "C" int bmnas_conv_num_counters(const bmnas_conv_params* p) { return (p->M + TM - 1) / TM; }

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
    const int N = p->B * p->L;
    dim3 grid((N + TN - 1) / TN, (p->M + TM - 1) / TM);
    k_conv_fwd<<<grid, GT, 0, (cudaStream_t)stream>>>(*p, N, (int)grid.x);
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
    const int N = p->B * p->L;
    dim3 grid((N + TN - 1) / TN, (p->K + TM - 1) / TM);
    k_conv_dgrad<<<grid, GT, 0, (cudaStream_t)stream>>>(*p, N);
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
    const int N = p->B * p->L;
    const int tiles = ((p->K + TN - 1) / TN) * ((p->M + TM - 1) / TM);
    int splits = p->splits;
    if (splits <= 0) {
        splits = (2 * kNumSMs + tiles - 1) / tiles;
        const int maxs = (N + 4 * TK - 1) / (4 * TK);  // at least 4 K-chunks per split
        if (splits > maxs) splits = maxs;
        if (splits < 1) splits = 1;
    }
    int chunkN = (N + splits - 1) / splits;
    chunkN = ((chunkN + TK - 1) / TK) * TK;
    splits = (N + chunkN - 1) / chunkN;
    dim3 grid((p->K + TN - 1) / TN, (p->M + TM - 1) / TM, splits);
    k_conv_wgrad<<<grid, GT, 0, (cudaStream_t)stream>>>(*p, N, chunkN);
    BMNAS_LAUNCH_CHECK();
    return BMNAS_OK;
}
