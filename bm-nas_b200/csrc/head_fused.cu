// Classifier head + loss + their backward in ONE launch -- bmnas_head_fused (include/bmnas_b200.h).
//
//   logits = x W^T + bias                                   central_classifier   (ntu_darts_searchable.py:100-101, 176-178)
//   loss   = mean CE(logits, labels) | mean BCE-with-logits  criterion            (ntu_darts_searchable.py:25, mmimdb_darts_searchable.py:22)
//   g      = d loss / d logits                               softmax - onehot | sigmoid - target, mean-scaled
//   gx     = g W                                             input gradient of the classifier (what loss.backward() hands
//                                                            to the fusion network's backward)
//
// At the reference batch this chain was five dependent launches on the critical path of every half step (linear fwd ->
// loss fwd -> fill -> loss bwd -> linear dX: ~40 us of a ~300 us half step at NTU, tools/timeline.py) for 47 MFLOP.
// Everything in it is sample-local except the mean, so ONE kernel does it:
//   * a thread-block CLUSTER of 8 CTAs owns a group of SB <= 8 samples; the reduction dimension K (2048 at NTU) is
//     split over the cluster in 128-column chunks (chunk j -> rank j % 8), so W (491 KB) is spread over ~128 SMs and read
//     once per sample group;
//   * phase 1: every CTA forms the partial logits of its chunks from shared memory (cp.async-staged W chunk, padded rows:
//     conflict-free 128-bit reads; a thread owns one class x two samples);
//   * the partials meet through DISTRIBUTED SHARED MEMORY: one cluster barrier, every CTA sums the 8 partials in rank
//     order (bit-identical logits in all 8 CTAs, deterministic), then computes softmax / sigmoid, the per-sample loss
//     and g redundantly; rank 0 stores logits, g and the group's loss partial;
//   * phase 2: gx for the CTA's own columns from the W chunks that are still resident in shared memory;
//   * the last cluster to finish (atomic ticket) sums the group partials in fixed order -> loss.
// The weight / bias gradients (gW = g^T x) only feed the optimiser: bmnas_linear_bwd computes them from g on the side
// branch of the graph, as before.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace bmnas {
namespace hd {

constexpr int CS = 8;            // CTAs per cluster = K split
constexpr int KC = 128;          // columns per chunk
constexpr int LDW = KC + 4;      // padded W row (floats): a quarter warp reading 8 consecutive rows hits 8 distinct 16-byte bank groups
constexpr int TH = 256;
constexpr int SBT = 8;           // samples per group (max)
constexpr int NMAX = 128;        // classes (max)
constexpr int MAXR = (NMAX * (SBT / 2) + TH - 1) / TH;   // (class, sample pair) outputs per thread in phase 1
constexpr size_t SMEM_BUDGET = 200 * 1024;

__device__ __forceinline__ void cp_async16(float* dst_smem, const float* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

__host__ __device__ inline size_t slot_floats(int N) { return (size_t)N * LDW + (size_t)SBT * KC; }

__global__ void __launch_bounds__(TH) k_head(const bmnas_head_params p, const int SB, const int n_chunks, const int nbuf) {
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) float smem[];
    __shared__ float part[SBT * NMAX];      // this CTA's partial logits [sample][class] (read by the whole cluster)
    __shared__ float lg[SBT * NMAX];        // logits, then g = d loss / d logits
    __shared__ float wl[SBT];               // per-sample loss terms
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rank = (int)cluster.block_rank();
    const int grp = blockIdx.y;
    const int N = p.N, K = p.K;
    const int b0 = grp * SB;
    const int nsv = min(SB, p.B - b0);
    const size_t slot = slot_floats(N);
    const int n_my = (n_chunks - rank + CS - 1) / CS;       // chunks rank, rank + CS, ...
    const bool resident = n_my <= nbuf;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);

    auto stage_w = [&](int i, int s) {
        const int k0 = (rank + i * CS) * KC, kc = min(KC, K - k0);
        float* ws = smem + (size_t)s * slot;
        for (int u = tid; u < N * (KC / 4); u += TH) {
            const int c = u >> 5, q = u & 31;
            float* d = ws + c * LDW + q * 4;
            if (q * 4 < kc) cp_async16(d, p.W + (long long)c * K + k0 + q * 4);
            else *reinterpret_cast<float4*>(d) = z4;
        }
    };
    auto stage_x = [&](int i, int s) {
        const int k0 = (rank + i * CS) * KC, kc = min(KC, K - k0);
        float* xs = smem + (size_t)s * slot + (size_t)N * LDW;
        for (int u = tid; u < SBT * (KC / 4); u += TH) {
            const int b = u >> 5, q = u & 31;
            float* d = xs + b * KC + q * 4;
            if (b < nsv && q * 4 < kc) cp_async16(d, p.x + (long long)(b0 + b) * K + k0 + q * 4);
            else *reinterpret_cast<float4*>(d) = z4;
        }
    };

    // ---- early section: the weights were last written by the optimiser of the previous half step (>= 2 kernels ago)
    const int n_first = resident ? n_my : min(n_my, 1);
    for (int i = 0; i < n_first; ++i) stage_w(i, i);
    pdl_wait();
    pdl_trigger();
    for (int i = 0; i < n_first; ++i) stage_x(i, i);
    cp_commit();

    // ---- phase 1: partial logits of this CTA's chunks; thread = (class c, sample pair bp) for up to MAXR outputs
    const int npair = (SB + 1) >> 1, nout = N * npair;
    float acc[MAXR][2];
    int oc[MAXR], obp[MAXR];
#pragma unroll
    for (int r = 0; r < MAXR; ++r) {
        acc[r][0] = acc[r][1] = 0.f;
        const int q = tid + r * TH;
        obp[r] = q < nout ? q / N : -1;
        oc[r] = q < nout ? q - obp[r] * N : 0;
    }
    for (int i = 0; i < n_my; ++i) {
        int s;
        if (resident) {
            s = i;
            if (i == 0) {
                cp_wait<0>();
                __syncthreads();
            }
        } else {
            s = i & 1;
            if (i + 1 < n_my) {
                stage_w(i + 1, (i + 1) & 1);
                stage_x(i + 1, (i + 1) & 1);
                cp_commit();
                cp_wait<1>();
            } else {
                cp_wait<0>();
            }
            __syncthreads();
        }
        const float* ws = smem + (size_t)s * slot;
        const float* xs = ws + (size_t)N * LDW;
#pragma unroll
        for (int r = 0; r < MAXR; ++r) {
            if (obp[r] < 0) continue;
            const float4* w4 = reinterpret_cast<const float4*>(ws + oc[r] * LDW);
            const float4* xa = reinterpret_cast<const float4*>(xs + (2 * obp[r]) * KC);
            const float4* xb = xa + KC / 4;         // sample 2*bp + 1 (a zero row when it does not exist: SBT is even)
            float a0 = acc[r][0], a1 = acc[r][1];
#pragma unroll 8
            for (int k4 = 0; k4 < KC / 4; ++k4) {
                const float4 w = w4[k4], u = xa[k4], v = xb[k4];
                a0 = fmaf(w.x, u.x, a0); a0 = fmaf(w.y, u.y, a0); a0 = fmaf(w.z, u.z, a0); a0 = fmaf(w.w, u.w, a0);
                a1 = fmaf(w.x, v.x, a1); a1 = fmaf(w.y, v.y, a1); a1 = fmaf(w.z, v.z, a1); a1 = fmaf(w.w, v.w, a1);
            }
            acc[r][0] = a0; acc[r][1] = a1;
        }
        if (!resident) __syncthreads();            // the slot is restaged two iterations later
    }
#pragma unroll
    for (int r = 0; r < MAXR; ++r) {
        if (obp[r] < 0) continue;
        part[(2 * obp[r]) * N + oc[r]] = acc[r][0];
        part[(2 * obp[r] + 1) * N + oc[r]] = acc[r][1];
    }
    // streaming mode: the first W chunk of phase 2 travels while the cluster meets
    if (!resident && p.gx) {
        stage_w(0, 0);
        cp_commit();
    }
    cluster.sync();

    // ---- logits = bias + sum of the cluster's partials, fixed rank order (every CTA gets the same bits)
    const float* rp[CS];
#pragma unroll
    for (int r = 0; r < CS; ++r) rp[r] = cluster.map_shared_rank(part, r);
    for (int o = tid; o < nsv * N; o += TH) {
        float v[CS];
#pragma unroll
        for (int r = 0; r < CS; ++r) v[r] = rp[r][o];
        float s = p.bias ? __ldg(p.bias + (o % N)) : 0.f;
#pragma unroll
        for (int r = 0; r < CS; ++r) s += v[r];
        lg[o] = s;
    }
    __syncthreads();

    // ---- loss and g: warp w owns sample w of the group (same arithmetic as k_loss_fwd, loss_optim.cu)
    if (warp < SBT) {
        float lsum = 0.f;
        if (warp < nsv) {
            const int b = b0 + warp;
            float* x = lg + warp * N;
            const bool store = rank == 0;
            if (p.kind == 0) {
                const float invB = 1.f / (float)p.B;
                float mx = -INFINITY;
                for (int j = lane; j < N; j += 32) mx = fmaxf(mx, x[j]);
                mx = warp_max(mx);
                float s = 0.f;
                for (int j = lane; j < N; j += 32) s += expf(x[j] - mx);
                s = warp_sum(s);
                const long long lab64 = p.labels[b];
                const bool bad = lab64 < 0 || lab64 >= (long long)N;     // NaN loss and gradient, as k_loss_fwd does
                const int lab = bad ? 0 : (int)lab64;
                const float lse = mx + logf(s);
                lsum = bad ? __int_as_float(0x7fc00000) : lse - x[lab];
                __syncwarp();
                const float inv = 1.f / s;
                for (int j = lane; j < N; j += 32) {
                    const float xv = x[j];
                    const float g = bad ? __int_as_float(0x7fc00000) : (expf(xv - mx) * inv - (j == lab ? 1.f : 0.f)) * invB;
                    if (store) {
                        p.logits[(long long)b * N + j] = xv;
                        if (p.glogits) p.glogits[(long long)b * N + j] = g;
                    }
                    x[j] = g;
                }
            } else {
                const float inv = 1.f / ((float)p.B * (float)N);
                for (int j = lane; j < N; j += 32) {
                    const float xv = x[j], t = p.targets[(long long)b * N + j];
                    lsum += fmaxf(xv, 0.f) - xv * t + log1pf(expf(-fabsf(xv)));
                    const float g = (sigmoidf_(xv) - t) * inv;
                    if (store) {
                        p.logits[(long long)b * N + j] = xv;
                        if (p.glogits) p.glogits[(long long)b * N + j] = g;
                    }
                    x[j] = g;
                }
                lsum = warp_sum(lsum);
            }
        }
        if (lane == 0) wl[warp] = lsum;
    }
    __syncthreads();
    if (rank == 0 && tid == 0) {
        float s = 0.f;
        for (int w = 0; w < nsv; ++w) s += wl[w];
        p.partials[grp] = s;
    }

    // ---- phase 2: gx[b][k] = sum_c g[b][c] W[c][k] for this CTA's columns; warp w -> samples 2w, 2w+1; lane -> one float4 column
    if (p.gx) {
        const bool act = 2 * warp < nsv;
        const bool two = 2 * warp + 1 < nsv;
        const float* g0 = lg + (2 * warp) * N;
        const float* g1 = g0 + N;                   // rows >= nsv of lg are never read as gradients (two == false)
        for (int i = 0; i < n_my; ++i) {
            int s;
            if (resident) {
                s = i;
            } else {
                s = i & 1;
                if (i + 1 < n_my) {
                    stage_w(i + 1, (i + 1) & 1);
                    cp_commit();
                    cp_wait<1>();
                } else {
                    cp_wait<0>();
                }
                __syncthreads();
            }
            const int k0 = (rank + i * CS) * KC + lane * 4;
            if (act && k0 < K) {
                const float* ws = smem + (size_t)s * slot + lane * 4;
                float4 a0 = z4, a1 = z4;
#pragma unroll 4
                for (int c = 0; c < N; ++c) {
                    const float4 w = *reinterpret_cast<const float4*>(ws + c * LDW);
                    const float u = g0[c], v = two ? g1[c] : 0.f;
                    a0.x = fmaf(u, w.x, a0.x); a0.y = fmaf(u, w.y, a0.y); a0.z = fmaf(u, w.z, a0.z); a0.w = fmaf(u, w.w, a0.w);
                    a1.x = fmaf(v, w.x, a1.x); a1.y = fmaf(v, w.y, a1.y); a1.z = fmaf(v, w.z, a1.z); a1.w = fmaf(v, w.w, a1.w);
                }
                *reinterpret_cast<float4*>(p.gx + (long long)(b0 + 2 * warp) * K + k0) = a0;
                if (two) *reinterpret_cast<float4*>(p.gx + (long long)(b0 + 2 * warp + 1) * K + k0) = a1;
            }
            if (!resident) __syncthreads();
        }
    }

    // ---- mean loss: the last group to arrive sums the group partials in fixed order
    if (rank == 0) {
        if (last_block(p.counter, gridDim.y)) {
            if (tid == 0) {
                float s = 0.f;
                for (unsigned g = 0; g < gridDim.y; ++g) s += ld_cg(p.partials + g);
                p.loss[0] = s / (p.kind == 0 ? (float)p.B : (float)p.B * (float)N);
            }
        }
    }
    cluster.sync();          // no CTA leaves while a peer may still read its partial logits
}

struct Geo {
    int SB, G, n_chunks, nbuf;
    size_t smem;
};

static bool geometry(const bmnas_head_params* p, Geo* g) {
    const int max_groups = kNumSMs / CS;                    // one wave of clusters
    int SB = (p->B + max_groups - 1) / max_groups;
    if (SB < 2) SB = 2;
    if (SB & 1) ++SB;                                       // phase 1 works on sample pairs
    if (SB > SBT) SB = SBT;
    g->SB = SB;
    g->G = (p->B + SB - 1) / SB;
    g->n_chunks = (p->K + KC - 1) / KC;
    const int n_my_max = (g->n_chunks + CS - 1) / CS;
    const size_t slot = slot_floats(p->N) * sizeof(float);
    int fit = (int)(SMEM_BUDGET / slot);
    if (fit < 2) return false;
    g->nbuf = n_my_max <= fit ? n_my_max : 2;
    g->smem = (size_t)g->nbuf * slot;
    return true;
}

static bool al16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }

}  // namespace hd
}  // namespace bmnas

using namespace bmnas;

extern "C" int bmnas_head_supported(const bmnas_head_params* p) {
    using namespace hd;
    if (!p || p->B < 1 || p->K < 4 || (p->K & 3) || p->N < 1 || p->N > NMAX) return 0;
    if (p->kind != 0 && p->kind != 1) return 0;
    if (!p->x || !p->W || !p->logits || !p->loss || !p->partials || !p->counter) return 0;
    if (p->kind == 0 ? !p->labels : !p->targets) return 0;
    if (!al16(p->x) || !al16(p->W) || (p->gx && !al16(p->gx))) return 0;
    // every sample group re-reads W: beyond a few waves of clusters the three-kernel path (W read once per 12 samples by
    // 96 CTAs, k_lin_fwd) is the better machine
    if (p->B > 8 * SBT * (kNumSMs / CS)) return 0;
    Geo g;
    return geometry(p, &g) ? 1 : 0;
}

extern "C" long long bmnas_head_partials_size(const bmnas_head_params* p) {
    using namespace hd;
    Geo g;
    if (!p || !geometry(p, &g)) return 0;
    return g.G;
}

extern "C" int bmnas_head_fused(const bmnas_head_params* p, void* stream) {
    using namespace hd;
    if (!bmnas_head_supported(p)) return BMNAS_EINVAL;
    BMNAS_DRY_RETURN();
    Geo g;
    geometry(p, &g);
    static size_t configured = 0;
    if (g.smem > configured) {
        if (cudaFuncSetAttribute(k_head, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BUDGET) != cudaSuccess) return BMNAS_ELAUNCH;
        configured = SMEM_BUDGET;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS, g.G);
    cfg.blockDim = dim3(TH);
    cfg.dynamicSmemBytes = g.smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = bmnas_pdl_flag ? 2 : 1;
    cudaLaunchKernelEx(&cfg, k_head, *p, g.SB, g.n_chunks, g.nbuf);
    BMNAS_LAUNCH_CHECK();
    return BMNAS_OK;
}
