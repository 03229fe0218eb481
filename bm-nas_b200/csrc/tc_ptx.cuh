// tcgen05 / TMEM / TMA / mbarrier PTX wrappers and the UMMA descriptor helpers shared by the tensor-core kernels
// (gemm_tc.cu: stand-alone conv GEMMs; mixed_tc.cu: the fused NodeMixedOp kernels).
#pragma once
#include "common.cuh"

namespace bmnas {
namespace tc {

constexpr int TCM = 128;  // accumulator rows per CTA = UMMA M
constexpr int KC = 32;    // tf32 reduction elements per 128-byte swizzled operand row (4 UMMA k-steps of 8)

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, 0x989680;\n\t"
            "selp.b32 %0, 1, 0, P1;\n\t"
            "}"
            : "=r"(done)
            : "r"(s32(bar)), "r"(phase)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(s32(bar)) : "memory");
}
// TMA bulk copy global -> shared; completion is signalled on `bar` as transaction bytes
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst_smem)),
                 "l"(src), "r"(bytes), "r"(s32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(bar)) : "memory");
}
// One lane of a CONVERGED warp.  The MMA / TMA issue loops must run warp-uniformly with only the instruction itself
// predicated on this: inside `if (threadIdx.x == 0)` the compiler has to assume divergent operands and wraps every
// tcgen05.mma in an R2UR + ELECT + BRA.U.ANY loop -- measured 157 cycles per MMA instead of 48 (N=64) / 64 (N=128)
// (tools/ubench/mma_rate.cu).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(dst_smem)), "r"(ncols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols));
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, tf32 inputs, fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns: thread t of the warp receives lane (base lane + t)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// tcgen05.ld of 16 columns without the wait (issue several, then ONE tcgen05.wait::ld)
__device__ __forceinline__ void tmem_ld16_raw(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
// sum of the same 16 columns of the first `used` of NA accumulators (accumulator a starts BNC columns after a-1)
template <int NA, int BNC>
__device__ __forceinline__ void tmem_ld16_sum(uint32_t taddr, float (&v)[16], int used) {
    tmem_ld16(taddr, v);
#pragma unroll
    for (int a = 1; a < NA; ++a) {
        if (a < used) {
            float w[16];
            tmem_ld16(taddr + (uint32_t)(a * BNC), w);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += w[i];
        }
    }
}
__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// [0,14) start>>4, [16,30) leading byte offset>>4 (unused for swizzled K-major, 1), [32,46) stride byte offset>>4
// (8 rows x 128 B = 1024), [46,48) version=1, [61,64) layout type (2 = SWIZZLE_128B).  One operand row holds the
// KC = 32 reduction elements of a slab in 128 contiguous bytes; inside each 8-row x 128 B atom the 16-byte chunk c
// of row r sits at chunk position c ^ (r & 7) (sw_off below).  A k-step of 8 tf32 advances the start address by
// 32 bytes inside the atom (the hardware applies the XOR to the address bits), atoms are 1024-byte aligned.
// (The SWIZZLE_NONE "interleave" layout this replaced fed the tensor core at ~40 B/clk: 130 cycles per
// 128x32x8 MMA, measured; see profiles/.)
__device__ __forceinline__ uint64_t kdesc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | (1ull << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ __forceinline__ uint32_t sw_off(int row, int kc) {
    return (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u + (uint32_t)((kc ^ (row & 7)) << 4);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c=f32 [4,6)=1, a=tf32 [7,10)=2, b=tf32 [10,13)=2,
// a/b K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// kind::f16 with bf16 operands: c=f32 [4,6)=1, a=bf16 [7,10)=1, b=bf16 [10,13)=1, K-major, N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ float tf32_hi(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}

// round-to-nearest tf32 of a FINITE fp32 value as two integer ops (ties away from zero, like cvt.rna): the compiler's
// cvt.rna.tf32.f32 expands to four (it also routes Inf / NaN around the add).  Used where the operand staging is bound by
// instruction issue (gemm_ws.cu, wgrad_ws.cu); the inputs there are activations / gradients of a finite forward pass.
__device__ __forceinline__ float tf32_rn_fast(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u); }
// explicit shared-space accesses (32-bit shared addresses): through a generic pointer carved out of the dynamic shared
// array the compiler emitted LD.E / ST.E (generic, long-scoreboard) in the operand staging loops
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t saddr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// hi_s / lo_s: shared addresses of the hi / lo operand images
template <bool X3>
__device__ __forceinline__ void put_chunk_fast(uint32_t hi_s, uint32_t lo_s, uint32_t off, float4 v) {
    const float4 h = make_float4(tf32_rn_fast(v.x), tf32_rn_fast(v.y), tf32_rn_fast(v.z), tf32_rn_fast(v.w));
    sts128(hi_s + off, h);
    if (X3) sts128(lo_s + off, make_float4(tf32_rn_fast(v.x - h.x), tf32_rn_fast(v.y - h.y), tf32_rn_fast(v.z - h.z), tf32_rn_fast(v.w - h.w)));
}

// write one 16-byte K-chunk (4 reduction elements of one operand row) as hi (and lo) tf32 values
template <bool X3>
__device__ __forceinline__ void put_chunk(uint8_t* hi_base, uint8_t* lo_base, uint32_t off, float4 v) {
    float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
    *reinterpret_cast<float4*>(hi_base + off) = h;
    // lo is rounded to tf32 HERE (round to nearest): left as fp32 the tensor core would drop its low 13 mantissa bits, a
    // truncation whose bias grows linearly with the length of the reduction instead of with its square root
    if (X3)
        *reinterpret_cast<float4*>(lo_base + off) =
            make_float4(tf32_hi(v.x - h.x), tf32_hi(v.y - h.y), tf32_hi(v.z - h.z), tf32_hi(v.w - h.w));
}


}  // namespace tc
}  // namespace bmnas
