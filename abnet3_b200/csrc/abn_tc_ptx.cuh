// abn_tc_ptx.cuh -- tcgen05 / TMEM / TMA / mbarrier building blocks shared by the tensor-core
// kernels (abn_tc2.cu: persistent grouped GEMM; abn_tc3.cu: forward pass with the activations
// resident in shared memory).  Inline PTX for sm_100a.
#pragma once
#include <cuda.h>
#include <stdlib.h>
#include <cuda_bf16.h>

#include "abn_common.cuh"

namespace abn {

constexpr int G_BM = 128;          // UMMA_M
constexpr int G_BK = 64;           // bf16 elements per K block (128-byte swizzle row / 64 K rows)
constexpr int G_UK = 16;           // UMMA_K
constexpr int G_EPI_WARPS = 8;     // two per TMEM lane quarter: each takes every other 32-column chunk
constexpr int G_THREADS = 64 + 32 * G_EPI_WARPS;
constexpr int G_OBUF = 2;          // output boxes in flight per epilogue warp (TMA store latency ~1 us)



// ------------------------------------------------------------------- PTX ---
__device__ __forceinline__ unsigned g_smem_u32(const void *p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void g_mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void g_mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void g_mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug must trap, not hang the GPU.
__device__ __forceinline__ void g_mbar_wait(unsigned bar, unsigned parity) {
    unsigned done = 0;
    for (unsigned spin = 0; spin < (1u << 27); ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void g_tma_2d(unsigned dst, const CUtensorMap *map, unsigned bar, int c0,
                                         int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(map)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void g_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void g_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void g_commit(unsigned bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 ::"r"(bar) : "memory");
}
__device__ __forceinline__ void g_mma(unsigned d_tmem, unsigned long long a_desc,
                                      unsigned long long b_desc, unsigned idesc, unsigned acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void g_ld32(unsigned taddr, float (&v)[32]) {
    unsigned r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
          "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}

__device__ __forceinline__ void g_ld16(unsigned taddr, float (&v)[16]) {
    unsigned r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}

// 64 consecutive accumulator columns of this warp's 32 TMEM lanes in one request
__device__ __forceinline__ void g_ld64(unsigned taddr, float (&v)[64]) {
    unsigned r[64];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr) : "memory");
#pragma unroll
    for (int j = 0; j < 64; ++j) v[j] = __uint_as_float(r[j]);
}

// cute::UMMA::SmemDescriptor, 128-byte swizzle.
//   K-major  tile [rows x 64 K]: rows of 128 B, 8-row groups 1024 B apart (SBO); a UMMA_K
//            step of 16 elements advances the start address by 32 B
//   MN-major tile [64 K rows x R]: per 64-element column block, 64 rows of 128 B; 8-row
//            K groups 1024 B apart (SBO), column blocks 8192 B apart (LBO); a UMMA_K step
//            of 16 rows advances the start address by 2048 B
__device__ __forceinline__ unsigned long long g_desc(unsigned smem_addr, int mn_major) {
    unsigned long long d = 0;
    d |= (unsigned long long)((smem_addr & 0x3FFFFu) >> 4);
    d |= (unsigned long long)(mn_major ? (8192 >> 4) : 1) << 16;     // leading byte offset
    d |= (unsigned long long)(1024 >> 4) << 32;                      // stride byte offset
    d |= (unsigned long long)1 << 46;                                // descriptor version (Blackwell)
    d |= (unsigned long long)2 << 61;                                // SWIZZLE_128B
    return d;
}
// cute::UMMA::InstrDescriptor for kind::f16: D fp32, A/B bf16, majors per operand
__device__ __forceinline__ unsigned g_idesc(int m, int n, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)a_mn << 15) | ((unsigned)b_mn << 16) |
           ((unsigned)(n >> 3) << 17) | ((unsigned)(m >> 4) << 24);
}

__device__ __forceinline__ unsigned g_pack_bf16(float lo, float hi);
__device__ __forceinline__ float g_tanh(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// one MUFU op per element: sigmoid(v) = 0.5 + 0.5 tanh(v / 2); well inside the bf16 tolerance
template <int ACT>
__device__ __forceinline__ float g_act(float v) {
    if (ACT == 1) return fmaf(0.5f, g_tanh(0.5f * v), 0.5f);
    if (ACT == 2) return g_tanh(v);
    if (ACT == 3) return v > 0.f ? v : 0.f;
    return v;
}
template <int ACT>
__device__ __forceinline__ float g_dact(float g, float y) {
    if (ACT == 1) return g * (y * (1.f - y));
    if (ACT == 2) return g * (1.f - y * y);
    if (ACT == 3) return y > 0.f ? g : 0.f;
    return g;
}
// bf16 outputs of sigmoid / tanh layers: two activations per MUFU op (tanh.approx.bf16x2);
// the argument is rounded to bf16 first, an error of the order of the output rounding
__device__ __forceinline__ unsigned g_tanh_bf16x2(unsigned x) {
    unsigned y;
    asm("tanh.approx.bf16x2 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}
template <int ACT>      // ACT 1: sigmoid, 2: tanh;  out[16] = packed bf16 pairs of act(v + bias)
__device__ __forceinline__ void g_bias_act32_packed(const float (&v)[32], const float *bs,
                                                    unsigned (&out)[16]) {
    const float4 *b4 = reinterpret_cast<const float4 *>(bs);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float4 b = b4[q];
        const float sc = ACT == 1 ? 0.5f : 1.f;
        const unsigned p0 = g_pack_bf16(sc * (v[4 * q] + b.x), sc * (v[4 * q + 1] + b.y));
        const unsigned p1 = g_pack_bf16(sc * (v[4 * q + 2] + b.z), sc * (v[4 * q + 3] + b.w));
        unsigned t0 = g_tanh_bf16x2(p0), t1 = g_tanh_bf16x2(p1);
        if (ACT == 1) {         // 0.5 t + 0.5 on both halves (bf16 0.5 = 0x3f00)
            asm("fma.rn.bf16x2 %0, %1, %2, %2;" : "=r"(t0) : "r"(t0), "r"(0x3f003f00u));
            asm("fma.rn.bf16x2 %0, %1, %2, %2;" : "=r"(t1) : "r"(t1), "r"(0x3f003f00u));
        }
        out[2 * q] = t0;
        out[2 * q + 1] = t1;
    }
}
template <int ACT>
__device__ __forceinline__ void g_bias_act32(float (&v)[32], const float *bs) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = g_act<ACT>(v[j] + bs[j]);
}
template <int ACT>
__device__ __forceinline__ void g_dact32(float (&v)[32], const uint4 (&y)[4]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const unsigned w[4] = {y[q].x, y[q].y, y[q].z, y[q].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            v[8 * q + 2 * e] = g_dact<ACT>(v[8 * q + 2 * e], __uint_as_float(w[e] << 16));
            v[8 * q + 2 * e + 1] = g_dact<ACT>(v[8 * q + 2 * e + 1], __uint_as_float(w[e] & 0xffff0000u));
        }
    }
}
__device__ __forceinline__ unsigned g_pack_bf16(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<unsigned *>(&h);
}

// ---- CTA-pair (cta_group::2) helpers: the two CTAs of a cluster share every B tile (each
// loads half of its columns), one thread of the leader CTA issues the 256-row MMAs
constexpr unsigned G_PEER_MASK = 0xFEFFFFFFu;      // shared-window address of the same object in CTA 0 of the pair
__device__ __forceinline__ unsigned g_cluster_rank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void g_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void g_tma_2d_pair(unsigned dst, const CUtensorMap *map, unsigned bar_cta0,
                                              int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(map)), "r"(bar_cta0), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void g_mma_pair(unsigned d_tmem, unsigned long long a_desc,
                                           unsigned long long b_desc, unsigned idesc, unsigned acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
// ---- the MMA issuer as a CONVERGENT warp --------------------------------------------------
// tcgen05.mma / tcgen05.commit take their descriptors from uniform registers.  Issued from a
// single-lane branch (`if (lane == 0)`) the operands live in per-thread registers and every
// instruction is wrapped in an ELECT / 7 x R2UR.BROADCAST / BRA.U.ANY loop: ~150 clocks of issue
// per 128-clock MMA (tools/mma_rate.cu), i.e. the tensor pipe idles half of the time.  Run by
// the whole warp on warp-uniform values, with elect.sync inside the asm block, the operands stay
// in uniform registers and one MMA costs a handful of issue slots.
__device__ __forceinline__ void g_mma_pair_warp(unsigned d_tmem, unsigned long long a_desc,
                                                unsigned long long b_desc, unsigned idesc, unsigned acc) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void g_tma_2d_pair_warp(unsigned dst, const CUtensorMap *map, unsigned bar_cta0,
                                                   int c0, int c1) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4}], [%2];\n\t}"
        ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(map)), "r"(bar_cta0), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void g_tma_2d_warp(unsigned dst, const CUtensorMap *map, unsigned bar, int c0, int c1) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4}], [%2];\n\t}"
        ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(map)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void g_mbar_expect_tx_warp(unsigned bar, unsigned bytes) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void g_mma_warp(unsigned d_tmem, unsigned long long a_desc,
                                           unsigned long long b_desc, unsigned idesc, unsigned acc) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void g_commit_warp(unsigned bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(bar) : "memory");
}
__device__ __forceinline__ void g_commit_pair_warp(unsigned bar) {      // arrives on `bar` in BOTH CTAs
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
        "[%0], %1;\n\t}" ::"r"(bar), "h"((unsigned short)3) : "memory");
}
// all 32 lanes poll; the loop lives inside the asm block, so the compiler sees no divergence
// (bounded like g_mbar_wait: a protocol bug traps)
__device__ __forceinline__ void g_mbar_wait_warp(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .u32 c;\n\t"
        "mov.u32 c, 0;\n"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "add.u32 c, c, 1;\n\t"
        "setp.lt.u32 p, c, 0x8000000;\n\t"
        "@p bra WAIT_LOOP;\n\t"
        "trap;\n"
        "WAIT_DONE:\n\t}"
        ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void g_mbar_wait_cluster_warp(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .u32 c;\n\t"
        "mov.u32 c, 0;\n"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "add.u32 c, c, 1;\n\t"
        "setp.lt.u32 p, c, 0x8000000;\n\t"
        "@p bra WAIT_LOOP;\n\t"
        "trap;\n"
        "WAIT_DONE:\n\t}"
        ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void g_commit_pair(unsigned bar) {      // arrives on `bar` in BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
                 "[%0], %1;" ::"r"(bar), "h"((unsigned short)3) : "memory");
}
__device__ __forceinline__ void g_mbar_arrive_cta0(unsigned bar) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];"
                 ::"r"(bar & G_PEER_MASK) : "memory");
}

__device__ __forceinline__ void g_tma_store_2d(const CUtensorMap *map, unsigned src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<unsigned long long>(map)), "r"(src), "r"(c0), "r"(c1) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void g_store_wait_read0() {
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void g_store_wait_all() {
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}


// ---- tensor maps (host) ---------------------------------------------------
typedef CUresult (*GEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                              const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                              const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline GEncodeFn g_encode_fn() {
    static GEncodeFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
                cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<GEncodeFn>(p);
    }
    return fn;
}

// bf16 array [rows, cols] with leading dimension ld (elements), cols contiguous;
// box = box_cols (inner) x box_rows, 128-byte swizzle, out-of-bounds elements read as 0
static inline int g_make_map(CUtensorMap *map, const void *ptr, long long rows, long long cols,
                             long long ld, int box_cols, int box_rows,
                             CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
    GEncodeFn fn = g_encode_fn();
    if (!fn) return set_error(ABN_EIO, "cuTensorMapEncodeTiled is not available");
    if ((reinterpret_cast<uintptr_t>(ptr) & 15u) || (ld & 7))
        return set_error(ABN_EINVAL, "bf16 operand must be 16-byte aligned with ld %% 8 == 0");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), dims, strides,
                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(ABN_EIO, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return ABN_OK;
}

static inline bool g_use_pdl() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("ABN_GEMM_NO_PDL"); v = (e && e[0] == '1') ? 0 : 1; }
    return v == 1;
}

}  // namespace abn
