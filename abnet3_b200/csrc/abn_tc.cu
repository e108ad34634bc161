// abn_tc.cu -- kernel (3), tensor-core path: the embedder's dense contractions
// on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM), fed
// by TMA, with the bias + activation epilogue fused.
//
// Reference behaviour served (paths relative to /root/reference):
//   forward   y = act(x W^T + b)      abnet3/model.py:133-170, :179-186
//   backward  dx = dz W, dW = dz^T x  autograd of the same blocks (abnet3/trainer.py:238)
//
// One kernel, `tc_gemm_kernel`, computes  C[M,N] = A[M,K] . B[N,K]^T  with A and B
// bf16, K-major (K contiguous), fp32 accumulation:
//   forward : A = x        [M, n_in],   B = W    [n_out, n_in]
//   dgrad   : A = dz       [M, n_out],  B = W^T  [n_in, n_out]
//   wgrad   : A = dz^T     [n_out, M],  B = x^T  [n_in, M]      (split over K = M)
// The transposed bf16 operands are written by the producing epilogues (a TMEM
// row-per-thread epilogue writes the transpose coalesced for free).
//
// CTA = 192 threads, one 128 x BN output tile:
//   warp 0      TMA producer: cp.async.bulk.tensor.2d, 128B swizzle, 3-stage mbarrier ring
//   warp 1      MMA issuer: one elected lane issues tcgen05.mma.cta_group::1.kind::f16
//               (M=128, N=BN, K=16) x4 per 64-wide K block, tcgen05.commit frees the stage
//   warps 2-5   epilogue: tcgen05.ld 32x32b (one TMEM lane = one output row per thread),
//               bias + activation (forward) or  x act'(y_below) (dgrad), staged through smem so
//               that the fp32 / bf16 / transposed-bf16 stores, the fp32 atomics (split-K wgrad)
//               and the bias-gradient column sums are all coalesced
// K and N tails are handled by TMA out-of-bounds zero fill; no operand padding
// beyond a leading dimension that is a multiple of 8 elements (16-byte rows).
#include <cuda.h>
#include <cuda_bf16.h>

#include "abn_common.cuh"

namespace abn {

constexpr int TC_BM = 128;        // UMMA_M
constexpr int TC_BK = 64;         // bf16 elements per K block = one 128-byte swizzle row
constexpr int TC_STAGES = 3;        // 96 KB of operand stages: two CTAs per SM
constexpr int TC_THREADS = 192;
constexpr int UMMA_K = 16;

enum { TC_EPI_BIAS_ACT = 0, TC_EPI_STORE = 1, TC_EPI_ATOMIC = 2, TC_EPI_DGRAD_ACT = 3 };

struct TcEpilogue {
    int mode, act;
    const float *bias;
    float *out_f32; long long ld_f32;
    __nv_bfloat16 *out_bf16; long long ld_bf16;
    __nv_bfloat16 *outT_bf16; long long ld_T;
    const __nv_bfloat16 *yprev; long long ld_yprev;   // DGRAD_ACT: forward output of the layer below
    float *db;                                        // column sums of the written values (+=)
};

// ------------------------------------------------------------------- PTX ---
__device__ __forceinline__ unsigned smem_u32_of(const void *p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
// Bounded wait: a protocol bug must trap, not hang the GPU.
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned done = 0;
    for (unsigned spin = 0; spin < (1u << 26); ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap *map, unsigned bar,
                                            int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(map)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(unsigned bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(unsigned d_tmem, unsigned long long a_desc,
                                            unsigned long long b_desc, unsigned idesc,
                                            unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld32(unsigned taddr, unsigned (&r)[32]) {
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
}

// K-major operand tile, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart
// (cute::UMMA::SmemDescriptor: start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version 1 <<46 |
//  layout SWIZZLE_128B (2) <<61)
__device__ __forceinline__ unsigned long long umma_desc_sw128(unsigned smem_addr) {
    unsigned long long d = 0;
    d |= (unsigned long long)((smem_addr & 0x3FFFFu) >> 4);
    d |= (unsigned long long)1 << 16;                 // leading byte offset (unused for SW128 K-major)
    d |= (unsigned long long)(1024 >> 4) << 32;       // stride byte offset: 8 rows x 128 B
    d |= (unsigned long long)1 << 46;                 // descriptor version (Blackwell)
    d |= (unsigned long long)2 << 61;                 // SWIZZLE_128B
    return d;
}

// cute::UMMA::InstrDescriptor for kind::f16: D fp32, A/B bf16, both K-major
__host__ __device__ constexpr unsigned umma_idesc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(n >> 3) << 17) |
           ((unsigned)(m >> 4) << 24);
}

// activation / derivative with the activation id as a compile-time constant;
// sigmoid = ex2 + rcp (two MUFU ops), well inside the bf16 tolerance of this path
template <int ACT>
__device__ __forceinline__ float tc_act(float v) {
    if (ACT == 1) return __fdividef(1.f, 1.f + __expf(-v));
    if (ACT == 2) { const float e = __expf(2.f * v); return __fdividef(e - 1.f, e + 1.f); }
    if (ACT == 3) return v > 0.f ? v : 0.f;
    return v;
}
template <int ACT>
__device__ __forceinline__ float tc_dact(float g, float y) {
    if (ACT == 1) return g * (y * (1.f - y));
    if (ACT == 2) return g * (1.f - y * y);
    if (ACT == 3) return y > 0.f ? g : 0.f;
    return g;
}

// ---------------------------------------------------------------- kernel ---
template <int BN, int ACT>
__global__ void __launch_bounds__(TC_THREADS, 2)
tc_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               int M, int N, int K, int k_blocks_per_split, const TcEpilogue ep) {
    constexpr unsigned A_BYTES = TC_BM * TC_BK * 2;       // 16 KB
    constexpr unsigned B_BYTES = BN * TC_BK * 2;
    constexpr unsigned STAGE_BYTES = A_BYTES + B_BYTES;
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte alignment for the 128B-swizzle atoms
    const unsigned base = (smem_u32_of(smem_raw) + 1023u) & ~1023u;
    const unsigned bars = base + TC_STAGES * STAGE_BYTES;  // full[4], empty[4], tmem_full, tmem_ptr
    const unsigned full0 = bars, empty0 = bars + 8 * TC_STAGES, tfull = bars + 16 * TC_STAGES;
    const unsigned tptr = tfull + 8;
    volatile unsigned *tptr_gen =
        reinterpret_cast<volatile unsigned *>(smem_raw + (tptr - smem_u32_of(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * BN;
    const int total_kb = (K + TC_BK - 1) / TC_BK;
    const int kb_beg = blockIdx.z * k_blocks_per_split;
    const int kb_end = min(total_kb, kb_beg + k_blocks_per_split);
    const int nkb = kb_end - kb_beg;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(tfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {      // TMEM allocation (power of two >= 32 columns), same warp frees it
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(tptr), "n"(BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = *tptr_gen;

    if (nkb > 0) {
        if (warp == 0) {
            if (lane == 0) {
                for (int i = 0; i < nkb; ++i) {
                    const int s = i % TC_STAGES;
                    const unsigned ph = (i / TC_STAGES) & 1;
                    mbar_wait(empty0 + 8 * s, ph ^ 1);
                    mbar_expect_tx(full0 + 8 * s, STAGE_BYTES);
                    const unsigned sa = base + s * STAGE_BYTES;
                    tma_load_2d(sa, &map_a, full0 + 8 * s, (kb_beg + i) * TC_BK, m0);
                    tma_load_2d(sa + A_BYTES, &map_b, full0 + 8 * s, (kb_beg + i) * TC_BK, n0);
                }
            }
        } else if (warp == 1) {
            if (lane == 0) {
                constexpr unsigned idesc = umma_idesc_bf16(TC_BM, BN);
                for (int i = 0; i < nkb; ++i) {
                    const int s = i % TC_STAGES;
                    const unsigned ph = (i / TC_STAGES) & 1;
                    mbar_wait(full0 + 8 * s, ph);
                    tc_fence_after();
                    const unsigned sa = base + s * STAGE_BYTES;
                    const unsigned long long da = umma_desc_sw128(sa);
                    const unsigned long long db = umma_desc_sw128(sa + A_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / UMMA_K; ++k) {
                        // advance 16 bf16 = 32 bytes along K inside the swizzle row: +2 in 16-byte units
                        tc_mma_bf16(tmem, da + 2ull * k, db + 2ull * k, idesc, (i | k) != 0);
                    }
                    tc_commit(empty0 + 8 * s);          // frees the smem stage when the MMAs retire
                }
                tc_commit(tfull);                       // accumulator complete
            }
        } else {
            // epilogue warps 2..5 (128 threads).  Phase 1: TMEM lane group = warp % 4, one
            // output row per thread -> bias/activation -> fp32 tile in smem (the operand stages
            // are dead once the accumulator is complete).  Phase 2: row pass, lanes along
            // columns => every global access is a contiguous 64-128 B segment.  Phase 3: column
            // pass, lanes along rows => the transposed bf16 copy and the bias-gradient column
            // sums are coalesced too.  CS_LD is odd: all three phases are bank-conflict free.
            constexpr int CS_LD = BN + 1;
            float *Cs = reinterpret_cast<float *>(smem_raw + (base - smem_u32_of(smem_raw)));
            const int lg = warp & 3;
            const int we = warp - 2;
            const int rloc = lg * 32 + lane;
            mbar_wait(tfull, 0);
            tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                unsigned r[32];
                tc_ld32(tmem + ((unsigned)(lg * 32) << 16) + (unsigned)c0, r);
                if (ep.mode == TC_EPI_BIAS_ACT) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int col = n0 + c0 + j;
                        const float bv = (ep.bias && col < N) ? __ldg(ep.bias + col) : 0.f;
                        Cs[rloc * CS_LD + c0 + j] = tc_act<ACT>(__uint_as_float(r[j]) + bv);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) Cs[rloc * CS_LD + c0 + j] = __uint_as_float(r[j]);
                }
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            // 4 rows x (BN/32) columns per iteration: all yprev loads are issued before
            // their first use, so their latency is paid once per 4 rows, not once per row
            for (int rb = we; rb < TC_BM; rb += 16) {
                if (m0 + rb >= M) break;
                float yv[4][BN / 32];
                if (ep.mode == TC_EPI_DGRAD_ACT) {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
#pragma unroll
                        for (int i = 0; i < BN / 32; ++i) {
                            const int grow = m0 + rb + 4 * q, gcol = n0 + lane + 32 * i;
                            yv[q][i] = (grow < M && gcol < N)
                                ? __bfloat162float(ep.yprev[(long long)grow * ep.ld_yprev + gcol])
                                : 0.f;
                        }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int rr = rb + 4 * q, grow = m0 + rr;
                    if (grow >= M) break;
#pragma unroll
                    for (int i = 0; i < BN / 32; ++i) {
                        const int c = lane + 32 * i, gcol = n0 + c;
                        if (gcol >= N) continue;
                        float v = Cs[rr * CS_LD + c];
                        if (ep.mode == TC_EPI_DGRAD_ACT) {
                            v = tc_dact<ACT>(v, yv[q][i]);
                            Cs[rr * CS_LD + c] = v;
                        }
                        if (ep.out_f32) {
                            float *dst = ep.out_f32 + (long long)grow * ep.ld_f32 + gcol;
                            if (ep.mode == TC_EPI_ATOMIC) atomicAdd(dst, v);
                            else *dst = v;
                        }
                        if (ep.out_bf16)
                            ep.out_bf16[(long long)grow * ep.ld_bf16 + gcol] = __float2bfloat16_rn(v);
                    }
                }
            }
            if (ep.outT_bf16 || ep.db) {
                asm volatile("bar.sync 1, 128;" ::: "memory");
                for (int c = we; c < BN; c += 4) {
                    const int gcol = n0 + c;
                    if (gcol >= N) break;
                    float sum = 0.f;
#pragma unroll
                    for (int i = 0; i < TC_BM / 32; ++i) {
                        const int rr = lane + 32 * i, grow = m0 + rr;
                        if (grow < M) {
                            const float v = Cs[rr * CS_LD + c];
                            sum += v;
                            if (ep.outT_bf16)
                                ep.outT_bf16[(long long)gcol * ep.ld_T + grow] = __float2bfloat16_rn(v);
                        }
                    }
                    if (ep.db) {
                        sum = warp_sum(sum);
                        if (lane == 0) atomicAdd(ep.db + gcol, sum);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(BN)
                     : "memory");
    }
}

// ------------------------------------------------ elementwise companions ---
// fp32 [rows, cols] -> bf16 [rows, ld_dst] and/or transposed bf16 [cols, ld_T]
__global__ void cast_bf16_kernel(const float *__restrict__ src, long long rows, int cols,
                                 long long ld_src, __nv_bfloat16 *__restrict__ dst,
                                 long long ld_dst, __nv_bfloat16 *__restrict__ dstT,
                                 long long ld_T) {
    __shared__ float tile[32][33];
    const long long r0 = (long long)blockIdx.y * 32;
    const int c0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
    for (int i = ty; i < 32; i += 8) {
        const long long r = r0 + i;
        const int c = c0 + tx;
        float v = 0.f;
        if (r < rows && c < cols) {
            v = src[r * ld_src + c];
            if (dst) dst[r * ld_dst + c] = __float2bfloat16_rn(v);
        }
        tile[i][tx] = v;
    }
    if (!dstT) return;
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i;
        const long long r = r0 + tx;
        if (c < cols && r < rows) dstT[(long long)c * ld_T + r] = __float2bfloat16_rn(tile[tx][i]);
    }
}

// dz = dy * act'(y): writes dz bf16 [m, ld], dz^T bf16 [n, ldT], db (+)= column sums
__global__ void act_backward_bf16_kernel(const float *__restrict__ y, const float *__restrict__ dy,
                                         long long m, int n, int act,
                                         __nv_bfloat16 *__restrict__ dz, long long ld,
                                         __nv_bfloat16 *__restrict__ dzT, long long ldT,
                                         float *__restrict__ db) {
    __shared__ float tile[32][33];
    __shared__ float colsum[32];
    const long long r0 = (long long)blockIdx.y * 32;
    const int c0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    if (threadIdx.x < 32) colsum[threadIdx.x] = 0.f;
    __syncthreads();
    float part = 0.f;
    for (int i = ty; i < 32; i += 8) {
        const long long r = r0 + i;
        const int c = c0 + tx;
        float g = 0.f;
        if (r < m && c < n) {
            const float yy = y[r * n + c];
            g = dy[r * n + c];
            switch (act) {
                case 1: g *= yy * (1.f - yy); break;
                case 2: g *= 1.f - yy * yy; break;
                case 3: g = yy > 0.f ? g : 0.f; break;
                default: break;
            }
            if (dz) dz[r * ld + c] = __float2bfloat16_rn(g);
        }
        tile[i][tx] = g;
        part += g;
    }
    if (db) atomicAdd(&colsum[tx], part);
    __syncthreads();
    if (db && threadIdx.x < 32 && c0 + threadIdx.x < n) atomicAdd(db + c0 + threadIdx.x, colsum[threadIdx.x]);
    if (!dzT) return;
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i;
        const long long r = r0 + tx;
        if (c < n && r < m) dzT[(long long)c * ldT + r] = __float2bfloat16_rn(tile[tx][i]);
    }
}

// ------------------------------------------------------------------ host ---
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
                cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D bf16 tensor [rows, cols] with leading dimension ld (elements); box = 64 cols x box_rows
static int make_map(CUtensorMap *map, const void *ptr, long long rows, long long cols,
                    long long ld, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return set_error(ABN_EIO, "cuTensorMapEncodeTiled is not available");
    if ((reinterpret_cast<uintptr_t>(ptr) & 15u) || (ld & 7))
        return set_error(ABN_EINVAL, "bf16 operand must be 16-byte aligned with ld %% 8 == 0");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), dims, strides,
                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(ABN_EIO, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return ABN_OK;
}

template <int BN, int ACT>
static int launch_tc(const CUtensorMap &ma, const CUtensorMap &mb, int M, int N, int K, int split_k,
                     const TcEpilogue &ep, cudaStream_t st) {
    constexpr unsigned smem = TC_STAGES * (TC_BM * TC_BK * 2 + BN * TC_BK * 2) + 1024 + 256;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(tc_gemm_kernel<BN, ACT>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem) != cudaSuccess)
            return set_error(ABN_EIO, "tc_gemm: cannot reserve %u bytes of shared memory", smem);
        configured = true;
    }
    const int total_kb = (K + TC_BK - 1) / TC_BK;
    if (split_k < 1) split_k = 1;
    if (split_k > total_kb) split_k = total_kb;
    const int per = (total_kb + split_k - 1) / split_k;
    split_k = (total_kb + per - 1) / per;
    dim3 grid((N + BN - 1) / BN, (M + TC_BM - 1) / TC_BM, split_k);
    tc_gemm_kernel<BN, ACT><<<grid, TC_THREADS, smem, st>>>(ma, mb, M, N, K, per, ep);
    return check_launch("abn_gemm_bf16_tn");
}

}  // namespace abn

using namespace abn;

extern "C" int abn_gemm_bf16_tn(const void *A, int64_t lda, const void *B, int64_t ldb, int M, int N,
                                int K, int epilogue, const float *bias, int act, float *out_f32,
                                int64_t ld_f32, void *out_bf16, int64_t ld_bf16, void *outT_bf16,
                                int64_t ld_T, const void *yprev, int64_t ld_yprev, float *db,
                                int split_k, abn_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    if (M == 0 || N == 0) return ABN_OK;
    if (!A || !B || M < 0 || N < 0 || K <= 0 || epilogue < 0 || epilogue > 3 || act < 0 || act > 3)
        return set_error(ABN_EINVAL, "abn_gemm_bf16_tn: bad argument");
    if (epilogue == TC_EPI_DGRAD_ACT && !yprev)
        return set_error(ABN_EINVAL, "abn_gemm_bf16_tn: dgrad epilogue needs yprev");
    if (epilogue == TC_EPI_ATOMIC && !out_f32)
        return set_error(ABN_EINVAL, "abn_gemm_bf16_tn: atomic epilogue needs out_f32");
    if (epilogue != TC_EPI_ATOMIC) split_k = 1;
    const int bn = N <= 64 ? 64 : 128;
    CUtensorMap ma, mb;
    if (int rc = make_map(&ma, A, M, K, lda, TC_BM)) return rc;
    if (int rc = make_map(&mb, B, N, K, ldb, bn)) return rc;
    TcEpilogue ep;
    ep.mode = epilogue; ep.act = act; ep.bias = bias;
    ep.out_f32 = out_f32; ep.ld_f32 = ld_f32;
    ep.out_bf16 = static_cast<__nv_bfloat16 *>(out_bf16); ep.ld_bf16 = ld_bf16;
    ep.outT_bf16 = static_cast<__nv_bfloat16 *>(outT_bf16); ep.ld_T = ld_T;
    ep.yprev = static_cast<const __nv_bfloat16 *>(yprev); ep.ld_yprev = ld_yprev; ep.db = db;
    cudaStream_t st = (cudaStream_t)stream;
    const int a = (epilogue == TC_EPI_BIAS_ACT || epilogue == TC_EPI_DGRAD_ACT) ? act : 0;
#define ABN_TC(BN_, ACT_) return launch_tc<BN_, ACT_>(ma, mb, M, N, K, split_k, ep, st)
    if (bn == 64) {
        switch (a) { case 1: ABN_TC(64, 1); case 2: ABN_TC(64, 2); case 3: ABN_TC(64, 3);
                     default: ABN_TC(64, 0); }
    }
    switch (a) { case 1: ABN_TC(128, 1); case 2: ABN_TC(128, 2); case 3: ABN_TC(128, 3);
                 default: ABN_TC(128, 0); }
#undef ABN_TC
}

extern "C" int abn_cast_bf16(const float *src, int64_t rows, int cols, int64_t ld_src, void *dst,
                             int64_t ld_dst, void *dstT, int64_t ld_T, abn_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    if (rows == 0 || cols == 0) return ABN_OK;
    if (!src || (!dst && !dstT) || rows < 0 || cols < 0)
        return set_error(ABN_EINVAL, "abn_cast_bf16: bad argument");
    dim3 grid((cols + 31) / 32, (unsigned)((rows + 31) / 32));
    cast_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        src, rows, cols, ld_src, static_cast<__nv_bfloat16 *>(dst), ld_dst,
        static_cast<__nv_bfloat16 *>(dstT), ld_T);
    return check_launch("abn_cast_bf16");
}

extern "C" int abn_act_backward_bf16(const float *y, const float *dy, int64_t m, int n, int act,
                                     void *dz, int64_t ld, void *dzT, int64_t ldT, float *db,
                                     abn_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    if (m == 0 || n == 0) return ABN_OK;
    if (!y || !dy || m < 0 || n < 0 || act < 0 || act > 3)
        return set_error(ABN_EINVAL, "abn_act_backward_bf16: bad argument");
    dim3 grid((n + 31) / 32, (unsigned)((m + 31) / 32));
    act_backward_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        y, dy, m, n, act, static_cast<__nv_bfloat16 *>(dz), ld, static_cast<__nv_bfloat16 *>(dzT),
        ldT, db);
    return check_launch("abn_act_backward_bf16");
}
