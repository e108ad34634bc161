// Raw tcgen05.mma issue/execute rate on sm_100a: one thread per CTA (pair) issues `iters`
// back-to-back MMAs on resident (zero) shared-memory operands, then commits and waits.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/mma_rate tools/mma_rate.cu
//   tools/bin/mma_rate            (prints ns and SM clocks per MMA for a few shapes / layouts)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long mk_desc(unsigned addr, int mn_major) {
    unsigned long long d = 0;
    d |= (unsigned long long)((addr & 0x3FFFFu) >> 4);
    d |= (unsigned long long)(mn_major ? (8192 >> 4) : 1) << 16;
    d |= (unsigned long long)(1024 >> 4) << 32;
    d |= (unsigned long long)1 << 46;
    d |= (unsigned long long)2 << 61;
    return d;
}
__device__ __forceinline__ unsigned mk_idesc(int m, int n, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)a_mn << 15) | ((unsigned)b_mn << 16) |
           ((unsigned)(n >> 3) << 17) | ((unsigned)(m >> 4) << 24);
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

// CG: cta_group (1 or 2).  a_mn / b_mn: operand majors.  n: MMA N.  ksteps: MMAs per "stage"
// (the descriptors advance like in the chain kernels), commit_every: MMAs per tcgen05.commit.
template <int CG>
__global__ void __launch_bounds__(576, 1) rate_kernel(int n, int a_mn, int b_mn, int iters, int commit_every,
                                                      int two_acc, int commit_mode, int spin_mode, long long *out, int fill, const unsigned char *src) {
    extern __shared__ unsigned char smem_raw[];
    const unsigned raw = smem_u32(smem_raw), base = (raw + 1023u) & ~1023u;
    __shared__ unsigned long long bar[16];
    __shared__ unsigned tptr;
    unsigned rank = 0;
    if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    // fill 0: zero operands; 1: random bf16 in [-1, 1) (the tensor pipe's rate is data dependent)
    for (unsigned i = threadIdx.x; i < 192 * 1024 / 16; i += blockDim.x) {
        unsigned w[4] = {0, 0, 0, 0};
        if (fill) {
            unsigned h = i * 2654435761u + blockIdx.x * 40503u + 12345u;
            for (int j = 0; j < 4; ++j) {
                h ^= h << 13; h ^= h >> 17; h ^= h << 5;
                const unsigned lo = 0x3f00u | (h & 0x80ffu), hi = 0x3f00u | ((h >> 16) & 0x80ffu);
                w[j] = lo | (hi << 16);
            }
        }
        reinterpret_cast<uint4 *>(smem_raw + (base - raw))[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    if (threadIdx.x == 0) {
        for (int b = 0; b < 16; ++b)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[b])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        if (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tptr)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tptr)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CG == 2) {
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = tptr;
    if (threadIdx.x == 0 && rank == 0) {
        const unsigned idesc = mk_idesc(128 * CG, n, a_mn, b_mn);
        const unsigned a_step = a_mn ? (2048 >> 4) : (32 >> 4), b_step = b_mn ? (2048 >> 4) : (32 >> 4);
        long long t0, t1, t2;
        unsigned c0, c1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        asm volatile("mov.u32 %0, %%clock;" : "=r"(c0));
        int stage = 0, since = 0, cbar = 0;
        for (int i = 0; i < iters; ++i) {
            const int k = i & 3;
            if (k == 0 && i) stage = stage == 4 ? 0 : stage + 1;
            const unsigned long long da = mk_desc(base + ((i >> 2) & 7) * 16384, a_mn) + (unsigned long long)(a_step * k);
            const unsigned long long db = mk_desc(base + 131072 + stage * 12288, b_mn) + (unsigned long long)(b_step * k);
            const unsigned d = tmem + ((two_acc && (i & 32)) ? 256 : 0);
            if (CG == 2)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"((unsigned)(i != 0)) : "memory");
            else
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"((unsigned)(i != 0)) : "memory");
            if (++since == commit_every) {
                since = 0;
                cbar = cbar == 4 ? 0 : cbar + 1;
                const unsigned cb = smem_u32(&bar[1 + cbar]);
                if (CG == 2 && commit_mode == 0)
                    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                                 ::"r"(cb), "h"((unsigned short)3) : "memory");
                else if (CG == 2 && commit_mode == 1)
                    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                                 ::"r"(cb), "h"((unsigned short)1) : "memory");
                else if (CG == 2)
                    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                                 ::"r"(cb) : "memory");
                else
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                                 ::"r"(cb) : "memory");
            }
        }
        if (CG == 2)
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                         ::"r"(smem_u32(&bar[0])), "h"((unsigned short)1) : "memory");
        else
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                         ::"r"(smem_u32(&bar[0])) : "memory");
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        mbar_wait(smem_u32(&bar[0]), 0);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t2));
        asm volatile("mov.u32 %0, %%clock;" : "=r"(c1));
        out[blockIdx.x * 4 + 0] = t1 - t0;
        out[blockIdx.x * 4 + 1] = t2 - t0;
        out[blockIdx.x * 4 + 2] = (long long)(c1 - c0);
    }
    // spin_mode 1: every other thread of the leader CTA polls the final barrier like the epilogue
    // warps of the chain kernels do (all lanes, mbarrier.try_wait in a loop); 2: one lane per warp
    if (threadIdx.x >= 64 && rank == 0 && (spin_mode == 1 || spin_mode == 2)) {
        if (spin_mode == 1 || (threadIdx.x & 31) == 0) mbar_wait(smem_u32(&bar[0]), 0);
        __syncwarp();
    }
    // spin_mode 3: one thread of EACH CTA streams 16 KB bulk copies global -> the B ring (what the TMA
    // producer of the chain kernels does: 16 KB per 4 MMAs), `depth` copies in flight; 4: the same
    // bytes written by st.shared.v4 from one warp; 5: both CTAs' warps 2-3 only READ shared memory
    if (spin_mode == 3 && threadIdx.x == 64) {
        unsigned ph = 0;
        for (int it2 = 0; it2 < iters / 4; ++it2) {
            const int slot = it2 & 3;
            const unsigned tb = smem_u32(&bar[8 + slot]);
            if (it2 >= 4) { mbar_wait(tb, ph); if (slot == 3) ph ^= 1; }
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tb), "r"(16384u) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(base + 131072u + slot * 16384u), "l"(src + (size_t)(it2 & 63) * 16384), "r"(16384u), "r"(tb) : "memory");
        }
        for (int slot = 0; slot < 4; ++slot) mbar_wait(smem_u32(&bar[8 + slot]), ph);
    }
    if (spin_mode == 4 && threadIdx.x >= 64 && threadIdx.x < 96) {
        const unsigned dst = base + 131072u + (threadIdx.x - 64) * 16;
        for (int it2 = 0; it2 < iters * 8; ++it2)       // 512 B per instruction, 32 instructions = 16 KB per 4 MMAs
            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dst + (it2 & 31) * 512), "r"(it2) : "memory");
    }
    if (spin_mode == 5 && threadIdx.x >= 64 && threadIdx.x < 96) {
        const unsigned srcs = base + 131072u + (threadIdx.x - 64) * 16;
        unsigned acc = 0;
        for (int it2 = 0; it2 < iters * 8; ++it2) {
            unsigned a, b, c, d;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(srcs + (it2 & 31) * 512) : "memory");
            acc += a ^ b ^ c ^ d;
        }
        if (acc == 0x12345u) out[4095] = acc;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CG == 2) {
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

template <int CG>
static void run(int grid, int n, int a_mn, int b_mn, int iters, int commit_every, int two_acc, int commit_mode, const char *what, int spin_mode = 0, int threads = 128, int fill = 0) {
    static unsigned char *src = nullptr;
    if (!src) { cudaMalloc(&src, 64 * 16384); cudaMemset(src, 0x3c, 64 * 16384); }
    long long *out;
    cudaMalloc(&out, 4096 * 8);
    cudaMemset(out, 0, 4096 * 8);
    const int smem = 193 * 1024 + 1024;
    cudaFuncSetAttribute(rate_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    for (int rep = 0; rep < 3; ++rep) cudaLaunchKernelEx(&cfg, rate_kernel<CG>, n, a_mn, b_mn, iters, commit_every, two_acc, commit_mode, spin_mode, out, fill, (const unsigned char *)src);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", what, cudaGetErrorString(e)); exit(1); }
    long long h[4096];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    double issue = 0, total = 0, clk = 0; int cnt = 0;
    for (int b = 0; b < grid; b += CG) { issue += h[b * 4]; total += h[b * 4 + 1]; clk += h[b * 4 + 2]; ++cnt; }
    const double flop = 2.0 * 128 * CG * n * 16;
    printf("%-52s grid %3d  issue %6.1f ns/MMA  complete %6.1f ns/MMA = %6.1f clk  -> %6.1f TFLOP/s per %s, %7.1f chip\n",
           what, grid, issue / cnt / iters, total / cnt / iters, clk / cnt / iters,
           flop / (total / cnt / iters) * 1e-3, CG == 2 ? "pair" : "SM ", flop / (total / cnt / iters) * 1e-3 * (148 / CG));
    cudaFree(out);
}

int main() {
    const int it = 4096;
    for (int grid : {2, 128}) {
        run<2>(grid, 256, 0, 0, it, 4, 1, 0, "cg2 N256 commit/4, no other traffic", 0, 128, 1);
        run<2>(grid, 256, 0, 0, it, 4, 1, 0, "cg2 N256 commit/4 + bulk copies 16 KB / 4 MMAs", 3, 128, 1);
        run<2>(grid, 256, 0, 0, it, 4, 1, 0, "cg2 N256 commit/4 + st.shared 16 KB / 4 MMAs", 4, 128, 1);
        run<2>(grid, 256, 0, 0, it, 4, 1, 0, "cg2 N256 commit/4 + ld.shared 16 KB / 4 MMAs", 5, 128, 1);
        run<2>(grid, 128, 0, 0, it, 4, 1, 0, "cg2 N128 commit/4 + bulk copies", 3, 128, 1);
    }
    return 0;
}
