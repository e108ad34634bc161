// abn_tc2.cu -- kernel (3), tensor-core path: every dense contraction of the embedder's
// training step on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM),
// fed by TMA, as ONE persistent kernel that walks a list of GEMM problems.
//
// Reference behaviour served (paths relative to /root/reference):
//   forward   y = act(x W^T + b)        abnet3/model.py:133-170, :179-186
//   backward  dz_below = (dz W) * act'(y_below),  dW = dz^T x,  db = colsum(dz)
//             -- the autograd of those blocks (abnet3/trainer.py:238)
//
// All operands are bf16 arrays in their NATURAL row-major layout -- activations and dz
// [rows, features], weights [n_out, n_in] (nn.Linear) -- no transposed copies exist:
//   forward  C[rows, n_out] = x[rows, n_in] . W[n_out, n_in]^T     A K-major,  B K-major
//   dgrad    C[rows, n_in]  = dz[rows, n_out] . W[n_out, n_in]     A K-major,  B MN-major
//   wgrad    C[n_out, n_in] = dz[rows, n_out]^T . x[rows, n_in]    A MN-major, B MN-major
// "MN-major" = the operand's M (or N) index is the contiguous one in memory; the UMMA
// shared-memory descriptor and instruction descriptor express it (cute::UMMA::Major::MN),
// TMA delivers it as 64-element x 64-row boxes with the 128-byte swizzle.
// db comes for free: activations carry a column of ones right after their last feature
// (inside the 8-element row padding), so column n_in of dz^T [x | 1] is colsum(dz).
//
// CTA = 320 threads, persistent over 128 x BN output tiles (x split-K):
//   warp 0      TMA producer (cp.async.bulk.tensor.2d, mbarrier ring of smem stages)
//   warp 1      MMA issuer: one lane, tcgen05.mma.cta_group::1.kind::f16 M=128 N<=256 K=16
//   warps 2-9   epilogue: tcgen05.ld (one TMEM lane = one output row per thread, two warps
//               per lane quarter taking alternate 32-column chunks) -> bias + activation |
//               x act'(y_below) -> packed bf16 / fp32 rows, 16-byte stores; or (wgrad)
//               16-byte vector fp32 reds straight from the registers
// The accumulator is double buffered in TMEM (2 x BN columns): the epilogue of tile i
// overlaps the MMAs of tile i+1.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_bf16.h>

#include "abn_tc_ptx.cuh"

namespace abn {

constexpr int G_MAXP = ABN_GEMM_MAX_GROUP;

enum { GE_BIAS_ACT = 0, GE_DACT = 1, GE_ATOMIC = 2 };

struct GProblem {
    CUtensorMap map_a, map_b;
    CUtensorMap map_c, map_y;            // bf16 output / y_below [M, ld]: 64 x 32 boxes, 128-byte swizzle
    int M, N, K;                    // N includes the ones column of a wgrad problem
    int a_mn, b_mn;
    int epi, act, out_f32, ones_col, tma_store;
    int n_cap;                      // N + ones_col: the columns a tile row covers
    int tiles_m, tiles_n, splits, kb_per_split;
    int tile_beg;
    const float *bias;
    void *out; long long ldo;
    const __nv_bfloat16 *yprev; long long ld_yprev;
    float *ones_out;
    int *signal;                    // [tiles_m]: +1 per CTA when a tile's output rows are in global memory
    const int *wait; int wait_count; // A rows of m-tile mt may be loaded once wait[mt] >= wait_count
};
struct GGroup {
    GProblem p[G_MAXP];
    int n_problems, total_tiles;
    int rot;                        // tile of CTA pair q in wave k: k * pairs + (q + rot * k) % pairs
    long long *trace;               // debug: per-CTA role timestamps (NULL in production)
#ifdef ABN_TC_DEBUG
    int dbg;                        // experiments (tools/build_dbg.sh): 1 no epilogue work, 2 no TMA
                                    // stores, 4 no tcgen05.ld, 8 no tcgen05.mma
#endif
};
#ifdef ABN_TC_DEBUG
#define G_DBG(g, bit) (((g).dbg & (bit)) != 0)
#else
#define G_DBG(g, bit) false
#endif

struct GTile { int pi, m0, n0, kb0, nkb, n_eff, mt; };

__device__ __forceinline__ void g_trace(const GGroup &g, unsigned it, int slot) {
    if (g.trace && it < 8) {
        long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        g.trace[((size_t)blockIdx.x * 8 + it) * 16 + slot] = t;
    }
}

__device__ __forceinline__ GTile g_decode(const GGroup &g, int tile, int bn, int ncta = 1, int rank = 0) {
    int pi = 0;
#pragma unroll
    for (int q = 1; q < G_MAXP; ++q)
        if (q < g.n_problems && tile >= g.p[q].tile_beg) pi = q;
    const GProblem &P = g.p[pi];
    const int local = tile - P.tile_beg;
    const int per_split = P.tiles_m * P.tiles_n;
    const int ks = local / per_split, r = local - ks * per_split;
    const int mt = r / P.tiles_n, nt = r - mt * P.tiles_n;
    GTile t;
    t.pi = pi;
    t.m0 = (mt * ncta + rank) * G_BM;
    t.mt = mt;
    t.n0 = nt * bn;
    const int total_kb = (P.K + G_BK - 1) / G_BK;
    t.kb0 = ks * P.kb_per_split;
    t.nkb = min(total_kb, t.kb0 + P.kb_per_split) - t.kb0;
    const int rem = P.n_cap - t.n0;
    t.n_eff = rem >= bn ? bn : ((rem + 15) & ~15);
    return t;
}

// Wait until `count` tiles have published their rows of a 256-row block, then order this
// thread's TMA reads after the observation.
__device__ __forceinline__ void g_wait_block(const int *ctr, int count) {
    unsigned spins = 0;
    int seen;
    do {
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory");
        if (++spins > (1u << 28)) __trap();
    } while (seen < count);
    asm volatile("fence.proxy.async;" ::: "memory");
}

// Static schedule: wave k hands position (q + rot k) mod pairs to CTA pair q.  With chained
// problems the positions whose dependency sits in the wave just before stall for an epilogue;
// rot != 0 rotates that bad luck over the pairs instead of charging the same ones every wave.
__device__ __forceinline__ int g_tile_of(const GGroup &g, int q, int pairs, unsigned k) {
    return (int)k * pairs + (q + g.rot * (int)k) % pairs;
}

// One warp's share of a bf16-output tile (forward: bias + activation; dgrad: x act'(y_below)).
// Everything the chunk loop needs sits in registers; EPI / ACT are compile-time so that the
// loop body is straight-line code (the problem descriptor lives in constant memory: reading
// it field by field inside the loop costs a dependent LDC + branch per decision).
struct GEpi {
    const CUtensorMap *map_c, *map_y;
    const float *bs;                // staged bias of this tile (forward)
    unsigned taddr, my_out, my_y, ybar;
    int n0, row0, n_eff, n_cap, N, ones_col, c_first, lane;
    int dbg;
};
template <int EPI, int ACT, int BN>
__device__ __forceinline__ void g_epi_bf16_tile(const GEpi &e, unsigned &nbox, unsigned &ycount) {
    // A warp works in steps of 64 columns: its 32 x 64 bf16 box has 128-byte rows (128-byte
    // swizzle) -- half as many, twice as large TMA requests as 32-column boxes, which matter
    // because the stores share the TMA path with the operand loads of the next tile.
    const int lane = e.lane;
    const unsigned swz = (unsigned)(lane & 7);          // 128-byte swizzle of row `lane`
#pragma unroll 1
    for (int c0 = e.c_first; c0 < BN; c0 += 128) {
        const int gcol0 = e.n0 + c0;
        if (gcol0 >= e.n_cap || c0 >= e.n_eff + 16) break;     // warp-uniform
        const unsigned sbuf = e.my_out + (nbox & 1u) * 4096u;
        uint4 yc[8];
        if (EPI == GE_DACT) {
            // this warp's 32 x 64 box of y_below arrived in the SAME buffer its outputs will
            // use (read it into registers first); the next step's box goes to the other
            // buffer once the store that last used it has been read out
            g_mbar_wait(e.ybar, ycount & 1u);
            ++ycount;
#pragma unroll
            for (int q = 0; q < 8; ++q)
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(yc[q].x), "=r"(yc[q].y), "=r"(yc[q].z), "=r"(yc[q].w)
                             : "r"(sbuf + lane * 128 + (((unsigned)q ^ swz) << 4)) : "memory");
            __syncwarp();
            const int gn = gcol0 + 128;
            if (lane == 0 && gn < e.n_cap && c0 + 128 < e.n_eff + 16 && c0 + 128 < BN) {
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                g_mbar_expect_tx(e.ybar, 4096u);
                g_tma_2d(e.my_out + ((nbox + 1) & 1u) * 4096u, e.map_y, e.ybar, gn, e.row0);
            }
        } else {
            // the store issued two steps ago must have left this buffer
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
        }
#pragma unroll
        for (int hseg = 0; hseg < 2; ++hseg) {             // the two 32-column halves of the step
            const int cc = c0 + 32 * hseg, gc = gcol0 + 32 * hseg;
            float v[32];
            if (G_DBG(e, 4)) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = (float)(cc + j);
            } else {
                g_ld32(e.taddr + cc, v);
            }
            unsigned pk[16];
            if (EPI == GE_BIAS_ACT && (ACT == 1 || ACT == 2)) {
                g_bias_act32_packed<ACT>(v, e.bs + cc, pk);
            } else {
                if (EPI == GE_BIAS_ACT) {
                    g_bias_act32<ACT>(v, e.bs + cc);
                } else {
                    const uint4 yh[4] = {yc[4 * hseg], yc[4 * hseg + 1], yc[4 * hseg + 2], yc[4 * hseg + 3]};
                    g_dact32<ACT>(v, yh);
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) pk[j] = g_pack_bf16(v[2 * j], v[2 * j + 1]);
            }
            if (e.ones_col && e.N >= gc && e.N < gc + 32) {
                const int jo = e.N - gc;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    if (2 * j == jo) pk[j] = (pk[j] & 0xffff0000u) | 0x3f80u;
                    if (2 * j + 1 == jo) pk[j] = (pk[j] & 0x0000ffffu) | 0x3f800000u;
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                             ::"r"(sbuf + lane * 128 + (((unsigned)(4 * hseg + q) ^ swz) << 4)),
                               "r"(pk[4 * q]), "r"(pk[4 * q + 1]), "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3])
                             : "memory");
        }
        // bf16 rows -> one TMA store of the box (full-line writes, clipped at M rows / ldo
        // columns by the tensor map)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0 && !G_DBG(e, 2)) g_tma_store_2d(e.map_c, sbuf, gcol0, e.row0);
        ++nbox;
    }
}

// ---------------------------------------------------------------- kernel ---
// WG: every problem of the group is a reduction (wgrad: fp32 reds straight from the registers) --
// the 64 KB of output staging become two more operand stages.  A wgrad tile is a long K loop
// over operands that both stream from L2 (32 KB per k-block and CTA): with 4 stages in flight the
// ring is latency bound (112 ns per MMA, tools/trace_tc2.py), with 6 it is not.
template <int BN, int NCTA, bool WG>
constexpr int g_stages() {
    return (G_BM * G_BK * 2 + (BN / NCTA) * G_BK * 2) > 32768 ? (WG ? 4 : 3) : (WG ? 6 : 4);
}
template <int BN, int NCTA, bool WG>
__global__ void __launch_bounds__(G_THREADS, 1)
tc_group_kernel(const __grid_constant__ GGroup g) {
    constexpr unsigned A_BYTES = G_BM * G_BK * 2;           // 16 KB: this CTA's 128 rows
    constexpr unsigned B_BYTES = (BN / NCTA) * G_BK * 2;    // this CTA's share of the B tile
    constexpr unsigned STAGE = A_BYTES + B_BYTES;
    constexpr int STAGES = g_stages<BN, NCTA, WG>();
    // per epilogue warp: two 32 x 64 bf16 boxes staged for TMA stores (y_below boxes land in them too)
    constexpr unsigned OUT_BYTES = WG ? 0u : G_EPI_WARPS * 2u * 4096u;
    const int rank = NCTA == 2 ? (int)g_cluster_rank() : 0;
    const int tile0 = blockIdx.x / NCTA, tile_step = gridDim.x / NCTA;
    extern __shared__ unsigned char smem_raw[];
    const unsigned raw = g_smem_u32(smem_raw);
    const unsigned base = (raw + 1023u) & ~1023u;           // 128B-swizzle atoms are 1024-byte aligned
    unsigned char *gen = smem_raw + (base - raw);
    const unsigned outs = base + STAGES * STAGE;            // 1024-byte aligned (stages are multiples of 16 KB)
    const unsigned bars = outs + OUT_BYTES;
    const unsigned full0 = bars, empty0 = bars + 8 * STAGES;
    const unsigned tfull0 = bars + 16 * STAGES, tempty0 = tfull0 + 16, tptr = tempty0 + 16;
    const unsigned ybar0 = bars + 16 * STAGES + 64;          // one mbarrier per epilogue warp
    volatile unsigned *tptr_gen = reinterpret_cast<volatile unsigned *>(
        gen + STAGES * STAGE + OUT_BYTES + 16 * STAGES + 32);
    float *bias_s = reinterpret_cast<float *>(gen + STAGES * STAGE + OUT_BYTES + 256);      // [2][BN]

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // (provably warp-uniform)
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { g_mbar_init(full0 + 8 * s, 1); g_mbar_init(empty0 + 8 * s, 1); }
        for (int w = 0; w < G_EPI_WARPS; ++w) g_mbar_init(ybar0 + 8 * w, 1);
        for (int b = 0; b < 2; ++b) {
            g_mbar_init(tfull0 + 8 * b, 1);
            g_mbar_init(tempty0 + 8 * b, NCTA * G_EPI_WARPS);      // the pair's epilogues report to the leader
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {      // the whole TMEM of the SM is ours (one CTA per SM): two accumulators
        if (NCTA == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                         ::"r"(tptr), "n"(2 * BN) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                         ::"r"(tptr), "n"(2 * BN) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    g_fence_before();
    if (NCTA == 2) g_cluster_sync(); else __syncthreads();
    g_fence_after();
    const unsigned tmem = *tptr_gen;
    // programmatic dependent launch: everything above overlapped the predecessor's tail; its
    // results are needed from here on, and our own successor may start its prologue now
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp == 0) {
        // ------------------------------------------------------ TMA producer
        // (a convergent warp on warp-uniform values, the issuing lane elected per instruction,
        // like the MMA issuer below: four TMA requests per k-block of a wgrad tile cost ~600
        // clocks of issue from a single-lane branch, more than the 512 clocks its MMAs take)
        {
            const int rank_u = NCTA == 2 ? (int)(blockIdx.x & 1) : 0;   // == %cluster_ctarank (clusters of 2 along x)
            unsigned s = 0, ephase = 1;                      // ring slot and the parity its `empty` wait uses
            for (unsigned pit = 0;; ++pit) {
                const int tile = g_tile_of(g, tile0, tile_step, pit);
                if (tile >= g.total_tiles) break;
                const GTile t = g_decode(g, tile, BN, NCTA, rank_u);
                const GProblem &P = g.p[t.pi];
                if (lane == 0) g_trace(g, pit, 0);
                // in-launch dependency: the rows of A are the output of an earlier problem of
                // this group (possibly computed by other CTAs).  Row problems (forward, dgrad)
                // need their own 256-row block; a reduction over the rows (wgrad) needs the blocks
                // its K range crosses and asks for each one when its first k-block comes up.
                const bool wait_rows = P.wait && P.epi != GE_ATOMIC;
                const bool wait_k = P.wait && P.epi == GE_ATOMIC;
                if (wait_rows) { g_wait_block(P.wait + t.mt, P.wait_count); __syncwarp(); }
                // this CTA's share of the B tile: n_eff / NCTA columns from nb0 on
                const int nb_cols = t.n_eff / NCTA, nb0 = t.n0 + rank_u * nb_cols;
                const int nbox_b = (nb_cols + 63) >> 6;
                const int a_mn = P.a_mn, b_mn = P.b_mn, nkb = t.nkb;
                const unsigned bytes = A_BYTES + (b_mn ? (unsigned)nbox_b * 8192u : B_BYTES);
                for (int i = 0; i < nkb; ++i) {
                    g_mbar_wait_warp(empty0 + 8 * s, ephase);
                    const unsigned sa = base + s * STAGE, sb = sa + A_BYTES;
                    const int k0 = (t.kb0 + i) * G_BK;
                    if (wait_k && (i == 0 || k0 % (G_BM * NCTA) == 0)) {
                        g_wait_block(P.wait + k0 / (G_BM * NCTA), P.wait_count);
                        __syncwarp();
                    }
                    if (NCTA == 2) {
                        // both CTAs' copies complete on the LEADER's barrier, which expects them all
                        const unsigned fb = (full0 + 8 * s) & G_PEER_MASK;
                        if (rank_u == 0) g_mbar_expect_tx_warp(full0 + 8 * s, 2u * bytes);
                        if (a_mn) {
                            g_tma_2d_pair_warp(sa, &P.map_a, fb, t.m0, k0);
                            g_tma_2d_pair_warp(sa + 8192, &P.map_a, fb, t.m0 + 64, k0);
                        } else {
                            g_tma_2d_pair_warp(sa, &P.map_a, fb, k0, t.m0);
                        }
                        if (b_mn) {
                            for (int j = 0; j < nbox_b; ++j)
                                g_tma_2d_pair_warp(sb + j * 8192, &P.map_b, fb, nb0 + 64 * j, k0);
                        } else {
                            g_tma_2d_pair_warp(sb, &P.map_b, fb, k0, nb0);
                        }
                    } else {
                        g_mbar_expect_tx_warp(full0 + 8 * s, bytes);
                        if (a_mn) {
                            g_tma_2d_warp(sa, &P.map_a, full0 + 8 * s, t.m0, k0);
                            g_tma_2d_warp(sa + 8192, &P.map_a, full0 + 8 * s, t.m0 + 64, k0);
                        } else {
                            g_tma_2d_warp(sa, &P.map_a, full0 + 8 * s, k0, t.m0);
                        }
                        if (b_mn) {
                            for (int j = 0; j < nbox_b; ++j)
                                g_tma_2d_warp(sb + j * 8192, &P.map_b, full0 + 8 * s, nb0 + 64 * j, k0);
                        } else {
                            g_tma_2d_warp(sb, &P.map_b, full0 + 8 * s, k0, nb0);
                        }
                    }
                    if (++s == STAGES) { s = 0; ephase ^= 1u; }
                }
                if (lane == 0) g_trace(g, pit, 1);
            }
        }
    } else if (warp == 1) {
        // -------------------------------------------------------- MMA issuer
        // the WHOLE warp runs the loop on warp-uniform values; the issuing lane is elected
        // inside g_mma*_warp / g_commit*_warp (abn_tc_ptx.cuh: why).  Of a pair, only the leader
        // CTA issues.
        if (NCTA == 1 || (blockIdx.x & 1) == 0) {
            const unsigned tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
            unsigned s = 0, fphase = 0;                     // ring slot and its phase
            for (unsigned it = 0;; ++it) {
                const int tile = g_tile_of(g, tile0, tile_step, it);
                if (tile >= g.total_tiles) break;
                const GTile t = g_decode(g, tile, BN, NCTA, 0);
                const GProblem &P = g.p[t.pi];
                const unsigned ab = it & 1;
                if (lane == 0) g_trace(g, it, 2);
                g_mbar_wait_warp(tempty0 + 8 * ab, ((it >> 1) & 1) ^ 1);      // epilogue(s) drained this accumulator
                g_fence_after();
                if (lane == 0) g_trace(g, it, 3);
                const unsigned idesc = g_idesc(G_BM * NCTA, t.n_eff, P.a_mn, P.b_mn);
                const unsigned d_tmem = tmem_u + ab * BN;
                const unsigned a_step = P.a_mn ? (2048 >> 4) : (32 >> 4);
                const unsigned b_step = P.b_mn ? (2048 >> 4) : (32 >> 4);
                const int a_mn = P.a_mn, b_mn = P.b_mn, nkb = t.nkb;
                for (int i = 0; i < nkb; ++i) {
                    g_mbar_wait_warp(full0 + 8 * s, fphase);
                    g_fence_after();
                    const unsigned sa = base + s * STAGE;
                    const unsigned long long da = g_desc(sa, a_mn);
                    const unsigned long long db = g_desc(sa + A_BYTES, b_mn);
#pragma unroll
                    for (int k = 0; k < G_BK / G_UK; ++k) {
                        if (G_DBG(g, 8)) continue;
                        if (NCTA == 2)
                            g_mma_pair_warp(d_tmem, da + (unsigned long long)(a_step * k),
                                            db + (unsigned long long)(b_step * k), idesc, (i | k) != 0);
                        else
                            g_mma_warp(d_tmem, da + (unsigned long long)(a_step * k),
                                       db + (unsigned long long)(b_step * k), idesc, (i | k) != 0);
                    }
                    // frees the smem stage (in both CTAs of a pair) when these MMAs retire
                    if (NCTA == 2) g_commit_pair_warp(empty0 + 8 * s); else g_commit_warp(empty0 + 8 * s);
                    if (++s == STAGES) { s = 0; fphase ^= 1u; }
                }
                // accumulator complete (each CTA's epilogue watches its own barrier)
                if (NCTA == 2) g_commit_pair_warp(tfull0 + 8 * ab); else g_commit_warp(tfull0 + 8 * ab);
                if (lane == 0) g_trace(g, it, 4);
            }
        }
    } else {
        // ---------------------------------------------------------- epilogue
        const int wq = warp & 3;                        // TMEM lane quarter of this warp
        const int half = (warp - 2) >> 2;               // which of the quarter's two warps
        const int et = (warp - 2) * 32 + lane;          // 0 .. 255
        unsigned nbox = 0, ycount = 0;                  // boxes stored / y_below boxes consumed by this warp
        for (unsigned it = 0;; ++it) {
            const int tile = g_tile_of(g, tile0, tile_step, it);
            if (tile >= g.total_tiles) break;
            const GTile t = g_decode(g, tile, BN, NCTA, rank);
            const GProblem &P = g.p[t.pi];
            const unsigned ab = it & 1;
            float *bs = bias_s + ab * BN;
            const int row = t.m0 + wq * 32 + lane;
            const int c_first = half * 32;
            const int ew = warp - 2;
            const unsigned my_out = outs + ew * 8192u;
            const unsigned my_y = my_out;
            const unsigned ybar = ybar0 + 8 * ew;
            const int row0 = t.m0 + wq * 32;                 // first row of this warp's 32 x 32 boxes
            if (P.epi == GE_BIAS_ACT) {
                for (int c = et; c < BN; c += 32 * G_EPI_WARPS)
                    bs[c] = (P.bias && t.n0 + c < P.N) ? __ldg(P.bias + t.n0 + c) : 0.f;
                asm volatile("bar.sync 1, %0;" ::"n"(32 * G_EPI_WARPS) : "memory");
            } else if (P.epi == GE_DACT) {
                // the first chunk's y_below does not depend on the accumulator: fetch it now
                if (lane == 0 && t.n0 + 2 * c_first < P.n_cap && 2 * c_first < t.n_eff + 16) {
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    g_mbar_expect_tx(ybar, 4096u);
                    g_tma_2d(my_out + (nbox & 1u) * 4096u, &P.map_y, ybar, t.n0 + 2 * c_first, row0);
                }
            }
            if (et == 0) g_trace(g, it, 5);
            g_mbar_wait(tfull0 + 8 * ab, (it >> 1) & 1);
            g_fence_after();
            if (et == 0) g_trace(g, it, 6);
            const unsigned taddr = tmem + ((unsigned)(wq * 32) << 16) + ab * BN;
            if (t.nkb <= 0 || (G_DBG(g, 1) && P.epi != GE_DACT)) {
                // nothing was accumulated (cannot happen with the host's split sizes)
            } else if (P.epi == GE_ATOMIC) {
                // fp32 reduction of a split-K partial: 16-byte vector reds, one output row per
                // thread (REDG issue is per lane-op: the vector form is 4x cheaper than scalars)
                float *out = static_cast<float *>(P.out);
                const int n_w = P.ones_out ? P.N - 1 : P.N;          // columns that belong to `out`
                const bool vec_ok = (P.ldo & 3) == 0;
                for (int c0 = c_first; c0 < t.n_eff; c0 += 64) {
                    float v[32];
                    g_ld32(taddr + c0, v);
                    const int gcol0 = t.n0 + c0;
                    if (row < P.M) {
                        float *op = out + (long long)row * P.ldo + gcol0;
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            if (vec_ok && gcol0 + 4 * q + 4 <= n_w) {
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                                             ::"l"(op + 4 * q), "f"(v[4 * q]), "f"(v[4 * q + 1]),
                                               "f"(v[4 * q + 2]), "f"(v[4 * q + 3]) : "memory");
                            } else {
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const int gcol = gcol0 + 4 * q + e;
                                    if (gcol < n_w) atomicAdd(op + 4 * q + e, v[4 * q + e]);
                                    else if (gcol == n_w && P.ones_out)
                                        atomicAdd(P.ones_out + row, v[4 * q + e]);
                                }
                            }
                        }
                    }
                }
            } else if (P.tma_store) {
                GEpi e;
                e.map_c = &P.map_c; e.map_y = &P.map_y; e.bs = bs;
                e.taddr = taddr; e.my_out = my_out; e.my_y = my_y; e.ybar = ybar;
                e.n0 = t.n0; e.row0 = row0; e.n_eff = t.n_eff; e.n_cap = P.n_cap; e.N = P.N;
                e.ones_col = P.ones_col; e.c_first = 2 * c_first; e.lane = lane;      // 64-column steps
#ifdef ABN_TC_DEBUG
                e.dbg = g.dbg;
#else
                e.dbg = 0;
#endif
                const int mode = P.epi * 4 + P.act;
                switch (mode) {
                    case 0: g_epi_bf16_tile<GE_BIAS_ACT, 0, BN>(e, nbox, ycount); break;
                    case 1: g_epi_bf16_tile<GE_BIAS_ACT, 1, BN>(e, nbox, ycount); break;
                    case 2: g_epi_bf16_tile<GE_BIAS_ACT, 2, BN>(e, nbox, ycount); break;
                    case 3: g_epi_bf16_tile<GE_BIAS_ACT, 3, BN>(e, nbox, ycount); break;
                    case 4: g_epi_bf16_tile<GE_DACT, 0, BN>(e, nbox, ycount); break;
                    case 5: g_epi_bf16_tile<GE_DACT, 1, BN>(e, nbox, ycount); break;
                    case 6: g_epi_bf16_tile<GE_DACT, 2, BN>(e, nbox, ycount); break;
                    default: g_epi_bf16_tile<GE_DACT, 3, BN>(e, nbox, ycount); break;
                }
            } else {
                // fp32 rows (the embeddings): bias + activation, 16-byte stores
                const int n_cap = P.n_cap, N = P.N, act = P.act;
                const long long ldo = P.ldo;
                float *outp = static_cast<float *>(P.out);
                const bool row_ok = row < P.M;
                for (int c0 = c_first; c0 < BN; c0 += 64) {
                    const int gcol0 = t.n0 + c0;
                    if (gcol0 >= n_cap || c0 >= t.n_eff + 16) break;     // warp-uniform
                    float v[32];
                    g_ld32(taddr + c0, v);
                    switch (act) {
                        case 1: g_bias_act32<1>(v, bs + c0); break;
                        case 2: g_bias_act32<2>(v, bs + c0); break;
                        case 3: g_bias_act32<3>(v, bs + c0); break;
                        default: g_bias_act32<0>(v, bs + c0); break;
                    }
                    if (row_ok) {
                        float *op = outp + (long long)row * ldo + gcol0;
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            if (gcol0 + 4 * q + 4 <= N && (ldo & 3) == 0)
                                *reinterpret_cast<float4 *>(op + 4 * q) =
                                    make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                            else
#pragma unroll
                                for (int e2 = 0; e2 < 4; ++e2)
                                    if (gcol0 + 4 * q + e2 < N) op[4 * q + e2] = v[4 * q + e2];
                        }
                    }
                }
            }
            g_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (NCTA == 2) g_mbar_arrive_cta0(tempty0 + 8 * ab); else g_mbar_arrive(tempty0 + 8 * ab);
            }
            if (P.signal) {
                // publish this CTA's rows of the tile: every warp's stores have landed, then one
                // release-increment of the m-tile's counter (consumers: producers of later problems)
                if (lane == 0) g_store_wait_all();          // TMA-stored boxes (lane 0 issued them)
                if (!P.tma_store) __threadfence();           // rows stored directly from registers
                asm volatile("bar.sync 1, %0;" ::"n"(32 * G_EPI_WARPS) : "memory");
                if (et == 0) {
                    __threadfence();
                    asm volatile("fence.proxy.async;" ::: "memory");
                    asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(P.signal + t.mt) : "memory");
                }
            }
            if (et == 0) g_trace(g, it, 7);
        }
    }
    if (warp >= 2 && lane == 0) g_store_wait_all();            // staged boxes are on their way out of smem
    g_fence_before();
    if (NCTA == 2) g_cluster_sync(); else __syncthreads();     // nobody leaves while the peer may still signal it
    if (warp == 2) {
        g_fence_after();
        if (NCTA == 2)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(2 * BN)
                         : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(2 * BN)
                         : "memory");
    }
}

// ------------------------------------------------------------------ host ---
template <int BN, int NCTA, bool WG>
static int g_launch(const GGroup &g, int sm_count, cudaStream_t st) {
    constexpr unsigned stage = G_BM * G_BK * 2 + (BN / NCTA) * G_BK * 2;
    constexpr int STAGES = g_stages<BN, NCTA, WG>();
    constexpr unsigned smem = STAGES * stage + (WG ? 0 : G_EPI_WARPS * 8192) + 256 + 2 * BN * 4 + 1024;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(tc_group_kernel<BN, NCTA, WG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem) != cudaSuccess)
            return set_error(ABN_EIO, "abn_gemm_bf16_group: cannot reserve %u bytes of shared memory",
                             smem);
        configured = true;
    }
    int grid = g.total_tiles * NCTA < sm_count ? g.total_tiles * NCTA : sm_count;
    grid -= grid % NCTA;
    if (const char *e = getenv("ABN_GEMM_GRID")) {       // experiments: fewer CTAs
        const int lim = atoi(e);
        if (lim >= NCTA && lim < grid) grid = lim - lim % NCTA;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(G_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = NCTA;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_use_pdl() ? 2 : 1;
    if (cudaLaunchKernelEx(&cfg, tc_group_kernel<BN, NCTA, WG>, g) != cudaSuccess)
        return check_launch("abn_gemm_bf16_group");
    return check_launch("abn_gemm_bf16_group");
}

}  // namespace abn

using namespace abn;

// debug hook (not in the public header): device buffer of 148 * 4 * 8 int64 timestamps
extern "C" { __attribute__((visibility("default"))) void *abn_gemm_trace_buffer = nullptr; }

extern "C" int abn_gemm_bf16_group(const abn_gemm_problem *problems, int n_problems,
                                   abn_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    if (n_problems == 0) return ABN_OK;
    if (!problems || n_problems < 0 || n_problems > G_MAXP)
        return set_error(ABN_EINVAL, "abn_gemm_bf16_group: 1..%d problems per call", G_MAXP);
    static int sm_count = 0;
    if (!sm_count) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    }
    int max_n = 0;
    for (int i = 0; i < n_problems; ++i) {
        const abn_gemm_problem &q = problems[i];
        if (!q.A || !q.B || !q.out || q.M <= 0 || q.N <= 0 || q.K <= 0 || q.epilogue < 0 ||
            q.epilogue > 2 || q.act < 0 || q.act > 3)
            return set_error(ABN_EINVAL, "abn_gemm_bf16_group: bad argument in problem %d", i);
        if (q.epilogue == GE_DACT && !q.yprev)
            return set_error(ABN_EINVAL, "abn_gemm_bf16_group: problem %d: act' epilogue needs yprev", i);
        if (q.epilogue == GE_ATOMIC && !q.out_f32)
            return set_error(ABN_EINVAL, "abn_gemm_bf16_group: problem %d: reduction output is fp32", i);
        const int n_cols = q.N + ((q.epilogue == GE_ATOMIC ? q.ones_out != nullptr : q.ones_col != 0) ? 1 : 0);
        if (n_cols > max_n) max_n = n_cols;
    }
    const int bn = max_n <= 128 ? 128 : 256;
    // CTA pairs (cta_group::2) halve the B traffic per MMA; ABN_GEMM_1CTA=1 keeps single CTAs
    static int ncta = 0;
    if (!ncta) {
        const char *e = getenv("ABN_GEMM_1CTA");
        ncta = (e && e[0] == '1') ? 1 : 2;
    }
    GGroup g;
    memset(&g, 0, sizeof(g));
    g.n_problems = n_problems;
    int tile = 0;
    for (int i = 0; i < n_problems; ++i) {
        const abn_gemm_problem &q = problems[i];
        GProblem &P = g.p[i];
        const int ones_in = (q.epilogue == GE_ATOMIC && q.ones_out) ? 1 : 0;
        P.M = q.M; P.N = q.N + ones_in; P.K = q.K;
        P.a_mn = q.a_mn ? 1 : 0; P.b_mn = q.b_mn ? 1 : 0;
        P.epi = q.epilogue; P.act = q.act; P.out_f32 = q.out_f32 ? 1 : 0;
        P.ones_col = (q.epilogue != GE_ATOMIC && q.ones_col) ? 1 : 0;
        P.bias = q.bias; P.out = q.out; P.ldo = q.ldo;
        P.yprev = static_cast<const __nv_bfloat16 *>(q.yprev); P.ld_yprev = q.ld_yprev;
        P.ones_out = ones_in ? q.ones_out : nullptr;
        P.signal = q.signal; P.wait = q.wait; P.wait_count = q.wait_count;
        if (!P.out_f32 && ((q.ldo & 7) || q.ldo < P.N + P.ones_col))
            return set_error(ABN_EINVAL, "abn_gemm_bf16_group: problem %d: bf16 output rows must be "
                             "padded to a multiple of 8 elements covering N%s", i,
                             P.ones_col ? " + 1" : "");
        // A: K-major [M, K] or MN-major [K, M];  B: K-major [N, K] or MN-major [K, N]
        int rc = P.a_mn ? g_make_map(&P.map_a, q.A, q.K, q.M, q.lda, 64, 64)
                        : g_make_map(&P.map_a, q.A, q.M, q.K, q.lda, G_BK, G_BM);
        if (rc) return rc;
        rc = P.b_mn ? g_make_map(&P.map_b, q.B, q.K, P.N, q.ldb, 64, 64)
                    : g_make_map(&P.map_b, q.B, P.N, q.K, q.ldb, G_BK, bn / ncta);
        if (rc) return rc;
        P.tma_store = (!P.out_f32 && q.epilogue != GE_ATOMIC) ? 1 : 0;
        if (P.tma_store) {
            rc = g_make_map(&P.map_c, q.out, q.M, q.ldo, q.ldo, 64, 32);
            if (rc) return rc;
        }
        if (q.epilogue == GE_DACT) {
            if (P.out_f32)
                return set_error(ABN_EINVAL, "abn_gemm_bf16_group: problem %d: the act' epilogue "
                                 "writes bf16", i);
            rc = g_make_map(&P.map_y, q.yprev, q.M, q.ld_yprev, q.ld_yprev, 64, 32);
            if (rc) return rc;
        }
        P.tiles_m = (P.M + G_BM * ncta - 1) / (G_BM * ncta);
        P.n_cap = P.N + P.ones_col;
        P.tiles_n = (P.n_cap + bn - 1) / bn;
        const int total_kb = (P.K + G_BK - 1) / G_BK;
        int splits = (q.epilogue == GE_ATOMIC && q.split_k > 1) ? q.split_k : 1;
        if (splits > total_kb) splits = total_kb;
        P.kb_per_split = (total_kb + splits - 1) / splits;
        P.splits = (total_kb + P.kb_per_split - 1) / P.kb_per_split;
        P.tile_beg = tile;
        tile += P.tiles_m * P.tiles_n * P.splits;
    }
    // wait_count 0 = "all tiles of the problem of this group that signals what I wait for"
    for (int i = 0; i < n_problems; ++i) {
        GProblem &P = g.p[i];
        if (!P.wait || P.wait_count > 0) continue;
        for (int j = 0; j < i; ++j)
            if (g.p[j].signal == P.wait) P.wait_count = g.p[j].tiles_n * g.p[j].splits * ncta;
        if (P.wait_count <= 0)
            return set_error(ABN_EINVAL, "abn_gemm_bf16_group: problem %d waits for a counter no "
                             "earlier problem of the group signals", i);
    }
    g.total_tiles = tile;
    g.trace = reinterpret_cast<long long *>(abn_gemm_trace_buffer);
    { const char *e = getenv("ABN_GEMM_ROT"); g.rot = e ? atoi(e) : 0; }
#ifdef ABN_TC_DEBUG
    { const char *e = getenv("ABN_GEMM_DBG"); g.dbg = e ? atoi(e) : 0; }
#endif
    cudaStream_t st = (cudaStream_t)stream;
    bool all_red = true;
    for (int i = 0; i < n_problems; ++i) all_red = all_red && problems[i].epilogue == GE_ATOMIC;
    { const char *e = getenv("ABN_GEMM_WG_STAGES"); if (e && e[0] == '0') all_red = false; }
    if (all_red) {
        if (ncta == 2)
            return bn == 128 ? g_launch<128, 2, true>(g, sm_count, st) : g_launch<256, 2, true>(g, sm_count, st);
        return bn == 128 ? g_launch<128, 1, true>(g, sm_count, st) : g_launch<256, 1, true>(g, sm_count, st);
    }
    if (ncta == 2)
        return bn == 128 ? g_launch<128, 2, false>(g, sm_count, st) : g_launch<256, 2, false>(g, sm_count, st);
    return bn == 128 ? g_launch<128, 1, false>(g, sm_count, st) : g_launch<256, 1, false>(g, sm_count, st);
}
