// abn_tc3.cu -- kernel (3b): the embedder's FORWARD pass (MODE 0) and its dz chain (MODE 1) with the
// activations resident on chip.  One launch runs every layer of  y = act(x W^T + b)  (abnet3/model.py:133-170,
// :179-186) for a 256-row block per CTA pair:
//
//   * the block's activations live in shared memory as the A operand of the next layer
//     ("slab": 8 k-blocks of 128 rows x 64 bf16, 128-byte swizzle = the layout TMA would have
//     produced), written there by the epilogue of the layer before -- they are never re-read
//     from L2, and the next layer does not wait for a global round trip;
//   * only the weights stream (TMA ring, this CTA's half of every 256-column B tile);
//   * the accumulators of a layer's two 256-column halves fill the 512 TMEM columns; the
//     epilogue drains them k-block by k-block and the next layer's MMAs start on a k-block
//     as soon as both CTAs have written it (per-k-block mbarriers on the leader CTA);
//   * every hidden layer's rows also leave through TMA stores from the slab (the backward
//     pass needs them), the last layer's rows as fp32 (the embeddings).
//
// Why (measured, tools/time_chain.py + tools/trace_tc2.py): the chained grouped GEMM is bound
// by L2 -> SM operand delivery (196 MB per forward pass at ~11 TB/s) and, with the layers'
// dependencies, by the latency of store -> signal -> load between layers (44 us against 31 us
// without dependencies).  Keeping A on chip removes half of the operand traffic and all of
// that latency.
//
// Round 2, second half (all measured with the kernel's own %globaltimer stamps, tools/trace_fused.py):
//   * the TMA producer and the MMA issuer are CONVERGENT warps on warp-uniform values (the issuing
//     lane is elected inside the asm block, abn_tc_ptx.cuh): from a single-lane branch every
//     tcgen05.mma cost ~150 clocks of issue (ELECT / R2UR loops) for 128 clocks of tensor work;
//   * the epilogue publishes a slab block with a CTA-scope release (`.release.cluster` is a
//     MEMBAR.ALL.GPU per block) and writes block cb once the layer's last tile has retired ITS
//     k-block cb (kfree barriers), not the whole tile;
//   * MODE 0 can compute the pair loss and the output layer's dz in the last layer's epilogue
//     (interleaved pair rows, f_epi_loss): the embeddings never reach HBM;
//   * MODE 1: a proxy fence orders the TMA refill of a warp's y_below box after the loads of the
//     previous one.
//
// Serves networks whose layer widths fit the slab: n_in <= 512, n_out (+ the ones column)
// <= 512; anything else takes the chained launch of abn_tc2.cu.
#include <stdlib.h>
#include <string.h>

#include "abn_tc_ptx.cuh"
#include "abn_drop.cuh"

namespace abn {

constexpr int F_MAXL = ABN_MLP_MAX_LAYERS;
constexpr int F_KB = 8;                         // slab k-blocks (8 x 64 = 512 features)
constexpr int F_STAGES_FWD = 5, F_STAGES_DGRAD = 4;     // B ring (dgrad: the y_below boxes take the rest)
constexpr unsigned F_SLAB_KB_BYTES = 128 * 64 * 2;      // 16 KB: this CTA's 128 rows of one k-block
constexpr unsigned F_B_BYTES = 128 * 64 * 2;            // this CTA's half of a 256-column B tile

struct FLayer {
    CUtensorMap map_w;              // forward: W [n_out, n_in], box {64 k, 128 rows};  dgrad: box {64 cols, 64 k rows}
    CUtensorMap map_out;            // bf16 output [rows, ld], box {64 cols, 32 rows}
    CUtensorMap map_y;              // dgrad: y_below [rows, ld] bf16, box {64 cols, 32 rows}
    const float *bias;
    float *out32; long long ld32;   // last layer: fp32 rows
    int n_in, n_out, act, ones_col, out_f32;
    int nkb, tiles_n, n_cap;        // GEMM view: K = 64 nkb, N = n_cap (forward: n_out + ones_col; dgrad: n_in)
    DropArgs drop;                  // forward: this layer's dropout; dgrad: the dropout of the layer below
};
// MODE 0 with the pair loss fused into the last layer's epilogue (abnet3/loss.py:46-67, :85-105 +
// the output layer's act'): rows are INTERLEAVED (row 2k = x1[k], row 2k + 1 = x2[k]) so both
// embeddings of a pair sit in neighbouring lanes of one epilogue warp.
struct FLoss {
    int on, kind, write_emb;
    float margin, scale;
    const float *y;                 // [rows / 2] labels
    float *loss;                    // += sum of the loss terms * scale
    __nv_bfloat16 *dz; long long ld_dz;     // dz of the output layer, bf16 [rows, ld_dz]
};
struct FChain {
    FLoss loss;
    CUtensorMap map_x;              // x [rows, n_in0] bf16, box {64, 128}
    FLayer L[F_MAXL];
    int n_layers, rows, tiles_m;
    long long *trace;               // debug: role timestamps per CTA and layer (NULL in production)
    int debug;                      // debug (tools/trace_fused.py): 1 skip the B-ring waits, 2 skip the slab waits
};

// debug timeline (tools/trace_fused.py): [cta][layer][slot]
__device__ __forceinline__ void f_trace(const FChain &ch, int layer, int slot) {
    if (ch.trace) {
        long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        ch.trace[((size_t)blockIdx.x * 8 + layer) * 16 + slot] = t;
    }
}

// second debug region: when each B stage was (a) requested by the producer, (b) seen full by the
// MMA issuer -- [which][cta][layer][tile][k-block] behind the 148 x 8 x 16 role stamps
__device__ __forceinline__ void f_trace_kb(const FChain &ch, int which, int layer, int nt, int kb) {
    if (ch.trace) {
        long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        ch.trace[148 * 8 * 16 + ((((size_t)which * 148 + blockIdx.x) * 8 + layer) * 2 + nt) * 8 + kb] = t;
    }
}

__device__ __forceinline__ void f_mbar_wait_cluster(unsigned bar, unsigned parity) {
    unsigned done = 0;
    for (unsigned spin = 0; spin < (1u << 27); ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}
// arrive on the LEADER CTA's copy of `bar` (own copy when this CTA is the leader).  What is published
// are this CTA's OWN shared-memory rows (st.shared + fence.proxy.async before the warp meets), which
// the pair's tensor core reads in place: CTA-scope release, the form CUTLASS's ClusterBarrier::arrive
// uses on remote barriers.  `.release.cluster` compiles to MEMBAR.ALL.GPU + ERRBAR in front of every
// arrive -- ~1 us per 64-column block with the TMA stores of the same warp in flight.
__device__ __forceinline__ void f_arrive_leader_release(unsigned bar) {
    asm volatile("mbarrier.arrive.release.cta.shared::cluster.b64 _, [%0];"
                 ::"r"(bar & G_PEER_MASK) : "memory");
}

__device__ __forceinline__ int f_n_eff(const FLayer &L, int nt) {
    const int rem = L.n_cap - nt * 256;        // (the ones column is computed as a zero and overwritten)
    return rem >= 256 ? 256 : ((rem + 15) & ~15);
}

// One 64-column block of a hidden layer: TMEM -> bias + activation -> bf16 -> this warp's 32
// rows of slab k-block cb (the next layer's A operand), then a TMA store of the same box.
// z -> keep ? z / (1 - p) : 0 on 32 consecutive columns starting at col0 (a multiple of 4)
__device__ __forceinline__ void f_drop32(float (&v)[32], const DropArgs &dr, unsigned long long dkey,
                                         long long row, int col0) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const unsigned long long bits = drop_bits4(dkey, row, (col0 >> 2) + q);
#pragma unroll
        for (int e = 0; e < 4; ++e)
            v[4 * q + e] = drop_keep_of(bits, e, dr.thresh) ? v[4 * q + e] * dr.inv_keep : 0.f;
    }
}

// `gate` (0 = none): an mbarrier (with `gate_parity`) that must have completed before the slab is
// written -- the commit of the layer's LAST accumulator: the tensor core reads the slab until
// then.  The TMEM load and the math of the first accumulator's blocks run while the second
// accumulator's MMAs are still in flight; only the stores wait.
template <int ACT, bool DROP>
__device__ __forceinline__ void f_epi_block(unsigned taddr, const float *bs, unsigned dst, int lane,
                                            int ones_at, const DropArgs &dr, unsigned long long dkey,
                                            long long row, int col0, unsigned gate, unsigned gate_parity,
                                            long long *tr = nullptr) {
    const unsigned swz = (unsigned)(lane & 7);
    float v64[64];
    g_ld64(taddr, v64);
    unsigned pk[32];
#pragma unroll
    for (int hseg = 0; hseg < 2; ++hseg) {
        float (&v)[32] = *reinterpret_cast<float (*)[32]>(&v64[32 * hseg]);
        unsigned (&pkh)[16] = *reinterpret_cast<unsigned (*)[16]>(&pk[16 * hseg]);
        if (DROP && dr.state) {     // Linear -> Dropout -> act (abnet3/model.py:136-141)
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += bs[32 * hseg + j];
            f_drop32(v, dr, dkey, row, col0 + 32 * hseg);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = g_act<ACT>(v[j]);
#pragma unroll
            for (int j = 0; j < 16; ++j) pkh[j] = g_pack_bf16(v[2 * j], v[2 * j + 1]);
        } else if (ACT == 1 || ACT == 2) {
            g_bias_act32_packed<ACT>(v, bs + 32 * hseg, pkh);
        } else {
            g_bias_act32<ACT>(v, bs + 32 * hseg);
#pragma unroll
            for (int j = 0; j < 16; ++j) pkh[j] = g_pack_bf16(v[2 * j], v[2 * j + 1]);
        }
        const int jo = ones_at - 32 * hseg;             // the column of ones, if it is in this half
        if (jo >= 0 && jo < 32) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                if (2 * j == jo) pkh[j] = (pkh[j] & 0xffff0000u) | 0x3f80u;
                if (2 * j + 1 == jo) pkh[j] = (pkh[j] & 0x0000ffffu) | 0x3f800000u;
            }
        }
    }
    if (tr) { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); tr[0] = t; }
    if (gate) {
        g_mbar_wait(gate, gate_parity);
        g_fence_after();
    }
    if (tr) { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); tr[1] = t; }
#pragma unroll
    for (int q = 0; q < 8; ++q)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                     ::"r"(dst + lane * 128 + (((unsigned)q ^ swz) << 4)),
                       "r"(pk[4 * q]), "r"(pk[4 * q + 1]), "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3])
                     : "memory");
}

// dgrad: TMEM -> x act'(y_below) -> bf16 -> slab rows (y_below's box sits in `ybuf`)
template <int ACT, bool DROP>
__device__ __forceinline__ void f_epi_block_d(unsigned taddr, const uint4 (&yc)[8], unsigned dst, int lane,
                                              const DropArgs &dr, unsigned long long dkey, long long row,
                                              int col0, unsigned gate, unsigned gate_parity) {
    const unsigned swz = (unsigned)(lane & 7);
    float v64[64];
    g_ld64(taddr, v64);
    unsigned pk[32];
#pragma unroll
    for (int hseg = 0; hseg < 2; ++hseg) {
        float (&v)[32] = *reinterpret_cast<float (*)[32]>(&v64[32 * hseg]);
        const uint4 yh[4] = {yc[4 * hseg], yc[4 * hseg + 1], yc[4 * hseg + 2], yc[4 * hseg + 3]};
        g_dact32<ACT>(v, yh);
        if (DROP && dr.state) f_drop32(v, dr, dkey, row, col0 + 32 * hseg);     // dz * keep / (1 - p)
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[16 * hseg + j] = g_pack_bf16(v[2 * j], v[2 * j + 1]);
    }
    if (gate) {
        g_mbar_wait(gate, gate_parity);
        g_fence_after();
    }
#pragma unroll
    for (int q = 0; q < 8; ++q)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                     ::"r"(dst + lane * 128 + (((unsigned)q ^ swz) << 4)),
                       "r"(pk[4 * q]), "r"(pk[4 * q + 1]), "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3])
                     : "memory");
}

// The last forward layer with the loss fused in: this warp holds rows row0 .. row0 + 31 (one per
// lane) of the 64-column block cb of the embeddings (n_out <= 128: blocks 0 and 1, taken by two
// warps of a lane quarter, which meet at named barrier 2 + wq).
//   pass 1  e = act(z + b); dot / |a|^2 partials with the partner row (lane ^ 1) by shuffle
//   ------  the two column blocks exchange their partials through shared memory
//   pass 2  dz = dL/dc * (partner * inv - k * e) * act'(e) -> bf16 rows; embeddings -> fp32 rows
// Same arithmetic per element as pair_loss_dz_vec_kernel (abn_fused.cu).  The loops stay ROLLED
// (16 columns per trip): this code runs once per row block, so a fully unrolled body is paid for
// in instruction-cache misses (27 KB of straight-line code: 10 us per launch, tools/
// trace_fused_loss.py), not in issue slots.
template <int ACT>
__device__ __noinline__ void f_epi_loss(const FLoss &F, int rows, int n_out, float *out32, long long ld32,
                                        unsigned taddr, const float *bs, float *exch, int wq, int cb,
                                        int nblk, int row0, int lane) {
    const int row = row0 + lane;
    const int col0 = 64 * cb;
    const int ncol = n_out - col0 < 64 ? n_out - col0 : 64;
    float dot = 0.f, na = 0.f;
#pragma unroll 1
    for (int c = 0; c < ncol; c += 16) {
        float v[16];
        g_ld16(taddr + c, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float e = g_act<ACT>(v[j] + bs[c + j]);
            const float pr = __shfl_xor_sync(0xffffffffu, e, 1);
            if (c + j < ncol) {
                dot = fmaf(e, pr, dot);
                na = fmaf(e, e, na);
            }
        }
    }
    if (nblk > 1) {                                       // the other block's partials of the same rows
        float *mine = exch + ((wq * 2 + cb) * 32 + lane) * 2;
        mine[0] = dot; mine[1] = na;
        asm volatile("bar.sync %0, 64;" ::"r"(2 + wq) : "memory");
        const float *other = exch + ((wq * 2 + (cb ^ 1)) * 32 + lane) * 2;
        // (block 0's sum first in both warps: the two halves of a row agree bit for bit)
        dot = cb == 0 ? dot + other[0] : other[0] + dot;
        na = cb == 0 ? na + other[1] : other[1] + na;
    }
    const float nb = __shfl_xor_sync(0xffffffffu, na, 1);
    constexpr float EPS = 1e-6f;                          // nn.CosineSimilarity(eps=1e-6), loss.py:44
    const float ra = sqrtf(na), rb = sqrtf(nb);
    const float an = fmaxf(ra, EPS), bn = fmaxf(rb, EPS);
    const float inv = 1.f / (an * bn);
    const float c = dot * inv;
    const bool live = row < rows;
    const float lab = live ? __ldg(F.y + (row >> 1)) : 0.f;
    float term, dldc;
    if (F.kind == 0) {            // coscos2, loss.py:59-62
        if (lab == 1.f)       { term = 0.5f * (1.f - c); dldc = -0.5f; }
        else if (lab == -1.f) { term = c * c;            dldc = 2.f * c; }
        else                  { term = c;                dldc = 1.f; }
    } else {                      // cosmargin, loss.py:98-101
        if (lab == 1.f)       { term = 1.f - c;          dldc = -1.f; }
        else if (lab == -1.f) { const float h = c - F.margin;
                                term = fmaxf(h, 0.f);    dldc = h > 0.f ? 1.f : 0.f; }
        else                  { term = c;                dldc = 1.f; }
    }
    if (cb == 0) {                // one term per pair: the even row of the pair, first column block
        float t = (live && !(lane & 1)) ? term : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0 && t != 0.f) atomicAdd(F.loss, t * F.scale);
    }
    const float g = dldc * F.scale;
    const float ka = ra > EPS ? c / (an * an) : 0.f;
    __nv_bfloat16 *zp = F.dz + (long long)row * F.ld_dz + col0;
    float *op = out32 + (long long)row * ld32 + col0;
#pragma unroll 1
    for (int c0 = 0; c0 < ncol; c0 += 16) {
        float v[16];
        g_ld16(taddr + c0, v);
        unsigned pk[8];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
            const float e0 = g_act<ACT>(v[j] + bs[c0 + j]), e1 = g_act<ACT>(v[j + 1] + bs[c0 + j + 1]);
            const float p0 = __shfl_xor_sync(0xffffffffu, e0, 1), p1 = __shfl_xor_sync(0xffffffffu, e1, 1);
            v[j] = e0; v[j + 1] = e1;
            pk[j >> 1] = g_pack_bf16(g_dact<ACT>(g * (p0 * inv - ka * e0), e0),
                                     g_dact<ACT>(g * (p1 * inv - ka * e1), e1));
        }
        if (!live) continue;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            if (c0 + 8 * q + 8 <= ncol) {
                *reinterpret_cast<uint4 *>(zp + c0 + 8 * q) =
                    make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
            } else {
#pragma unroll
                for (int e2 = 0; e2 < 8; ++e2)
                    if (c0 + 8 * q + e2 < ncol) {
                        const unsigned w = pk[4 * q + (e2 >> 1)];
                        reinterpret_cast<unsigned short *>(zp)[c0 + 8 * q + e2] =
                            (unsigned short)((e2 & 1) ? (w >> 16) : (w & 0xffffu));
                    }
            }
        }
        if (F.write_emb) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (c0 + j < ncol) op[c0 + j] = v[j];
        }
    }
}

// MODE 0: forward (B = W K-major, bias + activation, last layer fp32)
// MODE 1: dgrad   (B = W MN-major, x act'(y_below) with y_below fetched per warp by TMA)
constexpr int F_EW_FWD = 16, F_EW_DGRAD = 8;        // epilogue warps: NQ = EW / 4 share a TMEM lane quarter
template <int MODE, bool DROP, bool LOSS>
__global__ void __launch_bounds__(64 + 32 * (MODE == 0 ? F_EW_FWD : F_EW_DGRAD), 1)
mlp_chain_kernel(const __grid_constant__ FChain ch) {
    constexpr int EW = MODE == 0 ? F_EW_FWD : F_EW_DGRAD, NQ = EW / 4;
    constexpr int F_STAGES = MODE == 0 ? F_STAGES_FWD : F_STAGES_DGRAD;
    extern __shared__ unsigned char smem_raw[];
    if (threadIdx.x == 0) f_trace(ch, 7, 12);                           // (debug) kernel entry
    const unsigned raw = g_smem_u32(smem_raw);
    const unsigned base = (raw + 1023u) & ~1023u;
    unsigned char *gen = smem_raw + (base - raw);
    const unsigned slab = base;                                         // F_KB x 16 KB
    const unsigned ring = slab + F_KB * F_SLAB_KB_BYTES;                // F_STAGES x 16 KB
    const unsigned bars = ring + F_STAGES * F_B_BYTES;
    const unsigned bfull0 = bars, bempty0 = bars + 8 * F_STAGES;
    const unsigned xfull0 = bars + 16 * F_STAGES;                       // [F_KB] layer 0: TMA
    const unsigned sfull0 = xfull0 + 8 * F_KB;                          // [F_KB] later layers: epilogue
    const unsigned afull0 = sfull0 + 8 * F_KB, aempty0 = afull0 + 16;   // [2] accumulators
    const unsigned slab_free = aempty0 + 16, drained = slab_free + 8;
    const unsigned tptr = drained + 8;
    const unsigned ybar0 = tptr + 8;                                    // [F_EW_DGRAD] dgrad: y_below boxes
    const unsigned kfree0 = ybar0 + 8 * F_EW_DGRAD;                     // [F_KB] slab k-block no longer read by this layer
    constexpr unsigned BAR_BYTES = 16 * F_STAGES + 16 * F_KB + 32 + 16 + 16 + 8 * F_EW_DGRAD + 8 * F_KB;
    volatile unsigned *tptr_gen = reinterpret_cast<volatile unsigned *>(
        gen + F_KB * F_SLAB_KB_BYTES + F_STAGES * F_B_BYTES + (tptr - bars));
    float *bias_s = reinterpret_cast<float *>(
        gen + F_KB * F_SLAB_KB_BYTES + F_STAGES * F_B_BYTES + ((BAR_BYTES + 127u) & ~127u));  // [2][512] (forward)
    // dgrad: one 32 x 64 bf16 box of y_below per epilogue warp, 1024-byte aligned (128-byte swizzle)
    const unsigned ybuf0 = (bars + BAR_BYTES + 1023u) & ~1023u;

    const int rank = (int)g_cluster_rank();
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    // (the shuffle makes the warp index provably warp-uniform: role branches and everything
    // derived from it stay on the uniform datapath)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < F_STAGES; ++s) { g_mbar_init(bfull0 + 8 * s, 1); g_mbar_init(bempty0 + 8 * s, 1); }
        for (int k = 0; k < F_KB; ++k) {
            g_mbar_init(xfull0 + 8 * k, 1);
            g_mbar_init(sfull0 + 8 * k, 8);          // 4 row-quarter warps x 2 CTAs
            g_mbar_init(kfree0 + 8 * k, 1);
        }
        for (int a = 0; a < 2; ++a) {
            g_mbar_init(afull0 + 8 * a, 1);
            g_mbar_init(aempty0 + 8 * a, 2 * EW);
        }
        g_mbar_init(slab_free, 1);
        g_mbar_init(drained, EW);
        for (int w = 0; w < F_EW_DGRAD; ++w) g_mbar_init(ybar0 + 8 * w, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(tptr), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    g_fence_before();
    g_cluster_sync();
    g_fence_after();
    const unsigned tmem = *tptr_gen;
    if (threadIdx.x == 0) f_trace(ch, 7, 13);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (threadIdx.x == 0) f_trace(ch, 7, 14);

    if (warp == 0) {
        // -------------------------------------------------- TMA producer: x, then the weights
        // (a convergent warp on warp-uniform values, the issuing lane elected per instruction,
        // like the MMA issuer below)
        {
            const int rank_u = blockIdx.x & 1;          // == %cluster_ctarank (clusters of 2 along x)
            unsigned s = 0, ephase = 1, iter = 0;       // B ring slot and the parity its `empty` wait uses
            for (int rb = blockIdx.x >> 1; rb < ch.tiles_m; rb += gridDim.x >> 1, ++iter) {
                const int m0 = (rb * 2 + rank_u) * G_BM;
                if (iter > 0) {          // the slab is reused: MMAs and output stores of the last block are done
                    g_mbar_wait_warp(slab_free, (iter - 1) & 1);
                    g_mbar_wait_warp(drained, (iter - 1) & 1);
                }
                const int nkb0 = ch.L[0].nkb;
                for (int kb = 0; kb < nkb0; ++kb) {
                    if (rank_u == 0) g_mbar_expect_tx_warp(xfull0 + 8 * kb, 2u * F_SLAB_KB_BYTES);
                    g_tma_2d_pair_warp(slab + kb * F_SLAB_KB_BYTES, &ch.map_x, (xfull0 + 8 * kb) & G_PEER_MASK,
                                       kb * G_BK, m0);
                }
                for (int l = 0; l < ch.n_layers; ++l) {
                    const FLayer &L = ch.L[l];
                    const int nkb = L.nkb;
                    for (int nt = 0; nt < L.tiles_n; ++nt) {
                        const int nb_cols = f_n_eff(L, nt) >> 1, nb0 = nt * 256 + rank_u * nb_cols;
                        const int nbox = (nb_cols + 63) >> 6;
                        for (int kb = 0; kb < nkb; ++kb) {
                            g_mbar_wait_warp(bempty0 + 8 * s, ephase);
                            if (lane == 0) f_trace_kb(ch, 0, l, nt, kb);
                            const unsigned fb = (bfull0 + 8 * s) & G_PEER_MASK;
                            if (MODE == 0) {
                                if (rank_u == 0) g_mbar_expect_tx_warp(bfull0 + 8 * s, 2u * F_B_BYTES);
                                g_tma_2d_pair_warp(ring + s * F_B_BYTES, &L.map_w, fb, kb * G_BK, nb0);
                            } else {
                                if (rank_u == 0) g_mbar_expect_tx_warp(bfull0 + 8 * s, 2u * (unsigned)nbox * 8192u);
                                for (int j = 0; j < nbox; ++j)
                                    g_tma_2d_pair_warp(ring + s * F_B_BYTES + j * 8192, &L.map_w, fb,
                                                       nb0 + 64 * j, kb * G_BK);
                            }
                            if (++s == F_STAGES) { s = 0; ephase ^= 1u; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (leader CTA)
        // The WHOLE warp runs this loop on warp-uniform values (g_mma_pair_warp elects the
        // issuing lane inside the asm block): see abn_tc_ptx.cuh.
        if ((blockIdx.x & 1) == 0) {
            const unsigned tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
            unsigned aphase = 0, sphase = 0, iter = 0;      // phase bits per barrier
            unsigned s = 0, bphase = 0;                     // B ring slot and its phase
            for (int rb = blockIdx.x >> 1; rb < ch.tiles_m; rb += gridDim.x >> 1, ++iter) {
                for (int l = 0; l < ch.n_layers; ++l) {
                    const FLayer &L = ch.L[l];
                    const int nkb = L.nkb;
                    for (int nt = 0; nt < L.tiles_n; ++nt) {
                        if (lane == 0) f_trace(ch, l, 0 + 4 * nt);
                        g_mbar_wait_warp(aempty0 + 8 * nt, ((aphase >> nt) & 1) ^ 1);   // epilogues drained it
                        aphase ^= 1u << nt;
                        if (lane == 0) f_trace(ch, l, 1 + 4 * nt);
                        g_fence_after();
                        const unsigned idesc = g_idesc(2 * G_BM, f_n_eff(L, nt), 0, MODE);
                        const unsigned d_tmem = tmem_u + nt * 256;
                        unsigned long long da = g_desc(slab, 0);
                        for (int kb = 0; kb < nkb; ++kb, da += (F_SLAB_KB_BYTES >> 4)) {
                            if (nt == 0 && !(ch.debug & 2)) {   // this k-block of the layer's input is in both slabs
                                if (l == 0) {
                                    g_mbar_wait_warp(xfull0 + 8 * kb, iter & 1);
                                } else {
                                    g_mbar_wait_warp(sfull0 + 8 * kb, (sphase >> kb) & 1);
                                    sphase ^= 1u << kb;
                                }
                                g_fence_after();
                            }
                            if (!(ch.debug & 1)) g_mbar_wait_warp(bfull0 + 8 * s, bphase);
                            if (lane == 0) f_trace_kb(ch, 1, l, nt, kb);
                            g_fence_after();
                            const unsigned long long db = g_desc(ring + s * F_B_BYTES, MODE);
                            constexpr unsigned b_step = MODE ? (2048 >> 4) : (32 >> 4);
#pragma unroll
                            for (int k = 0; k < G_BK / G_UK; ++k)
                                g_mma_pair_warp(d_tmem, da + (unsigned long long)(2 * k),
                                                db + (unsigned long long)(b_step * k), idesc, (kb | k) != 0);
                            g_commit_pair_warp(bempty0 + 8 * s);
                            // the layer's LAST tile: once its MMAs on k-block kb retire, nothing reads
                            // slab block kb any more -- the epilogue may overwrite it (kfree)
                            if (nt + 1 == L.tiles_n && nt > 0) g_commit_pair_warp(kfree0 + 8 * kb);
                            if (++s == F_STAGES) { s = 0; bphase ^= 1u; }
                        }
                        g_commit_pair_warp(afull0 + 8 * nt);
                        if (lane == 0) f_trace(ch, l, 2 + 4 * nt);
                    }
                }
                g_commit_pair_warp(slab_free);
            }
        }
    } else {
        // ---------------------------------------------------------------------- epilogue
        const int ew = warp - 2;
        const int wq = warp & 3;                        // TMEM lane quarter of this warp
        const int half = ew >> 2;                       // takes the 64-column blocks cb = half, half + NQ, ..
        const int et = ew * 32 + lane;
        unsigned afphase = 0, kphase = 0, lcount = 0, ycount = 0;
        for (int rb = pair; rb < ch.tiles_m; rb += npairs) {
            const int m0 = (rb * 2 + rank) * G_BM;
            const int row0 = m0 + wq * 32;
            for (int l = 0; l < ch.n_layers; ++l, ++lcount) {
                const FLayer &L = ch.L[l];
                float *bs = bias_s + (lcount & 1u) * 512;
                const int nblk = (L.n_cap + 63) >> 6;
                const unsigned ybuf = ybuf0 + ew * 4096u, ybar = ybar0 + 8 * ew;
                const unsigned long long dkey = (DROP && L.drop.state) ? drop_key(L.drop) : 0ull;
                const long long drow = row0 + lane;
                if (MODE == 0) {
                    for (int c = et; c < 512; c += 32 * EW)
                        bs[c] = (L.bias && c < L.n_out) ? __ldg(L.bias + c) : 0.f;
                    if (et < 32) bias_s[1024 + et] = 0.f;
                }
                // this warp's earlier output boxes have left the slab rows it is about to rewrite
                if (lane == 0) {
                    g_store_wait_read0();
                    if (MODE == 1 && half < nblk) {      // the first block's y_below does not depend on the MMAs
                        g_mbar_expect_tx(ybar, 4096u);
                        g_tma_2d(ybuf, &L.map_y, ybar, half * 64, row0);
                    }
                }
                if (MODE == 0) asm volatile("bar.sync 1, %0;" ::"n"(32 * EW) : "memory");
                else __syncwarp();
                if (et == 0) f_trace(ch, l, 8);
                // accumulator a is read as soon as ITS MMAs are complete; slab block cb (which the
                // layer's last tile still reads) is written once that tile's MMAs on k-block cb have
                // retired -- the `gate` of the earlier accumulators' blocks
                const int a_last = L.tiles_n - 1;
                for (int a = 0; a < L.tiles_n; ++a) {
                    g_mbar_wait(afull0 + 8 * a, (afphase >> a) & 1);
                    afphase ^= 1u << a;
                    g_fence_after();
                    if (et == 0) f_trace(ch, l, 9);
                    for (int cb = 4 * a + half; cb < 4 * a + 4 && cb < nblk; cb += NQ) {
                        // blocks of the earlier accumulators are written while the last tile's MMAs
                        // still read the slab: block cb waits for THEIR k-block cb only (kfree)
                        unsigned gate = 0, gate_parity = 0;
                        if ((MODE == 1 || !L.out_f32) && a < a_last && cb < L.nkb) {
                            gate = kfree0 + 8 * cb;
                            gate_parity = (kphase >> cb) & 1;
                        }
                        const unsigned taddr = tmem + ((unsigned)(wq * 32) << 16) + cb * 64;
                        long long *dbg_tr = (ch.trace && et == 0 && a == 0)
                            ? ch.trace + ((size_t)blockIdx.x * 8 + l) * 16 + 12 : nullptr;
                        if (MODE == 1 || !L.out_f32) {
                            const unsigned dst = slab + cb * F_SLAB_KB_BYTES + wq * 4096u;
                            if (MODE == 0) {
                                const int ones_at = L.ones_col ? L.n_out - cb * 64 : -1;
                                switch (L.act) {
                                    case 1: f_epi_block<1, DROP>(taddr, bs + cb * 64, dst, lane, ones_at, L.drop, dkey, drow, cb * 64, gate, gate_parity, dbg_tr); break;
                                    case 2: f_epi_block<2, DROP>(taddr, bs + cb * 64, dst, lane, ones_at, L.drop, dkey, drow, cb * 64, gate, gate_parity, dbg_tr); break;
                                    case 3: f_epi_block<3, DROP>(taddr, bs + cb * 64, dst, lane, ones_at, L.drop, dkey, drow, cb * 64, gate, gate_parity, dbg_tr); break;
                                    default: f_epi_block<0, DROP>(taddr, bs + cb * 64, dst, lane, ones_at, L.drop, dkey, drow, cb * 64, gate, gate_parity, dbg_tr); break;
                                }
                            } else {
                                g_mbar_wait(ybar, ycount & 1u);
                                ++ycount;
                                uint4 yc[8];
#pragma unroll
                                for (int q = 0; q < 8; ++q)
                                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                                 : "=r"(yc[q].x), "=r"(yc[q].y), "=r"(yc[q].z), "=r"(yc[q].w)
                                                 : "r"(ybuf + lane * 128 + (((unsigned)q ^ (unsigned)(lane & 7)) << 4))
                                                 : "memory");
                                // every lane holds its y_below row: fetch the next box into the same
                                // buffer.  The TMA write is an async-proxy access: program order does
                                // not order it after these generic-proxy reads (with the tensor pipe
                                // saturating shared memory the loads can still be in flight when the
                                // box lands) -- the proxy fence does, lane by lane, before the warp meets
                                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                                __syncwarp();
                                if (lane == 0 && cb + NQ < nblk) {
                                    g_mbar_expect_tx(ybar, 4096u);
                                    g_tma_2d(ybuf, &L.map_y, ybar, (cb + NQ) * 64, row0);
                                }
                                switch (L.act) {
                                    case 1: f_epi_block_d<1, DROP>(taddr, yc, dst, lane, L.drop, dkey, drow, cb * 64, gate, gate_parity); break;
                                    case 2: f_epi_block_d<2, DROP>(taddr, yc, dst, lane, L.drop, dkey, drow, cb * 64, gate, gate_parity); break;
                                    case 3: f_epi_block_d<3, DROP>(taddr, yc, dst, lane, L.drop, dkey, drow, cb * 64, gate, gate_parity); break;
                                    default: f_epi_block_d<0, DROP>(taddr, yc, dst, lane, L.drop, dkey, drow, cb * 64, gate, gate_parity); break;
                                }
                            }
                            // generic-proxy writes -> visible to the async proxy (the next layer's
                            // MMAs and the TMA store), then: k-block cb is ready in this quarter
                            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                            __syncwarp();
                            if (lane == 0) {
                                if (l + 1 < ch.n_layers && cb < ch.L[l + 1].nkb)
                                    f_arrive_leader_release(sfull0 + 8 * cb);
                                g_tma_store_2d(&L.map_out, dst, cb * 64, row0);
                            }
                            if (dbg_tr) { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); dbg_tr[2] = t; }
                        } else if (MODE == 0 && LOSS) {
                            // the embeddings never leave the SM: loss + dz of the output layer here
                            float *exch = bias_s + 1024 + 32;
                            switch (L.act) {
                                case 1: f_epi_loss<1>(ch.loss, ch.rows, L.n_out, L.out32, L.ld32, taddr, bs + cb * 64, exch, wq, cb, nblk, row0, lane); break;
                                case 2: f_epi_loss<2>(ch.loss, ch.rows, L.n_out, L.out32, L.ld32, taddr, bs + cb * 64, exch, wq, cb, nblk, row0, lane); break;
                                case 3: f_epi_loss<3>(ch.loss, ch.rows, L.n_out, L.out32, L.ld32, taddr, bs + cb * 64, exch, wq, cb, nblk, row0, lane); break;
                                default: f_epi_loss<0>(ch.loss, ch.rows, L.n_out, L.out32, L.ld32, taddr, bs + cb * 64, exch, wq, cb, nblk, row0, lane); break;
                            }
                        } else {
                            // the embeddings: fp32 rows, 16-byte stores
                            const int row = row0 + lane;
#pragma unroll
                            for (int hseg = 0; hseg < 2; ++hseg) {
                                float v[32];
                                g_ld32(taddr + 32 * hseg, v);
                                const float *b2 = bs + cb * 64 + 32 * hseg;
                                if (DROP && L.drop.state) {
#pragma unroll
                                    for (int j = 0; j < 32; ++j) v[j] += b2[j];
                                    f_drop32(v, L.drop, dkey, drow, cb * 64 + 32 * hseg);
                                    b2 = bias_s + 1024;              // 32 zeros: the bias is in already
                                }
                                switch (L.act) {
                                    case 1: g_bias_act32<1>(v, b2); break;
                                    case 2: g_bias_act32<2>(v, b2); break;
                                    case 3: g_bias_act32<3>(v, b2); break;
                                    default: g_bias_act32<0>(v, b2); break;
                                }
                                const int gcol0 = cb * 64 + 32 * hseg;
                                if (row < ch.rows) {
                                    float *op = L.out32 + (long long)row * L.ld32 + gcol0;
#pragma unroll
                                    for (int q = 0; q < 8; ++q) {
                                        if (gcol0 + 4 * q + 4 <= L.n_out && (L.ld32 & 3) == 0)
                                            *reinterpret_cast<float4 *>(op + 4 * q) =
                                                make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                                        else
#pragma unroll
                                            for (int e2 = 0; e2 < 4; ++e2)
                                                if (gcol0 + 4 * q + e2 < L.n_out) op[4 * q + e2] = v[4 * q + e2];
                                    }
                                }
                            }
                        }
                    }
                    // this warp is done with accumulator a
                    g_fence_before();
                    __syncwarp();
                    if (lane == 0) g_mbar_arrive_cta0(aempty0 + 8 * a);
                    if (et == 0) f_trace(ch, l, 10 + a);
                }
                // (every kfree barrier of this layer's k-blocks completed one phase, waited for or not)
                if (L.tiles_n > 1) kphase ^= (1u << L.nkb) - 1u;
            }
            // the block's output boxes have been read out of the slab (the next block's x may land)
            if (lane == 0) {
                g_store_wait_read0();
                g_mbar_arrive(drained);
            }
        }
    }
    if (warp >= 2 && lane == 0) g_store_wait_all();
    g_fence_before();
    g_cluster_sync();
    if (threadIdx.x == 0) f_trace(ch, 7, 15);
    if (warp == 2) {
        g_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512)
                     : "memory");
    }
}

}  // namespace abn

using namespace abn;

extern "C" void *abn_gemm_trace_buffer;      // debug hook (abn_tc2.cu)
extern "C" { __attribute__((visibility("default"))) int abn_chain_debug = 0; }      // debug hook: FChain::debug of the next launches

namespace {

int f_sm_count() {
    static int sm_count = 0;
    if (!sm_count) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    }
    return sm_count;
}

template <int MODE, bool DROP, bool LOSS = false>
int f_launch(const FChain &ch, cudaStream_t st, const char *what) {
    // slab + weight ring + barriers + (forward: staged biases | dgrad: y_below boxes) + alignment slack
    constexpr int F_STAGES = MODE == 0 ? F_STAGES_FWD : F_STAGES_DGRAD;
    constexpr unsigned smem = F_KB * F_SLAB_KB_BYTES + F_STAGES * F_B_BYTES + 1024 +
                              (MODE == 0 ? 512 + 2 * 512 * 4 + 128 + (LOSS ? 2048 : 0) : 1024 + F_EW_DGRAD * 4096);
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(mlp_chain_kernel<MODE, DROP, LOSS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem) != cudaSuccess)
            return set_error(ABN_EIO, "%s: cannot reserve %u bytes of shared memory", what, smem);
        configured = true;
    }
    const int sms = f_sm_count();
    int grid = 2 * ch.tiles_m < sms ? 2 * ch.tiles_m : sms;
    grid -= grid % 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(64 + 32 * (MODE == 0 ? F_EW_FWD : F_EW_DGRAD));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_use_pdl() ? 2 : 1;
    cudaLaunchKernelEx(&cfg, mlp_chain_kernel<MODE, DROP, LOSS>, ch);
    return check_launch(what);
}

}  // namespace

static int f_forward(const void *x, int64_t ldx, int64_t rows, const abn_mlp_layer *layers, int n_layers,
                     const abn_mlp_loss *loss, abn_stream_t stream);

extern "C" int abn_mlp_forward_fused(const void *x, int64_t ldx, int64_t rows,
                                     const abn_mlp_layer *layers, int n_layers,
                                     abn_stream_t stream) {
    return f_forward(x, ldx, rows, layers, n_layers, nullptr, stream);
}

extern "C" int abn_mlp_forward_loss_fused(const void *x, int64_t ldx, int64_t rows,
                                          const abn_mlp_layer *layers, int n_layers,
                                          const abn_mlp_loss *loss, abn_stream_t stream) {
    if (!loss || !loss->y || !loss->loss || !loss->dz || (loss->kind != 0 && loss->kind != 1) ||
        (rows & 1) || loss->ld_dz < 8 || (loss->ld_dz & 7) || (reinterpret_cast<uintptr_t>(loss->dz) & 15))
        return set_error(ABN_EINVAL, "abn_mlp_forward_loss_fused: bad loss description (even, interleaved "
                         "rows; dz rows 16-byte aligned)");
    return f_forward(x, ldx, rows, layers, n_layers, loss, stream);
}

static int f_forward(const void *x, int64_t ldx, int64_t rows, const abn_mlp_layer *layers, int n_layers,
                     const abn_mlp_loss *loss, abn_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    if (rows == 0) return ABN_OK;
    if (!x || !layers || rows < 0 || n_layers < 1 || n_layers > F_MAXL)
        return set_error(ABN_EINVAL, "abn_mlp_forward_fused: 1..%d layers", F_MAXL);
    FChain ch;
    memset(&ch, 0, sizeof(ch));
    if (loss) {
        const abn_mlp_layer &last = layers[n_layers - 1];
        if (!last.out_f32 || last.n_out > 128 || loss->ld_dz < last.n_out || last.drop.p > 0.f)
            return set_error(ABN_EINVAL, "abn_mlp_forward_loss_fused: the last layer must write fp32 "
                             "embeddings of at most 128 columns, without dropout");
        ch.loss.on = 1;
        ch.loss.kind = loss->kind;
        ch.loss.write_emb = loss->write_embeddings ? 1 : 0;
        ch.loss.margin = loss->margin;
        ch.loss.scale = loss->scale;
        ch.loss.y = loss->y;
        ch.loss.loss = loss->loss;
        ch.loss.dz = static_cast<__nv_bfloat16 *>(loss->dz);
        ch.loss.ld_dz = loss->ld_dz;
    }
    ch.n_layers = n_layers;
    ch.rows = (int)rows;
    ch.tiles_m = (int)((rows + 2 * G_BM - 1) / (2 * G_BM));
    int rc = g_make_map(&ch.map_x, x, rows, layers[0].n_in, ldx, G_BK, G_BM);
    if (rc) return rc;
    for (int l = 0; l < n_layers; ++l) {
        const abn_mlp_layer &q = layers[l];
        FLayer &L = ch.L[l];
        const bool last = l == n_layers - 1;
        if (!q.W || !q.out || q.n_in <= 0 || q.n_out <= 0 || q.act < 0 || q.act > 3 ||
            q.n_in > F_KB * G_BK || q.n_out + (q.ones_col ? 1 : 0) > F_KB * G_BK ||
            (l > 0 && q.n_in != layers[l - 1].n_out) || (!last && q.out_f32) ||
            (q.out_f32 && q.ones_col))
            return set_error(ABN_EINVAL, "abn_mlp_forward_fused: layer %d: widths up to %d, hidden "
                             "outputs bf16, consecutive layers must fit", l, F_KB * G_BK);
        L.bias = q.bias;
        L.n_in = q.n_in; L.n_out = q.n_out; L.act = q.act;
        L.ones_col = q.ones_col ? 1 : 0; L.out_f32 = q.out_f32 ? 1 : 0;
        L.drop = drop_args(&q.drop);
        L.n_cap = q.n_out + L.ones_col;
        L.nkb = (q.n_in + G_BK - 1) / G_BK;
        L.tiles_n = (L.n_cap + 255) / 256;
        rc = g_make_map(&L.map_w, q.W, q.n_out, q.n_in, q.ldw, G_BK, 128);
        if (rc) return rc;
        if (L.out_f32) {
            L.out32 = static_cast<float *>(q.out);
            L.ld32 = q.ldo;
        } else {
            if ((q.ldo & 7) || q.ldo < L.n_cap)
                return set_error(ABN_EINVAL, "abn_mlp_forward_fused: layer %d: bf16 output rows must be "
                                 "padded to a multiple of 8 elements covering n_out%s", l,
                                 L.ones_col ? " + 1" : "");
            rc = g_make_map(&L.map_out, q.out, rows, q.ldo, q.ldo, 64, 32);
            if (rc) return rc;
        }
    }
    ch.trace = reinterpret_cast<long long *>(abn_gemm_trace_buffer);
    ch.debug = abn_chain_debug;
    bool any_drop = false;
    for (int l = 0; l < n_layers; ++l) any_drop |= ch.L[l].drop.state != nullptr;
    if (loss) {
        if (any_drop) return f_launch<0, true, true>(ch, (cudaStream_t)stream, "abn_mlp_forward_loss_fused");
        return f_launch<0, false, true>(ch, (cudaStream_t)stream, "abn_mlp_forward_loss_fused");
    }
    return any_drop ? f_launch<0, true>(ch, (cudaStream_t)stream, "abn_mlp_forward_fused")
                    : f_launch<0, false>(ch, (cudaStream_t)stream, "abn_mlp_forward_fused");
}

extern "C" int abn_mlp_dgrad_fused(const void *dz_top, int64_t ld_top, int64_t rows,
                                   const abn_mlp_dlayer *layers, int n_layers,
                                   abn_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    if (rows == 0 || n_layers == 0) return ABN_OK;
    if (!dz_top || !layers || rows < 0 || n_layers < 0 || n_layers > F_MAXL)
        return set_error(ABN_EINVAL, "abn_mlp_dgrad_fused: up to %d layers", F_MAXL);
    FChain ch;
    memset(&ch, 0, sizeof(ch));
    ch.n_layers = n_layers;
    ch.rows = (int)rows;
    ch.tiles_m = (int)((rows + 2 * G_BM - 1) / (2 * G_BM));
    // GEMM view of layer l: A = dz [rows, K = n_out], B = W [K = n_out, N = n_in] as stored
    int rc = g_make_map(&ch.map_x, dz_top, rows, layers[0].n_out, ld_top, G_BK, G_BM);
    if (rc) return rc;
    for (int l = 0; l < n_layers; ++l) {
        const abn_mlp_dlayer &q = layers[l];
        FLayer &L = ch.L[l];
        if (!q.W || !q.y_below || !q.dz_below || q.n_in <= 0 || q.n_out <= 0 || q.act_below < 0 ||
            q.act_below > 3 || q.n_in > F_KB * G_BK || q.n_out > F_KB * G_BK ||
            (l > 0 && q.n_out != layers[l - 1].n_in))
            return set_error(ABN_EINVAL, "abn_mlp_dgrad_fused: layer %d: widths up to %d, consecutive "
                             "layers must fit", l, F_KB * G_BK);
        if ((q.ld_dz & 7) || q.ld_dz < q.n_in || (q.ld_y & 7) || q.ld_y < q.n_in)
            return set_error(ABN_EINVAL, "abn_mlp_dgrad_fused: layer %d: bf16 rows must be padded to a "
                             "multiple of 8 elements covering n_in", l);
        L.n_in = q.n_in; L.n_out = q.n_out; L.act = q.act_below;
        L.drop = drop_args(&q.drop_below);
        L.n_cap = q.n_in;
        L.nkb = (q.n_out + G_BK - 1) / G_BK;
        L.tiles_n = (L.n_cap + 255) / 256;
        rc = g_make_map(&L.map_w, q.W, q.n_out, q.n_in, q.ldw, 64, 64);
        if (rc) return rc;
        rc = g_make_map(&L.map_out, q.dz_below, rows, q.ld_dz, q.ld_dz, 64, 32);
        if (rc) return rc;
        rc = g_make_map(&L.map_y, q.y_below, rows, q.ld_y, q.ld_y, 64, 32);
        if (rc) return rc;
    }
    ch.trace = reinterpret_cast<long long *>(abn_gemm_trace_buffer);
    ch.debug = abn_chain_debug;
    bool any_drop = false;
    for (int l = 0; l < n_layers; ++l) any_drop |= ch.L[l].drop.state != nullptr;
    return any_drop ? f_launch<1, true>(ch, (cudaStream_t)stream, "abn_mlp_dgrad_fused")
                    : f_launch<1, false>(ch, (cudaStream_t)stream, "abn_mlp_dgrad_fused");
}
