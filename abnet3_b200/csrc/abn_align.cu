// abn_align.cu -- kernels (1) cosine frame distance and (2) DTW wavefront + traceback, as TWO
// stages per pair: a distance kernel leaves the pair's float32 distance matrix in a "skew"
// (anti-diagonal-major) layout in a device workspace slot, a DTW kernel (one warp per pair) reads
// it back with coalesced loads.  (Round 1 began with one fused kernel per pair; three of its four
// warps idled while warp 0 swept the wavefront: 22.8 M -> 27.8 M pairs/s for the split.)
//
// Reference behaviour reproduced (paths relative to /root/reference):
//   cosine_distance     abnet3/utils.py:40-60   (float32 arithmetic, zero-norm
//                       rules :55-58, NaN / negative -> pair invalid :59)
//   DTW + traceback     external dtw.DTW called at abnet3/utils.py:149-151;
//                       recurrence and tie rule as oracle/dtw_oracle.c
//   get_dtw_alignment   abnet3/utils.py:147-153, batched over the pair list
//                       like abnet3/dataloader.py:183-206 / :642-653
//
// Design
// ------
// Token pairs are ragged (20-80 frames in the canonical corpus), and the distance stage is a
// small dense contraction per pair, so pairs are first bucketed on the device into SIZE CLASSES
// (ceil(n1/16), ceil(n2/16)); every class has its own kernel instantiation with a right-sized
// register tile and shared-memory footprint (more resident CTAs for short tokens, no wasted FMAs
// on padding beyond 16 frames).
//   generic class kernel (128 threads, persistent over the class's pair list):
//     HBM --cp.async 16 B--> smem K-chunks of both tokens (double buffered) --LDS.128-->
//     register-tiled fp32 FMA, accumulated per 40-wide chunk in a fixed order --> reciprocal
//     norms, acosf --> D in smem in skew layout --> 16-byte copies to the hand-over slot
//   stacked class kernel (7 x 40 frame stacks): one 160-byte 1-D TMA copy per extended frame,
//     40-deep Gram tile, 7-tap diagonal sums in the generic kernel's order (same bits)
//   DTW kernel (dtw_skew_kernel<G>): lane l owns rows G l .. G l + G - 1, float64 recurrence
//     with the oracle's tie order, 2-bit directions in smem, lane 0 walks them back, all lanes
//     write the global frame-index pairs
//   long tokens: long_tile_kernel (one CTA per tile of a pair's matrix) + dtw_band_kernel
//     (128-row bands, directions in global memory)
#include "abn_common.cuh"

namespace abn {

constexpr int AL_THREADS = 128;
constexpr int KC = 40;           // floats of K staged per chunk (one fbank frame of the 7-stack)
constexpr int KCP = KC + 4;      // smem row stride: 11 x 16 B, odd => LDS.128 conflict-free
constexpr int NM_SHORT = 96;                     // longest token of the fused class kernels (6 x 16)
constexpr int NM_LIMIT = ABN_MAX_TOKEN_FRAMES;   // longest token overall (tiled kernel)
constexpr int NCLS_SIDE = NM_SHORT / 16;         // 6
constexpr int NCLS = NCLS_SIDE * NCLS_SIDE;      // 36 size classes
constexpr int CLS_LONG = NCLS;                   // pseudo class: a token longer than NM_SHORT
constexpr int CLS_INVALID = NCLS + 1;            // pseudo class: invalid shape
constexpr int NBINS = NCLS + 2;
constexpr float PI_F = 3.14159274101257324f;     // float32(np.pi)

// workspace layout (bytes): counts[38] | class_off[39] | cursor[38] | order[n_pairs]
constexpr size_t WS_COUNTS = 0, WS_OFF = 256, WS_CURSOR = 512, WS_ORDER = 1024;

struct AlignArgs {
    const float *feat; int64_t n_rows; int dim;
    const int32_t *pair_tok; int n_pairs;
    const int64_t *path_off; int32_t *idx1; int32_t *idx2; int32_t *path_len;
    double *cost; uint8_t *valid;
    const int64_t *dist_off; float *dist_out;      // distance-only mode when dist_out != null
    const int32_t *order; const int32_t *class_off;
    int stack;     // 0: generic rows; S > 1: rows are S-frame stacks of dim/S-wide frames
    // distance -> DTW hand-over: the pair at sorted position `it` of the window [w0, w1)
    // keeps its distance matrix at dws + (it - w0) * slot_cells, in skew layout (below)
    float *dws; int w0, w1, slot_cells;
};

// Skew layout of an n1 x n2 matrix: anti-diagonal t = i + j holds its cells
// contiguously, rows ascending, diagonals back to back; cell (i, j) lives at
// skew_base(i + j) + i.  The DTW wavefront reads one diagonal per step, one row per lane:
// a contiguous, coalesced segment.  Diagonals 0 .. n1-2 are followed by one +inf GUARD
// float: it is what the lane of row t + 1 reads at step t -- the cell (t + 1, -1) left of
// the matrix -- so the wavefront needs no range test on its loads (every other
// out-of-range read lands on some finite distance or guard of the same slot, and no
// in-range cell ever consumes what is computed from it).  n1 * n2 + n1 - 1 floats.
__host__ __device__ __forceinline__ int skew_base(int t, int n1, int n2) {
    const int m = n1 < n2 ? n1 : n2, mx = n1 < n2 ? n2 : n1, T = n1 + n2 - 1;
    int off;
    if (t <= m) off = t * (t + 1) / 2;
    else if (t <= mx) off = m * (m + 1) / 2 + (t - m) * m;
    else off = n1 * n2 - (T - t) * (T - t + 1) / 2;
    const int ilo = t - n2 + 1 > 0 ? t - n2 + 1 : 0;
    const int guards = t < n1 - 1 ? t : n1 - 1;          // guards of the diagonals before t
    return off + guards - ilo;
}
__host__ __device__ __forceinline__ int skew_cells(int n1, int n2) { return n1 * n2 + n1 - 1; }
constexpr int SKEW_SLACK = 128;      // floats a wavefront may read past its matrix

__host__ __device__ constexpr unsigned a16(unsigned x) { return (x + 15u) & ~15u; }

template <int RA, int NCG>
struct ClassLayout {
    static constexpr int ROWS_A = 16 * RA, ROWS_B = 16 * NCG;
    static constexpr int LDD = ROWS_B + 2;   // (LDD-1) odd: the wavefront's lane stride
    static constexpr unsigned STAGE_BYTES = (ROWS_A + ROWS_B) * KCP * 4u;
    static constexpr unsigned DIRS_OFF = a16(ROWS_A * LDD * 4u);
    static constexpr unsigned ALIAS_END = DIRS_OFF + ROWS_A * ROWS_B;
    static constexpr unsigned REGION = 2u * STAGE_BYTES > ALIAS_END ? 2u * STAGE_BYTES : ALIAS_END;
    static constexpr unsigned NORMS_OFF = a16(REGION);
    static constexpr unsigned PATH_OFF = NORMS_OFF + (ROWS_A + ROWS_B) * 4u;
    static constexpr unsigned MISC_OFF = a16(PATH_OFF + 2u * (ROWS_A + ROWS_B));
    static constexpr unsigned TOTAL = MISC_OFF + 32u;
};

// class kernels: no row-major D, no DTW state -- the matrix leaves in skew layout
template <int RA, int NCG>
struct DistLayout {
    static constexpr int ROWS_A = 16 * RA, ROWS_B = 16 * NCG;
    static constexpr int LDD = ROWS_B + 2;
    static constexpr unsigned STAGE_BYTES = (ROWS_A + ROWS_B) * KCP * 4u;
    static constexpr unsigned D_BYTES = (ROWS_A * ROWS_B + ROWS_A) * 4u;
    static constexpr unsigned REGION = 2u * STAGE_BYTES > D_BYTES ? 2u * STAGE_BYTES : D_BYTES;
    static constexpr unsigned NORMS_OFF = a16(REGION);
    static constexpr unsigned TB_OFF = NORMS_OFF + (ROWS_A + ROWS_B) * 4u;
    static constexpr unsigned TOTAL = TB_OFF + (ROWS_A + ROWS_B) * 4u + 16u;
};

__device__ __forceinline__ unsigned smem_u32(const void *p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ float4 lds128(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}

// ---------------------------------------------------------------- staging --
// thread -> (16-byte piece = tid & 15, row = tid >> 4): a warp copies two rows
// of 160 contiguous bytes each per instruction
__device__ __forceinline__ void stage_chunk(unsigned buf_addr, const float *g1, const float *g2,
                                            int n1, int n2, int rows_a, int dim, int k0, int kc4,
                                            int tid) {
    const int piece = tid & 15, r0 = tid >> 4;
    if (piece < kc4) {
        const float *s1 = g1 + k0 + piece * 4;
        const unsigned d1 = buf_addr + piece * 16;
        for (int r = r0; r < n1; r += AL_THREADS / 16)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d1 + r * (KCP * 4)),
                         "l"(s1 + (size_t)r * dim) : "memory");
        const float *s2 = g2 + k0 + piece * 4;
        const unsigned d2 = buf_addr + rows_a * (KCP * 4) + piece * 16;
        for (int r = r0; r < n2; r += AL_THREADS / 16)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d2 + r * (KCP * 4)),
                         "l"(s2 + (size_t)r * dim) : "memory");
    }
}

__device__ __forceinline__ float row_sumsq(unsigned row_addr, int kc4) {
    float acc = 0.f;
    for (int k4 = 0; k4 < kc4; ++k4) {
        const float4 v = lds128(row_addr + k4 * 16);
        acc = fmaf(v.x, v.x, acc);
        acc = fmaf(v.y, v.y, acc);
        acc = fmaf(v.z, v.z, acc);
        acc = fmaf(v.w, v.w, acc);
    }
    return acc;
}

// utils.py:47-58 for one cell, in float32.  rx, ry are the RECIPROCAL row norms (0 for
// a zero row, which is how the zero-norm rules are recognised): cos = dot * (rx * ry)
// and arccos * (1/pi) replace the reference's two divisions -- at most 2 ulp apart in
// cos, well inside the parity bound, and ~30 instructions cheaper.
constexpr float INV_PI_F = 0.318309873342514038f;       // float32(1 / pi)
__device__ __forceinline__ float recip_norm(float sumsq) {
    const float n = sqrtf(sumsq);
    return n == 0.f ? 0.f : __frcp_rn(n);
}
__device__ __forceinline__ float cell_distance(float dot, float rx, float ry) {
    const float cs = __fmul_rn(dot, __fmul_rn(rx, ry));
    const float d = __fmul_rn(acosf(cs), INV_PI_F);
    if (rx == 0.f || ry == 0.f) return (rx == 0.f && ry == 0.f) ? 0.f : 1.f;
    return d;
}

__device__ __forceinline__ void skew_guards(float *Dsk, const int *tb, int n1, int tid);

// ------------------------------------------------------- distance (kernel 1)
// Thread grid 16 (rows) x 8 (cols): thread (ti, tj) owns rows ti + 16 r and
// columns tj + 8 c.  Inside a warp that is 8 consecutive rows x 4 consecutive
// columns, so every LDS.128 is one conflict-free wavefront with broadcast.
// Every dot product is the sum over the K chunks of a sequential fp32 FMA
// chain over the chunk: fixed order, and ~3x tighter than one long chain.
template <int RA, int NCG, typename L>
__device__ __forceinline__ void pair_distance(unsigned char *smem, const float *g1,
                                              const float *g2, int n1, int n2, int dim,
                                              float *dist_gmem, int ld_gmem, const int *tb,
                                              int &bad) {
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int ti = (warp >> 1) * 8 + (lane >> 2);
    const int tj = (warp & 1) * 4 + (lane & 3);
    const unsigned sbase = smem_u32(smem);
    float *norms = reinterpret_cast<float *>(smem + L::NORMS_OFF);
    float *Ds = reinterpret_cast<float *>(smem);

    float tot[RA][2 * NCG];
#pragma unroll
    for (int r = 0; r < RA; ++r)
#pragma unroll
        for (int c = 0; c < 2 * NCG; ++c) tot[r][c] = 0.f;

    // row-norm ownership: combined row index q in [0, n1 + n2)
    const int q0 = tid, q1 = tid + AL_THREADS;
    const int nq = n1 + n2;
    const int row0 = q0 < n1 ? q0 : L::ROWS_A + (q0 - n1);
    const int row1 = q1 < n1 ? q1 : L::ROWS_A + (q1 - n1);
    float ss0 = 0.f, ss1 = 0.f;

    const int nchunks = (dim + KC - 1) / KC;
    stage_chunk(sbase, g1, g2, n1, n2, L::ROWS_A, dim, 0, min(KC, dim) / 4, tid);
    cp_async_commit();
    for (int ch = 0; ch < nchunks; ++ch) {
        const int k0 = ch * KC;
        const int kc4 = min(KC, dim - k0) / 4;
        if (ch + 1 < nchunks) {
            const int k1 = k0 + KC;
            stage_chunk(sbase + ((ch + 1) & 1) * L::STAGE_BYTES, g1, g2, n1, n2, L::ROWS_A, dim, k1,
                        min(KC, dim - k1) / 4, tid);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const unsigned buf = sbase + (ch & 1) * L::STAGE_BYTES;
        if (q0 < nq) ss0 += row_sumsq(buf + row0 * (KCP * 4), kc4);
        if (q1 < nq) ss1 += row_sumsq(buf + row1 * (KCP * 4), kc4);

        float acc[RA][2 * NCG];
#pragma unroll
        for (int r = 0; r < RA; ++r)
#pragma unroll
            for (int c = 0; c < 2 * NCG; ++c) acc[r][c] = 0.f;
        const unsigned a_addr = buf + ti * (KCP * 4);
        const unsigned b_addr = buf + (L::ROWS_A + tj) * (KCP * 4);
#pragma unroll 2
        for (int k4 = 0; k4 < kc4; ++k4) {
            float4 a[RA];
#pragma unroll
            for (int r = 0; r < RA; ++r) a[r] = lds128(a_addr + (16 * r) * (KCP * 4) + k4 * 16);
#pragma unroll
            for (int cg = 0; cg < NCG; ++cg) {
                float4 b[2];
#pragma unroll
                for (int c = 0; c < 2; ++c)
                    b[c] = lds128(b_addr + (8 * (2 * cg + c)) * (KCP * 4) + k4 * 16);
#pragma unroll
                for (int r = 0; r < RA; ++r)
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        float s = acc[r][2 * cg + c];
                        s = fmaf(a[r].x, b[c].x, s);
                        s = fmaf(a[r].y, b[c].y, s);
                        s = fmaf(a[r].z, b[c].z, s);
                        s = fmaf(a[r].w, b[c].w, s);
                        acc[r][2 * cg + c] = s;
                    }
            }
        }
#pragma unroll
        for (int r = 0; r < RA; ++r)
#pragma unroll
            for (int c = 0; c < 2 * NCG; ++c) tot[r][c] += acc[r][c];
        __syncthreads();   // everyone done with this stage before it is refilled
    }
    if (q0 < nq) norms[row0] = recip_norm(ss0);
    if (q1 < nq) norms[row1] = recip_norm(ss1);
    __syncthreads();

    // epilogue: utils.py:47-58 in float32; the staging buffers are dead, D may overwrite them
    if (tb) skew_guards(Ds, tb, n1, tid);
#pragma unroll
    for (int r = 0; r < RA; ++r) {
        const int i = ti + 16 * r;
        if (i >= n1) continue;
        const float xn = norms[i];
#pragma unroll
        for (int c = 0; c < 2 * NCG; ++c) {
            const int j = tj + 8 * c;
            if (j >= n2) continue;
            const float yn = norms[L::ROWS_A + j];
            const float d = cell_distance(tot[r][c], xn, yn);
            if (!(d >= 0.f)) bad = 1;
            if (dist_gmem) dist_gmem[(size_t)i * ld_gmem + j] = d;
            else if (tb) Ds[tb[i + j] + i] = d;
            else Ds[i * L::LDD + j] = d;
        }
    }
}

// every thread: the pair's skew-base table, and after the matrix is complete its
// contiguous copy to the hand-over slot (128-bit accesses)
__device__ __forceinline__ void skew_table(int *tb, int n1, int n2, int tid) {
    for (int t = tid; t < n1 + n2 - 1; t += AL_THREADS) tb[t] = skew_base(t, n1, n2);
}
// the +inf guards after diagonals 0 .. n1-2 (call once the staging area under D is dead)
__device__ __forceinline__ void skew_guards(float *Dsk, const int *tb, int n1, int tid) {
    for (int t = tid; t < n1 - 1; t += AL_THREADS) Dsk[tb[t] + t + 1] = __int_as_float(0x7f800000);
}
__device__ __forceinline__ void skew_copy_out(const float *Dsk, float *dst, int cells, int tid) {
    const float4 *s4 = reinterpret_cast<const float4 *>(Dsk);
    float4 *d4 = reinterpret_cast<float4 *>(dst);
    for (int e = tid; e < (cells >> 2); e += AL_THREADS) d4[e] = s4[e];
    for (int e = (cells & ~3) + tid; e < cells; e += AL_THREADS) dst[e] = Dsk[e];
}

// ------------------------------------------------------------ DTW (kernel 2)
// One warp per pair.  Lane l owns rows G*l .. G*l+G-1; at step t every row i
// handles cell (i, t - i): an anti-diagonal sweep.  Within a lane the upper
// neighbour is a register; across lanes it is one fp64 shuffle per step.
// C[i,j] = D[i,j] + min(C[i-1,j-1], C[i-1,j], C[i,j-1]) with the oracle's tie
// order (diag, up, left): one add per cell, so results are bit-identical to
// the sequential recurrence for any float64 D.  The body is branch-free:
// out-of-range cells compute on a clamped address and are not committed.
template <typename DT, int G>
__device__ __forceinline__ double dtw_wavefront(const DT *D, int ldd, uint8_t *dirs, int ldr,
                                                int n1, int n2, int lane) {
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    double cur[G], prev[G];
#pragma unroll
    for (int g = 0; g < G; ++g) cur[g] = prev[g] = INF;
    double nbprev = INF;
    const int i0 = lane * G;
    const int T = n1 + n2 - 1;
    for (int t = 0; t < T; ++t) {
        double up0 = __shfl_up_sync(0xffffffffu, cur[G - 1], 1);
        up0 = lane == 0 ? INF : up0;
        const double dg0 = nbprev;
        nbprev = up0;
        double nw[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int i = i0 + g, j = t - i;
            const bool ok = (i < n1) & ((unsigned)j < (unsigned)n2);
            const int cell = ok ? i * ldd + j : 0;
            const double d = (double)D[cell];
            const double up = g == 0 ? up0 : cur[g - 1];
            const double dg = g == 0 ? dg0 : prev[g - 1];
            const double lf = cur[g];
            const bool up_le = up <= lf;
            const double m1 = up_le ? up : lf;
            const bool use_dg = dg <= m1;          // dg <= up && dg <= lf
            double m = use_dg ? dg : m1;
            const uint8_t dir = use_dg ? DIR_DIAG : (up_le ? DIR_UP : DIR_LEFT);
            m = (i | j) == 0 ? 0.0 : m;
            nw[g] = ok ? d + m : lf;
            if (ok) dirs[i * ldr + j] = dir;
        }
#pragma unroll
        for (int g = 0; g < G; ++g) { prev[g] = cur[g]; cur[g] = nw[g]; }
    }
    // C[n1-1, n2-1] lives in lane (n1-1)/G, slot (n1-1)%G
    double c = 0.0;
#pragma unroll
    for (int g = 0; g < G; ++g) c = ((n1 - 1) % G == g) ? cur[g] : c;
    return __shfl_sync(0xffffffffu, c, (n1 - 1) / G);
}

// lane 0: follow the stored directions from (n1-1, n2-1) back to (0, 0)
__device__ __forceinline__ int traceback(const uint8_t *dirs, int ldr, int n1, int n2,
                                         uint8_t *pb_i, uint8_t *pb_j) {
    int i = n1 - 1, j = n2 - 1, L = 1;
    pb_i[0] = (uint8_t)i; pb_j[0] = (uint8_t)j;
    while ((i | j) != 0) {
        const uint8_t d = dirs[i * ldr + j];
        i -= (d != DIR_LEFT);
        j -= (d != DIR_UP);
        pb_i[L] = (uint8_t)i; pb_j[L] = (uint8_t)j; ++L;
    }
    return L;
}

// ------------------------------------------------- DTW kernel (one warp per pair)
// Reads the skew-layout distance matrix a class kernel left in the hand-over slot:
// at step t lane l needs cells (G l + g, t - G l - g) = skew_base(t) + G l + g, G
// coalesced 4-byte loads, prefetched one 4-step block ahead.  Same float64 recurrence
// and tie order as dtw_wavefront above (C[-1][-1] = 0 replaces the (0,0) special case;
// out-of-range cells carry +inf instead of being skipped -- no in-range cell ever
// reads them).  Directions: 2 bits per cell, one byte per (row, 4 steps), in this
// warp's slice of shared memory; lane 0 walks them back, all lanes write the indices.
constexpr int DTW_WARPS = 4;

template <int G>
__global__ void __launch_bounds__(DTW_WARPS * 32)
dtw_skew_kernel(const AlignArgs a, int cls_beg, int cls_end, int t4_cap) {
    constexpr int R = 32 * G;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *mine = smem + warp * (t4_cap * R + 8 * t4_cap);
    uint8_t *dirs = mine;
    uint8_t *pb_i = mine + t4_cap * R;
    uint8_t *pb_j = pb_i + 4 * t4_cap;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    const float INF_F = __int_as_float(0x7f800000);
    const int beg = max(a.class_off[cls_beg], a.w0), end = min(a.class_off[cls_end], a.w1);
    const int i0 = lane * G;

    for (int it = beg + blockIdx.x * DTW_WARPS + warp; it < end; it += gridDim.x * DTW_WARPS) {
        const int p = a.order[it];
        if (!a.valid[p]) continue;            // NaN in the matrix: the class kernel dropped the pair
        const int4 tk = reinterpret_cast<const int4 *>(a.pair_tok)[p];
        const int s1 = tk.x, n1 = tk.y, s2 = tk.z, n2 = tk.w;
        const float *__restrict__ D = a.dws + (size_t)(it - a.w0) * a.slot_cells;
        const int T = n1 + n2 - 1, m = min(n1, n2);

        double cur[G], prev[G];
#pragma unroll
        for (int g = 0; g < G; ++g) cur[g] = prev[g] = INF;
        double nbprev = lane == 0 ? 0.0 : INF;          // C[-1][-1] = 0 feeds cell (0, 0)
        const float *Dl = D + i0;                       // this lane's rows
        asm volatile("" : "+l"(Dl));                    // keep it one register pair: loads are Dl[b + g]
        // skew bases of 32 consecutive steps live one per lane; a step's base is one shuffle
        int bases = 0;
#define ABN_DTW_LOAD(dst, t_first)                                                      \
        {                                                                               \
            if (((t_first) & 31) == 0) bases = skew_base(min((t_first) + lane, T - 1), n1, n2); \
            _Pragma("unroll") for (int u = 0; u < 4; ++u) {                             \
                const int b = __shfl_sync(FULL, bases, ((t_first) + u) & 31);           \
                _Pragma("unroll") for (int g = 0; g < G; ++g) dst[u][g] = __ldg(Dl + b + g); \
            }                                                                           \
        }
#define ABN_DTW_STEP(dc, u)                                                             \
        {                                                                               \
            double up0 = __shfl_up_sync(FULL, cur[G - 1], 1);                           \
            up0 = lane == 0 ? INF : up0;                                                \
            const double dg0 = nbprev;                                                  \
            nbprev = up0;                                                               \
            double nw[G];                                                               \
            _Pragma("unroll") for (int g = 0; g < G; ++g) {                             \
                const double d = (double)dc[u][g];                                      \
                const double up = g == 0 ? up0 : cur[g > 0 ? g - 1 : 0];                \
                const double dg = g == 0 ? dg0 : prev[g > 0 ? g - 1 : 0];               \
                const double lf = cur[g];                                               \
                const bool up_le = up <= lf;                                            \
                const double m1 = up_le ? up : lf;                                      \
                const bool use_dg = dg <= m1;          /* dg <= up && dg <= lf */       \
                const double mm = use_dg ? dg : m1;                                     \
                const unsigned dir = use_dg ? DIR_DIAG : (up_le ? DIR_UP : DIR_LEFT);   \
                nw[g] = d + mm;                                                         \
                bits[g] |= dir << (2 * (u));                                            \
            }                                                                           \
            _Pragma("unroll") for (int g = 0; g < G; ++g) { prev[g] = cur[g]; cur[g] = nw[g]; } \
        }
#define ABN_DTW_BLOCK(dc, dnext, t0)                                                    \
        {                                                                               \
            ABN_DTW_LOAD(dnext, (t0) + 4)                                               \
            unsigned bits[G];                                                           \
            _Pragma("unroll") for (int g = 0; g < G; ++g) bits[g] = 0;                  \
            if ((t0) + 4 <= T) {          /* warp-uniform: a full block, no per-step test */ \
                ABN_DTW_STEP(dc, 0) ABN_DTW_STEP(dc, 1) ABN_DTW_STEP(dc, 2) ABN_DTW_STEP(dc, 3) \
            } else {                      /* the last, partial block */                 \
                ABN_DTW_STEP(dc, 0)                                                     \
                if ((t0) + 1 < T) ABN_DTW_STEP(dc, 1)                                   \
                if ((t0) + 2 < T) ABN_DTW_STEP(dc, 2)                                   \
            }                                                                           \
            _Pragma("unroll") for (int g = 0; g < G; ++g)                               \
                dirs[((t0) >> 2) * R + i0 + g] = (uint8_t)bits[g];                      \
        }
        float da[4][G], db[4][G];
        ABN_DTW_LOAD(da, 0)
        for (int t0 = 0; t0 < T; t0 += 8) {
            ABN_DTW_BLOCK(da, db, t0)
            if (t0 + 4 < T) ABN_DTW_BLOCK(db, da, t0 + 4)
        }
#undef ABN_DTW_BLOCK
#undef ABN_DTW_STEP
#undef ABN_DTW_LOAD
        const int glast = (n1 - 1) % G;
        double c = cur[0];
        if (G > 1) c = glast == 1 ? cur[G > 1 ? 1 : 0] : c;
        if (G > 2) c = glast == 2 ? cur[G > 2 ? 2 : 0] : c;
        c = __shfl_sync(FULL, c, (n1 - 1) / G);
        __syncwarp();
        int len = 0;
        if (lane == 0) {
            int i = n1 - 1, j = n2 - 1;
            len = 1;
            pb_i[0] = (uint8_t)i; pb_j[0] = (uint8_t)j;
            while ((i | j) != 0 && len < T) {
                const int t = i + j;
                const unsigned d = ((unsigned)dirs[(t >> 2) * R + i] >> (2 * (t & 3))) & 3u;
                i -= (d != DIR_LEFT);
                j -= (d != DIR_UP);
                pb_i[len] = (uint8_t)i; pb_j[len] = (uint8_t)j; ++len;
            }
            a.path_len[p] = len;
            a.cost[p] = c;
        }
        len = __shfl_sync(FULL, len, 0);
        __syncwarp();
        const int64_t off = a.path_off[p];
        for (int k = lane; k < len; k += 32) {
            a.idx1[off + k] = s1 + (int)pb_i[len - 1 - k];
            a.idx2[off + k] = s2 + (int)pb_j[len - 1 - k];
        }
        __syncwarp();          // the slice is reused by this warp's next pair
    }
}

// ------------------------------------------------------------ class kernel --
// distance matrix of every pair of one size class -> skew layout -> hand-over slot
// (or, for abn_cosine_distance, row-major straight to the caller's buffer)
template <int RA, int NCG>
__global__ void __launch_bounds__(AL_THREADS)
align_class_kernel(const AlignArgs a) {
    using L = DistLayout<RA, NCG>;
    constexpr int CLS = (RA - 1) * NCLS_SIDE + (NCG - 1);
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x;
    const int beg = max(a.class_off[CLS], a.w0), end = min(a.class_off[CLS + 1], a.w1);
    const float *Dsk = reinterpret_cast<const float *>(smem);
    int *tb = reinterpret_cast<int *>(smem + L::TB_OFF);

    for (int it = beg + blockIdx.x; it < end; it += gridDim.x) {
        const int p = a.order[it];
        const int4 tk = reinterpret_cast<const int4 *>(a.pair_tok)[p];
        const int s1 = tk.x, n1 = tk.y, s2 = tk.z, n2 = tk.w;
        if (!a.dist_out) skew_table(tb, n1, n2, tid);
        int bad = 0;
        pair_distance<RA, NCG, L>(smem, a.feat + (size_t)s1 * a.dim, a.feat + (size_t)s2 * a.dim,
                                  n1, n2, a.dim,
                                  a.dist_out ? a.dist_out + a.dist_off[p] : nullptr, n2,
                                  a.dist_out ? nullptr : tb, bad);
        bad = __syncthreads_or(bad);
        if (tid == 0) {
            a.valid[p] = bad ? 0 : 1;     // utils.py:59 assert fails -> dataloader.py:190-191 drops the pair
            if (bad && !a.dist_out) { a.path_len[p] = 0; a.cost[p] = nan(""); }
        }
        if (!a.dist_out && !bad)
            skew_copy_out(Dsk, a.dws + (size_t)(it - a.w0) * a.slot_cells, skew_cells(n1, n2), tid);
        __syncthreads();   // smem is reused by the next pair
    }
}


// --------------------------------------- stacked-feature fast path
// When the table is an S = 7 frame stack of 40-wide frames (abnet3/features.py:135-159:
// row t = [x[t-3] .. x[t+3]], zeros outside the file), the 280-long dot product of
// rows i and j is a 7-tap DIAGONAL sum of 40-long dot products of un-stacked frames:
//     <row_i, row_j> = sum_{c=0..6} <x[i+c-3], y[j+c-3]> = sum_c G40[i+c][j+c]
// over the token's n + 6 "extended" frames, which all live inside the token's own
// rows (middle block of every row, left blocks of the first row, right blocks of the
// last one).  The generic kernel accumulates per 40-wide chunk in exactly this order,
// so this path returns the SAME BITS with 7x fewer FMAs and 6-7x fewer bytes read.
// The caller vouches for the structure (abn_stack_violations checks a table).
constexpr int STACK_S = 7, STACK_H = STACK_S / 2, STACK_F = KC;
constexpr int STACK_MAXN = NM_SHORT - 2 * STACK_H;       // 90: longest token of this path

template <int RA, int NCG>      // extended sizes: 16 RA >= n1 + 6, 16 NCG >= n2 + 6
struct StackLayout {
    static constexpr int ROWS_A = 16 * RA, ROWS_B = 16 * NCG;
    static constexpr int LDG = ROWS_B + 4;   // = 4 or 20 (mod 32): the diagonal reads are conflict-free
    static constexpr unsigned STAGE_BYTES = (ROWS_A + ROWS_B) * KCP * 4u;
    static constexpr unsigned D_BYTES = (ROWS_A * ROWS_B + ROWS_A) * 4u;
    static constexpr unsigned REGION0 = STAGE_BYTES > D_BYTES ? STAGE_BYTES : D_BYTES;
    static constexpr unsigned G_OFF = a16(REGION0);
    static constexpr int G_ROWS = ROWS_A + 8;      // the epilogue's diagonal runs may read past the tile
    static constexpr unsigned N40_OFF = a16(G_OFF + G_ROWS * LDG * 4u);
    static constexpr unsigned NORMS_OFF = N40_OFF + (ROWS_A + ROWS_B) * 4u;
    static constexpr unsigned TB_OFF = NORMS_OFF + (ROWS_A + ROWS_B) * 4u;
    static constexpr unsigned BAR_OFF = a16(TB_OFF + (ROWS_A + ROWS_B) * 4u);
    static constexpr unsigned TOTAL = BAR_OFF + 16u;
};

// extended frame e (0 .. n+5) of a token whose first row is `base`: where its 40 floats live
__device__ __forceinline__ const float *ext_frame(const float *base, int n, int dim, int e) {
    const int row = e < STACK_H ? 0 : (e < n + STACK_H ? e - STACK_H : n - 1);
    const int blk = e < STACK_H ? e : (e < n + STACK_H ? STACK_H : e - n + 1);
    return base + (size_t)row * dim + blk * STACK_F;
}

// bulk-copy engine (TMA, 1-D): one instruction moves a whole 160-byte frame and signals
// the CTA's mbarrier with the byte count
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes,
                                         unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes "
                 "[%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(unsigned bar, unsigned parity) {
    unsigned done = 0;
    for (unsigned spin = 0; spin < (1u << 28); ++spin) {      // bounded: a lost copy traps, never hangs
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}

// 7-tap diagonal sums + utils.py:47-58 -> D.  A thread owns RUNS of 8 cells along a
// diagonal, (i0 + k, j0 + k): the 14 Gram entries a run needs are loaded once, consecutive
// lanes take consecutive j0 (conflict-free).  Every sum adds its 7 taps in stack order,
// like the generic kernel's chunk loop.  Runs start at rows 8 b and columns -7 .. n2 - 1;
// cells outside the matrix are computed on whatever the loads returned and not stored.
template <bool TO_GMEM, typename L>
__device__ __forceinline__ int stack_epilogue(const float *Gs, const float *norms, const int *tb,
                                              unsigned dsk_addr, float *dist_gmem, int n1, int n2,
                                              int tid, int ld_gmem) {
    constexpr int RUN = 8;
    int bad = 0;
    const int W = n2 + RUN - 1;
    const int S = ((n1 + RUN - 1) / RUN) * W;
    const float invW = 1.0f / (float)W;
    for (int sidx = tid; sidx < S; sidx += AL_THREADS) {
        const int b = (int)(((float)sidx + 0.5f) * invW);
        const int i0 = b * RUN, j0 = sidx - b * W - (RUN - 1);
        const int klo = max(0, -j0), khi = min(n1 - i0, n2 - j0);     // valid cells: klo <= k < khi
        const float *gp = Gs + i0 * L::LDG + j0;
        float g[RUN + STACK_S - 1];
#pragma unroll
        for (int k = 0; k < RUN + STACK_S - 1; ++k) g[k] = gp[k * (L::LDG + 1)];
#pragma unroll
        for (int k = 0; k < RUN; ++k) {
            const int i = i0 + k, j = j0 + k;
            float tot = g[k];
#pragma unroll
            for (int d = 1; d < STACK_S; ++d) tot += g[k + d];
            const float dd = cell_distance(tot, norms[i], norms[L::ROWS_A + j]);
            if (k >= klo && k < khi) {
                if (!(dd >= 0.f)) bad = 1;
                if (TO_GMEM) dist_gmem[(size_t)i * ld_gmem + j] = dd;
                else asm volatile("st.shared.f32 [%0], %1;" ::"r"(dsk_addr + 4u * (unsigned)(tb[i + j] + i)),
                                  "f"(dd) : "memory");
            }
        }
    }
    return bad;
}

template <int RA, int NCG>
__global__ void __launch_bounds__(AL_THREADS)
align_stack_kernel(const AlignArgs a) {
    using L = StackLayout<RA, NCG>;
    constexpr int CLS = (RA - 1) * NCLS_SIDE + (NCG - 1);
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int ti = (warp >> 1) * 8 + (lane >> 2);
    const int tj = (warp & 1) * 4 + (lane & 3);
    const int beg = max(a.class_off[CLS], a.w0), end = min(a.class_off[CLS + 1], a.w1);
    const unsigned sbase = smem_u32(smem);
    const unsigned bar = sbase + L::BAR_OFF;
    float *Ds = reinterpret_cast<float *>(smem);
    float *Gs = reinterpret_cast<float *>(smem + L::G_OFF);
    float *n40 = reinterpret_cast<float *>(smem + L::N40_OFF);
    float *norms = reinterpret_cast<float *>(smem + L::NORMS_OFF);
    int *tb = reinterpret_cast<int *>(smem + L::TB_OFF);
    unsigned phase = 0;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    for (int it = beg + blockIdx.x; it < end; it += gridDim.x) {
        const int p = a.order[it];
        const int4 tk = reinterpret_cast<const int4 *>(a.pair_tok)[p];
        const int s1 = tk.x, n1 = tk.y, s2 = tk.z, n2 = tk.w;
        const int n1e = n1 + 2 * STACK_H, n2e = n2 + 2 * STACK_H;
        const float *g1 = a.feat + (size_t)s1 * a.dim, *g2 = a.feat + (size_t)s2 * a.dim;

        // 1. stage the extended frames of both tokens: one 160-byte bulk copy per frame
        if (tid == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar),
                         "r"((unsigned)(n1e + n2e) * (STACK_F * 4u)) : "memory");
        for (int e = tid; e < n1e + n2e; e += AL_THREADS) {
            const bool first = e < n1e;
            const int ee = first ? e : e - n1e;
            const float *src = first ? ext_frame(g1, n1, a.dim, ee) : ext_frame(g2, n2, a.dim, ee);
            bulk_g2s(sbase + ((first ? 0 : L::ROWS_A) + ee) * (KCP * 4), src, STACK_F * 4u, bar);
        }
        if (!a.dist_out) skew_table(tb, n1, n2, tid);
        mbar_wait_parity(bar, phase);
        phase ^= 1u;

        // 2. per-frame sums of squares and the 40-deep Gram tile G40 (register tiled)
        {
            const int nq = n1e + n2e;
            for (int q = tid; q < nq; q += AL_THREADS) {
                const int row = q < n1e ? q : L::ROWS_A + (q - n1e);
                n40[row] = row_sumsq(sbase + row * (KCP * 4), STACK_F / 4);
            }
        }
        float acc[RA][2 * NCG];
#pragma unroll
        for (int r = 0; r < RA; ++r)
#pragma unroll
            for (int c = 0; c < 2 * NCG; ++c) acc[r][c] = 0.f;
        {
            const unsigned a_addr = sbase + ti * (KCP * 4);
            const unsigned b_addr = sbase + (L::ROWS_A + tj) * (KCP * 4);
#pragma unroll 2
            for (int k4 = 0; k4 < STACK_F / 4; ++k4) {
                float4 av[RA];
#pragma unroll
                for (int r = 0; r < RA; ++r) av[r] = lds128(a_addr + (16 * r) * (KCP * 4) + k4 * 16);
#pragma unroll
                for (int cg = 0; cg < NCG; ++cg) {
                    float4 bv[2];
#pragma unroll
                    for (int c = 0; c < 2; ++c)
                        bv[c] = lds128(b_addr + (8 * (2 * cg + c)) * (KCP * 4) + k4 * 16);
#pragma unroll
                    for (int r = 0; r < RA; ++r)
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            float t = acc[r][2 * cg + c];
                            t = fmaf(av[r].x, bv[c].x, t);
                            t = fmaf(av[r].y, bv[c].y, t);
                            t = fmaf(av[r].z, bv[c].z, t);
                            t = fmaf(av[r].w, bv[c].w, t);
                            acc[r][2 * cg + c] = t;
                        }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < RA; ++r)
#pragma unroll
            for (int c = 0; c < 2 * NCG; ++c) Gs[(ti + 16 * r) * L::LDG + tj + 8 * c] = acc[r][c];
        __syncthreads();           // G40, n40 complete; the staging area is dead from here on

        // 3. reciprocal row norms: |row_i|^2 = sum_c |x[i+c-3]|^2, summed in stack order
        for (int q = tid; q < n1 + n2; q += AL_THREADS) {
            const int r0 = q < n1 ? q : L::ROWS_A + (q - n1);
            float ss = 0.f;
#pragma unroll
            for (int c = 0; c < STACK_S; ++c) ss += n40[r0 + c];
            norms[r0] = recip_norm(ss);
        }
        __syncthreads();

        // 4. distances -> D in skew layout over the dead staging area (or row-major to gmem)
        int bad;
        if (!a.dist_out) skew_guards(Ds, tb, n1, tid);
        if (a.dist_out)
            bad = stack_epilogue<true, L>(Gs, norms, tb, sbase, a.dist_out + a.dist_off[p], n1, n2,
                                          tid, n2);
        else
            bad = stack_epilogue<false, L>(Gs, norms, tb, sbase, nullptr, n1, n2, tid, n2);
        bad = __syncthreads_or(bad);
        if (tid == 0) {
            a.valid[p] = bad ? 0 : 1;
            if (bad && !a.dist_out) { a.path_len[p] = 0; a.cost[p] = nan(""); }
        }
        if (!a.dist_out && !bad)
            skew_copy_out(Ds, a.dws + (size_t)(it - a.w0) * a.slot_cells, skew_cells(n1, n2), tid);
        // the next pair's bulk copies (async proxy) overwrite what this pair read and wrote
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
    }
}

// rows r, r+1 of one file must overlap by dim - dim/S columns; counts the rows that do not
__global__ void stack_violations_kernel(const float *__restrict__ feat, int64_t n_rows, int dim,
                                        int stack, const uint8_t *__restrict__ last_row_of_file,
                                        unsigned long long *__restrict__ count) {
    const int warps_per_block = blockDim.x >> 5;
    const int64_t r = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    if (r + 1 >= n_rows || (last_row_of_file && last_row_of_file[r])) return;
    const int lane = threadIdx.x & 31, f = dim / stack;
    const float *a = feat + (size_t)r * dim + f, *b = feat + (size_t)(r + 1) * dim;
    int bad = 0;
    for (int k = lane; k < dim - f; k += 32)
        bad |= a[k] != b[k];      // value compare: +0 == -0; a NaN is a violation
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0 && bad) atomicAdd(count, 1ull);
}

// Rebuild the S-frame stack on the device from its middle blocks (the only part a host ->
// device upload of a verified stack needs to carry: S x fewer bytes over PCIe): block c of
// row t is the middle block of row t + c - S/2 when that row belongs to the same file, zeros
// otherwise (abnet3/features.py:135-159 pads with zeros at the file edges).
__global__ void restack_kernel(float *__restrict__ feat, int64_t n_rows, int dim, int stack,
                               const uint8_t *__restrict__ last_row_of_file) {
    const int warps_per_block = blockDim.x >> 5;
    const int64_t t = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    if (t >= n_rows) return;
    const int lane = threadIdx.x & 31, f = dim / stack, h = stack / 2;
    // file extent around t, looking at most h rows each way
    int back = 0, fwd = 0;
    while (back < h && t - back - 1 >= 0 && !(last_row_of_file && last_row_of_file[t - back - 1])) ++back;
    while (fwd < h && t + fwd + 1 < n_rows && !(last_row_of_file && last_row_of_file[t + fwd])) ++fwd;
    float *row = feat + (size_t)t * dim;
    for (int c = 0; c < stack; ++c) {
        if (c == h) continue;
        const int d = c - h;
        const bool inside = d < 0 ? -d <= back : d <= fwd;
        const float *src = feat + (size_t)(t + d) * dim + h * f;
        for (int k = lane; k < f; k += 32) row[c * f + k] = inside ? src[k] : 0.f;
    }
}

// ------------------------------------------------------- size-class bucketing
__device__ __forceinline__ int pair_class(const int4 tk, int64_t n_rows, int stack) {
    const int s1 = tk.x, n1 = tk.y, s2 = tk.z, n2 = tk.w;
    const bool ok = n1 > 0 && n2 > 0 && n1 <= NM_LIMIT && n2 <= NM_LIMIT && s1 >= 0 && s2 >= 0 &&
                    (int64_t)s1 + n1 <= n_rows && (int64_t)s2 + n2 <= n_rows;
    if (!ok) return CLS_INVALID;
    const int ext = stack ? 2 * STACK_H : 0;      // stack mode classes are on n + 6
    if (n1 + ext > NM_SHORT || n2 + ext > NM_SHORT) return CLS_LONG;
    return ((n1 + ext + 15) / 16 - 1) * NCLS_SIDE + ((n2 + ext + 15) / 16 - 1);
}

constexpr int BK_THREADS = 256, BK_ITEMS = 4;

__global__ void __launch_bounds__(BK_THREADS)
class_count_kernel(const AlignArgs a, int *__restrict__ counts) {
    __shared__ int hist[NBINS];
    if (threadIdx.x < NBINS) hist[threadIdx.x] = 0;
    __syncthreads();
    const int base = blockIdx.x * BK_THREADS * BK_ITEMS;
#pragma unroll
    for (int u = 0; u < BK_ITEMS; ++u) {
        const int p = base + u * BK_THREADS + threadIdx.x;
        if (p < a.n_pairs) {
            const int c = pair_class(reinterpret_cast<const int4 *>(a.pair_tok)[p], a.n_rows,
                                     a.stack);
            atomicAdd(&hist[c], 1);
            if (c == CLS_INVALID) {   // dataloader.py:184 / :188-191: the pair is skipped
                a.valid[p] = 0;
                if (!a.dist_out) { a.path_len[p] = 0; a.cost[p] = nan(""); }
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < NBINS && hist[threadIdx.x]) atomicAdd(&counts[threadIdx.x], hist[threadIdx.x]);
}

__global__ void class_scan_kernel(const int *__restrict__ counts, int *__restrict__ class_off) {
    if (threadIdx.x == 0) {
        int s = 0;
        for (int c = 0; c < NBINS; ++c) { class_off[c] = s; s += counts[c]; }
        class_off[NBINS] = s;
    }
}

__global__ void __launch_bounds__(BK_THREADS)
class_scatter_kernel(const AlignArgs a, const int *__restrict__ class_off,
                     int *__restrict__ cursor, int32_t *__restrict__ order) {
    __shared__ int hist[NBINS];
    __shared__ int gbase[NBINS];
    if (threadIdx.x < NBINS) hist[threadIdx.x] = 0;
    __syncthreads();
    const int base = blockIdx.x * BK_THREADS * BK_ITEMS;
    int cls[BK_ITEMS], rank[BK_ITEMS];
#pragma unroll
    for (int u = 0; u < BK_ITEMS; ++u) {
        const int p = base + u * BK_THREADS + threadIdx.x;
        cls[u] = -1;
        rank[u] = 0;
        if (p < a.n_pairs) {
            cls[u] = pair_class(reinterpret_cast<const int4 *>(a.pair_tok)[p], a.n_rows, a.stack);
            rank[u] = atomicAdd(&hist[cls[u]], 1);
            if (cls[u] == CLS_LONG && a.dist_out) a.valid[p] = 1;     // tiles clear it on a NaN
        }
    }
    __syncthreads();
    if (threadIdx.x < NBINS && hist[threadIdx.x])
        gbase[threadIdx.x] =
            class_off[threadIdx.x] + atomicAdd(&cursor[threadIdx.x], hist[threadIdx.x]);
    __syncthreads();
#pragma unroll
    for (int u = 0; u < BK_ITEMS; ++u) {
        const int p = base + u * BK_THREADS + threadIdx.x;
        if (cls[u] >= 0) order[gbase[cls[u]] + rank[u]] = p;
    }
}


// ------------------------------------------- long tokens, two-kernel path
// Long pairs follow the same two-stage scheme as the fused classes, with the matrix cut in
// tiles for the distance stage and in 128-row bands for the DTW stage:
//   long_tile_kernel<STACKED>  one CTA per (pair, tile): the tile's distances exactly like a
//       (6,6) class pair -- stacked tables through the 40-deep Gram + 7-tap diagonal sums
//       (90 x 90 tiles: 96 extended frames per side), generic tables through the 280-deep
//       register tile (96 x 96) -- written to the pair's skew-layout slot in the hand-over
//       workspace (anti-diagonal segments of a tile are contiguous there).  Tiles are
//       independent: the whole device works on the distance stage of a window of pairs.
//   dtw_band_kernel            one warp per pair.  Lane l owns rows 128 b + 4 l .. + 3 of band
//       b and sweeps the band's anti-diagonals over ALL columns (n2 + 127 steps, coalesced
//       skew-layout loads, same float64 recurrence and tie order as dtw_skew_kernel); the
//       band's last row is the next band's boundary (shared memory).  Directions go to a
//       per-warp scratch area in global memory (one coalesced 128-byte store per 4 steps),
//       lane 0 walks them back.
constexpr int LT_GEN = NM_SHORT;          // 96
constexpr int LT_STK = STACK_MAXN;        // 90
constexpr int BAND_G = 4, BAND_ROWS = 32 * BAND_G;
constexpr int LONG_SLACK = 2 * BAND_ROWS + 16;     // floats a band sweep may read past the matrix
constexpr int LONG_DTW_WARPS = 4;
constexpr int LONG_DTW_MAX_CTAS = 1024;

struct LongArgs {
    float *dws;             // hand-over slots of the window's long pairs
    unsigned *bad;          // one flag per slot: a NaN / negative distance was seen
    uint8_t *dirs;          // per-warp direction scratch of dtw_band_kernel
    size_t slot_cells;      // floats per slot
    size_t dirs_bytes;      // bytes per warp
    int lw0, lw1;           // window inside the CLS_LONG class (positions relative to its start)
    int tcap;               // tiles per side of the largest matrix
};

__host__ __device__ inline size_t long_dirs_bytes(int nmax) {
    const size_t bands = (size_t)(nmax + BAND_ROWS - 1) / BAND_ROWS;
    return bands * (size_t)((nmax + BAND_ROWS + 3) / 4 + 1) * BAND_ROWS;
}

template <bool STACKED, int R>
__global__ void __launch_bounds__(AL_THREADS)
long_tile_kernel(const AlignArgs a, const LongArgs la) {
    using LS = StackLayout<R, R>;
    using LG = DistLayout<R, R>;
    constexpr int LTS = STACKED ? 16 * R - 2 * STACK_H : 16 * R;
    constexpr unsigned TBG_OFF = STACKED ? LS::TOTAL : LG::TOTAL;     // int[32 R]: global bases per tile diagonal
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x;
    const unsigned sbase = smem_u32(smem);
    float *Ds = reinterpret_cast<float *>(smem);
    int *tb = reinterpret_cast<int *>(smem + (STACKED ? LS::TB_OFF : LG::TB_OFF));
    int *tbg = reinterpret_cast<int *>(smem + TBG_OFF);
    const int cbeg = a.class_off[CLS_LONG], cend = a.class_off[CLS_LONG + 1];
    const int lo = cbeg + la.lw0, hi = min(cend, cbeg + la.lw1);
    const int t2 = la.tcap * la.tcap;
    unsigned phase = 0;
    const unsigned bar = sbase + LS::BAR_OFF;
    if (STACKED) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    const long long n_work = (long long)max(hi - lo, 0) * t2;
    for (long long w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int k = (int)(w / t2), tile = (int)(w - (long long)k * t2);
        const int bi = tile / la.tcap, bj = tile - bi * la.tcap;
        const int p = a.order[lo + k];
        const int4 tk = reinterpret_cast<const int4 *>(a.pair_tok)[p];
        const int s1 = tk.x, n1 = tk.y, s2 = tk.z, n2 = tk.w;
        const int i0 = bi * LTS, j0 = bj * LTS;
        if (i0 >= n1 || j0 >= n2) continue;                 // CTA-uniform
        const int h = min(LTS, n1 - i0), wd = min(LTS, n2 - j0);
        const int slot = k;
        float *dst = la.dws ? la.dws + (size_t)slot * la.slot_cells : nullptr;
        const float *g1 = a.feat + (size_t)s1 * a.dim, *g2 = a.feat + (size_t)s2 * a.dim;
        int bad = 0;
        if (!a.dist_out) {
            skew_table(tb, h, wd, tid);                                       // tile-local layout
            for (int t = tid; t < h + wd - 1; t += AL_THREADS)                // where diagonal t of the tile starts
                tbg[t] = skew_base(i0 + j0 + t, n1, n2) + i0;                 // in the pair's slot (row i0)
        }
        if (STACKED) {
            const int he = h + 2 * STACK_H, we = wd + 2 * STACK_H;
            float *Gs = reinterpret_cast<float *>(smem + LS::G_OFF);
            float *n40 = reinterpret_cast<float *>(smem + LS::N40_OFF);
            float *norms = reinterpret_cast<float *>(smem + LS::NORMS_OFF);
            const int warp = tid >> 5, lane = tid & 31;
            const int ti = (warp >> 1) * 8 + (lane >> 2);
            const int tj = (warp & 1) * 4 + (lane & 3);
            if (tid == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar),
                             "r"((unsigned)(he + we) * (STACK_F * 4u)) : "memory");
            for (int e = tid; e < he + we; e += AL_THREADS) {
                const bool first = e < he;
                const int ee = first ? e : e - he;
                const float *src = first ? ext_frame(g1, n1, a.dim, i0 + ee) : ext_frame(g2, n2, a.dim, j0 + ee);
                bulk_g2s(sbase + ((first ? 0 : LS::ROWS_A) + ee) * (KCP * 4), src, STACK_F * 4u, bar);
            }
            mbar_wait_parity(bar, phase);
            phase ^= 1u;
            for (int q = tid; q < he + we; q += AL_THREADS) {
                const int row = q < he ? q : LS::ROWS_A + (q - he);
                n40[row] = row_sumsq(sbase + row * (KCP * 4), STACK_F / 4);
            }
            float acc[R][2 * R];
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int c = 0; c < 2 * R; ++c) acc[r][c] = 0.f;
            {
                const unsigned a_addr = sbase + ti * (KCP * 4);
                const unsigned b_addr = sbase + (LS::ROWS_A + tj) * (KCP * 4);
#pragma unroll 2
                for (int k4 = 0; k4 < STACK_F / 4; ++k4) {
                    float4 av[R];
#pragma unroll
                    for (int r = 0; r < R; ++r) av[r] = lds128(a_addr + (16 * r) * (KCP * 4) + k4 * 16);
#pragma unroll
                    for (int cg = 0; cg < R; ++cg) {
                        float4 bv[2];
#pragma unroll
                        for (int c = 0; c < 2; ++c)
                            bv[c] = lds128(b_addr + (8 * (2 * cg + c)) * (KCP * 4) + k4 * 16);
#pragma unroll
                        for (int r = 0; r < R; ++r)
#pragma unroll
                            for (int c = 0; c < 2; ++c) {
                                float t = acc[r][2 * cg + c];
                                t = fmaf(av[r].x, bv[c].x, t);
                                t = fmaf(av[r].y, bv[c].y, t);
                                t = fmaf(av[r].z, bv[c].z, t);
                                t = fmaf(av[r].w, bv[c].w, t);
                                acc[r][2 * cg + c] = t;
                            }
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int c = 0; c < 2 * R; ++c) Gs[(ti + 16 * r) * LS::LDG + tj + 8 * c] = acc[r][c];
            __syncthreads();
            for (int q = tid; q < h + wd; q += AL_THREADS) {
                const int r0 = q < h ? q : LS::ROWS_A + (q - h);
                float ss = 0.f;
#pragma unroll
                for (int c = 0; c < STACK_S; ++c) ss += n40[r0 + c];
                norms[r0] = recip_norm(ss);
            }
            __syncthreads();
            if (a.dist_out)
                bad = stack_epilogue<true, LS>(Gs, norms, tb, sbase,
                                               a.dist_out + a.dist_off[p] + (size_t)i0 * n2 + j0, h, wd, tid, n2);
            else
                bad = stack_epilogue<false, LS>(Gs, norms, tb, sbase, nullptr, h, wd, tid, wd);
        } else {
            pair_distance<R, R, LG>(
                smem, g1 + (size_t)i0 * a.dim, g2 + (size_t)j0 * a.dim, h, wd, a.dim,
                a.dist_out ? a.dist_out + a.dist_off[p] + (size_t)i0 * n2 + j0 : nullptr, n2,
                a.dist_out ? nullptr : tb, bad);
        }
        bad = __syncthreads_or(bad);
        if (bad && tid == 0) {
            if (a.dist_out) a.valid[p] = 0;
            else la.bad[slot] = 1u;
        }
        if (!a.dist_out) {
            // tile-local skew -> the pair's slot: diagonal t of the tile is one contiguous run of
            // rows rlo .. rhi in both layouts; one warp per diagonal
            const int warp = tid >> 5, lane = tid & 31;
            for (int t = warp; t < h + wd - 1; t += AL_THREADS / 32) {
                const int rlo = max(0, t - wd + 1), rhi = min(h - 1, t);
                const float *src = Ds + tb[t];
                float *d = dst + tbg[t];
                for (int r = rlo + lane; r <= rhi; r += 32) d[r] = src[r];
            }
            // the +inf guard after diagonal i - 1 of the pair (the cell left of row i)
            if (bj == 0)
                for (int r = tid; r < h; r += AL_THREADS) {
                    const int i = i0 + r;
                    if (i >= 1) dst[skew_base(i - 1, n1, n2) + i] = __int_as_float(0x7f800000);
                }
        }
        if (STACKED) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
    }
}

__global__ void __launch_bounds__(LONG_DTW_WARPS * 32)
dtw_band_kernel(const AlignArgs a, const LongArgs la, int nmax) {
    constexpr int G = BAND_G;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned per_warp = 2u * 8u * (unsigned)(nmax + 2) + 4u * 2u * (unsigned)nmax;
    unsigned char *mine = smem + warp * per_warp;
    double *topbuf0 = reinterpret_cast<double *>(mine);
    uint16_t *pb_i = reinterpret_cast<uint16_t *>(mine + 2u * 8u * (unsigned)(nmax + 2));
    uint16_t *pb_j = pb_i + 2 * nmax;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    const int cbeg = a.class_off[CLS_LONG], cend = a.class_off[CLS_LONG + 1];
    const int lo = cbeg + la.lw0, hi = min(cend, cbeg + la.lw1);
    const int gw = blockIdx.x * LONG_DTW_WARPS + warp;
    uint8_t *dirs = la.dirs + (size_t)gw * la.dirs_bytes;
    const unsigned band_bytes = (unsigned)((nmax + BAND_ROWS + 3) / 4 + 1) * BAND_ROWS;
    const int r0 = lane * G;

    for (int k = gw; lo + k < hi; k += gridDim.x * LONG_DTW_WARPS) {
        const int p = a.order[lo + k];
        const int4 tk = reinterpret_cast<const int4 *>(a.pair_tok)[p];
        const int s1 = tk.x, n1 = tk.y, s2 = tk.z, n2 = tk.w;
        if (la.bad[k] || n1 > nmax || n2 > nmax) {    // NaN in the matrix (utils.py:59), or max_frames was too small
            if (lane == 0) { a.valid[p] = 0; a.path_len[p] = 0; a.cost[p] = nan(""); }
            continue;
        }
        const float *__restrict__ D = la.dws + (size_t)k * la.slot_cells;
        const int T = n1 + n2 - 1;
        const int n_bands = (n1 + BAND_ROWS - 1) / BAND_ROWS;
        for (int j = lane; j <= n2; j += 32) topbuf0[j] = j == 0 ? 0.0 : INF;     // C[-1][-1] = 0 feeds (0, 0)
        __syncwarp();
        double cost = 0.0;
        for (int b = 0; b < n_bands; ++b) {
            const int ib = b * BAND_ROWS;
            const int rows = min(BAND_ROWS, n1 - ib);
            const int S = n2 + rows - 1;                       // band steps: cell (ib + r, s - r)
            const double *top = topbuf0 + (b & 1) * (nmax + 2);
            double *top_next = topbuf0 + ((b + 1) & 1) * (nmax + 2);
            if (lane == 0) top_next[0] = INF;                  // C[last row][-1]
            uint32_t *dirs32 = reinterpret_cast<uint32_t *>(dirs + (size_t)b * band_bytes);
            const int r_last = rows - 1;                       // the row whose values are the next band's top
            double cur[G], prev[G];
#pragma unroll
            for (int g = 0; g < G; ++g) cur[g] = prev[g] = INF;
            double nbprev = lane == 0 ? top[0] : INF;
            const float *Dl = D + ib + r0;
            asm volatile("" : "+l"(Dl));
            int bases = 0;
#define ABN_BAND_LOAD(dst, s_first)                                                         \
            {                                                                               \
                if (((s_first) & 31) == 0)                                                  \
                    bases = skew_base(min(ib + (s_first) + lane, T - 1), n1, n2);           \
                _Pragma("unroll") for (int u = 0; u < 4; ++u) {                             \
                    const int bs = __shfl_sync(FULL, bases, ((s_first) + u) & 31);          \
                    _Pragma("unroll") for (int g = 0; g < G; ++g) dst[u][g] = __ldg(Dl + bs + g); \
                }                                                                           \
            }
#define ABN_BAND_STEP(dc, u, s)                                                             \
            {                                                                               \
                double up0 = __shfl_up_sync(FULL, cur[G - 1], 1);                           \
                if (lane == 0) up0 = top[1 + min((s), n2 - 1)];                             \
                const double dg0 = nbprev;                                                  \
                nbprev = up0;                                                               \
                double nw[G];                                                               \
                _Pragma("unroll") for (int g = 0; g < G; ++g) {                             \
                    const double d = (double)dc[u][g];                                      \
                    const double up = g == 0 ? up0 : cur[g > 0 ? g - 1 : 0];                \
                    const double dg = g == 0 ? dg0 : prev[g > 0 ? g - 1 : 0];               \
                    const double lf = cur[g];                                               \
                    const bool up_le = up <= lf;                                            \
                    const double m1 = up_le ? up : lf;                                      \
                    const bool use_dg = dg <= m1;                                           \
                    const double mm = use_dg ? dg : m1;                                     \
                    const unsigned dir = use_dg ? DIR_DIAG : (up_le ? DIR_UP : DIR_LEFT);   \
                    nw[g] = d + mm;                                                         \
                    bits |= dir << (8 * g + 2 * (u));                                       \
                    const int jj = (s) - (r0 + g);                                          \
                    if (r0 + g == r_last && (unsigned)jj < (unsigned)n2) top_next[1 + jj] = nw[g]; \
                }                                                                           \
                _Pragma("unroll") for (int g = 0; g < G; ++g) { prev[g] = cur[g]; cur[g] = nw[g]; } \
            }
#define ABN_BAND_BLOCK(dc, dnext, s0)                                                       \
            {                                                                               \
                ABN_BAND_LOAD(dnext, (s0) + 4)                                              \
                unsigned bits = 0;                                                          \
                if ((s0) + 4 <= S) {                                                        \
                    ABN_BAND_STEP(dc, 0, (s0)) ABN_BAND_STEP(dc, 1, (s0) + 1)               \
                    ABN_BAND_STEP(dc, 2, (s0) + 2) ABN_BAND_STEP(dc, 3, (s0) + 3)           \
                } else {                                                                    \
                    ABN_BAND_STEP(dc, 0, (s0))                                              \
                    if ((s0) + 1 < S) ABN_BAND_STEP(dc, 1, (s0) + 1)                        \
                    if ((s0) + 2 < S) ABN_BAND_STEP(dc, 2, (s0) + 2)                        \
                }                                                                           \
                dirs32[((s0) >> 2) * 32 + lane] = bits;                                     \
            }
            float da[4][G], db[4][G];
            ABN_BAND_LOAD(da, 0)
            for (int s0 = 0; s0 < S; s0 += 8) {
                ABN_BAND_BLOCK(da, db, s0)
                if (s0 + 4 < S) ABN_BAND_BLOCK(db, da, s0 + 4)
            }
#undef ABN_BAND_BLOCK
#undef ABN_BAND_STEP
#undef ABN_BAND_LOAD
            if (b == n_bands - 1) {
                const int gl = r_last % G;
                double c = cur[0];
                c = gl == 1 ? cur[1] : c;
                c = gl == 2 ? cur[2] : c;
                c = gl == 3 ? cur[3] : c;
                cost = __shfl_sync(FULL, c, r_last / G);
            }
            __syncwarp();
        }
        __threadfence_block();
        __syncwarp();
        int len = 0;
        if (lane == 0) {
            int i = n1 - 1, j = n2 - 1;
            len = 1;
            pb_i[0] = (uint16_t)i; pb_j[0] = (uint16_t)j;
            while ((i | j) != 0 && len < T) {
                const int b = i >> 7, r = i & (BAND_ROWS - 1), s = j + r;
                const unsigned byte = *reinterpret_cast<volatile uint8_t *>(
                    dirs + (size_t)b * band_bytes + (size_t)(s >> 2) * BAND_ROWS + r);
                const unsigned d = (byte >> (2 * (s & 3))) & 3u;
                i -= (d != DIR_LEFT);
                j -= (d != DIR_UP);
                pb_i[len] = (uint16_t)i; pb_j[len] = (uint16_t)j; ++len;
            }
            a.path_len[p] = len;
            a.cost[p] = cost;
            a.valid[p] = 1;
        }
        len = __shfl_sync(FULL, len, 0);
        __syncwarp();
        const int64_t off = a.path_off[p];
        for (int q = lane; q < len; q += 32) {
            a.idx1[off + q] = s1 + (int)pb_i[len - 1 - q];
            a.idx2[off + q] = s2 + (int)pb_j[len - 1 - q];
        }
        __syncwarp();
    }
}

// Test hook: DTW on caller-supplied float64 matrices (bit-exact mode).
struct DtwLayout { int nm, ldd; unsigned dirs_off, path_off, misc_off, total; };
__host__ __device__ inline DtwLayout dtw_layout(int nm) {
    DtwLayout L;
    L.nm = nm;
    L.ldd = nm + 1;
    L.dirs_off = ((unsigned)nm * L.ldd * 8u + 15u) & ~15u;
    L.path_off = L.dirs_off + (unsigned)nm * nm;
    L.misc_off = (L.path_off + 4u * nm + 15u) & ~15u;
    L.total = L.misc_off + 32u;
    return L;
}

__global__ void __launch_bounds__(AL_THREADS)
dtw_from_dist_kernel(const double *__restrict__ dist, const int64_t *__restrict__ dist_off,
                     const int32_t *__restrict__ shape, int n_pairs,
                     const int64_t *__restrict__ path_off, int32_t *__restrict__ path1,
                     int32_t *__restrict__ path2, int32_t *__restrict__ path_len,
                     double *__restrict__ cost, uint8_t *__restrict__ valid, int nm) {
    extern __shared__ __align__(16) unsigned char smem[];
    const DtwLayout L = dtw_layout(nm);
    const int p = blockIdx.x;
    if (p >= n_pairs) return;
    const int n1 = shape[2 * p], n2 = shape[2 * p + 1];
    const int tid = threadIdx.x;
    if (n1 <= 0 || n2 <= 0 || n1 > nm || n2 > nm) {
        if (tid == 0) { path_len[p] = 0; cost[p] = nan(""); valid[p] = 0; }
        return;
    }
    double *Ds = reinterpret_cast<double *>(smem);
    uint8_t *dirs = smem + L.dirs_off;
    uint8_t *pb_i = smem + L.path_off;
    uint8_t *pb_j = pb_i + 2 * nm;
    int *misc = reinterpret_cast<int *>(smem + L.misc_off);
    const double *src = dist + dist_off[p];
    int bad = 0;
    for (int e = tid; e < n1 * n2; e += AL_THREADS) {
        const double d = src[e];
        if (!(d >= 0.0)) bad = 1;
        Ds[(e / n2) * L.ldd + (e % n2)] = d;
    }
    bad = __syncthreads_or(bad);
    if (bad) {
        if (tid == 0) { path_len[p] = 0; cost[p] = nan(""); valid[p] = 0; }
        return;
    }
    if (tid < 32) {
        const int g = (n1 + 31) >> 5;
        double c;
        if (g == 1)      c = dtw_wavefront<double, 1>(Ds, L.ldd, dirs, nm, n1, n2, tid);
        else if (g == 2) c = dtw_wavefront<double, 2>(Ds, L.ldd, dirs, nm, n1, n2, tid);
        else             c = dtw_wavefront<double, 3>(Ds, L.ldd, dirs, nm, n1, n2, tid);
        __syncwarp();
        if (tid == 0) {
            const int len = traceback(dirs, nm, n1, n2, pb_i, pb_j);
            misc[0] = len;
            path_len[p] = len;
            cost[p] = c;
            valid[p] = 1;
        }
    }
    __syncthreads();
    const int len = misc[0];
    const int64_t off = path_off[p];
    for (int k = tid; k < len; k += AL_THREADS) {
        path1[off + k] = (int)pb_i[len - 1 - k];
        path2[off + k] = (int)pb_j[len - 1 - k];
    }
}

// ------------------------------------------------ diff pairs / compaction --
// dataloader.py:208-231: truncation, or diagonal stretch
// rint(linspace(0, len_min-1, len_max)) with the LONGER token in X1 (first
// operand wins length ties in Python's min/max, so n1 == n2 reads token 1 on
// both sides).
__global__ void diff_pairs_kernel(const int32_t *__restrict__ pair_tok, int n_pairs, int stretch,
                                  const int64_t *__restrict__ out_off, int32_t *__restrict__ idx1,
                                  int32_t *__restrict__ idx2) {
    const int warps_per_block = blockDim.x >> 5;
    const int p = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    if (p >= n_pairs) return;
    const int lane = threadIdx.x & 31;
    const int4 tk = reinterpret_cast<const int4 *>(pair_tok)[p];
    const int s1 = tk.x, n1 = tk.y, s2 = tk.z, n2 = tk.w;
    if (n1 <= 0 || n2 <= 0) return;
    const int64_t off = out_off[p];
    if (!stretch) {
        const int m = min(n1, n2);
        for (int k = lane; k < m; k += 32) { idx1[off + k] = s1 + k; idx2[off + k] = s2 + k; }
    } else {
        const int smax = n1 >= n2 ? s1 : s2, smin = n1 <= n2 ? s1 : s2;
        const int lmax = max(n1, n2), lmin = min(n1, n2);
        // numpy.linspace(0, lmin-1, lmax): step = (lmin-1)/(lmax-1) in float64,
        // y[k] = k*step, last element forced to stop; rint = half-to-even
        const double step = lmax > 1 ? (double)(lmin - 1) / (double)(lmax - 1) : 0.0;
        for (int k = lane; k < lmax; k += 32) {
            double v = (double)k * step;
            if (k == lmax - 1 && lmax > 1) v = (double)(lmin - 1);
            idx1[off + k] = smax + k;
            idx2[off + k] = smin + (int)rint(v);
        }
    }
}

__global__ void compact_paths_kernel(const int32_t *__restrict__ src1,
                                     const int32_t *__restrict__ src2,
                                     const int64_t *__restrict__ src_off,
                                     const int64_t *__restrict__ dst_off,
                                     const int32_t *__restrict__ path_len, int n_pairs,
                                     int32_t *__restrict__ dst1, int32_t *__restrict__ dst2) {
    const int warps_per_block = blockDim.x >> 5;
    const int p = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    if (p >= n_pairs) return;
    const int lane = threadIdx.x & 31;
    const int len = path_len[p];
    const int64_t so = src_off[p], d0 = dst_off[p];
    for (int k = lane; k < len; k += 32) {
        dst1[d0 + k] = src1[so + k];
        dst2[d0 + k] = src2[so + k];
    }
}

// Batch generation: one warp copies one 280-float row with 128-bit accesses.
__global__ void gather_batch_kernel(const float *__restrict__ feat, int dim,
                                    const int32_t *__restrict__ idx1,
                                    const int32_t *__restrict__ idx2,
                                    const int8_t *__restrict__ y_in,
                                    const int64_t *__restrict__ sel, int64_t n,
                                    float *__restrict__ x1, float *__restrict__ x2,
                                    float *__restrict__ y_out) {
    const int warps_per_block = blockDim.x >> 5;
    const int64_t w = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    if (w >= 2 * n) return;
    const int lane = threadIdx.x & 31;
    const int64_t k = w >> 1;
    const int side = (int)(w & 1);
    const int64_t pos = sel ? sel[k] : k;
    const int32_t row = side ? idx2[pos] : idx1[pos];
    const float4 *src = reinterpret_cast<const float4 *>(feat + (size_t)row * dim);
    float4 *dst = reinterpret_cast<float4 *>((side ? x2 : x1) + (size_t)k * dim);
    for (int c = lane; c < dim / 4; c += 32) dst[c] = __ldg(src + c);
    if (side == 0 && lane == 0 && y_out) y_out[k] = y_in ? (float)y_in[pos] : 1.f;
}

// ----------------------------------------------------------- host dispatch --
struct ClassLaunch {
    void (*kernel)(const AlignArgs);
    unsigned smem;
    int grid;      // resident CTAs on the whole device (persistent launch)
};

static int prepare_kernel(ClassLaunch &cl, void (*kernel)(const AlignArgs), unsigned smem,
                          int sm_count, const char *what, int ra, int ncg) {
    cl.kernel = kernel;
    cl.smem = smem;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess)
        return set_error(ABN_EIO, "cudaFuncSetAttribute(smem=%u) failed for %s class %d,%d", smem,
                         what, ra, ncg);
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, AL_THREADS, smem) !=
            cudaSuccess || per_sm < 1)
        return set_error(ABN_EIO, "occupancy query failed for %s class %d,%d", what, ra, ncg);
    cl.grid = per_sm * sm_count;
    return ABN_OK;
}

template <int RA, int NCG>
static int prepare_class(ClassLaunch *generic, ClassLaunch *stacked, int sm_count) {
    constexpr int idx = (RA - 1) * NCLS_SIDE + (NCG - 1);
    if (int rc = prepare_kernel(generic[idx], align_class_kernel<RA, NCG>,
                                DistLayout<RA, NCG>::TOTAL, sm_count, "generic", RA, NCG))
        return rc;
    return prepare_kernel(stacked[idx], align_stack_kernel<RA, NCG>, StackLayout<RA, NCG>::TOTAL,
                          sm_count, "stacked", RA, NCG);
}

template <int RA>
static int prepare_row(ClassLaunch *g, ClassLaunch *s, int sm_count) {
    int rc = ABN_OK;
    if (!rc) rc = prepare_class<RA, 1>(g, s, sm_count);
    if (!rc) rc = prepare_class<RA, 2>(g, s, sm_count);
    if (!rc) rc = prepare_class<RA, 3>(g, s, sm_count);
    if (!rc) rc = prepare_class<RA, 4>(g, s, sm_count);
    if (!rc) rc = prepare_class<RA, 5>(g, s, sm_count);
    if (!rc) rc = prepare_class<RA, 6>(g, s, sm_count);
    return rc;
}

static int class_table(const ClassLaunch **generic, const ClassLaunch **stacked) {
    static ClassLaunch gtab[NCLS], stab[NCLS];
    static int state = 0;        // 0 = not built, 1 = ok  (one device per process)
    if (state == 0) {
        int dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        int rc = ABN_OK;
        if (!rc) rc = prepare_row<1>(gtab, stab, sms);
        if (!rc) rc = prepare_row<2>(gtab, stab, sms);
        if (!rc) rc = prepare_row<3>(gtab, stab, sms);
        if (!rc) rc = prepare_row<4>(gtab, stab, sms);
        if (!rc) rc = prepare_row<5>(gtab, stab, sms);
        if (!rc) rc = prepare_row<6>(gtab, stab, sms);
        if (rc) return rc;
        state = 1;
    }
    *generic = gtab;
    *stacked = stab;
    return ABN_OK;
}

// hand-over slot of one pair (floats): the largest matrix of the fused classes, or -- when the
// pair list may hold long tokens (above STACK_MAXN frames: the stacked kernels route those to
// the long path) -- the largest matrix overall
static bool may_be_long(int max_frames) { return max_frames > STACK_MAXN; }
static size_t slot_cells_for(int max_frames) {
    if (may_be_long(max_frames)) {
        const size_t n = (size_t)(max_frames < NM_SHORT ? NM_SHORT : max_frames);
        return (n * n + n + LONG_SLACK + 3) & ~(size_t)3;
    }
    const size_t ns = (size_t)max_frames;
    return (ns * ns + ns + SKEW_SLACK + 3) & ~(size_t)3;
}
static int long_dtw_warps(int n_pairs) {
    const int want = ((n_pairs > 0 ? n_pairs : 1) + LONG_DTW_WARPS - 1) / LONG_DTW_WARPS * LONG_DTW_WARPS;
    const int cap = LONG_DTW_MAX_CTAS * LONG_DTW_WARPS;
    return want < cap ? want : cap;
}
static size_t a256(size_t x) { return (x + 255) & ~(size_t)255; }
// workspace: header | order | [long path: bad flags | direction scratch] | hand-over slots
static size_t ws_bad_off(int n_pairs) {
    return a256(WS_ORDER + sizeof(int32_t) * (size_t)(n_pairs > 0 ? n_pairs : 0));
}
static size_t ws_dirs_off(int n_pairs) {
    return a256(ws_bad_off(n_pairs) + sizeof(unsigned) * (size_t)(n_pairs > 0 ? n_pairs : 0));
}
static size_t ws_dist_off(int n_pairs, int max_frames) {
    if (!may_be_long(max_frames)) return ws_bad_off(n_pairs);
    return a256(ws_dirs_off(n_pairs) + (size_t)long_dtw_warps(n_pairs) * long_dirs_bytes(max_frames));
}

struct DtwLaunch { int grid; unsigned smem; int t4_cap; };

template <int G>
static int launch_dtw(const AlignArgs &a, int ra, int ext, cudaStream_t st) {
    static DtwLaunch tab[2][NCLS_SIDE + 1] = {};
    DtwLaunch &d = tab[ext ? 1 : 0][ra];
    if (d.grid == 0) {
        // longest diagonal count of the row: (16 ra - ext) + (NM_SHORT - ext) - 1 steps
        d.t4_cap = (16 * ra + NM_SHORT - 2 * ext + 2) / 4;
        d.smem = DTW_WARPS * (unsigned)(d.t4_cap * 32 * G + 8 * d.t4_cap);
        // at most 20 KB: below the 48 KB every kernel may use without opting in (and the
        // opt-in attribute is per kernel, not per launch, so rows sharing a G must not set it)
        int per_sm = 0, dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dtw_skew_kernel<G>,
                                                          DTW_WARPS * 32, d.smem) != cudaSuccess ||
            per_sm < 1)
            return set_error(ABN_EIO, "occupancy query failed for the DTW kernel (G = %d)", G);
        d.grid = per_sm * sms;
    }
    const int want = (a.w1 - a.w0 + DTW_WARPS - 1) / DTW_WARPS;
    dtw_skew_kernel<G><<<d.grid < want ? d.grid : want, DTW_WARPS * 32, d.smem, st>>>(
        a, (ra - 1) * NCLS_SIDE, ra * NCLS_SIDE, d.t4_cap);
    if (cudaError_t e = cudaGetLastError())
        return set_error(ABN_EIO, "DTW kernel launch (G %d, row %d, grid %d, smem %u): %s", G, ra,
                         d.grid < want ? d.grid : want, d.smem, cudaGetErrorString(e));
    return ABN_OK;
}

static int run_align(AlignArgs a, int max_frames, void *workspace, size_t workspace_bytes,
                     cudaStream_t st, const char *who) {
    if (max_frames <= 0) return set_error(ABN_EINVAL, "%s: max_frames must be positive", who);
    if (max_frames > NM_LIMIT)
        return set_error(ABN_ERANGE, "%s: token of %d frames exceeds %d", who, max_frames, NM_LIMIT);
    const bool dist_only = a.dist_out != nullptr;
    const size_t slot = slot_cells_for(max_frames);
    const size_t d_off = ws_dist_off(a.n_pairs, max_frames);
    const size_t need = dist_only ? WS_ORDER + sizeof(int32_t) * (size_t)a.n_pairs
                                  : d_off + slot * sizeof(float);
    if (!workspace || workspace_bytes < need)
        return set_error(ABN_ENOMEM, "%s: workspace of at least %zu bytes needed, %zu given", who,
                         need, workspace_bytes);
    if (a.stack != 0 && (a.stack != STACK_S || a.dim != STACK_S * STACK_F))
        return set_error(ABN_EINVAL, "%s: the stacked fast path needs stack == %d and dim == %d "
                         "(got stack %d, dim %d); pass stack = 0 for the generic kernels", who,
                         STACK_S, STACK_S * STACK_F, a.stack, a.dim);
    const ClassLaunch *gtab = nullptr, *stab = nullptr;
    if (int rc = class_table(&gtab, &stab)) return rc;
    const ClassLaunch *tab = a.stack ? stab : gtab;
    const int ext = a.stack ? 2 * STACK_H : 0;
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    int *counts = reinterpret_cast<int *>(ws + WS_COUNTS);
    int *class_off = reinterpret_cast<int *>(ws + WS_OFF);
    int *cursor = reinterpret_cast<int *>(ws + WS_CURSOR);
    int32_t *order = reinterpret_cast<int32_t *>(ws + WS_ORDER);
    a.order = order;
    a.class_off = class_off;
    a.w0 = 0;
    a.w1 = a.n_pairs;
    a.dws = nullptr;
    a.slot_cells = (int)slot;
    cudaMemsetAsync(ws, 0, WS_ORDER, st);
    const int per_block = BK_THREADS * BK_ITEMS;
    const int blocks = (a.n_pairs + per_block - 1) / per_block;
    class_count_kernel<<<blocks, BK_THREADS, 0, st>>>(a, counts);
    class_scan_kernel<<<1, 32, 0, st>>>(counts, class_off);
    class_scatter_kernel<<<blocks, BK_THREADS, 0, st>>>(a, class_off, cursor, order);
    size_t chunk = dist_only ? (size_t)a.n_pairs : (workspace_bytes - d_off) / (slot * sizeof(float));
    if (chunk > (size_t)a.n_pairs) chunk = (size_t)a.n_pairs;
    if (max_frames + ext > NM_SHORT) {
        // long pairs: tile distance kernel -> skew slots -> band DTW kernel, in windows of
        // `chunk` pairs of the long class (its size is only known on the device: windows past
        // its end find nothing to do)
        // tiles: the (4,4) class (58 x 58 on a stacked table, 64 x 64 otherwise) -- five resident
        // CTAs per SM instead of two for the (6,6) class, and 7 x 58 covers 400 frames with less
        // padding than 5 x 90: 595 k pairs/s against 466 k at 400 frames (tools/time_long.py;
        // ABN_LONG_R = 5 | 6 selects the larger classes)
        static int tile_grid[2][3] = {{0, 0, 0}, {0, 0, 0}}, band_grid = 0, sms = 0;
        const int which = a.stack ? 1 : 0;
        const int lt_max = a.stack ? LT_STK : LT_GEN;
        const int nt = (max_frames + lt_max - 1) / lt_max;
        int R = 4;
        (void)nt;
        {   // experiment knob: ABN_LONG_R = 4 | 5 | 6 forces the tile class
            static int forced = -1;
            if (forced < 0) { const char *e = getenv("ABN_LONG_R"); forced = e ? atoi(e) : 0; }
            if (forced >= 4 && forced <= 6) R = forced;
        }
        const int lt = 16 * R - ext;
        typedef void (*TileKernel)(const AlignArgs, const LongArgs);
        static const TileKernel kernels[2][3] = {
            {long_tile_kernel<false, 4>, long_tile_kernel<false, 5>, long_tile_kernel<false, 6>},
            {long_tile_kernel<true, 4>, long_tile_kernel<true, 5>, long_tile_kernel<true, 6>}};
        static const unsigned smem_of[2][3] = {
            {DistLayout<4, 4>::TOTAL, DistLayout<5, 5>::TOTAL, DistLayout<6, 6>::TOTAL},
            {StackLayout<4, 4>::TOTAL, StackLayout<5, 5>::TOTAL, StackLayout<6, 6>::TOTAL}};
        const TileKernel tile_kernel = kernels[which][R - 4];
        const unsigned tile_smem = smem_of[which][R - 4] + 32u * R * 4u;
        if (!sms) {
            int dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        }
        if (!tile_grid[which][R - 4]) {
            if (cudaFuncSetAttribute(tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)tile_smem) != cudaSuccess)
                return set_error(ABN_EIO, "%s: cannot reserve %u bytes of shared memory", who, tile_smem);
            int per_sm = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tile_kernel, AL_THREADS, tile_smem);
            tile_grid[which][R - 4] = (per_sm > 0 ? per_sm : 1) * sms;
        }
        const unsigned band_smem = LONG_DTW_WARPS * (24u * (unsigned)max_frames + 32u);
        {
            static unsigned band_smem_set = 0;
            if (band_smem > band_smem_set) {
                if (cudaFuncSetAttribute(dtw_band_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)band_smem) != cudaSuccess)
                    return set_error(ABN_EIO, "%s: cannot reserve %u bytes of shared memory", who, band_smem);
                band_smem_set = band_smem;
                band_grid = 0;
            }
            if (!band_grid) {
                int per_sm = 0;
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dtw_band_kernel,
                                                              LONG_DTW_WARPS * 32, band_smem_set);
                band_grid = (per_sm > 0 ? per_sm : 1) * sms;
                if (band_grid > LONG_DTW_MAX_CTAS) band_grid = LONG_DTW_MAX_CTAS;
            }
        }
        LongArgs la{};
        la.tcap = (max_frames + lt - 1) / lt;
        la.slot_cells = slot;
        la.dirs_bytes = long_dirs_bytes(max_frames);
        la.bad = reinterpret_cast<unsigned *>(ws + ws_bad_off(a.n_pairs));
        la.dirs = ws + ws_dirs_off(a.n_pairs);
        la.dws = dist_only ? nullptr : reinterpret_cast<float *>(ws + d_off);
        const int max_warps = long_dtw_warps(a.n_pairs);
        for (size_t w0 = 0; w0 < (size_t)a.n_pairs; w0 += chunk) {
            la.lw0 = (int)w0;
            la.lw1 = (int)(w0 + chunk < (size_t)a.n_pairs ? w0 + chunk : (size_t)a.n_pairs);
            const int span = la.lw1 - la.lw0;
            if (!dist_only) cudaMemsetAsync(la.bad, 0, sizeof(unsigned) * (size_t)span, st);
            const long long work = (long long)span * la.tcap * la.tcap;
            const int tg = (long long)tile_grid[which][R - 4] < work ? tile_grid[which][R - 4] : (int)work;
            tile_kernel<<<tg, AL_THREADS, tile_smem, st>>>(a, la);
            if (cudaError_t e = cudaGetLastError())
                return set_error(ABN_EIO, "%s: long-token tile kernel launch (smem %u): %s", who,
                                 tile_smem, cudaGetErrorString(e));
            if (dist_only) continue;
            int bg = (span + LONG_DTW_WARPS - 1) / LONG_DTW_WARPS;
            if (bg > band_grid) bg = band_grid;
            if (bg * LONG_DTW_WARPS > max_warps) bg = max_warps / LONG_DTW_WARPS;
            dtw_band_kernel<<<bg, LONG_DTW_WARPS * 32, band_smem, st>>>(a, la, max_frames);
            if (cudaError_t e = cudaGetLastError())
                return set_error(ABN_EIO, "%s: long-token DTW kernel launch (smem %u): %s", who,
                                 band_smem, cudaGetErrorString(e));
        }
    }
    // The fused classes run in rounds over windows of the class-sorted pair order, as many
    // pairs per round as the workspace has hand-over slots for.  Per round and class row:
    // the row's distance kernels (large classes first), then ONE DTW launch over the row.
    const int side = ((max_frames + ext < NM_SHORT ? max_frames + ext : NM_SHORT) + 15) / 16;
    if (!dist_only) a.dws = reinterpret_cast<float *>(ws + d_off);
    for (size_t w0 = 0; w0 < (size_t)a.n_pairs; w0 += chunk) {
        a.w0 = (int)w0;
        a.w1 = (int)(w0 + chunk < (size_t)a.n_pairs ? w0 + chunk : (size_t)a.n_pairs);
        const int span = a.w1 - a.w0;
        for (int ra = side; ra >= 1; --ra) {
            for (int ncg = side; ncg >= 1; --ncg) {
                const ClassLaunch &cl = tab[(ra - 1) * NCLS_SIDE + (ncg - 1)];
                cl.kernel<<<cl.grid < span ? cl.grid : span, AL_THREADS, cl.smem, st>>>(a);
                if (cudaError_t e = cudaGetLastError())
                    return set_error(ABN_EIO, "%s: class (%d,%d) launch (grid %d, smem %u): %s", who,
                                     ra, ncg, cl.grid < span ? cl.grid : span, cl.smem,
                                     cudaGetErrorString(e));
            }
            if (dist_only) continue;
            int rc;
            if (ra <= 2) rc = launch_dtw<1>(a, ra, ext, st);
            else if (ra <= 4) rc = launch_dtw<2>(a, ra, ext, st);
            else rc = launch_dtw<3>(a, ra, ext, st);
            if (rc) return rc;
        }
    }
    return check_launch(who);
}

}  // namespace abn

// ------------------------------------------------------------------ C ABI --
using namespace abn;

extern "C" size_t abn_align_workspace_bytes(int n_pairs, int max_frames, int rounds) {
    if (n_pairs < 0) n_pairs = 0;
    if (max_frames < 1) max_frames = 1;
    if (rounds < 1) rounds = 1;
    size_t per_round = ((size_t)n_pairs + rounds - 1) / rounds;
    if (per_round < 1) per_round = 1;
    if (max_frames > NM_LIMIT) max_frames = NM_LIMIT;
    return ws_dist_off(n_pairs, max_frames) + per_round * slot_cells_for(max_frames) * sizeof(float);
}

// kernels one abn_align_pairs call enqueues (for launch accounting in benchmarks)
extern "C" int abn_align_launches(int n_pairs, int max_frames, int stack, size_t workspace_bytes) {
    if (n_pairs <= 0 || max_frames <= 0) return 0;
    const int ext = stack ? 2 * STACK_H : 0;
    if (max_frames > NM_LIMIT) max_frames = NM_LIMIT;
    const size_t slot = slot_cells_for(max_frames) * sizeof(float), d_off = ws_dist_off(n_pairs, max_frames);
    size_t chunk = workspace_bytes > d_off ? (workspace_bytes - d_off) / slot : 0;
    if (chunk < 1) return 0;
    if (chunk > (size_t)n_pairs) chunk = (size_t)n_pairs;
    const int rounds = (int)(((size_t)n_pairs + chunk - 1) / chunk);
    const int side = ((max_frames + ext < NM_SHORT ? max_frames + ext : NM_SHORT) + 15) / 16;
    return 3 + (max_frames + ext > NM_SHORT ? 2 * rounds : 0) + rounds * (side * side + side);
}

extern "C" int abn_align_pairs(const float *feat, int64_t n_rows, int dim,
                               const int32_t *pair_tok, int n_pairs, int max_frames, int stack,
                               const int64_t *path_off, int32_t *idx1, int32_t *idx2,
                               int32_t *path_len, double *cost, uint8_t *valid, void *workspace,
                               size_t workspace_bytes, abn_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    if (n_pairs == 0) return ABN_OK;
    if (!feat || !pair_tok || !path_off || !idx1 || !idx2 || !path_len || !cost || !valid ||
        n_pairs < 0 || dim <= 0 || (dim & 3))
        return set_error(ABN_EINVAL, "abn_align_pairs: bad argument (dim must be a multiple of 4)");
    AlignArgs a{};
    a.feat = feat; a.n_rows = n_rows; a.dim = dim; a.pair_tok = pair_tok; a.n_pairs = n_pairs;
    a.path_off = path_off; a.idx1 = idx1; a.idx2 = idx2; a.path_len = path_len; a.cost = cost;
    a.valid = valid; a.dist_off = nullptr; a.dist_out = nullptr; a.stack = stack;
    return run_align(a, max_frames, workspace, workspace_bytes, (cudaStream_t)stream,
                     "abn_align_pairs");
}

extern "C" int abn_cosine_distance(const float *feat, int64_t n_rows, int dim,
                                   const int32_t *pair_tok, int n_pairs, int max_frames, int stack,
                                   const int64_t *dist_off, float *dist, uint8_t *valid,
                                   void *workspace, size_t workspace_bytes, abn_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    if (n_pairs == 0) return ABN_OK;
    if (!feat || !pair_tok || !dist_off || !dist || !valid || n_pairs < 0 || dim <= 0 || (dim & 3))
        return set_error(ABN_EINVAL, "abn_cosine_distance: bad argument");
    AlignArgs a{};
    a.feat = feat; a.n_rows = n_rows; a.dim = dim; a.pair_tok = pair_tok; a.n_pairs = n_pairs;
    a.valid = valid; a.dist_off = dist_off; a.dist_out = dist; a.stack = stack;
    return run_align(a, max_frames, workspace, workspace_bytes, (cudaStream_t)stream,
                     "abn_cosine_distance");
}

extern "C" int abn_stack_upload(float *feat_dev, const float *feat_host, int64_t n_rows, int dim,
                                int stack, const uint8_t *last_row_of_file, abn_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    if (!feat_dev || !feat_host || n_rows < 0 || dim <= 0 || stack < 3 || !(stack & 1) || dim % stack ||
        ((dim / stack) & 3))
        return set_error(ABN_EINVAL, "abn_stack_upload: bad argument (odd stack >= 3 dividing dim)");
    if (n_rows == 0) return ABN_OK;
    const int f = dim / stack, h = stack / 2;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemcpy2DAsync(feat_dev + h * f, (size_t)dim * 4, feat_host + h * f,
                                      (size_t)dim * 4, (size_t)f * 4, (size_t)n_rows,
                                      cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess)
        return set_error(ABN_EIO, "abn_stack_upload: cudaMemcpy2DAsync: %s", cudaGetErrorString(e));
    const int wpb = 8;
    restack_kernel<<<(unsigned)((n_rows + wpb - 1) / wpb), wpb * 32, 0, st>>>(feat_dev, n_rows, dim,
                                                                              stack, last_row_of_file);
    return check_launch("abn_stack_upload");
}

namespace abn {
// row t of the stacked table from the un-stacked frames: block c = frames[t + c - h] inside the
// file, zeros outside (abnet3/features.py:135-159); one warp per row, 16-byte accesses
__global__ void stack_build_kernel(float *__restrict__ feat, const float *__restrict__ frames,
                                   int64_t n_rows, int f, int stack,
                                   const uint8_t *__restrict__ last_row_of_file) {
    const int warps_per_block = blockDim.x >> 5;
    const int64_t t = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    if (t >= n_rows) return;
    const int lane = threadIdx.x & 31, h = stack / 2, f4 = f >> 2;
    int back = 0, fwd = 0;
    while (back < h && t - back - 1 >= 0 && !(last_row_of_file && last_row_of_file[t - back - 1])) ++back;
    while (fwd < h && t + fwd + 1 < n_rows && !(last_row_of_file && last_row_of_file[t + fwd])) ++fwd;
    float4 *row = reinterpret_cast<float4 *>(feat + (size_t)t * f * stack);
    for (int e = lane; e < stack * f4; e += 32) {
        const int c = e / f4, k = e - c * f4, d = c - h;
        const bool inside = d < 0 ? -d <= back : d <= fwd;
        row[e] = inside ? __ldg(reinterpret_cast<const float4 *>(frames + (size_t)(t + d) * f) + k)
                        : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
}  // namespace abn

extern "C" int abn_stack_from_frames(float *feat_dev, const float *frames, int64_t n_rows, int f,
                                     int stack, const uint8_t *last_row_of_file, abn_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    if (!feat_dev || !frames || n_rows < 0 || f <= 0 || (f & 3) || stack < 3 || !(stack & 1))
        return set_error(ABN_EINVAL, "abn_stack_from_frames: bad argument (odd stack >= 3, f %% 4 == 0)");
    if (n_rows == 0) return ABN_OK;
    const int wpb = 8;
    stack_build_kernel<<<(unsigned)((n_rows + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
        feat_dev, frames, n_rows, f, stack, last_row_of_file);
    return check_launch("abn_stack_from_frames");
}

namespace abn {
// one warp per pair: direction of step k = (idx1[k] - idx1[k-1], idx2[k] - idx2[k-1])
__global__ void pack_directions_kernel(const int32_t *__restrict__ idx1, const int32_t *__restrict__ idx2,
                                       const int64_t *__restrict__ dst_off,
                                       const int32_t *__restrict__ path_len, int n_pairs,
                                       const int64_t *__restrict__ dir_off, uint8_t *__restrict__ dirs) {
    const int warps_per_block = blockDim.x >> 5;
    const int p = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    if (p >= n_pairs) return;
    const int lane = threadIdx.x & 31;
    const int nd = path_len[p] - 1;
    if (nd <= 0) return;
    const int32_t *a = idx1 + dst_off[p], *b = idx2 + dst_off[p];
    uint8_t *out = dirs + dir_off[p];
    for (int byte = lane; byte < (nd + 3) / 4; byte += 32) {
        unsigned v = 0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = 4 * byte + u + 1;
            if (k <= nd) {
                const int di = a[k] - a[k - 1], dj = b[k] - b[k - 1];
                const unsigned d = (di && dj) ? DIR_DIAG : (di ? DIR_UP : DIR_LEFT);
                v |= d << (2 * u);
            }
        }
        out[byte] = (uint8_t)v;
    }
}
}  // namespace abn

extern "C" int abn_pack_directions(const int32_t *idx1, const int32_t *idx2, const int64_t *dst_off,
                                   const int32_t *path_len, int n_pairs, const int64_t *dir_off,
                                   uint8_t *dirs, abn_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    if (n_pairs == 0) return ABN_OK;
    if (!idx1 || !idx2 || !dst_off || !path_len || !dir_off || !dirs || n_pairs < 0)
        return set_error(ABN_EINVAL, "abn_pack_directions: bad argument");
    const int wpb = 8;
    abn::pack_directions_kernel<<<(n_pairs + wpb - 1) / wpb, wpb * 32, 0, (cudaStream_t)stream>>>(
        idx1, idx2, dst_off, path_len, n_pairs, dir_off, dirs);
    return check_launch("abn_pack_directions");
}

extern "C" int abn_stack_violations(const float *feat, int64_t n_rows, int dim, int stack,
                                    const uint8_t *last_row_of_file, unsigned long long *count,
                                    abn_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    if (!feat || !count || n_rows < 0 || dim <= 0 || stack < 2 || dim % stack)
        return set_error(ABN_EINVAL, "abn_stack_violations: bad argument");
    if (n_rows < 2) return ABN_OK;
    const int wpb = 8;
    stack_violations_kernel<<<(unsigned)((n_rows + wpb - 1) / wpb), wpb * 32, 0,
                              (cudaStream_t)stream>>>(feat, n_rows, dim, stack, last_row_of_file,
                                                      count);
    return check_launch("abn_stack_violations");
}

extern "C" int abn_dtw_from_dist(const double *dist, const int64_t *dist_off, const int32_t *shape,
                                 int n_pairs, int max_frames, const int64_t *path_off,
                                 int32_t *path1, int32_t *path2, int32_t *path_len, double *cost,
                                 uint8_t *valid, abn_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    if (n_pairs == 0) return ABN_OK;
    if (!dist || !dist_off || !shape || !path_off || !path1 || !path2 || !path_len || !cost ||
        !valid || n_pairs < 0)
        return set_error(ABN_EINVAL, "abn_dtw_from_dist: bad argument");
    if (max_frames <= 0)
        return set_error(ABN_EINVAL, "abn_dtw_from_dist: max_frames must be positive");
    if (max_frames > NM_SHORT)
        return set_error(ABN_ERANGE, "abn_dtw_from_dist: matrix side %d exceeds %d", max_frames,
                         NM_SHORT);
    const int nm = ((max_frames < 16 ? 16 : max_frames) + 15) & ~15;
    const DtwLayout L = dtw_layout(nm);
    cudaFuncSetAttribute(dtw_from_dist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)L.total);
    dtw_from_dist_kernel<<<n_pairs, AL_THREADS, L.total, (cudaStream_t)stream>>>(
        dist, dist_off, shape, n_pairs, path_off, path1, path2, path_len, cost, valid, nm);
    return check_launch("abn_dtw_from_dist");
}

extern "C" int abn_diff_pairs(const int32_t *pair_tok, int n_pairs, int stretch,
                              const int64_t *out_off, int32_t *idx1, int32_t *idx2,
                              abn_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    if (n_pairs == 0) return ABN_OK;
    if (!pair_tok || !out_off || !idx1 || !idx2 || n_pairs < 0)
        return set_error(ABN_EINVAL, "abn_diff_pairs: bad argument");
    const int wpb = 8;
    diff_pairs_kernel<<<(n_pairs + wpb - 1) / wpb, wpb * 32, 0, (cudaStream_t)stream>>>(
        pair_tok, n_pairs, stretch, out_off, idx1, idx2);
    return check_launch("abn_diff_pairs");
}

extern "C" int abn_compact_paths(const int32_t *src1, const int32_t *src2, const int64_t *src_off,
                                 const int64_t *dst_off, const int32_t *path_len, int n_pairs,
                                 int32_t *dst1, int32_t *dst2, abn_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    if (n_pairs == 0) return ABN_OK;
    if (!src1 || !src2 || !src_off || !dst_off || !path_len || !dst1 || !dst2 || n_pairs < 0)
        return set_error(ABN_EINVAL, "abn_compact_paths: bad argument");
    const int wpb = 8;
    compact_paths_kernel<<<(n_pairs + wpb - 1) / wpb, wpb * 32, 0, (cudaStream_t)stream>>>(
        src1, src2, src_off, dst_off, path_len, n_pairs, dst1, dst2);
    return check_launch("abn_compact_paths");
}

namespace abn {
__global__ void store_scalar64_kernel(const long long *src, long long *dst) { *dst = *src; }
}  // namespace abn

extern "C" int abn_store_scalar64(const int64_t *src, int64_t *dst_host_mapped, abn_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    if (!src || !dst_host_mapped) return set_error(ABN_EINVAL, "abn_store_scalar64: bad argument");
    abn::store_scalar64_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const long long *>(src), reinterpret_cast<long long *>(dst_host_mapped));
    return check_launch("abn_store_scalar64");
}

extern "C" int abn_gather_batch(const float *feat, int dim, const int32_t *idx1,
                                const int32_t *idx2, const int8_t *y_in, const int64_t *sel,
                                int64_t n, float *x1, float *x2, float *y_out,
                                abn_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    if (n == 0) return ABN_OK;
    if (!feat || !idx1 || !idx2 || !x1 || !x2 || n < 0 || dim <= 0 || (dim & 3))
        return set_error(ABN_EINVAL, "abn_gather_batch: bad argument");
    const int wpb = 8;
    const int64_t warps = 2 * n;
    gather_batch_kernel<<<(unsigned)((warps + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
        feat, dim, idx1, idx2, y_in, sel, n, x1, x2, y_out);
    return check_launch("abn_gather_batch");
}
