// abn_align.cu -- kernels (1) cosine frame distance and (2) DTW wavefront +
// traceback, fused: one CTA aligns one token pair; the distance matrix, the
// accumulated costs and the traceback directions never leave the SM.
//
// Reference behaviour reproduced (paths relative to /root/reference):
//   cosine_distance     abnet3/utils.py:40-60   (float32 arithmetic, zero-norm
//                       rules :55-58, NaN / negative -> pair invalid :59)
//   DTW + traceback     external dtw.DTW called at abnet3/utils.py:149-151;
//                       recurrence and tie rule as oracle/dtw_oracle.c
//   get_dtw_alignment   abnet3/utils.py:147-153, batched over the pair list
//                       like abnet3/dataloader.py:183-206 / :642-653
//
// Data flow per CTA (128 threads):
//   HBM --cp.async 16 B--> smem K-chunks of both tokens (double buffered)
//       --LDS.128--> register-tiled fp32 FMA (fixed k order => run-to-run and
//       GPU-count invariant) --> norms, divide, acosf --> D in smem (aliases
//       the staging buffers) --> warp 0: anti-diagonal wavefront in fp64 with
//       warp shuffles, 1 byte direction per cell in smem --> lane 0 traceback
//       --> all threads write global frame-index pairs.
#include "abn_common.cuh"

namespace abn {

constexpr int AL_THREADS = 128;
constexpr int KC = 40;           // floats of K staged per chunk (one fbank frame of the 7-stack)
constexpr int KCP = KC + 4;      // smem row stride: 11 x 16 B, odd => LDS.128 conflict-free
constexpr int KCP4 = KCP / 4;
constexpr int NM_LIMIT = 96;     // longest token of the single-tile kernel
constexpr float PI_F = 3.14159274101257324f;   // float32(np.pi)

struct AlignLayout {
    int nm, ldd;
    unsigned stage_bytes, dirs_off, norms_off, path_off, misc_off, total;
};

__host__ __device__ inline AlignLayout align_layout(int nm, int dist_elem_bytes) {
    AlignLayout L;
    L.nm = nm;
    L.ldd = nm + 2;  // (ldd - 1) odd: the wavefront's lane stride is bank-conflict free
    L.stage_bytes = 2u * nm * KCP * 4u;
    unsigned dbytes = (unsigned)nm * L.ldd * dist_elem_bytes;
    L.dirs_off = (dbytes + 15u) & ~15u;
    unsigned alias_end = L.dirs_off + (unsigned)nm * nm;
    unsigned region = 2u * L.stage_bytes;
    if (alias_end > region) region = alias_end;
    L.norms_off = (region + 15u) & ~15u;
    L.path_off = L.norms_off + 2u * nm * 4u;
    L.misc_off = L.path_off + 2u * nm * 2u;
    L.total = L.misc_off + 32u;
    return L;
}

// ---------------------------------------------------------------- staging --
__device__ __forceinline__ void stage_chunk(float *buf, const float *g1, const float *g2,
                                            int n1, int n2, int nm, int dim, int k0, int kc4,
                                            int tid) {
    const int piece = tid & 15, r0 = tid >> 4;
    if (piece < kc4) {
        const float *s1 = g1 + k0 + piece * 4;
        float *d1 = buf + piece * 4;
        for (int r = r0; r < n1; r += AL_THREADS / 16)
            cp_async16(d1 + r * KCP, s1 + (size_t)r * dim);
        const float *s2 = g2 + k0 + piece * 4;
        float *d2 = buf + nm * KCP + piece * 4;
        for (int r = r0; r < n2; r += AL_THREADS / 16)
            cp_async16(d2 + r * KCP, s2 + (size_t)r * dim);
    }
}

__device__ __forceinline__ float row_sumsq(const float *row, int kc4, float acc) {
    const float4 *p = reinterpret_cast<const float4 *>(row);
    for (int k4 = 0; k4 < kc4; ++k4) {
        float4 v = p[k4];
        acc = fmaf(v.x, v.x, acc);
        acc = fmaf(v.y, v.y, acc);
        acc = fmaf(v.z, v.z, acc);
        acc = fmaf(v.w, v.w, acc);
    }
    return acc;
}

// ------------------------------------------------------- distance (kernel 1)
// Thread grid 16 (rows) x 8 (cols): thread (ti, tj) owns rows ti + 16 r and
// columns tj + 8 c.  Inside a warp that is 8 consecutive rows x 4 consecutive
// columns, so every LDS.128 is one conflict-free wavefront with broadcast.
template <int RA, int NCG>   // NCG = 16-column groups
__device__ __forceinline__ void pair_distance(unsigned char *smem, const AlignLayout &L,
                                              const float *g1, const float *g2, int n1, int n2,
                                              int dim, float *dist_smem, float *dist_gmem,
                                              int &bad) {
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int ti = (warp >> 1) * 8 + (lane >> 2);
    const int tj = (warp & 1) * 4 + (lane & 3);
    const int nm = L.nm;
    float *stage[2] = {reinterpret_cast<float *>(smem),
                       reinterpret_cast<float *>(smem + L.stage_bytes)};
    float *norms = reinterpret_cast<float *>(smem + L.norms_off);

    float acc[RA][2 * NCG];
#pragma unroll
    for (int r = 0; r < RA; ++r)
#pragma unroll
        for (int c = 0; c < 2 * NCG; ++c) acc[r][c] = 0.f;

    // row-norm ownership: combined row index q in [0, n1 + n2)
    const int q0 = tid, q1 = tid + AL_THREADS;
    const int nq = n1 + n2;
    const int row0 = q0 < n1 ? q0 : nm + (q0 - n1);
    const int row1 = q1 < n1 ? q1 : nm + (q1 - n1);
    float ss0 = 0.f, ss1 = 0.f;

    const int nchunks = (dim + KC - 1) / KC;
    stage_chunk(stage[0], g1, g2, n1, n2, nm, dim, 0, min(KC, dim) / 4, tid);
    cp_async_commit();
    for (int ch = 0; ch < nchunks; ++ch) {
        const int k0 = ch * KC;
        const int kc4 = min(KC, dim - k0) / 4;
        if (ch + 1 < nchunks) {
            const int k1 = k0 + KC;
            stage_chunk(stage[(ch + 1) & 1], g1, g2, n1, n2, nm, dim, k1,
                        min(KC, dim - k1) / 4, tid);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float *buf = stage[ch & 1];
        if (q0 < nq) ss0 = row_sumsq(buf + row0 * KCP, kc4, ss0);
        if (q1 < nq) ss1 = row_sumsq(buf + row1 * KCP, kc4, ss1);

        const float4 *A4 = reinterpret_cast<const float4 *>(buf) + ti * KCP4;
        const float4 *B4 = reinterpret_cast<const float4 *>(buf) + (nm + tj) * KCP4;
#pragma unroll 2
        for (int k4 = 0; k4 < kc4; ++k4) {
            float4 a[RA];
#pragma unroll
            for (int r = 0; r < RA; ++r) a[r] = A4[(16 * r) * KCP4 + k4];
#pragma unroll
            for (int cg = 0; cg < NCG; ++cg) {
                float4 b[2];
#pragma unroll
                for (int c = 0; c < 2; ++c) b[c] = B4[(8 * (2 * cg + c)) * KCP4 + k4];
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
        __syncthreads();   // everyone done with stage[ch & 1] before it is refilled
    }
    if (q0 < nq) norms[row0] = sqrtf(ss0);
    if (q1 < nq) norms[row1] = sqrtf(ss1);
    __syncthreads();

    // epilogue: utils.py:47-58 in float32; the staging buffers are dead, D may overwrite them
#pragma unroll
    for (int r = 0; r < RA; ++r) {
        const int i = ti + 16 * r;
        if (i >= n1) continue;
        const float xn = norms[i];
#pragma unroll
        for (int c = 0; c < 2 * NCG; ++c) {
            const int j = tj + 8 * c;
            if (j >= n2) continue;
            const float yn = norms[nm + j];
            float d;
            if (xn == 0.f || yn == 0.f) {
                d = (xn == 0.f && yn == 0.f) ? 0.f : 1.f;
            } else {
                const float cs = __fdiv_rn(acc[r][c], __fmul_rn(xn, yn));
                d = __fdiv_rn(acosf(cs), PI_F);
            }
            if (!(d >= 0.f)) bad = 1;
            if (dist_gmem) dist_gmem[(size_t)i * n2 + j] = d;
            else dist_smem[i * L.ldd + j] = d;
        }
    }
}

__device__ __forceinline__ void dispatch_distance(unsigned char *smem, const AlignLayout &L,
                                                  const float *g1, const float *g2, int n1,
                                                  int n2, int dim, float *dist_smem,
                                                  float *dist_gmem, int &bad) {
    const int ra = (n1 + 15) >> 4;     // 1..6
    const int ncg = (n2 + 15) >> 4;    // 1..6
#define ABN_CASE(RA_, NCG_)                                                                \
    case (RA_) * 8 + (NCG_):                                                               \
        pair_distance<RA_, NCG_>(smem, L, g1, g2, n1, n2, dim, dist_smem, dist_gmem, bad); \
        break;
#define ABN_ROW(RA_) ABN_CASE(RA_, 1) ABN_CASE(RA_, 2) ABN_CASE(RA_, 3) \
                     ABN_CASE(RA_, 4) ABN_CASE(RA_, 5) ABN_CASE(RA_, 6)
    switch (ra * 8 + ncg) {
        ABN_ROW(1) ABN_ROW(2) ABN_ROW(3) ABN_ROW(4) ABN_ROW(5) ABN_ROW(6)
        default: break;
    }
#undef ABN_ROW
#undef ABN_CASE
}

// ------------------------------------------------------------ DTW (kernel 2)
// One warp per pair.  Lane l owns rows G*l .. G*l+G-1; at step t every row i
// handles cell (i, t - i): an anti-diagonal sweep.  Within a lane the upper
// neighbour is a register; across lanes it is one fp64 shuffle per step.
// C[i,j] = D[i,j] + min(C[i-1,j-1], C[i-1,j], C[i,j-1]) with the oracle's tie
// order (diag, up, left): one add per cell, so results are bit-identical to
// the sequential recurrence for any float64 D.
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
        if (lane == 0) up0 = INF;
        const double dg0 = nbprev;
        nbprev = up0;
        double nw[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int i = i0 + g, j = t - i;
            const double up = g == 0 ? up0 : cur[g - 1];
            const double dg = g == 0 ? dg0 : prev[g - 1];
            const double lf = cur[g];
            nw[g] = lf;
            if (i < n1 && j >= 0 && j < n2) {
                const double d = (double)D[i * ldd + j];
                uint8_t dir;
                double m;
                if (dg <= up && dg <= lf) { dir = DIR_DIAG; m = dg; }
                else if (up <= lf)        { dir = DIR_UP;   m = up; }
                else                      { dir = DIR_LEFT; m = lf; }
                if ((i | j) == 0) m = 0.0;
                nw[g] = d + m;
                dirs[i * ldr + j] = dir;
            }
        }
#pragma unroll
        for (int g = 0; g < G; ++g) { prev[g] = cur[g]; cur[g] = nw[g]; }
    }
    // C[n1-1, n2-1] lives in lane (n1-1)/G, slot (n1-1)%G
    double c = 0.0;
#pragma unroll
    for (int g = 0; g < G; ++g)
        if ((n1 - 1) % G == g) c = cur[g];
    return __shfl_sync(0xffffffffu, c, (n1 - 1) / G);
}

template <typename DT>
__device__ __forceinline__ double dtw_dispatch(const DT *D, int ldd, uint8_t *dirs, int ldr,
                                               int n1, int n2, int lane) {
    const int g = (n1 + 31) >> 5;
    if (g == 1) return dtw_wavefront<DT, 1>(D, ldd, dirs, ldr, n1, n2, lane);
    if (g == 2) return dtw_wavefront<DT, 2>(D, ldd, dirs, ldr, n1, n2, lane);
    return dtw_wavefront<DT, 3>(D, ldd, dirs, ldr, n1, n2, lane);
}

// lane 0: follow the stored directions from (n1-1, n2-1) back to (0, 0)
__device__ __forceinline__ int traceback(const uint8_t *dirs, int ldr, int n1, int n2,
                                         uint8_t *pb_i, uint8_t *pb_j) {
    int i = n1 - 1, j = n2 - 1, L = 0;
    pb_i[0] = (uint8_t)i; pb_j[0] = (uint8_t)j; L = 1;
    while (i > 0 || j > 0) {
        const uint8_t d = dirs[i * ldr + j];
        i -= (d != DIR_LEFT);
        j -= (d != DIR_UP);
        pb_i[L] = (uint8_t)i; pb_j[L] = (uint8_t)j; ++L;
    }
    return L;
}

// --------------------------------------------------------------- kernels ---
__global__ void __launch_bounds__(AL_THREADS, 3)
align_pairs_kernel(const float *__restrict__ feat, int64_t n_rows, int dim,
                   const int32_t *__restrict__ pair_tok, int n_pairs,
                   const int64_t *__restrict__ path_off, int32_t *__restrict__ idx1,
                   int32_t *__restrict__ idx2, int32_t *__restrict__ path_len,
                   double *__restrict__ cost, uint8_t *__restrict__ valid, int nm,
                   const int64_t *__restrict__ dist_off, float *__restrict__ dist_out) {
    // dist_out != nullptr: distance-only mode (abn_cosine_distance) -- D goes to
    // global memory and the DTW stage is skipped.
    extern __shared__ __align__(16) unsigned char smem[];
    const AlignLayout L = align_layout(nm, 4);
    const int p = blockIdx.x;
    if (p >= n_pairs) return;
    const int4 tk = reinterpret_cast<const int4 *>(pair_tok)[p];
    const int s1 = tk.x, n1 = tk.y, s2 = tk.z, n2 = tk.w;
    const int tid = threadIdx.x;
    const bool shape_ok = n1 > 0 && n2 > 0 && n1 <= nm && n2 <= nm && s1 >= 0 && s2 >= 0 &&
                          (int64_t)s1 + n1 <= n_rows && (int64_t)s2 + n2 <= n_rows;
    if (!shape_ok) {   // dataloader.py:184 / :188-191: the pair is skipped
        if (tid == 0) {
            valid[p] = 0;
            if (!dist_out) { path_len[p] = 0; cost[p] = nan(""); }
        }
        return;
    }
    float *Ds = reinterpret_cast<float *>(smem);
    uint8_t *dirs = smem + L.dirs_off;
    uint8_t *pb_i = smem + L.path_off;
    uint8_t *pb_j = pb_i + 2 * nm;
    int *misc = reinterpret_cast<int *>(smem + L.misc_off);

    int bad = 0;
    dispatch_distance(smem, L, feat + (size_t)s1 * dim, feat + (size_t)s2 * dim, n1, n2, dim, Ds,
                      dist_out ? dist_out + dist_off[p] : nullptr, bad);
    bad = __syncthreads_or(bad);
    if (dist_out) {
        if (tid == 0) valid[p] = bad ? 0 : 1;
        return;
    }
    if (bad) {         // utils.py:59 assert fails -> dataloader.py:190-191 drops the pair
        if (tid == 0) { path_len[p] = 0; cost[p] = nan(""); valid[p] = 0; }
        return;
    }
    if (tid < 32) {
        const double c = dtw_dispatch<float>(Ds, L.ldd, dirs, nm, n1, n2, tid);
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
        idx1[off + k] = s1 + (int)pb_i[len - 1 - k];
        idx2[off + k] = s2 + (int)pb_j[len - 1 - k];
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
        const double c = dtw_dispatch<double>(Ds, L.ldd, dirs, nm, n1, n2, tid);
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
        // y[k] = k*step (+0), last element forced to stop; rint = half-to-even
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

}  // namespace abn

// ------------------------------------------------------------------ C ABI --
using namespace abn;

static int pick_nm(const char *who, int max_frames, int *nm) {
    if (max_frames <= 0) return set_error(ABN_EINVAL, "%s: max_frames must be positive", who);
    if (max_frames > NM_LIMIT)
        return set_error(ABN_ERANGE, "%s: token of %d frames exceeds %d", who, max_frames,
                         NM_LIMIT);
    *nm = ((max_frames < 16 ? 16 : max_frames) + 15) & ~15;
    return ABN_OK;
}

extern "C" int abn_align_pairs(const float *feat, int64_t n_rows, int dim,
                               const int32_t *pair_tok, int n_pairs, int max_frames,
                               const int64_t *path_off, int32_t *idx1, int32_t *idx2,
                               int32_t *path_len, double *cost, uint8_t *valid,
                               abn_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    if (n_pairs == 0) return ABN_OK;
    if (!feat || !pair_tok || !path_off || !idx1 || !idx2 || !path_len || !cost || !valid ||
        n_pairs < 0 || dim <= 0 || (dim & 3))
        return set_error(ABN_EINVAL, "abn_align_pairs: bad argument (dim must be a multiple of 4)");
    cudaStream_t st = (cudaStream_t)stream;
    int nm = 0;
    if (int rc = pick_nm("abn_align_pairs", max_frames, &nm)) return rc;
    const AlignLayout L = align_layout(nm, 4);
    cudaFuncSetAttribute(align_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)L.total);
    align_pairs_kernel<<<n_pairs, AL_THREADS, L.total, st>>>(
        feat, n_rows, dim, pair_tok, n_pairs, path_off, idx1, idx2, path_len, cost, valid, nm,
        nullptr, nullptr);
    return check_launch("abn_align_pairs");
}

extern "C" int abn_cosine_distance(const float *feat, int64_t n_rows, int dim,
                                   const int32_t *pair_tok, int n_pairs, int max_frames,
                                   const int64_t *dist_off, float *dist, uint8_t *valid,
                                   abn_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    if (n_pairs == 0) return ABN_OK;
    if (!feat || !pair_tok || !dist_off || !dist || !valid || n_pairs < 0 || dim <= 0 || (dim & 3))
        return set_error(ABN_EINVAL, "abn_cosine_distance: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    int nm = 0;
    if (int rc = pick_nm("abn_cosine_distance", max_frames, &nm)) return rc;
    const AlignLayout L = align_layout(nm, 4);
    cudaFuncSetAttribute(align_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)L.total);
    align_pairs_kernel<<<n_pairs, AL_THREADS, L.total, st>>>(
        feat, n_rows, dim, pair_tok, n_pairs, nullptr, nullptr, nullptr, nullptr, nullptr, valid,
        nm, dist_off, dist);
    return check_launch("abn_cosine_distance");
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
    cudaStream_t st = (cudaStream_t)stream;
    int nm = 0;
    if (int rc = pick_nm("abn_dtw_from_dist", max_frames, &nm)) return rc;
    const DtwLayout L = dtw_layout(nm);
    cudaFuncSetAttribute(dtw_from_dist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)L.total);
    dtw_from_dist_kernel<<<n_pairs, AL_THREADS, L.total, st>>>(
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
