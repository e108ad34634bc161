// abn_fused.cu -- the memory-bound companions of the tensor-core training step, each
// fused with the format conversion its consumer needs so that no fp32 intermediate makes
// a round trip through HBM:
//   abn_gather_batch_bf16    batch generation (abnet3/dataloader.py:204-205, :673-684)
//                            straight into the bf16 A operand of the first layer
//   abn_pair_loss_dz         coscos2 / cosmargin value (abnet3/loss.py:46-67, :85-105) and
//                            the gradient w.r.t. the PRE-activation of the output layer,
//                            dz = dL/de * act'(e), as bf16 rows (the wgrad / dgrad operand)
//   abn_optimizer_step_fused optimizer.step() of abnet3/trainer.py:240 over the flat
//                            bucket + the bf16 operand copy of every weight matrix + the
//                            zeroing of the gradient bucket for the next step's reductions
#include <cuda.h>
#include <cuda_bf16.h>
#include <string.h>

#include "abn_common.cuh"
#include "abn_drop.cuh"

namespace abn {

// one warp per gathered row; 16-byte loads, 8-byte bf16x4 stores.
// cursor (nullable, int64[2] = {next table row, ticket}): without `sel` the batch is the table rows
// cursor[0] .. cursor[0]+n-1 and the last block to finish advances cursor[0] by n -- an epoch of
// fixed-size batches over a shuffled frame-pair table (abnet3/dataloader.py:686-739) replays as
// ONE CUDA graph with no host-side bookkeeping per batch.  table_rows > 0: a batch that would run
// past the table is skipped (the cursor still advances): the gather of batch k + 1 is issued as a
// prefetch beside the training step of batch k.  loss_acc (nullable, double[1]):
// the previous step's loss (word 0 of zero_me) is added to it before it is cleared, i.e. the
// `train_loss += loss.data[0]` of abnet3/trainer.py:242 without a host synchronisation per step.
__global__ void gather_bf16_kernel(const float *__restrict__ feat, int dim,
                                   const int32_t *__restrict__ idx1,
                                   const int32_t *__restrict__ idx2,
                                   const int8_t *__restrict__ y_in,
                                   const int8_t *__restrict__ y2_in,
                                   const int64_t *__restrict__ sel, int64_t *cursor, int64_t table_rows,
                                   int64_t n,
                                   __nv_bfloat16 *__restrict__ xb, int64_t ldx,
                                   float *__restrict__ y_out, float *__restrict__ y2_out,
                                   unsigned *__restrict__ zero_me, int zero_words,
                                   double *__restrict__ loss_acc, int interleave) {
    pdl_wait();
    if (zero_me && blockIdx.x == 0) {
        if (loss_acc && threadIdx.x == 0) loss_acc[0] += (double)__uint_as_float(zero_me[0]);
        __syncthreads();
        for (int i = threadIdx.x; i < zero_words; i += blockDim.x) zero_me[i] = 0u;
    }
    const int warps_per_block = blockDim.x >> 5;
    const int64_t w = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const int64_t base = (cursor && !sel) ? *reinterpret_cast<volatile int64_t *>(cursor) : 0;
    // a PREFETCH past the end of the table (the sweep's last step) gathers nothing
    const bool in_table = !(cursor && !sel && table_rows > 0 && base + n > table_rows);
    if (w < 2 * n && in_table) {
        const int lane = threadIdx.x & 31;
        const int64_t k = w >> 1;
        const int side = (int)(w & 1);
        const int64_t pos = sel ? sel[k] : base + k;
        const int32_t row = side ? idx2[pos] : idx1[pos];
        const float4 *src = reinterpret_cast<const float4 *>(feat + (size_t)row * dim);
        uint2 *dst = reinterpret_cast<uint2 *>(xb + (size_t)(interleave ? w : (side ? n + k : k)) * ldx);
        const int nv = dim >> 2;
        if (nv <= 96) {            // (280-wide rows: 70 float4) every load in flight before the first store
            float4 v[3];
#pragma unroll
            for (int u = 0; u < 3; ++u)
                if (lane + 32 * u < nv) v[u] = __ldg(src + lane + 32 * u);
#pragma unroll
            for (int u = 0; u < 3; ++u)
                if (lane + 32 * u < nv) {
                    __nv_bfloat162 lo = __floats2bfloat162_rn(v[u].x, v[u].y), hi = __floats2bfloat162_rn(v[u].z, v[u].w);
                    dst[lane + 32 * u] = make_uint2(*reinterpret_cast<unsigned *>(&lo), *reinterpret_cast<unsigned *>(&hi));
                }
        } else {
            for (int c = lane; c < nv; c += 32) {
                const float4 v = __ldg(src + c);
                __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
                dst[c] = make_uint2(*reinterpret_cast<unsigned *>(&lo), *reinterpret_cast<unsigned *>(&hi));
            }
        }
        if (side == 0 && lane == 0) {
            if (y_out) y_out[k] = y_in ? (float)y_in[pos] : 1.f;
            if (y2_out) y2_out[k] = y2_in ? (float)y2_in[pos] : 1.f;
        }
    }
    if (cursor && !sel) {          // every block has read cursor[0] before it takes its ticket
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned long long t = atomicAdd(reinterpret_cast<unsigned long long *>(cursor + 1), 1ull);
            if (t == (unsigned long long)gridDim.x - 1) {
                cursor[1] = 0;
                cursor[0] = base + n;
                __threadfence();
            }
        }
    }
}

constexpr int LZ_WARPS = 8;
constexpr float LZ_EPS = 1e-6f;      // nn.CosineSimilarity(eps=1e-6), loss.py:44

__device__ __forceinline__ float dact_of(float y, int act) {
    if (act == 1) return y * (1.f - y);
    if (act == 2) return 1.f - y * y;
    if (act == 3) return y > 0.f ? 1.f : 0.f;
    return 1.f;
}

// one warp per frame pair (same arithmetic as pair_loss_kernel in abn_nn.cu)
__global__ void __launch_bounds__(LZ_WARPS * 32)
pair_loss_dz_kernel(const float *__restrict__ e1, const float *__restrict__ e2,
                    const float *__restrict__ y, int64_t n, int dim, int64_t ld, int kind,
                    float margin, float scale, int act, float *__restrict__ loss,
                    __nv_bfloat16 *__restrict__ dz1, __nv_bfloat16 *__restrict__ dz2,
                    int64_t ld_dz, const DropArgs drop, long long row2_off, int col_off) {
    __shared__ float wsum[LZ_WARPS];
    pdl_wait();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long dkey = drop.state ? drop_key(drop) : 0ull;
    float local = 0.f;
    for (int64_t row = (int64_t)blockIdx.x * LZ_WARPS + warp; row < n;
         row += (int64_t)gridDim.x * LZ_WARPS) {
        const float *a = e1 + row * ld, *b = e2 + row * ld;
        float dot = 0.f, na = 0.f, nb = 0.f;
        for (int k = lane; k < dim; k += 32) {
            const float x = a[k], z = b[k];
            dot = fmaf(x, z, dot); na = fmaf(x, x, na); nb = fmaf(z, z, nb);
        }
        dot = warp_sum(dot); na = warp_sum(na); nb = warp_sum(nb);
        const float ra = sqrtf(na), rb = sqrtf(nb);
        const float an = fmaxf(ra, LZ_EPS), bn = fmaxf(rb, LZ_EPS);
        const float inv = 1.f / (an * bn);
        const float c = dot * inv;
        const float lab = y[row];
        float term, dldc;
        if (kind == 0) {          // coscos2, loss.py:59-62
            if (lab == 1.f)       { term = 0.5f * (1.f - c); dldc = -0.5f; }
            else if (lab == -1.f) { term = c * c;            dldc = 2.f * c; }
            else                  { term = c;                dldc = 1.f; }
        } else {                  // cosmargin, loss.py:98-101
            if (lab == 1.f)       { term = 1.f - c;          dldc = -1.f; }
            else if (lab == -1.f) { const float h = c - margin;
                                    term = fmaxf(h, 0.f);    dldc = h > 0.f ? 1.f : 0.f; }
            else                  { term = c;                dldc = 1.f; }
        }
        if (lane == 0) local += term;
        const float g = dldc * scale;
        const float ka = ra > LZ_EPS ? c / (an * an) : 0.f;
        const float kb = rb > LZ_EPS ? c / (bn * bn) : 0.f;
        __nv_bfloat16 *ga = dz1 + row * ld_dz, *gb = dz2 + row * ld_dz;
        for (int k = lane; k < dim; k += 32) {
            const float x = a[k], z = b[k];      // L1 hits: read a moment ago
            float da = g * (z * inv - ka * x) * dact_of(x, act);
            float db = g * (x * inv - kb * z) * dact_of(z, act);
            if (drop.state) {       // the output layer's dropout (abnet3/model.py:136-141)
                da = drop_keep(dkey, row, col_off + k, drop.thresh) ? da * drop.inv_keep : 0.f;
                db = drop_keep(dkey, row2_off + row, col_off + k, drop.thresh) ? db * drop.inv_keep : 0.f;
            }
            ga[k] = __float2bfloat16_rn(da);
            gb[k] = __float2bfloat16_rn(db);
        }
    }
    if (lane == 0) wsum[warp] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < LZ_WARPS; ++w) s += wsum[w];
        if (s != 0.f) atomicAdd(loss, s * scale);
    }
}

// Vector form for dim <= 128, dim % 4 == 0, 16-byte aligned rows: EIGHT lanes per frame pair (four
// pairs per warp at a time), float4 loads kept in registers between the reduction and the
// gradient pass, 8-byte bf16x4 stores.  Same arithmetic per element as the kernel above; the
// three sums of a pair are reduced over 8 lanes instead of 32 (a different fp32 summation
// order: the parity tests bound both against the oracle).
template <bool DROP>
__global__ void __launch_bounds__(LZ_WARPS * 32)
pair_loss_dz_vec_kernel(const float *__restrict__ e1, const float *__restrict__ e2,
                        const float *__restrict__ y, int64_t n, int dim, int64_t ld, int kind,
                        float margin, float scale, int act, float *__restrict__ loss,
                        __nv_bfloat16 *__restrict__ dz1, __nv_bfloat16 *__restrict__ dz2,
                        int64_t ld_dz, const DropArgs drop, long long row2_off, int col_off) {
    __shared__ float wsum[LZ_WARPS];
    pdl_wait();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane & 7, slot = lane >> 3;
    const unsigned long long dkey = (DROP && drop.state) ? drop_key(drop) : 0ull;
    const int nv = dim >> 2;
    float local = 0.f;
    for (int64_t row0 = ((int64_t)blockIdx.x * LZ_WARPS + warp) * 4; row0 < n;
         row0 += (int64_t)gridDim.x * LZ_WARPS * 4) {
        const int64_t row = row0 + slot;
        const bool live = row < n;
        const float4 *a4 = reinterpret_cast<const float4 *>(e1 + (live ? row : 0) * ld);
        const float4 *b4 = reinterpret_cast<const float4 *>(e2 + (live ? row : 0) * ld);
        float4 av[4], bv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int c = sub + 8 * u;
            av[u] = (live && c < nv) ? a4[c] : make_float4(0.f, 0.f, 0.f, 0.f);
            bv[u] = (live && c < nv) ? b4[c] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float dot = 0.f, na = 0.f, nb = 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            dot = fmaf(av[u].x, bv[u].x, dot); na = fmaf(av[u].x, av[u].x, na); nb = fmaf(bv[u].x, bv[u].x, nb);
            dot = fmaf(av[u].y, bv[u].y, dot); na = fmaf(av[u].y, av[u].y, na); nb = fmaf(bv[u].y, bv[u].y, nb);
            dot = fmaf(av[u].z, bv[u].z, dot); na = fmaf(av[u].z, av[u].z, na); nb = fmaf(bv[u].z, bv[u].z, nb);
            dot = fmaf(av[u].w, bv[u].w, dot); na = fmaf(av[u].w, av[u].w, na); nb = fmaf(bv[u].w, bv[u].w, nb);
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            dot += __shfl_xor_sync(0xffffffffu, dot, o);
            na += __shfl_xor_sync(0xffffffffu, na, o);
            nb += __shfl_xor_sync(0xffffffffu, nb, o);
        }
        const float ra = sqrtf(na), rb = sqrtf(nb);
        const float an = fmaxf(ra, LZ_EPS), bn = fmaxf(rb, LZ_EPS);
        const float inv = 1.f / (an * bn);
        const float c = dot * inv;
        const float lab = live ? y[row] : 0.f;
        float term, dldc;
        if (kind == 0) {          // coscos2, loss.py:59-62
            if (lab == 1.f)       { term = 0.5f * (1.f - c); dldc = -0.5f; }
            else if (lab == -1.f) { term = c * c;            dldc = 2.f * c; }
            else                  { term = c;                dldc = 1.f; }
        } else {                  // cosmargin, loss.py:98-101
            if (lab == 1.f)       { term = 1.f - c;          dldc = -1.f; }
            else if (lab == -1.f) { const float h = c - margin;
                                    term = fmaxf(h, 0.f);    dldc = h > 0.f ? 1.f : 0.f; }
            else                  { term = c;                dldc = 1.f; }
        }
        if (live && sub == 0) local += term;
        const float g = dldc * scale;
        const float ka = ra > LZ_EPS ? c / (an * an) : 0.f;
        const float kb = rb > LZ_EPS ? c / (bn * bn) : 0.f;
        if (live) {
            uint2 *ga = reinterpret_cast<uint2 *>(dz1 + row * ld_dz);
            uint2 *gb = reinterpret_cast<uint2 *>(dz2 + row * ld_dz);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int cc = sub + 8 * u;
                if (cc < nv) {
                    const float4 x = av[u], z = bv[u];
                    float da[4] = {g * (z.x * inv - ka * x.x) * dact_of(x.x, act),
                                   g * (z.y * inv - ka * x.y) * dact_of(x.y, act),
                                   g * (z.z * inv - ka * x.z) * dact_of(x.z, act),
                                   g * (z.w * inv - ka * x.w) * dact_of(x.w, act)};
                    float db[4] = {g * (x.x * inv - kb * z.x) * dact_of(z.x, act),
                                   g * (x.y * inv - kb * z.y) * dact_of(z.y, act),
                                   g * (x.z * inv - kb * z.z) * dact_of(z.z, act),
                                   g * (x.w * inv - kb * z.w) * dact_of(z.w, act)};
                    if (DROP && drop.state) {       // the output layer's dropout (abnet3/model.py:136-141)
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int col = col_off + 4 * cc + e;
                            da[e] = drop_keep(dkey, row, col, drop.thresh) ? da[e] * drop.inv_keep : 0.f;
                            db[e] = drop_keep(dkey, row2_off + row, col, drop.thresh) ? db[e] * drop.inv_keep : 0.f;
                        }
                    }
                    __nv_bfloat162 a0 = __floats2bfloat162_rn(da[0], da[1]);
                    __nv_bfloat162 a1 = __floats2bfloat162_rn(da[2], da[3]);
                    __nv_bfloat162 b0 = __floats2bfloat162_rn(db[0], db[1]);
                    __nv_bfloat162 b1 = __floats2bfloat162_rn(db[2], db[3]);
                    ga[cc] = make_uint2(*reinterpret_cast<unsigned *>(&a0), *reinterpret_cast<unsigned *>(&a1));
                    gb[cc] = make_uint2(*reinterpret_cast<unsigned *>(&b0), *reinterpret_cast<unsigned *>(&b1));
                }
            }
        }
    }
    local += __shfl_xor_sync(0xffffffffu, local, 8);
    local += __shfl_xor_sync(0xffffffffu, local, 16);
    if (lane == 0) wsum[warp] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < LZ_WARPS; ++w) s += wsum[w];
        if (s != 0.f) atomicAdd(loss, s * scale);
    }
}

struct SegTable { abn_param_segment s[ABN_MAX_PARAM_SEGMENTS]; int n; };

// grid.y = segment; same update rules as optimizer_kernel (abn_nn.cu)
__global__ void optimizer_fused_kernel(float *__restrict__ p, float *__restrict__ g,
                                       float *__restrict__ s0, float *__restrict__ s1, int kind,
                                       float lr, float momentum, float gscale, float bc1,
                                       float bc2_sqrt, int zero_grad, const SegTable tab) {
    pdl_wait();
    const abn_param_segment sg = tab.s[blockIdx.y];
    __nv_bfloat16 *wb = static_cast<__nv_bfloat16 *>(sg.bf16);
    if (((sg.offset | sg.count) & 3) == 0 && (!wb || ((sg.n_in | sg.ld) & 3) == 0)) {
        // 4 consecutive elements per thread: 16-byte accesses on every array
        for (int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; j < sg.count;
             j += (int64_t)gridDim.x * blockDim.x * 4) {
            const int64_t i = sg.offset + j;
            const float4 g4 = *reinterpret_cast<const float4 *>(g + i);
            const float4 w4 = *reinterpret_cast<const float4 *>(p + i);
            float w[4] = {w4.x, w4.y, w4.z, w4.w};
            const float gr[4] = {g4.x * gscale, g4.y * gscale, g4.z * gscale, g4.w * gscale};
            if (kind == 0) {
                if (momentum != 0.f) {
                    const float4 m4 = *reinterpret_cast<const float4 *>(s0 + i);
                    float b[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) { b[e] = momentum * b[e] + gr[e]; w[e] -= lr * b[e]; }
                    *reinterpret_cast<float4 *>(s0 + i) = make_float4(b[0], b[1], b[2], b[3]);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) w[e] -= lr * gr[e];
                }
            } else {
                const float4 a4 = *reinterpret_cast<const float4 *>(s0 + i);
                const float4 b4 = *reinterpret_cast<const float4 *>(s1 + i);
                float sa[4] = {a4.x, a4.y, a4.z, a4.w}, sb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (kind == 1) {        // torch.optim.Adadelta(rho=0.9, eps=1e-6)
                        const float rho = 0.9f, eps = 1e-6f;
                        const float sq = rho * sa[e] + (1.f - rho) * gr[e] * gr[e];
                        const float stdv = sqrtf(sq + eps);
                        const float delta = sqrtf(sb[e] + eps) / stdv * gr[e];
                        sa[e] = sq;
                        sb[e] = rho * sb[e] + (1.f - rho) * delta * delta;
                        w[e] -= lr * delta;
                    } else {                // torch.optim.Adam(betas=(0.9, 0.999), eps=1e-8)
                        const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
                        const float m = b1 * sa[e] + (1.f - b1) * gr[e];
                        const float v = b2 * sb[e] + (1.f - b2) * gr[e] * gr[e];
                        sa[e] = m; sb[e] = v;
                        const float denom = sqrtf(v) / bc2_sqrt + eps;
                        w[e] -= (lr / bc1) * (m / denom);
                    }
                }
                *reinterpret_cast<float4 *>(s0 + i) = make_float4(sa[0], sa[1], sa[2], sa[3]);
                *reinterpret_cast<float4 *>(s1 + i) = make_float4(sb[0], sb[1], sb[2], sb[3]);
            }
            *reinterpret_cast<float4 *>(p + i) = make_float4(w[0], w[1], w[2], w[3]);
            if (zero_grad) *reinterpret_cast<float4 *>(g + i) = make_float4(0.f, 0.f, 0.f, 0.f);
            if (wb) {
                const int64_t r = j / sg.n_in;
                __nv_bfloat162 lo = __floats2bfloat162_rn(w[0], w[1]), hi = __floats2bfloat162_rn(w[2], w[3]);
                *reinterpret_cast<uint2 *>(wb + r * sg.ld + (j - r * sg.n_in)) =
                    make_uint2(*reinterpret_cast<unsigned *>(&lo), *reinterpret_cast<unsigned *>(&hi));
            }
        }
        return;
    }
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < sg.count;
         j += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = sg.offset + j;
        const float grad = g[i] * gscale;
        float w = p[i];
        if (kind == 0) {            // torch.optim.SGD(momentum, dampening=0)
            float buf = grad;
            if (momentum != 0.f) { buf = momentum * s0[i] + grad; s0[i] = buf; }
            w -= lr * buf;
        } else if (kind == 1) {     // torch.optim.Adadelta(rho=0.9, eps=1e-6)
            const float rho = 0.9f, eps = 1e-6f;
            const float sq = rho * s0[i] + (1.f - rho) * grad * grad;
            const float stdv = sqrtf(sq + eps);
            const float delta = sqrtf(s1[i] + eps) / stdv * grad;
            s0[i] = sq;
            s1[i] = rho * s1[i] + (1.f - rho) * delta * delta;
            w -= lr * delta;
        } else {                    // torch.optim.Adam(betas=(0.9, 0.999), eps=1e-8)
            const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
            const float m = b1 * s0[i] + (1.f - b1) * grad;
            const float v = b2 * s1[i] + (1.f - b2) * grad * grad;
            s0[i] = m; s1[i] = v;
            const float denom = sqrtf(v) / bc2_sqrt + eps;
            w -= (lr / bc1) * (m / denom);
        }
        p[i] = w;
        if (zero_grad) g[i] = 0.f;
        if (wb) {
            const int64_t r = j / sg.n_in;
            wb[r * sg.ld + (j - r * sg.n_in)] = __float2bfloat16_rn(w);
        }
    }
}

}  // namespace abn

using namespace abn;

extern "C" int abn_gather_step_bf16(const float *feat, int dim, const int32_t *idx1,
                                    const int32_t *idx2, const int8_t *y_in, const int8_t *y2_in,
                                    const int64_t *sel, int64_t *cursor, int64_t table_rows,
                                    int64_t n, void *xb,
                                    int64_t ldx, float *y_out, float *y2_out, void *zero_me,
                                    int zero_words, double *loss_acc, int interleave,
                                    abn_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    if (n == 0) return ABN_OK;
    if (!feat || !idx1 || !idx2 || !xb || n < 0 || dim <= 0 || (dim & 3) || ldx < dim || (ldx & 3) ||
        (loss_acc && (!zero_me || zero_words < 1)))
        return set_error(ABN_EINVAL, "abn_gather_step_bf16: bad argument");
    const int wpb = 8;
    const int64_t warps = 2 * n;
    launch_pdl(gather_bf16_kernel, dim3((unsigned)((warps + wpb - 1) / wpb)), dim3(wpb * 32), (cudaStream_t)stream,
        feat, dim, idx1, idx2, y_in, y2_in, sel, cursor, table_rows, n, static_cast<__nv_bfloat16 *>(xb), ldx, y_out,
        y2_out, static_cast<unsigned *>(zero_me), zero_me ? zero_words : 0, loss_acc, interleave);
    return check_launch("abn_gather_step_bf16");
}

extern "C" int abn_gather_batch_bf16(const float *feat, int dim, const int32_t *idx1,
                                     const int32_t *idx2, const int8_t *y_in, const int64_t *sel,
                                     int64_t n, void *xb, int64_t ldx, float *y_out,
                                     void *zero_me, int zero_words, abn_stream_t stream) {
    return abn_gather_step_bf16(feat, dim, idx1, idx2, y_in, nullptr, sel, nullptr, 0, n, xb, ldx, y_out,
                                nullptr, zero_me, zero_words, nullptr, 0, stream);
}

extern "C" int abn_pair_loss_dz(const float *e1, const float *e2, const float *y, int64_t n, int dim,
                                int64_t ld, int kind, float margin, float scale, int act,
                                float *loss, void *dz1, void *dz2, int64_t ld_dz,
                                abn_stream_t stream) {
    return abn_pair_loss_dz_drop(e1, e2, y, n, dim, ld, kind, margin, scale, act, loss, dz1, dz2, ld_dz,
                                 nullptr, 0, 0, stream);
}

extern "C" int abn_pair_loss_dz_drop(const float *e1, const float *e2, const float *y, int64_t n, int dim,
                                     int64_t ld, int kind, float margin, float scale, int act,
                                     float *loss, void *dz1, void *dz2, int64_t ld_dz,
                                     const abn_dropout *drop, int64_t row2_offset, int col_offset,
                                     abn_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    const DropArgs da = drop_args(drop);
    const long long r2 = (long long)row2_offset;
    if (n == 0) return ABN_OK;
    if (ld == 0) ld = dim;
    if (!e1 || !e2 || !y || !loss || !dz1 || !dz2 || n < 0 || dim <= 0 || ld < dim ||
        ld_dz < dim || (kind != 0 && kind != 1) || act < 0 || act > 3)
        return set_error(ABN_EINVAL, "abn_pair_loss_dz: bad argument");
    const bool vec = dim <= 128 && !(dim & 3) && !(ld & 3) && !(ld_dz & 3) &&
                     !(((uintptr_t)e1 | (uintptr_t)e2) & 15) && !(((uintptr_t)dz1 | (uintptr_t)dz2) & 7);
    if (vec) {
        int64_t blocks = (n + 4 * LZ_WARPS - 1) / (4 * LZ_WARPS);
        if (blocks > 148 * 8) blocks = 148 * 8;
        if (da.state)
            launch_pdl(pair_loss_dz_vec_kernel<true>, dim3((unsigned)blocks), dim3(LZ_WARPS * 32),
                (cudaStream_t)stream, e1, e2, y, n, dim, ld, kind, margin, scale, act, loss,
                static_cast<__nv_bfloat16 *>(dz1), static_cast<__nv_bfloat16 *>(dz2), ld_dz, da, r2, col_offset);
        else
            launch_pdl(pair_loss_dz_vec_kernel<false>, dim3((unsigned)blocks), dim3(LZ_WARPS * 32),
                (cudaStream_t)stream, e1, e2, y, n, dim, ld, kind, margin, scale, act, loss,
                static_cast<__nv_bfloat16 *>(dz1), static_cast<__nv_bfloat16 *>(dz2), ld_dz, da, r2, col_offset);
        return check_launch("abn_pair_loss_dz");
    }
    int64_t blocks = (n + LZ_WARPS - 1) / LZ_WARPS;
    if (blocks > 148 * 8) blocks = 148 * 8;
    launch_pdl(pair_loss_dz_kernel, dim3((unsigned)blocks), dim3(LZ_WARPS * 32), (cudaStream_t)stream,
        e1, e2, y, n, dim, ld, kind, margin, scale, act, loss, static_cast<__nv_bfloat16 *>(dz1),
        static_cast<__nv_bfloat16 *>(dz2), ld_dz, da, r2, col_offset);
    return check_launch("abn_pair_loss_dz");
}

extern "C" int abn_optimizer_step_fused(float *param, float *grad, float *state0, float *state1,
                                        int kind, float lr, float momentum, float grad_scale,
                                        int64_t step, const abn_param_segment *segments,
                                        int n_segments, int zero_grad, abn_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    if (n_segments == 0) return ABN_OK;
    if (!param || !grad || !segments || n_segments < 0 || n_segments > ABN_MAX_PARAM_SEGMENTS ||
        kind < 0 || kind > 2 || (kind == 0 && momentum != 0.f && !state0) ||
        (kind >= 1 && (!state0 || !state1)) || step < 1)
        return set_error(ABN_EINVAL, "abn_optimizer_step_fused: bad argument (at most %d segments)",
                         ABN_MAX_PARAM_SEGMENTS);
    SegTable tab;
    tab.n = n_segments;
    int64_t longest = 0;
    for (int i = 0; i < n_segments; ++i) {
        tab.s[i] = segments[i];
        if (segments[i].count < 0 || (segments[i].bf16 && segments[i].n_in <= 0))
            return set_error(ABN_EINVAL, "abn_optimizer_step_fused: bad segment %d", i);
        if (segments[i].count > longest) longest = segments[i].count;
    }
    const float bc1 = 1.f - powf(0.9f, (float)step);
    const float bc2s = sqrtf(1.f - powf(0.999f, (float)step));
    int64_t bx = (longest / 4 + 255) / 256;
    if (bx > 148 * 4) bx = 148 * 4;
    if (bx < 1) bx = 1;
    dim3 grid((unsigned)bx, (unsigned)n_segments);
    launch_pdl(optimizer_fused_kernel, grid, dim3(256), (cudaStream_t)stream,
        param, grad, state0, state1, kind, lr, momentum, grad_scale, bc1, bc2s, zero_grad, tab);
    return check_launch("abn_optimizer_step_fused");
}

// ------------------------------------------------------------------------
// Data-parallel step without a library collective: the gradient all-reduce is FUSED into
// the optimizer kernel over NVLink peer memory.  Every rank maps its peers' gradient
// buckets (CUDA IPC) and a small flag block; per step
//   dp_optimizer_kernel            block 0 sets ready[rank] = step (= done[rank] + 1) in every peer's
//                                  flags; all blocks wait until ready[p] >= step for all p, then every thread
//                                  sums its elements over the local + peer buckets (one-shot
//                                  reads over NVLink / NVSwitch), applies the update, writes
//                                  the bf16 weight copy; the last block sets done[rank] = step
//                                  in every peer's flags
//   dp_grad_reset_kernel           (before the next step's wgrad reductions) waits until
//                                  done[p] >= done[rank] for all p -- nobody still reads this
//                                  rank's bucket -- and clears it
// The flags are monotonically increasing step numbers kept on the device, so the whole
// sequence replays inside a CUDA graph.  Replaces dist.all_reduce + optimizer.step().
namespace abn {

constexpr int DP_MAX_WORLD = ABN_DP_MAX_WORLD;
enum { DPF_READY = 0, DPF_DONE = DP_MAX_WORLD, DPF_COUNTER = 2 * DP_MAX_WORLD, DPF_TICKET = 2 * DP_MAX_WORLD + 1 };

struct DpPeers {
    const float *grad[DP_MAX_WORLD];          // every rank's gradient bucket (own one included)
    unsigned long long *flags[DP_MAX_WORLD];  // every rank's flag block
    int rank, world;
};

__device__ __forceinline__ unsigned long long dp_ld_acquire(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void dp_st_release(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// bounded spin: a missing peer must trap, not hang the GPU forever
__device__ __forceinline__ void dp_wait_all(const unsigned long long *slots, int world,
                                            unsigned long long want) {
    for (int p = 0; p < world; ++p) {
        unsigned long long spins = 0;
        while (dp_ld_acquire(slots + p) < want) {
            if (++spins > (1ull << 31)) __trap();
        }
    }
}

// 16-byte uncached loads from a (peer) bucket
__device__ __forceinline__ float4 dp_ld4(const float *p) {
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ float dp_update(float w, float grad, float *s0, float *s1, int64_t i,
                                           int kind, float lr, float momentum, float bc1,
                                           float bc2_sqrt) {
    if (kind == 0) {            // torch.optim.SGD(momentum, dampening=0)
        float buf = grad;
        if (momentum != 0.f) { buf = momentum * s0[i] + grad; s0[i] = buf; }
        return w - lr * buf;
    }
    if (kind == 1) {            // torch.optim.Adadelta(rho=0.9, eps=1e-6)
        const float rho = 0.9f, eps = 1e-6f;
        const float sq = rho * s0[i] + (1.f - rho) * grad * grad;
        const float stdv = sqrtf(sq + eps);
        const float delta = sqrtf(s1[i] + eps) / stdv * grad;
        s0[i] = sq;
        s1[i] = rho * s1[i] + (1.f - rho) * delta * delta;
        return w - lr * delta;
    }
    const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;     // torch.optim.Adam
    const float m = b1 * s0[i] + (1.f - b1) * grad;
    const float v = b2 * s1[i] + (1.f - b2) * grad * grad;
    s0[i] = m; s1[i] = v;
    const float denom = sqrtf(v) / bc2_sqrt + eps;
    return w - (lr / bc1) * (m / denom);
}

__global__ void dp_optimizer_kernel(float *__restrict__ p, float *__restrict__ s0,
                                    float *__restrict__ s1, int kind, float lr, float momentum,
                                    float gscale, float bc1, float bc2_sqrt, const SegTable tab,
                                    const DpPeers pe) {
    unsigned long long *mine = pe.flags[pe.rank];
    __shared__ unsigned long long step_s;
    if (threadIdx.x == 0) {
        // this step's number: one more than the last step this rank completed (done[rank] is
        // only advanced by the LAST block of a launch, after every block has read it)
        const unsigned long long step = dp_ld_acquire(mine + DPF_DONE + pe.rank) + 1;
        if (blockIdx.x == 0 && blockIdx.y == 0) {
            // the stream order put every gradient reduction of this rank before this kernel
            __threadfence_system();
            for (int q = 0; q < pe.world; ++q) dp_st_release(pe.flags[q] + DPF_READY + pe.rank, step);
        }
        dp_wait_all(mine + DPF_READY, pe.world, step);       // every rank's gradients are complete
        step_s = step;
    }
    __syncthreads();
    const abn_param_segment sg = tab.s[blockIdx.y];
    __nv_bfloat16 *wb = static_cast<__nv_bfloat16 *>(sg.bf16);
    const bool vec = ((sg.offset | sg.count) & 3) == 0 && (!wb || (sg.n_in & 3) == 0);
    if (vec) {
        // 4 consecutive elements per thread: 16-byte loads from every bucket, all in flight
        for (int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; j < sg.count;
             j += (int64_t)gridDim.x * blockDim.x * 4) {
            const int64_t i = sg.offset + j;
            float4 part[DP_MAX_WORLD];
#pragma unroll
            for (int r = 0; r < DP_MAX_WORLD; ++r)
                if (r < pe.world) part[r] = dp_ld4(pe.grad[r] + i);
            float4 gs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int r = 0; r < DP_MAX_WORLD; ++r)       // rank order: the same sum on every rank
                if (r < pe.world) { gs.x += part[r].x; gs.y += part[r].y; gs.z += part[r].z; gs.w += part[r].w; }
            const float4 w4 = *reinterpret_cast<const float4 *>(p + i);
            float w[4] = {w4.x, w4.y, w4.z, w4.w};
            const float g[4] = {gs.x * gscale, gs.y * gscale, gs.z * gscale, gs.w * gscale};
#pragma unroll
            for (int e = 0; e < 4; ++e)
                w[e] = dp_update(w[e], g[e], s0, s1, i + e, kind, lr, momentum, bc1, bc2_sqrt);
            *reinterpret_cast<float4 *>(p + i) = make_float4(w[0], w[1], w[2], w[3]);
            if (wb) {
                const int64_t r = j / sg.n_in;
                __nv_bfloat162 lo = __floats2bfloat162_rn(w[0], w[1]), hi = __floats2bfloat162_rn(w[2], w[3]);
                *reinterpret_cast<uint2 *>(wb + r * sg.ld + (j - r * sg.n_in)) =
                    make_uint2(*reinterpret_cast<unsigned *>(&lo), *reinterpret_cast<unsigned *>(&hi));
            }
        }
    } else {
        for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < sg.count;
             j += (int64_t)gridDim.x * blockDim.x) {
            const int64_t i = sg.offset + j;
            float gsum = 0.f;
            for (int r = 0; r < pe.world; ++r) gsum += __ldcv(pe.grad[r] + i);
            const float w = dp_update(p[i], gsum * gscale, s0, s1, i, kind, lr, momentum, bc1, bc2_sqrt);
            p[i] = w;
            if (wb) {
                const int64_t r = j / sg.n_in;
                wb[r * sg.ld + (j - r * sg.n_in)] = __float2bfloat16_rn(w);
            }
        }
    }
    // the last block to finish tells every peer that this rank no longer reads their buckets
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned long long total = (unsigned long long)gridDim.x * gridDim.y;
        const unsigned long long t = atomicAdd(mine + DPF_TICKET, 1ull);
        if (t == total - 1) {
            mine[DPF_TICKET] = 0;
            __threadfence_system();
            for (int q = 0; q < pe.world; ++q)
                dp_st_release(pe.flags[q] + DPF_DONE + pe.rank, step_s);
        }
    }
}

__global__ void dp_grad_reset_kernel(float *__restrict__ g, int64_t n, const DpPeers pe) {
    const unsigned long long *mine = pe.flags[pe.rank];
    if (threadIdx.x == 0) dp_wait_all(mine + DPF_DONE, pe.world, dp_ld_acquire(mine + DPF_DONE + pe.rank));
    __syncthreads();
    float4 *g4 = reinterpret_cast<float4 *>(g);
    const int64_t n4 = n >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
         i += (int64_t)gridDim.x * blockDim.x)
        g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        g[i] = 0.f;
}

static int dp_fill(DpPeers &pe, const abn_dp_peers *peers) {
    if (!peers || peers->world < 2 || peers->world > DP_MAX_WORLD || peers->rank < 0 ||
        peers->rank >= peers->world)
        return set_error(ABN_EINVAL, "data-parallel peers: world must be 2..%d", DP_MAX_WORLD);
    pe.rank = peers->rank;
    pe.world = peers->world;
    for (int r = 0; r < peers->world; ++r) {
        if (!peers->grad[r] || !peers->flags[r])
            return set_error(ABN_EINVAL, "data-parallel peers: rank %d is not mapped", r);
        pe.grad[r] = static_cast<const float *>(peers->grad[r]);
        pe.flags[r] = static_cast<unsigned long long *>(peers->flags[r]);
    }
    return ABN_OK;
}

}  // namespace abn

extern "C" int abn_ipc_export(const void *ptr, unsigned char *handle64, int64_t *offset) {
    if (int rc = require_sm100()) return rc;
    if (!ptr || !handle64 || !offset) return set_error(ABN_EINVAL, "abn_ipc_export: bad argument");
    typedef CUresult (*RangeFn)(CUdeviceptr *, size_t *, CUdeviceptr);
    static RangeFn fn = nullptr;
    if (!fn) {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return set_error(ABN_EIO, "abn_ipc_export: cuMemGetAddressRange is not available");
        fn = reinterpret_cast<RangeFn>(f);
    }
    CUdeviceptr base = 0;
    size_t size = 0;
    if (fn(&base, &size, (CUdeviceptr)ptr) != CUDA_SUCCESS)
        return set_error(ABN_EIO, "abn_ipc_export: cannot find the allocation of %p", ptr);
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, reinterpret_cast<void *>(base));
    if (e != cudaSuccess)
        return set_error(ABN_EIO, "abn_ipc_export: cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    static_assert(sizeof(h) == 64, "CUDA IPC handles are 64 bytes");
    memcpy(handle64, &h, 64);
    *offset = (int64_t)((CUdeviceptr)ptr - base);
    return ABN_OK;
}

extern "C" int abn_ipc_import(const unsigned char *handle64, int64_t offset, void **ptr) {
    if (int rc = require_sm100()) return rc;
    if (!handle64 || !ptr) return set_error(ABN_EINVAL, "abn_ipc_import: bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void *base = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_error(ABN_EIO, "abn_ipc_import: cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
    }
    *ptr = static_cast<unsigned char *>(base) + offset;
    return ABN_OK;
}

extern "C" int abn_dp_optimizer_step(float *param, float *state0, float *state1, int kind, float lr,
                                     float momentum, float grad_scale, int64_t step,
                                     const abn_param_segment *segments, int n_segments,
                                     const abn_dp_peers *peers, abn_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    if (n_segments == 0) return ABN_OK;
    if (!param || !segments || n_segments < 0 || n_segments > ABN_MAX_PARAM_SEGMENTS || kind < 0 ||
        kind > 2 || (kind == 0 && momentum != 0.f && !state0) || (kind >= 1 && (!state0 || !state1)) ||
        step < 1)
        return set_error(ABN_EINVAL, "abn_dp_optimizer_step: bad argument");
    DpPeers pe;
    if (int rc = dp_fill(pe, peers)) return rc;
    SegTable tab;
    tab.n = n_segments;
    int64_t longest = 0;
    for (int i = 0; i < n_segments; ++i) {
        tab.s[i] = segments[i];
        if (segments[i].count > longest) longest = segments[i].count;
    }
    const float bc1 = 1.f - powf(0.9f, (float)step);
    const float bc2s = sqrtf(1.f - powf(0.999f, (float)step));
    int64_t bx = (longest / 4 + 255) / 256;
    if (bx > 148 * 2) bx = 148 * 2;
    if (bx < 1) bx = 1;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)bx, (unsigned)n_segments);
    dp_optimizer_kernel<<<grid, 256, 0, st>>>(param, state0, state1, kind, lr, momentum, grad_scale,
                                              bc1, bc2s, tab, pe);
    return check_launch("abn_dp_optimizer_step");
}

extern "C" int abn_dp_grad_reset(float *grad, int64_t n, const abn_dp_peers *peers,
                                 abn_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    if (!grad || n < 0) return set_error(ABN_EINVAL, "abn_dp_grad_reset: bad argument");
    DpPeers pe;
    if (int rc = dp_fill(pe, peers)) return rc;
    int64_t bx = (n / 4 + 255) / 256;
    if (bx > 148 * 4) bx = 148 * 4;
    if (bx < 1) bx = 1;
    dp_grad_reset_kernel<<<(unsigned)bx, 256, 0, (cudaStream_t)stream>>>(grad, n, pe);
    return check_launch("abn_dp_grad_reset");
}

// ------------------------------------------------------------------------
// Two-shot variant, all transfers are WRITES over NVLink (posted, no round trips):
//   phase 0  every rank pushes slice j of its gradient bucket into owner j's receive
//            buffer (recv[j][rank]) and then raises pushed[rank] = step on every peer
//   phase 1  the owner of slice r sums its own slice and the W-1 received ones in rank
//            order, applies the optimizer to that slice and pushes the updated fp32
//            parameters into EVERY rank's parameter bucket; raises updated[rank] = step
//   phase 2  once all slices have arrived: bf16 operand copies of the weights, gradient
//            bucket cleared for the next step's reductions
// One kernel, one block per SM (every block must be resident: they wait on flags that
// other blocks of the same grid help to raise).  Per rank and step: 2 (W-1)/W x the bucket
// leaves over NVLink, as in a ring all-reduce, in two hops instead of 2 (W-1).
namespace abn {

enum { DPF_PUSHED = 0, DPF_UPDATED = DP_MAX_WORLD, DPF_STEP = 2 * DP_MAX_WORLD,
       DPF_TICKET_A = 2 * DP_MAX_WORLD + 1, DPF_TICKET_B = 2 * DP_MAX_WORLD + 2,
       DPF_TICKET_C = 2 * DP_MAX_WORLD + 3 };

// debug timeline (tools/dp_trace.py): globaltimer stamps of block 0, [step % 64][8]
__device__ long long *g_dp_trace = nullptr;
__device__ __forceinline__ void dp_stamp(unsigned long long step, int slot) {
    if (g_dp_trace && blockIdx.x == 0 && threadIdx.x == 0) {
        long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_dp_trace[(step & 63ull) * 8 + slot] = t;
    }
}

struct DpPush {
    float *param[DP_MAX_WORLD];
    float *recv[DP_MAX_WORLD];
    unsigned long long *flags[DP_MAX_WORLD];
    int rank, world;
    long long n, cap;
    int one_shot;
};

__device__ __forceinline__ void dp_st4(float *p, float4 v) {
    asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
// all blocks have passed: the last one to arrive raises `slot` = step on every rank -- one thread
// per peer, so the W release stores (an NVLink round trip each) go out side by side
__device__ __forceinline__ void dp_grid_raise(const DpPush &pp, int ticket, int slot,
                                              unsigned long long step) {
    __shared__ int last_s;
    // the block's (remote) writes are ordered before the barrier; ONE system-scope fence by the
    // thread that takes the ticket publishes them (fences are cumulative) -- 512 membar.sys per
    // block cost 5-8 us per raise (tools/dp_trace.py)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        unsigned long long *mine = pp.flags[pp.rank];
        const unsigned long long t = atomicAdd(mine + ticket, 1ull);
        last_s = t == (unsigned long long)gridDim.x - 1;
        if (last_s) mine[ticket] = 0;
    }
    __syncthreads();
    if (last_s && threadIdx.x < pp.world) {
        __threadfence_system();
        dp_st_release(pp.flags[threadIdx.x] + slot + pp.rank, step);
    }
}
// every rank has raised `slots[p] >= want`: one polling thread per peer
__device__ __forceinline__ void dp_wait_all_par(const unsigned long long *slots, int world,
                                                unsigned long long want) {
    if (threadIdx.x < world) {
        unsigned long long spins = 0;
        while (dp_ld_acquire(slots + threadIdx.x) < want)
            if (++spins > (1ull << 31)) __trap();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(512)
dp_push_kernel(float *__restrict__ grad, float *__restrict__ s0, float *__restrict__ s1, int kind,
               float lr, float momentum, float gscale, float bc1, float bc2_sqrt,
               const SegTable tab, const DpPush pp) {
    unsigned long long *mine = pp.flags[pp.rank];
    __shared__ unsigned long long step_s;
    if (threadIdx.x == 0) step_s = dp_ld_acquire(mine + DPF_STEP) + 1;
    __syncthreads();
    const unsigned long long step = step_s;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long gstride = (long long)gridDim.x * blockDim.x;
    const int W = pp.world, R = pp.rank;

    dp_stamp(step, 0);
    // phase 0: slice j of my bucket -> owner j's receive row R (one pass over the bucket: every
    // thread has stores in flight to several owners)
    for (long long i = 4 * gtid; i < pp.n; i += 4 * gstride) {
        const int j = (int)(i / pp.cap);
        if (j == R) continue;
        dp_st4(pp.recv[j] + (long long)R * pp.cap + (i - (long long)j * pp.cap),
               *reinterpret_cast<const float4 *>(grad + i));
    }
    dp_stamp(step, 1);
    dp_grid_raise(pp, DPF_TICKET_A, DPF_PUSHED, step);
    dp_stamp(step, 2);

    // phase 1: reduce + update my slice, push the new parameters to everybody
    dp_wait_all_par(mine + DPF_PUSHED, W, step);
    dp_stamp(step, 3);
    {
        const long long lo = R * pp.cap, hi = min(pp.n, lo + pp.cap);
        const float *rv = pp.recv[R];
        for (long long i = lo + 4 * gtid; i < hi; i += 4 * gstride) {
            float4 gs = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int p = 0; p < W; ++p) {         // rank order
                const float4 v = p == R ? *reinterpret_cast<const float4 *>(grad + i)
                                        : dp_ld4(rv + (long long)p * pp.cap + (i - lo));
                gs.x += v.x; gs.y += v.y; gs.z += v.z; gs.w += v.w;
            }
            const float4 w4 = *reinterpret_cast<const float4 *>(pp.param[R] + i);
            float w[4] = {w4.x, w4.y, w4.z, w4.w};
            const float g[4] = {gs.x * gscale, gs.y * gscale, gs.z * gscale, gs.w * gscale};
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (i + e < pp.n)
                    w[e] = dp_update(w[e], g[e], s0, s1, i + e, kind, lr, momentum, bc1, bc2_sqrt);
            const float4 out = make_float4(w[0], w[1], w[2], w[3]);
            for (int q = 0; q < W; ++q) dp_st4(pp.param[q] + i, out);
        }
    }
    dp_stamp(step, 4);
    dp_grid_raise(pp, DPF_TICKET_B, DPF_UPDATED, step);
    dp_stamp(step, 5);

    // phase 2: all slices are in: bf16 operand copies, gradient bucket cleared
    dp_wait_all_par(mine + DPF_UPDATED, W, step);
    dp_stamp(step, 6);
    const float *pl = pp.param[R];
    for (int sidx = 0; sidx < tab.n; ++sidx) {
        const abn_param_segment sg = tab.s[sidx];
        __nv_bfloat16 *wb = static_cast<__nv_bfloat16 *>(sg.bf16);
        if (!wb) continue;
        for (long long j = gtid; j < sg.count; j += gstride) {
            const long long r = j / sg.n_in;
            wb[r * sg.ld + (j - r * sg.n_in)] = __float2bfloat16_rn(__ldcv(pl + sg.offset + j));
        }
    }
    for (long long i = 4 * gtid; i < pp.n; i += 4 * gstride)
        *reinterpret_cast<float4 *>(grad + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    dp_stamp(step, 7);
    // the step counter advances once every block has read it
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long t = atomicAdd(mine + DPF_TICKET_C, 1ull);
        if (t == (unsigned long long)gridDim.x - 1) {
            mine[DPF_TICKET_C] = 0;
            __threadfence();
            mine[DPF_STEP] = step;
        }
    }
}

// One-shot variant: every rank pushes its WHOLE gradient bucket into every peer's receive row
// (posted NVLink writes, (W - 1) x the bucket per rank), ONE all-to-all flag exchange, then each
// rank sums the W copies in rank order and applies the optimizer to the whole bucket locally --
// bf16 weight copies and the gradient reset in the same pass, no second hop.  The exchange is
// latency bound (two-shot: two fence + flag + wait round trips, measured +24 us at 2 GPUs and
// +62 us at 8), so one hop wins although it moves W/2 x the bytes.  Receive rows are double
// buffered by step parity: a rank can only be one flag exchange ahead of its slowest peer.
__global__ void __launch_bounds__(512)
dp_push1_kernel(float *__restrict__ grad, float *__restrict__ s0, float *__restrict__ s1, int kind,
                float lr, float momentum, float gscale, float bc1, float bc2_sqrt,
                const SegTable tab, const DpPush pp) {
    unsigned long long *mine = pp.flags[pp.rank];
    __shared__ unsigned long long step_s;
    if (threadIdx.x == 0) step_s = dp_ld_acquire(mine + DPF_STEP) + 1;
    __syncthreads();
    const unsigned long long step = step_s;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long gstride = (long long)gridDim.x * blockDim.x;
    const int W = pp.world, R = pp.rank;
    const long long half = (long long)(step & 1ull) * W * pp.cap;       // this step's receive rows

    dp_stamp(step, 0);
    for (long long i = 4 * gtid; i < pp.n; i += 4 * gstride) {
        const float4 v = *reinterpret_cast<const float4 *>(grad + i);
        for (int q = 0; q < W; ++q)
            if (q != R) dp_st4(pp.recv[q] + half + (long long)R * pp.cap + i, v);
    }
    dp_stamp(step, 1);
    dp_grid_raise(pp, DPF_TICKET_A, DPF_PUSHED, step);
    dp_stamp(step, 2);
    dp_wait_all_par(mine + DPF_PUSHED, W, step);
    dp_stamp(step, 3);

    const float *rv = pp.recv[R] + half;
    float *pl = pp.param[R];
    for (int sidx = 0; sidx < tab.n; ++sidx) {
        const abn_param_segment sg = tab.s[sidx];
        __nv_bfloat16 *wb = static_cast<__nv_bfloat16 *>(sg.bf16);
        const bool vec = ((sg.offset | sg.count) & 3) == 0 && (!wb || ((sg.n_in | sg.ld) & 3) == 0);
        if (vec) {
            for (long long j = 4 * gtid; j < sg.count; j += 4 * gstride) {
                const long long i = sg.offset + j;
                float4 gs = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int p = 0; p < W; ++p) {         // rank order: the same bits on every rank
                    const float4 v = p == R ? *reinterpret_cast<const float4 *>(grad + i)
                                            : dp_ld4(rv + (long long)p * pp.cap + i);
                    gs.x += v.x; gs.y += v.y; gs.z += v.z; gs.w += v.w;
                }
                const float4 w4 = *reinterpret_cast<const float4 *>(pl + i);
                float w[4] = {w4.x, w4.y, w4.z, w4.w};
                const float g[4] = {gs.x * gscale, gs.y * gscale, gs.z * gscale, gs.w * gscale};
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    w[e] = dp_update(w[e], g[e], s0, s1, i + e, kind, lr, momentum, bc1, bc2_sqrt);
                *reinterpret_cast<float4 *>(pl + i) = make_float4(w[0], w[1], w[2], w[3]);
                *reinterpret_cast<float4 *>(grad + i) = make_float4(0.f, 0.f, 0.f, 0.f);
                if (wb) {
                    const long long r = j / sg.n_in;
                    __nv_bfloat162 lo = __floats2bfloat162_rn(w[0], w[1]), hi = __floats2bfloat162_rn(w[2], w[3]);
                    *reinterpret_cast<uint2 *>(wb + r * sg.ld + (j - r * sg.n_in)) =
                        make_uint2(*reinterpret_cast<unsigned *>(&lo), *reinterpret_cast<unsigned *>(&hi));
                }
            }
        } else {
            for (long long j = gtid; j < sg.count; j += gstride) {
                const long long i = sg.offset + j;
                float gsum = 0.f;
                for (int p = 0; p < W; ++p)
                    gsum += p == R ? grad[i] : __ldcv(rv + (long long)p * pp.cap + i);
                const float w = dp_update(pl[i], gsum * gscale, s0, s1, i, kind, lr, momentum, bc1, bc2_sqrt);
                pl[i] = w;
                grad[i] = 0.f;
                if (wb) {
                    const long long r = j / sg.n_in;
                    wb[r * sg.ld + (j - r * sg.n_in)] = __float2bfloat16_rn(w);
                }
            }
        }
    }
    // the step counter advances once every block has read it
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long t = atomicAdd(mine + DPF_TICKET_C, 1ull);
        if (t == (unsigned long long)gridDim.x - 1) {
            mine[DPF_TICKET_C] = 0;
            __threadfence();
            mine[DPF_STEP] = step;
        }
    }
}

// ------------------------------------------------------------------------
// Two-shot exchange with the FLAG INSIDE THE DATA (the "LL" protocol of collective libraries):
// every element travels as one 8-byte store {float value, uint32 step}; the receiver spins on the
// element itself until its step field is the current step.  A naturally aligned 8-byte store is
// single-copy atomic, so no system fence, no grid-wide ticket and no separate flag round trip are
// needed -- the fence + flag + poll hops of the kernels above cost ~10 us each
// (tools/dp_trace.py), which is most of what the exchange costs at this bucket size.
//   phase 0  every element outside my slice -> its owner's receive row (recvA[owner][me][k])
//   phase 1  my slice: sum the W contributions in rank order (spinning on the W - 1 remote ones),
//            optimizer update, new parameter -> my bucket and -> every peer's parameter inbox
//            (recvB[peer][i]); bf16 operand copy; gradient cleared
//   phase 2  the other slices: spin on my inbox, store the parameter, bf16 operand copy
// Blocks never wait for each other.  Flow control is implicit: a rank starts step s + 1 only after
// it has received every slice of step s, and an owner sends a slice of step s only after it has
// consumed every contribution of step s -- so a slot is never overwritten before it was read.
// Twice the bytes of the plain form (1.4 MB more per rank and step at 8 GPUs): bandwidth is not
// what bounds this exchange.
__device__ __forceinline__ unsigned long long ll_pack(float v, unsigned step) {
    return ((unsigned long long)step << 32) | (unsigned long long)__float_as_uint(v);
}
// four consecutive elements: two 16-byte stores (each 8-byte half is atomic on its own)
__device__ __forceinline__ void ll_store4(unsigned long long *p, const float (&v)[4], unsigned step) {
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(ll_pack(v[0], step)),
                 "l"(ll_pack(v[1], step)) : "memory");
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(p + 2), "l"(ll_pack(v[2], step)),
                 "l"(ll_pack(v[3], step)) : "memory");
}
__device__ __forceinline__ void ll_load4_raw(const unsigned long long *p, unsigned long long (&w)[4]) {
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(w[0]), "=l"(w[1]) : "l"(p) : "memory");
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(w[2]), "=l"(w[3]) : "l"(p + 2) : "memory");
}
__device__ __forceinline__ bool ll_ready4(const unsigned long long (&w)[4], unsigned step) {
    return (unsigned)(w[0] >> 32) == step && (unsigned)(w[1] >> 32) == step &&
           (unsigned)(w[2] >> 32) == step && (unsigned)(w[3] >> 32) == step;
}
// spin until all four elements carry this step's flag (a missing peer must trap, not hang forever)
__device__ __forceinline__ void ll_wait4(const unsigned long long *p, unsigned long long (&w)[4], unsigned step) {
    unsigned long long spins = 0;
    while (!ll_ready4(w, step)) {
        if (++spins > (1ull << 28)) __trap();
        ll_load4_raw(p, w);
    }
}
// bf16 operand copies of parameters i .. i+3 (segments start and end on multiples of 4 or the
// elements are handled one by one)
__device__ __forceinline__ void ll_bf16_4(const SegTable &tab, long long i, const float (&w)[4]) {
    for (int sidx = 0; sidx < tab.n; ++sidx) {
        const abn_param_segment sg = tab.s[sidx];
        if (i + 3 < sg.offset || i >= sg.offset + sg.count || !sg.bf16) continue;
        __nv_bfloat16 *wb = static_cast<__nv_bfloat16 *>(sg.bf16);
        if (i >= sg.offset && i + 4 <= sg.offset + sg.count && !((sg.n_in | sg.ld | (i - sg.offset)) & 3)) {
            const long long j = i - sg.offset, r = j / sg.n_in;
            __nv_bfloat162 lo2 = __floats2bfloat162_rn(w[0], w[1]), hi2 = __floats2bfloat162_rn(w[2], w[3]);
            *reinterpret_cast<uint2 *>(wb + r * sg.ld + (j - r * sg.n_in)) =
                make_uint2(*reinterpret_cast<unsigned *>(&lo2), *reinterpret_cast<unsigned *>(&hi2));
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const long long j = i + e - sg.offset;
                if (j >= 0 && j < sg.count) {
                    const long long r = j / sg.n_in;
                    wb[r * sg.ld + (j - r * sg.n_in)] = __float2bfloat16_rn(w[e]);
                }
            }
        }
    }
}

__global__ void __launch_bounds__(512)
dp_ll_kernel(float *__restrict__ grad, float *__restrict__ s0, float *__restrict__ s1, int kind,
             float lr, float momentum, float gscale, float bc1, float bc2_sqrt,
             const SegTable tab, const DpPush pp) {
    unsigned long long *mine = pp.flags[pp.rank];
    __shared__ unsigned long long step_s;
    if (threadIdx.x == 0) step_s = dp_ld_acquire(mine + DPF_STEP) + 1;
    __syncthreads();
    const unsigned step = (unsigned)step_s;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long gstride = (long long)gridDim.x * blockDim.x;
    const int W = pp.world, R = pp.rank;
    const long long lo = (long long)R * pp.cap, hi = min(pp.n, lo + pp.cap);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    dp_stamp(step_s, 0);
    // phase 0: four consecutive elements per thread (n, cap and lo are multiples of 4)
    for (long long i = 4 * gtid; i < pp.n; i += 4 * gstride) {
        if (i >= lo && i < hi) continue;
        const int j = (int)(i / pp.cap);
        unsigned long long *dst = reinterpret_cast<unsigned long long *>(pp.recv[j]) +
                                  (long long)R * pp.cap + (i - (long long)j * pp.cap);
        const float4 g4 = *reinterpret_cast<const float4 *>(grad + i);
        const float v[4] = {g4.x, g4.y, g4.z, g4.w};
        ll_store4(dst, v, step);
        *reinterpret_cast<float4 *>(grad + i) = zero4;
    }
    dp_stamp(step_s, 1);
    // phase 1: every remote contribution of the four elements is requested before the first is checked
    {
        const unsigned long long *rv = reinterpret_cast<const unsigned long long *>(pp.recv[R]);
        float *pl = pp.param[R];
        for (long long i = lo + 4 * gtid; i < hi; i += 4 * gstride) {
            unsigned long long part[DP_MAX_WORLD][4];
#pragma unroll
            for (int p = 0; p < DP_MAX_WORLD; ++p)
                if (p < W && p != R) ll_load4_raw(rv + (long long)p * pp.cap + (i - lo), part[p]);
            const float4 own = *reinterpret_cast<const float4 *>(grad + i);
            *reinterpret_cast<float4 *>(grad + i) = zero4;
            float gs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int p = 0; p < DP_MAX_WORLD; ++p) {          // rank order
                if (p >= W) continue;
                if (p == R) {
                    gs[0] += own.x; gs[1] += own.y; gs[2] += own.z; gs[3] += own.w;
                } else {
                    ll_wait4(rv + (long long)p * pp.cap + (i - lo), part[p], step);
#pragma unroll
                    for (int e = 0; e < 4; ++e) gs[e] += __uint_as_float((unsigned)part[p][e]);
                }
            }
            const float4 w4 = *reinterpret_cast<const float4 *>(pl + i);
            float w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (i + e < pp.n)
                    w[e] = dp_update(w[e], gs[e] * gscale, s0, s1, i + e, kind, lr, momentum, bc1, bc2_sqrt);
            *reinterpret_cast<float4 *>(pl + i) = make_float4(w[0], w[1], w[2], w[3]);
            for (int q = 0; q < W; ++q)
                if (q != R)
                    ll_store4(reinterpret_cast<unsigned long long *>(pp.recv[q]) + (long long)W * pp.cap + i,
                              w, step);
            ll_bf16_4(tab, i, w);
        }
    }
    dp_stamp(step_s, 2);
    // phase 2: the other owners' slices from my parameter inbox
    {
        const unsigned long long *inbox = reinterpret_cast<const unsigned long long *>(pp.recv[R]) +
                                          (long long)W * pp.cap;
        float *pl = pp.param[R];
        for (long long i = 4 * gtid; i < pp.n; i += 4 * gstride) {
            if (i >= lo && i < hi) continue;
            unsigned long long raw[4];
            ll_load4_raw(inbox + i, raw);
            ll_wait4(inbox + i, raw, step);
            const float w[4] = {__uint_as_float((unsigned)raw[0]), __uint_as_float((unsigned)raw[1]),
                                __uint_as_float((unsigned)raw[2]), __uint_as_float((unsigned)raw[3])};
            *reinterpret_cast<float4 *>(pl + i) = make_float4(w[0], w[1], w[2], w[3]);
            ll_bf16_4(tab, i, w);
        }
    }
    dp_stamp(step_s, 3);
    // the step counter advances once every block has read it
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long t = atomicAdd(mine + DPF_TICKET_C, 1ull);
        if (t == (unsigned long long)gridDim.x - 1) {
            mine[DPF_TICKET_C] = 0;
            __threadfence();
            mine[DPF_STEP] = step_s;
        }
    }
}

}  // namespace abn

extern "C" int abn_dp_set_trace(long long *buffer) {       // debug hook: device [64][8] int64, or NULL
    if (int rc = require_sm100()) return rc;
    cudaError_t e = cudaMemcpyToSymbol(abn::g_dp_trace, &buffer, sizeof(buffer));
    return e == cudaSuccess ? ABN_OK : set_error(ABN_EIO, "abn_dp_set_trace: %s", cudaGetErrorString(e));
}

extern "C" int abn_dp_push_step(float *grad, float *state0, float *state1, int kind, float lr,
                                float momentum, float grad_scale, int64_t step,
                                const abn_param_segment *segments, int n_segments,
                                const abn_dp_push *peers, abn_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    if (!grad || !segments || !peers || n_segments <= 0 || n_segments > ABN_MAX_PARAM_SEGMENTS ||
        kind < 0 || kind > 2 || (kind == 0 && momentum != 0.f && !state0) ||
        (kind >= 1 && (!state0 || !state1)) || step < 1)
        return set_error(ABN_EINVAL, "abn_dp_push_step: bad argument");
    if (peers->world < 2 || peers->world > DP_MAX_WORLD || peers->rank < 0 ||
        peers->rank >= peers->world || peers->n <= 0 || (peers->n & 3) || (peers->slice_cap & 3) ||
        peers->slice_cap * peers->world < peers->n || (peers->one_shot == 1 && peers->slice_cap < peers->n))
        return set_error(ABN_EINVAL, "abn_dp_push_step: world 2..%d, n and slice_cap multiples of 4, "
                         "world * slice_cap >= n", DP_MAX_WORLD);
    DpPush pp;
    pp.rank = peers->rank; pp.world = peers->world; pp.n = peers->n; pp.cap = peers->slice_cap;
    pp.one_shot = peers->one_shot;
    for (int r = 0; r < peers->world; ++r) {
        if (!peers->param[r] || !peers->recv[r] || !peers->flags[r])
            return set_error(ABN_EINVAL, "abn_dp_push_step: rank %d is not mapped", r);
        pp.param[r] = static_cast<float *>(peers->param[r]);
        pp.recv[r] = static_cast<float *>(peers->recv[r]);
        pp.flags[r] = static_cast<unsigned long long *>(peers->flags[r]);
    }
    SegTable tab;
    tab.n = n_segments;
    for (int i = 0; i < n_segments; ++i) tab.s[i] = segments[i];
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const float bc1 = 1.f - powf(0.9f, (float)step);
    const float bc2s = sqrtf(1.f - powf(0.999f, (float)step));
    if (pp.one_shot == 2)
        dp_ll_kernel<<<sms, 512, 0, (cudaStream_t)stream>>>(grad, state0, state1, kind, lr, momentum,
                                                            grad_scale, bc1, bc2s, tab, pp);
    else if (pp.one_shot)
        dp_push1_kernel<<<sms, 512, 0, (cudaStream_t)stream>>>(grad, state0, state1, kind, lr, momentum,
                                                               grad_scale, bc1, bc2s, tab, pp);
    else
        dp_push_kernel<<<sms, 512, 0, (cudaStream_t)stream>>>(grad, state0, state1, kind, lr, momentum,
                                                              grad_scale, bc1, bc2s, tab, pp);
    return check_launch("abn_dp_push_step");
}
