// abn_nn.cu -- kernel (4) fused pair loss + gradient, the fp32 SIMT parity path
// of kernel (3) (embedder MLP layers), and the fused optimizer step.
//
// Reference behaviour reproduced (paths relative to /root/reference):
//   coscos2 / cosmargin      abnet3/loss.py:46-67, :85-105 (+ autograd backward)
//   Linear -> act blocks     abnet3/model.py:133-170, forward :179-186
//   optimizer.step()         abnet3/trainer.py:68-87, :240 (torch.optim semantics)
#include "abn_common.cuh"
#include "abn_drop.cuh"

namespace abn {

// ------------------------------------------------------------- loss (4) ----
// One warp per frame pair: dot, |a|^2, |b|^2 by shuffle reduction, then the
// loss term and both gradients in the same pass (rows stay in registers for
// dim <= 128; longer rows are re-read from L1/L2).
//   c = sum_k (a_k / max(|a|,eps)) (b_k / max(|b|,eps))          (torch >= 1.12)
//   dc/da = b / (an bn) - [|a| > eps] c a / an^2     (an = max(|a|, eps))
constexpr int LOSS_WARPS = 8;
constexpr float COS_EPS = 1e-6f;

__global__ void __launch_bounds__(LOSS_WARPS * 32)
pair_loss_kernel(const float *__restrict__ e1, const float *__restrict__ e2,
                 const float *__restrict__ y, int64_t n, int dim, int64_t ld, int kind,
                 float margin, float scale, float *__restrict__ loss, float *__restrict__ de1,
                 float *__restrict__ de2) {
    __shared__ float wsum[LOSS_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float local = 0.f;
    for (int64_t row = (int64_t)blockIdx.x * LOSS_WARPS + warp; row < n;
         row += (int64_t)gridDim.x * LOSS_WARPS) {
        const float *a = e1 + row * ld, *b = e2 + row * ld;
        float av[4], bv[4];
        float dot = 0.f, na = 0.f, nb = 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = lane + 32 * u;
            av[u] = k < dim ? a[k] : 0.f;
            bv[u] = k < dim ? b[k] : 0.f;
            dot = fmaf(av[u], bv[u], dot);
            na = fmaf(av[u], av[u], na);
            nb = fmaf(bv[u], bv[u], nb);
        }
        for (int k = lane + 128; k < dim; k += 32) {
            const float x = a[k], z = b[k];
            dot = fmaf(x, z, dot); na = fmaf(x, x, na); nb = fmaf(z, z, nb);
        }
        dot = warp_sum(dot); na = warp_sum(na); nb = warp_sum(nb);
        const float ra = sqrtf(na), rb = sqrtf(nb);
        const float an = fmaxf(ra, COS_EPS), bn = fmaxf(rb, COS_EPS);
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
        if (de1) {
            const float g = dldc * scale;
            const float ka = ra > COS_EPS ? c / (an * an) : 0.f;
            const float kb = rb > COS_EPS ? c / (bn * bn) : 0.f;
            float *ga = de1 + row * ld, *gb = de2 + row * ld;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int k = lane + 32 * u;
                if (k < dim) {
                    ga[k] = g * (bv[u] * inv - ka * av[u]);
                    gb[k] = g * (av[u] * inv - kb * bv[u]);
                }
            }
            for (int k = lane + 128; k < dim; k += 32) {
                const float x = a[k], z = b[k];
                ga[k] = g * (z * inv - ka * x);
                gb[k] = g * (x * inv - kb * z);
            }
        }
    }
    if (lane == 0) wsum[warp] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < LOSS_WARPS; ++w) s += wsum[w];
        if (s != 0.f) atomicAdd(loss, s * scale);
    }
}

// ------------------------------------------------ SIMT fp32 GEMM (parity) ---
// C[M,N] = sum_k A(m,k) B(k,n), with A and B each either k-contiguous or
// m/n-contiguous, 128x64x16 tiles, 256 threads, 8x4 register tile per thread,
// k order fixed and sequential so results are deterministic.
//   AK: A stored [M][K] (k contiguous)   else [K][M]
//   BK: B stored [N][K] (k contiguous)   else [K][N]
constexpr int GM = 128, GN = 64, GK = 16, GT = 256;

enum { EPI_BIAS_ACT = 0, EPI_STORE = 1, EPI_ACCUM = 2, EPI_ADD = 3 };

__device__ __forceinline__ float act_fwd(float v, int act) {
    switch (act) {
        case 1: return 1.f / (1.f + expf(-v));
        case 2: return tanhf(v);
        case 3: return v > 0.f ? v : 0.f;
        default: return v;
    }
}


// 4 consecutive floats starting at p, of which the first `nvalid` exist;
// vectorised when all four exist and p is 16-byte aligned.
__device__ __forceinline__ float4 load4_guard(const float *p, int nvalid) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (nvalid >= 4 && ((reinterpret_cast<uintptr_t>(p) & 15u) == 0)) {
        v = *reinterpret_cast<const float4 *>(p);
    } else {
        if (nvalid > 0) v.x = p[0];
        if (nvalid > 1) v.y = p[1];
        if (nvalid > 2) v.z = p[2];
        if (nvalid > 3) v.w = p[3];
    }
    return v;
}

template <bool AK, bool BK>
__global__ void __launch_bounds__(GT)
sgemm_kernel(const float *__restrict__ A, const float *__restrict__ B, float *__restrict__ C,
             int M, int N, int K, int lda, int ldb, int ldc, int epi,
             const float *__restrict__ bias, int act, int ksplit_len,
             const DropArgs drop = DropArgs{nullptr, 0u, 1.f, 0}, long long drop_row0 = 0) {
    __shared__ float As[2][GK][GM + 4];
    __shared__ float Bs[2][GK][GN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
    const int kbeg = blockIdx.z * ksplit_len;
    const int kend = min(K, kbeg + ksplit_len);
    const int tx = tid & 15, ty = tid >> 4;      // 16 x 16 threads; thread tile 8 (m) x 4 (n)

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    auto load_tiles = [&](int buf, int k0) {
        // A tile: GM x GK = 2048 floats, 8 per thread
        if (AK) {
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                const int e = tid + it * GT;          // 512 float4: row = e / 4, kq = e % 4
                const int r = e >> 2, kq = (e & 3) * 4;
                const int gm = m0 + r, gk = k0 + kq;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (gm < M) v = load4_guard(A + (size_t)gm * lda + gk, kend - gk);
                As[buf][kq][r] = v.x; As[buf][kq + 1][r] = v.y;
                As[buf][kq + 2][r] = v.z; As[buf][kq + 3][r] = v.w;
            }
        } else {
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                const int e = tid + it * GT;          // row k = e / 32, mq = e % 32
                const int kk = e >> 5, mq = (e & 31) * 4;
                const int gk = k0 + kk, gm = m0 + mq;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (gk < kend) v = load4_guard(A + (size_t)gk * lda + gm, M - gm);
                *reinterpret_cast<float4 *>(&As[buf][kk][mq]) = v;
            }
        }
        // B tile: GN x GK = 1024 floats, 4 per thread
        if (BK) {
            const int r = tid >> 2, kq = (tid & 3) * 4;
            const int gn = n0 + r, gk = k0 + kq;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gn < N) v = load4_guard(B + (size_t)gn * ldb + gk, kend - gk);
            Bs[buf][kq][r] = v.x; Bs[buf][kq + 1][r] = v.y;
            Bs[buf][kq + 2][r] = v.z; Bs[buf][kq + 3][r] = v.w;
        } else {
            const int kk = tid >> 4, nq = (tid & 15) * 4;
            const int gk = k0 + kk, gn = n0 + nq;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gk < kend) v = load4_guard(B + (size_t)gk * ldb + gn, N - gn);
            *reinterpret_cast<float4 *>(&Bs[buf][kk][nq]) = v;
        }
    };

    const int nk = (kend - kbeg + GK - 1) / GK;
    if (nk > 0) load_tiles(0, kbeg);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) load_tiles(buf ^ 1, kbeg + (kt + 1) * GK);
#pragma unroll
        for (int kk = 0; kk < GK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[buf][kk][tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }

    const bool dropping = epi == EPI_BIAS_ACT && drop.state != nullptr;
    const unsigned long long dkey = dropping ? drop_key(drop) : 0ull;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (gm >= M) continue;
        // this thread's four columns n0 + 4 tx .. + 3 share one draw (abn_drop.cuh)
        const unsigned long long dbits = dropping ? drop_bits4(dkey, drop_row0 + gm, (n0 >> 2) + tx) : 0ull;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + tx * 4 + j;
            if (gn >= N) continue;
            float v = acc[i][j];
            float *dst = C + (size_t)gm * ldc + gn;
            if (epi == EPI_BIAS_ACT) {
                float z = v + (bias ? bias[gn] : 0.f);
                if (dropping) z = drop_keep_of(dbits, j, drop.thresh) ? z * drop.inv_keep : 0.f;
                *dst = act_fwd(z, act);
            }
            else if (epi == EPI_STORE) *dst = v;
            else if (epi == EPI_ADD) *dst += v;
            else atomicAdd(dst, v);
        }
    }
}

// dz = dy * act'(y) in place, and db (+)= column sums of dz.
__global__ void act_backward_kernel(const float *__restrict__ y, float *__restrict__ dy, int64_t m,
                                    int n, int act, float *__restrict__ db, int rows_per_block,
                                    const DropArgs drop, long long drop_row0) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= n) return;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
    const int64_t r1 = min(m, r0 + rows_per_block);
    const unsigned long long dkey = drop.state ? drop_key(drop) : 0ull;
    float s = 0.f;
    for (int64_t r = r0; r < r1; ++r) {
        const size_t o = (size_t)r * n + col;
        const float yy = y[o];
        float g = dy[o];
        switch (act) {
            case 1: g *= yy * (1.f - yy); break;
            case 2: g *= 1.f - yy * yy; break;
            case 3: g = yy > 0.f ? g : 0.f; break;
            default: break;
        }
        if (drop.state) g = drop_keep(dkey, drop_row0 + r, col, drop.thresh) ? g * drop.inv_keep : 0.f;
        dy[o] = g;
        s += g;
    }
    if (db) atomicAdd(db + col, s);
}

// --------------------------------------------------------------- optimizer --
__global__ void optimizer_kernel(float *__restrict__ p, const float *__restrict__ g,
                                 float *__restrict__ s0, float *__restrict__ s1, int64_t n,
                                 int kind, float lr, float momentum, float gscale, float bc1,
                                 float bc2_sqrt) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
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
    }
}

}  // namespace abn

using namespace abn;

extern "C" int abn_pair_loss(const float *e1, const float *e2, const float *y, int64_t n, int dim,
                             int64_t ld, int kind, float margin, float scale, float *loss,
                             float *de1, float *de2, abn_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    if (n == 0) return ABN_OK;
    if (ld == 0) ld = dim;
    if (!e1 || !e2 || !y || !loss || n < 0 || dim <= 0 || ld < dim || (kind != 0 && kind != 1) ||
        ((de1 == nullptr) != (de2 == nullptr)))
        return set_error(ABN_EINVAL, "abn_pair_loss: bad argument");
    int64_t blocks = (n + LOSS_WARPS - 1) / LOSS_WARPS;
    if (blocks > 148 * 8) blocks = 148 * 8;
    pair_loss_kernel<<<(unsigned)blocks, LOSS_WARPS * 32, 0, (cudaStream_t)stream>>>(
        e1, e2, y, n, dim, ld, kind, margin, scale, loss, de1, de2);
    return check_launch("abn_pair_loss");
}

namespace abn {
// y = act(x W^T + b) on the fp32 SIMT path
int simt_linear_forward(const float *x, const float *W, const float *b, int64_t m, int n_in,
                        int n_out, int act, float *y, const abn_dropout *drop, int64_t row_offset,
                        cudaStream_t st) {
    dim3 grid((n_out + GN - 1) / GN, (unsigned)((m + GM - 1) / GM), 1);
    sgemm_kernel<true, true><<<grid, GT, 0, st>>>(x, W, y, (int)m, n_out, n_in, n_in, n_in, n_out,
                                                  EPI_BIAS_ACT, b, act, n_in, drop_args(drop),
                                                  (long long)row_offset);
    return check_launch("abn_linear_forward(simt)");
}

int launch_act_backward(const float *y, float *dy, int64_t m, int n_out, int act, float *db,
                        const abn_dropout *drop, int64_t row_offset, cudaStream_t st) {
    const int rows_per_block = 128;
    dim3 grid((n_out + 127) / 128, (unsigned)((m + rows_per_block - 1) / rows_per_block));
    act_backward_kernel<<<grid, 128, 0, st>>>(y, dy, m, n_out, act, db, rows_per_block,
                                              drop_args(drop), (long long)row_offset);
    return check_launch("abn_linear_backward(act)");
}

// the keep mask of one layer, 1 byte per element (tests: mask replay in the oracle)
__global__ void dropout_mask_kernel(const DropArgs drop, int64_t rows, int cols, uint8_t *__restrict__ mask) {
    const unsigned long long key = drop_key(drop);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < rows * cols;
         e += (int64_t)gridDim.x * blockDim.x)
        mask[e] = drop_keep(key, e / cols, (int)(e % cols), drop.thresh) ? 1 : 0;
}

int simt_linear_backward(const float *x, const float *W, const float *y, float *dy, int64_t m,
                         int n_in, int n_out, int act, int accumulate, float *dx, float *dW,
                         float *db, const abn_dropout *drop, int64_t row_offset, cudaStream_t st) {
    const int acc_dx = accumulate & 2;
    accumulate &= 1;
    if (!accumulate) {
        if (db) cudaMemsetAsync(db, 0, sizeof(float) * n_out, st);
        if (dW) cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)n_out * n_in, st);
    }
    if (int rc = launch_act_backward(y, dy, m, n_out, act, db, drop, row_offset, st)) return rc;
    if (dx) {   // dx[m, n_in] = dz[m, n_out] @ W[n_out, n_in]
        dim3 grid((n_in + GN - 1) / GN, (unsigned)((m + GM - 1) / GM), 1);
        sgemm_kernel<true, false><<<grid, GT, 0, st>>>(dy, W, dx, (int)m, n_in, n_out, n_out, n_in,
                                                       n_in, acc_dx ? EPI_ADD : EPI_STORE, nullptr,
                                                       0, n_out);
        if (int rc = check_launch("abn_linear_backward(dgrad)")) return rc;
    }
    if (dW) {   // dW[n_out, n_in] += dz^T[n_out, m] @ x[m, n_in], split over m
        int tiles = ((n_out + GM - 1) / GM) * ((n_in + GN - 1) / GN);
        int splits = (2 * 148 + tiles - 1) / tiles;
        int64_t len = (m + splits - 1) / splits;
        len = (len + GK - 1) / GK * GK;
        if (len < GK) len = GK;
        splits = (int)((m + len - 1) / len);
        dim3 grid((n_in + GN - 1) / GN, (n_out + GM - 1) / GM, splits);
        sgemm_kernel<false, false><<<grid, GT, 0, st>>>(dy, x, dW, n_out, n_in, (int)m, n_out,
                                                        n_in, n_in, EPI_ACCUM, nullptr, 0,
                                                        (int)len);
        if (int rc = check_launch("abn_linear_backward(wgrad)")) return rc;
    }
    return ABN_OK;
}
}  // namespace abn

extern "C" int abn_optimizer_step(float *param, const float *grad, float *state0, float *state1,
                                  int64_t n, int kind, float lr, float momentum, float grad_scale,
                                  int64_t step, abn_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    if (n == 0) return ABN_OK;
    if (!param || !grad || n < 0 || kind < 0 || kind > 2 ||
        (kind == 0 && momentum != 0.f && !state0) || (kind >= 1 && (!state0 || !state1)) ||
        step < 1)
        return set_error(ABN_EINVAL, "abn_optimizer_step: bad argument");
    const float bc1 = 1.f - powf(0.9f, (float)step);
    const float bc2s = sqrtf(1.f - powf(0.999f, (float)step));
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    optimizer_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        param, grad, state0, state1, n, kind, lr, momentum, grad_scale, bc1, bc2s);
    return check_launch("abn_optimizer_step");
}

extern "C" int abn_dropout_mask(const abn_dropout *drop, int64_t rows, int cols, uint8_t *mask,
                                abn_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    if (!drop || !drop->state || !mask || rows < 0 || cols <= 0 || drop->p <= 0.f || drop->p >= 1.f)
        return set_error(ABN_EINVAL, "abn_dropout_mask: bad argument (0 < p < 1, state on the device)");
    if (rows == 0) return ABN_OK;
    int64_t blocks = (rows * cols + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    abn::dropout_mask_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(abn::drop_args(drop), rows,
                                                                                 cols, mask);
    return check_launch("abn_dropout_mask");
}
