// abn_tc.cu -- operand preparation for the tensor-core path (abn_tc2.cu): fp32 -> bf16
// copies of weight matrices (after load_state_dict) and of fp32 input batches.
#include <cuda_bf16.h>

#include "abn_common.cuh"

namespace abn {

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

}  // namespace abn

using namespace abn;

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

