// abn_common.cuh -- shared helpers for the sm_100a kernels and the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <math.h>

#include "../../include/abnet3_b200.h"

namespace abn {

// thread-local last-error text, returned by abn_last_error()
char *err_buf();
int set_error(int code, const char *fmt, ...);

inline int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        return set_error(ABN_EIO, "%s: %s", what, cudaGetErrorString(e));
    return ABN_OK;
}

// The library is sm_100a-only by design: no other code path exists.
int require_sm100();

// direction codes (shared with oracle/dtw_oracle.c)
enum : uint8_t { DIR_DIAG = 0, DIR_UP = 1, DIR_LEFT = 2 };

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() {
    asm volatile("cp.async.commit_group;\n" ::);
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace abn
