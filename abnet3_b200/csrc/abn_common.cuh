// abn_common.cuh -- shared helpers for the sm_100a kernels and the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
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

// Programmatic dependent launch for the small kernels of the training step: the kernel may be
// scheduled while its predecessor drains; pdl_wait() (first statement of the kernel) blocks
// until the predecessor's memory is visible, and lets the successor start its own prologue.
__device__ __forceinline__ void pdl_wait() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    static int on = -1;                 // ABN_NO_PDL=1: plain stream order
    if (on < 0) { const char *e = getenv("ABN_NO_PDL"); on = (e && e[0] == '1') ? 0 : 1; }
    cfg.numAttrs = on ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace abn
