// abn_capi.cu -- library-wide pieces of the C ABI: version, error text, the
// sm_100 gate, and the precision dispatch of the embedder layers.
#include "abn_common.cuh"

namespace abn {

char *err_buf() {
    static thread_local char buf[512] = "";
    return buf;
}

int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

// No fallback by design: anything but an sm_100 device is refused loudly.
int require_sm100() {
    static thread_local int cached = -1;
    if (cached == 0) return ABN_OK;
    int dev = 0, major = 0, minor = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) {
        cudaGetLastError();
        return set_error(ABN_ENOSYS, "no usable CUDA device: this library has no CPU path");
    }
    if (major != 10)
        return set_error(ABN_ENOSYS, "device is sm_%d%d; this library is built for sm_100a only",
                         major, minor);
    cached = 0;
    return ABN_OK;
}

int simt_linear_forward(const float *x, const float *W, const float *b, int64_t m, int n_in,
                        int n_out, int act, float *y, const abn_dropout *drop, int64_t row_offset,
                        cudaStream_t st);
int simt_linear_backward(const float *x, const float *W, const float *y, float *dy, int64_t m,
                         int n_in, int n_out, int act, int accumulate, float *dx, float *dW,
                         float *db, const abn_dropout *drop, int64_t row_offset, cudaStream_t st);

}  // namespace abn

using namespace abn;

extern "C" int abn_version(void) { return 100; /* 0.1.0 */ }

extern "C" const char *abn_last_error(void) { return err_buf(); }

extern "C" int abn_device_info(int *sm_count, int *cc_major, int *cc_minor,
                               size_t *smem_optin_bytes) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();
        return set_error(ABN_ENOSYS, "no usable CUDA device");
    }
    if (sm_count) { cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev); *sm_count = v; }
    if (cc_major) { cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev); *cc_major = v; }
    if (cc_minor) { cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev); *cc_minor = v; }
    if (smem_optin_bytes) {
        cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        *smem_optin_bytes = (size_t)v;
    }
    return ABN_OK;
}

extern "C" int abn_linear_forward_drop(const float *x, const float *W, const float *b, int64_t m,
                                       int n_in, int n_out, int act, int precision, float *y,
                                       const abn_dropout *drop, int64_t row_offset,
                                       abn_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    if (m == 0) return ABN_OK;
    if (!x || !W || !y || m < 0 || m > 0x7fffffff || n_in <= 0 || n_out <= 0 || act < 0 || act > 3 ||
        (drop && (drop->p < 0.f || drop->p >= 1.f)))
        return set_error(ABN_EINVAL, "abn_linear_forward: bad argument");
    if (precision == 0)
        return simt_linear_forward(x, W, b, m, n_in, n_out, act, y, drop, row_offset,
                                   (cudaStream_t)stream);
    return set_error(ABN_EINVAL, "abn_linear_forward: unknown precision %d", precision);
}

extern "C" int abn_linear_forward(const float *x, const float *W, const float *b, int64_t m,
                                  int n_in, int n_out, int act, int precision, float *y,
                                  abn_stream_t stream) {
    return abn_linear_forward_drop(x, W, b, m, n_in, n_out, act, precision, y, nullptr, 0, stream);
}

extern "C" int abn_linear_backward_drop(const float *x, const float *W, const float *y, float *dy,
                                        int64_t m, int n_in, int n_out, int act, int precision,
                                        int accumulate, float *dx, float *dW, float *db,
                                        const abn_dropout *drop, int64_t row_offset,
                                        abn_stream_t stream) {
    if (int rc = require_sm100()) return rc;
    if (m == 0) return ABN_OK;
    if (!x || !W || !y || !dy || m < 0 || m > 0x7fffffff || n_in <= 0 || n_out <= 0 || act < 0 ||
        act > 3 || (drop && (drop->p < 0.f || drop->p >= 1.f)))
        return set_error(ABN_EINVAL, "abn_linear_backward: bad argument");
    if (precision == 0)
        return simt_linear_backward(x, W, y, dy, m, n_in, n_out, act, accumulate, dx, dW, db, drop,
                                    row_offset, (cudaStream_t)stream);
    return set_error(ABN_EINVAL, "abn_linear_backward: unknown precision %d", precision);
}

extern "C" int abn_linear_backward(const float *x, const float *W, const float *y, float *dy,
                                   int64_t m, int n_in, int n_out, int act, int precision,
                                   int accumulate, float *dx, float *dW, float *db,
                                   abn_stream_t stream) {
    return abn_linear_backward_drop(x, W, y, dy, m, n_in, n_out, act, precision, accumulate, dx, dW,
                                    db, nullptr, 0, stream);
}
