// Error reporting, device queries and ABI versioning for libamp_b200.so.
#include <cstdarg>
#include <cstdio>

#include "amp_internal.h"

namespace amp {

static thread_local char g_error[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char *what) {
    snprintf(g_error, sizeof(g_error), "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    // clear the sticky-free error state so the next call reports its own failure
    (void)cudaGetLastError();
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return AMP_ENODEV;
    if (e == cudaErrorMemoryAllocation) return AMP_ENOMEM;
    return AMP_ECUDA;
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace amp

extern "C" {

int amp_b200_abi_version(void) { return AMP_B200_ABI_VERSION; }

const char *amp_last_error(void) { return amp::g_error; }

int amp_set_device(int device) {
    AMP_CUDA_TRY(cudaSetDevice(device));
    return AMP_OK;
}

int amp_device_info(int *sms, int *cc_major, int *cc_minor) {
    int dev = 0;
    AMP_CUDA_TRY(cudaGetDevice(&dev));
    int v = 0;
    if (sms) {
        AMP_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
        *sms = v;
    }
    if (cc_major) {
        AMP_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev));
        *cc_major = v;
    }
    if (cc_minor) {
        AMP_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev));
        *cc_minor = v;
    }
    return AMP_OK;
}

}  // extern "C"
