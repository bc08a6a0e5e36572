// abi.cu -- error plumbing and device checks shared by every entry point of libhmm_b200.so.
#include "common.cuh"

#include <string.h>

namespace hmmb200 {

static thread_local char g_last_error[512] = "";

char *last_error_buf() { return g_last_error; }

int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(HMMB200_ELAUNCH, "%s: %s", what, cudaGetErrorString(e));
    return HMMB200_OK;
}

int require_sm100() {
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_error(HMMB200_ENODEVICE, "no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e));
    }
    // per-device cache; benign race (idempotent writes)
    static int cached[64];
    if (dev >= 0 && dev < 64 && cached[dev] != 0) return cached[dev] > 0 ? HMMB200_OK
        : set_error(HMMB200_ENODEVICE, "device %d is not compute capability 10.x", dev);
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_error(HMMB200_ENODEVICE, "cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
    }
    if (dev >= 0 && dev < 64) cached[dev] = (major == 10) ? 1 : -1;
    if (major != 10) return set_error(HMMB200_ENODEVICE, "device %d is sm_%d0, this library is built for sm_100a only", dev, major);
    return HMMB200_OK;
}

}  // namespace hmmb200

using namespace hmmb200;

HMMB200_EXPORT int hmmb200_abi_version(void) { return HMMB200_ABI_VERSION; }

HMMB200_EXPORT const char *hmmb200_last_error(void) { return last_error_buf(); }

HMMB200_EXPORT int hmmb200_device_check(int ordinal) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return set_error(HMMB200_ENODEVICE, "no CUDA device visible (this library has no CPU fallback)");
    }
    int dev = ordinal;
    if (dev < 0) {
        e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return set_error(HMMB200_ENODEVICE, "cudaGetDevice: %s", cudaGetErrorString(e));
    }
    if (dev >= count) return set_error(HMMB200_EINVAL, "device ordinal %d out of range (%d devices)", dev, count);
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) return set_error(HMMB200_ENODEVICE, "cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
    if (major != 10) return set_error(HMMB200_ENODEVICE, "device %d is sm_%d0; sm_100a required", dev, major);
    return HMMB200_OK;
}
