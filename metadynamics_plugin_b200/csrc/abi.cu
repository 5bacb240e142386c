// abi.cu -- library-level pieces of the C ABI: version, last-error text, device query.
#include "common.cuh"

namespace metad {
static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
    g_last_error = buf;
    return METAD_ERR_CUDA;
}

int device_sm_count() {
    static int cached = 0;
    if (!cached) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            return 148;  // B200
    }
    return cached;
}
}  // namespace metad

extern "C" int metad_version(void) { return 100; }
extern "C" const char* metad_last_error(void) { return metad::g_last_error.c_str(); }
