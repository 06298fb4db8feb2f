// librpst: version, error plumbing, tuning registry, device attributes.
#include "common.cuh"

#include <string.h>

namespace rpst {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return RPST_ERR_CUDA;
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

int set_adain_tuning(const char* name, int64_t v, bool set, int64_t* out);

}  // namespace rpst

extern "C" int rpst_version(void) { return RPST_VERSION; }

extern "C" const char* rpst_last_error(void) { return rpst::g_err; }

extern "C" int rpst_set_tuning(const char* name, int64_t value) {
    if (name && rpst::set_adain_tuning(name, value, true, nullptr)) return RPST_OK;
    rpst::set_error("unknown tuning knob '%s'", name ? name : "(null)");
    return RPST_ERR_INVALID;
}

extern "C" int64_t rpst_get_tuning(const char* name) {
    int64_t v = -1;
    if (name && rpst::set_adain_tuning(name, 0, false, &v)) return v;
    return -1;
}
