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
int set_watchdog_adain(unsigned long long ns);
int set_watchdog_cov(unsigned long long ns);
int set_watchdog_flash(unsigned long long ns);
int set_watchdog_gemm(unsigned long long ns);
int set_watchdog_sanet(unsigned long long ns);
int set_watchdog_seg(unsigned long long ns);
int set_watchdog_pwconv(unsigned long long ns);

static int64_t g_watchdog_ms = 4000;

// the limit lives in one device variable per translation unit and per device: write it everywhere
static int apply_watchdog(int64_t ms) {
    const unsigned long long ns = ms <= 0 ? 0ull : (unsigned long long)ms * 1000000ull;
    int count = 0, cur = 0;
    g_watchdog_ms = ms <= 0 ? 0 : ms;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return RPST_OK;   // no device: nothing to configure
    cudaGetDevice(&cur);
    int rc = 0;
    for (int d = 0; d < count; ++d) {
        if (cudaSetDevice(d) != cudaSuccess) continue;
        rc |= set_watchdog_adain(ns);
        rc |= set_watchdog_cov(ns);
        rc |= set_watchdog_flash(ns);
        rc |= set_watchdog_gemm(ns);
        rc |= set_watchdog_sanet(ns);
        rc |= set_watchdog_seg(ns);
        rc |= set_watchdog_pwconv(ns);
    }
    cudaSetDevice(cur);
    if (rc) {
        set_error("watchdog_ms: cudaMemcpyToSymbol failed");
        return RPST_ERR_CUDA;
    }
    return RPST_OK;
}

}  // namespace rpst

extern "C" int rpst_version(void) { return RPST_VERSION; }

namespace rpst { int64_t ns_flagged_total(); }

extern "C" const char* rpst_last_error(void) { return rpst::g_err; }

extern "C" int rpst_set_tuning(const char* name, int64_t value) {
    if (name && !strcmp(name, "watchdog_ms")) return rpst::apply_watchdog(value);
    if (name && rpst::set_adain_tuning(name, value, true, nullptr)) return RPST_OK;
    rpst::set_error("unknown tuning knob '%s'", name ? name : "(null)");
    return RPST_ERR_INVALID;
}

extern "C" int64_t rpst_get_tuning(const char* name) {
    int64_t v = -1;
    if (name && !strcmp(name, "watchdog_ms")) return rpst::g_watchdog_ms;
    if (name && !strcmp(name, "wct_ns_flagged")) return rpst::ns_flagged_total();   // read-only counter (synchronises)
    if (name && rpst::set_adain_tuning(name, 0, false, &v)) return v;
    return -1;
}
