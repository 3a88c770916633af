// pixelrec_b200 -- ABI housekeeping: version, thread-local error string, cached device attributes.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace pr {

static thread_local char g_err[512] = "";

void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// kernel-variant switches (bit mask), overridable with PR_TUNE for A/B measurements on the GPU box:
//   1 = add_ln backward as per-warp TMA row pipelines, 2 = L2 prefetch of a warp's next row in the register LN kernels
static int g_tune = -1;
int tune() {
    if (g_tune < 0) {
        const char* e = getenv("PR_TUNE");
        g_tune = e ? atoi(e) : PR_TUNE_DEFAULT;
        if (g_tune < 0) g_tune = 0;
    }
    return g_tune;
}
void set_tune(int mask) { g_tune = mask; }

static const unsigned long long* g_seed_dev = nullptr;
const unsigned long long* seed_device() { return g_seed_dev; }
void set_seed_device(const unsigned long long* p) { g_seed_dev = p; }

}  // namespace pr

extern "C" int pr_version(void) { return PR_ABI_VERSION; }
extern "C" const char* pr_last_error_string(void) { return pr::g_err; }
extern "C" int pr_sm_count(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        pr::set_last_error("pr_sm_count: no CUDA device");
        return PR_ERR_UNSUPPORTED;
    }
    return pr::sm_count();
}
extern "C" int pr_set_tuning(int mask) {
    if (mask >= 0) pr::set_tune(mask);
    return pr::tune();
}
extern "C" int pr_set_seed_device(const uint64_t* seed_offset_dev) {
#ifdef PR_SEED_DEV
    pr::set_seed_device((const unsigned long long*)seed_offset_dev);
    return PR_OK;
#else
    if (seed_offset_dev == nullptr) return PR_OK;
    pr::set_last_error("pr_set_seed_device: this build has no device-side seeds (rebuild with -DPR_SEED_DEV)");
    return PR_ERR_UNSUPPORTED;
#endif
}
extern "C" int pr_set_device(int device) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        pr::set_last_error("pr_set_device(%d): %s", device, cudaGetErrorString(e));
        return (int)e;
    }
    return PR_OK;
}
