// libtmx.so — error state, version, per-device init.
#include "tmx_common.cuh"
#include <cstring>

namespace tmx {

static thread_local char g_err[512] = "";
static int g_inited[64] = {0};
static int g_sms[64] = {0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return TMX_OK;
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return TMX_ECUDA;
}

static int current_device() {
    int d = -1;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return -1;
    return d;
}

int require_init() {
    int d = current_device();
    if (d < 0 || !g_inited[d]) {
        set_error("tmx_init(device) has not been called for the current device (%d)", d);
        return TMX_ENOINIT;
    }
    return TMX_OK;
}

int sm_count() {
    int d = current_device();
    return d >= 0 && g_sms[d] > 0 ? g_sms[d] : 148;
}

int attn_init();      // attention.cu: function attributes + driver entry point for tensor maps
int groupnorm_init(); // groupnorm.cu
int routed_init();    // routed_linear.cu
int linear_init();    // linear.cu

}  // namespace tmx

extern "C" int tmx_version(void) { return TMX_VERSION; }

extern "C" const char* tmx_last_error(void) { return tmx::g_err; }

extern "C" int tmx_init(int device) {
    using namespace tmx;
    TMX_REQUIRE(device >= 0 && device < 64, TMX_EINVAL, "tmx_init: bad device %d", device);
    TMX_CUDA(cudaSetDevice(device));
    cudaDeviceProp p;
    TMX_CUDA(cudaGetDeviceProperties(&p, device));
    TMX_REQUIRE(p.major == 10, TMX_EARCH,
                "tmx_init: device %d is sm_%d%d; libtmx is built for sm_100a only (no fallback)",
                device, p.major, p.minor);
    g_sms[device] = p.multiProcessorCount;
    int rc = attn_init();
    if (rc) return rc;
    rc = groupnorm_init();
    if (rc) return rc;
    rc = routed_init();
    if (rc) return rc;
    rc = linear_init();
    if (rc) return rc;
    g_inited[device] = 1;
    return TMX_OK;
}
