// upr_api.cu -- library-level entry points of the C ABI (include/upretinex_b200.h).
#include "upr_common.cuh"

namespace upr {

static std::mutex g_side_pool_mutex;
static SidePool g_side_pools[64];

std::mutex& side_pool_mutex() { return g_side_pool_mutex; }

SidePool* side_pool()     // call with side_pool_mutex() held
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    SidePool& p = g_side_pools[dev];
    if (!p.ready) {
        bool ok = cudaEventCreateWithFlags(&p.fork, cudaEventDisableTiming) == cudaSuccess;
        for (int i = 0; i < 2 && ok; ++i)
            ok = cudaStreamCreateWithFlags(&p.s[i], cudaStreamNonBlocking) == cudaSuccess &&
                 cudaEventCreateWithFlags(&p.join[i], cudaEventDisableTiming) == cudaSuccess;
        if (!ok) { (void)cudaGetLastError(); return nullptr; }
        p.ready = true;
    }
    return &p;
}

}  // namespace upr

extern "C" {

const char* upr_version(void) { return "upretinex_b200 0.1.0 (sm_100a)"; }

const char* upr_status_string(int status)
{
    switch (status) {
        case UPR_OK: return "UPR_OK";
        case UPR_E_NULL: return "UPR_E_NULL: null pointer argument";
        case UPR_E_SHAPE: return "UPR_E_SHAPE: non-positive or unsupported shape";
        case UPR_E_WORKSPACE: return "UPR_E_WORKSPACE: workspace too small or misaligned";
        case UPR_E_PARAM: return "UPR_E_PARAM: bad scalar parameter";
        case UPR_E_DEVICE: return "UPR_E_DEVICE: current device is not sm_100";
        default: break;
    }
    return status > 0 ? cudaGetErrorString(static_cast<cudaError_t>(status)) : "UPR_E_?: unknown status";
}

int upr_device_check(void)
{
    int dev = 0, major = 0;
    UPR_CUDA_TRY(cudaGetDevice(&dev));
    UPR_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    return major == 10 ? UPR_OK : UPR_E_DEVICE;
}

}  // extern "C"
