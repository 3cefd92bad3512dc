// upr_api.cu -- library-level entry points of the C ABI (include/upretinex_b200.h).
#include "upr_common.cuh"

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
