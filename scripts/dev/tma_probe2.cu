// canonical libcu++ TMA sample (CUDA programming guide) as a probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
constexpr int SW = 64, SH = 8;
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void kernel(const __grid_constant__ CUtensorMap tensor_map, int x, int y, float* out)
{
    __shared__ alignas(128) float smem_buffer[SH][SW];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
    } else {
        token = bar.arrive();
    }
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < SW * SH; i += blockDim.x) out[i] = smem_buffer[i / SW][i % SW];
}
int main(int argc, char** argv)
{
    const int X = argc > 1 ? atoi(argv[1]) : 64, Y = argc > 2 ? atoi(argv[2]) : 16;
    const int w = 1920, h = 1080;
    std::vector<float> hx(size_t(w) * h);
    for (size_t i = 0; i < hx.size(); ++i) hx[i] = float(i % 100003);
    float *dx, *dout;
    cudaMalloc(&dx, hx.size() * 4); cudaMalloc(&dout, SW * SH * 4);
    cudaMemcpy(dx, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    CUtensorMap tmap; memset(&tmap, 0, sizeof tmap);
    const cuuint64_t gdim[2] = {w, h};
    const cuuint64_t gstride[1] = {cuuint64_t(w) * 4};
    const cuuint32_t box[2] = {SW, SH};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = reinterpret_cast<EncodeTiledFn>(p)(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dx, gdim, gstride, box, estr,
        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d\n", int(r));
    kernel<<<1, 128>>>(tmap, X, Y, dout);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    std::vector<float> ho(SW * SH);
    cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int rr = 0; rr < SH; ++rr) for (int c = 0; c < SW; ++c) { const int gy = Y + rr, gx = X + c; const float ref = (gy < 0 || gy >= h || gx < 0 || gx >= w) ? 0.f : hx[size_t(gy) * w + gx]; bad += ho[rr * SW + c] != ref; }
    printf("mismatches %d\n", bad);
    return 0;
}
