// standalone probe of the TMA tile load used by upr_ext.cu (development aid)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
#ifndef BW
#define BW 160
#endif
#ifndef BH
#define BH 62
#endif
#ifndef L2P
#define L2P CU_TENSOR_MAP_L2_PROMOTION_L2_128B
#endif
__global__ void probe(const __grid_constant__ CUtensorMap tmap, float* out, int x0, int y0, int z)
{
    extern __shared__ __align__(128) unsigned char raw[];
    float* tile = reinterpret_cast<float*>(raw);
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        printf("smem tile addr %u bar addr %u\n", smem_u32(tile), smem_u32(&bar));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(BW * BH * 4) : "memory");
        printf("expect_tx done; tmap generic addr %p\n", (const void*)&tmap);
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(smem_u32(tile)), "l"(&tmap), "r"(x0), "r"(y0), "r"(z), "r"(smem_u32(&bar)) : "memory");
    }
    uint32_t done = 0; int spin = 0;
    while (!done && spin < (1 << 20)) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        ++spin;
    }
    if (threadIdx.x == 0) printf("done %u after %d spins\n", done, spin);
    for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = done ? tile[i] : -1.0f;
}
int main()
{
    const int w = 1920, h = 1080, planes = 3;
    std::vector<float> hx(size_t(w) * h * planes);
    for (size_t i = 0; i < hx.size(); ++i) hx[i] = float(i % 100003);
    float *dx, *dout;
    cudaMalloc(&dx, hx.size() * 4); cudaMalloc(&dout, BW * BH * 4);
    cudaMemcpy(dx, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    printf("entry point: err %d q %d p %p\n", int(e), int(q), p);
    CUtensorMap tmap; memset(&tmap, 0, sizeof tmap);
    const cuuint64_t gdim[3] = {w, h, planes};
    const cuuint64_t gstride[2] = {cuuint64_t(w) * 4, cuuint64_t(w) * h * 4};
    const cuuint32_t box[3] = {BW, BH, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = reinterpret_cast<EncodeTiledFn>(p)(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dx, gdim, gstride, box, estr,
        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, L2P, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode result %d\n", int(r));
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, BW * BH * 4 + 1024);
    for (int t = 0; t < 2; ++t) {
        const int x0 = t == 0 ? 113 : -15, y0 = t == 0 ? 200 : -15, z = 1;
        probe<<<1, 256, BW * BH * 4 + 1024>>>(tmap, dout, x0, y0, z);
        e = cudaDeviceSynchronize();
        printf("kernel: %s\n", cudaGetErrorString(e));
        std::vector<float> ho(BW * BH);
        cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int rr = 0; rr < BH; ++rr) for (int c = 0; c < BW; ++c) {
            const int gy = y0 + rr, gx = x0 + c;
            const float ref = (gy < 0 || gy >= h || gx < 0 || gx >= w) ? 0.0f : hx[size_t(z) * w * h + size_t(gy) * w + gx];
            if (ho[rr * BW + c] != ref) { if (bad < 5) printf("mismatch r%d c%d got %f ref %f\n", rr, c, ho[rr * BW + c], ref); ++bad; }
        }
        printf("tile %d: %d mismatches\n", t, bad);
    }
    return 0;
}
