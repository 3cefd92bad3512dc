// l2_probe.cu -- how much of a producer kernel's output is still L2-resident when the next kernel reads it back, while a larger
// input stream passes through the same L2?  (Design probe for the frame-at-a-time content-aware pipeline: P1 reads a 99.5 MB 4K
// frame and leaves a 33 / 66 MB intermediate that P2 and P3 re-read.)
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o l2_probe scripts/dev/l2_probe.cu && ./l2_probe
//
// producer: reads X (sx MB; ld.global.nc.L1::no_allocate with or without an L2::evict_first policy) and writes T (st MB; plain or
//           L2::evict_last stores);   consumer: reads T forwards or backwards (+ optionally streams X again) and sums it.
// Printed: consumer time and the bandwidth it corresponds to; > 6.5 TB/s means L2 hits.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ float4 ld_stream(const float4* p, bool ef, unsigned long long pol)
{
    float4 v;
    if (ef) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    else asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// mode bit 0: evict_first on the X loads; bit 1: evict_last on the T stores
__global__ void producer(const float4* __restrict__ X, size_t nx, float4* __restrict__ T, size_t nt, int mode)
{
    unsigned long long pol_ef, pol_el;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_ef));
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_el));
    const size_t stride = size_t(gridDim.x) * blockDim.x, i0 = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    // every thread block walks X and T proportionally, so that T is produced WHILE X streams through (like a real kernel)
    const size_t ratio = nt ? (nx + nt - 1) / nt : 0;
    float acc = 0.f;
    for (size_t i = i0; i < (nt ? nt : nx); i += stride) {
        if (nt) {
            for (size_t k = 0; k < ratio; ++k) {
                const size_t j = k * nt + i;
                if (j < nx) { const float4 v = ld_stream(X + j, mode & 1, pol_ef); acc += v.x + v.y + v.z + v.w; }
            }
            const float4 o = make_float4(acc, acc, acc, acc);
            if (mode & 2) asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(T + i), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w), "l"(pol_el) : "memory");
            else T[i] = o;
        } else {
            const float4 v = ld_stream(X + i, mode & 1, pol_ef); acc += v.x;
        }
    }
    if (acc == 123.456f) printf("x");
}

// dir: 0 forwards, 1 backwards
__global__ void consumer(const float4* __restrict__ T, size_t nt, const float4* __restrict__ X, size_t nx, int dir, int mode, float* sink)
{
    unsigned long long pol_ef;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_ef));
    const size_t stride = size_t(gridDim.x) * blockDim.x, i0 = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const size_t ratio = nx ? (nx + nt - 1) / nt : 0;
    float acc = 0.f;
    for (size_t i = i0; i < nt; i += stride) {
        const size_t ii = dir ? nt - 1 - i : i;
        const float4 v = __ldcg(T + ii);
        acc += v.x + v.y + v.z + v.w;
        for (size_t k = 0; k < ratio; ++k) {
            const size_t j = k * nt + ii;
            if (j < nx) { const float4 u = ld_stream(X + j, mode & 1, pol_ef); acc += u.x; }
        }
    }
    if (acc == 123.456f) *sink = acc;
}

int main()
{
    const size_t MB = 1 << 20;
    const size_t maxx = 400 * MB, maxt = 128 * MB;
    float4 *X, *T, *F; float* sink;
    CK(cudaMalloc(&X, maxx)); CK(cudaMalloc(&T, maxt)); CK(cudaMalloc(&F, 512 * MB)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(X, 0, maxx)); CK(cudaMemset(T, 0, maxt));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int grid = 148 * 8, blk = 256;
    int l2 = 0, persist = 0; cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, 0); cudaDeviceGetAttribute(&persist, cudaDevAttrMaxPersistingL2CacheSize, 0);
    printf("L2 %d MB, max persisting %d MB\n", l2 >> 20, persist >> 20);
    printf("%6s %6s %5s %4s %6s | %10s %10s | %10s %10s\n", "st_MB", "sx_MB", "mode", "dir", "cons_x", "prod_ms", "prod_GB/s", "cons_ms", "cons_GB/s");
    const int sts[] = {16, 33, 50, 66, 100};
    const int sxs[] = {0, 100, 200};
    for (int st : sts) for (int sx : sxs) for (int mode = 0; mode < 4; ++mode) for (int dir = 0; dir < 2; ++dir) for (int cx = 0; cx < 2; ++cx) {
        if (sx == 0 && (mode & 1)) continue;
        if (cx && sx == 0) continue;
        if (cx && dir) continue;
        const size_t nt = st * MB / 16, nx = sx * MB / 16;
        float best_p = 1e9f, best_c = 1e9f;
        for (int rep = 0; rep < 4; ++rep) {
            CK(cudaMemsetAsync(F, rep, 512 * MB));   // flush L2
            CK(cudaEventRecord(e0));
            producer<<<grid, blk>>>(X, nx, T, nt, mode);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float tp; CK(cudaEventElapsedTime(&tp, e0, e1));
            CK(cudaEventRecord(e0));
            consumer<<<grid, blk>>>(T, nt, X, cx ? nx : 0, dir, mode, sink);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float tc; CK(cudaEventElapsedTime(&tc, e0, e1));
            if (tp < best_p) best_p = tp;
            if (tc < best_c) best_c = tc;
        }
        const double pb = double(st + sx) * MB, cb = double(st + (cx ? sx : 0)) * MB;
        printf("%6d %6d %5d %4d %6d | %10.4f %10.0f | %10.4f %10.0f\n", st, sx, mode, dir, cx, best_p, pb / best_p / 1e6, best_c, cb / best_c / 1e6);
    }
    // three-kernel chain in one stream, no sync between (what the pipeline would do): P(x->T) , C(T + x) , C(T + x) per "frame", 8 frames
    printf("\nchain: 8 frames x {producer(x 100 MB -> T st MB), consumer(T + x), consumer(T + 2x)} back to back; DRAM-only bound for comparison\n");
    for (int st : {33, 66}) for (int mode : {0, 1, 3}) {
        const size_t nt = st * MB / 16, nx = 100 * MB / 16;
        float best = 1e9f;
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaMemsetAsync(F, rep, 512 * MB));
            CK(cudaEventRecord(e0));
            for (int f = 0; f < 8; ++f) {
                producer<<<grid, blk>>>(X + (f % 3) * nx, nx, T, nt, mode);
                consumer<<<grid, blk>>>(T, nt, X + (f % 3) * nx, nx, 1, mode, sink);
                consumer<<<grid, blk>>>(T, nt, X + ((f + 1) % 3) * nx, nx, 0, mode, sink);
            }
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float t; CK(cudaEventElapsedTime(&t, e0, e1));
            if (t < best) best = t;
        }
        const double all = 8.0 * (3 * 100 + 3 * st) * MB, dram_min = 8.0 * (3 * 100) * MB;
        printf("st %3d mode %d: %.3f ms; all-bytes rate %.0f GB/s; if T never touched DRAM the x streams alone would be %.0f GB/s\n", st, mode, best,
               all / best / 1e6, dram_min / best / 1e6);
    }
    return 0;
}
