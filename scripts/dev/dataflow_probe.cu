// dataflow_probe.cu -- would a persistent, frame-at-a-time pipeline keep the content-aware intermediate in L2?
//
// One cooperative kernel emulates the memory behaviour of the three content-aware phases per 4K frame
//   A: read x (99.5 MB), write T (33 MB)        B: read T, read x        C: read T, read e (99.5 MB), write o (99.5 MB)
// with grid-wide barriers between phases (16 frames = 48 barriers).  Variants:
//   reuse=1  T lives in ONE 33 MB slot per group that every frame overwrites (what the pipeline would do: L2 hits expected)
//   reuse=0  every (frame, phase) touches a fresh T region (no reuse possible: the DRAM baseline for the same instruction stream)
//   groups=2 even / odd CTAs work on different frames, half a phase apart (issue-bound and memory-bound phases overlap;
//            two T slots live)
//   hint=1   x / e / o streams carry L2::evict_first
// Printed: ms for 16 frames and the equivalent GB/s over the 48 B/px of compulsory DRAM traffic.
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ float4 ldx(const float4* p, int hint, unsigned long long pol)
{
    float4 v;
    if (hint) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    else asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void stx(float4* p, float4 v, int hint, unsigned long long pol)
{
    if (hint) asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
    else asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// group-local barrier: `bar` counts arrivals; generation = arrivals / nblk
__device__ __forceinline__ void group_barrier(unsigned* bar, unsigned nblk, unsigned& gen)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        ++gen;
        while (*((volatile unsigned*)bar) < gen * nblk) { }
        __threadfence();
    }
    __syncthreads();
}

// plane4 = float4 per 33 MB plane; x/e/o frames are 3 planes
__global__ void __launch_bounds__(256) pipeline(const float4* __restrict__ X, const float4* __restrict__ E, float4* __restrict__ O, float4* __restrict__ T,
                                                size_t plane4, int frames, int reuse, int groups, int hint, unsigned* bars, float* sink)
{
    unsigned long long pol;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    const int grp = blockIdx.x % groups;
    const unsigned nblk = gridDim.x / groups;
    const size_t gtid = size_t(blockIdx.x / groups) * blockDim.x + threadIdx.x, gstride = size_t(nblk) * blockDim.x;
    unsigned gen = 0;
    float acc = 0.f;
    for (int f = grp; f < frames; f += groups) {
        const float4* x = X + size_t(f % 4) * 3 * plane4;
        const float4* e = E + size_t(f % 4) * 3 * plane4;
        float4* o = O + size_t(f % 4) * 3 * plane4;
        float4* tA = T + size_t(reuse ? grp : (3 * f + 0)) * plane4;
        float4* tB = T + size_t(reuse ? grp : (3 * f + 1)) * plane4;
        float4* tC = T + size_t(reuse ? grp : (3 * f + 2)) * plane4;
        // A
        for (size_t i = gtid; i < plane4; i += gstride) {
            const float4 a = ldx(x + i, hint, pol), b = ldx(x + plane4 + i, hint, pol), c = ldx(x + 2 * plane4 + i, hint, pol);
            tA[i] = make_float4(a.x + b.x, a.y + c.y, b.z, c.w);
        }
        group_barrier(bars + grp * 32, nblk, gen);
        // B
        for (size_t i = gtid; i < plane4; i += gstride) {
            const float4 t = __ldcg(tB + i);
            const float4 a = ldx(x + i, hint, pol), b = ldx(x + plane4 + i, hint, pol), c = ldx(x + 2 * plane4 + i, hint, pol);
            acc += t.x + a.x + b.y + c.z;
        }
        group_barrier(bars + grp * 32, nblk, gen);
        // C
        for (size_t i = gtid; i < plane4; i += gstride) {
            const float4 t = __ldcg(tC + i);
            const float4 a = ldx(e + i, hint, pol), b = ldx(e + plane4 + i, hint, pol), c = ldx(e + 2 * plane4 + i, hint, pol);
            stx(o + i, make_float4(a.x * t.x, a.y, a.z, a.w), hint, pol);
            stx(o + plane4 + i, make_float4(b.x * t.y, b.y, b.z, b.w), hint, pol);
            stx(o + 2 * plane4 + i, make_float4(c.x * t.z, c.y, c.z, c.w), hint, pol);
        }
        group_barrier(bars + grp * 32, nblk, gen);
    }
    if (acc == 123.456f) *sink = acc;
}

int main()
{
    const size_t plane4 = size_t(2160) * 3840 / 4;          // float4 per plane (33.2 MB)
    const int frames = 16;
    float4 *X, *E, *O, *T; unsigned* bars; float* sink;
    CK(cudaMalloc(&X, 4 * 3 * plane4 * 16)); CK(cudaMalloc(&E, 4 * 3 * plane4 * 16)); CK(cudaMalloc(&O, 4 * 3 * plane4 * 16));
    CK(cudaMalloc(&T, size_t(3 * frames) * plane4 * 16)); CK(cudaMalloc(&bars, 1024)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(X, 0, 4 * 3 * plane4 * 16)); CK(cudaMemset(E, 0, 4 * 3 * plane4 * 16)); CK(cudaMemset(T, 0, size_t(3 * frames) * plane4 * 16));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pipeline, 256, 0));
    printf("resident CTAs per SM: %d\n", per_sm);
    const double px = double(frames) * 2160 * 3840;
    printf("%7s %6s %6s %5s | %8s %12s %14s\n", "ctas/sm", "reuse", "groups", "hint", "ms", "GB/s@48B/px", "GB/s@64B/px");
    cudaStream_t st; CK(cudaStreamCreate(&st));
    for (int persist : {0, 1}) for (int cps : {4, 6}) for (int groups : {1, 2}) for (int reuse : {0, 1}) for (int hint : {0}) {
        if (cps > per_sm) continue;
        if (persist && !reuse) continue;
        if (persist) {
            // pin the T slots (33 MB per group) in the persisting L2 carve-out
            CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, size_t(79) << 20));
            cudaStreamAttrValue av{};
            av.accessPolicyWindow.base_ptr = T;
            av.accessPolicyWindow.num_bytes = size_t(groups) * plane4 * 16;
            av.accessPolicyWindow.hitRatio = 1.0f;
            av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            CK(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av));
        }
        int grid = 148 * cps; grid -= grid % groups;
        float best = 1e9f;
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaMemset(bars, 0, 1024));
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0, st));
            int fr = frames, r = reuse, g = groups, h = hint; size_t p4 = plane4;
            void* args[] = {&X, &E, &O, &T, &p4, &fr, &r, &g, &h, &bars, &sink};
            CK(cudaLaunchCooperativeKernel((void*)pipeline, dim3(grid), dim3(256), args, 0, st));
            CK(cudaEventRecord(e1, st)); CK(cudaEventSynchronize(e1));
            float t; CK(cudaEventElapsedTime(&t, e0, e1));
            if (t < best) best = t;
        }
        printf("%7d %6d %6d %5d | %8.3f %12.0f %14.0f  persist=%d\n", cps, reuse, groups, hint, best, 48.0 * px / best / 1e6, 64.0 * px / best / 1e6, persist);
        if (persist) {
            cudaStreamAttrValue av{};
            av.accessPolicyWindow.num_bytes = 0;
            CK(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av));
            CK(cudaCtxResetPersistingL2Cache());
        }
    }
    // barrier cost alone: same kernel on a tiny plane (48 group barriers, no data)
    {
        CK(cudaMemset(bars, 0, 1024));
        int grid = 148 * 4, fr = frames, r = 1, g = 1, h = 0; size_t p4 = 1024;
        void* args[] = {&X, &E, &O, &T, &p4, &fr, &r, &g, &h, &bars, &sink};
        CK(cudaLaunchCooperativeKernel((void*)pipeline, dim3(grid), dim3(256), args, 0, st));
        CK(cudaDeviceSynchronize());
        CK(cudaMemset(bars, 0, 1024));
        CK(cudaEventRecord(e0, st));
        CK(cudaLaunchCooperativeKernel((void*)pipeline, dim3(grid), dim3(256), args, 0, st));
        CK(cudaEventRecord(e1, st)); CK(cudaEventSynchronize(e1));
        float t; CK(cudaEventElapsedTime(&t, e0, e1));
        printf("48 barriers on an empty pipeline (592 CTAs): %.3f ms\n", t);
    }
    return 0;
}
