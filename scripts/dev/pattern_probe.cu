// DRAM access-pattern probe (development): read / write 64 x 3 x 1080 x 1920 f32 planes
//   tile pattern: one CTA per (frame, 240 x 135 tile): 960-byte row segments, the pattern of the CLAHE kernels
//   band pattern: one CTA per (frame, 17 full-width rows): 7680-byte rows
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pattern_probe pattern_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int H = 1080, W = 1920, N = 64, W4 = W / 4;
constexpr size_t PLANE4 = size_t(H) * W4;

__global__ void __launch_bounds__(256, 3) rd_tile(const float4* __restrict__ in, float* out)
{
    extern __shared__ char pad[];
    const int tile = blockIdx.x, f = blockIdx.y, ty = tile / 8, tx = tile % 8;
    const int lr = threadIdx.x / 60, lc = threadIdx.x % 60;
    float acc = 0.f;
    if (lr < 4) {
        const float4* p = in + size_t(f) * 3 * PLANE4 + size_t(ty * 135) * W4 + tx * 60 + lc;
        for (int r = lr; r < 135; r += 4) {
            const float4 a = __ldcs(p + size_t(r) * W4), b = __ldcs(p + PLANE4 + size_t(r) * W4), c = __ldcs(p + 2 * PLANE4 + size_t(r) * W4);
            acc += a.x + a.w + b.y + c.z;
        }
    }
    if (acc == 123.456f) out[0] = acc + pad[0];
}
__global__ void __launch_bounds__(256, 3) rd_band(const float4* __restrict__ in, float* out, int rows)
{
    extern __shared__ char pad[];
    const int band = blockIdx.x, f = blockIdx.y;
    float acc = 0.f;
    const float4* p = in + size_t(f) * 3 * PLANE4 + size_t(band * rows) * W4;
    const int r1 = min(rows, H - band * rows);
    for (int r = 0; r < r1; ++r)
        for (int c = threadIdx.x; c < W4; c += 256) {
            const float4 a = __ldcs(p + size_t(r) * W4 + c), b = __ldcs(p + PLANE4 + size_t(r) * W4 + c), cc = __ldcs(p + 2 * PLANE4 + size_t(r) * W4 + c);
            acc += a.x + a.w + b.y + cc.z;
        }
    if (acc == 123.456f) out[0] = acc + pad[0];
}
__global__ void __launch_bounds__(256, 3) wr_tile(float4* __restrict__ o)
{
    extern __shared__ char pad[];
    const int tile = blockIdx.x, f = blockIdx.y, ty = tile / 8, tx = tile % 8;
    const int lr = threadIdx.x / 60, lc = threadIdx.x % 60;
    if (lr < 4) {
        float4* p = o + size_t(f) * 3 * PLANE4 + size_t(ty * 135) * W4 + tx * 60 + lc;
        const float4 v = make_float4(1.f, 2.f, 3.f, float(lc));
        for (int r = lr; r < 135; r += 4) { __stcs(p + size_t(r) * W4, v); __stcs(p + PLANE4 + size_t(r) * W4, v); __stcs(p + 2 * PLANE4 + size_t(r) * W4, v); }
    }
}
__global__ void __launch_bounds__(256, 3) wr_band(float4* __restrict__ o, int rows)
{
    extern __shared__ char pad[];
    const int band = blockIdx.x, f = blockIdx.y;
    float4* p = o + size_t(f) * 3 * PLANE4 + size_t(band * rows) * W4;
    const int r1 = min(rows, H - band * rows);
    const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    for (int r = 0; r < r1; ++r)
        for (int c = threadIdx.x; c < W4; c += 256) { __stcs(p + size_t(r) * W4 + c, v); __stcs(p + PLANE4 + size_t(r) * W4 + c, v); __stcs(p + 2 * PLANE4 + size_t(r) * W4 + c, v); }
}
// K1-like traffic: read 3 f32 planes, write 3 u8 planes (one u32 per 4 pixels), tile or band pattern, loads one row ahead
__global__ void __launch_bounds__(256, 3) k1_tile(const float4* __restrict__ in, unsigned* __restrict__ lab)
{
    extern __shared__ char pad[];
    const int tile = blockIdx.x, f = blockIdx.y, ty = tile / 8, tx = tile % 8;
    const int lr = threadIdx.x / 60, lc = threadIdx.x % 60;
    if (lr >= 4) return;
    const size_t base = size_t(f) * 3 * PLANE4 + size_t(ty * 135) * W4 + tx * 60 + lc;
    const float4* p = in + base;
    unsigned* q = lab + base;
    float4 a = __ldcs(p + size_t(lr) * W4), b = __ldcs(p + PLANE4 + size_t(lr) * W4), c = __ldcs(p + 2 * PLANE4 + size_t(lr) * W4);
    for (int r = lr; r < 135; r += 4) {
        float4 na = a, nb = b, nc = c;
        if (r + 4 < 135) { na = __ldcs(p + size_t(r + 4) * W4); nb = __ldcs(p + PLANE4 + size_t(r + 4) * W4); nc = __ldcs(p + 2 * PLANE4 + size_t(r + 4) * W4); }
        q[size_t(r) * W4] = __float_as_uint(a.x + a.w); q[PLANE4 + size_t(r) * W4] = __float_as_uint(b.y); q[2 * PLANE4 + size_t(r) * W4] = __float_as_uint(c.z);
        a = na; b = nb; c = nc;
    }
}
__global__ void __launch_bounds__(256, 3) k1_band(const float4* __restrict__ in, unsigned* __restrict__ lab, int rows)
{
    extern __shared__ char pad[];
    const int band = blockIdx.x, f = blockIdx.y;
    const size_t base = size_t(f) * 3 * PLANE4 + size_t(band * rows) * W4;
    const float4* p = in + base;
    unsigned* q = lab + base;
    const int r1 = min(rows, H - band * rows);
    for (int r = 0; r < r1; ++r) {
        const int c0 = threadIdx.x, c1 = threadIdx.x + 256;
        const float4 a = __ldcs(p + size_t(r) * W4 + c0), b = __ldcs(p + PLANE4 + size_t(r) * W4 + c0), c = __ldcs(p + 2 * PLANE4 + size_t(r) * W4 + c0);
        float4 a2 = a, b2 = b, c2 = c;
        if (c1 < W4) { a2 = __ldcs(p + size_t(r) * W4 + c1); b2 = __ldcs(p + PLANE4 + size_t(r) * W4 + c1); c2 = __ldcs(p + 2 * PLANE4 + size_t(r) * W4 + c1); }
        q[size_t(r) * W4 + c0] = __float_as_uint(a.x + a.w); q[PLANE4 + size_t(r) * W4 + c0] = __float_as_uint(b.y); q[2 * PLANE4 + size_t(r) * W4 + c0] = __float_as_uint(c.z);
        if (c1 < W4) { q[size_t(r) * W4 + c1] = __float_as_uint(a2.x + a2.w); q[PLANE4 + size_t(r) * W4 + c1] = __float_as_uint(b2.y); q[2 * PLANE4 + size_t(r) * W4 + c1] = __float_as_uint(c2.z); }
    }
}
// K3-like traffic: read 3 u8 planes (one u32 per 4 pixels), write 3 f32 planes
__global__ void __launch_bounds__(512, 2) k3_tile(const unsigned* __restrict__ lab, float4* __restrict__ o)
{
    const int tile = blockIdx.x, f = blockIdx.y, ty = tile / 8, tx = tile % 8;
    const int lr = threadIdx.x / 60, lc = threadIdx.x % 60;   // 480 threads = 8 rows x 60 columns
    const size_t base = size_t(f) * 3 * PLANE4 + size_t(ty * 135) * W4 + tx * 60 + lc;
    for (int r = lr; r < 135; r += 8) {
        const unsigned a = __ldg(lab + base + size_t(r) * W4), b = __ldg(lab + base + PLANE4 + size_t(r) * W4), c = __ldg(lab + base + 2 * PLANE4 + size_t(r) * W4);
        const float4 v = make_float4(float(a & 255), float(b & 255), float(c & 255), float(a >> 24));
        __stcs(o + base + size_t(r) * W4, v); __stcs(o + base + PLANE4 + size_t(r) * W4, v); __stcs(o + base + 2 * PLANE4 + size_t(r) * W4, v);
    }
}
__global__ void __launch_bounds__(512, 2) k3_band(const unsigned* __restrict__ lab, float4* __restrict__ o, int rows)
{
    const int band = blockIdx.x, f = blockIdx.y;
    const size_t base = size_t(f) * 3 * PLANE4 + size_t(band * rows) * W4 + threadIdx.x;   // 480 threads = one full row
    const int r1 = min(rows, H - band * rows);
    for (int r = 0; r < r1; ++r) {
        const unsigned a = __ldg(lab + base + size_t(r) * W4), b = __ldg(lab + base + PLANE4 + size_t(r) * W4), c = __ldg(lab + base + 2 * PLANE4 + size_t(r) * W4);
        const float4 v = make_float4(float(a & 255), float(b & 255), float(c & 255), float(a >> 24));
        __stcs(o + base + size_t(r) * W4, v); __stcs(o + base + PLANE4 + size_t(r) * W4, v); __stcs(o + base + 2 * PLANE4 + size_t(r) * W4, v);
    }
}
template <typename F> float timeit(F f)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    cudaEventRecord(a);
    for (int i = 0; i < 10; ++i) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / 10;
}
int main()
{
    float4 *in, *o; float* out;
    const size_t bytes = size_t(N) * 3 * PLANE4 * 16;
    cudaMalloc(&in, bytes); cudaMalloc(&o, bytes); cudaMalloc(&out, 4);
    cudaMemset(in, 0, bytes);
    const size_t smem = 70656;
    cudaFuncSetAttribute(rd_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    cudaFuncSetAttribute(rd_band, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    cudaFuncSetAttribute(wr_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    cudaFuncSetAttribute(wr_band, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    const double gb = bytes / 1e9;
    float t;
    t = timeit([&] { rd_tile<<<dim3(64, N), 256, smem>>>(in, out); }); printf("read  tile pattern          %.3f ms  %.0f GB/s\n", t, gb / t * 1e3);
    for (int rows : {17, 8, 4}) {
        t = timeit([&] { rd_band<<<dim3((H + rows - 1) / rows, N), 256, smem>>>(in, out, rows); }); printf("read  band pattern %2d rows  %.3f ms  %.0f GB/s\n", rows, t, gb / t * 1e3);
    }
    t = timeit([&] { wr_tile<<<dim3(64, N), 256, smem>>>(o); }); printf("write tile pattern          %.3f ms  %.0f GB/s\n", t, gb / t * 1e3);
    for (int rows : {17, 8, 4}) {
        t = timeit([&] { wr_band<<<dim3((H + rows - 1) / rows, N), 256, smem>>>(o, rows); }); printf("write band pattern %2d rows  %.3f ms  %.0f GB/s\n", rows, t, gb / t * 1e3);
    }
    cudaFuncSetAttribute(k1_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    cudaFuncSetAttribute(k1_band, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    unsigned* lab = reinterpret_cast<unsigned*>(o);
    t = timeit([&] { k1_tile<<<dim3(64, N), 256, smem>>>(in, lab); }); printf("K1-like tile pattern        %.3f ms  (12 B/px read + 3 B/px written)\n", t);
    for (int rows : {17, 8, 4}) {
        t = timeit([&] { k1_band<<<dim3((H + rows - 1) / rows, N), 256, smem>>>(in, lab, rows); }); printf("K1-like band pattern %2d rows %.3f ms\n", rows, t);
    }
    const unsigned* labr = reinterpret_cast<const unsigned*>(in);
    t = timeit([&] { k3_tile<<<dim3(64, N), 480>>>(labr, o); }); printf("K3-like tile pattern        %.3f ms  (3 B/px read + 12 B/px written)\n", t);
    for (int rows : {17, 8}) {
        t = timeit([&] { k3_band<<<dim3((H + rows - 1) / rows, N), 480>>>(labr, o, rows); }); printf("K3-like band pattern %2d rows %.3f ms\n", rows, t);
    }
    t = timeit([&] { cudaMemcpyAsync(o, in, bytes, cudaMemcpyDeviceToDevice); }); printf("memcpy d2d (r+w)            %.3f ms  %.0f GB/s\n", t, 2 * gb / t * 1e3);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
