// upr_saliency.cu -- content-aware saliency and attention maps (sm_100a).
//
// Replaces ContentAwareEnhancer.compute_saliency_map / compute_attention_map
// (/root/reference/enhancers/content_aware.py:19-59, :61-91):
//   gray = BGR2GRAY(u8 quantise(x))                     OpenCV fixed point
//   lap  = |cv2.Laplacian(gray, CV_64F)|                4-neighbour, BORDER_REFLECT_101 (integers <= 1020)
//   blur = cv2.GaussianBlur(lap, (15,15), 0)            sigma 2.6, separable, fp64, BORDER_REFLECT_101
//   sal  = float32((blur - min) / (max - min + 1e-8))   per image
//   att  = sal * (1 / (luma(x) + 0.1)); att = (att - min) / (max - min + 1e-8)   fp32, per image
//
// K5 k_saliency_blur: one CTA per 64x64 output tile.  The reflect-101 extension commutes with the (mirror
//    symmetric) Laplacian and Gaussian, so the whole chain is evaluated on the reflected plane: gray
//    (tile+8 halo, u8) -> |lap| (tile+7, u16) -> row pass (fp64, symmetric taps paired so the integer
//    pair sums are exact) -> column pass (fp64) -> un-normalised blur stored as fp32, per-CTA fp64 min/max
//    folded into per-image ordered-integer atomics (exact, order independent).
// K6 k_attention_raw: sal normalise + luma division fused, fp32 min/max of the raw attention per image.
// K7 k_normalize: (v - min) / (max - min + 1e-8) in place (used for both maps).
#include <algorithm>
#include <cmath>

#include "upr_common.cuh"

namespace upr {

constexpr int kSalThreads = 256;
constexpr int kSalTile = 64;
constexpr int kGW = kSalTile + 16;  // gray   80 x 80
constexpr int kLW = kSalTile + 14;  // |lap|  78 x 78
constexpr int kRH = kSalTile + 14;  // row pass: 78 rows x 64 cols fp64

struct GaussTaps {
    double t[8];  // t[d] = weight at distance d from the centre of the 15-tap kernel (kernel parameter -> constant bank)
};

__device__ __forceinline__ int reflect101_s(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

// order-preserving map double <-> unsigned 64 (for exact atomic min/max)
__device__ __forceinline__ unsigned long long dbl_key(double v)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_dbl(unsigned long long k)
{
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}
__device__ __forceinline__ unsigned flt_key(float v)
{
    const unsigned b = __float_as_uint(v);
    return (b >> 31) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_flt(unsigned k)
{
    const unsigned b = (k >> 31) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(b);
}

struct SalMinMax {           // per image, 32 bytes
    unsigned long long blur_min, blur_max;   // ordered keys of fp64 values
    unsigned att_min, att_max;               // ordered keys of fp32 values
    unsigned pad0, pad1;
};

__global__ void k_sal_reset(SalMinMax* mm, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    mm[i].blur_min = ~0ull; mm[i].blur_max = 0ull;
    mm[i].att_min = ~0u; mm[i].att_max = 0u;
    mm[i].pad0 = mm[i].pad1 = 0;
}

__global__ void __launch_bounds__(kSalThreads)
k_saliency_blur(const float* __restrict__ x, int h, int w, int tiles_x, float* __restrict__ blur_out, SalMinMax* __restrict__ mm,
                const GaussTaps taps)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    double* s_row = reinterpret_cast<double*>(s_raw);                                  // [kRH][kSalTile]
    unsigned short* s_lap = reinterpret_cast<unsigned short*>(s_row + kRH * kSalTile);  // [kLW][kLW]
    unsigned char* s_gray = reinterpret_cast<unsigned char*>(s_lap + kLW * kLW);        // [kGW][kGW]
    __shared__ double s_mn[kSalThreads / 32], s_mx[kSalThreads / 32];

    const int tid = threadIdx.x;
    const int f = blockIdx.y;
    const int tyi = blockIdx.x / tiles_x, txi = blockIdx.x - tyi * tiles_x;
    const int x0 = txi * kSalTile, y0 = tyi * kSalTile;
    const long long plane = (long long)h * w;
    const float* img = x + (long long)f * 3 * plane;

    // gray on the reflected plane: s_gray[r][c] = gray(reflect(y0-8+r), reflect(x0-8+c))
    for (int i = tid; i < kGW * kGW; i += kSalThreads) {
        const int r = i / kGW, c = i - r * kGW;
        const int gy = reflect101_s(y0 - 8 + r, h), gx = reflect101_s(x0 - 8 + c, w);
        const long long o = (long long)gy * w + gx;
        const int qr = quantize_u8(__ldg(img + o)), qg = quantize_u8(__ldg(img + plane + o)),
                  qb = quantize_u8(__ldg(img + 2 * plane + o));
        s_gray[i] = (unsigned char)((qr * 9798 + qg * 19235 + qb * 3735 + 16384) >> 15);
    }
    __syncthreads();
    // |laplacian| at plane coordinates (y0-7+r, x0-7+c)
    for (int i = tid; i < kLW * kLW; i += kSalThreads) {
        const int r = i / kLW, c = i - r * kLW;
        const unsigned char* g = s_gray + (r + 1) * kGW + c + 1;
        const int v = int(g[-kGW]) + int(g[kGW]) + int(g[-1]) + int(g[1]) - 4 * int(g[0]);
        s_lap[i] = (unsigned short)abs(v);
    }
    __syncthreads();
    // row pass: rows y0-7 .. y0+70, columns x0 .. x0+63
    for (int i = tid; i < kRH * kSalTile; i += kSalThreads) {
        const int r = i / kSalTile, c = i - r * kSalTile;
        const unsigned short* p = s_lap + r * kLW + c + 7;
        double acc = taps.t[0] * double(int(p[0]));
#pragma unroll
        for (int d = 1; d <= 7; ++d) acc = __fma_rn(taps.t[d], double(int(p[-d]) + int(p[d])), acc);
        s_row[i] = acc;
    }
    __syncthreads();
    // column pass + store + min/max
    double mn = INFINITY, mx = -INFINITY;
    for (int i = tid; i < kSalTile * kSalTile; i += kSalThreads) {
        const int r = i / kSalTile, c = i - r * kSalTile;
        const int gy = y0 + r, gx = x0 + c;
        if (gy >= h || gx >= w) continue;
        const double* p = s_row + (r + 7) * kSalTile + c;
        double acc = taps.t[0] * p[0];
#pragma unroll
        for (int d = 1; d <= 7; ++d) acc = __fma_rn(taps.t[d], p[-d * kSalTile] + p[d * kSalTile], acc);
        blur_out[(long long)f * plane + (long long)gy * w + gx] = float(acc);
        mn = fmin(mn, acc);
        mx = fmax(mx, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((tid & 31) == 0) { s_mn[tid >> 5] = mn; s_mx[tid >> 5] = mx; }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int k = 1; k < kSalThreads / 32; ++k) { mn = fmin(mn, s_mn[k]); mx = fmax(mx, s_mx[k]); }
        atomicMin(&mm[f].blur_min, dbl_key(mn));
        atomicMax(&mm[f].blur_max, dbl_key(mx));
    }
}

// sal = float((double(blur) - min) / (max - min + 1e-8)); optional attention raw value + its fp32 min/max
template <bool kAttention>
__global__ void __launch_bounds__(kSalThreads)
k_sal_normalize(const float* __restrict__ blur, const float* __restrict__ x, float* __restrict__ out, long long plane,
                SalMinMax* __restrict__ mm)
{
    __shared__ float s_mn[kSalThreads / 32], s_mx[kSalThreads / 32];
    const int f = blockIdx.y;
    const double bmn = key_dbl(mm[f].blur_min), bmx = key_dbl(mm[f].blur_max);
    const double den = bmx - bmn + 1e-8;
    const float* img = x + (long long)f * 3 * plane;
    float mn = INFINITY, mx = -INFINITY;
    const long long stride = (long long)gridDim.x * kSalThreads;
    for (long long i = (long long)blockIdx.x * kSalThreads + threadIdx.x; i < plane; i += stride) {
        const float s = float((double(blur[(long long)f * plane + i]) - bmn) / den);
        if (kAttention) {
            const float lum = __fadd_rn(__fadd_rn(__fmul_rn(0.299f, __ldg(img + i)), __fmul_rn(0.587f, __ldg(img + plane + i))),
                                        __fmul_rn(0.114f, __ldg(img + 2 * plane + i)));
            const float a = __fmul_rn(s, __fdiv_rn(1.0f, __fadd_rn(lum, 0.1f)));
            out[(long long)f * plane + i] = a;
            mn = fminf(mn, a);
            mx = fmaxf(mx, a);
        } else {
            out[(long long)f * plane + i] = s;
        }
    }
    if constexpr (kAttention) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) { s_mn[threadIdx.x >> 5] = mn; s_mx[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 1; k < kSalThreads / 32; ++k) { mn = fminf(mn, s_mn[k]); mx = fmaxf(mx, s_mx[k]); }
        atomicMin(&mm[f].att_min, flt_key(mn));
        atomicMax(&mm[f].att_max, flt_key(mx));
    }
    }
}

__global__ void __launch_bounds__(kSalThreads)
k_att_normalize(float* __restrict__ att, long long plane, const SalMinMax* __restrict__ mm)
{
    const int f = blockIdx.y;
    const float mn = key_flt(mm[f].att_min), mx = key_flt(mm[f].att_max);
    const float den = __fadd_rn(__fsub_rn(mx, mn), 1e-8f);
    const long long stride = (long long)gridDim.x * kSalThreads;
    for (long long i = (long long)blockIdx.x * kSalThreads + threadIdx.x; i < plane; i += stride) {
        float* p = att + (long long)f * plane + i;
        *p = __fdiv_rn(__fsub_rn(*p, mn), den);
    }
}

static GaussTaps sal_taps()
{
    // cv2.getGaussianKernel(15, sigma=0.3*((15-1)*0.5-1)+0.8 = 2.6), fp64
    double k[15], sum = 0.0;
    const double sigma = 0.3 * ((15 - 1) * 0.5 - 1) + 0.8;
    for (int i = 0; i < 15; ++i) { const double d = i - 7; k[i] = std::exp(-0.5 * d * d / (sigma * sigma)); sum += k[i]; }
    GaussTaps g;
    for (int d = 0; d < 8; ++d) g.t[d] = k[7 + d] / sum;
    return g;
}

static size_t sal_ws_bytes(int n, int h, int w)
{
    return align_up(size_t(n) * sizeof(SalMinMax), 256) + align_up(size_t(n) * h * w * sizeof(float), 256);
}

// mode 0: saliency only -> out ; mode 1: attention -> out (saliency is an internal temporary)
static int sal_run(int mode, const float* x, int n, int h, int w, float* out, void* ws, size_t ws_bytes, cudaStream_t s)
{
    if (n < 0 || n > 65535 || h <= 0 || w <= 0) return UPR_E_SHAPE;
    if (n == 0) return UPR_OK;
    if (!x || !out || !ws) return UPR_E_NULL;
    if (ws_bytes < sal_ws_bytes(n, h, w) || (reinterpret_cast<uintptr_t>(ws) & 255u)) return UPR_E_WORKSPACE;
    static const GaussTaps taps = sal_taps();
    auto* mm = static_cast<SalMinMax*>(ws);
    auto* blur = reinterpret_cast<float*>(static_cast<unsigned char*>(ws) + align_up(size_t(n) * sizeof(SalMinMax), 256));
    const long long plane = (long long)h * w;
    k_sal_reset<<<(n + 127) / 128, 128, 0, s>>>(mm, n);
    UPR_LAUNCH_CHECK();
    const int tiles_x = (w + kSalTile - 1) / kSalTile, tiles_y = (h + kSalTile - 1) / kSalTile;
    const size_t smem = size_t(kRH) * kSalTile * sizeof(double) + size_t(kLW) * kLW * 2 + size_t(kGW) * kGW;
    static bool attr_set = false;
    if (!attr_set) {
        UPR_CUDA_TRY(cudaFuncSetAttribute(k_saliency_blur, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        attr_set = true;
    }
    k_saliency_blur<<<dim3(tiles_x * tiles_y, n), kSalThreads, smem, s>>>(x, h, w, tiles_x, blur, mm, taps);
    UPR_LAUNCH_CHECK();
    const int parts = int(std::max<long long>(1, std::min<long long>((plane + kSalThreads * 4 - 1) / (kSalThreads * 4),
                                                                     (8LL * kNumSMsB200 + n - 1) / n)));
    if (mode == 0) {
        k_sal_normalize<false><<<dim3(parts, n), kSalThreads, 0, s>>>(blur, x, out, plane, mm);
        UPR_LAUNCH_CHECK();
    } else {
        k_sal_normalize<true><<<dim3(parts, n), kSalThreads, 0, s>>>(blur, x, out, plane, mm);
        UPR_LAUNCH_CHECK();
        k_att_normalize<<<dim3(parts, n), kSalThreads, 0, s>>>(out, plane, mm);
        UPR_LAUNCH_CHECK();
    }
    return UPR_OK;
}

}  // namespace upr

extern "C" {

size_t upr_saliency_workspace_bytes(int n, int h, int w)
{
    if (n < 0 || h <= 0 || w <= 0) return 0;
    return upr::sal_ws_bytes(std::max(n, 1), h, w);
}

int upr_saliency_f32(const float* x_nchw, int n, int h, int w, float* sal_n1hw, void* workspace, size_t workspace_bytes,
                     upr_stream_t stream)
{
    return upr::sal_run(0, x_nchw, n, h, w, sal_n1hw, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int upr_attention_f32(const float* x_nchw, int n, int h, int w, float* att_n1hw, void* workspace, size_t workspace_bytes,
                      upr_stream_t stream)
{
    return upr::sal_run(1, x_nchw, n, h, w, att_n1hw, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
