// upr_ext.cu -- EXTENSION ops: operations the north-star names but the reference does not
// contain -- a generic separable Gaussian blur, a Gaussian pyramid level, log-domain single/multi-scale Retinex
// (SSR/MSR: sum_s w_s * (log(I + eps) - log(G_s * I + eps))) and gamma.  They have NO reference parity target; the oracle
// is OpenCV / NumPy directly (cv2.GaussianBlur, cv2.pyrDown, np.log, np.power) and nothing in the reference entry points
// calls them, so default outputs are unchanged.
//
// k_gauss_tile: persistent CTAs; the input tile (128 x 32 outputs + halo up to 15) is staged in shared memory by TMA
// (cp.async.bulk.tensor.3d over a (w, h, planes) tensor map: out-of-image elements arrive as zeros, planes never bleed
// into each other), double buffered -- the tile of work item k+1 is in flight while item k is filtered.  Border CTAs
// then patch the halo with BORDER_REFLECT_101 values from inside the tile.  One staged tile serves every scale of an
// MSR call (the reference-style chain of cv2.GaussianBlur calls would re-read the image once per scale).
// Without TMA (rows not 16-byte aligned, tiny images) the same kernel fills the tile with reflect-indexed loads.
#include <cuda.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "upr_common.cuh"

namespace upr {

constexpr int kGtThreads = 256;
constexpr int kGtTW = 128, kGtTH = 32, kGtMaxR = 15;
constexpr int kGtColOrg = 16;                     // the staged tile starts 16 columns left of the output tile: TMA needs a 16-byte
                                                  // aligned start in the innermost dimension (x0 - radius faults: probed on the box)
constexpr int kGtInW = kGtColOrg + kGtTW + kGtMaxR + 1;   // 160: row pitch of the staged tile (640 B, a multiple of 16)
constexpr int kGtInH = kGtTH + 2 * kGtMaxR;       // 62
constexpr int kGtStageFloats = kGtInW * kGtInH;
constexpr int kGtMaxScales = 4;

struct GaussSpec {
    int nscales;
    int radius[kGtMaxScales];
    float taps[kGtMaxScales][kGtMaxR + 1];   // taps[s][d] = weight at distance d
    float weight[kGtMaxScales];              // MSR weights
    float eps;
    int mode;                                // 0 = blur (scale 0), 1 = MSR
    int rmax;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t done = 0;
    for (int spin = 0; !done; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (spin > (1 << 24)) __trap();   // a TMA that never lands must fail loudly, not hang the device
    }
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, int x, int y, int z, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ int ext_reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

// items = planes * tiles_y * tiles_x ; out may alias nothing (x is read through halos of neighbouring tiles)
template <bool kTma>
__global__ void __launch_bounds__(kGtThreads, 2)
k_gauss_tile(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ x, float* __restrict__ out, int planes, int h,
             int w, int tiles_x, int tiles_y, const GaussSpec spec)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* s_in = reinterpret_cast<float*>(smem_raw);                       // [2][kGtInH][kGtInW]
    float* s_tmp = s_in + 2 * kGtStageFloats;                               // [kGtInH][kGtTW]  horizontally filtered rows
    float* s_acc = s_tmp + kGtInH * kGtTW;                                  // [kGtTH][kGtTW]   MSR accumulator
    __shared__ __align__(8) uint64_t s_bar[2];

    const int tid = threadIdx.x;
    const int R = spec.rmax;
    const long long nitems = (long long)planes * tiles_x * tiles_y;
    const uint32_t stage_bytes = uint32_t(kGtInW) * uint32_t(kGtTH + 2 * R) * 4u;   // box = kGtInW x (TH + 2R) x 1

    auto decode = [&](long long item, int& pl, int& x0, int& y0) {
        pl = int(item / (tiles_x * tiles_y));
        const int rem = int(item - (long long)pl * tiles_x * tiles_y);
        const int tyi = rem / tiles_x, txi = rem - tyi * tiles_x;
        x0 = txi * kGtTW;
        y0 = tyi * kGtTH;
    };
    auto issue = [&](long long item, int stage) {   // one thread
        int pl, x0, y0;
        decode(item, pl, x0, y0);
        mbar_expect_tx(&s_bar[stage], stage_bytes);
        tma_load_3d(s_in + stage * kGtStageFloats, &tmap, x0 - kGtColOrg, y0 - R, pl, &s_bar[stage]);
    };

    if (kTma) {
        if (tid == 0) {
            mbar_init(&s_bar[0], 1);
            mbar_init(&s_bar[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0 && blockIdx.x < nitems) issue(blockIdx.x, 0);
    }

    uint32_t phase[2] = {0u, 0u};
    int stage = 0;
    for (long long item = blockIdx.x; item < nitems; item += gridDim.x, stage ^= 1) {
        int pl, x0, y0;
        decode(item, pl, x0, y0);
        float* tile = s_in + stage * kGtStageFloats;
        const int rows_in = kGtTH + 2 * R;
        if (kTma) {
            const long long nxt = item + gridDim.x;
            if (tid == 0 && nxt < nitems) issue(nxt, stage ^ 1);   // the other stage was released by the barrier below
            mbar_wait(&s_bar[stage], phase[stage]);
            phase[stage] ^= 1u;
            // BORDER_REFLECT_101: TMA delivered zeros outside the image; columns first, then whole rows
            const bool edge = (x0 - R < 0) || (x0 + kGtTW + R > w) || (y0 - R < 0) || (y0 + kGtTH + R > h);
            const int xo = x0 - kGtColOrg;   // image column of tile column 0
            if (edge) {
                for (int i = tid; i < rows_in * kGtInW; i += kGtThreads) {
                    const int r = i / kGtInW, c = i - r * kGtInW;
                    const int gy = y0 - R + r, gx = xo + c;
                    if (gy >= 0 && gy < h && (gx < 0 || gx >= w)) {
                        const int sx = ext_reflect101(gx, w) - xo;
                        if (sx >= 0 && sx < kGtInW) tile[r * kGtInW + c] = tile[r * kGtInW + sx];
                    }
                }
                __syncthreads();
                for (int i = tid; i < rows_in * kGtInW; i += kGtThreads) {
                    const int r = i / kGtInW, c = i - r * kGtInW;
                    const int gy = y0 - R + r;
                    if (gy < 0 || gy >= h) {
                        const int sy = ext_reflect101(gy, h) - (y0 - R);
                        if (sy >= 0 && sy < rows_in) tile[r * kGtInW + c] = tile[sy * kGtInW + c];
                    }
                }
                __syncthreads();
            }
        } else {
            const float* src = x + (long long)pl * h * w;
            for (int i = tid; i < rows_in * kGtInW; i += kGtThreads) {
                const int r = i / kGtInW, c = i - r * kGtInW;
                const int gy = ext_reflect101(y0 - R + r, h), gx = ext_reflect101(x0 - kGtColOrg + c, w);
                tile[i] = __ldg(src + (long long)gy * w + gx);
            }
            __syncthreads();
        }

        for (int s = 0; s < spec.nscales; ++s) {
            const int r = spec.radius[s];
            const float* tp = spec.taps[s];
            // horizontal pass over the rows this scale needs: tile rows [R - r, R + TH + r)
            const int hrows = kGtTH + 2 * r;
            for (int i = tid; i < hrows * (kGtTW / 4); i += kGtThreads) {
                const int rr = i / (kGtTW / 4), c4 = (i - rr * (kGtTW / 4)) * 4;
                const float* p = tile + (R - r + rr) * kGtInW + kGtColOrg + c4;      // centre of output column c4
                // sliding windows: L = p[-d .. 3-d], Rw = p[d .. 3+d]; one new shared load per side and tap distance
                float L0 = p[0], L1 = p[1], L2 = p[2], L3 = p[3];
                float R0 = L0, R1 = L1, R2 = L2, R3 = L3;
                float a0 = tp[0] * L0, a1 = tp[0] * L1, a2 = tp[0] * L2, a3 = tp[0] * L3;
                for (int d = 1; d <= r; ++d) {
                    const float t = tp[d];
                    L3 = L2; L2 = L1; L1 = L0; L0 = p[-d];
                    R0 = R1; R1 = R2; R2 = R3; R3 = p[3 + d];
                    a0 = fmaf(t, L0 + R0, a0);
                    a1 = fmaf(t, L1 + R1, a1);
                    a2 = fmaf(t, L2 + R2, a2);
                    a3 = fmaf(t, L3 + R3, a3);
                }
                *reinterpret_cast<float4*>(s_tmp + rr * kGtTW + c4) = make_float4(a0, a1, a2, a3);
            }
            __syncthreads();
            // vertical pass -> output rows; s_tmp row (rr + r) is the centre of output row rr
            for (int i = tid; i < kGtTH * (kGtTW / 4); i += kGtThreads) {
                const int rr = i / (kGtTW / 4), c4 = (i - rr * (kGtTW / 4)) * 4;
                const float* p = s_tmp + (rr + r) * kGtTW + c4;
                float4 acc = *reinterpret_cast<const float4*>(p);
                acc.x *= tp[0]; acc.y *= tp[0]; acc.z *= tp[0]; acc.w *= tp[0];
                for (int d = 1; d <= r; ++d) {
                    const float4 u = *reinterpret_cast<const float4*>(p - d * kGtTW);
                    const float4 v = *reinterpret_cast<const float4*>(p + d * kGtTW);
                    const float t = tp[d];
                    acc.x = fmaf(t, u.x + v.x, acc.x);
                    acc.y = fmaf(t, u.y + v.y, acc.y);
                    acc.z = fmaf(t, u.z + v.z, acc.z);
                    acc.w = fmaf(t, u.w + v.w, acc.w);
                }
                float res[4] = {acc.x, acc.y, acc.z, acc.w};
                if (spec.mode == 1) {
                    const float* ctr = tile + (R + rr) * kGtInW + kGtColOrg + c4;
                    float* ap = s_acc + rr * kGtTW + c4;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float term = spec.weight[s] * (logf(ctr[k] + spec.eps) - logf(res[k] + spec.eps));
                        res[k] = s == 0 ? term : ap[k] + term;
                        ap[k] = res[k];
                    }
                }
                if (spec.mode == 0 || s == spec.nscales - 1) {
                    const int gy = y0 + rr, gx = x0 + c4;
                    if (gy < h) {
                        float* dst = out + (long long)pl * h * w + (long long)gy * w + gx;
                        if (gx + 3 < w && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
                            __stcs(reinterpret_cast<float4*>(dst), make_float4(res[0], res[1], res[2], res[3]));
                        } else {
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                if (gx + k < w) dst[k] = res[k];
                        }
                    }
                }
            }
            // generic-proxy accesses to the tile / scratch are ordered before the next asynchronous (TMA) write
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
        }
    }
}

// cv2.pyrDown: 5x5 [1 4 6 4 1]/16 separable, BORDER_REFLECT_101, output ((w+1)/2, (h+1)/2); OpenCV's float path:
// row = 6 s[2x] + 4 (s[2x-1] + s[2x+1]) + s[2x-2] + s[2x+2], then the same over rows, times 1/256.
__global__ void __launch_bounds__(256)
k_pyr_down(const float* __restrict__ x, float* __restrict__ out, int planes, int h, int w, int oh, int ow)
{
    const long long total = (long long)planes * oh * ow;
    for (long long it = (long long)blockIdx.x * 256 + threadIdx.x; it < total; it += (long long)gridDim.x * 256) {
        const int pl = int(it / ((long long)oh * ow));
        const int rem = int(it - (long long)pl * oh * ow);
        const int oy = rem / ow, ox = rem - oy * ow;
        const float* src = x + (long long)pl * h * w;
        float rowv[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const float* rp = src + (long long)ext_reflect101(2 * oy - 2 + j, h) * w;
            const float s0 = __ldg(rp + ext_reflect101(2 * ox - 2, w)), s1 = __ldg(rp + ext_reflect101(2 * ox - 1, w));
            const float s2 = __ldg(rp + ext_reflect101(2 * ox, w)), s3 = __ldg(rp + ext_reflect101(2 * ox + 1, w));
            const float s4 = __ldg(rp + ext_reflect101(2 * ox + 2, w));
            rowv[j] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(s2, 6.0f), __fmul_rn(__fadd_rn(s1, s3), 4.0f)), s0), s4);
        }
        const float v = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(rowv[2], 6.0f), __fmul_rn(__fadd_rn(rowv[1], rowv[3]), 4.0f)), rowv[0]), rowv[4]);
        out[it] = __fmul_rn(v, 1.0f / 256.0f);
    }
}

__global__ void __launch_bounds__(256)
k_gamma(const float* __restrict__ x, float* __restrict__ out, long long count, float gamma)
{
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < count; i += (long long)gridDim.x * 256) {
        const float v = fminf(fmaxf(x[i], 0.0f), 1.0f);
        out[i] = powf(v, gamma);
    }
}

// ---- host -------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// cv2.getGaussianKernel(ksize, sigma, CV_32F): fixed tables for ksize <= 7 with sigma <= 0, else exp() in double, normalised
static void gaussian_taps(int ksize, double sigma, float* half /* r+1 */)
{
    static const float small_tab[4][7] = {{1.f}, {0.25f, 0.5f, 0.25f}, {0.0625f, 0.25f, 0.375f, 0.25f, 0.0625f},
                                          {0.03125f, 0.109375f, 0.21875f, 0.28125f, 0.21875f, 0.109375f, 0.03125f}};
    const int r = ksize / 2;
    if (sigma <= 0 && ksize <= 7) {
        for (int d = 0; d <= r; ++d) half[d] = small_tab[ksize >> 1][r + d];
        return;
    }
    const double sg = sigma > 0 ? sigma : ((ksize - 1) * 0.5 - 1) * 0.3 + 0.8;
    const double scale2x = -0.5 / (sg * sg);
    double k[2 * kGtMaxR + 1], sum = 0.0;
    for (int i = 0; i < ksize; ++i) {
        const double xx = i - (ksize - 1) * 0.5;
        k[i] = std::exp(scale2x * xx * xx);
        sum += k[i];
    }
    for (int d = 0; d <= r; ++d) half[d] = float(k[r + d] * (1.0 / sum));
}

static int gauss_launch(const float* x, float* out, int planes, int h, int w, const GaussSpec& spec, cudaStream_t s)
{
    const int tiles_x = (w + kGtTW - 1) / kGtTW, tiles_y = (h + kGtTH - 1) / kGtTH;
    const long long nitems = (long long)planes * tiles_x * tiles_y;
    const size_t smem = size_t(2 * kGtStageFloats + kGtInH * kGtTW + kGtTH * kGtTW) * sizeof(float);
    static std::atomic<unsigned long long> mt{0}, mp{0};
    UPR_CUDA_TRY(ensure_dynamic_smem(k_gauss_tile<true>, smem, mt));
    UPR_CUDA_TRY(ensure_dynamic_smem(k_gauss_tile<false>, smem, mp));
    const int grid = int(std::min<long long>(nitems, 2LL * kNumSMsB200));
    CUtensorMap tmap;
    std::memset(&tmap, 0, sizeof tmap);
    bool tma = (w % 4 == 0) && aligned16(x) && w >= 2 * kGtMaxR + 2 && h >= 2 * kGtMaxR + 2 && encode_tiled_fn() != nullptr;
    if (tma) {
        const cuuint64_t gdim[3] = {cuuint64_t(w), cuuint64_t(h), cuuint64_t(planes)};
        const cuuint64_t gstride[2] = {cuuint64_t(w) * 4, cuuint64_t(w) * cuuint64_t(h) * 4};
        const cuuint32_t box[3] = {cuuint32_t(kGtInW), cuuint32_t(kGtTH + 2 * spec.rmax), 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        const CUresult r = encode_tiled_fn()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(x), gdim, gstride, box, estr,
                                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) tma = false;
    }
    if (tma) k_gauss_tile<true><<<grid, kGtThreads, smem, s>>>(tmap, x, out, planes, h, w, tiles_x, tiles_y, spec);
    else k_gauss_tile<false><<<grid, kGtThreads, smem, s>>>(tmap, x, out, planes, h, w, tiles_x, tiles_y, spec);
    UPR_LAUNCH_CHECK();
    return UPR_OK;
}

}  // namespace upr

extern "C" {

int upr_ext_gaussian_blur_f32(const float* x, float* out, int planes, int h, int w, int ksize, double sigma, upr_stream_t stream)
{
    if (planes < 0 || h <= 0 || w <= 0 || ksize < 1 || (ksize & 1) == 0 || ksize > 2 * upr::kGtMaxR + 1) return UPR_E_SHAPE;
    if (planes == 0) return UPR_OK;
    if (!x || !out || x == out) return UPR_E_NULL;
    upr::GaussSpec spec{};
    spec.nscales = 1; spec.mode = 0; spec.radius[0] = ksize / 2; spec.rmax = ksize / 2; spec.eps = 0.0f;
    upr::gaussian_taps(ksize, sigma, spec.taps[0]);
    return upr::gauss_launch(x, out, planes, h, w, spec, static_cast<cudaStream_t>(stream));
}

int upr_ext_msr_f32(const float* x, float* out, int planes, int h, int w, int nscales, const int* ksizes, const double* sigmas,
                    const float* weights, float eps, upr_stream_t stream)
{
    if (planes < 0 || h <= 0 || w <= 0 || nscales < 1 || nscales > upr::kGtMaxScales) return UPR_E_SHAPE;
    if (planes == 0) return UPR_OK;
    if (!x || !out || !ksizes || !sigmas || !weights || x == out) return UPR_E_NULL;
    upr::GaussSpec spec{};
    spec.nscales = nscales; spec.mode = 1; spec.eps = eps; spec.rmax = 0;
    for (int s = 0; s < nscales; ++s) {
        if (ksizes[s] < 1 || (ksizes[s] & 1) == 0 || ksizes[s] > 2 * upr::kGtMaxR + 1) return UPR_E_PARAM;
        spec.radius[s] = ksizes[s] / 2;
        spec.rmax = std::max(spec.rmax, spec.radius[s]);
        spec.weight[s] = weights[s];
        upr::gaussian_taps(ksizes[s], sigmas[s], spec.taps[s]);
    }
    return upr::gauss_launch(x, out, planes, h, w, spec, static_cast<cudaStream_t>(stream));
}

int upr_ext_pyr_down_f32(const float* x, float* out, int planes, int h, int w, upr_stream_t stream)
{
    if (planes < 0 || h <= 0 || w <= 0) return UPR_E_SHAPE;
    if (planes == 0) return UPR_OK;
    if (!x || !out) return UPR_E_NULL;
    const int oh = (h + 1) / 2, ow = (w + 1) / 2;
    const long long total = (long long)planes * oh * ow;
    const int grid = int(std::min<long long>((total + 255) / 256, 16LL * upr::kNumSMsB200));
    upr::k_pyr_down<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, out, planes, h, w, oh, ow);
    UPR_LAUNCH_CHECK();
    return UPR_OK;
}

int upr_ext_gamma_f32(const float* x, float* out, long long count, float gamma, upr_stream_t stream)
{
    if (count < 0) return UPR_E_SHAPE;
    if (count == 0) return UPR_OK;
    if (!x || !out) return UPR_E_NULL;
    const int grid = int(std::min<long long>((count + 255) / 256, 16LL * upr::kNumSMsB200));
    upr::k_gamma<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, out, count, gamma);
    UPR_LAUNCH_CHECK();
    return UPR_OK;
}

}  // extern "C"
