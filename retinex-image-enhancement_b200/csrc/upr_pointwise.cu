// upr_pointwise.cu -- the three pointwise epilogues of the classical path (sm_100a).
//
//   a8  Retinex decomposition + recombination   /root/reference/models/model.py:405-413, :442
//         R = x / (illu + eps);  enhanced = R*e + (1 - R)*e^2
//   a5  multi-scale gain + clamp                /root/reference/enhancers/multi_scale.py:97-98
//         out = clamp(enhanced * gain[image], 0, 1)
//   a7  attention gain + clamp                  /root/reference/enhancers/content_aware.py:119-120
//         out = clamp(enhanced * (1 + 0.2*att), 0, 1)
//
// All three are pure HBM streams: one 128-bit load per operand per thread, evict-first
// both ways (nothing is re-read), a persistent grid of 148 x kCtasPerSm CTAs walking the
// frame-major index space.  Every product/sum is an explicitly rounded fp32 operation
// (__fmul_rn/__fadd_rn/__fdiv_rn): torch eager evaluates each Python operator as its own
// kernel, so there is no FMA contraction in the reference and none here -> bit-exact.
#include "upr_common.cuh"

namespace upr {

constexpr int kPwThreads = 256;
constexpr int kPwCtasPerSm = 8;

__device__ __forceinline__ float clamp01_keep_nan(float v)
{
    // torch.clamp propagates NaN; fminf/fmaxf would not
    return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
}

__device__ __forceinline__ void recombine1(float x, float d, float e, float& r, float& o)
{
    r = __fdiv_rn(x, d);
    o = __fadd_rn(__fmul_rn(r, e), __fmul_rn(__fsub_rn(1.0f, r), __fmul_rn(e, e)));
}

// ---- a8 ---------------------------------------------------------------------------------
// plane4 = H*W/4 ; one work item = 4 consecutive pixels of one frame (all three channels)
template <bool kWriteRefl, bool kWriteEnh>
__global__ void __launch_bounds__(kPwThreads)
k_recombine_vec(const float* __restrict__ x, const float* __restrict__ illu, const float* __restrict__ e,
                float* __restrict__ refl, float* __restrict__ enh, long long plane4, long long items, float eps)
{
    const uint64_t pol = policy_evict_first();
    const long long stride = (long long)gridDim.x * kPwThreads;
    for (long long it = (long long)blockIdx.x * kPwThreads + threadIdx.x; it < items; it += stride) {
        const long long f = it / plane4, p = it - f * plane4;
        const long long o3 = (f * 3 * plane4 + p) * 4, o1 = (f * plane4 + p) * 4;
        const float4 il = ld_stream_f4(illu + o1, pol);
        const float d[4] = {__fadd_rn(il.x, eps), __fadd_rn(il.y, eps), __fadd_rn(il.z, eps), __fadd_rn(il.w, eps)};
        float4 xv[3], ev[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            xv[c] = ld_stream_f4(x + o3 + c * plane4 * 4, pol);
            if (kWriteEnh) ev[c] = ld_stream_f4(e + o3 + c * plane4 * 4, pol);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float4 r, o;
            if (kWriteEnh) {
                recombine1(xv[c].x, d[0], ev[c].x, r.x, o.x);
                recombine1(xv[c].y, d[1], ev[c].y, r.y, o.y);
                recombine1(xv[c].z, d[2], ev[c].z, r.z, o.z);
                recombine1(xv[c].w, d[3], ev[c].w, r.w, o.w);
                st_stream_f4(enh + o3 + c * plane4 * 4, o, pol);
            } else {
                r.x = __fdiv_rn(xv[c].x, d[0]);
                r.y = __fdiv_rn(xv[c].y, d[1]);
                r.z = __fdiv_rn(xv[c].z, d[2]);
                r.w = __fdiv_rn(xv[c].w, d[3]);
            }
            if (kWriteRefl) st_stream_f4(refl + o3 + c * plane4 * 4, r, pol);
        }
    }
}

template <bool kWriteRefl, bool kWriteEnh>
__global__ void __launch_bounds__(kPwThreads)
k_recombine_scalar(const float* __restrict__ x, const float* __restrict__ illu, const float* __restrict__ e,
                   float* __restrict__ refl, float* __restrict__ enh, long long plane, long long items, float eps)
{
    const long long stride = (long long)gridDim.x * kPwThreads;
    for (long long it = (long long)blockIdx.x * kPwThreads + threadIdx.x; it < items; it += stride) {
        const long long f = it / plane, p = it - f * plane;
        const float d = __fadd_rn(illu[it], eps);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const long long o = (f * 3 + c) * plane + p;
            float r, v;
            if (kWriteEnh) {
                recombine1(x[o], d, e[o], r, v);
                enh[o] = v;
            } else {
                r = __fdiv_rn(x[o], d);
            }
            if (kWriteRefl) refl[o] = r;
        }
    }
}

// ---- a5 / a7 ----------------------------------------------------------------------------
// kPerPixel == false: gain[f] per frame (a5);  true: gain = 1 + 0.2*att[f][p] per pixel (a7)
// One 128-bit load and one store per trip, all CTAs sweeping the batch linearly inside a ~5 MB window: 6.2-6.3 TB/s on 16 - 192 4K
// frames (a device-to-device copy: 6.5-6.6).  Measured and NOT kept (scripts/dev/clamp_probe.py): 2 / 4 loads in flight per thread
// (window x2 / x4) 5.87 / 5.76 TB/s; a (parts, frame) grid without the index divisions but with every frame open at once 5.70 TB/s
// -- this stream is limited by DRAM page locality, not by bytes in flight or by its 64-bit index arithmetic.
template <bool kPerPixel>
__global__ void __launch_bounds__(kPwThreads)
k_gain_clamp_vec(const float* __restrict__ enh, const float* __restrict__ gain, float* __restrict__ out, int c,
                 long long plane4, long long items)
{
    const uint64_t pol = policy_evict_first();
    const long long stride = (long long)gridDim.x * kPwThreads;
    for (long long it = (long long)blockIdx.x * kPwThreads + threadIdx.x; it < items; it += stride) {
        const long long fc = it / plane4, p = it - fc * plane4;
        const long long f = fc / c;
        const float4 v = ld_stream_f4(enh + it * 4, pol);
        float4 g;
        if (kPerPixel) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(gain + (f * plane4 + p) * 4));
            g.x = __fadd_rn(1.0f, __fmul_rn(0.2f, a.x));
            g.y = __fadd_rn(1.0f, __fmul_rn(0.2f, a.y));
            g.z = __fadd_rn(1.0f, __fmul_rn(0.2f, a.z));
            g.w = __fadd_rn(1.0f, __fmul_rn(0.2f, a.w));
        } else {
            const float s = __ldg(gain + f);
            g = make_float4(s, s, s, s);
        }
        float4 o;
        o.x = clamp01_keep_nan(__fmul_rn(v.x, g.x));
        o.y = clamp01_keep_nan(__fmul_rn(v.y, g.y));
        o.z = clamp01_keep_nan(__fmul_rn(v.z, g.z));
        o.w = clamp01_keep_nan(__fmul_rn(v.w, g.w));
        st_stream_f4(out + it * 4, o, pol);
    }
}

template <bool kPerPixel>
__global__ void __launch_bounds__(kPwThreads)
k_gain_clamp_scalar(const float* __restrict__ enh, const float* __restrict__ gain, float* __restrict__ out, int c,
                    long long plane, long long items)
{
    const long long stride = (long long)gridDim.x * kPwThreads;
    for (long long it = (long long)blockIdx.x * kPwThreads + threadIdx.x; it < items; it += stride) {
        const long long fc = it / plane, p = it - fc * plane;
        const long long f = fc / c;
        const float g = kPerPixel ? __fadd_rn(1.0f, __fmul_rn(0.2f, gain[f * plane + p])) : gain[f];
        out[it] = clamp01_keep_nan(__fmul_rn(enh[it], g));
    }
}

static inline int pw_grid(long long items)
{
    const long long want = (items + kPwThreads - 1) / kPwThreads;
    return int(std::max<long long>(1, std::min<long long>(want, (long long)kNumSMsB200 * kPwCtasPerSm)));
}

template <bool R, bool E>
static int recombine_launch(const float* x, const float* illu, const float* e, float* refl, float* enh, int n, int h,
                            int w, float eps, cudaStream_t s)
{
    const long long plane = (long long)h * w;
    const bool vec = plane % 4 == 0 && aligned16(x) && aligned16(illu) && (!E || (aligned16(e) && aligned16(enh))) &&
                     (!R || aligned16(refl));
    if (vec) {
        const long long items = (long long)n * (plane / 4);
        k_recombine_vec<R, E><<<pw_grid(items), kPwThreads, 0, s>>>(x, illu, e, refl, enh, plane / 4, items, eps);
    } else {
        const long long items = (long long)n * plane;
        k_recombine_scalar<R, E><<<pw_grid(items), kPwThreads, 0, s>>>(x, illu, e, refl, enh, plane, items, eps);
    }
    UPR_LAUNCH_CHECK();
    return UPR_OK;
}

template <bool P>
static int gain_launch(const float* enh, const float* gain, float* out, int n, int c, int h, int w, cudaStream_t s)
{
    const long long plane = (long long)h * w;
    const bool vec = plane % 4 == 0 && aligned16(enh) && aligned16(out) && (!P || aligned16(gain));
    if (vec) {
        const long long items = (long long)n * c * (plane / 4);
        k_gain_clamp_vec<P><<<pw_grid(items), kPwThreads, 0, s>>>(enh, gain, out, c, plane / 4, items);
    } else {
        const long long items = (long long)n * c * plane;
        k_gain_clamp_scalar<P><<<pw_grid(items), kPwThreads, 0, s>>>(enh, gain, out, c, plane, items);
    }
    UPR_LAUNCH_CHECK();
    return UPR_OK;
}

int scale_clamp_launch(const float* enh, const float* gain, float* out, int n, int c, int h, int w, cudaStream_t s)
{
    return gain_launch<false>(enh, gain, out, n, c, h, w, s);
}

// ---- save_image quantiser (enhancers/simple_enhance.py:65-100) -----------------------------------
// [n][c][h][w] f32 -> [n][h][w][c] u8 with (clip(x, 0, 1) * 255).astype(uint8): fp32 product, truncation.  NaN clips to NaN and
// casts to 0 on the host; fmaxf(NaN, 0) = 0 gives the same byte here.  c = 1 (illumination maps) or 3 (frames).  One thread
// per 4 pixels: c 128-bit loads, 4c contiguous bytes stored.
template <int C>
__global__ void __launch_bounds__(kPwThreads)
k_quantize_u8(const float* __restrict__ x, unsigned char* __restrict__ out, long long plane, long long items)
{
    const long long plane4 = plane / 4;
    const long long stride = (long long)gridDim.x * kPwThreads;
    for (long long it = (long long)blockIdx.x * kPwThreads + threadIdx.x; it < items; it += stride) {
        const long long f = it / plane4, p = it - f * plane4;
        unsigned q[C][4];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float4 v = __ldcs(reinterpret_cast<const float4*>(x + (f * C + c) * plane) + p);
            const float a[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) q[c][k] = unsigned(__float2int_rz(__fmul_rn(fminf(fmaxf(a[k], 0.0f), 1.0f), 255.0f)));
        }
        unsigned char* o = out + (f * plane + p * 4) * C;
        if (C == 1) {
            *reinterpret_cast<unsigned*>(o) = q[0][0] | (q[0][1] << 8) | (q[0][2] << 16) | (q[0][3] << 24);
        } else {
            unsigned* o32 = reinterpret_cast<unsigned*>(o);   // 12 bytes: R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3
            o32[0] = q[0][0] | (q[1][0] << 8) | (q[2][0] << 16) | (q[0][1] << 24);
            o32[1] = q[1][1] | (q[2][1] << 8) | (q[0][2] << 16) | (q[1][2] << 24);
            o32[2] = q[2][2] | (q[0][3] << 8) | (q[1][3] << 16) | (q[2][3] << 24);
        }
    }
}

template <int C>
__global__ void __launch_bounds__(kPwThreads)
k_quantize_u8_scalar(const float* __restrict__ x, unsigned char* __restrict__ out, long long plane, long long items)
{
    const long long stride = (long long)gridDim.x * kPwThreads;
    for (long long it = (long long)blockIdx.x * kPwThreads + threadIdx.x; it < items; it += stride) {
        const long long f = it / plane, p = it - f * plane;
#pragma unroll
        for (int c = 0; c < C; ++c)
            out[it * C + c] = (unsigned char)__float2int_rz(__fmul_rn(fminf(fmaxf(x[(f * C + c) * plane + p], 0.0f), 1.0f), 255.0f));
    }
}

}  // namespace upr

extern "C" {

int upr_quantize_u8_f32(const float* x_nchw, unsigned char* out_nhwc, int n, int c, int h, int w, upr_stream_t stream)
{
    if (n < 0 || h <= 0 || w <= 0 || (c != 1 && c != 3)) return UPR_E_SHAPE;
    if (n == 0) return UPR_OK;
    if (!x_nchw || !out_nhwc) return UPR_E_NULL;
    auto s = static_cast<cudaStream_t>(stream);
    const long long plane = (long long)h * w;
    if (plane % 4 == 0 && upr::aligned16(x_nchw) && (reinterpret_cast<uintptr_t>(out_nhwc) & 3u) == 0) {
        const long long items = (long long)n * (plane / 4);
        if (c == 1) upr::k_quantize_u8<1><<<upr::pw_grid(items), upr::kPwThreads, 0, s>>>(x_nchw, out_nhwc, plane, items);
        else upr::k_quantize_u8<3><<<upr::pw_grid(items), upr::kPwThreads, 0, s>>>(x_nchw, out_nhwc, plane, items);
    } else {
        const long long items = (long long)n * plane;
        if (c == 1) upr::k_quantize_u8_scalar<1><<<upr::pw_grid(items), upr::kPwThreads, 0, s>>>(x_nchw, out_nhwc, plane, items);
        else upr::k_quantize_u8_scalar<3><<<upr::pw_grid(items), upr::kPwThreads, 0, s>>>(x_nchw, out_nhwc, plane, items);
    }
    UPR_LAUNCH_CHECK();
    return UPR_OK;
}

int upr_retinex_recombine_f32(const float* x, const float* illu, const float* e, float* refl, float* enh, int n, int h,
                              int w, float eps, upr_stream_t stream)
{
    if (n < 0 || h <= 0 || w <= 0) return UPR_E_SHAPE;
    if (n == 0) return UPR_OK;
    if (!x || !illu || !e || !enh) return UPR_E_NULL;
    auto s = static_cast<cudaStream_t>(stream);
    return refl ? upr::recombine_launch<true, true>(x, illu, e, refl, enh, n, h, w, eps, s)
                : upr::recombine_launch<false, true>(x, illu, e, nullptr, enh, n, h, w, eps, s);
}

int upr_retinex_decompose_f32(const float* x, const float* illu, float* refl, int n, int h, int w, float eps,
                              upr_stream_t stream)
{
    if (n < 0 || h <= 0 || w <= 0) return UPR_E_SHAPE;
    if (n == 0) return UPR_OK;
    if (!x || !illu || !refl) return UPR_E_NULL;
    return upr::recombine_launch<true, false>(x, illu, nullptr, refl, nullptr, n, h, w, eps,
                                              static_cast<cudaStream_t>(stream));
}

int upr_scale_clamp_f32(const float* enh, const float* gain_per_image, float* out, int n, int c, int h, int w,
                        upr_stream_t stream)
{
    if (n < 0 || c <= 0 || h <= 0 || w <= 0) return UPR_E_SHAPE;
    if (n == 0) return UPR_OK;
    if (!enh || !gain_per_image || !out) return UPR_E_NULL;
    return upr::gain_launch<false>(enh, gain_per_image, out, n, c, h, w, static_cast<cudaStream_t>(stream));
}

int upr_attention_apply_f32(const float* enh, const float* att, float* out, int n, int c, int h, int w,
                            upr_stream_t stream)
{
    if (n < 0 || c <= 0 || h <= 0 || w <= 0) return UPR_E_SHAPE;
    if (n == 0) return UPR_OK;
    if (!enh || !att || !out) return UPR_E_NULL;
    return upr::gain_launch<true>(enh, att, out, n, c, h, w, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
