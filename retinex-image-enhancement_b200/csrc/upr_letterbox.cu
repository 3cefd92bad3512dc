// upr_letterbox.cu -- YOLO-style letterbox of the enhance drivers' --max_size path on the GPU.
//
// Replaces utils/letterbox.py:9-102 as called from enhancers/simple_enhance.py:43-58 (scaleup=False, auto=True):
//   u8 = (x*255).astype(uint8)                      numpy cast semantics (wrap, NaN -> 0; see quantize_u8)
//   cv2.resize(u8, (rw, rh), INTER_LINEAR)          when the size changes: OpenCV's 8-bit fixed-point bilinear
//                                                   (11-bit coefficients; rows at 2^11 scale, columns via >>4, >>16, +2, >>2)
//   cv2.copyMakeBorder(top, bottom, left, right, BORDER_CONSTANT, 114)
//   .astype(float32) / 255
// The fixed-point recipe was pinned against the cv2 4.13 binary (IPP on and off) for DOWN-scaling, which is all the
// drivers ever ask for (scaleup=False): 0 mismatches on 16 shape pairs (tests/test_oracle_pin.py).  cv2's up-scaling
// differs from this recipe by 1 LSB on ~0.04 % of the pixels, so up-scaling is refused here (UPR_E_PARAM) and stays on
// the host.  Input is either the reference's f32 CHW tensor or the decoded file itself (u8 HWC, 3 B/px over PCIe
// instead of 12).
#include <algorithm>

#include "upr_common.cuh"

namespace upr {

struct LbGeom {
    int n, c, h, w;          // source
    int rh, rw;              // resized (unpadded) size
    int top, left, oh, ow;   // placement in the output
    double scale_x, scale_y; // w / rw, h / rh
    float pad[4];            // border value / 255 per channel
    int resize;
};

__device__ __forceinline__ void lb_coeff(int d, double scale, int sn, int& s, int& a0, int& a1)
{
    float f = float((d + 0.5) * scale - 0.5);   // OpenCV: double product, stored as float
    s = int(floorf(f));
    f = __fsub_rn(f, float(s));
    if (s < 0) { f = 0.0f; s = 0; }
    if (s >= sn - 1) { f = 0.0f; s = sn - 1; }
    a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.0f, f), 2048.0f));   // saturate_cast<short>(cvRound(.)): |v| <= 2048
    a1 = __float2int_rn(__fmul_rn(f, 2048.0f));
}

template <bool kU8Hwc>
__device__ __forceinline__ int lb_px(const void* in, const LbGeom& g, int f, int ch, int y, int x)
{
    if (kU8Hwc) return static_cast<const unsigned char*>(in)[((size_t(f) * g.h + y) * g.w + x) * g.c + ch];
    return quantize_u8(__ldg(static_cast<const float*>(in) + ((size_t(f) * g.c + ch) * g.h + y) * g.w + x));
}

template <bool kU8Hwc>
__global__ void __launch_bounds__(256)
k_letterbox(const void* __restrict__ in, float* __restrict__ out, const LbGeom g)
{
    const long long total = (long long)g.n * g.oh * g.ow;
    for (long long it = (long long)blockIdx.x * 256 + threadIdx.x; it < total; it += (long long)gridDim.x * 256) {
        const int f = int(it / ((long long)g.oh * g.ow));
        const int rem = int(it - (long long)f * g.oh * g.ow);
        const int oy = rem / g.ow, ox = rem - oy * g.ow;
        const int iy = oy - g.top, ix = ox - g.left;
        float* o = out + (size_t(f) * g.c * g.oh + oy) * g.ow + ox;
        const size_t oplane = size_t(g.oh) * g.ow;
        if (iy < 0 || iy >= g.rh || ix < 0 || ix >= g.rw) {
            for (int ch = 0; ch < g.c; ++ch) o[ch * oplane] = g.pad[min(ch, 3)];
            continue;
        }
        if (!g.resize) {
            for (int ch = 0; ch < g.c; ++ch) o[ch * oplane] = __fdiv_rn(float(lb_px<kU8Hwc>(in, g, f, ch, iy, ix)), 255.0f);
            continue;
        }
        int sx, a0, a1, sy, b0, b1;
        lb_coeff(ix, g.scale_x, g.w, sx, a0, a1);
        lb_coeff(iy, g.scale_y, g.h, sy, b0, b1);
        const int x1 = min(sx + 1, g.w - 1), y1 = min(sy + 1, g.h - 1);
        for (int ch = 0; ch < g.c; ++ch) {
            const int s0 = lb_px<kU8Hwc>(in, g, f, ch, sy, sx) * a0 + lb_px<kU8Hwc>(in, g, f, ch, sy, x1) * a1;
            const int s1 = lb_px<kU8Hwc>(in, g, f, ch, y1, sx) * a0 + lb_px<kU8Hwc>(in, g, f, ch, y1, x1) * a1;
            int v = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2;
            v = min(max(v, 0), 255);
            o[ch * oplane] = __fdiv_rn(float(v), 255.0f);
        }
    }
}

static int lb_run(bool u8, const void* in, float* out, int n, int c, int h, int w, int rh, int rw, int top, int left, int oh,
                  int ow, const unsigned char* pad, cudaStream_t s)
{
    if (n < 0 || c <= 0 || c > 4 || h <= 0 || w <= 0 || rh <= 0 || rw <= 0 || top < 0 || left < 0 || oh < top + rh || ow < left + rw)
        return UPR_E_SHAPE;
    if (rh > h || rw > w) return UPR_E_PARAM;   // up-scaling: not bit-exact with this recipe, stays on the host
    if (n == 0) return UPR_OK;
    if (!in || !out) return UPR_E_NULL;
    LbGeom g{};
    g.n = n; g.c = c; g.h = h; g.w = w; g.rh = rh; g.rw = rw; g.top = top; g.left = left; g.oh = oh; g.ow = ow;
    g.scale_x = double(w) / rw; g.scale_y = double(h) / rh;
    g.resize = (rh != h || rw != w);
    for (int i = 0; i < 4; ++i) g.pad[i] = float(pad ? pad[std::min(i, c - 1)] : 114) / 255.0f;
    const long long total = (long long)n * oh * ow;
    const int grid = int(std::min<long long>((total + 255) / 256, 16LL * kNumSMsB200));
    if (u8) k_letterbox<true><<<grid, 256, 0, s>>>(in, out, g);
    else k_letterbox<false><<<grid, 256, 0, s>>>(in, out, g);
    UPR_LAUNCH_CHECK();
    return UPR_OK;
}

}  // namespace upr

extern "C" {

int upr_letterbox_f32(const float* in_nchw, float* out_nchw, int n, int c, int h, int w, int rh, int rw, int top, int left,
                      int oh, int ow, const unsigned char* pad_value, upr_stream_t stream)
{
    return upr::lb_run(false, in_nchw, out_nchw, n, c, h, w, rh, rw, top, left, oh, ow, pad_value, static_cast<cudaStream_t>(stream));
}

int upr_letterbox_u8_f32(const unsigned char* in_nhwc, float* out_nchw, int n, int c, int h, int w, int rh, int rw, int top,
                         int left, int oh, int ow, const unsigned char* pad_value, upr_stream_t stream)
{
    return upr::lb_run(true, in_nhwc, out_nchw, n, c, h, w, rh, rw, top, left, oh, ow, pad_value, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
