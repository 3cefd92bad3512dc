// upr_stats.cu -- per-image statistics kernels (sm_100a).
//
//   a3   brightness histogram       /root/reference/enhancers/adaptive_params.py:24-68
//          u8 gray (OpenCV BGR2GRAY fixed point) of the trunc-quantised image, 256 bins.
//          mean/std/dark/mid/bright ratios are exact functions of this histogram (host side).
//   a9   texture complexity         /root/reference/losses/loss.py:523-583
//          'tv'           mean|dx| + mean|dy| per image
//          'edge_density' frac(Sobel magnitude > 1.5 * mean magnitude), gray = channel mean,
//                         reflect(-101) padding
//   a10  dynamic smoothness weight  /root/reference/losses/loss.py:704-720
//          w = clamp(w0 * (1 - 0.8 * mean_B(c)), 0.1, 5.0)   from the (all-reduced) [sum c, B]
//
// Reductions: fp64 per-thread accumulators -> warp shuffles -> one partial per CTA written
// to the workspace -> the last CTA of an image (ticket) adds the partials in index order.
// No floating-point atomics: results are deterministic run to run.
#include <algorithm>

#include "upr_common.cuh"

namespace upr {

constexpr int kStThreads = 256;

__device__ __forceinline__ double block_sum(double v, double* s_red)
{
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_red[wid] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < kStThreads / 32; ++i) t += s_red[i];
    return t;
}

// true (block-uniform) in exactly one CTA per image: the one that arrived last
__device__ __forceinline__ bool last_cta_of(unsigned* ticket, unsigned nparts, int* s_flag)
{
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(ticket, 1u);
        *s_flag = (t == nparts - 1);
        if (t == nparts - 1) *ticket = 0;  // self-cleaning: ready for the next call on this workspace
    }
    __syncthreads();
    const bool last = *s_flag != 0;
    if (last) __threadfence();
    return last;
}

// Data-parallel exchange of the batch statistics over NVLink peer memory, fused into the statistics kernel (the
// "compute step followed by a collective" of this path: losses/loss.py:710 needs the batch mean over ALL ranks).
// Every rank owns one symmetric-memory buffer (torch.distributed._symmetric_memory; peers[] holds the W mapped base
// addresses, identical layout everywhere):
//     float    slot[2][kPeerMax][2]   [parity][source rank] = {sum of complexity, image count}
//     unsigned flag[2][kPeerMax]      [parity][source rank] = sequence number of the call that wrote the slot
// The thread that finishes the local batch stores its pair into slot[seq & 1][rank] of EVERY rank (P2P stores), fences
// system-wide, raises the flags, then waits for the W flags of its own buffer and adds the W pairs in rank order -- all
// ranks add the same numbers in the same order, so every rank derives the bit-identical weight.  Two parities: a rank can
// be at most one call ahead of its slowest peer (it needs that peer's flag of the current call to finish).
// A peer that does not arrive within the time-out (wall clock, %globaltimer; default 10 minutes like NCCL's, settable with
// upr_peer_set_timeout_ms) does NOT trap the kernel: the waiting rank records {seq, missing rank} in the status word at the end of
// its own buffer, returns NaN statistics / weight (which poison the loss visibly) and the host raises from upr_peer_status /
// DynamicSmoothWeight.check_peers().  Every rank must call once per step, in lockstep.
constexpr int kPeerMax = 64;
constexpr size_t kPeerDataBytes = size_t(2) * kPeerMax * 2 * sizeof(float) + size_t(2) * kPeerMax * sizeof(unsigned);
constexpr size_t kPeerStatusOff = (kPeerDataBytes + 15) / 16 * 16;     // unsigned[2]: {seq of the failed call, rank waited for}
constexpr size_t kPeerBufBytes = kPeerStatusOff + 16;
static unsigned long long g_peer_timeout_ns = 600ull * 1000ull * 1000ull * 1000ull;

struct PeerXchg {
    const unsigned long long* peers;   // device array [world] of peer buffer addresses; nullptr = single process
    int rank, world;
    unsigned seq;                      // >= 1, strictly increasing per call on this buffer set
    float w0;                          // weight_smooth
    float* weight_out;                 // device scalar, may be nullptr
    unsigned long long timeout_ns;     // wall-clock bound on the wait for a peer's flag
};

__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ float dyn_weight(float sum, float count, float w0)
{
    const float avg = __fdiv_rn(sum, count);
    const float wv = __fmul_rn(w0, __fsub_rn(1.0f, __fmul_rn(avg, 0.8f)));
    return wv < 0.1f ? 0.1f : (wv > 5.0f ? 5.0f : wv);
}

__device__ __forceinline__ void peer_exchange(const PeerXchg& px, float& s, float& cnt)
{
    const int par = int(px.seq & 1u);
    for (int r = 0; r < px.world; ++r) {
        volatile float* slot = reinterpret_cast<volatile float*>(px.peers[r]) + (size_t(par) * kPeerMax + px.rank) * 2;
        slot[0] = s;
        slot[1] = cnt;
    }
    __threadfence_system();
    for (int r = 0; r < px.world; ++r) {
        volatile unsigned* flag = reinterpret_cast<volatile unsigned*>(px.peers[r] + size_t(2) * kPeerMax * 2 * sizeof(float)) +
                                  size_t(par) * kPeerMax + px.rank;
        *flag = px.seq;
    }
    const unsigned long long mine = px.peers[px.rank];
    volatile unsigned* myflags = reinterpret_cast<volatile unsigned*>(mine + size_t(2) * kPeerMax * 2 * sizeof(float)) + size_t(par) * kPeerMax;
    const unsigned long long t_start = global_timer_ns();
    for (int r = 0; r < px.world; ++r) {
        unsigned spins = 0;
        while (myflags[r] != px.seq) {
            if ((++spins & 1023u) == 0 && global_timer_ns() - t_start > px.timeout_ns) {
                // a peer that never arrives (skipped step, crashed rank) must neither hang nor fault the device
                volatile unsigned* status = reinterpret_cast<volatile unsigned*>(mine + kPeerStatusOff);
                status[1] = unsigned(r);
                status[0] = px.seq;
                s = cnt = __int_as_float(0x7fc00000);
                return;
            }
        }
    }
    __threadfence_system();
    volatile float* myslots = reinterpret_cast<volatile float*>(mine) + size_t(par) * kPeerMax * 2;
    float ts = 0.0f, tc = 0.0f;
    for (int r = 0; r < px.world; ++r) {
        ts = __fadd_rn(ts, myslots[2 * r]);
        tc = __fadd_rn(tc, myslots[2 * r + 1]);
    }
    s = ts;
    cnt = tc;
}

// Called by thread 0 of the CTA that finished image f: the image that finishes LAST adds the
// per-image values in index order (deterministic) into stats2 = [sum of complexity, image count],
// the two numbers the data-parallel all-reduce carries (loss.py:710 batch mean).  With a peer table the all-reduce
// happens right here and the dynamic smoothness weight is written too.
__device__ __forceinline__ void finish_batch(unsigned* batch_ticket, int n, const float* per_image, float* stats2,
                                             const PeerXchg px = PeerXchg{nullptr, 0, 1, 0u, 0.0f, nullptr, 0ull})
{
    if (!stats2) return;
    __threadfence();
    const unsigned t = atomicAdd(batch_ticket, 1u);
    if (t != unsigned(n - 1)) return;
    *batch_ticket = 0;
    __threadfence();
    float s = 0.0f;
    for (int i = 0; i < n; ++i) s = __fadd_rn(s, __ldcg(per_image + i));
    float cnt = float(n);
    if (px.peers) peer_exchange(px, s, cnt);
    stats2[0] = s;
    stats2[1] = cnt;
    if (px.weight_out) px.weight_out[0] = dyn_weight(s, cnt, px.w0);
}

// -----------------------------------------------------------------------------------------
// a3: brightness histogram.  grid = (parts, n); per-warp private shared histograms.
// -----------------------------------------------------------------------------------------
__device__ __forceinline__ int gray_u8(int r, int g, int b) { return (r * 9798 + g * 19235 + b * 3735 + 16384) >> 15; }

__global__ void __launch_bounds__(kStThreads)
k_brightness_hist(const float* __restrict__ in, unsigned* __restrict__ hist, long long plane)
{
    __shared__ unsigned s_h[kStThreads / 32][256];
    const int tid = threadIdx.x, wid = tid >> 5;
    for (int i = tid; i < (kStThreads / 32) * 256; i += kStThreads) (&s_h[0][0])[i] = 0;
    __syncthreads();
    const float* R = in + (long long)blockIdx.y * 3 * plane;
    const uint64_t pol = policy_evict_first();
    const long long stride = (long long)gridDim.x * kStThreads;
    if (plane % 4 == 0 && aligned16(R)) {
        const long long p4 = plane / 4;
        for (long long i = (long long)blockIdx.x * kStThreads + tid; i < p4; i += stride) {
            const float4 r = ld_stream_f4(R + i * 4, pol), g = ld_stream_f4(R + plane + i * 4, pol),
                         b = ld_stream_f4(R + 2 * plane + i * 4, pol);
            atomicAdd(&s_h[wid][gray_u8(quantize_u8(r.x), quantize_u8(g.x), quantize_u8(b.x))], 1u);
            atomicAdd(&s_h[wid][gray_u8(quantize_u8(r.y), quantize_u8(g.y), quantize_u8(b.y))], 1u);
            atomicAdd(&s_h[wid][gray_u8(quantize_u8(r.z), quantize_u8(g.z), quantize_u8(b.z))], 1u);
            atomicAdd(&s_h[wid][gray_u8(quantize_u8(r.w), quantize_u8(g.w), quantize_u8(b.w))], 1u);
        }
    } else {
        for (long long i = (long long)blockIdx.x * kStThreads + tid; i < plane; i += stride)
            atomicAdd(&s_h[wid][gray_u8(quantize_u8(R[i]), quantize_u8(R[plane + i]), quantize_u8(R[2 * plane + i]))], 1u);
    }
    __syncthreads();
    unsigned t = 0;
#pragma unroll
    for (int k = 0; k < kStThreads / 32; ++k) t += s_h[k][tid];
    if (t) atomicAdd(&hist[(long long)blockIdx.y * 256 + tid], t);
}

// -----------------------------------------------------------------------------------------
// a9 'tv'.  grid = (parts, n).  A work item is one row segment of 4 pixels of one channel.
// -----------------------------------------------------------------------------------------
// workspace: partial [n][parts][2] fp64, tickets [n] + 1 batch ticket, mean magnitude [n] fp64
__global__ void __launch_bounds__(kStThreads)
k_texture_tv(const float* __restrict__ x, int c, int h, int w, double* __restrict__ partial,
             unsigned* __restrict__ tickets, float* __restrict__ per_image, float* __restrict__ stats2, const PeerXchg px)
{
    __shared__ double s_red[kStThreads / 32];
    __shared__ int s_flag;
    const int tid = threadIdx.x;
    const int f = blockIdx.y, parts = gridDim.x;
    const long long plane = (long long)h * w;
    const float* img = x + (long long)f * c * plane;
    double sh = 0.0, sv = 0.0;
    const long long rows = (long long)c * h;  // (channel,row) pairs
    const long long stride = (long long)parts * kStThreads;
    if (w % 4 == 0 && aligned16(img)) {
        const int w4 = w / 4;
        const long long items = rows * w4;
        for (long long it = (long long)blockIdx.x * kStThreads + tid; it < items; it += stride) {
            const long long row = it / w4;
            const int q = int(it - row * w4);
            const int y = int(row % h);
            const float* p = img + row * w + q * 4;
            const float4 a = __ldg(reinterpret_cast<const float4*>(p));
            float hsum = fabsf(a.x - a.y) + fabsf(a.y - a.z) + fabsf(a.z - a.w);
            if (q + 1 < w4) hsum += fabsf(a.w - __ldg(p + 4));
            sh += double(hsum);
            if (y + 1 < h) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(p + w));
                sv += double(fabsf(a.x - b.x) + fabsf(a.y - b.y) + fabsf(a.z - b.z) + fabsf(a.w - b.w));
            }
        }
    } else {
        const long long items = rows * w;
        for (long long it = (long long)blockIdx.x * kStThreads + tid; it < items; it += stride) {
            const long long row = it / w;
            const int xx = int(it - row * w);
            const int y = int(row % h);
            const float a = img[it];
            if (xx + 1 < w) sh += double(fabsf(a - img[it + 1]));
            if (y + 1 < h) sv += double(fabsf(a - img[it + w]));
        }
    }
    sh = block_sum(sh, s_red);
    sv = block_sum(sv, s_red);
    if (tid == 0) {
        partial[((long long)f * parts + blockIdx.x) * 2 + 0] = sh;
        partial[((long long)f * parts + blockIdx.x) * 2 + 1] = sv;
    }
    if (!last_cta_of(tickets + f, parts, &s_flag)) return;
    if (tid == 0) {
        double th = 0.0, tv = 0.0;
        for (int k = 0; k < parts; ++k) {
            th += __ldcg(&partial[((long long)f * parts + k) * 2 + 0]);
            tv += __ldcg(&partial[((long long)f * parts + k) * 2 + 1]);
        }
        // torch.mean of an empty slice is NaN (w == 1 or h == 1): 0/0 reproduces that
        const float mh = float(th / (double(c) * h * (w - 1)));
        const float mv = float(tv / (double(c) * (h - 1) * w));
        const float cx = __fadd_rn(mh, mv);
        per_image[f] = cx;
        finish_batch(tickets + gridDim.y, gridDim.y, per_image, stats2, px);
    }
}

// -----------------------------------------------------------------------------------------
// a9 'edge_density'.  Pass 1 sums the Sobel magnitude, pass 2 recomputes it with the same
// device function (bit-identical) and counts pixels above 1.5 * mean.
// -----------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect101_dev(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

__device__ __forceinline__ float gray_mean(const float* __restrict__ img, long long plane, long long off, int c)
{
    if (c == 1) return __ldg(img + off);
    float s = 0.0f;
    for (int ch = 0; ch < c; ++ch) s = __fadd_rn(s, __ldg(img + ch * plane + off));
    return __fdiv_rn(s, float(c));
}

__device__ __forceinline__ float sobel_mag(const float* __restrict__ img, long long plane, int c, int h, int w, int y, int x)
{
    const int ym = reflect101_dev(y - 1, h), yp = reflect101_dev(y + 1, h);
    const int xm = reflect101_dev(x - 1, w), xp = reflect101_dev(x + 1, w);
    const float a = gray_mean(img, plane, (long long)ym * w + xm, c), b = gray_mean(img, plane, (long long)ym * w + x, c),
                cc = gray_mean(img, plane, (long long)ym * w + xp, c);
    const float d = gray_mean(img, plane, (long long)y * w + xm, c), f = gray_mean(img, plane, (long long)y * w + xp, c);
    const float g = gray_mean(img, plane, (long long)yp * w + xm, c), hh = gray_mean(img, plane, (long long)yp * w + x, c),
                k = gray_mean(img, plane, (long long)yp * w + xp, c);
    const float gx = __fadd_rn(__fadd_rn(__fsub_rn(cc, a), __fmul_rn(2.0f, __fsub_rn(f, d))), __fsub_rn(k, g));
    const float gy = __fadd_rn(__fadd_rn(__fsub_rn(g, a), __fmul_rn(2.0f, __fsub_rn(hh, b))), __fsub_rn(k, cc));
    return __fsqrt_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)));
}

template <int kPass>
__global__ void __launch_bounds__(kStThreads)
k_texture_edge(const float* __restrict__ x, int c, int h, int w, double* __restrict__ partial,
               unsigned* __restrict__ tickets, double* __restrict__ mean_mag, float* __restrict__ per_image,
               float* __restrict__ stats2, const PeerXchg px)
{
    __shared__ double s_red[kStThreads / 32];
    __shared__ int s_flag;
    const int tid = threadIdx.x;
    const int f = blockIdx.y, parts = gridDim.x;
    const long long plane = (long long)h * w;
    const float* img = x + (long long)f * c * plane;
    float thr = 0.0f;
    if (kPass == 2) thr = __fmul_rn(float(__ldcg(&mean_mag[f])), 1.5f);
    double acc = 0.0;
    const long long stride = (long long)parts * kStThreads;
    for (long long it = (long long)blockIdx.x * kStThreads + tid; it < plane; it += stride) {
        const int y = int(it / w), xx = int(it - (long long)y * w);
        const float m = sobel_mag(img, plane, c, h, w, y, xx);
        acc += (kPass == 1) ? double(m) : (m > thr ? 1.0 : 0.0);
    }
    acc = block_sum(acc, s_red);
    if (tid == 0) partial[(long long)f * parts + blockIdx.x] = acc;
    if (!last_cta_of(tickets + f, parts, &s_flag)) return;
    if (tid == 0) {
        double t = 0.0;
        for (int k = 0; k < parts; ++k) t += __ldcg(&partial[(long long)f * parts + k]);
        if (kPass == 1) {
            mean_mag[f] = double(float(t / double(plane)));  // torch.mean result is fp32
        } else {
            const float cx = float(t / double(plane));
            per_image[f] = cx;
            finish_batch(tickets + gridDim.y, gridDim.y, per_image, stats2, px);
        }
    }
}

// -----------------------------------------------------------------------------------------
// N3: EdgeAwareSmoothnessLoss.forward (losses/loss.py:136-176), the term the dynamic weight of a10 multiplies
// (:724), with its gradient w.r.t. the illumination map.  Only illu_map carries a gradient; everything derived from
// img_low is a no-grad image statistic:
//     wh = exp(-lambda * mean_c |S[..., :-1] - S[..., 1:]|)          [B,1,H,W-1]     (wv likewise, [B,1,H-1,W])
//     E  = Sobel magnitude of mean_c S, reflect padding              [B,1,H,W]       (sobel_mag above: same as a9)
//     fh = 1 + alpha * mean(E[..., :W-1], dim=-1)                    [B,1,H,1]       avg_pool2d((1,W-1), stride 1)[..., :-1]
//     fv = 1 + alpha * mean(E[..., :H-1, :], dim=-2)                 [B,1,1,W]
//     loss = mean(wh * fh * |I[..., :-1] - I[..., 1:]|) + mean(wv * fv * |I[..., :-1, :] - I[..., 1:, :]|)
// (yes: the reference's pooling windows make the edge factors ROW and COLUMN means of the edge map).
// Three launches: edge map + row means, column means, loss + d loss / d I in one pass over S and I.
// -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kStThreads)
k_smooth_edge_rows(const float* __restrict__ s_img, int cs, int h, int w, float* __restrict__ edge, float* __restrict__ rowmean)
{
    // one warp per image row
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int y = blockIdx.x * (kStThreads / 32) + wid, f = blockIdx.y;
    if (y >= h) return;
    const long long plane = (long long)h * w;
    const float* img = s_img + (long long)f * cs * plane;
    float* e = edge + (long long)f * plane + (long long)y * w;
    double acc = 0.0;
    for (int x = lane; x < w; x += 32) {
        const float m = sobel_mag(img, plane, cs, h, w, y, x);
        e[x] = m;
        if (x < w - 1) acc += double(m);
    }
    acc = warp_sum(acc);
    if (lane == 0) rowmean[(long long)f * h + y] = float(acc / double(w - 1));
}

__global__ void __launch_bounds__(kStThreads)
k_smooth_edge_cols(const float* __restrict__ edge, int h, int w, float* __restrict__ colmean)
{
    const int x = blockIdx.x * kStThreads + threadIdx.x, f = blockIdx.y;
    if (x >= w) return;
    const float* e = edge + (long long)f * h * w + x;
    double acc = 0.0;
    for (int y = 0; y < h - 1; ++y) acc += double(__ldg(e + (long long)y * w));
    colmean[(long long)f * w + x] = float(acc / double(h - 1));
}

__device__ __forceinline__ float sgnf(float v) { return v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : 0.0f); }

// exp(-lambda * mean_c |S(p) - S(q)|)
__device__ __forceinline__ float smooth_weight(const float* __restrict__ img, long long plane, int cs, long long p, long long q, float lambda)
{
    float s = 0.0f;
    for (int c = 0; c < cs; ++c) s = __fadd_rn(s, fabsf(__fsub_rn(__ldg(img + c * plane + p), __ldg(img + c * plane + q))));
    return expf(__fmul_rn(-lambda, __fdiv_rn(s, float(cs))));
}

__global__ void __launch_bounds__(kStThreads)
k_smooth_loss(const float* __restrict__ illu, const float* __restrict__ s_img, int n, int ci, int cs, int h, int w, float lambda,
              float alpha, const float* __restrict__ rowmean, const float* __restrict__ colmean, float* __restrict__ grad,
              double* __restrict__ partial, unsigned* __restrict__ tickets, float* __restrict__ loss3)
{
    __shared__ double s_red[kStThreads / 32];
    __shared__ int s_flag;
    const int tid = threadIdx.x;
    const int f = blockIdx.y, parts = gridDim.x;
    const long long plane = (long long)h * w;
    const float* img = s_img + (long long)f * cs * plane;
    const float* il = illu + (long long)f * ci * plane;
    float* gr = grad ? grad + (long long)f * ci * plane : nullptr;
    const double inv_nh = 1.0 / (double(n) * ci * h * (w - 1)), inv_nv = 1.0 / (double(n) * ci * (h - 1) * w);
    const float g_h = float(inv_nh), g_v = float(inv_nv);
    double sh = 0.0, sv = 0.0;
    const long long stride = (long long)parts * kStThreads;
    for (long long p = (long long)blockIdx.x * kStThreads + tid; p < plane; p += stride) {
        const int y = int(p / w), x = int(p - (long long)y * w);
        const float fh = __fadd_rn(1.0f, __fmul_rn(alpha, __ldg(rowmean + (long long)f * h + y)));
        const float fv = __fadd_rn(1.0f, __fmul_rn(alpha, __ldg(colmean + (long long)f * w + x)));
        // the four edges that touch this pixel: right (y,x), left (y,x-1), down (y,x), up (y-1,x)
        const bool has_r = x + 1 < w, has_l = x > 0, has_d = y + 1 < h, has_u = y > 0;
        const float wr = has_r ? __fmul_rn(smooth_weight(img, plane, cs, p, p + 1, lambda), fh) : 0.0f;
        const float wd = has_d ? __fmul_rn(smooth_weight(img, plane, cs, p, p + w, lambda), fv) : 0.0f;
        float wl = 0.0f, wu = 0.0f;
        if (gr) {
            wl = has_l ? __fmul_rn(smooth_weight(img, plane, cs, p - 1, p, lambda), fh) : 0.0f;
            wu = has_u ? __fmul_rn(smooth_weight(img, plane, cs, p - w, p, lambda), fv) : 0.0f;
        }
        for (int c = 0; c < ci; ++c) {
            const float* ic = il + c * plane;
            const float v = __ldg(ic + p);
            const float dr = has_r ? __fsub_rn(v, __ldg(ic + p + 1)) : 0.0f;
            const float dd = has_d ? __fsub_rn(v, __ldg(ic + p + w)) : 0.0f;
            sh += double(__fmul_rn(wr, fabsf(dr)));
            sv += double(__fmul_rn(wd, fabsf(dd)));
            if (gr) {
                const float dl = has_l ? __fsub_rn(__ldg(ic + p - 1), v) : 0.0f;
                const float du = has_u ? __fsub_rn(__ldg(ic + p - w), v) : 0.0f;
                const float gh = __fsub_rn(__fmul_rn(wr, sgnf(dr)), __fmul_rn(wl, sgnf(dl)));
                const float gv = __fsub_rn(__fmul_rn(wd, sgnf(dd)), __fmul_rn(wu, sgnf(du)));
                gr[c * plane + p] = __fadd_rn(__fmul_rn(gh, g_h), __fmul_rn(gv, g_v));
            }
        }
    }
    sh = block_sum(sh, s_red);
    sv = block_sum(sv, s_red);
    if (tid == 0) {
        partial[((long long)f * parts + blockIdx.x) * 2 + 0] = sh;
        partial[((long long)f * parts + blockIdx.x) * 2 + 1] = sv;
    }
    if (!last_cta_of(tickets + f, parts, &s_flag)) return;
    if (tid == 0) {
        // image totals into the first slot of the image (ordered sum), then the image that finishes last adds the images in order
        double th = 0.0, tv = 0.0;
        for (int k = 0; k < parts; ++k) {
            th += __ldcg(&partial[((long long)f * parts + k) * 2 + 0]);
            tv += __ldcg(&partial[((long long)f * parts + k) * 2 + 1]);
        }
        partial[(long long)f * parts * 2 + 0] = th;
        partial[(long long)f * parts * 2 + 1] = tv;
        __threadfence();
        unsigned* batch_ticket = tickets + n;
        if (atomicAdd(batch_ticket, 1u) != unsigned(n - 1)) return;
        *batch_ticket = 0;
        __threadfence();
        double ah = 0.0, av = 0.0;
        for (int i = 0; i < n; ++i) {
            ah += __ldcg(&partial[(long long)i * parts * 2 + 0]);
            av += __ldcg(&partial[(long long)i * parts * 2 + 1]);
        }
        const float lh = float(ah * inv_nh), lv = float(av * inv_nv);
        loss3[0] = __fadd_rn(lh, lv);
        loss3[1] = lh;
        loss3[2] = lv;
    }
}

// -----------------------------------------------------------------------------------------
// N3, second part: the three statistics losses of the ENHANCED image in one read of (enhanced, low):
//   exposure  (AdaptiveExposureLoss, losses/loss.py:29-58):  mean |avg_pool16(mean_c R) - E|,  E = base + (0.8 - base)(1 - mean(mean_c S))
//   colour    (ColorLoss, :351-368):                         (mr-mg)^2 + (mr-mb)^2 + (mg-mb)^2 of the batch channel means of R
//   spatial   (SpatialConsistencyLoss, :404-427):            mean((dx R - dx S)^2) + mean((dy R - dy S)^2)
// Forward: one pass for the sums, one warp per 16x16 patch for the patch means, a one-CTA finish.  Backward: one pass that
// forms  u_exp * d exp + u_col * d col + u_spa * d spa  per pixel from the saved statistics (u = upstream gradients).
// saved[0..7] = {mr, mg, mb, mean(gray S), E, dx residual scale 2/Nh, dy residual scale 2/Nv, 1/(M*P*P*3)}, then the patch means.
// -----------------------------------------------------------------------------------------
constexpr int kEnhSums = 6;   // sum R, sum G, sum B (enhanced), sum of mean_c S, sum (dx R - dx S)^2, sum (dy R - dy S)^2

__global__ void __launch_bounds__(kStThreads)
k_enh_sums(const float* __restrict__ enh, const float* __restrict__ low, int n, int h, int w, double* __restrict__ partial,
           unsigned* __restrict__ tickets, float base_target, float target_coef, int hp, int wp, int patch, float* __restrict__ saved,
           float* __restrict__ losses3)
{
    __shared__ double s_red[kStThreads / 32];
    __shared__ int s_flag;
    const int tid = threadIdx.x;
    const int f = blockIdx.y, parts = gridDim.x;
    const long long plane = (long long)h * w;
    const float* e = enh + (long long)f * 3 * plane;
    const float* l = low + (long long)f * 3 * plane;
    double acc[kEnhSums] = {0, 0, 0, 0, 0, 0};
    const long long stride = (long long)parts * kStThreads;
    for (long long p = (long long)blockIdx.x * kStThreads + tid; p < plane; p += stride) {
        const int y = int(p / w), x = int(p - (long long)y * w);
        const bool has_r = x + 1 < w, has_d = y + 1 < h;
        float lsum = 0.0f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float ev = __ldg(e + c * plane + p), lv = __ldg(l + c * plane + p);
            acc[c] += double(ev);
            lsum = __fadd_rn(lsum, lv);
            if (has_r) {
                const float d = __fsub_rn(__fsub_rn(ev, __ldg(e + c * plane + p + 1)), __fsub_rn(lv, __ldg(l + c * plane + p + 1)));
                acc[4] += double(__fmul_rn(d, d));
            }
            if (has_d) {
                const float d = __fsub_rn(__fsub_rn(ev, __ldg(e + c * plane + p + w)), __fsub_rn(lv, __ldg(l + c * plane + p + w)));
                acc[5] += double(__fmul_rn(d, d));
            }
        }
        acc[3] += double(__fdiv_rn(lsum, 3.0f));
    }
#pragma unroll
    for (int k = 0; k < kEnhSums; ++k) {
        const double t = block_sum(acc[k], s_red);
        if (tid == 0) partial[((long long)f * parts + blockIdx.x) * kEnhSums + k] = t;
    }
    if (!last_cta_of(tickets + f, parts, &s_flag)) return;
    if (tid == 0) {
        double tot[kEnhSums] = {0, 0, 0, 0, 0, 0};
        for (int q = 0; q < parts; ++q)
            for (int k = 0; k < kEnhSums; ++k) tot[k] += __ldcg(&partial[((long long)f * parts + q) * kEnhSums + k]);
        for (int k = 0; k < kEnhSums; ++k) partial[(long long)f * parts * kEnhSums + k] = tot[k];
        __threadfence();
        unsigned* batch_ticket = tickets + n;
        if (atomicAdd(batch_ticket, 1u) != unsigned(n - 1)) return;
        *batch_ticket = 0;
        __threadfence();
        double all[kEnhSums] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < n; ++i)
            for (int k = 0; k < kEnhSums; ++k) all[k] += __ldcg(&partial[(long long)i * parts * kEnhSums + k]);
        const double npix = double(n) * h * w;
        const float mr = float(all[0] / npix), mg = float(all[1] / npix), mb = float(all[2] / npix);
        const float gmean = float(all[3] / npix);
        // base + (0.8 - base) * (1 - mean): the coefficient is formed in double on the host, as Python does at loss.py:47
        const float target = __fadd_rn(base_target, __fmul_rn(target_coef, __fsub_rn(1.0f, gmean)));
        const double nh = double(n) * 3 * h * (w - 1), nv = double(n) * 3 * (h - 1) * w;
        const float drg = __fsub_rn(mr, mg), drb = __fsub_rn(mr, mb), dgb = __fsub_rn(mg, mb);
        saved[0] = mr; saved[1] = mg; saved[2] = mb; saved[3] = gmean; saved[4] = target;
        saved[5] = float(2.0 / nh); saved[6] = float(2.0 / nv);
        saved[7] = float(1.0 / (double(n) * hp * wp * patch * patch * 3));
        losses3[1] = __fadd_rn(__fadd_rn(__fmul_rn(drg, drg), __fmul_rn(drb, drb)), __fmul_rn(dgb, dgb));
        losses3[2] = __fadd_rn(float(all[4] / nh), float(all[5] / nv));
    }
}

// one warp per patch: pm = avg_pool(mean_c R)
__global__ void __launch_bounds__(kStThreads)
k_enh_patch_means(const float* __restrict__ enh, int h, int w, int hp, int wp, int patch, float* __restrict__ pm)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long k = (long long)blockIdx.x * (kStThreads / 32) + wid;      // patch index inside the frame
    const int f = blockIdx.y;
    if (k >= (long long)hp * wp) return;
    const int py = int(k / wp), px = int(k - (long long)py * wp);
    const long long plane = (long long)h * w;
    const float* e = enh + (long long)f * 3 * plane + (long long)py * patch * w + px * patch;
    double acc = 0.0;
    for (int i = lane; i < patch * patch; i += 32) {
        const int yy = i / patch, xx = i - yy * patch;
        const long long o = (long long)yy * w + xx;
        const float g = __fdiv_rn(__fadd_rn(__fadd_rn(__ldg(e + o), __ldg(e + plane + o)), __ldg(e + 2 * plane + o)), 3.0f);
        acc += double(g);
    }
    acc = warp_sum(acc);
    if (lane == 0) pm[(long long)f * hp * wp + k] = float(acc / double(patch * patch));
}

__global__ void __launch_bounds__(kStThreads)
k_enh_exposure_finish(const float* __restrict__ pm, long long m, const float* __restrict__ saved, float* __restrict__ losses3)
{
    __shared__ double s_red[kStThreads / 32];
    const float target = __ldcg(saved + 4);
    double acc = 0.0;
    for (long long i = threadIdx.x; i < m; i += kStThreads) acc += double(fabsf(__fsub_rn(__ldg(pm + i), target)));
    acc = block_sum(acc, s_red);
    if (threadIdx.x == 0) losses3[0] = float(acc / double(m));
}

__global__ void __launch_bounds__(kStThreads)
k_enh_grad(const float* __restrict__ enh, const float* __restrict__ low, int h, int w, int hp, int wp, int patch,
           const float* __restrict__ saved, const float* __restrict__ pm, const float* __restrict__ upstream3, float* __restrict__ grad)
{
    const int f = blockIdx.y;
    const long long plane = (long long)h * w;
    const float* e = enh + (long long)f * 3 * plane;
    const float* l = low + (long long)f * 3 * plane;
    float* g = grad + (long long)f * 3 * plane;
    const float u_exp = __ldg(upstream3 + 0), u_col = __ldg(upstream3 + 1), u_spa = __ldg(upstream3 + 2);
    const float mr = __ldg(saved + 0), mg = __ldg(saved + 1), mb = __ldg(saved + 2), target = __ldg(saved + 4);
    const float sh = __ldg(saved + 5), sv = __ldg(saved + 6), se = __ldg(saved + 7);
    const float inv_npix = float(1.0 / (double(gridDim.y) * h * w));
    const float drg = __fsub_rn(mr, mg), drb = __fsub_rn(mr, mb), dgb = __fsub_rn(mg, mb);
    // d colour / d mean_c, spread over the pixels of the channel
    const float gc[3] = {__fmul_rn(__fmul_rn(2.0f, __fadd_rn(drg, drb)), inv_npix),
                         __fmul_rn(__fmul_rn(2.0f, __fsub_rn(dgb, drg)), inv_npix),
                         __fmul_rn(__fmul_rn(-2.0f, __fadd_rn(drb, dgb)), inv_npix)};
    const long long stride = (long long)gridDim.x * kStThreads;
    for (long long p = (long long)blockIdx.x * kStThreads + threadIdx.x; p < plane; p += stride) {
        const int y = int(p / w), x = int(p - (long long)y * w);
        const bool has_r = x + 1 < w, has_l = x > 0, has_d = y + 1 < h, has_u = y > 0;
        float ge = 0.0f;
        if (y < hp * patch && x < wp * patch) {
            const float d = __fsub_rn(__ldg(pm + (long long)f * hp * wp + (long long)(y / patch) * wp + x / patch), target);
            ge = __fmul_rn(d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f), se);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float* ec = e + c * plane;
            const float* lc = l + c * plane;
            const float ev = __ldg(ec + p), lv = __ldg(lc + p);
            float gs = 0.0f;
            if (has_r) gs = __fadd_rn(gs, __fmul_rn(sh, __fsub_rn(__fsub_rn(ev, __ldg(ec + p + 1)), __fsub_rn(lv, __ldg(lc + p + 1)))));
            if (has_l) gs = __fsub_rn(gs, __fmul_rn(sh, __fsub_rn(__fsub_rn(__ldg(ec + p - 1), ev), __fsub_rn(__ldg(lc + p - 1), lv))));
            if (has_d) gs = __fadd_rn(gs, __fmul_rn(sv, __fsub_rn(__fsub_rn(ev, __ldg(ec + p + w)), __fsub_rn(lv, __ldg(lc + p + w)))));
            if (has_u) gs = __fsub_rn(gs, __fmul_rn(sv, __fsub_rn(__fsub_rn(__ldg(ec + p - w), ev), __fsub_rn(__ldg(lc + p - w), lv))));
            g[c * plane + p] = __fadd_rn(__fadd_rn(__fmul_rn(u_exp, ge), __fmul_rn(u_col, gc[c])), __fmul_rn(u_spa, gs));
        }
    }
}

__global__ void k_dynamic_weight(const float* __restrict__ stats2, float w0, float* __restrict__ out)
{
    out[0] = dyn_weight(stats2[0], stats2[1], w0);
}

static int tex_parts(int n, long long items_per_image)
{
    const long long by_work = std::max<long long>(1, items_per_image / (kStThreads * 4));
    const long long by_fill = (4LL * kNumSMsB200 + n - 1) / n;
    return int(std::max<long long>(1, std::min<long long>(std::min(by_work, by_fill), 1024)));
}

struct TexLayout {
    size_t off_partial, off_tickets, off_mean, total;
};
// The ticket words come FIRST, at offsets that do not depend on the batch size: the kernels leave them at zero, and a
// workspace that was zero-filled once may then be reused for ANY later batch size (a trainer's last, smaller batch).  With
// the tickets behind the n-dependent partial sums, a call with a different n found stale partial sums where it expected
// clean tickets and summed an incomplete set of partials.
constexpr size_t kTicketRegionBytes = (size_t(65535) + 2) * sizeof(unsigned);   // n <= 65535 image tickets + 1 batch ticket
static TexLayout tex_layout(int n)
{
    TexLayout L;
    L.off_tickets = 0;
    L.off_mean = align_up(kTicketRegionBytes, 256);
    L.off_partial = align_up(L.off_mean + size_t(n) * sizeof(double), 256);
    L.total = align_up(L.off_partial + size_t(n) * 1024 * 2 * sizeof(double), 256);
    return L;
}

}  // namespace upr

extern "C" {

int upr_brightness_hist_f32(const float* in_nchw, int n, int h, int w, uint32_t* hist256_per_image, upr_stream_t stream)
{
    if (n < 0 || h <= 0 || w <= 0) return UPR_E_SHAPE;
    if (n == 0) return UPR_OK;
    if (!in_nchw || !hist256_per_image) return UPR_E_NULL;
    auto s = static_cast<cudaStream_t>(stream);
    const long long plane = (long long)h * w;
    UPR_CUDA_TRY(cudaMemsetAsync(hist256_per_image, 0, size_t(n) * 256 * sizeof(uint32_t), s));
    const int parts = upr::tex_parts(n, plane / 4 + 1);
    for (int f0 = 0; f0 < n; f0 += 65535) {
        const int nf = std::min(n - f0, 65535);
        upr::k_brightness_hist<<<dim3(parts, nf), upr::kStThreads, 0, s>>>(in_nchw + (long long)f0 * 3 * plane,
                                                                           hist256_per_image + (long long)f0 * 256, plane);
        UPR_LAUNCH_CHECK();
    }
    return UPR_OK;
}

size_t upr_texture_workspace_bytes(int n)
{
    if (n < 0) return 0;
    return upr::tex_layout(std::max(n, 1)).total;
}

int upr_texture_workspace_init(void* workspace, size_t workspace_bytes, int n, upr_stream_t stream)
{
    if (n < 0) return UPR_E_SHAPE;
    if (!workspace) return UPR_E_NULL;
    const upr::TexLayout lay = upr::tex_layout(std::max(n, 1));
    if (workspace_bytes < lay.total) return UPR_E_WORKSPACE;
    UPR_CUDA_TRY(cudaMemsetAsync(workspace, 0, lay.total, static_cast<cudaStream_t>(stream)));
    return UPR_OK;
}

static int texture_run(int method, const float* x, int n, int c, int h, int w, float* per_image, float* stats2,
                       void* ws, size_t ws_bytes, cudaStream_t s,
                       const upr::PeerXchg px = upr::PeerXchg{nullptr, 0, 1, 0u, 0.0f, nullptr, 0ull})
{
    if (n < 0 || n > 65535 || c <= 0 || h <= 0 || w <= 0) return UPR_E_SHAPE;
    if (n == 0) return UPR_OK;
    if (!x || !per_image || !ws) return UPR_E_NULL;
    const upr::TexLayout lay = upr::tex_layout(n);
    if (ws_bytes < lay.total || (reinterpret_cast<uintptr_t>(ws) & 255u)) return UPR_E_WORKSPACE;
    auto* base = static_cast<unsigned char*>(ws);
    auto* partial = reinterpret_cast<double*>(base + lay.off_partial);
    auto* tickets = reinterpret_cast<unsigned*>(base + lay.off_tickets);
    auto* mean = reinterpret_cast<double*>(base + lay.off_mean);
    const long long plane = (long long)h * w;
    {
        const int nf = n, f0 = 0;
        const float* xf = x;
        if (method == 0) {
            const int parts = upr::tex_parts(n, (long long)c * plane / 4 + 1);
            upr::k_texture_tv<<<dim3(parts, nf), upr::kStThreads, 0, s>>>(xf, c, h, w, partial + (long long)f0 * parts * 2,
                                                                         tickets + f0, per_image + f0, stats2, px);
            UPR_LAUNCH_CHECK();
        } else {
            const int parts = upr::tex_parts(n, plane);
            upr::k_texture_edge<1><<<dim3(parts, nf), upr::kStThreads, 0, s>>>(xf, c, h, w, partial + (long long)f0 * parts,
                                                                              tickets + f0, mean + f0, per_image + f0, nullptr, px);
            UPR_LAUNCH_CHECK();
            upr::k_texture_edge<2><<<dim3(parts, nf), upr::kStThreads, 0, s>>>(xf, c, h, w, partial + (long long)f0 * parts,
                                                                              tickets + f0, mean + f0, per_image + f0, stats2, px);
            UPR_LAUNCH_CHECK();
        }
    }
    return UPR_OK;
}

int upr_texture_tv_f32(const float* x, int n, int c, int h, int w, float* per_image, float* batch_stats2, void* workspace,
                       size_t workspace_bytes, upr_stream_t stream)
{
    return texture_run(0, x, n, c, h, w, per_image, batch_stats2, workspace, workspace_bytes,
                       static_cast<cudaStream_t>(stream));
}

int upr_texture_edge_density_f32(const float* x, int n, int c, int h, int w, float* per_image, float* batch_stats2,
                                 void* workspace, size_t workspace_bytes, upr_stream_t stream)
{
    return texture_run(1, x, n, c, h, w, per_image, batch_stats2, workspace, workspace_bytes,
                       static_cast<cudaStream_t>(stream));
}

size_t upr_smooth_loss_workspace_bytes(int n, int h, int w)
{
    if (n < 0 || h < 2 || w < 2) return 0;
    const size_t nn = size_t(std::max(n, 1));
    return upr::align_up(nn * 1024 * 2 * sizeof(double), 256) + upr::align_up((nn + 1) * sizeof(unsigned), 256) +
           upr::align_up(nn * h * sizeof(float), 256) + upr::align_up(nn * w * sizeof(float), 256) +
           upr::align_up(nn * size_t(h) * w * sizeof(float), 256);
}

int upr_edge_smooth_loss_f32(const float* illu, const float* img_low, int n, int ci, int cs, int h, int w, float lambda_val, float alpha,
                             float* loss3, float* grad_illu, void* workspace, size_t workspace_bytes, upr_stream_t stream)
{
    using namespace upr;
    if (n <= 0 || n > 65535 || ci <= 0 || cs <= 0 || h < 2 || w < 2) return UPR_E_SHAPE;
    if (!illu || !img_low || !loss3 || !workspace) return UPR_E_NULL;
    if (workspace_bytes < upr_smooth_loss_workspace_bytes(n, h, w) || (reinterpret_cast<uintptr_t>(workspace) & 255u)) return UPR_E_WORKSPACE;
    auto s = static_cast<cudaStream_t>(stream);
    auto* base = static_cast<unsigned char*>(workspace);
    auto* partial = reinterpret_cast<double*>(base);
    size_t off = align_up(size_t(n) * 1024 * 2 * sizeof(double), 256);
    auto* tickets = reinterpret_cast<unsigned*>(base + off);
    off += align_up((size_t(n) + 1) * sizeof(unsigned), 256);
    auto* rowmean = reinterpret_cast<float*>(base + off);
    off += align_up(size_t(n) * h * sizeof(float), 256);
    auto* colmean = reinterpret_cast<float*>(base + off);
    off += align_up(size_t(n) * w * sizeof(float), 256);
    auto* edge = reinterpret_cast<float*>(base + off);
    // the tickets are self-cleaning, but this workspace is not required to arrive zero-filled: clear them (n + 1 words)
    UPR_CUDA_TRY(cudaMemsetAsync(tickets, 0, (size_t(n) + 1) * sizeof(unsigned), s));
    const int rows_per_cta = kStThreads / 32;
    k_smooth_edge_rows<<<dim3((h + rows_per_cta - 1) / rows_per_cta, n), kStThreads, 0, s>>>(img_low, cs, h, w, edge, rowmean);
    UPR_LAUNCH_CHECK();
    k_smooth_edge_cols<<<dim3((w + kStThreads - 1) / kStThreads, n), kStThreads, 0, s>>>(edge, h, w, colmean);
    UPR_LAUNCH_CHECK();
    const int parts = tex_parts(n, (long long)h * w);
    k_smooth_loss<<<dim3(parts, n), kStThreads, 0, s>>>(illu, img_low, n, ci, cs, h, w, lambda_val, alpha, rowmean, colmean, grad_illu,
                                                         partial, tickets, loss3);
    UPR_LAUNCH_CHECK();
    return UPR_OK;
}

size_t upr_enh_losses_workspace_bytes(int n)
{
    if (n < 0) return 0;
    const size_t nn = size_t(std::max(n, 1));
    return upr::align_up(nn * 1024 * upr::kEnhSums * sizeof(double), 256) + upr::align_up((nn + 1) * sizeof(unsigned), 256);
}

size_t upr_enh_losses_saved_floats(int n, int h, int w, int patch)
{
    if (n < 0 || patch <= 0 || h < patch || w < patch) return 0;
    return 8 + size_t(std::max(n, 1)) * (h / patch) * (w / patch);
}

int upr_enh_losses_f32(const float* enhanced, const float* img_low, int n, int h, int w, double base_target, int patch, float* losses3,
                       float* saved, void* workspace, size_t workspace_bytes, upr_stream_t stream)
{
    using namespace upr;
    if (n <= 0 || n > 65535 || patch <= 0 || h < patch || w < patch || h < 2 || w < 2) return UPR_E_SHAPE;
    if (!enhanced || !img_low || !losses3 || !saved || !workspace) return UPR_E_NULL;
    if (workspace_bytes < upr_enh_losses_workspace_bytes(n) || (reinterpret_cast<uintptr_t>(workspace) & 255u)) return UPR_E_WORKSPACE;
    auto s = static_cast<cudaStream_t>(stream);
    auto* base = static_cast<unsigned char*>(workspace);
    auto* partial = reinterpret_cast<double*>(base);
    auto* tickets = reinterpret_cast<unsigned*>(base + align_up(size_t(n) * 1024 * kEnhSums * sizeof(double), 256));
    UPR_CUDA_TRY(cudaMemsetAsync(tickets, 0, (size_t(n) + 1) * sizeof(unsigned), s));
    const int hp = h / patch, wp = w / patch;
    const int parts = tex_parts(n, (long long)h * w);
    k_enh_sums<<<dim3(parts, n), kStThreads, 0, s>>>(enhanced, img_low, n, h, w, partial, tickets, float(base_target), float(0.8 - base_target),
                                                     hp, wp, patch, saved, losses3);
    UPR_LAUNCH_CHECK();
    const long long per = (long long)hp * wp;
    const int wpc = kStThreads / 32;
    k_enh_patch_means<<<dim3(unsigned((per + wpc - 1) / wpc), n), kStThreads, 0, s>>>(enhanced, h, w, hp, wp, patch, saved + 8);
    UPR_LAUNCH_CHECK();
    k_enh_exposure_finish<<<1, kStThreads, 0, s>>>(saved + 8, per * n, saved, losses3);
    UPR_LAUNCH_CHECK();
    return UPR_OK;
}

int upr_enh_losses_grad_f32(const float* enhanced, const float* img_low, int n, int h, int w, int patch, const float* saved,
                            const float* upstream3, float* grad_enhanced, upr_stream_t stream)
{
    using namespace upr;
    if (n <= 0 || n > 65535 || patch <= 0 || h < patch || w < patch || h < 2 || w < 2) return UPR_E_SHAPE;
    if (!enhanced || !img_low || !saved || !upstream3 || !grad_enhanced) return UPR_E_NULL;
    const int parts = tex_parts(n, (long long)h * w);
    k_enh_grad<<<dim3(parts, n), kStThreads, 0, static_cast<cudaStream_t>(stream)>>>(enhanced, img_low, h, w, h / patch, w / patch, patch, saved,
                                                                                     saved + 8, upstream3, grad_enhanced);
    UPR_LAUNCH_CHECK();
    return UPR_OK;
}

size_t upr_peer_stats_buffer_bytes(void) { return upr::kPeerBufBytes; }

int upr_peer_set_timeout_ms(double ms)
{
    if (!(ms > 0.0) || ms > 1e12) return UPR_E_PARAM;
    upr::g_peer_timeout_ns = (unsigned long long)(ms * 1e6);
    return UPR_OK;
}

int upr_peer_status(const void* own_buffer_dev, unsigned* status2_host, upr_stream_t stream)
{
    if (!own_buffer_dev || !status2_host) return UPR_E_NULL;
    auto s = static_cast<cudaStream_t>(stream);
    UPR_CUDA_TRY(cudaMemcpyAsync(status2_host, static_cast<const char*>(own_buffer_dev) + upr::kPeerStatusOff, 2 * sizeof(unsigned),
                                 cudaMemcpyDeviceToHost, s));
    UPR_CUDA_TRY(cudaStreamSynchronize(s));
    return UPR_OK;
}

int upr_texture_weight_peer_f32(const float* x, int n, int c, int h, int w, int method, float* per_image, float* batch_stats2,
                                void* workspace, size_t workspace_bytes, const unsigned long long* peer_buffers_dev, int rank,
                                int world, unsigned seq, float weight_smooth, float* weight_out, upr_stream_t stream)
{
    if (method != 0 && method != 1) return UPR_E_PARAM;
    if (!batch_stats2 || !weight_out) return UPR_E_NULL;
    if (peer_buffers_dev && (world < 1 || world > upr::kPeerMax || rank < 0 || rank >= world || seq == 0)) return UPR_E_PARAM;
    if (n <= 0) return UPR_E_SHAPE;   // every rank must contribute to the exchange
    const upr::PeerXchg px{peer_buffers_dev, rank, world, seq, weight_smooth, weight_out, upr::g_peer_timeout_ns};
    return texture_run(method, x, n, c, h, w, per_image, batch_stats2, workspace, workspace_bytes,
                       static_cast<cudaStream_t>(stream), px);
}

int upr_dynamic_smooth_weight_f32(const float* batch_stats2, float weight_smooth, float* weight_out, upr_stream_t stream)
{
    if (!batch_stats2 || !weight_out) return UPR_E_NULL;
    upr::k_dynamic_weight<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(batch_stats2, weight_smooth, weight_out);
    UPR_LAUNCH_CHECK();
    return UPR_OK;
}

}  // extern "C"
