// upr_host.cu -- host-buffer entry point of the CLAHE-in-Lab op.
//
// The reference's apply_clahe_enhancement (/root/reference/enhancers/adaptive_params.py:121-169) takes a tensor
// that may live on the host and returns a HOST tensor (:164-167).  upr_clahe_lab_f32_host is the same
// contract at the C ABI: host f32 NCHW in, host f32 NCHW out.  Frames are independent, so the batch is cut
// into chunks that ride a ring of kSlots streams: H2D(chunk i+1) overlaps the kernels of chunk i and the
// D2H of chunk i-1 (PCIe is full duplex; B200 has separate copy engines per direction).  The device staging
// buffers and workspaces belong to a per-device pool that grows on demand and is released by
// upr_host_pool_release().  Pinned (page-locked) host buffers give truly asynchronous copies; pageable
// buffers work but serialise inside the driver.
#include <algorithm>
#include <mutex>

#include "upr_common.cuh"

namespace upr {

constexpr int kSlots = 3;
constexpr int kMaxDevices = 16;
constexpr size_t kChunkTargetBytes = size_t(96) << 20;  // f32 input bytes per chunk

struct HostSlot {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    void* d_in = nullptr;
    void* d_out = nullptr;
    void* ws = nullptr;
    size_t img_bytes = 0, ws_bytes = 0;
    bool busy = false;
};

struct HostPool {
    std::mutex mu;
    HostSlot slot[kSlots];
};

static HostPool g_pool[kMaxDevices];

static int slot_reserve(HostSlot& s, size_t img_bytes, size_t ws_bytes)
{
    if (!s.stream) {
        UPR_CUDA_TRY(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        UPR_CUDA_TRY(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    }
    if (s.img_bytes < img_bytes) {
        if (s.d_in) cudaFree(s.d_in);
        if (s.d_out) cudaFree(s.d_out);
        s.d_in = s.d_out = nullptr;
        s.img_bytes = 0;
        UPR_CUDA_TRY(cudaMalloc(&s.d_in, img_bytes));
        UPR_CUDA_TRY(cudaMalloc(&s.d_out, img_bytes));
        s.img_bytes = img_bytes;
    }
    if (s.ws_bytes < ws_bytes) {
        if (s.ws) cudaFree(s.ws);
        s.ws = nullptr;
        s.ws_bytes = 0;
        UPR_CUDA_TRY(cudaMalloc(&s.ws, ws_bytes));
        s.ws_bytes = ws_bytes;
    }
    return UPR_OK;
}

// mode 0: f32 NCHW -> f32 NCHW (upr_clahe_lab_f32); mode 1: packed u8 HWC -> packed u8 HWC (upr_clahe_lab_u8)
static int host_pipeline(int mode, const void* in_host, void* out_host, int n, int h, int w, double clip_limit, int tiles_x,
                         int tiles_y, int frames_per_chunk)
{
    if (n < 0 || h <= 0 || w <= 0 || tiles_x <= 0 || tiles_y <= 0) return UPR_E_SHAPE;
    if (n == 0) return UPR_OK;
    if (!in_host || !out_host) return UPR_E_NULL;
    int dev = 0;
    UPR_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) return UPR_E_DEVICE;
    const size_t frame_bytes = size_t(3) * h * w * (mode == 0 ? sizeof(float) : 1);
    int chunk = frames_per_chunk > 0 ? frames_per_chunk : int(std::max<size_t>(1, kChunkTargetBytes / (size_t(3) * h * w * sizeof(float))));
    chunk = std::min(chunk, n);
    const size_t ws_bytes = upr_clahe_workspace_bytes(chunk, h, w, tiles_x, tiles_y);
    if (ws_bytes == 0) return UPR_E_SHAPE;

    HostPool& pool = g_pool[dev];
    std::lock_guard<std::mutex> lock(pool.mu);
    for (auto& s : pool.slot) {
        const int rc = slot_reserve(s, size_t(chunk) * frame_bytes, ws_bytes);
        if (rc) return rc;
    }
    int rc = UPR_OK;
    int idx = 0;
    // The first upload and the last download overlap with nothing: a half-sized first and last chunk shortens both ends of the
    // pipeline (64 x 1080p, chunks of 2: head and tail of one frame).
    const int edge = n > 2 * chunk ? std::max(1, chunk / 2) : chunk;
    for (int f0 = 0, nf = 0; f0 < n && rc == UPR_OK; f0 += nf, ++idx) {
        HostSlot& s = pool.slot[idx % kSlots];
        const int left = n - f0;
        nf = idx == 0 ? edge : chunk;
        if (left - nf < edge && left > edge) nf = left - edge;      // leave exactly `edge` frames for the last chunk
        nf = std::min(nf, left);
        const size_t bytes = size_t(nf) * frame_bytes;
        if (s.busy) {  // the slot's previous D2H must have drained before its buffers are reused
            cudaError_t e = cudaEventSynchronize(s.done);
            if (e != cudaSuccess) { rc = int(e); break; }
            s.busy = false;
        }
        cudaError_t e = cudaMemcpyAsync(s.d_in, static_cast<const char*>(in_host) + size_t(f0) * frame_bytes, bytes,
                                        cudaMemcpyHostToDevice, s.stream);
        if (e != cudaSuccess) { rc = int(e); break; }
        rc = mode == 0 ? upr_clahe_lab_f32(static_cast<const float*>(s.d_in), static_cast<float*>(s.d_out), nf, h, w, clip_limit,
                                           tiles_x, tiles_y, s.ws, s.ws_bytes, s.stream)
                       : upr_clahe_lab_u8(static_cast<const unsigned char*>(s.d_in), static_cast<unsigned char*>(s.d_out), nf, h, w,
                                          clip_limit, tiles_x, tiles_y, s.ws, s.ws_bytes, s.stream);
        if (rc) break;
        e = cudaMemcpyAsync(static_cast<char*>(out_host) + size_t(f0) * frame_bytes, s.d_out, bytes, cudaMemcpyDeviceToHost, s.stream);
        if (e != cudaSuccess) { rc = int(e); break; }
        e = cudaEventRecord(s.done, s.stream);
        if (e != cudaSuccess) { rc = int(e); break; }
        s.busy = true;
    }
    for (auto& s : pool.slot) {
        if (s.busy) {
            cudaError_t e = cudaEventSynchronize(s.done);
            if (e != cudaSuccess && rc == UPR_OK) rc = int(e);
            s.busy = false;
        }
    }
    return rc;
}

}  // namespace upr

extern "C" {

int upr_clahe_lab_f32_host(const float* in_host, float* out_host, int n, int h, int w, double clip_limit, int tiles_x,
                           int tiles_y, int frames_per_chunk)
{
    return upr::host_pipeline(0, in_host, out_host, n, h, w, clip_limit, tiles_x, tiles_y, frames_per_chunk);
}

int upr_clahe_lab_u8_host(const unsigned char* in_host, unsigned char* out_host, int n, int h, int w, double clip_limit, int tiles_x,
                          int tiles_y, int frames_per_chunk)
{
    return upr::host_pipeline(1, in_host, out_host, n, h, w, clip_limit, tiles_x, tiles_y, frames_per_chunk);
}

int upr_host_pool_release(void)
{
    using namespace upr;
    int dev = 0;
    UPR_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) return UPR_E_DEVICE;
    HostPool& pool = g_pool[dev];
    std::lock_guard<std::mutex> lock(pool.mu);
    for (auto& s : pool.slot) {
        if (s.stream) cudaStreamSynchronize(s.stream);
        if (s.d_in) cudaFree(s.d_in);
        if (s.d_out) cudaFree(s.d_out);
        if (s.ws) cudaFree(s.ws);
        if (s.done) cudaEventDestroy(s.done);
        if (s.stream) cudaStreamDestroy(s.stream);
        s = HostSlot{};
    }
    return UPR_OK;
}

}  // extern "C"
