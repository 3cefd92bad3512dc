/*
 * upretinex_b200.h -- C ABI of libupretinex_b200.so, the B200 (sm_100a) implementation of the
 * classical image-processing hot path of UP-Retinex (xh92117/Retinex-image-Enhancement).
 *
 * The reference has no FFI of its own: its hot path is Python that calls OpenCV / torch on
 * the host.  Each entry point below replaces the arithmetic of one reference function
 * (cited as file:line relative to the reference tree); INTEGRATION.md shows the ctypes stub
 * a maintainer of the reference would add at each call site.
 *
 * Conventions
 *   - All image pointers are DEVICE pointers owned by the caller unless the name ends in
 *     `_host`.  Images are planar NCHW, f32 in [0,1] (values outside are handled exactly
 *     like the reference handles them).
 *   - Every call is asynchronous on the supplied stream (a cudaStream_t passed as void*),
 *     never synchronises, never allocates device memory (workspaces are caller-owned; their
 *     size comes from the matching *_workspace_bytes()).  The `_host` entry points are the
 *     exception: they own a pinned/device staging pool and return after the result is in
 *     the host buffer.
 *   - Return value: 0 = OK, < 0 = invalid argument (UPR_E_*), > 0 = cudaError_t.
 *   - Thread-safe and re-entrant across streams; a workspace must not be shared by calls
 *     that may run concurrently.
 *   - There is no CPU fallback: without an sm_100 device every call fails.
 */
#ifndef UPRETINEX_B200_H
#define UPRETINEX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define UPR_API
#else
#define UPR_API __attribute__((visibility("default")))
#endif

typedef void* upr_stream_t; /* cudaStream_t */

enum {
    UPR_OK = 0,
    UPR_E_NULL = -1,      /* null pointer argument */
    UPR_E_SHAPE = -2,     /* non-positive or unsupported shape */
    UPR_E_WORKSPACE = -3, /* workspace too small or misaligned */
    UPR_E_PARAM = -4,     /* bad scalar parameter */
    UPR_E_DEVICE = -5     /* current device is not sm_100 */
};

UPR_API const char* upr_version(void);
UPR_API const char* upr_status_string(int status);
/* 0 if the current CUDA device is a compute-capability 10.x part, UPR_E_DEVICE otherwise. */
UPR_API int upr_device_check(void);

/* ---- a1: CLAHE in Lab ---------------------------------------------------------------
 * Replaces AdaptiveParameterAdjuster.apply_clahe_enhancement
 * (enhancers/adaptive_params.py:121-169): unclipped trunc-quantise -> sRGB u8 -> OpenCV
 * 8-bit Lab -> CLAHE(clip_limit, tiles) on L -> Lab -> sRGB u8 -> /255.
 * Bit-exact with the reference for every frame of the batch (the reference is batch 1). */
UPR_API size_t upr_clahe_workspace_bytes(int n, int h, int w, int tiles_x, int tiles_y);
UPR_API int upr_clahe_lab_f32(const float* in_nchw, float* out_nchw, int n, int h, int w,
                              double clip_limit, int tiles_x, int tiles_y,
                              void* workspace, size_t workspace_bytes, upr_stream_t stream);
/* Copies the per-tile raw histograms ([n][tiles_y*tiles_x][256] int32), LUTs
 * ([n][tiles_y*tiles_x][256] u8) and/or the u8 Lab intermediate ([n][3][h][w]) of the LAST
 * upr_clahe_lab_f32 call that used `workspace` into caller device buffers (any may be NULL). */
UPR_API int upr_clahe_debug_dump(const void* workspace, int n, int h, int w, int tiles_x, int tiles_y,
                                 int32_t* hist_out, uint8_t* lut_out, uint8_t* lab_out,
                                 upr_stream_t stream);
/* Host copies of the constant tables the kernels use (for tests/audits). Any may be NULL.
 * gamma[256] u16, cbrt[2048] u16, labyf[256] u32 (ify | y<<16), invgamma[4096] u8. */
UPR_API int upr_get_tables(uint16_t* gamma, uint16_t* cbrt, uint32_t* labyf, uint8_t* invgamma);

#ifdef __cplusplus
}
#endif
#endif /* UPRETINEX_B200_H */
