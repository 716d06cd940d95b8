// orbb_preview.cu -- the JPEG preview the reference ships to its UI with every processed frame:
// host-side mirror of reference src/SlamGpuPipeline/buildStream.cpp:266-277 (nvJPEG encoder set-up: quality 90,
// 4:2:0), :491-521 (gray frame copied into three planes, keypoints painted into the G plane by overlay_keypoints,
// nvjpegEncodeImage) and :613-621 (bitstream retrieved to host memory), and of kernel_overlay_keypoints
// (src/cuda/post_processing.cu:45-70).  Differences, all deliberate: the planes and the encoder state are allocated
// once (the reference mallocs the host image per frame), one kernel fills the three planes and a second paints the
// keypoints (the reference issues three cudaMemcpy2DAsync + a kernel + two stream syncs before the encode), the
// overlay clamps at the LOWER image edge too (the reference only tests the upper one and can write before the plane),
// and the keypoint count is read on the device.  nvJPEG is library code (as in the reference); the two kernels are ours.
#include <dlfcn.h>
#include <nvjpeg.h>

#include <algorithm>
#include <cstdio>
#include <new>

#include "orbb_internal.cuh"

// libnvjpeg is opened on first use (dlopen), not linked: liborbb200.so itself stays sm_100a-only code with no
// dependency beyond the CUDA runtime, and a box without nvJPEG loses only the preview (orbb_preview_create fails).
namespace {
struct NvJpegApi {
    void *lib = nullptr;
    nvjpegStatus_t (*CreateSimple)(nvjpegHandle_t *) = nullptr;
    nvjpegStatus_t (*Destroy)(nvjpegHandle_t) = nullptr;
    nvjpegStatus_t (*EncoderStateCreate)(nvjpegHandle_t, nvjpegEncoderState_t *, cudaStream_t) = nullptr;
    nvjpegStatus_t (*EncoderStateDestroy)(nvjpegEncoderState_t) = nullptr;
    nvjpegStatus_t (*EncoderParamsCreate)(nvjpegHandle_t, nvjpegEncoderParams_t *, cudaStream_t) = nullptr;
    nvjpegStatus_t (*EncoderParamsDestroy)(nvjpegEncoderParams_t) = nullptr;
    nvjpegStatus_t (*EncoderParamsSetQuality)(nvjpegEncoderParams_t, const int, cudaStream_t) = nullptr;
    nvjpegStatus_t (*EncoderParamsSetSamplingFactors)(nvjpegEncoderParams_t, const nvjpegChromaSubsampling_t, cudaStream_t) = nullptr;
    nvjpegStatus_t (*EncodeImage)(nvjpegHandle_t, nvjpegEncoderState_t, const nvjpegEncoderParams_t, const nvjpegImage_t *,
                                  nvjpegInputFormat_t, int, int, cudaStream_t) = nullptr;
    nvjpegStatus_t (*EncodeRetrieveBitstream)(nvjpegHandle_t, nvjpegEncoderState_t, unsigned char *, size_t *, cudaStream_t) = nullptr;
};
NvJpegApi *nvjpeg_api() {
    static NvJpegApi api;
    static bool tried = false;
    if (tried) return api.lib ? &api : nullptr;
    tried = true;
    for (const char *name : {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12", "/usr/local/cuda/lib64/libnvjpeg.so"}) {
        api.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
        if (api.lib) break;
    }
    if (!api.lib) return nullptr;
    bool ok = true;
#define ORBB_NVJ(f) ok = ok && (*reinterpret_cast<void **>(&api.f) = dlsym(api.lib, "nvjpeg" #f)) != nullptr
    ORBB_NVJ(CreateSimple); ORBB_NVJ(Destroy); ORBB_NVJ(EncoderStateCreate); ORBB_NVJ(EncoderStateDestroy);
    ORBB_NVJ(EncoderParamsCreate); ORBB_NVJ(EncoderParamsDestroy); ORBB_NVJ(EncoderParamsSetQuality);
    ORBB_NVJ(EncoderParamsSetSamplingFactors); ORBB_NVJ(EncodeImage); ORBB_NVJ(EncodeRetrieveBitstream);
#undef ORBB_NVJ
    if (!ok) { dlclose(api.lib); api.lib = nullptr; return nullptr; }
    return &api;
}
}  // namespace

struct orbb_preview {
    int w = 0, h = 0, device = 0, quality = 90;
    size_t pitch = 0;
    uint8_t *d_planes = nullptr;  // 3 planes of h rows x pitch bytes (R, G, B: all the gray frame, G with the overlay)
    nvjpegHandle_t nv = nullptr;
    nvjpegEncoderState_t state = nullptr;
    nvjpegEncoderParams_t params = nullptr;
    char err[160] = {0};
};

namespace {

// gray frame -> three identical planes, 16 bytes per thread (the reference: 3 x cudaMemcpy2DAsync, buildStream.cpp:494-508)
__global__ void __launch_bounds__(256) k_preview_planes(const uint8_t *__restrict__ gray, size_t gray_pitch, int w, int h,
                                                        uint8_t *__restrict__ planes, size_t pitch) {
    const int xq = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (16 * xq >= w) return;
    const uint8_t *src = gray + (size_t)y * gray_pitch + 16 * xq;
    uint4 v;
    if (16 * xq + 16 <= w && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) v = __ldg(reinterpret_cast<const uint4 *>(src));
    else {
        uint8_t b[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) b[i] = 16 * xq + i < w ? __ldg(src + i) : 0;
        v = *reinterpret_cast<uint4 *>(b);
    }
    const size_t plane = pitch * (size_t)h;
    uint8_t *dst = planes + (size_t)y * pitch + 16 * xq;  // pitch is a multiple of 16
    *reinterpret_cast<uint4 *>(dst) = v;
    *reinterpret_cast<uint4 *>(dst + plane) = v;
    *reinterpret_cast<uint4 *>(dst + 2 * plane) = v;
}

// kernel_overlay_keypoints (post_processing.cu:45-70): for (int x = pos.x - 1; x < pos.x + 1; x++) -- the int start
// truncates, the float bound does not -- same for y; every pixel of that block becomes 255 in the G plane
__global__ void __launch_bounds__(128) k_preview_overlay(uint8_t *__restrict__ plane_g, size_t pitch, int w, int h,
                                                         const uint8_t *__restrict__ xy, int xy_stride, int n_host,
                                                         const int *__restrict__ n_dev) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = n_dev ? min(*n_dev, n_host) : n_host;
    if (i >= n) return;
    const float px = *reinterpret_cast<const float *>(xy + (size_t)i * xy_stride);
    const float py = *reinterpret_cast<const float *>(xy + (size_t)i * xy_stride + 4);
    for (int x = (int)__fsub_rn(px, 1.0f); (float)x < __fadd_rn(px, 1.0f); ++x)
        for (int y = (int)__fsub_rn(py, 1.0f); (float)y < __fadd_rn(py, 1.0f); ++y)
            if (x >= 0 && y >= 0 && x < w && y < h) plane_g[(size_t)y * pitch + x] = 255;
}

}  // namespace

#define PCK(p, call)                                                                                          \
    do {                                                                                                      \
        cudaError_t e__ = (call);                                                                             \
        if (e__ != cudaSuccess) {                                                                             \
            snprintf((p)->err, sizeof((p)->err), "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return ORBB_ERR_CUDA;                                                                             \
        }                                                                                                     \
    } while (0)
#define PNV(p, call)                                                                                          \
    do {                                                                                                      \
        nvjpegStatus_t s__ = (call);                                                                          \
        if (s__ != NVJPEG_STATUS_SUCCESS) {                                                                   \
            snprintf((p)->err, sizeof((p)->err), "%s:%d %s: nvjpeg status %d", __FILE__, __LINE__, #call, (int)s__); \
            return ORBB_ERR_CUDA;                                                                             \
        }                                                                                                     \
    } while (0)

extern "C" int orbb_preview_destroy(orbb_preview *p) {
    if (!p) return ORBB_OK;
    cudaSetDevice(p->device);
    if (p->params) nvjpeg_api()->EncoderParamsDestroy(p->params);  // non-null only if the library was found
    if (p->state) nvjpeg_api()->EncoderStateDestroy(p->state);
    if (p->nv) nvjpeg_api()->Destroy(p->nv);
    if (p->d_planes) cudaFree(p->d_planes);
    delete p;
    return ORBB_OK;
}

static int preview_init(orbb_preview *p) {
    if (!nvjpeg_api()) { snprintf(p->err, sizeof(p->err), "libnvjpeg.so.12 not found (dlopen)"); return ORBB_ERR_NO_DEVICE; }
    PCK(p, cudaSetDevice(p->device));
    PCK(p, cudaMalloc(&p->d_planes, 3 * p->pitch * (size_t)p->h));
    PNV(p, nvjpeg_api()->CreateSimple(&p->nv));
    PNV(p, nvjpeg_api()->EncoderStateCreate(p->nv, &p->state, nullptr));
    PNV(p, nvjpeg_api()->EncoderParamsCreate(p->nv, &p->params, nullptr));
    PNV(p, nvjpeg_api()->EncoderParamsSetQuality(p->params, p->quality, nullptr));
    PNV(p, nvjpeg_api()->EncoderParamsSetSamplingFactors(p->params, NVJPEG_CSS_420, nullptr));
    return ORBB_OK;
}

extern "C" int orbb_preview_create(orbb_preview **out, int width, int height, int quality, int device) {
    if (!out || width < 16 || height < 16 || width > 16384 || height > 16384 || quality < 1 || quality > 100) return ORBB_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return ORBB_ERR_NO_DEVICE; }
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return ORBB_ERR_NO_DEVICE;
    if (device >= ndev) return ORBB_ERR_NO_DEVICE;
    orbb_preview *p = new (std::nothrow) orbb_preview();
    if (!p) return ORBB_ERR_INVALID;
    p->w = width; p->h = height; p->device = device; p->quality = quality;
    p->pitch = ((size_t)width + 255) & ~(size_t)255;
    const int rc = preview_init(p);
    if (rc) {
        fprintf(stderr, "orbb_preview_create: %s\n", p->err);
        orbb_preview_destroy(p);
        return rc;
    }
    *out = p;
    return ORBB_OK;
}

extern "C" const char *orbb_preview_last_error(const orbb_preview *p) { return p ? p->err : ""; }

// The device part (planes + overlay), asynchronous on the stream; split out so that tests can check it exactly.
static int preview_fill(orbb_preview *p, const uint8_t *d_gray, size_t gray_pitch, const void *d_xy, int xy_stride, int n_kp,
                        const int32_t *d_n_kp, cudaStream_t st) {
    dim3 grid(((p->w + 15) / 16 + 255) / 256, p->h);
    k_preview_planes<<<grid, 256, 0, st>>>(d_gray, gray_pitch, p->w, p->h, p->d_planes, p->pitch);
    PCK(p, cudaGetLastError());
    if (d_xy && n_kp > 0) {
        k_preview_overlay<<<(n_kp + 127) / 128, 128, 0, st>>>(p->d_planes + p->pitch * (size_t)p->h, p->pitch, p->w, p->h,
                                                              static_cast<const uint8_t *>(d_xy), xy_stride, n_kp, d_n_kp);
        PCK(p, cudaGetLastError());
    }
    return ORBB_OK;
}

extern "C" int orbb_preview_encode_host(orbb_preview *p, const uint8_t *d_gray, size_t gray_pitch, const void *d_xy,
                                        int xy_stride, int n_kp, const int32_t *d_n_kp, uint8_t *h_jpeg, size_t capacity,
                                        size_t *length, void *stream) {
    if (!p || !d_gray || !length || gray_pitch < (size_t)p->w || n_kp < 0 || (d_xy && (xy_stride < 8 || (xy_stride & 3))))
        return ORBB_ERR_INVALID;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    PCK(p, cudaSetDevice(p->device));
    int rc = preview_fill(p, d_gray, gray_pitch, d_xy, xy_stride, n_kp, d_n_kp, st);
    if (rc) return rc;
    nvjpegImage_t img = {};
    for (int c = 0; c < 3; ++c) { img.channel[c] = p->d_planes + (size_t)c * p->pitch * p->h; img.pitch[c] = p->pitch; }
    PNV(p, nvjpeg_api()->EncodeImage(p->nv, p->state, p->params, &img, NVJPEG_INPUT_RGB, p->w, p->h, st));
    size_t len = 0;
    PNV(p, nvjpeg_api()->EncodeRetrieveBitstream(p->nv, p->state, nullptr, &len, st));
    *length = len;
    if (!h_jpeg) return ORBB_OK;  // size query
    if (len > capacity) return ORBB_ERR_CAPACITY;
    PNV(p, nvjpeg_api()->EncodeRetrieveBitstream(p->nv, p->state, h_jpeg, &len, st));
    PCK(p, cudaStreamSynchronize(st));
    return ORBB_OK;
}

extern "C" int orbb_preview_debug_planes(orbb_preview *p, const uint8_t *d_gray, size_t gray_pitch, const void *d_xy,
                                         int xy_stride, int n_kp, const int32_t *d_n_kp, uint8_t *h_planes /*[3][h][w]*/) {
    if (!p || !d_gray || !h_planes || gray_pitch < (size_t)p->w || n_kp < 0) return ORBB_ERR_INVALID;
    PCK(p, cudaSetDevice(p->device));
    const int rc = preview_fill(p, d_gray, gray_pitch, d_xy, xy_stride, n_kp, d_n_kp, 0);
    if (rc) return rc;
    PCK(p, cudaMemcpy2D(h_planes, p->w, p->d_planes, p->pitch, p->w, (size_t)3 * p->h, cudaMemcpyDeviceToHost));
    return ORBB_OK;
}
