// ref_gpu_bench.cu -- REFERENCE-COMPILED TIMING BASELINE.  TEST/BENCH INFRASTRUCTURE ONLY, never on the product path.
//
// The reference's own front-end kernels (reference src/cuda/{gaussian_blur_3x3,pyramid,fast,nms,orb}.cu and
// src/cuda_common.cpp) compile for sm_100a once the missing cuda-samples header is stubbed
// (oracle/ref_shim_include/helper_cuda.h).  oracle/Makefile's `ref_gpu` target compiles THOSE FILES where they lie under
// /root/reference together with this driver into oracle/_ref/ref_gpu_bench (git-ignored; no reference source is copied).
// This file only (a) allocates the buffers the way the reference's slot thread does
// (src/SlamGpuPipeline/buildStream.cpp:244-332: cudaMallocPitch image + float response per level, one SoA block
// pos/score/level, angle, 32-byte and 32-bit descriptors, 64 KB LUT) and (b) issues the reference's stage sequence in
// its order (buildStream.cpp:424-460: gaussian_blur_3x3 -> pyramid_create_levels -> detect -> compute_fast_angle ->
// calc_orb, plus the SoA D2H of :462-466), timing it with CUDA events.
//
// It computes a DIFFERENT algorithm from the product (FAST-12 float SAD score on a 3x3-blurred single level, one
// keypoint per 32x32 cell, un-steered BRIEF squeezed to 32 bits), so it is a timing
// baseline only: "the reference's kernels recompiled for sm_100a", next to the product on the same frame size.
//
// usage: ref_gpu_bench <gray.raw> <w> <h> <iters>   -> one JSON line on stdout
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../SlamGpuPipeline/defines.h"
#include "fast.cuh"
#include "nms.cuh"
#include "orb.cuh"
#include "pyramid.cuh"

using namespace Jetracer;

struct Slot {
    int w, h, n_cells;
    unsigned char *d_gray; size_t gray_pitch;
    std::vector<pyramid_t> pyramid;
    unsigned char *d_lut;
    float *d_grid; float2 *d_pos; float *d_score; int *d_level;
    float *d_angle; unsigned char *d_desc_tmp; uint32_t *d_desc;
    float *h_grid;
    cudaStream_t stream;
};

static void slot_create(Slot &s, int w, int h) {
    s.w = w; s.h = h;
    s.n_cells = ((w + 31) / 32) * ((h + 31) / 32);
    checkCudaErrors(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    checkCudaErrors(cudaMallocPitch((void **)&s.d_gray, &s.gray_pitch, w, h));
    int lw = w, lh = h;
    for (int i = 0; i < PYRAMID_LEVELS; ++i) {
        pyramid_t L;
        if (i) { lw /= 2; lh /= 2; }
        L.image_width = lw; L.image_height = lh;
        checkCudaErrors(cudaMallocPitch((void **)&L.image, &L.image_pitch, lw, lh));
        checkCudaErrors(cudaMallocPitch((void **)&L.response, &L.response_pitch, lw * sizeof(float), lh));
        checkCudaErrors(cudaMemset2D(L.image, L.image_pitch, 0, lw, lh));
        checkCudaErrors(cudaMemset2D(L.response, L.response_pitch, 0, lw * sizeof(float), lh));
        s.pyramid.push_back(L);
    }
    checkCudaErrors(cudaMalloc((void **)&s.d_lut, 64 * 1024));
    checkCudaErrors(cudaMalloc((void **)&s.d_grid, s.n_cells * 4 * sizeof(float)));
    checkCudaErrors(cudaMemset(s.d_grid, 0, s.n_cells * 4 * sizeof(float)));
    s.d_pos = (float2 *)s.d_grid;
    s.d_score = s.d_grid + 2 * s.n_cells;
    s.d_level = (int *)(s.d_grid + 3 * s.n_cells);
    checkCudaErrors(cudaMalloc((void **)&s.d_angle, s.n_cells * sizeof(float)));
    checkCudaErrors(cudaMalloc((void **)&s.d_desc_tmp, s.n_cells * 32));
    checkCudaErrors(cudaMalloc((void **)&s.d_desc, s.n_cells * sizeof(uint32_t)));
    checkCudaErrors(cudaMallocHost((void **)&s.h_grid, s.n_cells * 4 * sizeof(float)));
    fast_gpu_calculate_lut(s.d_lut, FAST_MIN_ARC_LENGTH);  // default stream, as the reference does
    loadPattern();
    checkCudaErrors(cudaDeviceSynchronize());
}

enum { ST_BLUR, ST_PYR, ST_DETECT, ST_ANGLE, ST_ORB, ST_D2H, ST_N };

// the reference's stage sequence for one frame; ev (optional) receives ST_N + 1 events
static void slot_frame(Slot &s, cudaEvent_t *ev) {
    if (ev) cudaEventRecord(ev[0], s.stream);
    gaussian_blur_3x3(s.pyramid[0].image, (int)s.pyramid[0].image_pitch, s.d_gray, (int)s.gray_pitch, s.w, s.h, s.stream);
    if (ev) cudaEventRecord(ev[1], s.stream);
    pyramid_create_levels(s.pyramid, s.stream);
    if (ev) cudaEventRecord(ev[2], s.stream);
    detect(s.pyramid, s.d_lut, FAST_EPSILON, s.d_pos, s.d_score, s.d_level, s.stream);
    if (ev) cudaEventRecord(ev[3], s.stream);
    compute_fast_angle(s.d_angle, s.d_pos, s.pyramid[0].image, (int)s.pyramid[0].image_pitch, s.w, s.h, s.n_cells, s.stream);
    if (ev) cudaEventRecord(ev[4], s.stream);
    calc_orb(s.d_angle, s.d_pos, s.d_desc_tmp, s.d_desc, s.pyramid[0].image, (int)s.pyramid[0].image_pitch, s.w, s.h,
             s.n_cells, s.stream);
    if (ev) cudaEventRecord(ev[5], s.stream);
    checkCudaErrors(cudaMemcpyAsync(s.h_grid, s.d_grid, s.n_cells * 4 * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
    if (ev) cudaEventRecord(ev[6], s.stream);
}

int main(int argc, char **argv) {
    if (argc < 5) { std::fprintf(stderr, "usage: %s gray.raw w h iters\n", argv[0]); return 2; }
    const int w = std::atoi(argv[2]), h = std::atoi(argv[3]), iters = std::max(8, std::atoi(argv[4]));
    std::vector<unsigned char> img((size_t)w * h);
    FILE *f = std::fopen(argv[1], "rb");
    if (!f || std::fread(img.data(), 1, img.size(), f) != img.size()) { std::fprintf(stderr, "cannot read frame\n"); return 2; }
    std::fclose(f);

    Slot s;
    slot_create(s, w, h);
    checkCudaErrors(cudaMemcpy2D(s.d_gray, s.gray_pitch, img.data(), w, w, h, cudaMemcpyHostToDevice));

    for (int i = 0; i < 5; ++i) slot_frame(s, nullptr);
    checkCudaErrors(cudaStreamSynchronize(s.stream));
    int n_kp = 0;
    for (int i = 0; i < s.n_cells; ++i) n_kp += s.h_grid[2 * s.n_cells + i] > 0.f;

    cudaEvent_t e0, e1, ev[ST_N + 1];
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (auto &e : ev) cudaEventCreate(&e);

    // (1) one frame at a time, synchronised after each -- how the reference's slot thread runs (1 frame in flight)
    std::vector<float> lat(iters);
    for (int i = 0; i < iters; ++i) {
        cudaEventRecord(e0, s.stream);
        slot_frame(s, nullptr);
        cudaEventRecord(e1, s.stream);
        checkCudaErrors(cudaStreamSynchronize(s.stream));
        cudaEventElapsedTime(&lat[i], e0, e1);
    }
    std::sort(lat.begin(), lat.end());

    // (2) frames issued back to back on the slot's stream (launch-rate / kernel bound, no host sync between frames)
    float ms_b2b = 0.f;
    cudaEventRecord(e0, s.stream);
    for (int i = 0; i < iters; ++i) slot_frame(s, nullptr);
    cudaEventRecord(e1, s.stream);
    checkCudaErrors(cudaStreamSynchronize(s.stream));
    cudaEventElapsedTime(&ms_b2b, e0, e1);

    // (3) per-stage split (events between stages; medians over iters)
    std::vector<float> st[ST_N];
    for (int i = 0; i < iters; ++i) {
        slot_frame(s, ev);
        checkCudaErrors(cudaStreamSynchronize(s.stream));
        for (int k = 0; k < ST_N; ++k) { float t; cudaEventElapsedTime(&t, ev[k], ev[k + 1]); st[k].push_back(t); }
    }
    float med[ST_N];
    for (int k = 0; k < ST_N; ++k) { std::sort(st[k].begin(), st[k].end()); med[k] = st[k][st[k].size() / 2]; }
    checkCudaErrors(cudaGetLastError());

    std::printf("{\"w\": %d, \"h\": %d, \"levels\": %d, \"cells\": %d, \"keypoints\": %d, \"iters\": %d, "
                "\"frame_latency_us\": %.2f, \"frame_latency_best_us\": %.2f, \"back_to_back_us_per_frame\": %.2f, "
                "\"stages_us\": {\"gaussian_blur_3x3\": %.2f, \"pyramid_create_levels\": %.2f, \"detect\": %.2f, "
                "\"compute_fast_angle\": %.2f, \"calc_orb\": %.2f, \"d2h\": %.2f}}\n",
                w, h, (int)PYRAMID_LEVELS, s.n_cells, n_kp, iters, 1e3f * lat[iters / 2], 1e3f * lat[0],
                1e3f * ms_b2b / iters, 1e3f * med[ST_BLUR], 1e3f * med[ST_PYR], 1e3f * med[ST_DETECT], 1e3f * med[ST_ANGLE],
                1e3f * med[ST_ORB], 1e3f * med[ST_D2H]);
    return 0;
}
