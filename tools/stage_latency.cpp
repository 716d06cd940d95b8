// One RGB-D frame through orbb_rgbd_stage_submit + orbb_rgbd_stage_wait from a C++ host, the way the reference's
// pipeline thread would call it (no Python wrapper in the timed region).  Frames come from a file written by
// tools/stage_latency_probe.py: 4 gray frames (u8) followed by 4 depth frames (u16).
//   g++ -O2 -std=c++17 -Iinclude -I/usr/local/cuda/include tools/stage_latency.cpp -o tools/_build/stage_latency \
//       -Ljetracer-orbslam2_b200 -lorbb200 -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/jetracer-orbslam2_b200
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#include "orbb200.h"

int main(int argc, char **argv) {
    if (argc < 5) { std::fprintf(stderr, "usage: %s frames.bin width height nfeatures [iters]\n", argv[0]); return 2; }
    const int w = std::atoi(argv[2]), h = std::atoi(argv[3]), nf = std::atoi(argv[4]), iters = argc > 5 ? std::atoi(argv[5]) : 200;
    const size_t fb = (size_t)w * h, db = 2 * fb;
    uint8_t *gray = nullptr;
    uint16_t *depth = nullptr;
    if (cudaHostAlloc((void **)&gray, 4 * fb, cudaHostAllocDefault) != cudaSuccess ||
        cudaHostAlloc((void **)&depth, 4 * db, cudaHostAllocDefault) != cudaSuccess) return 3;
    FILE *f = std::fopen(argv[1], "rb");
    if (!f || std::fread(gray, 1, 4 * fb, f) != 4 * fb || std::fread(depth, 1, 4 * db, f) != 4 * db) return 4;
    std::fclose(f);
    orbb_rgbd_config cfg;
    std::memset(&cfg, 0, sizeof cfg);
    cfg.orb = {nf, 1.2f, 8, 20, 7};
    cfg.max_batch = 1;
    cfg.depth_intrin = {w, h, w * 0.5f + 3.7f, h * 0.5f - 2.2f, 0.502f * w, 0.502f * w, ORBB_DISTORTION_BROWN_CONRADY, {0, 0, 0, 0, 0}};
    cfg.image_intrin = {w, h, w * 0.5f - 5.1f, h * 0.5f + 4.3f, 0.72f * w, 0.725f * w, ORBB_DISTORTION_INVERSE_BROWN_CONRADY, {0, 0, 0, 0, 0}};
    const float rot[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, tr[3] = {0.0148f, 0.0002f, 0.0003f};
    std::memcpy(cfg.depth_to_image.rotation, rot, sizeof rot);
    std::memcpy(cfg.depth_to_image.translation, tr, sizeof tr);
    cfg.depth_scale = 0.001f; cfg.max_pixel_distance = 2.0f; cfg.max_hamming_distance = 64;
    orbb_rgbd_stage *s = nullptr;
    int rc = orbb_rgbd_stage_create(&s, &cfg, 0);
    if (rc) { std::fprintf(stderr, "create: %d\n", rc); return 5; }
    orbb_slam_frames out;
    for (int i = 0; i < 10; ++i) {
        const int t = orbb_rgbd_stage_submit(s, gray + (i % 4) * fb, depth + (i % 4) * fb, 1, nullptr);
        if (t < 0 || orbb_rgbd_stage_wait(s, t, &out)) return 6;
    }
    std::vector<double> lat, sub;
    for (int i = 0; i < iters; ++i) {
        const auto t0 = std::chrono::steady_clock::now();
        const int t = orbb_rgbd_stage_submit(s, gray + (i % 4) * fb, depth + (i % 4) * fb, 1, nullptr);
        const auto t1 = std::chrono::steady_clock::now();
        if (t < 0 || orbb_rgbd_stage_wait(s, t, &out)) return 6;
        const auto t2 = std::chrono::steady_clock::now();
        lat.push_back(std::chrono::duration<double, std::micro>(t2 - t0).count());
        sub.push_back(std::chrono::duration<double, std::micro>(t1 - t0).count());
    }
    std::sort(lat.begin(), lat.end()); std::sort(sub.begin(), sub.end());
    std::printf("%dx%d %d kp (C++ host): one RGB-D frame submit+wait median %.1f us, best %.1f us (host issue inside submit %.1f us), "
                "keypoints %d, valid %d, matched %d\n", w, h, nf, lat[lat.size() / 2], lat[0], sub[sub.size() / 2],
                out.keypoints_count[0], out.valid_keypoints_num[0], out.matched_keypoints_num[0]);
    orbb_rgbd_stage_destroy(s);
    return 0;
}
