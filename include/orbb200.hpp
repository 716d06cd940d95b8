// orbb200.hpp -- header-only C++ host mirror over the C ABI (include/orbb200.h).
//
// The reference's host code is C++ (SlamGpuPipeline::buildStream, reference
// src/SlamGpuPipeline/buildStream.cpp:190-680), so this is what a maintainer includes:
//   * orbb200::ORBextractor -- the upstream ORB-SLAM2 surface the reference names in
//     src_trash1/orb_extractor.cpp:6-8: ctor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST),
//     operator()(image, mask, keypoints, descriptors), GetLevels/GetScaleFactors/... getters.
//     OpenCV headers are optional (define ORBB_WITH_OPENCV to get the cv::InputArray overload).
//   * namespace Jetracer -- the stage functions with the reference's names
//     (src/cuda/pyramid.cuh:20-21, fast.cuh:42-48, nms.cuh:11-15, orb.cuh:9-37, post_processing.cuh:40-51),
//     taking the handle (which owns what the reference passed as std::vector<pyramid_t> + raw buffers).
// Errors become std::runtime_error (the reference aborts the process, src/cuda_common.h:69-95).
#ifndef ORBB200_HPP
#define ORBB200_HPP

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "orbb200.h"

#ifdef ORBB_WITH_OPENCV
#include <opencv2/core.hpp>
#endif

namespace orbb200 {

using KeyPoint = orbb_keypoint;  // layout-identical to cv::KeyPoint (28 bytes)

inline void check(int rc, const orbb_handle *h, const char *what) {
    if (rc >= 0) return;
    std::string msg = std::string(what) + ": " + orbb_strerror(rc);
    if (rc == ORBB_ERR_CUDA && h) msg += std::string(" (") + orbb_last_cuda_error(h) + ")";
    throw std::runtime_error(msg);
}

class ORBextractor {
public:
    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, int width, int height,
                 int max_batch = 1, int device = -1)
        : width_(width), height_(height), max_batch_(max_batch) {
        orbb_params p{nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST};
        check(orbb_create(&h_, &p, width, height, max_batch, device), nullptr, "orbb_create");
        nlevels_ = orbb_get_levels(h_);
        scale_factor_ = scaleFactor;
        max_kp_ = orbb_max_keypoints_per_frame(h_);
        mvScaleFactor.resize(nlevels_); mvInvScaleFactor.resize(nlevels_);
        mvLevelSigma2.resize(nlevels_); mvInvLevelSigma2.resize(nlevels_);
        check(orbb_get_scale_factors(h_, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(),
                                     mvInvLevelSigma2.data()), h_, "orbb_get_scale_factors");
        kp_.resize((size_t)max_kp_ * max_batch_);
        desc_.resize((size_t)max_kp_ * max_batch_ * ORBB_DESC_BYTES);
        counts_.resize(max_batch_);
    }
    ~ORBextractor() { orbb_destroy(h_); }
    ORBextractor(const ORBextractor &) = delete;
    ORBextractor &operator=(const ORBextractor &) = delete;

    // ORBextractor::operator()(image, mask, keypoints, descriptors); mask is ignored (as upstream).
    // descriptors: keypoints.size() x 32 bytes, row-major.
    void operator()(const uint8_t *image, size_t pitch, const uint8_t * /*mask*/, std::vector<KeyPoint> &keypoints,
                    std::vector<uint8_t> &descriptors, void *cuda_stream = nullptr) {
        check(orbb_extract_batch_host(h_, image, pitch, pitch * (size_t)height_, 1, kp_.data(), desc_.data(),
                                      counts_.data(), max_kp_, cuda_stream), h_, "orbb_extract_batch_host");
        const int n = counts_[0];
        keypoints.assign(kp_.begin(), kp_.begin() + n);
        descriptors.assign(desc_.begin(), desc_.begin() + (size_t)n * ORBB_DESC_BYTES);
    }
    // batch form: frames [n][height][pitch]; outputs per frame
    void extract(const uint8_t *frames, size_t pitch, size_t frame_stride, int n,
                 std::vector<std::vector<KeyPoint>> &keypoints, std::vector<std::vector<uint8_t>> &descriptors,
                 void *cuda_stream = nullptr) {
        check(orbb_extract_batch_host(h_, frames, pitch, frame_stride, n, kp_.data(), desc_.data(), counts_.data(),
                                      max_kp_, cuda_stream), h_, "orbb_extract_batch_host");
        keypoints.resize(n); descriptors.resize(n);
        for (int f = 0; f < n; ++f) {
            const size_t o = (size_t)f * max_kp_;
            keypoints[f].assign(kp_.begin() + o, kp_.begin() + o + counts_[f]);
            descriptors[f].assign(desc_.begin() + o * 32, desc_.begin() + (o + counts_[f]) * 32);
        }
    }
#ifdef ORBB_WITH_OPENCV
    void operator()(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint> &keypoints,
                    cv::OutputArray descriptors) {
        (void)mask;
        cv::Mat im = image.getMat();
        if (im.empty()) return;
        CV_Assert(im.type() == CV_8UC1 && im.cols == width_ && im.rows == height_);
        std::vector<KeyPoint> kp; std::vector<uint8_t> d;
        (*this)(im.data, im.step, nullptr, kp, d);
        keypoints.resize(kp.size());
        for (size_t i = 0; i < kp.size(); ++i)
            keypoints[i] = cv::KeyPoint(kp[i].x, kp[i].y, kp[i].size, kp[i].angle, kp[i].response, kp[i].octave, kp[i].class_id);
        if (kp.empty()) { descriptors.release(); return; }
        descriptors.create((int)kp.size(), 32, CV_8U);
        std::memcpy(descriptors.getMat().data, d.data(), d.size());
    }
#endif
    int GetLevels() const { return nlevels_; }
    float GetScaleFactor() const { return scale_factor_; }
    std::vector<float> GetScaleFactors() const { return mvScaleFactor; }
    std::vector<float> GetInverseScaleFactors() const { return mvInvScaleFactor; }
    std::vector<float> GetScaleSigmaSquares() const { return mvLevelSigma2; }
    std::vector<float> GetInverseScaleSigmaSquares() const { return mvInvLevelSigma2; }
    // mvImagePyramid[level] of the last batch (device memory, padded layout)
    orbb_level ImagePyramidLevel(int level, int frame = 0) const {
        orbb_level l{};
        check(orbb_get_level(h_, frame, level, &l), h_, "orbb_get_level");
        return l;
    }
    orbb_handle *handle() const { return h_; }
    int max_keypoints() const { return max_kp_; }

private:
    orbb_handle *h_ = nullptr;
    int width_, height_, max_batch_, nlevels_ = 0, max_kp_ = 0;
    float scale_factor_ = 1.2f;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
    std::vector<KeyPoint> kp_;
    std::vector<uint8_t> desc_;
    std::vector<int32_t> counts_;
};

// The SlamGpuPipeline slot body (reference src/SlamGpuPipeline/buildStream.cpp:345-660) as an object: what the slot
// thread owns instead of its hand-allocated buffers, streams and the per-frame slam_frame_t mallocs.
class RgbdFrameStage {
public:
    RgbdFrameStage(const orbb_rgbd_config &cfg, int device = -1) {
        check(orbb_rgbd_stage_create(&s_, &cfg, device), nullptr, "orbb_rgbd_stage_create");
    }
    ~RgbdFrameStage() { orbb_rgbd_stage_destroy(s_); }
    RgbdFrameStage(const RgbdFrameStage &) = delete;
    RgbdFrameStage &operator=(const RgbdFrameStage &) = delete;
    // enqueue a batch of consecutive frames (pinned host memory); returns a ticket
    int submit(const uint8_t *gray, const uint16_t *depth, int n_frames, const double *T_w2c_prev_curr = nullptr) {
        const int t = orbb_rgbd_stage_submit(s_, gray, depth, n_frames, T_w2c_prev_curr);
        check(t, orbb_rgbd_stage_handle(s_), "orbb_rgbd_stage_submit");
        return t;
    }
    // blocks until the batch is in host memory; the views stay valid until the second-next submit
    orbb_slam_frames wait(int ticket) {
        orbb_slam_frames f{};
        check(orbb_rgbd_stage_wait(s_, ticket, &f), orbb_rgbd_stage_handle(s_), "orbb_rgbd_stage_wait");
        return f;
    }
    void reset() { check(orbb_rgbd_stage_reset(s_), orbb_rgbd_stage_handle(s_), "orbb_rgbd_stage_reset"); }
    orbb_handle *handle() const { return orbb_rgbd_stage_handle(s_); }

private:
    orbb_rgbd_stage *s_ = nullptr;
};

}  // namespace orbb200

// ---- the reference's stage names (namespace Jetracer), async on the given stream -------------------------
namespace Jetracer {
inline void upload_frames(orbb_handle *h, const uint8_t *d_images, size_t pitch, size_t frame_stride, int n_frames,
                          void *stream) {
    orbb200::check(orbb_stage_upload(h, d_images, pitch, frame_stride, n_frames, stream), h, "upload_frames");
}
inline void pyramid_create_levels(orbb_handle *h, void *stream) {
    orbb200::check(orbb_pyramid_create_levels(h, stream), h, "pyramid_create_levels");
}
inline void detect(orbb_handle *h, void *stream) { orbb200::check(orbb_detect(h, stream), h, "detect"); }
inline void gaussian_blur(orbb_handle *h, void *stream) { orbb200::check(orbb_gaussian_blur(h, stream), h, "gaussian_blur"); }
// compute_fast_angle + calc_orb fused: angles (degrees) and 256-bit descriptors into caller-owned device arrays
inline void compute_fast_angle_and_calc_orb(orbb_handle *h, orbb_keypoint *d_kp, uint8_t *d_desc, int32_t *d_counts,
                                            int max_kp, void *stream) {
    orbb200::check(orbb_compute_angle_and_orb(h, d_kp, d_desc, d_counts, max_kp, stream), h, "calc_orb");
}
inline void match_keypoints(orbb_handle *h, const uint8_t *d_query, int nq, const uint8_t *d_train, int nt, int k,
                            float ratio, int32_t *d_idx, int32_t *d_dist, uint8_t *d_accept, int32_t *d_naccept,
                            void *stream) {
    orbb200::check(orbb_match_knn(h, d_query, nq, d_train, nt, k, ratio, d_idx, d_dist, d_accept, d_naccept, stream), h,
                   "match_keypoints");
}
// the reference's own gate (post_processing.cuh:40-51: max_pixel_distance, max_hamming_distance): keypoint arrays
// are passed as they are (positions are the first two floats of every orbb_keypoint)
inline void match_keypoints(orbb_handle *h, const uint8_t *d_desc_prev, const orbb_keypoint *d_kp_prev, int n_prev,
                            const uint8_t *d_desc_curr, const orbb_keypoint *d_kp_curr, int n_curr,
                            int max_pixel_distance, int max_hamming_distance, int32_t *d_idx, int32_t *d_dist,
                            int32_t *d_keypoints_num_matched, void *stream) {
    orbb200::check(orbb_match_windowed(h, d_desc_prev, d_kp_prev, (int)sizeof(orbb_keypoint), n_prev, d_desc_curr,
                                       d_kp_curr, (int)sizeof(orbb_keypoint), n_curr, (float)max_pixel_distance,
                                       max_hamming_distance, d_idx, d_dist, d_keypoints_num_matched, stream),
                   h, "match_keypoints (windowed)");
}
// src/cuda/cuda-align.cuh:41-50: the int2 pixel map and the device copies of the intrinsics are gone
inline void align_depth_to_other(orbb_handle *h, uint32_t *d_aligned_out, const uint16_t *d_depth_in, float depth_scale,
                                 const orbb_intrinsics &depth_intrin, const orbb_intrinsics &other_intrin,
                                 const orbb_extrinsics &depth_to_other, int n_frames, void *stream) {
    orbb200::check(orbb_align_depth_to_other(h, d_depth_in, n_frames, depth_scale, &depth_intrin, &other_intrin,
                                             &depth_to_other, d_aligned_out, stream), h, "align_depth_to_other");
}
// src/cuda/cuda-align.cuh:52-65: counts stay on the device (the reference round-trips h_valid_keypoints_num)
inline void keypoint_pixel_to_point(orbb_handle *h, const uint32_t *d_aligned_depth, const orbb_intrinsics &rgb_intrin,
                                    int n_frames, const orbb_keypoint *d_kp_in, const uint8_t *d_descriptors_in,
                                    const int32_t *d_keypoints_num, int max_kp, orbb_keypoint *d_kp_out,
                                    uint8_t *d_descriptors_out, double *d_points, int32_t *d_valid_keypoints_num,
                                    void *stream) {
    orbb200::check(orbb_keypoint_pixel_to_point(h, d_aligned_depth, &rgb_intrin, n_frames, d_kp_in, d_descriptors_in,
                                                d_keypoints_num, max_kp, d_kp_out, d_descriptors_out, d_points,
                                                d_valid_keypoints_num, stream), h, "keypoint_pixel_to_point");
}

// ---- the reference's OWN signatures (src/cuda/orb.cuh:9-37), for a call site that is left as it is -------
// The reference's free functions take no context; the handle they need here is bound once per slot thread
// (thread-local, like the per-slot buffers of buildStream.cpp:208-341) with Jetracer::bind(handle).  Only compiled
// where the CUDA vector types are visible (the reference's translation units include <cuda_runtime.h>).
inline orbb_handle *&bound_handle() {
    static thread_local orbb_handle *h = nullptr;
    return h;
}
inline void bind(orbb_handle *h) { bound_handle() = h; }
// src/cuda/orb.cuh:37: the pattern is uploaded by orbb_create; kept so the call in SlamGpuPipeline.cpp:52 still links
inline void loadPattern() {}
#if defined(__VECTOR_TYPES_H__) || defined(__CUDACC__)
// src/cuda/orb.cuh:9-16.  Angles are DEGREES (upstream IC_Angle + fastAtan2); the reference kernel wrote radians
// and then treated them as degrees, so no caller depends on the unit.
inline void compute_fast_angle(float *d_keypoints_angle, float2 *d_keypoints_pos, unsigned char *image, int image_pitch,
                               int image_width, int image_height, int keypoints_num, cudaStream_t stream) {
    orbb200::check(orbb_compute_fast_angle(bound_handle(), d_keypoints_angle, reinterpret_cast<const float *>(d_keypoints_pos),
                                           image, image_pitch, image_width, image_height, keypoints_num, stream),
                   bound_handle(), "compute_fast_angle");
}
// src/cuda/orb.cuh:18-27.  d_descriptors receives EIGHT uint32_t per keypoint (the full 256 bits) instead of the
// reference's one; d_descriptors_tmp is unused (may be nullptr); `image` must be the smoothed image.
inline void calc_orb(float *d_keypoints_angle, float2 *d_keypoints_pos, unsigned char * /*d_descriptors_tmp*/,
                     uint32_t *d_descriptors, unsigned char *image, int image_pitch, int image_width, int image_height,
                     int keypoints_num, cudaStream_t stream) {
    orbb200::check(orbb_calc_orb(bound_handle(), d_keypoints_angle, reinterpret_cast<const float *>(d_keypoints_pos),
                                 reinterpret_cast<uint8_t *>(d_descriptors), image, image_pitch, image_width, image_height,
                                 keypoints_num, stream),
                   bound_handle(), "calc_orb");
}
#endif
}  // namespace Jetracer

#endif  // ORBB200_HPP
