/*
 * orb_oracle.h -- CPU ORACLE for the ORB front-end.  TEST INFRASTRUCTURE ONLY.
 *
 * This is the parity checker for the CUDA product in jetracer-orbslam2_b200/csrc.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product never links, imports or falls back to anything in oracle/.
 *
 * PARITY PINNING.  The reference repository (dsvua/jetracer-orbslam2) names
 * src_trash1/orb_extractor.cpp as its CPU ORBextractor, but that file is an 11-line stub
 * (src_trash1/orb_extractor.cpp:1-11) and the repo carries no tests or golden vectors for
 * this path, so the *reference itself* leaves parity unpinned.  The oracle therefore restates
 * upstream raulmur/ORB_SLAM2 src/ORBextractor.cc (un-vendored, no version pin in the
 * reference) plus the OpenCV primitives it calls.  Those primitives ARE pinned here: every one
 * (resize INTER_LINEAR, copyMakeBorder REFLECT_101, FAST 9_16 + NMS, GaussianBlur 7x7 s=2,
 * fastAtan2) is checked bit-exactly against the OpenCV 4.13 build in this image
 * (tests/test_oracle_vs_cv2.py), and the whole extractor is checked against an independent
 * Python/cv2 restatement whose outputs are committed under tests/golden/.
 *
 * The only reference-pinned datum on the path is the rBRIEF pattern table
 * (src/cuda/orb.cuh:39-297), shared through include/orb_pattern_31.inc.
 *
 * Stage names follow the reference's stage interface (src/cuda/{pyramid,fast,nms,orb,
 * post_processing}.cuh) and the upstream ORBextractor member functions.
 */
#ifndef ORB_ORACLE_H
#define ORB_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORBO_EDGE_THRESHOLD 19
#define ORBO_HALF_PATCH 15
#define ORBO_PATCH 31
#define ORBO_MAX_LEVELS 16

/* == cv::KeyPoint, 28 bytes */
typedef struct {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} orbo_keypoint;

typedef struct {
    int32_t nfeatures;
    float scale_factor;
    int32_t nlevels, ini_th_fast, min_th_fast;
} orbo_params;

/* candidate produced by the per-cell FAST stage: coordinates are relative to
 * (minBorderX, minBorderY) = (16,16) of the level ROI, exactly as upstream hands them to
 * DistributeOctTree. */
typedef struct {
    int32_t x, y, response;
} orbo_candidate;

typedef struct orbo_ctx orbo_ctx;

/* ---- primitives (each checked against cv2 in tests/test_oracle_vs_cv2.py) ---- */
void orbo_resize_linear_u8(const uint8_t *src, int sw, int sh, size_t sp, uint8_t *dst, int dw,
                           int dh, size_t dp);
/* fill the b-pixel frame around an (w x h) ROI that sits at (b,b) inside `padded` */
void orbo_border_reflect101(uint8_t *padded, int w, int h, size_t pitch, int b);
/* cv::FAST(window, thr, nonmax=true, TYPE_9_16); returns count, order row-major */
int orbo_fast9_window(const uint8_t *win, int cw, int ch, size_t pitch, int thr, int nms,
                      orbo_candidate *out, int max_out);
/* threshold-free arc score m (corner at t <=> m > t; response = m-1); 0 outside [3,w-3)x[3,h-3) */
void orbo_fast_score_map(const uint8_t *img, int w, int h, size_t pitch, uint8_t *score,
                         size_t score_pitch);
void orbo_gaussian_blur7(const uint8_t *src, int w, int h, size_t sp, uint8_t *dst, size_t dp);
float orbo_fast_atan2(float y, float x);
/* brute-force Hamming k-NN (k = 1 or 2) over 32-byte descriptors; ties -> lowest train idx.
 * out_idx/out_dist are [nq][2]; accept[q] = (k==2 ? d1 < ratio*d2 : 1). returns #accepted */
int orbo_match_knn(const uint8_t *q, int nq, const uint8_t *t, int nt, int k, float ratio,
                   int32_t *out_idx, int32_t *out_dist, uint8_t *accept);

/* windowed 1-NN with Hamming cutoff (reference src/cuda/post_processing.cu:92-200 semantics, ties -> lowest
 * train index): out_idx/out_dist [nq], -1 when unmatched. returns #matched */
int orbo_match_windowed(const uint8_t *q, const float *q_xy, int nq, const uint8_t *t, const float *t_xy, int nt,
                        float max_px, int max_hamming, int32_t *out_idx, int32_t *out_dist);

/* ---- extractor context (upstream ORBextractor object) ---- */
int orbo_create(orbo_ctx **out, const orbo_params *p, int width, int height);
void orbo_destroy(orbo_ctx *c);
int orbo_nlevels(const orbo_ctx *c);
/* geometry: arrays of nlevels entries */
void orbo_get_geometry(const orbo_ctx *c, int32_t *lw, int32_t *lh, float *scale,
                       float *inv_scale, int32_t *nfeat_per_level);
void orbo_get_umax(const orbo_ctx *c, int32_t *umax16);

/* ORBextractor::operator()(image, mask[ignored], keypoints, descriptors).
 * Returns the keypoint count (levels concatenated 0..n-1, upstream list order inside a level)
 * or a negative error.  out_desc is count x 32 bytes. */
int orbo_extract(orbo_ctx *c, const uint8_t *img, size_t pitch, orbo_keypoint *out_kp,
                 uint8_t *out_desc, int max_kp);

/* ---- stage-level access to the last orbo_extract / orbo_compute_pyramid call ---- */
int orbo_compute_pyramid(orbo_ctx *c, const uint8_t *img, size_t pitch);
/* padded level: (w+38) x (h+38), returns pointer to padded origin */
const uint8_t *orbo_level_padded(const orbo_ctx *c, int level, int *pw, int *ph, size_t *pitch);
/* per-cell FAST(20 -> 7 fallback) candidates of a level, upstream order. returns count */
int orbo_level_candidates(orbo_ctx *c, int level, orbo_candidate *out, int max_out);
/* DistributeOctTree on an explicit candidate list; out indices into `cand`. returns count */
int orbo_distribute_octree(const orbo_candidate *cand, int n, int minX, int maxX, int minY,
                           int maxY, int N, int32_t *out_index, int max_out);
/* blurred ROI of a level (w x h, contiguous) as used for descriptors */
int orbo_level_blurred(orbo_ctx *c, int level, uint8_t *out);
float orbo_ic_angle(const orbo_ctx *c, int level, float x, float y);
void orbo_descriptor(const uint8_t *blurred, size_t pitch, float x, float y, float angle_deg,
                     uint8_t *desc32);

/* CPU baseline helper: run orbo_extract over n_frames frames with n_threads pthreads
 * (each thread owns a context).  Returns total keypoints or negative error. */
long orbo_extract_many(const orbo_params *p, const uint8_t *frames, int w, int h, size_t pitch,
                       size_t frame_stride, int n_frames, int n_threads, int32_t *counts);
long orbo_match_many(const uint8_t *q, int nq, const uint8_t *t, int nt, int k, float ratio,
                     int n_threads, int32_t *out_idx, int32_t *out_dist, uint8_t *accept);

/* ---- RGB-D association stages (rgbd_oracle.c; reference src/cuda/cuda-align.cu, post_processing.cu) ---- */
typedef struct {
    int32_t width, height;
    float ppx, ppy, fx, fy;
    int32_t model;
    float coeffs[5];
} orbo_intrinsics; /* == rs2_intrinsics */
typedef struct {
    float rotation[9];
    float translation[3];
} orbo_extrinsics; /* == rs2_extrinsics */
void orbo_align_depth_to_other(const uint16_t *depth, float depth_scale, const orbo_intrinsics *di,
                               const orbo_intrinsics *oi, const orbo_extrinsics *ex, uint32_t *out);
int orbo_keypoint_pixel_to_point(const uint32_t *aligned, const orbo_intrinsics *oi, const orbo_keypoint *kp,
                                 const uint8_t *desc, int n, orbo_keypoint *kp_out, uint8_t *desc_out, double *points);
void orbo_reproject_points(const double *points, int n, const double *T, const orbo_intrinsics *intrin, float *pos_out);
int orbo_search_by_projection(const uint8_t *q_desc, const float *q_uv, const orbo_keypoint *q_kp, int nq,
                              const uint8_t *t_desc, const orbo_keypoint *t_kp, int nt, const float *scale_factors,
                              int n_levels, float th, int th_high, int check_orientation, int32_t *out_idx,
                              int32_t *out_dist);
int orbo_compute_stereo_matches(orbo_ctx *left, orbo_ctx *right, const orbo_keypoint *kl, const uint8_t *dl, int n_left,
                                const orbo_keypoint *kr, const uint8_t *dr, int n_right, float mbf, float fx, float *uright,
                                float *depth);
void orbo_rgb_to_grayscale(const uint8_t *src, size_t src_pitch, int cols, int rows, uint8_t *dst, size_t dst_pitch);
int orbo_compact_pairs(const int32_t *idx, int nq, const double *q_points, const double *t_points, const float *t_xy,
                       int t_xy_stride_floats, double *prev_out, double *curr_out, uint16_t *x_out, uint16_t *y_out);

#ifdef __cplusplus
}
#endif
#endif
