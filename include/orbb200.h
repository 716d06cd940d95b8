/*
 * orbb200.h -- C ABI of the B200-native ORB front-end (liborbb200.so).
 *
 * Drop-in boundary for the ORB path of dsvua/jetracer-orbslam2.  Plain C: pointers, sizes and a
 * cudaStream_t passed as void*.  No torch / OpenCV / Eigen types.  Every entry point names the
 * reference interface (file:line under the reference tree) it replaces.
 *
 * Conventions:
 *   - return 0 on success, a negative orbb_status otherwise; nothing aborts the process
 *     (the reference's CUDA_API_CALL/checkCudaErrors exit(), src/cuda_common.h:69-95);
 *   - the caller owns inputs and outputs; the handle owns all scratch, allocated once in
 *     orbb_create (the reference cudaMallocs per frame, src/SlamGpuPipeline/buildStream.cpp:354-370);
 *   - one handle per (host thread, device); all work is enqueued on the stream given, no hidden
 *     synchronisation in the *_device entry points;
 *   - there is NO CPU fallback: without a CUDA device orbb_create returns ORBB_ERR_NO_DEVICE.
 */
#ifndef ORBB200_H
#define ORBB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORBB_VERSION 100
#define ORBB_MAX_LEVELS 16
#define ORBB_EDGE_THRESHOLD 19 /* upstream EDGE_THRESHOLD; border of every padded level */
#define ORBB_DESC_BYTES 32

typedef enum {
    ORBB_OK = 0,
    ORBB_ERR_INVALID = -1,     /* bad argument */
    ORBB_ERR_NO_DEVICE = -2,   /* no CUDA device / wrong arch: the product has no CPU path */
    ORBB_ERR_CUDA = -3,        /* a CUDA runtime call failed (see orbb_last_cuda_error) */
    ORBB_ERR_TOO_SMALL = -4,   /* a pyramid level would be < 62 px (upstream divides by nCols==0) */
    ORBB_ERR_CAPACITY = -5,    /* batch/keypoint capacity of the handle or output exceeded */
    ORBB_ERR_SHAPE = -6        /* image shape differs from the one the handle was created for */
} orbb_status;

/* == cv::KeyPoint (28 bytes): what ORBextractor::operator() fills.  Replaces the SoA
 * float2 pos / float score / int level triple of src/SlamGpuPipeline/buildStream.cpp:289-296. */
typedef struct {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} orbb_keypoint;

/* ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST); replaces the compile-time
 * knobs of src/SlamGpuPipeline/defines.h:2-9 (PYRAMID_LEVELS, FAST_EPSILON, ...). */
typedef struct {
    int32_t nfeatures;
    float scale_factor;
    int32_t nlevels, ini_th_fast, min_th_fast;
} orbb_params;

/* Level descriptor handed out by orbb_get_level; mirrors Jetracer::pyramid_t
 * (src/cuda/pyramid.cuh:9-18) but with the 19-px reflect-101 frame and no float response map. */
typedef struct {
    int32_t width, height;   /* ROI size */
    int32_t pitch;           /* bytes per row of the padded buffer */
    int32_t roi_offset;      /* byte offset of ROI pixel (0,0) from `padded` row start + 19 rows */
    uint8_t *padded;         /* device pointer to padded pixel (0,0): (width+38) x (height+38) */
    uint8_t *blurred;        /* device pointer to the 7x7-blurred ROI (same pitch, ROI origin) */
    float scale, inv_scale;  /* mvScaleFactor / mvInvScaleFactor */
    int32_t nfeatures;       /* mnFeaturesPerLevel */
} orbb_level;

typedef struct orbb_handle orbb_handle;

/* ---------------------------------------------------------------- lifetime */
/* Replaces the buffer prologue of SlamGpuPipeline::buildStream (buildStream.cpp:208-341) and
 * loadPattern() (src/cuda/orb.cu:218-225).  device < 0 => current device. */
int orbb_create(orbb_handle **out, const orbb_params *params, int width, int height, int max_batch,
                int device); /* max_batch <= 65535 (ORBB_ERR_CAPACITY beyond: the frame index is a grid dimension) */
int orbb_destroy(orbb_handle *h);
const char *orbb_strerror(int status);
const char *orbb_last_cuda_error(const orbb_handle *h);

/* ---------------------------------------------------------------- geometry / getters
 * ORBextractor::GetLevels/GetScaleFactors/GetInverseScaleFactors/GetScaleSigmaSquares/
 * GetInverseScaleSigmaSquares.  Arrays must hold nlevels entries (NULL to skip). */
int orbb_get_levels(const orbb_handle *h);
int orbb_get_scale_factors(const orbb_handle *h, float *scale, float *inv_scale, float *sigma2,
                           float *inv_sigma2);
int orbb_get_features_per_level(const orbb_handle *h, int32_t *nfeat);
int orbb_max_keypoints_per_frame(const orbb_handle *h); /* capacity to allocate per frame */
/* number of CUDA kernels this handle has launched so far (bench.py reports it as gpu_launches) */
long long orbb_get_launch_count(const orbb_handle *h);
/* mvImagePyramid[level] of frame `frame` in the last batch (device pointers into the handle) */
int orbb_get_level(const orbb_handle *h, int frame, int level, orbb_level *out);

/* ---------------------------------------------------------------- whole extractor
 * ORBextractor::operator()(image, mask[ignored], keypoints, descriptors) over a batch of frames;
 * replaces the stage-calling block of buildStream (buildStream.cpp:416-460).
 *   d_images : DEVICE pointer, frame f row r at d_images + f*frame_stride + r*pitch, u8 gray
 *   d_kp     : DEVICE [n_frames][max_kp] orbb_keypoint
 *   d_desc   : DEVICE [n_frames][max_kp][32]
 *   d_counts : DEVICE [n_frames] int32
 * Keypoints of a frame are concatenated by level (0..nlevels-1); inside a level they are in
 * quadtree Z-order (deterministic).  Asynchronous on `cuda_stream`. */
int orbb_extract_batch_device(orbb_handle *h, const uint8_t *d_images, size_t pitch,
                              size_t frame_stride, int n_frames, orbb_keypoint *d_kp,
                              uint8_t *d_desc, int32_t *d_counts, int max_kp, void *cuda_stream);
/* Same with HOST buffers (pinned or pageable): H2D of the frames, extraction, D2H of
 * keypoints/descriptors/counts, then synchronises the stream.  This is the call a
 * SlamGpuPipeline slot thread makes per batch (buildStream.cpp:399-406,462-466). */
int orbb_extract_batch_host(orbb_handle *h, const uint8_t *h_images, size_t pitch,
                            size_t frame_stride, int n_frames, orbb_keypoint *h_kp, uint8_t *h_desc,
                            int32_t *h_counts, int max_kp, void *cuda_stream);

/* Asynchronous form of orbb_extract_batch_host for double-buffered capture loops: enqueues the whole
 * pipeline (chunked H2D -> kernels -> D2H on the handle's internal streams) and returns a ticket (>= 0)
 * without waiting; orbb_wait(ticket) blocks until that batch's outputs are in the host buffers.  At most
 * two batches are in flight per handle (the call blocks on ticket-2); input and output host buffers of a
 * batch must stay untouched until its ticket has been waited for.  Pinned host memory is needed for the
 * copies to overlap (pageable memory works but serialises). */
int orbb_extract_batch_host_async(orbb_handle *h, const uint8_t *h_images, size_t pitch,
                                  size_t frame_stride, int n_frames, orbb_keypoint *h_kp,
                                  uint8_t *h_desc, int32_t *h_counts, int max_kp, void *cuda_stream);
int orbb_wait(orbb_handle *h, int ticket);

/* ---------------------------------------------------------------- stage interface
 * Same stage names as the reference's free functions in namespace Jetracer.  All take the batch
 * resident in the handle (set by orbb_stage_upload or a previous stage), are async on stream and may be
 * re-run on the resident batch (orbb_detect / orbb_detect_fast restart the candidate lists). */
/* copy frames into level 0 (+ reflect-101 frame); precedes pyramid_create_levels */
int orbb_stage_upload(orbb_handle *h, const uint8_t *d_images, size_t pitch, size_t frame_stride,
                      int n_frames, void *cuda_stream);
/* Jetracer::pyramid_create_levels (src/cuda/pyramid.cuh:20-21): levels 1..n-1, x(1/scale)
 * bilinear 11-bit fixed point == cv::resize INTER_LINEAR, + border */
int orbb_pyramid_create_levels(orbb_handle *h, void *cuda_stream);
/* Jetracer::detect + grid_nms (src/cuda/fast.cuh:42-48, nms.cuh:11-15): FAST-9 score, per-cell
 * 3x3 NMS with the ini/min threshold fallback, then DistributeOctTree selection per level */
int orbb_detect(orbb_handle *h, void *cuda_stream);
/* the two halves of orbb_detect, exposed so a host stage (or bench.py) can time/overlap them:
 * FAST score + cell NMS + threshold fallback -> candidate lists; then quadtree selection */
int orbb_detect_fast(orbb_handle *h, void *cuda_stream);
int orbb_detect_distribute(orbb_handle *h, void *cuda_stream);
/* Jetracer::gaussian_blur_3x3 slot (src/cuda/orb.cuh:29-35), now 7x7 sigma=2 integer Gaussian */
int orbb_gaussian_blur(orbb_handle *h, void *cuda_stream);
/* Jetracer::compute_fast_angle + calc_orb (src/cuda/orb.cuh:9-27): IC_Angle (degrees) and the
 * 256-bit steered BRIEF; writes the final keypoint/descriptor arrays */
int orbb_compute_angle_and_orb(orbb_handle *h, orbb_keypoint *d_kp, uint8_t *d_desc,
                               int32_t *d_counts, int max_kp, void *cuda_stream);

/* The reference calls the two halves separately (call sites src/SlamGpuPipeline/buildStream.cpp:442-460), on
 * caller-owned arrays: these two entry points keep that shape, so the call site only swaps names.
 *
 * Jetracer::compute_fast_angle (src/cuda/orb.cuh:9-16, kernel orb.cu:77-142): upstream IC_Angle for n keypoints of
 * ONE image: integer moments over the 31-px disc centred on (cvRound(x), cvRound(y)), angle = cv::fastAtan2 in
 * DEGREES [0, 360) (the reference kernel returns atan2f radians).  d_pos_xy: DEVICE float2[n] pixel positions in the
 * coordinates of `d_image` (e.g. an orbb_get_level() ROI: padded + 19 * pitch + 19); a keypoint whose disc leaves the
 * image gets -1 (cv::KeyPoint's "no orientation").  Async on stream. */
int orbb_compute_fast_angle(orbb_handle *h, float *d_angle, const float *d_pos_xy, const uint8_t *d_image,
                            int image_pitch, int image_width, int image_height, int keypoints_num,
                            void *cuda_stream);
/* Jetracer::calc_orb (src/cuda/orb.cuh:18-27, kernels orb.cu:17-75 + the 32-bit squeeze :145-169, dropped): upstream
 * computeOrbDescriptor for n keypoints of ONE image that is ALREADY smoothed (orbb_gaussian_blur's output,
 * orbb_level::blurred): 256-bit steered BRIEF, d_desc DEVICE [n][32] (bit k of byte i = test 8i+k); there is no
 * d_descriptors_tmp.  d_angle in degrees as written by orbb_compute_fast_angle.  A keypoint closer than 18 px to the
 * image edge (upstream keeps >= 19) gets an all-zero descriptor.  Async on stream. */
int orbb_calc_orb(orbb_handle *h, const float *d_angle, const float *d_pos_xy, uint8_t *d_desc,
                  const uint8_t *d_blurred_image, int image_pitch, int image_width, int image_height,
                  int keypoints_num, void *cuda_stream);
/* The SoA result of Jetracer::detect (float2 d_pos / float d_score / int d_level, buildStream.cpp:289-296, 434-440)
 * for the batch resident in the handle, after orbb_detect: per frame the selected keypoints of level 0, then level
 * 1, ...; d_pos_xy [n_frames][max_kp] float2 in the coordinates of the keypoint's OWN level ROI (unscaled),
 * d_score [n_frames][max_kp] (cv::FAST response), d_level [n_frames][max_kp], d_level_counts
 * [n_frames][nlevels] (keypoints per level), d_counts [n_frames].  Any output may be NULL.  Async on stream. */
int orbb_detect_export(orbb_handle *h, float *d_pos_xy, float *d_score, int32_t *d_level, int32_t *d_level_counts,
                       int32_t *d_counts, int max_kp, void *cuda_stream);

/* ---------------------------------------------------------------- matcher
 * Replaces Jetracer::match_keypoints (src/cuda/post_processing.cuh:40-51): brute-force Hamming
 * k-NN over 256-bit descriptors, k in {1,2}, ties -> lowest train index, accept iff
 * (k==1) or d1 < ratio*d2.  d_idx/d_dist are [nq][2] int32 (second column -1 when k==1 or nt<2),
 * d_accept [nq] u8 (may be NULL), d_naccept one int32 (may be NULL).  The all-pairs distances are computed on the
 * tensor cores (descriptor bits as +1 / -1 int8: dot = 256 - 2 * Hamming; tcgen05 MMAs with TMEM accumulators by
 * default, see orbb_debug_matcher_kind for the other two kernels); every kernel gives the same results.  The handle
 * supplies the split-T scratch and the 64 MB image of the expanded train set (train sets up to 262 144 rows use it),
 * both allocated once in orbb_create (query sets larger than the scratch holds are processed in chunks): no allocation,
 * no synchronisation.  The scratch is per handle, so matcher calls on one handle must share a stream (or be ordered
 * by the caller).  Async on stream. */
int orbb_match_knn(orbb_handle *h, const uint8_t *d_query, int nq, const uint8_t *d_train, int nt,
                   int k, float ratio, int32_t *d_idx, int32_t *d_dist, uint8_t *d_accept,
                   int32_t *d_naccept, void *cuda_stream);
/* Batch form (BASELINE cfg 5: every frame of a batch against one descriptor map): d_query is an extraction output
 * [n_frames][max_kp][32] with DEVICE counts [n_frames]; all frames are matched against the same train set.  Outputs
 * keep the inputs' fixed stride -- d_idx / d_dist [n_frames][max_kp][2], d_accept [n_frames][max_kp] -- and rows
 * past a frame's count report -1 / 0, so the result can be gathered across GPUs as it is.  No host round trip of
 * the counts.  Async on stream. */
int orbb_match_knn_batch(orbb_handle *h, const uint8_t *d_query, const int32_t *d_q_counts, int n_frames,
                         int max_kp, const uint8_t *d_train, int nt, int k, float ratio, int32_t *d_idx,
                         int32_t *d_dist, uint8_t *d_accept, int32_t *d_naccept, void *cuda_stream);
/* Segmented form: nseg independent (query set, train set) pairs, e.g. left/right or t/t+1 frames.
 * q_offsets/t_offsets are DEVICE int32 [nseg+1] row offsets into d_query/d_train; idx values are
 * relative to the segment's train set.  The launch geometry comes from the HOST-side sizes the caller states
 * (nothing is read back from the device, nothing synchronises): nq_total = q_offsets[nseg], max_q_per_seg /
 * max_t_per_seg = upper bounds of the segment sizes (rows beyond them are not matched).  ORBB_ERR_CAPACITY when
 * nq_total exceeds what the handle's scratch holds (about 2.4 M + max_batch * max_kp rows). */
int orbb_match_knn_segmented(orbb_handle *h, const uint8_t *d_query, const int32_t *d_q_offsets,
                             const uint8_t *d_train, const int32_t *d_t_offsets, int nseg, int nq_total,
                             int max_q_per_seg, int max_t_per_seg, int k, float ratio, int32_t *d_idx,
                             int32_t *d_dist, uint8_t *d_accept, void *cuda_stream);

/* Windowed (guided) matcher with the reference's own match_keypoints semantics
 * (src/cuda/post_processing.cu:92-200, call site src/SlamGpuPipeline/buildStream.cpp:545-548): for every
 * query keypoint, among the train keypoints with |dx| <= max_px and |dy| <= max_px (float compare), the one
 * with the smallest Hamming distance, accepted iff distance < max_hamming; ties -> lowest train index (the
 * reference's rotated scan order makes its tie-break thread dependent; ours is fixed).  Here on the full
 * 256-bit descriptors.  Positions are read as two consecutive floats (x, y) every `xy_stride` bytes, so an
 * orbb_keypoint array (stride 28) or a packed float2 array (stride 8) can be passed directly.
 * d_idx/d_dist: [nq] int32 (-1 when unmatched); d_nmatched: one int32 (may be NULL).  Async on stream. */
int orbb_match_windowed(orbb_handle *h, const uint8_t *d_query, const void *d_query_xy, int q_xy_stride, int nq,
                        const uint8_t *d_train, const void *d_train_xy, int t_xy_stride, int nt, float max_px,
                        int max_hamming, int32_t *d_idx, int32_t *d_dist, int32_t *d_nmatched,
                        void *cuda_stream);

/* ---------------------------------------------------------------- RGB-D association
 * The step right after descriptors in SlamGpuPipeline::buildStream (buildStream.cpp:376-394, 468-487, 523-556):
 * depth aligned to the image the keypoints live in, keypoints lifted to 3-D, previous-frame points reprojected
 * and matched inside a pixel window, matched 3-D pairs compacted for the pose solver.  Arithmetic follows the
 * reference kernels (a CUDA copy of librealsense's rsutil.h) with un-fused IEEE float32/float64 operations. */

/* == rs2_intrinsics / rs2_extrinsics of librealsense2 (rs_types.h), bit-compatible, so the structs the reference
 * uploads in SlamGpuPipeline::upload_intristics (SlamGpuPipeline.cpp:60-91) can be passed as they are. */
typedef enum {
    ORBB_DISTORTION_NONE = 0,
    ORBB_DISTORTION_MODIFIED_BROWN_CONRADY = 1,
    ORBB_DISTORTION_INVERSE_BROWN_CONRADY = 2,
    ORBB_DISTORTION_FTHETA = 3, /* not supported (atan/tan chain is not reproducible bit-exactly): ORBB_ERR_INVALID */
    ORBB_DISTORTION_BROWN_CONRADY = 4,
    ORBB_DISTORTION_KANNALA_BRANDT4 = 5
} orbb_distortion;
typedef struct {
    int32_t width, height;
    float ppx, ppy, fx, fy;
    int32_t model; /* orbb_distortion */
    float coeffs[5];
} orbb_intrinsics;
typedef struct {
    float rotation[9]; /* column-major 3x3 */
    float translation[3];
} orbb_extrinsics;

/* Jetracer::align_depth_to_other (src/cuda/cuda-align.cuh:41-50, kernels cuda-align.cu:122-286, launcher :366-399):
 * every depth pixel with a non-zero value is deprojected at its top-left (-0.5) and bottom-right (+0.5) corner,
 * transformed, projected into the other image; the rectangle between the two rounded corners receives
 * min(raw depth); pixels nothing maps to are 0.  One fused scatter kernel (no int2 pixel map in HBM) + a vectorised
 * sentinel->0 pass.  d_depth: DEVICE [n_frames][depth.height][depth.width] u16; d_aligned_out: DEVICE
 * [n_frames][other.height][other.width] u32.  Intrinsics/extrinsics are HOST structs (passed by value to the
 * kernel).  Async on stream. */
int orbb_align_depth_to_other(orbb_handle *h, const uint16_t *d_depth, int n_frames, float depth_scale,
                              const orbb_intrinsics *depth_intrin, const orbb_intrinsics *other_intrin,
                              const orbb_extrinsics *depth_to_other, uint32_t *d_aligned_out, void *cuda_stream);

/* Jetracer::keypoint_pixel_to_point (src/cuda/cuda-align.cuh:52-65, kernel cuda-align.cu:282-364, launcher :401-443):
 * keep the keypoints with aligned depth > 1 and response > 1, lift them to 3-D (float64, raw depth units) and
 * compact keypoints, descriptors and points.  Differences from the reference, both deliberate: the depth is read at
 * (int(x+0.5), int(y+0.5)) -- the reference indexes the column with pos.y (cuda-align.cu:332) -- and
 * the compaction keeps input order (the reference's atomicAdd order is nondeterministic).
 * Inputs are the outputs of orbb_extract_batch_device ([n_frames][max_kp] keypoints, [..][32] descriptors,
 * [n_frames] counts); outputs have the same strides; d_points is [n_frames][max_kp][3] float64;
 * d_valid_counts [n_frames].  In-place (out == in) is NOT allowed.  Async on stream. */
int orbb_keypoint_pixel_to_point(orbb_handle *h, const uint32_t *d_aligned_depth, const orbb_intrinsics *other_intrin,
                                 int n_frames, const orbb_keypoint *d_kp_in, const uint8_t *d_desc_in,
                                 const int32_t *d_counts_in, int max_kp, orbb_keypoint *d_kp_out, uint8_t *d_desc_out,
                                 double *d_points, int32_t *d_valid_counts, void *cuda_stream);

/* kernel_reproject_prev_points (src/cuda/post_processing.cu:72-90): pos = project(T * [p;1]) with T a column-major
 * 4x4 float64 (Eigen::Matrix4d layout).  d_T: DEVICE [n_frames][16] or NULL for identity.  d_pos_out:
 * [n_frames][max_kp] float2.  Async on stream. */
int orbb_reproject_points(orbb_handle *h, const double *d_points, const int32_t *d_counts, int n_frames, int max_kp,
                          const double *d_T, const orbb_intrinsics *intrin, float *d_pos_out, void *cuda_stream);

/* kernel_match_keypoints over a batch of frame pairs (src/cuda/post_processing.cu:92-200): frame f's query set
 * (descriptors d_query[f][..], positions d_query_xy, q_counts[f] rows) against its train set; same gate and tie rule as
 * orbb_match_windowed.  Then the matched 3-D pairs are compacted in query order (the reference's atomicAdd order is
 * nondeterministic): d_prev_matched/d_curr_matched [n_frames][max_kp][3] float64 gathered from d_query_points /
 * d_train_points, d_xy_u16 [n_frames][2][max_kp] (x row then y row, uint16_t(train pos), as :193-194),
 * d_nmatched [n_frames].  Point/xy outputs may be NULL.  d_idx/d_dist: [n_frames][max_kp].  Async on stream. */
int orbb_match_windowed_batch(orbb_handle *h, const uint8_t *d_query, const float *d_query_xy, const int32_t *d_q_counts,
                              const uint8_t *d_train, const void *d_train_xy, int t_xy_stride,
                              const int32_t *d_t_counts, int n_frames, int max_kp, float max_px, int max_hamming,
                              int32_t *d_idx, int32_t *d_dist, const double *d_query_points,
                              const double *d_train_points, double *d_prev_matched, double *d_curr_matched,
                              uint16_t *d_xy_u16, int32_t *d_nmatched, void *cuda_stream);

/* Guided matcher with ORB-SLAM2's SearchByProjection gates (upstream raulmur/ORB_SLAM2
 * src/ORBmatcher.cc SearchByProjection(Frame&, const Frame&, th, bMono) + ComputeThreeMaxima, un-vendored, no pin in
 * the reference): for every query keypoint of frame f, projected to (u, v) in the train image (orbb_reproject_points),
 *   radius = th * mvScaleFactors[query octave]; candidates = train keypoints with |u - x| < radius, |v - y| < radius
 *   (strict, as Frame::GetFeaturesInArea) and query octave - 1 <= train octave <= query octave + 1;
 *   best = smallest 256-bit Hamming distance (ties -> lowest train index; upstream: grid-cell traversal order);
 *   accepted iff best <= th_high (ORBmatcher::TH_HIGH = 100);
 *   check_orientation != 0: rot = query angle - train angle (+360 if negative), bin = round(rot / 30) (30 -> 0), only
 *   matches in the three most populated bins survive (second/third bin dropped below 10 % of the first).
 * Not modelled (SLAM state, not front-end): the "train keypoint already has a map point" skip, the stereo uRight check
 * and the forward/backward octave modes.  Arrays as in orbb_match_windowed_batch; d_nmatched [n_frames] counts the
 * survivors.  Async on stream. */
int orbb_match_projection_batch(orbb_handle *h, const uint8_t *d_query_desc, const float *d_query_uv,
                                const orbb_keypoint *d_query_kp, const int32_t *d_q_counts, const uint8_t *d_train_desc,
                                const orbb_keypoint *d_train_kp, const int32_t *d_t_counts, int n_frames, int max_kp,
                                float th, int th_high, int check_orientation, int32_t *d_idx, int32_t *d_dist,
                                int32_t *d_nmatched, void *cuda_stream);

/* Stereo association of rectified pairs with ORB-SLAM2 Frame::ComputeStereoMatches semantics (SURVEY.md 8f-3;
 * upstream raulmur/ORB_SLAM2 src/Frame.cc, un-vendored; the reference has no stereo path).  The pairs are the frames
 * (2p, 2p+1) = (left, right) of the batch last extracted with this handle -- their pyramids must still be resident --
 * and d_kp / d_desc / d_counts are that extraction's outputs.  Per LEFT keypoint: d_uright and d_depth
 * ([n_pairs][max_kp] float, -1 when there is no stereo match; depth = bf / disparity), d_nstereo [n_pairs] matches
 * surviving the median-SAD filter.  bf = fx * baseline (upstream mbf).  Row band, octave band, TH_HIGH / thOrbDist,
 * 11x11 SAD sub-pixel search and parabola fit as upstream.  n_pairs * max_kp must not exceed (max_batch / 2) *
 * orbb_max_keypoints_per_frame (scratch sized in orbb_create; ORBB_ERR_CAPACITY otherwise).  Async on stream. */
int orbb_compute_stereo_matches(orbb_handle *h, const orbb_keypoint *d_kp, const uint8_t *d_desc, const int32_t *d_counts,
                                int max_kp, int n_pairs, float bf, float fx, float *d_uright, float *d_depth,
                                int32_t *d_nstereo, void *cuda_stream);

/* Jetracer::rgb_to_grayscale (src/cuda/cuda_RGB_to_Grayscale.cuh, kernel cuda_RGB_to_Grayscale.cu:10-24, call site
 * buildStream.cpp:416-422) for a batch: gray = floor((B*0.07 + G*0.72 + R*0.21) + 0.5) in float64, every operation
 * rounded on its own, interleaved RGB8 in.  DEVICE pointers; rgb_pitch must be a multiple of 4.  Async on stream.
 * The output can be fed to orbb_extract_batch_device / orbb_stage_upload. */
int orbb_rgb_to_grayscale(orbb_handle *h, const uint8_t *d_rgb, size_t rgb_pitch, size_t rgb_frame_stride, int width,
                          int height, int n_frames, uint8_t *d_gray, size_t gray_pitch, size_t gray_frame_stride,
                          void *cuda_stream);

/* ---------------------------------------------------------------- RGB-D frame stage
 * The per-frame body of SlamGpuPipeline::buildStream (src/SlamGpuPipeline/buildStream.cpp:345-660) rewritten around
 * the handle, for batches of consecutive frames of one camera stream: pinned H2D, depth alignment on its own stream
 * next to the extraction (as :376-394 / :399-460), depth gate + 3-D lift (:468-487), reprojection + windowed match
 * against the previous frame + pair compaction (:523-556), results D2H into stage-owned pinned memory.  Nothing is
 * allocated per frame and nothing synchronises except orbb_rgbd_stage_wait (the reference cudaMallocs 3+5 buffers
 * and synchronises four times per frame).  Frame f of a batch is matched against frame f-1; frame 0 against the last
 * frame of the previous batch (kept on the device), or against nothing after create/reset.  JPEG preview, overlay and
 * the (disabled) pose maths of the reference are out of scope. */
typedef struct orbb_rgbd_stage orbb_rgbd_stage;
typedef struct {
    orbb_params orb;
    int32_t max_batch;
    orbb_intrinsics depth_intrin;   /* depth frames are depth_intrin.width x height, u16 */
    orbb_intrinsics image_intrin;   /* gray frames are image_intrin.width x height, u8; keypoints live here */
    orbb_extrinsics depth_to_image;
    float depth_scale;              /* rs2 depth units -> metres; only "is it zero" matters (cuda-align.cu:139) */
    float max_pixel_distance;       /* match_keypoints gate, 2 in the reference (buildStream.cpp:545-548) */
    int32_t max_hamming_distance;   /* 4 of 32 bits in the reference; here out of 256 */
} orbb_rgbd_config;
/* == the per-frame fields of slam_frame_t (src/SlamGpuPipeline/types.h:25-65) for a batch; HOST pointers into
 * stage-owned pinned memory, valid until the second-next submit. Row stride of every per-keypoint array: max_kp. */
typedef struct {
    int32_t n_frames, max_kp;
    const int32_t *keypoints_count;       /* [n] extracted keypoints (slam_frame_t::keypoints_count) */
    const int32_t *valid_keypoints_num;   /* [n] after the depth gate (h_valid_keypoints_num) */
    const int32_t *matched_keypoints_num; /* [n] matched against the previous frame (h_matched_keypoints_num) */
    const orbb_keypoint *keypoints;       /* [n][max_kp] depth-gated keypoints (d_pos + score + level) */
    const uint8_t *descriptors;           /* [n][max_kp][32] (256-bit, not the reference's 32-bit squeeze) */
    const double *points;                 /* [n][max_kp][3] (h_points) */
    const double *previous_matched_points, *current_matched_points; /* [n][max_kp][3] */
    const uint16_t *matched_xy;           /* [n][2][max_kp]: keypoints_x row, keypoints_y row */
} orbb_slam_frames;
int orbb_rgbd_stage_create(orbb_rgbd_stage **out, const orbb_rgbd_config *cfg, int device);
int orbb_rgbd_stage_destroy(orbb_rgbd_stage *s);
/* forget the previous frame (start of a new sequence) */
int orbb_rgbd_stage_reset(orbb_rgbd_stage *s);
/* the extractor handle inside the stage (getters, launch count, last CUDA error) */
orbb_handle *orbb_rgbd_stage_handle(orbb_rgbd_stage *s);
/* Enqueue n_frames consecutive frames: HOST gray [n][h][w] u8, HOST depth [n][dh][dw] u16 (pinned memory lets the
 * copies overlap the previous batch's kernels), h_T = per-frame T_w2c_prev_curr, [n][16] column-major float64
 * (Eigen::Matrix4d) or NULL for identity (the reference forces identity, buildStream.cpp:538).  Returns a ticket.
 * At most two batches are in flight; input buffers must stay untouched until the ticket has been waited for. */
int orbb_rgbd_stage_submit(orbb_rgbd_stage *s, const uint8_t *h_gray, const uint16_t *h_depth, int n_frames,
                           const double *h_T);
int orbb_rgbd_stage_wait(orbb_rgbd_stage *s, int ticket, orbb_slam_frames *out);

/* ---------------------------------------------------------------- wire format
 * The BSON document the reference sends per processed frame (src/WebSocket/WebSocketCom.cpp:164-184, writer
 * src/WebSocket/bson.cpp:46-130): int32 ax, ay, az, width, height, channels; binary (subtype 0x80) keypoints_x,
 * keypoints_y (uint16 each, n_matched of them: orbb_slam_frames::matched_xy rows) and image (the caller's JPEG bytes,
 * may be empty).  Host-only; byte-identical to the reference's Bson class.  orbb_slam_frame_bson_size gives the
 * exact message size; orbb_slam_frame_to_bson returns the bytes written or a negative orbb_status. */
size_t orbb_slam_frame_bson_size(int n_matched, size_t image_bytes);
long long orbb_slam_frame_to_bson(int32_t ax, int32_t ay, int32_t az, int32_t width, int32_t height, int32_t channels,
                                  const uint16_t *keypoints_x, const uint16_t *keypoints_y, int n_matched,
                                  const uint8_t *image, size_t image_bytes, uint8_t *out, size_t out_capacity);

/* ---------------------------------------------------------------- JPEG preview
 * The image the reference attaches to every frame message: its gray frame as three equal planes, the keypoints painted
 * into the G plane (Jetracer::overlay_keypoints, src/cuda/post_processing.cu:45-70: the 2x2 block
 * [int(x-1), x+1) x [int(y-1), y+1) set to 255), encoded by nvJPEG with quality 90 and 4:2:0 subsampling
 * (src/SlamGpuPipeline/buildStream.cpp:266-277, 491-521, 613-621).  The planes and the encoder state live in the
 * object (the reference mallocs per frame); libnvjpeg is opened with dlopen on first use, so the rest of the library
 * does not depend on it (ORBB_ERR_NO_DEVICE from create when it is missing).  The bytes go into
 * orbb_slam_frame_to_bson(image = ...). */
typedef struct orbb_preview orbb_preview;
int orbb_preview_create(orbb_preview **out, int width, int height, int quality, int device);
int orbb_preview_destroy(orbb_preview *p);
const char *orbb_preview_last_error(const orbb_preview *p);
/* d_gray: DEVICE u8 frame (row pitch gray_pitch).  d_xy: DEVICE keypoint positions, two floats (x, y) every xy_stride
 * bytes -- an orbb_keypoint array (stride 28) or packed float2 (stride 8) -- or NULL for no overlay; n_kp = number of
 * positions, further limited by *d_n_kp (DEVICE int32, e.g. one entry of the extraction's counts) when that is not
 * NULL.  Writes at most `capacity` bytes of JPEG to HOST memory and its length to *length (h_jpeg == NULL: only the
 * length).  Blocks until the bitstream is on the host (the encoder needs the stream drained, as in the reference). */
int orbb_preview_encode_host(orbb_preview *p, const uint8_t *d_gray, size_t gray_pitch, const void *d_xy, int xy_stride,
                             int n_kp, const int32_t *d_n_kp, uint8_t *h_jpeg, size_t capacity, size_t *length,
                             void *cuda_stream);
/* parity access: the three planes handed to the encoder, HOST [3][height][width].  Synchronises. */
int orbb_preview_debug_planes(orbb_preview *p, const uint8_t *d_gray, size_t gray_pitch, const void *d_xy, int xy_stride,
                              int n_kp, const int32_t *d_n_kp, uint8_t *h_planes);

/* ---------------------------------------------------------------- debug / parity access
 * Download stage outputs of frame `frame` of the last batch to HOST memory (synchronises). */
/* fill every stateless scratch buffer of the handle with `value` (0..255): extraction results must not change
 * (stand-in for compute-sanitizer initcheck where the sanitizer cannot run).  Synchronises. */
int orbb_debug_poison(orbb_handle *h, int value);
/* measured POPC issue rate (lanes per clock per SM) of this device: the matcher's roofline denominator */
int orbb_debug_popc_rate(orbb_handle *h, double *popc_per_clk_per_sm);
/* int8 tensor-core MMAs (m16n8k32) per clock per SM of this GPU, from a register-only microbenchmark: the roof of the
 * tensor-core matcher (16 descriptor pairs per MMA).  Synchronises. */
int orbb_debug_imma_rate(orbb_handle *h, double *imma_per_clk_per_sm);
/* which brute-force matcher kernel this process runs (read once from the environment): 0 = XOR / POPC (ORBB_MATCH_POPC=1),
 * 1 = warp-level int8 MMA (ORBB_MATCH_UMMA=0), 2 = tcgen05 int8 MMA with TMEM accumulators (default).  Same results. */
int orbb_debug_matcher_kind(void);
/* padded level, contiguous (w+38) x (h+38) */
int orbb_debug_get_padded(orbb_handle *h, int frame, int level, uint8_t *host_out);
/* blurred ROI, contiguous w x h */
int orbb_debug_get_blurred(orbb_handle *h, int frame, int level, uint8_t *host_out);
/* FAST arc score map of the ROI (w x h): m if m > min(ini,min) threshold else 0 (0 outside the
 * tested range [19,w-19) x [19,h-19)); recomputed by the same device code as orbb_detect */
int orbb_debug_get_scores(orbb_handle *h, int frame, int level, uint8_t *host_out);
/* per-cell FAST candidates handed to the quadtree: triples (x,y,response), coordinates relative
 * to (16,16) like upstream vToDistributeKeys; order unspecified. returns count (or <0) */
int orbb_debug_get_candidates(orbb_handle *h, int frame, int level, int32_t *host_xyr, int max_n);
/* quadtree-selected keys of a level: triples (x,y,response) relative to (16,16). returns count */
int orbb_debug_get_selected(orbb_handle *h, int frame, int level, int32_t *host_xyr, int max_n);
/* run only DistributeOctTree on a caller-supplied candidate list (host triples) for the geometry
 * of `level` with quota N; returns count, writes selected triples */
int orbb_debug_distribute(orbb_handle *h, int level, const int32_t *host_xyr, int n, int quota,
                          int32_t *host_out_xyr, int max_out);

#ifdef __cplusplus
}
#endif
#endif /* ORBB200_H */
