/*
 * rgbd_oracle.c -- CPU ORACLE for the RGB-D association stages.  TEST INFRASTRUCTURE ONLY (see orb_oracle.h:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may load it; the product never does).
 *
 * Restates, in plain C with the expressions written as the reference writes them:
 *   align_depth_to_other       reference src/cuda/cuda-align.cu:122-286 (kernels), :366-399 (launcher)
 *   keypoint_pixel_to_point    reference src/cuda/cuda-align.cu:282-364
 *   reproject_prev_points      reference src/cuda/post_processing.cu:10-43, 72-90
 *   matched-pair compaction    reference src/cuda/post_processing.cu:176-197
 * which are themselves copies of librealsense2's rsutil.h (rs2_project_point_to_pixel, rs2_deproject_pixel_to_point,
 * rs2_transform_point_to_point; librealsense 2.42 per the reference Dockerfile:60-75, not vendored).
 *
 * PARITY PINNING: the reference has no test or golden vector for these stages, and its own build lets nvcc contract
 * a*b+c into FMAs, so the low bits of its float results are not pinned by anything.  This oracle fixes them as
 * "every operation rounded on its own, C evaluation order" (compiled with -ffp-contract=off), which is what
 * librealsense's CPU rsutil.h computes on x86-64.  Deliberate differences from the reference:
 * the keypoint depth lookup uses (x, y) instead of (y, y); compactions keep input order.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "orb_oracle.h"

enum { MODEL_NONE = 0, MODEL_MODIFIED_BC = 1, MODEL_INVERSE_BC = 2, MODEL_FTHETA = 3, MODEL_BC = 4 };

/* CUDA float->int conversion (cvt.rzi.s32.f32): truncate, saturate, NaN -> 0 */
static int f2i_cuda(float v) {
    if (v != v) return 0;
    if (v >= 2147483648.0f) return INT32_MAX;
    if (v <= -2147483648.0f) return INT32_MIN;
    return (int)v;
}

/* cuda-align.cu:26-56 */
static void project_point_to_pixel(float pixel[2], const orbo_intrinsics *intrin, const float point[3]) {
    float x = point[0] / point[2], y = point[1] / point[2];
    if (intrin->model == MODEL_MODIFIED_BC) {
        float r2 = x * x + y * y;
        float f = 1 + intrin->coeffs[0] * r2 + intrin->coeffs[1] * r2 * r2 + intrin->coeffs[4] * r2 * r2 * r2;
        x *= f;
        y *= f;
        float dx = x + 2 * intrin->coeffs[2] * x * y + intrin->coeffs[3] * (r2 + 2 * x * x);
        float dy = y + 2 * intrin->coeffs[3] * x * y + intrin->coeffs[2] * (r2 + 2 * y * y);
        x = dx;
        y = dy;
    }
    pixel[0] = x * intrin->fx + intrin->ppx;
    pixel[1] = y * intrin->fy + intrin->ppy;
}

/* cuda-align.cu:58-83 */
static void deproject_pixel_to_point(float point[3], const orbo_intrinsics *intrin, const float pixel[2], float depth) {
    float x = (pixel[0] - intrin->ppx) / intrin->fx;
    float y = (pixel[1] - intrin->ppy) / intrin->fy;
    if (intrin->model == MODEL_INVERSE_BC) {
        float r2 = x * x + y * y;
        float f = 1 + intrin->coeffs[0] * r2 + intrin->coeffs[1] * r2 * r2 + intrin->coeffs[4] * r2 * r2 * r2;
        float ux = x * f + 2 * intrin->coeffs[2] * x * y + intrin->coeffs[3] * (r2 + 2 * x * x);
        float uy = y * f + 2 * intrin->coeffs[3] * x * y + intrin->coeffs[2] * (r2 + 2 * y * y);
        x = ux;
        y = uy;
    }
    point[0] = depth * x;
    point[1] = depth * y;
    point[2] = depth;
}

/* cuda-align.cu:85-110 */
static void deproject_pixel_to_point_double(double *point, const orbo_intrinsics *intrin, const float pixel[2], float depth) {
    double x = (pixel[0] - intrin->ppx) / intrin->fx;
    double y = (pixel[1] - intrin->ppy) / intrin->fy;
    if (intrin->model == MODEL_INVERSE_BC) {
        double r2 = x * x + y * y;
        double f = 1 + intrin->coeffs[0] * r2 + intrin->coeffs[1] * r2 * r2 + intrin->coeffs[4] * r2 * r2 * r2;
        double ux = x * f + 2 * intrin->coeffs[2] * x * y + intrin->coeffs[3] * (r2 + 2 * x * x);
        double uy = y * f + 2 * intrin->coeffs[3] * x * y + intrin->coeffs[2] * (r2 + 2 * y * y);
        x = ux;
        y = uy;
    }
    double depth_d = (double)depth;
    point[0] = depth_d * x;
    point[1] = depth_d * y;
    point[2] = depth_d;
}

/* cuda-align.cu:112-120 */
static void transform_point_to_point(float to_point[3], const orbo_extrinsics *extrin, const float from_point[3]) {
    to_point[0] = extrin->rotation[0] * from_point[0] + extrin->rotation[3] * from_point[1] + extrin->rotation[6] * from_point[2] + extrin->translation[0];
    to_point[1] = extrin->rotation[1] * from_point[0] + extrin->rotation[4] * from_point[1] + extrin->rotation[7] * from_point[2] + extrin->translation[1];
    to_point[2] = extrin->rotation[2] * from_point[0] + extrin->rotation[5] * from_point[1] + extrin->rotation[8] * from_point[2] + extrin->translation[2];
}

/* kernel_map_depth_to_other + kernel_reset_to_max + kernel_depth_to_other + kernel_reset_to_zero
 * (cuda-align.cu:122-286).  out: other.height x other.width u32. */
void orbo_align_depth_to_other(const uint16_t *depth, float depth_scale, const orbo_intrinsics *di,
                               const orbo_intrinsics *oi, const orbo_extrinsics *ex, uint32_t *out) {
    const size_t n_out = (size_t)oi->width * oi->height;
    for (size_t i = 0; i < n_out; ++i) out[i] = 9999999u;
    for (int depth_y = 0; depth_y < di->height; ++depth_y)
        for (int depth_x = 0; depth_x < di->width; ++depth_x) {
            const uint16_t raw = depth[(size_t)depth_y * di->width + depth_x];
            float depth_val = raw * depth_scale;
            int p[2][2] = {{-1, -1}, {-1, -1}};
            for (int block_index = 0; block_index < 2; ++block_index) {
                float shift = block_index ? 0.5 : -0.5;
                if (depth_val != 0) {
                    float depth_pixel[2] = {depth_x + shift, depth_y + shift}, depth_point[3], other_point[3], other_pixel[2];
                    deproject_pixel_to_point(depth_point, di, depth_pixel, depth_val);
                    transform_point_to_point(other_point, ex, depth_point);
                    project_point_to_pixel(other_pixel, oi, other_point);
                    p[block_index][0] = f2i_cuda(other_pixel[0] + 0.5f);
                    p[block_index][1] = f2i_cuda(other_pixel[1] + 0.5f);
                }
            }
            if (p[0][0] < 0 || p[0][1] < 0 || p[1][0] >= oi->width || p[1][1] >= oi->height) continue;
            for (int y = p[0][1]; y <= p[1][1]; ++y)
                for (int x = p[0][0]; x <= p[1][0]; ++x) {
                    uint32_t *o = &out[(size_t)y * oi->width + x];
                    if (raw < *o) *o = raw;
                }
        }
    for (size_t i = 0; i < n_out; ++i)
        if (out[i] == 9999999u) out[i] = 0;
}

/* kernel_keypoint_pixel_to_point (cuda-align.cu:282-364) with the (x, y) lookup and input-order compaction */
int orbo_keypoint_pixel_to_point(const uint32_t *aligned, const orbo_intrinsics *oi, const orbo_keypoint *kp,
                                 const uint8_t *desc, int n, orbo_keypoint *kp_out, uint8_t *desc_out, double *points) {
    int m = 0;
    for (int idx = 0; idx < n; ++idx) {
        const float pos[2] = {kp[idx].x, kp[idx].y};
        const int xi = (int)(pos[0] + 0.5), yi = (int)(pos[1] + 0.5);
        int depth = 0;
        if (xi >= 0 && yi >= 0 && xi < oi->width && yi < oi->height) depth = (int)aligned[(size_t)yi * oi->width + xi];
        const float score = kp[idx].response;
        if (depth > 1 && score > 1.0f) {
            deproject_pixel_to_point_double(points + 3 * (size_t)m, oi, pos, (float)depth);
            memcpy(desc_out + 32 * (size_t)m, desc + 32 * (size_t)idx, 32);
            kp_out[m] = kp[idx];
            ++m;
        }
    }
    return m;
}

/* kernel_reproject_prev_points (post_processing.cu:72-90) + project_point_to_pixel_double (:10-43).
 * T: column-major 4x4 (Eigen::Matrix4d) or NULL for identity; row.dot(v) is reduced pairwise like Eigen's
 * fixed-size redux: (a0 + a1) + (a2 + a3). */
void orbo_reproject_points(const double *points, int n, const double *T, const orbo_intrinsics *intrin, float *pos_out) {
    for (int idx = 0; idx < n; ++idx) {
        const double *p = points + 3 * (size_t)idx;
        double e[3] = {p[0], p[1], p[2]};
        if (T)
            for (int r = 0; r < 3; ++r) e[r] = (T[r] * p[0] + T[4 + r] * p[1]) + (T[8 + r] * p[2] + T[12 + r] * 1.0);
        float x = e[0] / e[2], y = e[1] / e[2];
        float pt[3] = {x, y, 1.0f}, pixel[2];
        project_point_to_pixel(pixel, intrin, pt); /* x / 1.0f == x exactly */
        pos_out[2 * idx] = pixel[0];
        pos_out[2 * idx + 1] = pixel[1];
    }
}

/* the is_matched tail of kernel_match_keypoints (post_processing.cu:176-197), in query order */
int orbo_compact_pairs(const int32_t *idx, int nq, const double *q_points, const double *t_points, const float *t_xy,
                       int t_xy_stride_floats, double *prev_out, double *curr_out, uint16_t *x_out, uint16_t *y_out) {
    int m = 0;
    for (int q = 0; q < nq; ++q) {
        const int t = idx[q];
        if (t < 0) continue;
        memcpy(prev_out + 3 * (size_t)m, q_points + 3 * (size_t)q, 3 * sizeof(double));
        memcpy(curr_out + 3 * (size_t)m, t_points + 3 * (size_t)t, 3 * sizeof(double));
        x_out[m] = (uint16_t)t_xy[(size_t)t * t_xy_stride_floats];
        y_out[m] = (uint16_t)t_xy[(size_t)t * t_xy_stride_floats + 1];
        ++m;
    }
    return m;
}

/* kernel_rgb_to_grayscale (reference src/cuda/cuda_RGB_to_Grayscale.cu:10-24), the expression as written there */
void orbo_rgb_to_grayscale(const uint8_t *src, size_t src_pitch, int cols, int rows, uint8_t *dst, size_t dst_pitch) {
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            float R, G, B;
            R = (float)(src[y * src_pitch + x * 3 + 0]);
            G = (float)(src[y * src_pitch + x * 3 + 1]);
            B = (float)(src[y * src_pitch + x * 3 + 2]);
            dst[y * dst_pitch + x] = (uint8_t)floor((B * 0.07 + G * 0.72 + R * 0.21) + 0.5);
        }
}
