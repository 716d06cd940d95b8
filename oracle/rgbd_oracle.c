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
#include <stdlib.h>
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

/* ---- camera model pieces.  The evaluation order of every sum and product below is the one C gives the
 * reference's expressions (left to right, integer constants promoted to the floating type of the expression);
 * compiled with -ffp-contract=off nothing is fused. */

/* radial polynomial 1 + k1 r2 + k2 r2^2 + k3 r2^3 (k3 is coeffs[4]) -- cuda-align.cu:39, :74 */
static float radial_f32(const float *k, float r2) { return 1 + k[0] * r2 + k[1] * r2 * r2 + k[4] * r2 * r2 * r2; }
static double radial_f64(const float *k, double r2) { return 1 + k[0] * r2 + k[1] * r2 * r2 + k[4] * r2 * r2 * r2; }

/* tangential terms added to a coordinate that has (forward model) or has not (inverse model) been scaled by the
 * radial factor; p1 = coeffs[2], p2 = coeffs[3] -- cuda-align.cu:42-43, :75-76 */
static float tangential_x32(const float *k, float base, float x, float y, float r2) {
    return base + 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x);
}
static float tangential_y32(const float *k, float base, float x, float y, float r2) {
    return base + 2 * k[3] * x * y + k[2] * (r2 + 2 * y * y);
}

/* normalised image coordinates -> pixel, with the forward (MODIFIED_BROWN_CONRADY) distortion -- cuda-align.cu:26-56 */
static void normalized_to_pixel(const orbo_intrinsics *cam, float x, float y, float *u, float *v) {
    if (cam->model == MODEL_MODIFIED_BC) {
        const float r2 = x * x + y * y;
        const float f = radial_f32(cam->coeffs, r2);
        x *= f;
        y *= f;
        const float xd = tangential_x32(cam->coeffs, x, x, y, r2), yd = tangential_y32(cam->coeffs, y, x, y, r2);
        x = xd;
        y = yd;
    }
    *u = x * cam->fx + cam->ppx;
    *v = y * cam->fy + cam->ppy;
}

/* pixel + depth -> 3-D point, with the INVERSE_BROWN_CONRADY undistortion -- cuda-align.cu:58-83 */
static void pixel_to_point_f32(const orbo_intrinsics *cam, float u, float v, float depth, float out[3]) {
    float x = (u - cam->ppx) / cam->fx;
    float y = (v - cam->ppy) / cam->fy;
    if (cam->model == MODEL_INVERSE_BC) {
        const float r2 = x * x + y * y;
        const float f = radial_f32(cam->coeffs, r2);
        const float xu = tangential_x32(cam->coeffs, x * f, x, y, r2), yu = tangential_y32(cam->coeffs, y * f, x, y, r2);
        x = xu;
        y = yu;
    }
    out[0] = depth * x;
    out[1] = depth * y;
    out[2] = depth;
}

/* the float64 twin used for the keypoints (float32 normalisation, float64 afterwards) -- cuda-align.cu:85-110 */
static void pixel_to_point_f64(const orbo_intrinsics *cam, float u, float v, float depth, double out[3]) {
    double x = (u - cam->ppx) / cam->fx;
    double y = (v - cam->ppy) / cam->fy;
    if (cam->model == MODEL_INVERSE_BC) {
        const float *k = cam->coeffs;
        const double r2 = x * x + y * y;
        const double f = radial_f64(k, r2);
        const double xu = x * f + 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x);
        const double yu = y * f + 2 * k[3] * x * y + k[2] * (r2 + 2 * y * y);
        x = xu;
        y = yu;
    }
    const double z = (double)depth;
    out[0] = z * x;
    out[1] = z * y;
    out[2] = z;
}

/* rigid transform, column-major rotation -- cuda-align.cu:112-120 */
static void rigid_transform(const orbo_extrinsics *e, const float p[3], float q[3]) {
    for (int r = 0; r < 3; ++r)
        q[r] = e->rotation[r] * p[0] + e->rotation[3 + r] * p[1] + e->rotation[6 + r] * p[2] + e->translation[r];
}

/* kernel_map_depth_to_other + kernel_reset_to_max + kernel_depth_to_other + kernel_reset_to_zero
 * (cuda-align.cu:122-286).  out: other.height x other.width u32. */
void orbo_align_depth_to_other(const uint16_t *depth, float depth_scale, const orbo_intrinsics *di,
                               const orbo_intrinsics *oi, const orbo_extrinsics *ex, uint32_t *out) {
    const size_t n_out = (size_t)oi->width * oi->height;
    const uint32_t UNSET = 9999999u; /* kernel_reset_to_max, cuda-align.cu:254-265 */
    for (size_t i = 0; i < n_out; ++i) out[i] = UNSET;
    for (int dy = 0; dy < di->height; ++dy)
        for (int dx = 0; dx < di->width; ++dx) {
            const uint16_t raw = depth[(size_t)dy * di->width + dx];
            const float metres = raw * depth_scale;
            if (metres == 0) continue; /* no depth: mapped pixel stays (-1,-1) and is rejected below (:139-141) */
            int corner[2][2];
            for (int c = 0; c < 2; ++c) { /* c = blockIdx.z of kernel_map_depth_to_other: top-left, bottom-right */
                const float shift = c ? 0.5 : -0.5;
                float p[3], q[3], u, v;
                pixel_to_point_f32(di, dx + shift, dy + shift, metres, p);
                rigid_transform(ex, p, q);
                normalized_to_pixel(oi, q[0] / q[2], q[1] / q[2], &u, &v);
                corner[c][0] = f2i_cuda(u + 0.5f);
                corner[c][1] = f2i_cuda(v + 0.5f);
            }
            if (corner[0][0] < 0 || corner[0][1] < 0 || corner[1][0] >= oi->width || corner[1][1] >= oi->height) continue;
            for (int y = corner[0][1]; y <= corner[1][1]; ++y)
                for (int x = corner[0][0]; x <= corner[1][0]; ++x) {
                    uint32_t *o = &out[(size_t)y * oi->width + x];
                    if (raw < *o) *o = raw; /* atomicMin, :246 */
                }
        }
    for (size_t i = 0; i < n_out; ++i) /* kernel_reset_to_zero, :267-279 */
        if (out[i] == UNSET) out[i] = 0;
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
            pixel_to_point_f64(oi, pos[0], pos[1], (float)depth, points + 3 * (size_t)m);
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
        const float x = e[0] / e[2], y = e[1] / e[2]; /* float64 division, then narrowed (post_processing.cu:16) */
        float pixel[2];
        normalized_to_pixel(intrin, x, y, &pixel[0], &pixel[1]);
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

/* ORB-SLAM2 ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th) gates + ComputeThreeMaxima (upstream
 * raulmur/ORB_SLAM2 src/ORBmatcher.cc, not vendored by the reference, no version pin): restated from the published
 * algorithm.  Front-end part only: no map-point bookkeeping, no stereo check, symmetric octave band.  Ties -> lowest
 * train index (upstream: grid traversal order of Frame::GetFeaturesInArea).  Returns the number of surviving matches. */
static int hamming256(const uint8_t *a, const uint8_t *b) {
    int d = 0;
    for (int i = 0; i < 32; ++i) d += __builtin_popcount((unsigned)(a[i] ^ b[i]));
    return d;
}

int orbo_search_by_projection(const uint8_t *q_desc, const float *q_uv, const orbo_keypoint *q_kp, int nq,
                              const uint8_t *t_desc, const orbo_keypoint *t_kp, int nt, const float *scale_factors,
                              int n_levels, float th, int th_high, int check_orientation, int32_t *out_idx,
                              int32_t *out_dist) {
    enum { HISTO_LENGTH = 30 };
    int hist[HISTO_LENGTH] = {0};
    const float factor = 1.0f / HISTO_LENGTH;
    for (int i = 0; i < nq; ++i) {
        out_idx[i] = -1;
        out_dist[i] = -1;
        int oct = q_kp[i].octave;
        if (oct < 0) oct = 0;
        if (oct > n_levels - 1) oct = n_levels - 1;
        const float radius = th * scale_factors[oct];
        const int min_level = oct - 1, max_level = oct + 1;
        int best_dist = 256, best_idx = -1;
        for (int j = 0; j < nt; ++j) {
            const float distx = t_kp[j].x - q_uv[2 * i], disty = t_kp[j].y - q_uv[2 * i + 1];
            if (!(fabsf(distx) < radius && fabsf(disty) < radius)) continue;
            if (t_kp[j].octave < min_level || t_kp[j].octave > max_level) continue;
            const int dist = hamming256(q_desc + 32 * (size_t)i, t_desc + 32 * (size_t)j);
            if (dist < best_dist) { best_dist = dist; best_idx = j; }
        }
        if (best_idx >= 0 && best_dist <= th_high) {
            out_idx[i] = best_idx;
            out_dist[i] = best_dist;
        }
    }
    int n = 0;
    int keep[HISTO_LENGTH];
    for (int b = 0; b < HISTO_LENGTH; ++b) keep[b] = 1;
    int *bins = 0;
    if (check_orientation) {
        bins = (int *)malloc(sizeof(int) * (size_t)(nq > 0 ? nq : 1));
        for (int i = 0; i < nq; ++i) {
            if (out_idx[i] < 0) continue;
            float rot = q_kp[i].angle - t_kp[out_idx[i]].angle;
            if (rot < 0.0) rot += 360.0f;
            int bin = (int)round(rot * factor);
            if (bin == HISTO_LENGTH) bin = 0;
            bins[i] = bin;
            hist[bin]++;
        }
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
        for (int b = 0; b < HISTO_LENGTH; ++b) {
            const int s = hist[b];
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = b; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = b; }
            else if (s > max3) { max3 = s; ind3 = b; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
        for (int b = 0; b < HISTO_LENGTH; ++b) keep[b] = (b == ind1 || b == ind2 || b == ind3);
    }
    for (int i = 0; i < nq; ++i) {
        if (out_idx[i] < 0) continue;
        if (check_orientation && !keep[bins[i]]) { out_idx[i] = -1; out_dist[i] = -1; }
        else ++n;
    }
    free(bins);
    return n;
}

/* ORB-SLAM2 Frame::ComputeStereoMatches (upstream raulmur/ORB_SLAM2 src/Frame.cc, not vendored by the reference, no
 * version pin): restated from the published algorithm.  left / right are oracle contexts whose pyramids hold the two
 * rectified images (orbo_extract or orbo_compute_pyramid was called on them); the keypoints / descriptors are passed
 * explicitly so that any ordering can be checked.  uright / depth: one float per left keypoint, -1 = no match.
 * Returns the number of matches left after the median-SAD filter. */
typedef struct { int dist, idx; } sad_idx;
static int cmp_sad_idx(const void *a, const void *b) {
    const sad_idx *x = (const sad_idx *)a, *y = (const sad_idx *)b;
    if (x->dist != y->dist) return x->dist < y->dist ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx);
}

int orbo_compute_stereo_matches(orbo_ctx *left, orbo_ctx *right, const orbo_keypoint *kl, const uint8_t *dl, int n_left,
                                const orbo_keypoint *kr, const uint8_t *dr, int n_right, float mbf, float fx, float *uright,
                                float *depth) {
    enum { TH_HIGH = 100, TH_LOW = 50 };
    const int nlev = orbo_nlevels(left);
    int32_t lw[ORBO_MAX_LEVELS], lh[ORBO_MAX_LEVELS], nf[ORBO_MAX_LEVELS];
    float sf[ORBO_MAX_LEVELS], isf[ORBO_MAX_LEVELS];
    orbo_get_geometry(left, lw, lh, sf, isf, nf);
    const int thOrbDist = (TH_HIGH + TH_LOW) / 2;
    const float mb = mbf / fx;
    const float minZ = mb, minD = 0, maxD = mbf / minZ;
    int *minr = (int *)malloc(sizeof(int) * (size_t)(n_right > 0 ? n_right : 1));
    int *maxr = (int *)malloc(sizeof(int) * (size_t)(n_right > 0 ? n_right : 1));
    for (int iR = 0; iR < n_right; ++iR) {
        const float kpY = kr[iR].y;
        const float r = 2.0f * sf[kr[iR].octave];
        maxr[iR] = (int)ceil(kpY + r);
        minr[iR] = (int)floor(kpY - r);
    }
    sad_idx *vDistIdx = (sad_idx *)malloc(sizeof(sad_idx) * (size_t)(n_left > 0 ? n_left : 1));
    int nmatch = 0;
    for (int iL = 0; iL < n_left; ++iL) {
        uright[iL] = -1.0f;
        depth[iL] = -1.0f;
        const int levelL = kl[iL].octave;
        const float vL = kl[iL].y, uL = kl[iL].x;
        const int row = (int)vL;
        const float minU = uL - maxD, maxU = uL - minD;
        if (maxU < 0) continue;
        int bestDist = TH_HIGH, bestIdxR = -1;
        for (int iR = 0; iR < n_right; ++iR) {
            if (row < minr[iR] || row > maxr[iR]) continue; /* vRowIndices[vL] membership */
            if (kr[iR].octave < levelL - 1 || kr[iR].octave > levelL + 1) continue;
            const float uR = kr[iR].x;
            if (uR >= minU && uR <= maxU) {
                const int dist = hamming256(dl + 32 * (size_t)iL, dr + 32 * (size_t)iR);
                if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
            }
        }
        if (bestIdxR < 0 || !(bestDist < thOrbDist)) continue;
        const float uR0 = kr[bestIdxR].x;
        const float scaleFactor = isf[levelL];
        const float scaleduL = round(kl[iL].x * scaleFactor);
        const float scaledvL = round(kl[iL].y * scaleFactor);
        const float scaleduR0 = round(uR0 * scaleFactor);
        const int w = 5, L = 5;
        int pw, ph;
        size_t pl, pr;
        const uint8_t *imL = orbo_level_padded(left, levelL, &pw, &ph, &pl) + (size_t)ORBO_EDGE_THRESHOLD * pl + ORBO_EDGE_THRESHOLD;
        const uint8_t *imR = orbo_level_padded(right, levelL, &pw, &ph, &pr) + (size_t)ORBO_EDGE_THRESHOLD * pr + ORBO_EDGE_THRESHOLD;
        const int cols = lw[levelL];
        int bestSad = INT32_MAX, bestincR = 0;
        float vDists[11];
        const float iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1;
        if (iniu < 0 || endu >= cols) continue;
        const int xl = (int)scaleduL, yl = (int)scaledvL, xr = (int)scaleduR0;
        const float centerL = imL[(ptrdiff_t)yl * (ptrdiff_t)pl + xl];
        for (int incR = -L; incR <= +L; ++incR) {
            const float centerR = imR[(ptrdiff_t)yl * (ptrdiff_t)pr + xr + incR];
            float dist = 0; /* cv::norm(IL, IR, NORM_L1) of the centre-subtracted float patches: exact integers */
            for (int dy = -w; dy <= w; ++dy)
                for (int dx = -w; dx <= w; ++dx) {
                    const float a = (float)imL[(ptrdiff_t)(yl + dy) * (ptrdiff_t)pl + xl + dx] - centerL;
                    const float b = (float)imR[(ptrdiff_t)(yl + dy) * (ptrdiff_t)pr + xr + incR + dx] - centerR;
                    dist += fabsf(a - b);
                }
            if (dist < bestSad) { bestSad = (int)dist; bestincR = incR; }
            vDists[L + incR] = dist;
        }
        if (bestincR == -L || bestincR == L) continue;
        const float dist1 = vDists[L + bestincR - 1], dist2 = vDists[L + bestincR], dist3 = vDists[L + bestincR + 1];
        const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
        if (deltaR < -1 || deltaR > 1) continue;
        float bestuR = sf[levelL] * ((float)scaleduR0 + (float)bestincR + deltaR);
        float disparity = (uL - bestuR);
        if (disparity >= minD && disparity < maxD) {
            if (disparity <= 0) {
                disparity = 0.01;
                bestuR = uL - 0.01;
            }
            depth[iL] = mbf / disparity;
            uright[iL] = bestuR;
            vDistIdx[nmatch].dist = bestSad;
            vDistIdx[nmatch].idx = iL;
            ++nmatch;
        }
    }
    int kept = nmatch;
    if (nmatch > 0) {
        qsort(vDistIdx, (size_t)nmatch, sizeof(sad_idx), cmp_sad_idx);
        const float median = vDistIdx[nmatch / 2].dist;
        const float thDist = 1.5f * 1.4f * median;
        for (int i = nmatch - 1; i >= 0; --i) {
            if (vDistIdx[i].dist < thDist) break;
            uright[vDistIdx[i].idx] = -1;
            depth[vDistIdx[i].idx] = -1;
            --kept;
        }
    }
    free(minr); free(maxr); free(vDistIdx);
    (void)nlev; (void)lh; (void)nf;
    return kept;
}
