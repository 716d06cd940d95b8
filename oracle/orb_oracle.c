/*
 * orb_oracle.c -- CPU ORACLE (test infrastructure, see orb_oracle.h header comment).
 *
 * Restates upstream ORB-SLAM2 ORBextractor (raulmur/ORB_SLAM2 src/ORBextractor.cc, not vendored
 * by the reference; SURVEY.md Appendix A is the spec) and the OpenCV 4.13 primitives it calls.
 * Reference anchors: the stage order of src/SlamGpuPipeline/buildStream.cpp:416-460,523-556;
 * the stage signatures in src/cuda/{pyramid,fast,nms,orb,post_processing}.cuh; the pattern
 * table src/cuda/orb.cuh:39-297; the extractor surface named by src_trash1/orb_extractor.cpp:6-8.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: no FMA contraction, so float32
 * arithmetic is reproducible on the GPU with __fmul_rn/__fadd_rn).
 */
#include "orb_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

static const int8_t k_pattern[1024] = {
#include "../include/orb_pattern_31.inc"
};

/* ---------- OpenCV scalar helpers ---------- */
static inline int cv_round_f(float v) { return (int)lrintf(v); }  /* round-half-even */
static inline int cv_round_d(double v) { return (int)lrint(v); }
static inline int cv_floor_f(float v) { return (int)floorf(v); }
static inline int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p;
        else p = 2 * len - 2 - p;
    }
    return p;
}

/* ---------- cv::resize(8U, INTER_LINEAR) -- SURVEY A.2 ----------
 * OpenCV imgproc/resize.cpp: scale = 1/((double)dst/src); fx=(float)((dx+.5)*scale-.5);
 * sx=floor(fx); fx-=sx; clamp (sx<0 -> 0,fx=0; sx>=sw-1 -> sw-1,fx=0); 11-bit weights via
 * saturate_cast<short>(w*2048) (round-half-even); rows clipped, row weights NOT zeroed;
 * dst = (((b0*(T0>>4))>>16) + ((b1*(T1>>4))>>16) + 2) >> 2.
 * Exact-2x shrink is rerouted by OpenCV to the INTER_AREA fast path: (a+b+c+d+2)>>2. */
void orbo_resize_linear_u8(const uint8_t *src, int sw, int sh, size_t sp, uint8_t *dst, int dw,
                           int dh, size_t dp) {
    if (sw == 2 * dw && sh == 2 * dh) {
        for (int y = 0; y < dh; ++y) {
            const uint8_t *r0 = src + (size_t)(2 * y) * sp, *r1 = r0 + sp;
            for (int x = 0; x < dw; ++x)
                dst[(size_t)y * dp + x] =
                    (uint8_t)((r0[2 * x] + r0[2 * x + 1] + r1[2 * x] + r1[2 * x + 1] + 2) >> 2);
        }
        return;
    }
    double inv_x = (double)dw / sw, inv_y = (double)dh / sh;
    double scale_x = 1. / inv_x, scale_y = 1. / inv_y;
    int *xofs = (int *)malloc(sizeof(int) * (size_t)dw);
    short *alpha = (short *)malloc(sizeof(short) * 2 * (size_t)dw);
    int *t0 = (int *)malloc(sizeof(int) * (size_t)dw), *t1 = (int *)malloc(sizeof(int) * (size_t)dw);
    for (int dx = 0; dx < dw; ++dx) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = cv_floor_f(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        alpha[2 * dx] = (short)cv_round_f((1.f - fx) * 2048.f);
        alpha[2 * dx + 1] = (short)cv_round_f(fx * 2048.f);
    }
    for (int dy = 0; dy < dh; ++dy) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = cv_floor_f(fy);
        fy -= sy;
        int b0 = (short)cv_round_f((1.f - fy) * 2048.f), b1 = (short)cv_round_f(fy * 2048.f);
        int y0 = sy < 0 ? 0 : (sy > sh - 1 ? sh - 1 : sy);
        int y1 = sy + 1 < 0 ? 0 : (sy + 1 > sh - 1 ? sh - 1 : sy + 1);
        const uint8_t *r0 = src + (size_t)y0 * sp, *r1 = src + (size_t)y1 * sp;
        for (int dx = 0; dx < dw; ++dx) {
            int sx = xofs[dx], sx1 = sx + 1 < sw ? sx + 1 : sw - 1;
            int a0 = alpha[2 * dx], a1 = alpha[2 * dx + 1];
            t0[dx] = r0[sx] * a0 + r0[sx1] * a1;
            t1[dx] = r1[sx] * a0 + r1[sx1] * a1;
        }
        for (int dx = 0; dx < dw; ++dx)
            dst[(size_t)dy * dp + dx] =
                (uint8_t)((((b0 * (t0[dx] >> 4)) >> 16) + ((b1 * (t1[dx] >> 4)) >> 16) + 2) >> 2);
    }
    free(xofs); free(alpha); free(t0); free(t1);
}

/* cv::copyMakeBorder(BORDER_REFLECT_101 [+ISOLATED]) in place around the ROI -- A.2 */
void orbo_border_reflect101(uint8_t *padded, int w, int h, size_t pitch, int b) {
    for (int py = 0; py < h + 2 * b; ++py) {
        int sy = reflect101(py - b, h) + b;
        for (int px = 0; px < w + 2 * b; ++px) {
            if (py >= b && py < h + b && px >= b && px < w + b) continue;
            int sx = reflect101(px - b, w) + b;
            padded[(size_t)py * pitch + px] = padded[(size_t)sy * pitch + sx];
        }
    }
}

/* ---------- cv::FAST TYPE_9_16 -- SURVEY A.3 ---------- */
static const int k_ring_dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int k_ring_dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

/* OpenCV cornerScore<16>: returns max(threshold, A, B) - 1 with
 * A = max over 9-arcs of min(v - ring), B = max over 9-arcs of min(ring - v). */
static int corner_score16(const uint8_t *p, size_t pitch, int threshold) {
    int d[25], v = p[0];
    for (int k = 0; k < 25; ++k)
        d[k] = v - p[(ptrdiff_t)k_ring_dy[k & 15] * (ptrdiff_t)pitch + k_ring_dx[k & 15]];
    int a0 = threshold;
    for (int k = 0; k < 16; k += 2) {
        int a = d[k + 1] < d[k + 2] ? d[k + 1] : d[k + 2];
        a = a < d[k + 3] ? a : d[k + 3];
        if (a <= a0) continue;
        for (int j = 4; j <= 8; ++j) a = a < d[k + j] ? a : d[k + j];
        int c = a < d[k] ? a : d[k];
        a0 = a0 > c ? a0 : c;
        c = a < d[k + 9] ? a : d[k + 9];
        a0 = a0 > c ? a0 : c;
    }
    int b0 = -a0;
    for (int k = 0; k < 16; k += 2) {
        int b = d[k + 1] > d[k + 2] ? d[k + 1] : d[k + 2];
        for (int j = 3; j <= 5; ++j) b = b > d[k + j] ? b : d[k + j];
        if (b >= b0) continue;
        for (int j = 6; j <= 8; ++j) b = b > d[k + j] ? b : d[k + j];
        int c = b > d[k] ? b : d[k];
        b0 = b0 < c ? b0 : c;
        c = b > d[k + 9] ? b : d[k + 9];
        b0 = b0 < c ? b0 : c;
    }
    return -b0 - 1;
}

static int is_corner9(const uint8_t *p, size_t pitch, int t) {
    int v = p[0], lo = v - t, hi = v + t;
    unsigned dark = 0, bright = 0;
    for (int k = 0; k < 16; ++k) {
        int r = p[(ptrdiff_t)k_ring_dy[k] * (ptrdiff_t)pitch + k_ring_dx[k]];
        if (r < lo) dark |= 1u << k;
        if (r > hi) bright |= 1u << k;
    }
    for (int pass = 0; pass < 2; ++pass) {
        unsigned m = pass ? bright : dark;
        m |= m << 16;
        int run = 0;
        for (int k = 0; k < 25; ++k) {
            if (m & (1u << k)) { if (++run >= 9) return 1; }
            else run = 0;
        }
    }
    return 0;
}

int orbo_fast9_window(const uint8_t *win, int cw, int ch, size_t pitch, int thr, int nms,
                      orbo_candidate *out, int max_out) {
    if (cw < 7 || ch < 7) return 0;
    thr = thr < 0 ? 0 : (thr > 255 ? 255 : thr);
    /* score rows with a zero frame; 0 = "not a corner at this threshold" (a corner scores >= thr, and a
     * corner with thr == 0 and score 0 cannot beat anything in the strict NMS either) */
    const int sp = cw + 2;
    uint8_t stack_buf[80 * 80];
    uint8_t *sc = (size_t)sp * (ch + 2) <= sizeof(stack_buf) ? stack_buf : (uint8_t *)malloc((size_t)sp * (ch + 2));
    uint8_t *is = NULL;
    memset(sc, 0, (size_t)sp * (ch + 2));
    if (thr == 0 || !nms) is = (uint8_t *)calloc((size_t)cw * ch, 1);
    for (int y = 3; y < ch - 3; ++y) {
        const uint8_t *row = win + (size_t)y * pitch;
        for (int x = 3; x < cw - 3; ++x) {
            const uint8_t *p = row + x;
            const int v = p[0], lo = v - thr, hi = v + thr;
            /* any 9-arc contains one pixel of every antipodal pair (same prune as cv::FAST) */
            const int r0 = p[3 * (ptrdiff_t)pitch], r8 = p[-3 * (ptrdiff_t)pitch], r4 = p[3], r12 = p[-3];
            const int dark = ((r0 < lo) | (r8 < lo)) & ((r4 < lo) | (r12 < lo));
            const int bright = ((r0 > hi) | (r8 > hi)) & ((r4 > hi) | (r12 > hi));
            if (!(dark | bright)) continue;
            if (!is_corner9(p, pitch, thr)) continue;
            sc[(y + 1) * sp + x + 1] = (uint8_t)corner_score16(p, pitch, thr);
            if (is) is[y * cw + x] = 1;
        }
    }
    int n = 0;
    for (int y = 3; y < ch - 3; ++y)
        for (int x = 3; x < cw - 3; ++x) {
            const uint8_t *q = sc + (y + 1) * sp + x + 1;
            const int s = q[0];
            if (is ? !is[y * cw + x] : s == 0) continue;
            if (nms && !(s > q[-1] && s > q[1] && s > q[-sp - 1] && s > q[-sp] && s > q[-sp + 1] &&
                         s > q[sp - 1] && s > q[sp] && s > q[sp + 1]))
                continue;
            if (n < max_out) { out[n].x = x; out[n].y = y; out[n].response = s; }
            ++n;
        }
    if (sc != stack_buf) free(sc);
    free(is);
    return n;
}

void orbo_fast_score_map(const uint8_t *img, int w, int h, size_t pitch, uint8_t *score,
                         size_t score_pitch) {
    for (int y = 0; y < h; ++y) memset(score + (size_t)y * score_pitch, 0, (size_t)w);
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            int m = corner_score16(img + (size_t)y * pitch + x, pitch, 0) + 1;
            score[(size_t)y * score_pitch + x] = (uint8_t)(m < 0 ? 0 : (m > 255 ? 255 : m));
        }
}

/* ---------- cv::GaussianBlur(7x7, 2, 2, REFLECT_101) on 8U, OpenCV 4.13 -- A.6 ----------
 * fixed-point separable kernel [18,34,48,56,48,34,18]/256, dst = (V + 32768) >> 16 */
void orbo_gaussian_blur7(const uint8_t *src, int w, int h, size_t sp, uint8_t *dst, size_t dp) {
    static const uint32_t k[7] = {18, 34, 48, 56, 48, 34, 18};
    uint16_t *hbuf = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)w * h);
    int *xi = (int *)malloc(sizeof(int) * (size_t)(w + 6));
    for (int x = -3; x < w + 3; ++x) xi[x + 3] = reflect101(x, w);
    for (int y = 0; y < h; ++y) {
        const uint8_t *r = src + (size_t)y * sp;
        uint16_t *o = hbuf + (size_t)y * w;
        for (int x = 0; x < w; ++x) {
            if (x >= 3 && x < w - 3) {
                const uint8_t *p = r + x - 3;
                o[x] = (uint16_t)(18 * (p[0] + p[6]) + 34 * (p[1] + p[5]) + 48 * (p[2] + p[4]) + 56 * p[3]);
            } else {
                uint32_t s = 0;
                for (int i = 0; i < 7; ++i) s += k[i] * r[xi[x + i]];
                o[x] = (uint16_t)s;
            }
        }
    }
    for (int y = 0; y < h; ++y) {
        const uint16_t *rr[7];
        for (int i = 0; i < 7; ++i) rr[i] = hbuf + (size_t)reflect101(y + i - 3, h) * w;
        uint8_t *o = dst + (size_t)y * dp;
        for (int x = 0; x < w; ++x) {
            const uint32_t s = 18u * ((uint32_t)rr[0][x] + rr[6][x]) + 34u * ((uint32_t)rr[1][x] + rr[5][x]) +
                               48u * ((uint32_t)rr[2][x] + rr[4][x]) + 56u * rr[3][x];
            o[x] = (uint8_t)((s + 32768u) >> 16);
        }
    }
    free(hbuf); free(xi);
}

/* cv::fastAtan2 (degrees), float32, no FMA -- A.5 */
float orbo_fast_atan2(float y, float x) {
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale,
                p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    float ax = fabsf(x), ay = fabsf(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

/* ---------- matcher -- A.7 ---------- */
static inline int hamming256(const uint8_t *a, const uint8_t *b) {
    uint64_t x[4], y[4];
    memcpy(x, a, 32); memcpy(y, b, 32);
    return __builtin_popcountll(x[0] ^ y[0]) + __builtin_popcountll(x[1] ^ y[1]) +
           __builtin_popcountll(x[2] ^ y[2]) + __builtin_popcountll(x[3] ^ y[3]);
}

int orbo_match_knn(const uint8_t *q, int nq, const uint8_t *t, int nt, int k, float ratio,
                   int32_t *out_idx, int32_t *out_dist, uint8_t *accept) {
    int nacc = 0;
    for (int i = 0; i < nq; ++i) {
        int b1 = -1, b2 = -1, d1 = 257, d2 = 257;
        for (int j = 0; j < nt; ++j) {
            int d = hamming256(q + (size_t)i * 32, t + (size_t)j * 32);
            if (d < d1) { d2 = d1; b2 = b1; d1 = d; b1 = j; }
            else if (d < d2) { d2 = d; b2 = j; }
        }
        out_idx[2 * i] = b1; out_idx[2 * i + 1] = b2;
        out_dist[2 * i] = b1 >= 0 ? d1 : -1; out_dist[2 * i + 1] = b2 >= 0 ? d2 : -1;
        int ok = b1 >= 0;
        if (k == 2) ok = (b2 >= 0) && ((float)d1 < ratio * (float)d2);
        if (accept) accept[i] = (uint8_t)ok;
        nacc += ok;
    }
    return nacc;
}

/* reference kernel_match_keypoints semantics (src/cuda/post_processing.cu:156-171) on 256-bit descriptors */
int orbo_match_windowed(const uint8_t *q, const float *q_xy, int nq, const uint8_t *t, const float *t_xy, int nt,
                        float max_px, int max_hamming, int32_t *out_idx, int32_t *out_dist) {
    int nm = 0;
    for (int i = 0; i < nq; ++i) {
        int best = -1, bd = max_hamming;
        for (int j = 0; j < nt; ++j) {
            if (fabsf(q_xy[2 * i] - t_xy[2 * j]) <= max_px && fabsf(q_xy[2 * i + 1] - t_xy[2 * j + 1]) <= max_px) {
                const int d = hamming256(q + (size_t)i * 32, t + (size_t)j * 32);
                if (d < bd) { bd = d; best = j; }
            }
        }
        out_idx[i] = best;
        out_dist[i] = best >= 0 ? bd : -1;
        nm += best >= 0;
    }
    return nm;
}

/* ---------- extractor context ---------- */
struct orbo_ctx {
    orbo_params p;
    int w, h, nlevels;
    int lw[ORBO_MAX_LEVELS], lh[ORBO_MAX_LEVELS], nfeat[ORBO_MAX_LEVELS];
    float sf[ORBO_MAX_LEVELS], inv_sf[ORBO_MAX_LEVELS];
    uint8_t *lvl[ORBO_MAX_LEVELS]; /* padded */
    size_t pitch[ORBO_MAX_LEVELS];
    int umax[16];
};

int orbo_create(orbo_ctx **out, const orbo_params *p, int width, int height) {
    if (!out || !p || p->nlevels < 1 || p->nlevels > ORBO_MAX_LEVELS || p->scale_factor <= 1.0f)
        return -1;
    orbo_ctx *c = (orbo_ctx *)calloc(1, sizeof(*c));
    c->p = *p; c->w = width; c->h = height; c->nlevels = p->nlevels;
    c->sf[0] = 1.0f;
    for (int i = 1; i < c->nlevels; ++i) c->sf[i] = c->sf[i - 1] * p->scale_factor;
    for (int i = 0; i < c->nlevels; ++i) c->inv_sf[i] = 1.0f / c->sf[i];
    float factor = 1.0f / p->scale_factor;
    float nd = p->nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)c->nlevels));
    int sum = 0;
    for (int l = 0; l < c->nlevels - 1; ++l) {
        c->nfeat[l] = cv_round_f(nd);
        sum += c->nfeat[l];
        nd *= factor;
    }
    c->nfeat[c->nlevels - 1] = p->nfeatures - sum > 0 ? p->nfeatures - sum : 0;
    for (int l = 0; l < c->nlevels; ++l) {
        c->lw[l] = cv_round_f((float)width * c->inv_sf[l]);
        c->lh[l] = cv_round_f((float)height * c->inv_sf[l]);
        /* upstream divides by nCols = (w-32)/30; reject shapes where that is zero */
        if (c->lw[l] - 32 < 30 || c->lh[l] - 32 < 30) { free(c); return -2; }
        c->pitch[l] = (size_t)c->lw[l] + 2 * ORBO_EDGE_THRESHOLD;
        c->lvl[l] = (uint8_t *)calloc(c->pitch[l] * (size_t)(c->lh[l] + 2 * ORBO_EDGE_THRESHOLD), 1);
    }
    /* umax -- A.1 */
    int vmax = (int)floor(ORBO_HALF_PATCH * sqrt(2.f) / 2 + 1);
    int vmin = (int)ceil(ORBO_HALF_PATCH * sqrt(2.f) / 2);
    const double hp2 = ORBO_HALF_PATCH * ORBO_HALF_PATCH;
    for (int v = 0; v <= vmax; ++v) c->umax[v] = cv_round_d(sqrt(hp2 - v * v));
    for (int v = ORBO_HALF_PATCH, v0 = 0; v >= vmin; --v) {
        while (c->umax[v0] == c->umax[v0 + 1]) ++v0;
        c->umax[v] = v0;
        ++v0;
    }
    *out = c;
    return 0;
}

void orbo_destroy(orbo_ctx *c) {
    if (!c) return;
    for (int l = 0; l < c->nlevels; ++l) free(c->lvl[l]);
    free(c);
}

int orbo_nlevels(const orbo_ctx *c) { return c->nlevels; }

void orbo_get_geometry(const orbo_ctx *c, int32_t *lw, int32_t *lh, float *scale,
                       float *inv_scale, int32_t *nfeat) {
    for (int l = 0; l < c->nlevels; ++l) {
        if (lw) lw[l] = c->lw[l];
        if (lh) lh[l] = c->lh[l];
        if (scale) scale[l] = c->sf[l];
        if (inv_scale) inv_scale[l] = c->inv_sf[l];
        if (nfeat) nfeat[l] = c->nfeat[l];
    }
}

void orbo_get_umax(const orbo_ctx *c, int32_t *u) { for (int i = 0; i < 16; ++i) u[i] = c->umax[i]; }

static inline uint8_t *roi(const orbo_ctx *c, int l) {
    return c->lvl[l] + (size_t)ORBO_EDGE_THRESHOLD * c->pitch[l] + ORBO_EDGE_THRESHOLD;
}

/* ORBextractor::ComputePyramid -- A.2 */
int orbo_compute_pyramid(orbo_ctx *c, const uint8_t *img, size_t pitch) {
    for (int y = 0; y < c->h; ++y) memcpy(roi(c, 0) + (size_t)y * c->pitch[0], img + (size_t)y * pitch, (size_t)c->w);
    orbo_border_reflect101(c->lvl[0], c->lw[0], c->lh[0], c->pitch[0], ORBO_EDGE_THRESHOLD);
    for (int l = 1; l < c->nlevels; ++l) {
        orbo_resize_linear_u8(roi(c, l - 1), c->lw[l - 1], c->lh[l - 1], c->pitch[l - 1], roi(c, l),
                              c->lw[l], c->lh[l], c->pitch[l]);
        orbo_border_reflect101(c->lvl[l], c->lw[l], c->lh[l], c->pitch[l], ORBO_EDGE_THRESHOLD);
    }
    return 0;
}

const uint8_t *orbo_level_padded(const orbo_ctx *c, int l, int *pw, int *ph, size_t *pitch) {
    if (pw) *pw = c->lw[l] + 2 * ORBO_EDGE_THRESHOLD;
    if (ph) *ph = c->lh[l] + 2 * ORBO_EDGE_THRESHOLD;
    if (pitch) *pitch = c->pitch[l];
    return c->lvl[l];
}

/* ComputeKeyPointsOctTree, candidate part -- A.3 (literal window loop) */
int orbo_level_candidates(orbo_ctx *c, int l, orbo_candidate *out, int max_out) {
    const int cols = c->lw[l], rows = c->lh[l];
    const int minBX = ORBO_EDGE_THRESHOLD - 3, minBY = minBX;
    const int maxBX = cols - ORBO_EDGE_THRESHOLD + 3, maxBY = rows - ORBO_EDGE_THRESHOLD + 3;
    const float W = 30;
    const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
    const int nCols = (int)(width / W), nRows = (int)(height / W);
    const int wCell = (int)ceilf(width / nCols), hCell = (int)ceilf(height / nRows);
    const uint8_t *im = roi(c, l);
    const size_t pitch = c->pitch[l];
    int n = 0;
    orbo_candidate *tmp = (orbo_candidate *)malloc(sizeof(orbo_candidate) * 64 * 64);
    for (int i = 0; i < nRows; ++i) {
        const float iniY = (float)(minBY + i * hCell);
        float maxY = iniY + hCell + 6;
        if (iniY >= maxBY - 3) continue;
        if (maxY > maxBY) maxY = (float)maxBY;
        for (int j = 0; j < nCols; ++j) {
            const float iniX = (float)(minBX + j * wCell);
            float maxX = iniX + wCell + 6;
            if (iniX >= maxBX - 6) continue;
            if (maxX > maxBX) maxX = (float)maxBX;
            const int x0 = (int)iniX, y0 = (int)iniY, cw = (int)maxX - x0, ch = (int)maxY - y0;
            const uint8_t *win = im + (size_t)y0 * pitch + x0;
            int k = orbo_fast9_window(win, cw, ch, pitch, c->p.ini_th_fast, 1, tmp, 64 * 64);
            if (k == 0) k = orbo_fast9_window(win, cw, ch, pitch, c->p.min_th_fast, 1, tmp, 64 * 64);
            for (int q = 0; q < k; ++q) {
                if (n < max_out) {
                    out[n].x = tmp[q].x + j * wCell;
                    out[n].y = tmp[q].y + i * hCell;
                    out[n].response = tmp[q].response;
                }
                ++n;
            }
        }
    }
    free(tmp);
    return n;
}

/* ---------- ORBextractor::DistributeOctTree / ExtractorNode::DivideNode -- A.4 ---------- */
typedef struct Node {
    int ulx, uly, urx, bry; /* UL=(ulx,uly) UR=(urx,uly) BL=(ulx,bry) BR=(urx,bry) */
    int *keys, nkeys, no_more;
    long seq; /* creation sequence: the oracle's definition of upstream's pointer tie-break */
    struct Node *prev, *next;
} Node;

typedef struct { Node *head, *tail; int size; long next_seq; } NodeList;

static Node *node_new(NodeList *L, int cap) {
    Node *n = (Node *)calloc(1, sizeof(Node));
    n->keys = (int *)malloc(sizeof(int) * (size_t)(cap > 0 ? cap : 1));
    n->seq = L->next_seq++;
    return n;
}
static void list_push_front(NodeList *L, Node *n) {
    n->prev = NULL; n->next = L->head;
    if (L->head) L->head->prev = n; else L->tail = n;
    L->head = n; L->size++;
}
static void list_push_back(NodeList *L, Node *n) {
    n->next = NULL; n->prev = L->tail;
    if (L->tail) L->tail->next = n; else L->head = n;
    L->tail = n; L->size++;
}
static Node *list_erase(NodeList *L, Node *n) { /* returns next */
    Node *nx = n->next;
    if (n->prev) n->prev->next = n->next; else L->head = n->next;
    if (n->next) n->next->prev = n->prev; else L->tail = n->prev;
    L->size--;
    free(n->keys); free(n);
    return nx;
}

static void divide_node(NodeList *L, const Node *p, const orbo_candidate *cand, Node *ch[4]) {
    const int halfX = (int)ceilf((float)(p->urx - p->ulx) / 2);
    const int halfY = (int)ceilf((float)(p->bry - p->uly) / 2);
    for (int k = 0; k < 4; ++k) ch[k] = node_new(L, p->nkeys);
    ch[0]->ulx = p->ulx;         ch[0]->uly = p->uly;         ch[0]->urx = p->ulx + halfX; ch[0]->bry = p->uly + halfY;
    ch[1]->ulx = p->ulx + halfX; ch[1]->uly = p->uly;         ch[1]->urx = p->urx;         ch[1]->bry = p->uly + halfY;
    ch[2]->ulx = p->ulx;         ch[2]->uly = p->uly + halfY; ch[2]->urx = p->ulx + halfX; ch[2]->bry = p->bry;
    ch[3]->ulx = p->ulx + halfX; ch[3]->uly = p->uly + halfY; ch[3]->urx = p->urx;         ch[3]->bry = p->bry;
    const int midx = ch[0]->urx, midy = ch[0]->bry;
    for (int i = 0; i < p->nkeys; ++i) {
        const orbo_candidate *kp = &cand[p->keys[i]];
        int k = (kp->x < midx) ? ((kp->y < midy) ? 0 : 2) : ((kp->y < midy) ? 1 : 3);
        ch[k]->keys[ch[k]->nkeys++] = p->keys[i];
    }
    for (int k = 0; k < 4; ++k) if (ch[k]->nkeys == 1) ch[k]->no_more = 1;
}

typedef struct { int count; Node *node; } SizeNode;
static int cmp_sizenode(const void *a, const void *b) {
    const SizeNode *x = (const SizeNode *)a, *y = (const SizeNode *)b;
    if (x->count != y->count) return x->count < y->count ? -1 : 1;
    /* upstream compares the node POINTERS here (allocator dependent); oracle: creation seq */
    return x->node->seq < y->node->seq ? -1 : (x->node->seq > y->node->seq ? 1 : 0);
}

int orbo_distribute_octree(const orbo_candidate *cand, int n, int minX, int maxX, int minY,
                           int maxY, int N, int32_t *out_index, int max_out) {
    const int nIni = (int)roundf((float)(maxX - minX) / (maxY - minY));
    if (nIni < 1) return -3;
    const float hX = (float)(maxX - minX) / nIni;
    NodeList L = {0};
    Node **ini = (Node **)malloc(sizeof(Node *) * (size_t)nIni);
    for (int i = 0; i < nIni; ++i) {
        Node *ni = node_new(&L, n);
        ni->ulx = (int)(hX * (float)i); ni->urx = (int)(hX * (float)(i + 1));
        ni->uly = 0; ni->bry = maxY - minY;
        list_push_back(&L, ni);
        ini[i] = ni;
    }
    for (int i = 0; i < n; ++i) {
        Node *r = ini[(int)((float)cand[i].x / hX)];
        r->keys[r->nkeys++] = i;
    }
    free(ini);
    for (Node *it = L.head; it;) {
        if (it->nkeys == 1) { it->no_more = 1; it = it->next; }
        else if (it->nkeys == 0) it = list_erase(&L, it);
        else it = it->next;
    }
    int finish = 0, cap = 16, nsn = 0;
    SizeNode *sn = (SizeNode *)malloc(sizeof(SizeNode) * (size_t)cap);
#define SN_PUSH(cnt, nd) do { if (nsn == cap) { cap *= 2; sn = (SizeNode *)realloc(sn, sizeof(SizeNode) * (size_t)cap); } \
                              sn[nsn].count = (cnt); sn[nsn].node = (nd); ++nsn; } while (0)
    while (!finish) {
        int prevSize = L.size, nToExpand = 0;
        nsn = 0;
        for (Node *it = L.head; it;) {
            if (it->no_more) { it = it->next; continue; }
            Node *ch[4];
            divide_node(&L, it, cand, ch);
            for (int k = 0; k < 4; ++k) {
                if (ch[k]->nkeys > 0) {
                    list_push_front(&L, ch[k]);
                    if (ch[k]->nkeys > 1) { ++nToExpand; SN_PUSH(ch[k]->nkeys, ch[k]); }
                } else { free(ch[k]->keys); free(ch[k]); }
            }
            it = list_erase(&L, it);
        }
        if (L.size >= N || L.size == prevSize) {
            finish = 1;
        } else if (L.size + nToExpand * 3 > N) {
            while (!finish) {
                prevSize = L.size;
                int nprev = nsn;
                SizeNode *prev = (SizeNode *)malloc(sizeof(SizeNode) * (size_t)(nprev > 0 ? nprev : 1));
                memcpy(prev, sn, sizeof(SizeNode) * (size_t)nprev);
                nsn = 0;
                qsort(prev, (size_t)nprev, sizeof(SizeNode), cmp_sizenode);
                for (int j = nprev - 1; j >= 0; --j) {
                    Node *ch[4];
                    divide_node(&L, prev[j].node, cand, ch);
                    for (int k = 0; k < 4; ++k) {
                        if (ch[k]->nkeys > 0) {
                            list_push_front(&L, ch[k]);
                            if (ch[k]->nkeys > 1) SN_PUSH(ch[k]->nkeys, ch[k]);
                        } else { free(ch[k]->keys); free(ch[k]); }
                    }
                    list_erase(&L, prev[j].node);
                    if (L.size >= N) break;
                }
                free(prev);
                if (L.size >= N || L.size == prevSize) finish = 1;
            }
        }
    }
#undef SN_PUSH
    free(sn);
    int cnt = 0;
    for (Node *it = L.head; it; it = it->next) {
        int best = it->keys[0], maxr = cand[best].response;
        for (int k = 1; k < it->nkeys; ++k)
            if (cand[it->keys[k]].response > maxr) { best = it->keys[k]; maxr = cand[best].response; }
        if (cnt < max_out) out_index[cnt] = best;
        ++cnt;
    }
    while (L.head) list_erase(&L, L.head);
    return cnt;
}

/* IC_Angle -- A.5 */
float orbo_ic_angle(const orbo_ctx *c, int l, float x, float y) {
    const size_t step = c->pitch[l];
    const uint8_t *center = roi(c, l) + (size_t)cv_round_f(y) * step + cv_round_f(x);
    int m01 = 0, m10 = 0;
    for (int u = -ORBO_HALF_PATCH; u <= ORBO_HALF_PATCH; ++u) m10 += u * center[u];
    for (int v = 1; v <= ORBO_HALF_PATCH; ++v) {
        int v_sum = 0, d = c->umax[v];
        for (int u = -d; u <= d; ++u) {
            int vp = center[u + (ptrdiff_t)v * (ptrdiff_t)step], vm = center[u - (ptrdiff_t)v * (ptrdiff_t)step];
            v_sum += vp - vm;
            m10 += u * (vp + vm);
        }
        m01 += v * v_sum;
    }
    return orbo_fast_atan2((float)m01, (float)m10);
}

int orbo_level_blurred(orbo_ctx *c, int l, uint8_t *out) {
    orbo_gaussian_blur7(roi(c, l), c->lw[l], c->lh[l], c->pitch[l], out, (size_t)c->lw[l]);
    return 0;
}

/* computeOrbDescriptor -- A.6; oracle definition a=(float)cos((double)rad), b=(float)sin(...) */
void orbo_descriptor(const uint8_t *blurred, size_t pitch, float x, float y, float angle_deg,
                     uint8_t *desc) {
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    const float angle = angle_deg * factorPI;
    const float a = (float)cos((double)angle), b = (float)sin((double)angle);
    const uint8_t *center = blurred + (size_t)cv_round_f(y) * pitch + cv_round_f(x);
    const int8_t *pat = k_pattern;
    for (int i = 0; i < 32; ++i, pat += 32) {
        int val = 0;
        for (int k = 0; k < 8; ++k) {
            const int8_t *p0 = pat + 4 * k, *p1 = p0 + 2;
            int t0 = center[(ptrdiff_t)cv_round_f(p0[0] * b + p0[1] * a) * (ptrdiff_t)pitch + cv_round_f(p0[0] * a - p0[1] * b)];
            int t1 = center[(ptrdiff_t)cv_round_f(p1[0] * b + p1[1] * a) * (ptrdiff_t)pitch + cv_round_f(p1[0] * a - p1[1] * b)];
            val |= (t0 < t1) << k;
        }
        desc[i] = (uint8_t)val;
    }
}

/* ORBextractor::operator() */
int orbo_extract(orbo_ctx *c, const uint8_t *img, size_t pitch, orbo_keypoint *out_kp,
                 uint8_t *out_desc, int max_kp) {
    orbo_compute_pyramid(c, img, pitch);
    int total = 0;
    for (int l = 0; l < c->nlevels; ++l) {
        const int cols = c->lw[l], rows = c->lh[l];
        const int minBX = ORBO_EDGE_THRESHOLD - 3, minBY = minBX;
        const int maxBX = cols - ORBO_EDGE_THRESHOLD + 3, maxBY = rows - ORBO_EDGE_THRESHOLD + 3;
        int cap = cols * rows / 4 + 64;
        orbo_candidate *cand = (orbo_candidate *)malloc(sizeof(orbo_candidate) * (size_t)cap);
        int n = orbo_level_candidates(c, l, cand, cap);
        if (n > cap) { free(cand); return -4; }
        int32_t *sel = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n + 1));
        int ns = n ? orbo_distribute_octree(cand, n, minBX, maxBX, minBY, maxBY, c->nfeat[l], sel, n) : 0;
        if (ns < 0) { free(cand); free(sel); return ns; }
        const int scaledPatchSize = (int)(ORBO_PATCH * c->sf[l]);
        uint8_t *blur = NULL;
        if (ns > 0) {
            blur = (uint8_t *)malloc((size_t)cols * rows);
            orbo_level_blurred(c, l, blur);
        }
        for (int i = 0; i < ns; ++i) {
            if (total >= max_kp) { free(cand); free(sel); free(blur); return -5; }
            orbo_keypoint kp;
            kp.x = (float)(cand[sel[i]].x + minBX);
            kp.y = (float)(cand[sel[i]].y + minBY);
            kp.size = (float)scaledPatchSize;
            kp.response = (float)cand[sel[i]].response;
            kp.octave = l;
            kp.class_id = -1;
            kp.angle = orbo_ic_angle(c, l, kp.x, kp.y);
            if (out_desc) orbo_descriptor(blur, (size_t)cols, kp.x, kp.y, kp.angle, out_desc + (size_t)total * 32);
            if (l != 0) { kp.x *= c->sf[l]; kp.y *= c->sf[l]; }
            out_kp[total++] = kp;
        }
        free(cand); free(sel); free(blur);
    }
    return total;
}

/* ---------- multi-threaded drivers for the CPU baseline ---------- */
typedef struct {
    const orbo_params *p; const uint8_t *frames; int w, h; size_t pitch, stride;
    int n_frames, tid, nthreads; int32_t *counts; long total; int err;
} ExtractJob;

static void *extract_worker(void *arg) {
    ExtractJob *j = (ExtractJob *)arg;
    orbo_ctx *c = NULL;
    if (orbo_create(&c, j->p, j->w, j->h)) { j->err = -1; return NULL; }
    int cap = j->p->nfeatures + 16 * ORBO_MAX_LEVELS + 64;
    orbo_keypoint *kp = (orbo_keypoint *)malloc(sizeof(orbo_keypoint) * (size_t)cap);
    uint8_t *desc = (uint8_t *)malloc((size_t)cap * 32);
    for (int f = j->tid; f < j->n_frames; f += j->nthreads) {
        int n = orbo_extract(c, j->frames + (size_t)f * j->stride, j->pitch, kp, desc, cap);
        if (n < 0) { j->err = n; break; }
        if (j->counts) j->counts[f] = n;
        j->total += n;
    }
    free(kp); free(desc); orbo_destroy(c);
    return NULL;
}

long orbo_extract_many(const orbo_params *p, const uint8_t *frames, int w, int h, size_t pitch,
                       size_t frame_stride, int n_frames, int n_threads, int32_t *counts) {
    if (n_threads < 1) n_threads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    ExtractJob *jobs = (ExtractJob *)calloc((size_t)n_threads, sizeof(ExtractJob));
    for (int t = 0; t < n_threads; ++t) {
        jobs[t] = (ExtractJob){p, frames, w, h, pitch, frame_stride, n_frames, t, n_threads, counts, 0, 0};
        pthread_create(&th[t], NULL, extract_worker, &jobs[t]);
    }
    long total = 0; int err = 0;
    for (int t = 0; t < n_threads; ++t) {
        pthread_join(th[t], NULL);
        total += jobs[t].total;
        if (jobs[t].err) err = jobs[t].err;
    }
    free(th); free(jobs);
    return err ? err : total;
}

typedef struct {
    const uint8_t *q, *t; int nq, nt, k; float ratio; int q0, q1;
    int32_t *idx, *dist; uint8_t *acc; long nacc;
} MatchJob;

static void *match_worker(void *arg) {
    MatchJob *j = (MatchJob *)arg;
    if (j->q1 > j->q0)
        j->nacc = orbo_match_knn(j->q + (size_t)j->q0 * 32, j->q1 - j->q0, j->t, j->nt, j->k, j->ratio,
                                 j->idx + 2 * (size_t)j->q0, j->dist + 2 * (size_t)j->q0,
                                 j->acc ? j->acc + j->q0 : NULL);
    return NULL;
}

long orbo_match_many(const uint8_t *q, int nq, const uint8_t *t, int nt, int k, float ratio,
                     int n_threads, int32_t *out_idx, int32_t *out_dist, uint8_t *accept) {
    if (n_threads < 1) n_threads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    MatchJob *jobs = (MatchJob *)calloc((size_t)n_threads, sizeof(MatchJob));
    int per = (nq + n_threads - 1) / n_threads;
    for (int i = 0; i < n_threads; ++i) {
        int q0 = i * per, q1 = q0 + per > nq ? nq : q0 + per;
        if (q0 > nq) q0 = nq;
        jobs[i] = (MatchJob){q, t, nq, nt, k, ratio, q0, q1, out_idx, out_dist, accept, 0};
        pthread_create(&th[i], NULL, match_worker, &jobs[i]);
    }
    long total = 0;
    for (int i = 0; i < n_threads; ++i) { pthread_join(th[i], NULL); total += jobs[i].nacc; }
    free(th); free(jobs);
    return total;
}
