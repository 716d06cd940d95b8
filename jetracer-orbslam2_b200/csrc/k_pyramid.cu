// k_pyramid.cu -- ComputePyramid stage (replaces Jetracer::pyramid_create_levels,
// reference src/cuda/pyramid.cu:31-84 and its 2x box kernel :6-29) and the 7x7 Gaussian that
// replaces gaussian_blur_3x3 (reference src/cuda/gaussian_blur_3x3.cu:15-72).
//
// Semantics = upstream ORBextractor::ComputePyramid on OpenCV 4.13:
// level l is cv::resize(level l-1, INTER_LINEAR) in 11-bit fixed point, each level framed by a
// 19-px BORDER_REFLECT_101 border.  The weight/offset tables are built once on the host exactly
// like OpenCV builds them (float32/double mix), so device arithmetic is integer only.
//
// Roofline: HBM-bound.  One thread produces 4 adjacent bytes of a padded row (one 32-bit
// coalesced store); the border is produced by the same kernel by reflecting the output coordinate
// and recomputing the pixel (no second pass over the ROI).
#include "orbb_internal.cuh"

namespace orbb {

__device__ __forceinline__ int reflect101(int p, int len) {
    // one reflection is enough: |p| excursion is <= 19 and every level is >= 62 px
    p = p < 0 ? -p : p;
    return p >= len ? 2 * len - 2 - p : p;
}

// ---- level 0: copy the caller's frames into the padded layout + border
__global__ void __launch_bounds__(128) k_level0(const uint8_t *__restrict__ in, size_t in_pitch,
                                                size_t in_stride, LevelDev L, int aligned4) {
    const int word = blockIdx.x * blockDim.x + threadIdx.x;
    const int py = blockIdx.y, frame = blockIdx.z;
    if (word * 4 >= L.pitch) return;
    const int y = reflect101(py - ORBB_BORDER, L.h);
    const uint8_t *src = in + (size_t)frame * in_stride + (size_t)y * in_pitch;
    const int b0 = word * 4;
    uint32_t out = 0;
    const int x0 = b0 - ORBB_ROI_X0;
    if (aligned4 && x0 >= 0 && x0 + 3 < L.w) {
        out = *reinterpret_cast<const uint32_t *>(src + x0);
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int px = b0 + k - ORBB_PAD_X0;
            if (px >= 0 && px < L.w + 2 * ORBB_BORDER)
                out |= (uint32_t)src[reflect101(px - ORBB_BORDER, L.w)] << (8 * k);
        }
    }
    *reinterpret_cast<uint32_t *>(L.img + (size_t)frame * L.frame_stride + (size_t)py * L.pitch + b0) = out;
}

// ---- level l from level l-1
__global__ void __launch_bounds__(128) k_resize(const LevelDev *__restrict__ levels, int l) {
    const LevelDev &L = levels[l];
    const LevelDev &S = levels[l - 1];
    const int word = blockIdx.x * blockDim.x + threadIdx.x;
    const int py = blockIdx.y, frame = blockIdx.z;
    if (word * 4 >= L.pitch) return;
    const int y = reflect101(py - ORBB_BORDER, L.h);
    const uint8_t *sroi = S.img + (size_t)frame * S.frame_stride + (size_t)ORBB_BORDER * S.pitch + ORBB_ROI_X0;
    const int b0 = word * 4;
    uint32_t out = 0;
    if (L.area2x) {
        const uint8_t *r0 = sroi + (size_t)(2 * y) * S.pitch, *r1 = r0 + S.pitch;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int px = b0 + k - ORBB_PAD_X0;
            if (px >= 0 && px < L.w + 2 * ORBB_BORDER) {
                const int x = reflect101(px - ORBB_BORDER, L.w);
                const int v = (r0[2 * x] + r0[2 * x + 1] + r1[2 * x] + r1[2 * x + 1] + 2) >> 2;
                out |= (uint32_t)v << (8 * k);
            }
        }
    } else {
        const int2 yr = __ldg(&L.yrows[y]);
        const short2 bw = __ldg(&L.ybeta[y]);
        const uint8_t *r0 = sroi + (size_t)yr.x * S.pitch, *r1 = sroi + (size_t)yr.y * S.pitch;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int px = b0 + k - ORBB_PAD_X0;
            if (px >= 0 && px < L.w + 2 * ORBB_BORDER) {
                const int x = reflect101(px - ORBB_BORDER, L.w);
                const int sx = __ldg(&L.xofs[x]);
                const int sx1 = min(sx + 1, S.w - 1);
                const short2 a = __ldg(&L.xalpha[x]);
                const int t0 = r0[sx] * a.x + r0[sx1] * a.y;
                const int t1 = r1[sx] * a.x + r1[sx1] * a.y;
                const int v = ((((int)bw.x * (t0 >> 4)) >> 16) + (((int)bw.y * (t1 >> 4)) >> 16) + 2) >> 2;
                out |= (uint32_t)v << (8 * k);
            }
        }
    }
    *reinterpret_cast<uint32_t *>(L.img + (size_t)frame * L.frame_stride + (size_t)py * L.pitch + b0) = out;
}

// ---- 7x7 sigma=2 Gaussian of the ROI, OpenCV 4.13 fixed point: k = [18,34,48,56,48,34,18]/256 per
// pass, dst = (V + 32768) >> 16.  Reads the padded level, so REFLECT_101 at the ROI
// edge is already materialised by the border.  Tile: 64 x 32 outputs per CTA of 256 threads.
#define BLUR_TW 64
#define BLUR_TH 32
__global__ void __launch_bounds__(256) k_blur(const LevelDev *__restrict__ levels, const TileEntry *__restrict__ tiles) {
    __shared__ __align__(16) uint8_t s_in[(BLUR_TH + 6) * 80];          // cols -8..71 (aligned), rows -3..34
    __shared__ __align__(16) uint16_t s_h[(BLUR_TH + 6) * BLUR_TW];     // horizontal pass, 8.8 fixed point
    const TileEntry t = tiles[blockIdx.x];
    const LevelDev &L = levels[t.level];
    const int frame = blockIdx.y;
    const int x0 = t.tx * BLUR_TW, y0 = t.ty * BLUR_TH;
    const uint8_t *roi = L.img + (size_t)frame * L.frame_stride + (size_t)ORBB_BORDER * L.pitch + ORBB_ROI_X0;
    // load rows y0-3 .. y0+34, cols x0-8 .. x0+71 as 20 words per row; clamp rows to the padded range
    for (int i = threadIdx.x; i < (BLUR_TH + 6) * 20; i += 256) {
        const int r = i / 20, wd = i - r * 20;
        int yy = y0 - 3 + r;
        yy = min(yy, L.h + ORBB_BORDER - 1);
        int xx = x0 - 8 + wd * 4;
        uint32_t v = 0;
        if (xx + 3 < L.pitch - ORBB_ROI_X0)  // stay inside the row allocation
            v = *reinterpret_cast<const uint32_t *>(roi + (ptrdiff_t)yy * L.pitch + xx);
        reinterpret_cast<uint32_t *>(s_in)[r * 20 + wd] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (BLUR_TH + 6) * BLUR_TW; i += 256) {
        const int r = i / BLUR_TW, c = i - r * BLUR_TW;
        const uint8_t *p = s_in + r * 80 + c + 8 - 3;
        const int s = 18 * (p[0] + p[6]) + 34 * (p[1] + p[5]) + 48 * (p[2] + p[4]) + 56 * p[3];
        s_h[i] = (uint16_t)s;
    }
    __syncthreads();
    // vertical pass: each thread makes 4 adjacent outputs of one row -> one 32-bit store
    for (int i = threadIdx.x; i < BLUR_TH * (BLUR_TW / 4); i += 256) {
        const int r = i / (BLUR_TW / 4), c4 = (i - r * (BLUR_TW / 4)) * 4;
        const int y = y0 + r, x = x0 + c4;
        if (y >= L.h || x >= L.w) continue;
        uint32_t out = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint16_t *q = s_h + r * BLUR_TW + c4 + k;
            const uint32_t v = 18u * (q[0] + q[6 * BLUR_TW]) + 34u * (q[BLUR_TW] + q[5 * BLUR_TW]) +
                               48u * (q[2 * BLUR_TW] + q[4 * BLUR_TW]) + 56u * q[3 * BLUR_TW];
            out |= ((v + 32768u) >> 16) << (8 * k);
        }
        *reinterpret_cast<uint32_t *>(L.blur + (size_t)frame * L.blur_stride + (size_t)y * L.pitch + x) = out;
    }
}

// ---------------------------------------------------------------- host launchers
cudaError_t launch_level0(const uint8_t *d_in, size_t pitch, size_t stride, const LevelDev &L0, int n_frames,
                          cudaStream_t st) {
    const int words = L0.pitch / 4;
    dim3 grid((words + 127) / 128, L0.rows, n_frames);
    const int aligned4 = ((reinterpret_cast<uintptr_t>(d_in) | pitch | stride) & 3) == 0;
    k_level0<<<grid, 128, 0, st>>>(d_in, pitch, stride, L0, aligned4);
    return cudaGetLastError();
}

cudaError_t launch_resize(const LevelDev *d_levels, const LevelDev &Lh, int l, int n_frames, cudaStream_t st) {
    const int words = Lh.pitch / 4;
    dim3 grid((words + 127) / 128, Lh.rows, n_frames);
    k_resize<<<grid, 128, 0, st>>>(d_levels, l);
    return cudaGetLastError();
}

cudaError_t launch_blur(const LevelDev *d_levels, const TileEntry *d_tiles, int n_tiles, int n_frames,
                        cudaStream_t st) {
    dim3 grid(n_tiles, n_frames);
    k_blur<<<grid, 256, 0, st>>>(d_levels, d_tiles);
    return cudaGetLastError();
}

}  // namespace orbb
