// k_orb.cu -- IC_Angle + steered-BRIEF descriptor, one warp per keypoint.
// Replaces Jetracer::compute_fast_angle / calc_orb (reference src/cuda/orb.cu:77-142, :17-75; the
// lossy 32-bit squeeze of :145-169 is dropped: descriptors stay 256 bit).
// Semantics:
//   IC_Angle  : integer moments m10/m01 over the 31-px disc (umax table) of the UN-blurred level,
//               angle = cv::fastAtan2 in degrees, float32 with explicit _rn ops (no FMA contraction);
//   descriptor: a = (float)cos((double)rad), b = (float)sin((double)rad); sample the 7x7-blurred
//               level at cvRound(x*b + y*a), cvRound(x*a - y*b); lane i builds byte i from pattern
//               points 16i..16i+15 (table from reference src/cuda/orb.cuh:39-297).
// The same kernel writes the final cv::KeyPoint record (pt scaled by mvScaleFactor[level]).
#include "orbb_internal.cuh"

namespace orbb {

__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale,
                p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    const float ax = fabsf(x), ay = fabsf(y);
    const float eps = 2.2204460492503131e-16f;  // (float)DBL_EPSILON
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

#define ORB_WARPS 8
#define PATCH_R 19                      // rotated pattern reach: |(13,13)| = 18.4 -> 19
#define PATCH_ROWS (2 * PATCH_R + 1)    // 39
#define PATCH_PITCH 44                  // 11 aligned words cover 39 columns at any byte phase

// umax[|v|] of upstream's circular patch, as a compile-time function of the unrolled row
__device__ __forceinline__ constexpr int umax_of(int av) {
    return (int)((0x3689ABCDDEEEFFFFull >> (4 * av)) & 15ull);
}

__global__ void __launch_bounds__(ORB_WARPS * 32)
k_angle_orb(const LevelDev *__restrict__ levels, int n_levels, const int *__restrict__ sel_count,
            const int8_t *__restrict__ pattern, const int *__restrict__ slot_level, const int *__restrict__ slot_base,
            int n_slots, orbb_keypoint *__restrict__ out_kp, uint8_t *__restrict__ out_desc,
            int *__restrict__ out_counts, int max_kp, int frame_base) {
    __shared__ __align__(16) uint8_t s_patch[ORB_WARPS][PATCH_ROWS * PATCH_PITCH];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gslot = blockIdx.x * ORB_WARPS + warp;
    const int frame = blockIdx.y + frame_base;
    if (gslot >= n_slots) return;
    const int level = slot_level[gslot], slot = gslot - slot_base[level];
    const int *cnt = sel_count + frame * n_levels;
    const LevelDev &L = levels[level];
    const int my_cnt = min(cnt[level], L.sel_cap);
    int off = 0;
    for (int l = 0; l < level; ++l) off += min(cnt[l], levels[l].sel_cap);
    if (level == 0 && slot == 0 && lane == 0) {
        int total = my_cnt;
        for (int l = 1; l < n_levels; ++l) total += min(cnt[l], levels[l].sel_cap);
        out_counts[frame] = min(total, max_kp);
    }
    if (slot >= my_cnt || off + slot >= max_kp) return;

    const uint32_t c = L.sel[(size_t)frame * L.sel_cap + slot];
    const int x = (int)(c & 0xfffu) + ORBB_MIN_BORDER, y = (int)((c >> 12) & 0xfffu) + ORBB_MIN_BORDER;
    const int resp = (int)(c >> 24) - 1;  // cv::FAST response = arc score - 1

    // ---- stage the 39x39 blurred patch with aligned 32-bit loads
    uint8_t *patch = s_patch[warp];
    const int xa = (x - PATCH_R) & ~3, poff = (x - PATCH_R) - xa;
    {
        const uint8_t *src = L.blur + (size_t)frame * L.blur_stride + (size_t)(y - PATCH_R) * L.pitch + xa;
        const int lr = lane >> 4, lw = lane & 15;  // 2 rows per warp instruction, 11 of 16 lanes active
        if (lw < 11) {
            const uint8_t *g = src + (size_t)lr * L.pitch + 4 * lw;
            const size_t step = 2 * (size_t)L.pitch;
            uint32_t v[(PATCH_ROWS + 1) / 2];
#pragma unroll
            for (int r = 0; r < (PATCH_ROWS + 1) / 2; ++r)  // all 20 loads in flight before the first store
                v[r] = (2 * r + lr < PATCH_ROWS) ? *reinterpret_cast<const uint32_t *>(g + r * step) : 0u;
#pragma unroll
            for (int r = 0; r < (PATCH_ROWS + 1) / 2; ++r)
                if (2 * r + lr < PATCH_ROWS) reinterpret_cast<uint32_t *>(patch + (2 * r + lr) * PATCH_PITCH)[lw] = v[r];
        }
    }

    // ---- IC_Angle: lane = column u (coalesced row reads), rows unrolled with their compile-time umax
    const int u = lane - 15, au = u < 0 ? -u : u;
    const int pitch = L.pitch;
    const uint8_t *rowp = L.img + (size_t)frame * L.frame_stride + (size_t)(y + ORBB_BORDER - 15) * pitch + ORBB_ROI_X0 + x + u;
    int colsum = 0, m01 = 0;
#pragma unroll
    for (int v = -15; v <= 15; ++v) {
        const int d = umax_of(v < 0 ? -v : v);
        const int val = (au <= d) ? (int)*rowp : 0;  // lane 31 has au = 16 > d
        rowp += pitch;
        colsum += val;
        m01 += v * val;
    }
    int m10 = u * colsum;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m10 += __shfl_xor_sync(0xffffffffu, m10, o);
        m01 += __shfl_xor_sync(0xffffffffu, m01, o);
    }
    const float angle = fast_atan2_deg((float)m01, (float)m10);

    // ---- steered BRIEF on the staged patch
    const float factor_pi = (float)(3.14159265358979323846 / 180.0);  // == (float)(CV_PI/180.f)
    const float rad = __fmul_rn(angle, factor_pi);
    double sd, cd;
    sincos((double)rad, &sd, &cd);
    const float a = (float)cd, b = (float)sd;
    const int8_t *pat = pattern + lane * 32;
    const int4 q0 = *reinterpret_cast<const int4 *>(pat), q1 = *reinterpret_cast<const int4 *>(pat + 16);
    const int w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
    __syncwarp();
    const uint8_t *pc = patch + PATCH_R * PATCH_PITCH + PATCH_R + poff;
    int val = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float x0 = (float)(int8_t)(w[k] & 0xff), y0 = (float)(int8_t)((w[k] >> 8) & 0xff);
        const float x1 = (float)(int8_t)((w[k] >> 16) & 0xff), y1 = (float)(int8_t)((w[k] >> 24) & 0xff);
        const int ry0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
        const int rx0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
        const int ry1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
        const int rx1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
        const int t0 = pc[ry0 * PATCH_PITCH + rx0], t1 = pc[ry1 * PATCH_PITCH + rx1];
        val |= (t0 < t1) << k;
    }
    const size_t o = (size_t)frame * max_kp + off + slot;
    out_desc[o * 32 + lane] = (uint8_t)val;
    if (lane < 7) {  // 28-byte cv::KeyPoint written as 7 coalesced words
        float f;
        switch (lane) {
            case 0: f = level ? __fmul_rn((float)x, L.scale) : (float)x; break;
            case 1: f = level ? __fmul_rn((float)y, L.scale) : (float)y; break;
            case 2: f = L.patch_size; break;
            case 3: f = angle; break;
            case 4: f = (float)resp; break;
            case 5: f = __int_as_float(level); break;
            default: f = __int_as_float(-1); break;
        }
        reinterpret_cast<float *>(out_kp + o)[lane] = f;
    }
}

cudaError_t launch_angle_orb(const LevelDev *d_levels, int n_levels, const int *d_sel_count, const int8_t *d_pattern,
                             const int *d_slot_level, const int *d_slot_base, int n_slots, int frame_base,
                             int n_frames, orbb_keypoint *d_kp, uint8_t *d_desc, int *d_counts, int max_kp, cudaStream_t st) {
    dim3 grid((n_slots + ORB_WARPS - 1) / ORB_WARPS, n_frames);
    k_angle_orb<<<grid, ORB_WARPS * 32, 0, st>>>(d_levels, n_levels, d_sel_count, d_pattern, d_slot_level, d_slot_base,
                                                 n_slots, d_kp, d_desc, d_counts, max_kp, frame_base);
    return cudaGetLastError();
}

}  // namespace orbb
