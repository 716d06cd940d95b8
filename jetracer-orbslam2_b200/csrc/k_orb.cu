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

__global__ void __launch_bounds__(ORB_WARPS * 32)
k_angle_orb(const LevelDev *__restrict__ levels, int n_levels, const int *__restrict__ sel_count,
            const int8_t *__restrict__ pattern, const int *__restrict__ slot_level, const int *__restrict__ slot_base,
            int n_slots, orbb_keypoint *__restrict__ out_kp, uint8_t *__restrict__ out_desc,
            int *__restrict__ out_counts, int max_kp) {
    const int lane = threadIdx.x & 31;
    const int gslot = blockIdx.x * ORB_WARPS + (threadIdx.x >> 5);
    const int frame = blockIdx.y;
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

    // ---- IC_Angle: lane r handles disc row v = r - 15
    const uint8_t *center = L.img + (size_t)frame * L.frame_stride + (size_t)(y + ORBB_BORDER) * L.pitch + ORBB_ROI_X0 + x;
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const int v = lane - 15, av = v < 0 ? -v : v;
        const int d = (int)((0x3689ABCDDEEEFFFFull >> (4 * av)) & 15ull);  // umax[|v|]
        const uint8_t *row = center + (ptrdiff_t)v * L.pitch;
        int s = 0, sx = 0;
        for (int u = -d; u <= d; ++u) {
            const int val = row[u];
            s += val;
            sx += u * val;
        }
        m10 = sx;
        m01 = v * s;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m10 += __shfl_xor_sync(0xffffffffu, m10, o);
        m01 += __shfl_xor_sync(0xffffffffu, m01, o);
    }
    const float angle = fast_atan2_deg((float)m01, (float)m10);

    // ---- steered BRIEF on the blurred level
    const float factor_pi = (float)(3.14159265358979323846 / 180.0);  // == (float)(CV_PI/180.f)
    const float rad = __fmul_rn(angle, factor_pi);
    const float a = (float)cos((double)rad), b = (float)sin((double)rad);
    const uint8_t *bc = L.blur + (size_t)frame * L.blur_stride + (size_t)y * L.pitch + x;
    const int8_t *pat = pattern + lane * 32;
    const int4 q0 = *reinterpret_cast<const int4 *>(pat), q1 = *reinterpret_cast<const int4 *>(pat + 16);
    const int w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
    int val = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float x0 = (float)(int8_t)(w[k] & 0xff), y0 = (float)(int8_t)((w[k] >> 8) & 0xff);
        const float x1 = (float)(int8_t)((w[k] >> 16) & 0xff), y1 = (float)(int8_t)((w[k] >> 24) & 0xff);
        const int ry0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
        const int rx0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
        const int ry1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
        const int rx1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
        const int t0 = bc[(ptrdiff_t)ry0 * L.pitch + rx0], t1 = bc[(ptrdiff_t)ry1 * L.pitch + rx1];
        val |= (t0 < t1) << k;
    }
    const size_t o = (size_t)frame * max_kp + off + slot;
    out_desc[o * 32 + lane] = (uint8_t)val;
    if (lane == 0) {
        orbb_keypoint kp;
        kp.x = level ? __fmul_rn((float)x, L.scale) : (float)x;
        kp.y = level ? __fmul_rn((float)y, L.scale) : (float)y;
        kp.size = L.patch_size;
        kp.angle = angle;
        kp.response = (float)resp;
        kp.octave = level;
        kp.class_id = -1;
        out_kp[o] = kp;
    }
}

cudaError_t launch_angle_orb(const LevelDev *d_levels, int n_levels, const int *d_sel_count, const int8_t *d_pattern,
                             const int *d_slot_level, const int *d_slot_base, int n_slots, int n_frames,
                             orbb_keypoint *d_kp, uint8_t *d_desc, int *d_counts, int max_kp, cudaStream_t st) {
    dim3 grid((n_slots + ORB_WARPS - 1) / ORB_WARPS, n_frames);
    k_angle_orb<<<grid, ORB_WARPS * 32, 0, st>>>(d_levels, n_levels, d_sel_count, d_pattern, d_slot_level, d_slot_base,
                                                 n_slots, d_kp, d_desc, d_counts, max_kp);
    return cudaGetLastError();
}

}  // namespace orbb
