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

#define ORB_WARPS 4
// keypoint slots per warp = template parameter KPW: the per-keypoint scalar math runs one slot per lane, so large
// batches use 8 (fewest instructions per keypoint); small batches use 2 or 1 so that a lone frame's ~1000 keypoints
// spread over all SMs instead of queueing 8 deep behind 40 CTAs (848x480 single frame: 18.7 -> see DESIGN.md)
#define PATCH_R 19                      // rotated pattern reach: |(13,13)| = 18.4 -> 19
#define PATCH_ROWS (2 * PATCH_R + 1)    // 39
#define PATCH_PITCH 44                  // 11 aligned words cover 39 columns at any byte phase

// umax[|v|] of upstream's circular patch, as a compile-time function of the unrolled row
__device__ __forceinline__ constexpr int umax_of(int av) {
    return (int)((0x3689ABCDDEEEFFFFull >> (4 * av)) & 15ull);
}

// warp = KPW consecutive keypoint slots, warp-synchronous (no block barriers):
//   phase 0  lanes 0..7: slot -> (level, x, y, response, output index), kept in registers and broadcast by shuffle
//   phase A  per slot: IC_Angle moments, lane = patch column (coalesced row reads of the un-blurred level)
//   phase S  lanes 0..7: fastAtan2, sincos(double) and the cv::KeyPoint record -- the scalar chain that a
//            warp-per-keypoint kernel would execute 32x redundantly now runs once per 8 keypoints
//   phase B  per slot: stage the 39x39 blurred patch, lane i builds descriptor byte i (its 16 pattern points sit in
//            registers as floats, converted once per warp)
template <int ORB_KPW>
__global__ void __launch_bounds__(ORB_WARPS * 32, 8)
k_angle_orb(const LevelDev *__restrict__ levels, int n_levels, const int *__restrict__ sel_count,
            const int8_t *__restrict__ pattern, const int *__restrict__ slot_level, const int *__restrict__ slot_base,
            int n_slots, orbb_keypoint *__restrict__ out_kp, uint8_t *__restrict__ out_desc,
            int *__restrict__ out_counts, int max_kp, int frame_base) {
    __shared__ __align__(16) uint8_t s_patch[ORB_WARPS][PATCH_ROWS * PATCH_PITCH];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int frame = blockIdx.y + frame_base;
    const int *cnt = sel_count + frame * n_levels;

    // ---- phase 0
    int my_o = -1, my_x = 0, my_y = 0, my_level = 0, my_resp = 0;
    {
        const int gslot = (blockIdx.x * ORB_WARPS + warp) * ORB_KPW + lane;
        if (lane < ORB_KPW && gslot < n_slots) {
            my_level = slot_level[gslot];
            const int slot = gslot - slot_base[my_level];
            const LevelDev &L = levels[my_level];
            const int my_cnt = min(cnt[my_level], L.sel_cap);
            int off = 0;
            for (int l = 0; l < my_level; ++l) off += min(cnt[l], levels[l].sel_cap);
            if (gslot == 0) {
                int total = my_cnt;
                for (int l = 1; l < n_levels; ++l) total += min(cnt[l], levels[l].sel_cap);
                out_counts[frame] = min(total, max_kp);
            }
            if (slot < my_cnt && off + slot < max_kp) {
                const uint32_t c = L.sel[(size_t)frame * L.sel_cap + slot];
                my_x = (int)(c & 0xfffu) + ORBB_MIN_BORDER;
                my_y = (int)((c >> 12) & 0xfffu) + ORBB_MIN_BORDER;
                my_resp = (int)(c >> 24) - 1;  // cv::FAST response = arc score - 1
                my_o = off + slot;
            }
        }
    }
    const unsigned live = __ballot_sync(FULL, my_o >= 0);
    if (live == 0) return;

    // ---- phase A: IC_Angle moments; rows unrolled with their compile-time umax
    const int u = lane - 15, au = u < 0 ? -u : u;
    int my_m01 = 0, my_m10 = 0;
#pragma unroll 1
    for (int i = 0; i < ORB_KPW; ++i) {
        if (!((live >> i) & 1u)) continue;
        const LevelDev &L = levels[__shfl_sync(FULL, my_level, i)];
        const int pitch = L.pitch;
        const uint8_t *rowp = L.img + (size_t)frame * L.frame_stride +
                              (size_t)(__shfl_sync(FULL, my_y, i) + ORBB_BORDER - 15) * pitch + ORBB_ROI_X0 +
                              __shfl_sync(FULL, my_x, i) + u;
        int colsum = 0, m01 = 0;
#pragma unroll
        for (int v = -15; v <= 15; ++v) {
            const int d = umax_of(v < 0 ? -v : v);
            const int val = (au <= d) ? (int)*rowp : 0;  // lane 31 has au = 16 > d
            rowp += pitch;
            colsum += val;
            m01 += v * val;
        }
        const int m10 = __reduce_add_sync(FULL, u * colsum);
        m01 = __reduce_add_sync(FULL, m01);
        if (lane == i) { my_m01 = m01; my_m10 = m10; }
    }

    // ---- phase S
    float my_a = 0.f, my_b = 0.f;
    if (my_o >= 0) {
        const float angle = fast_atan2_deg((float)my_m01, (float)my_m10);
        const float factor_pi = (float)(3.14159265358979323846 / 180.0);  // == (float)(CV_PI/180.f)
        const float rad = __fmul_rn(angle, factor_pi);
        double sd, cd;
        sincos((double)rad, &sd, &cd);
        my_a = (float)cd; my_b = (float)sd;
        const LevelDev &L = levels[my_level];
        float *kp = reinterpret_cast<float *>(out_kp + (size_t)frame * max_kp + my_o);
        kp[0] = my_level ? __fmul_rn((float)my_x, L.scale) : (float)my_x;
        kp[1] = my_level ? __fmul_rn((float)my_y, L.scale) : (float)my_y;
        kp[2] = L.patch_size;
        kp[3] = angle;
        kp[4] = (float)my_resp;
        kp[5] = __int_as_float(my_level);
        kp[6] = __int_as_float(-1);
    }
    // pattern points of this lane's descriptor byte, as floats
    float px0[8], py0[8], px1[8], py1[8];
    {
        const int8_t *pat = pattern + lane * 32;
        const int4 q0 = *reinterpret_cast<const int4 *>(pat), q1 = *reinterpret_cast<const int4 *>(pat + 16);
        const int w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            px0[k] = (float)(int8_t)(w[k] & 0xff); py0[k] = (float)(int8_t)((w[k] >> 8) & 0xff);
            px1[k] = (float)(int8_t)((w[k] >> 16) & 0xff); py1[k] = (float)(int8_t)((w[k] >> 24) & 0xff);
        }
    }

    // ---- phase B: steered BRIEF on the staged blurred patch
    uint8_t *patch = s_patch[warp];
#pragma unroll 1
    for (int i = 0; i < ORB_KPW; ++i) {
        if (!((live >> i) & 1u)) continue;
        const LevelDev &L = levels[__shfl_sync(FULL, my_level, i)];
        const int x = __shfl_sync(FULL, my_x, i), y = __shfl_sync(FULL, my_y, i), o = __shfl_sync(FULL, my_o, i);
        const float a = __shfl_sync(FULL, my_a, i), b = __shfl_sync(FULL, my_b, i);
        const int xa = (x - PATCH_R) & ~3, poff = (x - PATCH_R) - xa;
        __syncwarp();  // the previous slot's reads of the patch are done
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
        __syncwarp();
        const uint8_t *pc = patch + PATCH_R * PATCH_PITCH + PATCH_R + poff;
        int val = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int ry0 = __float2int_rn(__fadd_rn(__fmul_rn(px0[k], b), __fmul_rn(py0[k], a)));
            const int rx0 = __float2int_rn(__fsub_rn(__fmul_rn(px0[k], a), __fmul_rn(py0[k], b)));
            const int ry1 = __float2int_rn(__fadd_rn(__fmul_rn(px1[k], b), __fmul_rn(py1[k], a)));
            const int rx1 = __float2int_rn(__fsub_rn(__fmul_rn(px1[k], a), __fmul_rn(py1[k], b)));
            const int t0 = pc[ry0 * PATCH_PITCH + rx0], t1 = pc[ry1 * PATCH_PITCH + rx1];
            val |= (t0 < t1) << k;
        }
        out_desc[((size_t)frame * max_kp + o) * 32 + lane] = (uint8_t)val;
    }
}

cudaError_t launch_angle_orb(const LevelDev *d_levels, int n_levels, const int *d_sel_count, const int8_t *d_pattern,
                             const int *d_slot_level, const int *d_slot_base, int n_slots, int frame_base,
                             int n_frames, orbb_keypoint *d_kp, uint8_t *d_desc, int *d_counts, int max_kp, cudaStream_t st) {
    const long long slots = (long long)n_slots * n_frames;
    const int kpw = slots >= 32768 ? 8 : (slots >= 8192 ? 2 : 1);
    const int per_cta = ORB_WARPS * kpw;
    dim3 grid((n_slots + per_cta - 1) / per_cta, n_frames);
#define ORBB_LAUNCH_ORB(K)                                                                                             \
    k_angle_orb<K><<<grid, ORB_WARPS * 32, 0, st>>>(d_levels, n_levels, d_sel_count, d_pattern, d_slot_level, d_slot_base, \
                                                    n_slots, d_kp, d_desc, d_counts, max_kp, frame_base)
    if (kpw == 8) ORBB_LAUNCH_ORB(8);
    else if (kpw == 2) ORBB_LAUNCH_ORB(2);
    else ORBB_LAUNCH_ORB(1);
#undef ORBB_LAUNCH_ORB
    return cudaGetLastError();
}

// ---------------------------------------------------------------- the reference's two-call shape
// Jetracer::compute_fast_angle / calc_orb take caller-owned arrays (float2 positions, one image) and are called
// separately (reference src/cuda/orb.cuh:9-27, call sites src/SlamGpuPipeline/buildStream.cpp:442-460).  These two
// kernels keep that shape with upstream arithmetic; the fused k_angle_orb above is the batch path.  Warp = keypoint.
__global__ void __launch_bounds__(128)
k_fast_angle_pos(float *__restrict__ angle, const float2 *__restrict__ pos, const uint8_t *__restrict__ img, int pitch, int w,
                 int h, int n) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, k = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (k >= n) return;
    const float2 p = pos[k];
    const int cx = __float2int_rn(p.x), cy = __float2int_rn(p.y);  // cvRound
    if (cx < 15 || cy < 15 || cx >= w - 15 || cy >= h - 15) {       // the disc would leave the image
        if (lane == 0) angle[k] = -1.0f;
        return;
    }
    const int u = lane - 15, au = u < 0 ? -u : u;
    const uint8_t *rowp = img + (size_t)(cy - 15) * pitch + cx + (lane < 31 ? u : 0);
    int colsum = 0, m01 = 0;
#pragma unroll
    for (int v = -15; v <= 15; ++v) {
        const int d = umax_of(v < 0 ? -v : v);
        const int val = (au <= d) ? (int)__ldg(rowp) : 0;  // lane 31 has au = 16 > d
        rowp += pitch;
        colsum += val;
        m01 += v * val;
    }
    const int m10 = __reduce_add_sync(FULL, u * colsum);
    m01 = __reduce_add_sync(FULL, m01);
    if (lane == 0) angle[k] = fast_atan2_deg((float)m01, (float)m10);
}

__global__ void __launch_bounds__(128)
k_calc_orb_pos(const float *__restrict__ angle, const float2 *__restrict__ pos, uint8_t *__restrict__ desc,
               const uint8_t *__restrict__ img, int pitch, int w, int h, int n, const int8_t *__restrict__ pattern) {
    const int lane = threadIdx.x & 31, k = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (k >= n) return;
    const float2 p = pos[k];
    const int cx = __float2int_rn(p.x), cy = __float2int_rn(p.y);
    if (cx < 18 || cy < 18 || cx > w - 19 || cy > h - 19) {  // rotated pattern reach: cvRound(13 * sqrt 2) = 18
        desc[(size_t)k * 32 + lane] = 0;
        return;
    }
    const float factor_pi = (float)(3.14159265358979323846 / 180.0);
    const float rad = __fmul_rn(angle[k], factor_pi);
    double sd, cd;
    sincos((double)rad, &sd, &cd);
    const float a = (float)cd, b = (float)sd;
    const int8_t *pat = pattern + lane * 32;
    const int4 q0 = *reinterpret_cast<const int4 *>(pat), q1 = *reinterpret_cast<const int4 *>(pat + 16);
    const int wd[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
    const uint8_t *pc = img + (size_t)cy * pitch + cx;
    int val = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float x0 = (float)(int8_t)(wd[j] & 0xff), y0 = (float)(int8_t)((wd[j] >> 8) & 0xff);
        const float x1 = (float)(int8_t)((wd[j] >> 16) & 0xff), y1 = (float)(int8_t)((wd[j] >> 24) & 0xff);
        const int ry0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
        const int rx0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
        const int ry1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
        const int rx1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
        const int t0 = __ldg(pc + (ptrdiff_t)ry0 * pitch + rx0), t1 = __ldg(pc + (ptrdiff_t)ry1 * pitch + rx1);
        val |= (t0 < t1) << j;
    }
    desc[(size_t)k * 32 + lane] = (uint8_t)val;
}

// SoA view of the quadtree selection (the reference's detect() outputs): thread = (frame, slot)
__global__ void __launch_bounds__(128)
k_detect_export(const LevelDev *__restrict__ levels, int n_levels, const int *__restrict__ sel_count,
                const int *__restrict__ slot_level, const int *__restrict__ slot_base, int n_slots, float2 *__restrict__ pos,
                float *__restrict__ score, int *__restrict__ level, int *__restrict__ level_counts, int *__restrict__ counts,
                int max_kp) {
    const int gslot = blockIdx.x * blockDim.x + threadIdx.x, frame = blockIdx.y;
    if (gslot >= n_slots) return;
    const int *cnt = sel_count + frame * n_levels;
    const int l = slot_level[gslot], slot = gslot - slot_base[l];
    const LevelDev &L = levels[l];
    int off = 0;
    for (int i = 0; i < l; ++i) off += min(cnt[i], levels[i].sel_cap);
    if (gslot == 0) {
        int total = 0;
        for (int i = 0; i < n_levels; ++i) {
            const int c = min(cnt[i], levels[i].sel_cap);
            if (level_counts) level_counts[frame * n_levels + i] = max(0, min(c, max_kp - total));
            total += c;
        }
        if (counts) counts[frame] = min(total, max_kp);
    }
    if (slot >= min(cnt[l], L.sel_cap) || off + slot >= max_kp) return;
    const uint32_t c = L.sel[(size_t)frame * L.sel_cap + slot];
    const size_t o = (size_t)frame * max_kp + off + slot;
    if (pos) pos[o] = make_float2((float)((int)(c & 0xfffu) + ORBB_MIN_BORDER), (float)((int)((c >> 12) & 0xfffu) + ORBB_MIN_BORDER));
    if (score) score[o] = (float)((int)(c >> 24) - 1);
    if (level) level[o] = l;
}

cudaError_t launch_fast_angle_pos(float *d_angle, const float *d_pos, const uint8_t *d_img, int pitch, int w, int h, int n,
                                  cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    k_fast_angle_pos<<<(n + 3) / 4, 128, 0, st>>>(d_angle, reinterpret_cast<const float2 *>(d_pos), d_img, pitch, w, h, n);
    return cudaGetLastError();
}

cudaError_t launch_calc_orb_pos(const float *d_angle, const float *d_pos, uint8_t *d_desc, const uint8_t *d_img, int pitch, int w,
                                int h, int n, const int8_t *d_pattern, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    k_calc_orb_pos<<<(n + 3) / 4, 128, 0, st>>>(d_angle, reinterpret_cast<const float2 *>(d_pos), d_desc, d_img, pitch, w, h, n,
                                                d_pattern);
    return cudaGetLastError();
}

cudaError_t launch_detect_export(const LevelDev *d_levels, int n_levels, const int *d_sel_count, const int *d_slot_level,
                                 const int *d_slot_base, int n_slots, int n_frames, float *d_pos, float *d_score, int *d_level,
                                 int *d_level_counts, int *d_counts, int max_kp, cudaStream_t st) {
    dim3 grid((n_slots + 127) / 128, n_frames);
    k_detect_export<<<grid, 128, 0, st>>>(d_levels, n_levels, d_sel_count, d_slot_level, d_slot_base, n_slots,
                                          reinterpret_cast<float2 *>(d_pos), d_score, d_level, d_level_counts, d_counts, max_kp);
    return cudaGetLastError();
}

}  // namespace orbb
