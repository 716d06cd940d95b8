// k_rgbd.cu -- RGB-D association: the stages that follow the descriptors in SlamGpuPipeline::buildStream
// (reference src/SlamGpuPipeline/buildStream.cpp:376-394, 468-487, 523-556).
//   k_align_scatter / k_align_finish  replace Jetracer::align_depth_to_other (src/cuda/cuda-align.cu:122-286,
//                                     launcher :366-399): four kernels and a 16 B/pixel int2 map in the reference,
//                                     here one scatter kernel that maps both corners of a depth pixel in registers
//                                     and issues the RED.MIN itself, plus a 128-bit sentinel->0 pass;
//   k_kp_to_point                     replaces kernel_keypoint_pixel_to_point (cuda-align.cu:282-364): depth gate,
//                                     float64 deprojection, order-preserving block compaction;
//   k_reproject                       replaces kernel_reproject_prev_points (src/cuda/post_processing.cu:72-90);
//   k_compact_pairs                   the compaction tail of kernel_match_keypoints (post_processing.cu:176-197).
// Arithmetic = the reference's rsutil.h copies (cuda-align.cu:26-120, post_processing.cu:10-43), every float32 and
// float64 operation rounded on its own (_rn intrinsics: the library is built with -fmad=false, the intrinsics make
// the intent explicit), so that a plain C restatement (oracle/rgbd_oracle.c, -ffp-contract=off) reproduces it bit
// for bit.  Bound: the scatter is L2-atomic / HBM bound (2 B read + ~4 RED per depth pixel), everything else is
// launch-latency sized (about 1000 keypoints per frame).
#include "orbb_internal.cuh"

namespace orbb {

struct AlignArgs {
    orbb_intrinsics d, o;
    orbb_extrinsics e;
    float depth_scale;
};

__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }

// radial factor 1 + c0 r2 + c1 r2 r2 + c4 r2 r2 r2, left to right as the reference writes it
__device__ __forceinline__ float radial_f(const float *c, float r2) {
    return fadd(fadd(fadd(1.0f, fmul(c[0], r2)), fmul(fmul(c[1], r2), r2)), fmul(fmul(fmul(c[4], r2), r2), r2));
}

// rs2_deproject_pixel_to_point (cuda-align.cu:58-83)
__device__ __forceinline__ void deproject(float pt[3], const orbb_intrinsics &in, float px, float py, float depth) {
    float x = __fdiv_rn(__fsub_rn(px, in.ppx), in.fx);
    float y = __fdiv_rn(__fsub_rn(py, in.ppy), in.fy);
    if (in.model == ORBB_DISTORTION_INVERSE_BROWN_CONRADY) {
        const float r2 = fadd(fmul(x, x), fmul(y, y));
        const float f = radial_f(in.coeffs, r2);
        const float ux = fadd(fadd(fmul(x, f), fmul(fmul(fmul(2.0f, in.coeffs[2]), x), y)),
                              fmul(in.coeffs[3], fadd(r2, fmul(fmul(2.0f, x), x))));
        const float uy = fadd(fadd(fmul(y, f), fmul(fmul(fmul(2.0f, in.coeffs[3]), x), y)),
                              fmul(in.coeffs[2], fadd(r2, fmul(fmul(2.0f, y), y))));
        x = ux; y = uy;
    }
    pt[0] = fmul(depth, x); pt[1] = fmul(depth, y); pt[2] = depth;
}

// the distortion + pinhole tail shared by rs2_project_point_to_pixel (cuda-align.cu:26-56) and its double-input
// twin (post_processing.cu:10-43)
__device__ __forceinline__ void project_xy(float pix[2], const orbb_intrinsics &in, float x, float y) {
    if (in.model == ORBB_DISTORTION_MODIFIED_BROWN_CONRADY) {
        const float r2 = fadd(fmul(x, x), fmul(y, y));
        const float f = radial_f(in.coeffs, r2);
        x = fmul(x, f); y = fmul(y, f);
        const float dx = fadd(fadd(x, fmul(fmul(fmul(2.0f, in.coeffs[2]), x), y)),
                              fmul(in.coeffs[3], fadd(r2, fmul(fmul(2.0f, x), x))));
        const float dy = fadd(fadd(y, fmul(fmul(fmul(2.0f, in.coeffs[3]), x), y)),
                              fmul(in.coeffs[2], fadd(r2, fmul(fmul(2.0f, y), y))));
        x = dx; y = dy;
    }
    pix[0] = fadd(fmul(x, in.fx), in.ppx);
    pix[1] = fadd(fmul(y, in.fy), in.ppy);
}

// ---- depth -> other image.  thread = one depth pixel of one frame, CTA = 32 x 8 pixels.
// Without a distortion model on the depth side the normalised corner coordinates (c -+ 0.5 - pp) / f depend only on
// the column (row): the CTA computes its 33 + 9 distinct values once (42 IEEE divisions instead of 1024) and the
// pixels read them from shared memory -- the same operations on the same operands, so the result is unchanged.
__global__ void __launch_bounds__(256) k_align_scatter(const uint16_t *__restrict__ depth, uint32_t *__restrict__ out,
                                                       const AlignArgs A) {
    __shared__ float s_nx[33], s_ny[9];
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    const bool plain = A.d.model != ORBB_DISTORTION_INVERSE_BROWN_CONRADY;
    if (plain) {
        const int t = threadIdx.y * 32 + threadIdx.x;
        // corner k of the tile: pixel (x0 + k) - 0.5, which is also pixel (x0 + k - 1) + 0.5 (both sums are exact)
        if (t < 33) s_nx[t] = __fdiv_rn(__fsub_rn(fadd((float)(blockIdx.x * 32 + t), -0.5f), A.d.ppx), A.d.fx);
        else if (t < 42) s_ny[t - 33] = __fdiv_rn(__fsub_rn(fadd((float)(blockIdx.y * 8 + t - 33), -0.5f), A.d.ppy), A.d.fy);
        __syncthreads();
    }
    if (x >= A.d.width || y >= A.d.height) return;
    const size_t f = blockIdx.z;
    const unsigned raw = depth[(f * A.d.height + y) * A.d.width + x];
    const float dv = fmul((float)(int)raw, A.depth_scale);
    if (dv == 0.0f) return;  // no depth data: nothing is written (cuda-align.cu:139-141)
    int cx[2], cy[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {  // top-left (-0.5) and bottom-right (+0.5) corner of the depth pixel
        float p[3], q[3], pix[2];
        if (plain) {
            p[0] = fmul(dv, s_nx[threadIdx.x + c]); p[1] = fmul(dv, s_ny[threadIdx.y + c]); p[2] = dv;
        } else {
            const float shift = c ? 0.5f : -0.5f;
            deproject(p, A.d, fadd((float)x, shift), fadd((float)y, shift), dv);
        }
        const float *R = A.e.rotation, *T = A.e.translation;
        q[0] = fadd(fadd(fadd(fmul(R[0], p[0]), fmul(R[3], p[1])), fmul(R[6], p[2])), T[0]);
        q[1] = fadd(fadd(fadd(fmul(R[1], p[0]), fmul(R[4], p[1])), fmul(R[7], p[2])), T[1]);
        q[2] = fadd(fadd(fadd(fmul(R[2], p[0]), fmul(R[5], p[1])), fmul(R[8], p[2])), T[2]);
        project_xy(pix, A.o, __fdiv_rn(q[0], q[2]), __fdiv_rn(q[1], q[2]));
        cx[c] = (int)fadd(pix[0], 0.5f);  // cvt.rzi: truncation, saturating, NaN -> 0
        cy[c] = (int)fadd(pix[1], 0.5f);
    }
    if (cx[0] < 0 || cy[0] < 0 || cx[1] >= A.o.width || cy[1] >= A.o.height) return;
    uint32_t *o = out + f * A.o.width * A.o.height;
    for (int yy = cy[0]; yy <= cy[1]; ++yy)
        for (int xx = cx[0]; xx <= cx[1]; ++xx) atomicMin(o + (size_t)yy * A.o.width + xx, raw);
}

// pixels nothing mapped to still hold the 0xFFFFFFFF the buffer was initialised with: they become 0
__global__ void __launch_bounds__(256) k_align_finish(uint32_t *__restrict__ out, size_t n) {
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n) {
        uint4 v = *reinterpret_cast<uint4 *>(out + i);
        if (v.x == 0xFFFFFFFFu || v.y == 0xFFFFFFFFu || v.z == 0xFFFFFFFFu || v.w == 0xFFFFFFFFu) {
            v.x = v.x == 0xFFFFFFFFu ? 0u : v.x; v.y = v.y == 0xFFFFFFFFu ? 0u : v.y;
            v.z = v.z == 0xFFFFFFFFu ? 0u : v.z; v.w = v.w == 0xFFFFFFFFu ? 0u : v.w;
            *reinterpret_cast<uint4 *>(out + i) = v;
        }
    } else {
        for (size_t k = i; k < n; ++k)
            if (out[k] == 0xFFFFFFFFu) out[k] = 0u;
    }
}

// ---- order-preserving compaction helper: CTA-wide exclusive rank of `keep` within a chunk of blockDim.x items
__device__ __forceinline__ int block_rank(bool keep, int *s_warp, int *chunk_total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    __syncthreads();  // s_warp is reused between chunks
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < nw; ++w) {
        const int c = s_warp[w];
        before += w < warp ? c : 0;
        total += c;
    }
    *chunk_total = total;
    return before + __popc(m & ((1u << lane) - 1u));
}

// deproject_pixel_to_point_double (cuda-align.cu:85-110): float32 normalisation, float64 afterwards
__device__ __forceinline__ void lift_point(const orbb_intrinsics &in, float kx, float ky, int depth, double *p) {
    double x = (double)__fdiv_rn(__fsub_rn(kx, in.ppx), in.fx);
    double y = (double)__fdiv_rn(__fsub_rn(ky, in.ppy), in.fy);
    if (in.model == ORBB_DISTORTION_INVERSE_BROWN_CONRADY) {
        const double c0 = in.coeffs[0], c1 = in.coeffs[1], c2 = in.coeffs[2], c3 = in.coeffs[3], c4 = in.coeffs[4];
        const double r2 = dadd(dmul(x, x), dmul(y, y));
        const double fr = dadd(dadd(dadd(1.0, dmul(c0, r2)), dmul(dmul(c1, r2), r2)), dmul(dmul(dmul(c4, r2), r2), r2));
        const double ux = dadd(dadd(dmul(x, fr), dmul(dmul(dmul(2.0, c2), x), y)), dmul(c3, dadd(r2, dmul(dmul(2.0, x), x))));
        const double uy = dadd(dadd(dmul(y, fr), dmul(dmul(dmul(2.0, c3), x), y)), dmul(c2, dadd(r2, dmul(dmul(2.0, y), y))));
        x = ux; y = uy;
    }
    const double dd = (double)(float)depth;
    p[0] = dmul(dd, x); p[1] = dmul(dd, y); p[2] = dd;
}

// ---- keypoints -> 3-D points.  CTA = one frame; 1024 threads, so ~1000 keypoints are one or two chunks (each chunk is
// a chain of two dependent global loads and two barriers: a lone frame took 15.6 us with 256 threads, five chunks)
#define KP_THREADS 1024
__global__ void __launch_bounds__(KP_THREADS)
k_kp_to_point(const uint32_t *__restrict__ aligned, const orbb_intrinsics in, const orbb_keypoint *__restrict__ kp_in,
              const uint4 *__restrict__ desc_in, const int *__restrict__ counts_in, int max_kp,
              orbb_keypoint *__restrict__ kp_out, uint4 *__restrict__ desc_out, double *__restrict__ points,
              int *__restrict__ valid_counts) {
    __shared__ int s_warp[KP_THREADS / 32];
    const size_t f = blockIdx.x;
    const int n = min(counts_in[f], max_kp);
    const uint32_t *dimg = aligned + f * in.width * in.height;
    int base = 0;
    // two chunks per round: both chunks' keypoint and depth loads are in flight before the first barrier (the ranking
    // barriers would otherwise serialise one dependent load chain per chunk)
    for (int i0 = 0; i0 < n; i0 += 2 * KP_THREADS) {
        orbb_keypoint k[2];
        int depth[2] = {0, 0};
        bool keep[2] = {false, false};
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = i0 + u * KP_THREADS + threadIdx.x;
            k[u] = kp_in[f * max_kp + min(i, n - 1)];
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = i0 + u * KP_THREADS + threadIdx.x;
            const int xi = (int)((double)k[u].x + 0.5), yi = (int)((double)k[u].y + 0.5);
            if (i < n && xi >= 0 && yi >= 0 && xi < in.width && yi < in.height) depth[u] = (int)dimg[(size_t)yi * in.width + xi];
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = i0 + u * KP_THREADS + threadIdx.x;
            keep[u] = i < n && depth[u] > 1 && k[u].response > 1.0f;
            if (u == 1 && i0 + KP_THREADS >= n) break;  // block-uniform: no second chunk
            int total;
            const int r = base + block_rank(keep[u], s_warp, &total);
            if (keep[u]) {
                const size_t o = f * max_kp + r;
                kp_out[o] = k[u];
                desc_out[2 * o] = desc_in[2 * (f * max_kp + i)];
                desc_out[2 * o + 1] = desc_in[2 * (f * max_kp + i) + 1];
                lift_point(in, k[u].x, k[u].y, depth[u], points + 3 * o);
            }
            base += total;
        }
    }
    if (threadIdx.x == 0) valid_counts[f] = base;
}

// ---- the same for SMALL batches (a lone frame): CTA = one chunk of 128 keypoints of one frame.  A chunk's output offset is
// the number of valid keypoints before it, which the CTA counts itself -- the gate of a keypoint is two loads (position and
// response, then the aligned depth), all of them in flight at once -- so ten CTAs on ten SMs each run ONE short chain
// instead of one CTA walking two 1024-keypoint rounds with their ranking barriers (12 us -> see profiles/).
#define KPC_THREADS 128
__global__ void __launch_bounds__(KPC_THREADS)
k_kp_to_point_chunks(const uint32_t *__restrict__ aligned, const orbb_intrinsics in, const orbb_keypoint *__restrict__ kp_in,
                     const uint4 *__restrict__ desc_in, const int *__restrict__ counts_in, int max_kp,
                     orbb_keypoint *__restrict__ kp_out, uint4 *__restrict__ desc_out, double *__restrict__ points,
                     int *__restrict__ valid_counts) {
    __shared__ int s_warp[KPC_THREADS / 32], s_before[KPC_THREADS / 32];
    const size_t f = blockIdx.y;
    const int n = min(counts_in[f], max_kp);
    const int c0 = blockIdx.x * KPC_THREADS;
    if (c0 >= n && !(n == 0 && blockIdx.x == 0)) return;  // block-uniform
    const uint32_t *dimg = aligned + f * in.width * in.height;
    const orbb_keypoint *kin = kp_in + f * max_kp;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    auto gate = [&](float x, float y, float response, bool in_range) -> int {  // the aligned depth if the keypoint is kept, else 0
        const int xi = (int)((double)x + 0.5), yi = (int)((double)y + 0.5);
        int d = 0;
        if (in_range && xi >= 0 && yi >= 0 && xi < in.width && yi < in.height) d = (int)dimg[(size_t)yi * in.width + xi];
        return (d > 1 && response > 1.0f) ? d : 0;
    };
    // own keypoint: the whole record and its descriptor travel with the gate's loads
    const int i = c0 + threadIdx.x;
    const bool mine = i < n;
    const orbb_keypoint k = kin[mine ? i : 0];
    const uint4 d0 = desc_in[2 * (f * max_kp + (mine ? i : 0))], d1 = desc_in[2 * (f * max_kp + (mine ? i : 0)) + 1];
    // valid keypoints of the chunks before this one
    int before = 0;
#pragma unroll 4
    for (int j = threadIdx.x; j < c0; j += KPC_THREADS) before += gate(kin[j].x, kin[j].y, kin[j].response, true) != 0;
    const int depth = gate(k.x, k.y, k.response, mine);
    const unsigned m = __ballot_sync(0xffffffffu, depth != 0);
    before = __reduce_add_sync(0xffffffffu, before);
    if (lane == 0) { s_warp[warp] = __popc(m); s_before[warp] = before; }
    __syncthreads();
    int base = 0, own_before = 0, own_total = 0;
#pragma unroll
    for (int w2 = 0; w2 < KPC_THREADS / 32; ++w2) {
        base += s_before[w2];
        own_before += w2 < warp ? s_warp[w2] : 0;
        own_total += s_warp[w2];
    }
    if (depth != 0) {
        const size_t o = f * max_kp + base + own_before + __popc(m & ((1u << lane) - 1u));
        kp_out[o] = k;
        desc_out[2 * o] = d0; desc_out[2 * o + 1] = d1;
        lift_point(in, k.x, k.y, depth, points + 3 * o);
    }
    if (threadIdx.x == 0 && c0 + KPC_THREADS >= n) valid_counts[f] = base + own_total;  // the frame's last chunk
}

// ---- reprojection of the previous frame's points.  thread = one point.
__global__ void __launch_bounds__(128)
k_reproject(const double *__restrict__ points, const int *__restrict__ counts, int max_kp, const double *__restrict__ T,
            const orbb_intrinsics in, float2 *__restrict__ pos_out) {
    const size_t f = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= min(counts[f], max_kp)) return;
    const double *p = points + 3 * (f * max_kp + i);
    double e[3] = {p[0], p[1], p[2]};
    if (T) {
        // Eigen::Matrix4d * Vector4d, column-major; a fixed-size row.dot(v) reduces pairwise: (a0 + a1) + (a2 + a3)
        const double *M = T + 16 * f;
#pragma unroll
        for (int r = 0; r < 3; ++r)
            e[r] = dadd(dadd(dmul(M[r], p[0]), dmul(M[4 + r], p[1])), dadd(dmul(M[8 + r], p[2]), M[12 + r]));
    }
    float pix[2];
    project_xy(pix, in, (float)__ddiv_rn(e[0], e[2]), (float)__ddiv_rn(e[1], e[2]));
    pos_out[f * max_kp + i] = make_float2(pix[0], pix[1]);
}

// ---- matched 3-D pairs, compacted in query order.  CTA = one frame pair.
__global__ void __launch_bounds__(KP_THREADS)
k_compact_pairs(const int *__restrict__ idx, const int *__restrict__ q_counts, int max_kp,
                const double *__restrict__ q_points, const double *__restrict__ t_points,
                const uint8_t *__restrict__ t_xy, int t_stride, double *__restrict__ prev_out,
                double *__restrict__ curr_out, uint16_t *__restrict__ xy_out, int *__restrict__ n_matched) {
    __shared__ int s_warp[KP_THREADS / 32];
    const size_t f = blockIdx.x;
    const int n = min(q_counts[f], max_kp);
    int base = 0;
    for (int i0 = 0; i0 < n; i0 += KP_THREADS) {
        const int i = i0 + threadIdx.x;
        const int t = i < n ? idx[f * max_kp + i] : -1;
        int total;
        const int r = base + block_rank(t >= 0, s_warp, &total);
        if (t >= 0) {
            const size_t o = f * max_kp + r, q = f * max_kp + i, tt = f * max_kp + t;
            if (prev_out && q_points) { prev_out[3 * o] = q_points[3 * q]; prev_out[3 * o + 1] = q_points[3 * q + 1]; prev_out[3 * o + 2] = q_points[3 * q + 2]; }
            if (curr_out && t_points) { curr_out[3 * o] = t_points[3 * tt]; curr_out[3 * o + 1] = t_points[3 * tt + 1]; curr_out[3 * o + 2] = t_points[3 * tt + 2]; }
            if (xy_out) {
                const float *p = reinterpret_cast<const float *>(t_xy + tt * t_stride);
                xy_out[(2 * f) * max_kp + r] = (uint16_t)p[0];
                xy_out[(2 * f + 1) * max_kp + r] = (uint16_t)p[1];
            }
        }
        base += total;
    }
    if (threadIdx.x == 0 && n_matched) n_matched[f] = base;
}

// ---- the same for SMALL batches: CTA = one chunk of 128 queries of one frame pair; the chunk's output offset (matches
// before it) is one int load per earlier query, counted by the CTA itself (see k_kp_to_point_chunks).
__global__ void __launch_bounds__(KPC_THREADS)
k_compact_pairs_chunks(const int *__restrict__ idx, const int *__restrict__ q_counts, int max_kp,
                       const double *__restrict__ q_points, const double *__restrict__ t_points,
                       const uint8_t *__restrict__ t_xy, int t_stride, double *__restrict__ prev_out,
                       double *__restrict__ curr_out, uint16_t *__restrict__ xy_out, int *__restrict__ n_matched) {
    __shared__ int s_warp[KPC_THREADS / 32], s_before[KPC_THREADS / 32];
    const size_t f = blockIdx.y;
    const int n = min(q_counts[f], max_kp);
    const int c0 = blockIdx.x * KPC_THREADS;
    if (c0 >= n && !(n == 0 && blockIdx.x == 0)) return;  // block-uniform
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = c0 + threadIdx.x;
    const int t = i < n ? idx[f * max_kp + i] : -1;
    int before = 0;
#pragma unroll 4
    for (int j = threadIdx.x; j < c0; j += KPC_THREADS) before += idx[f * max_kp + j] >= 0;
    // the pair's records do not depend on the rank: fetch them next to the counting loads
    double qp[3] = {0, 0, 0}, tp[3] = {0, 0, 0};
    float tx = 0.f, ty = 0.f;
    if (t >= 0) {
        const size_t q = f * max_kp + i, tt = f * max_kp + t;
        if (prev_out && q_points) { qp[0] = q_points[3 * q]; qp[1] = q_points[3 * q + 1]; qp[2] = q_points[3 * q + 2]; }
        if (curr_out && t_points) { tp[0] = t_points[3 * tt]; tp[1] = t_points[3 * tt + 1]; tp[2] = t_points[3 * tt + 2]; }
        if (xy_out) {
            const float *p = reinterpret_cast<const float *>(t_xy + tt * t_stride);
            tx = p[0]; ty = p[1];
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, t >= 0);
    before = __reduce_add_sync(0xffffffffu, before);
    if (lane == 0) { s_warp[warp] = __popc(m); s_before[warp] = before; }
    __syncthreads();
    int base = 0, own_before = 0, own_total = 0;
#pragma unroll
    for (int w2 = 0; w2 < KPC_THREADS / 32; ++w2) {
        base += s_before[w2];
        own_before += w2 < warp ? s_warp[w2] : 0;
        own_total += s_warp[w2];
    }
    if (t >= 0) {
        const int r = base + own_before + __popc(m & ((1u << lane) - 1u));
        const size_t o = f * max_kp + r;
        if (prev_out && q_points) { prev_out[3 * o] = qp[0]; prev_out[3 * o + 1] = qp[1]; prev_out[3 * o + 2] = qp[2]; }
        if (curr_out && t_points) { curr_out[3 * o] = tp[0]; curr_out[3 * o + 1] = tp[1]; curr_out[3 * o + 2] = tp[2]; }
        if (xy_out) {
            xy_out[(2 * f) * max_kp + r] = (uint16_t)tx;
            xy_out[(2 * f + 1) * max_kp + r] = (uint16_t)ty;
        }
    }
    if (threadIdx.x == 0 && c0 + KPC_THREADS >= n && n_matched) n_matched[f] = base + own_total;
}

// ---- A lone frame's depth gate + 3-D lift + windowed match in ONE launch (the RGB-D stage's tail for max_batch = 1).
// k_kp_to_point_chunks followed by k_match_windowed were two dependent launches of ~8 and ~11 us for ~1 us of work each;
// the match only needs to know WHICH current keypoints survive the depth gate and where they land after compaction,
// and that costs a CTA two loads per keypoint.  So every CTA (= four previous-frame points, one per warp, as in
// k_match_windowed) gates ALL current keypoints itself and builds the compacted position table in shared memory --
// compacted index, position, and the raw index under which the descriptor still lives -- and CTA c < ceil(n / 128)
// additionally writes chunk c of the compacted keypoints / descriptors / 3-D points (the work of k_kp_to_point_chunks).
// Results are identical to the two-kernel form: same gate, same compaction order, same packed (distance, index) key.
#define GM_THREADS 128
#define GM_MAX_KP 2048   // 16 gate slots per thread; 20 KB of shared memory
__global__ void __launch_bounds__(GM_THREADS)
k_gate_match_lone(const uint32_t *__restrict__ aligned, const orbb_intrinsics in, const orbb_keypoint *__restrict__ kp_raw,
                  const uint4 *__restrict__ desc_raw, const int *__restrict__ count_raw, int max_kp,
                  orbb_keypoint *__restrict__ kp_out, uint4 *__restrict__ desc_out, double *__restrict__ points,
                  int *__restrict__ valid_out, int *__restrict__ count_blk,
                  const uint4 *__restrict__ q_desc, const float2 *__restrict__ q_pos, const int *__restrict__ q_count,
                  float max_px, int max_hamming, int *__restrict__ out_idx, int *__restrict__ out_dist) {
    __shared__ float2 s_txy[GM_MAX_KP];
    __shared__ uint16_t s_raw[GM_MAX_KP];
    __shared__ int s_cnt[GM_MAX_KP / 32 + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = min(count_raw[0], max_kp), nq = min(q_count[0], max_kp);
    const int nchunks = (n + GM_THREADS - 1) / GM_THREADS;
    const int q = blockIdx.x * (GM_THREADS / 32) + warp;
    if (blockIdx.x * (GM_THREADS / 32) >= nq && (int)blockIdx.x >= nchunks && blockIdx.x != 0) return;  // block-uniform
    // the query's descriptor and position travel with the gate's loads
    uint4 qa = make_uint4(0u, 0u, 0u, 0u), qb = qa;
    float2 qp = make_float2(0.f, 0.f);
    const bool active = q < nq;
    if (active) { qa = q_desc[(size_t)q * 2]; qb = q_desc[(size_t)q * 2 + 1]; qp = q_pos[q]; }
    // ---- gate of every current keypoint: item i = c * 128 + thread, all loads of all chunks in flight
    float kx[GM_MAX_KP / GM_THREADS], ky[GM_MAX_KP / GM_THREADS], kr[GM_MAX_KP / GM_THREADS];
    int depth[GM_MAX_KP / GM_THREADS];
#pragma unroll
    for (int c = 0; c < GM_MAX_KP / GM_THREADS; ++c) {
        const int i = c * GM_THREADS + threadIdx.x;
        kx[c] = ky[c] = kr[c] = 0.f;
        if (c < nchunks && i < n) { kx[c] = kp_raw[i].x; ky[c] = kp_raw[i].y; kr[c] = kp_raw[i].response; }
    }
#pragma unroll
    for (int c = 0; c < GM_MAX_KP / GM_THREADS; ++c) {
        const int i = c * GM_THREADS + threadIdx.x;
        const int xi = (int)((double)kx[c] + 0.5), yi = (int)((double)ky[c] + 0.5);
        depth[c] = 0;
        if (c < nchunks && i < n && xi >= 0 && yi >= 0 && xi < in.width && yi < in.height) depth[c] = (int)aligned[(size_t)yi * in.width + xi];
    }
    unsigned keep = 0;  // bit c: item kept; rank[c]: its place among the kept items of its warp step
    int rank[GM_MAX_KP / GM_THREADS];
#pragma unroll
    for (int c = 0; c < GM_MAX_KP / GM_THREADS; ++c) {
        const bool k = c < nchunks && depth[c] > 1 && kr[c] > 1.0f;
        const unsigned m = __ballot_sync(0xffffffffu, k);
        rank[c] = __popc(m & ((1u << lane) - 1u));
        if (k) keep |= 1u << c;
        if (lane == 0 && c < nchunks) s_cnt[c * 4 + warp] = __popc(m);
    }
    __syncthreads();
    if (warp == 0) {  // exclusive prefix over the (chunk, warp) counts in raw order: <= 64 entries, two per lane
        const int e0 = 2 * lane, e1 = 2 * lane + 1, ne = nchunks * 4;
        const int v0 = e0 < ne ? s_cnt[e0] : 0, v1 = e1 < ne ? s_cnt[e1] : 0;
        int inc = v0 + v1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        const int total = __shfl_sync(0xffffffffu, inc, 31);
        __syncwarp();
        if (e0 < ne) s_cnt[e0] = inc - v0 - v1;
        if (e1 < ne) s_cnt[e1] = inc - v1;
        if (lane == 0) s_cnt[GM_MAX_KP / 32] = total;
    }
    __syncthreads();
    const int nt = s_cnt[GM_MAX_KP / 32];
#pragma unroll
    for (int c = 0; c < GM_MAX_KP / GM_THREADS; ++c) {
        if (!((keep >> c) & 1u)) continue;
        const int i = c * GM_THREADS + threadIdx.x, ci = s_cnt[c * 4 + warp] + rank[c];
        s_txy[ci] = make_float2(kx[c], ky[c]);
        s_raw[ci] = (uint16_t)i;
        if ((int)blockIdx.x == c) {  // this CTA writes chunk c of the compacted frame
            kp_out[ci] = kp_raw[i];
            desc_out[2 * ci] = desc_raw[2 * (size_t)i];
            desc_out[2 * ci + 1] = desc_raw[2 * (size_t)i + 1];
            lift_point(in, kx[c], ky[c], depth[c], points + 3 * (size_t)ci);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { valid_out[0] = nt; if (count_blk) count_blk[0] = count_raw[0]; }
    __syncthreads();
    if (!active) return;
    // ---- windowed 1-NN of this warp's previous-frame point against the gated keypoints (k_match_windowed)
    unsigned best = 0xffffffffu;
    for (int tb = 0; tb < nt; tb += 32) {
        const int i = tb + lane;
        if (i < nt) {
            const float2 pt = s_txy[i];
            if (fabsf(__fsub_rn(qp.x, pt.x)) <= max_px && fabsf(__fsub_rn(qp.y, pt.y)) <= max_px) {
                const size_t t = s_raw[i];
                const uint4 a = __ldg(desc_raw + t * 2), b = __ldg(desc_raw + t * 2 + 1);
                const int d = __popc(qa.x ^ a.x) + __popc(qa.y ^ a.y) + __popc(qa.z ^ a.z) + __popc(qa.w ^ a.w) +
                              __popc(qb.x ^ b.x) + __popc(qb.y ^ b.y) + __popc(qb.z ^ b.z) + __popc(qb.w ^ b.w);
                if (d < max_hamming) best = min(best, ((unsigned)d << 16) | (unsigned)i);
            }
        }
    }
    best = __reduce_min_sync(0xffffffffu, best);
    if (lane == 0) {
        const bool ok = best != 0xffffffffu;
        out_idx[q] = ok ? (int)(best & 0xffffu) : -1;
        out_dist[q] = ok ? (int)(best >> 16) : -1;
    }
}

cudaError_t launch_gate_match_lone(const uint32_t *d_aligned, const orbb_intrinsics &in, const orbb_keypoint *kp_raw,
                                   const uint8_t *desc_raw, const int *count_raw, int max_kp, orbb_keypoint *kp_out,
                                   uint8_t *desc_out, double *points, int *valid_out, int *count_blk, const uint8_t *q_desc,
                                   const float *q_pos, const int *q_count, float max_px, int max_hamming, int *out_idx,
                                   int *out_dist, cudaStream_t st) {
    if (max_kp > GM_MAX_KP) return cudaErrorInvalidValue;
    const int grid = (max_kp + GM_THREADS / 32 - 1) / (GM_THREADS / 32);
    k_gate_match_lone<<<grid, GM_THREADS, 0, st>>>(d_aligned, in, kp_raw, reinterpret_cast<const uint4 *>(desc_raw), count_raw,
                                                   max_kp, kp_out, reinterpret_cast<uint4 *>(desc_out), points, valid_out, count_blk,
                                                   reinterpret_cast<const uint4 *>(q_desc), reinterpret_cast<const float2 *>(q_pos),
                                                   q_count, max_px, max_hamming, out_idx, out_dist);
    return cudaGetLastError();
}

// ---- RGB8 -> gray (reference cuda_RGB_to_Grayscale.cu:10-24).  thread = 4 pixels: three aligned 32-bit loads, one
// 32-bit store (the reference: 3 byte loads + 1 byte store per thread).  float64 like the reference's expression.
__global__ void __launch_bounds__(256) k_rgb_to_gray(const uint8_t *__restrict__ rgb, size_t rgb_pitch, size_t rgb_stride,
                                                     int width, int height, uint8_t *__restrict__ gray, size_t gray_pitch,
                                                     size_t gray_stride, int aligned) {
    const int q = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    const int x = 4 * q;
    if (x >= width || y >= height) return;
    const uint8_t *src = rgb + blockIdx.z * rgb_stride + (size_t)y * rgb_pitch + 3 * (size_t)x;
    uint8_t *dst = gray + blockIdx.z * gray_stride + (size_t)y * gray_pitch + x;
    uint8_t px[12];
    const int n = min(4, width - x);
    if (aligned && n == 4) {
        const uint32_t *s32 = reinterpret_cast<const uint32_t *>(src);
        const uint32_t a = __ldg(s32), b = __ldg(s32 + 1), c = __ldg(s32 + 2);
#pragma unroll
        for (int k = 0; k < 4; ++k) { px[k] = (a >> (8 * k)) & 255u; px[4 + k] = (b >> (8 * k)) & 255u; px[8 + k] = (c >> (8 * k)) & 255u; }
    } else {
#pragma unroll
        for (int k = 0; k < 12; ++k) px[k] = k < 3 * n ? src[k] : 0;
    }
    uint32_t out = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double R = (double)(float)px[3 * k], G = (double)(float)px[3 * k + 1], B = (double)(float)px[3 * k + 2];
        const double v = dadd(dadd(dadd(dmul(B, 0.07), dmul(G, 0.72)), dmul(R, 0.21)), 0.5);
        out |= (uint32_t)(uint8_t)(int)floor(v) << (8 * k);
    }
    if (aligned && n == 4) *reinterpret_cast<uint32_t *>(dst) = out;
    else
        for (int k = 0; k < n; ++k) dst[k] = (uint8_t)(out >> (8 * k));
}

// ---------------------------------------------------------------- host launchers
cudaError_t launch_align(const uint16_t *d_depth, int n_frames, float depth_scale, const orbb_intrinsics &di,
                         const orbb_intrinsics &oi, const orbb_extrinsics &ex, uint32_t *d_out, cudaStream_t st) {
    const size_t n = (size_t)n_frames * oi.width * oi.height;
    cudaError_t e = cudaMemsetAsync(d_out, 0xFF, n * sizeof(uint32_t), st);
    if (e != cudaSuccess) return e;
    AlignArgs A;
    A.d = di; A.o = oi; A.e = ex; A.depth_scale = depth_scale;
    dim3 grid((di.width + 31) / 32, (di.height + 7) / 8, n_frames), block(32, 8);
    k_align_scatter<<<grid, block, 0, st>>>(d_depth, d_out, A);
    k_align_finish<<<(unsigned)((n / 4 + 256) / 256), 256, 0, st>>>(d_out, n);
    return cudaGetLastError();
}

cudaError_t launch_kp_to_point(const uint32_t *d_aligned, const orbb_intrinsics &in, int n_frames, const orbb_keypoint *kp_in,
                               const uint8_t *desc_in, const int *counts_in, int max_kp, orbb_keypoint *kp_out,
                               uint8_t *desc_out, double *points, int *valid, cudaStream_t st) {
    if (n_frames <= 4) {  // small batches: one CTA per 128-keypoint chunk
        dim3 grid((max_kp + KPC_THREADS - 1) / KPC_THREADS, n_frames);
        k_kp_to_point_chunks<<<grid, KPC_THREADS, 0, st>>>(d_aligned, in, kp_in, reinterpret_cast<const uint4 *>(desc_in), counts_in,
                                                           max_kp, kp_out, reinterpret_cast<uint4 *>(desc_out), points, valid);
        return cudaGetLastError();
    }
    k_kp_to_point<<<n_frames, KP_THREADS, 0, st>>>(d_aligned, in, kp_in, reinterpret_cast<const uint4 *>(desc_in), counts_in,
                                                   max_kp, kp_out, reinterpret_cast<uint4 *>(desc_out), points, valid);
    return cudaGetLastError();
}

cudaError_t launch_reproject(const double *points, const int *counts, int n_frames, int max_kp, const double *T,
                             const orbb_intrinsics &in, float *pos_out, cudaStream_t st) {
    dim3 grid((max_kp + 127) / 128, n_frames);
    k_reproject<<<grid, 128, 0, st>>>(points, counts, max_kp, T, in, reinterpret_cast<float2 *>(pos_out));
    return cudaGetLastError();
}

cudaError_t launch_compact_pairs(const int *idx, const int *q_counts, int n_frames, int max_kp, const double *q_points,
                                 const double *t_points, const void *t_xy, int t_stride, double *prev_out,
                                 double *curr_out, uint16_t *xy_out, int *n_matched, cudaStream_t st) {
    if (n_frames <= 4) {  // small batches: one CTA per 128-query chunk
        dim3 grid((max_kp + KPC_THREADS - 1) / KPC_THREADS, n_frames);
        k_compact_pairs_chunks<<<grid, KPC_THREADS, 0, st>>>(idx, q_counts, max_kp, q_points, t_points,
                                                             static_cast<const uint8_t *>(t_xy), t_stride, prev_out, curr_out,
                                                             xy_out, n_matched);
        return cudaGetLastError();
    }
    k_compact_pairs<<<n_frames, KP_THREADS, 0, st>>>(idx, q_counts, max_kp, q_points, t_points,
                                                     static_cast<const uint8_t *>(t_xy), t_stride, prev_out, curr_out, xy_out,
                                                     n_matched);
    return cudaGetLastError();
}

// Stage bookkeeping between batches in ONE launch (it used to be five device-to-device copies, 2-3 us each inside a graph):
// the previous batch's last frame (row `carry` of the keypoint / descriptor / point / valid-count arrays) becomes row 0, and
// the extraction's per-frame counts are copied into the result block.  Rows are multiples of 4 bytes.
__global__ void k_stage_carry(uint32_t *kp, uint32_t *desc, uint32_t *pts, int *valid, int carry, int kp_w, int desc_w, int pts_w,
                              int *counts_dst, const int *counts_src, int n_counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_counts) counts_dst[i] = counts_src[i];
    if (carry <= 0) return;
    if (i == 0) valid[0] = valid[carry];
    if (i < kp_w) kp[i] = kp[(size_t)carry * kp_w + i];
    else if (i < kp_w + desc_w) desc[i - kp_w] = desc[(size_t)carry * desc_w + i - kp_w];
    else if (i < kp_w + desc_w + pts_w) pts[i - kp_w - desc_w] = pts[(size_t)carry * pts_w + i - kp_w - desc_w];
}

cudaError_t launch_stage_carry(orbb_keypoint *kp, uint8_t *desc, double *pts, int *valid, int carry, int max_kp, int *counts_dst,
                               const int *counts_src, int n_counts, cudaStream_t st) {
    const int kp_w = (int)(sizeof(orbb_keypoint) / 4) * max_kp, desc_w = 8 * max_kp, pts_w = 6 * max_kp;
    const int total = carry > 0 ? kp_w + desc_w + pts_w : 0;
    const int n = total > n_counts ? total : n_counts;
    k_stage_carry<<<(n + 255) / 256, 256, 0, st>>>(reinterpret_cast<uint32_t *>(kp), reinterpret_cast<uint32_t *>(desc),
                                                   reinterpret_cast<uint32_t *>(pts), valid, carry, kp_w, desc_w, pts_w, counts_dst,
                                                   counts_src, n_counts);
    return cudaGetLastError();
}

cudaError_t launch_rgb_to_gray(const uint8_t *d_rgb, size_t rgb_pitch, size_t rgb_stride, int w, int h, int n_frames,
                               uint8_t *d_gray, size_t gray_pitch, size_t gray_stride, cudaStream_t st) {
    const int aligned = ((reinterpret_cast<uintptr_t>(d_rgb) | rgb_pitch | rgb_stride | reinterpret_cast<uintptr_t>(d_gray) |
                          gray_pitch | gray_stride) & 3) == 0;
    dim3 grid(((w + 3) / 4 + 31) / 32, (h + 7) / 8, n_frames), block(32, 8);
    k_rgb_to_gray<<<grid, block, 0, st>>>(d_rgb, rgb_pitch, rgb_stride, w, h, d_gray, gray_pitch, gray_stride, aligned);
    return cudaGetLastError();
}

}  // namespace orbb
