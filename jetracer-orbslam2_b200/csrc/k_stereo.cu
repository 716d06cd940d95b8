// k_stereo.cu -- stereo association of a batch of rectified pairs with ORB-SLAM2 Frame::ComputeStereoMatches
// semantics (SURVEY.md 8f-3 "stereo epipolar-band matching for cfg 2/3"; upstream raulmur/ORB_SLAM2 src/Frame.cc,
// un-vendored, no pin in the reference -- the reference itself has no stereo path):
//   row band   right keypoint R is a candidate of left keypoint L iff floor(yR - r) <= (int)yL <= ceil(yR + r),
//              r = 2 * scale[octave R]; octave band +-1; uL - maxD <= uR <= uL (maxD = bf / (bf / fx));
//   descriptor best Hamming distance < TH_HIGH (100), ties -> lowest right index (upstream's row-table order);
//   sub-pixel  if best < (TH_HIGH + TH_LOW) / 2 = 75: 11x11 SAD of the centre-subtracted patches on the left key's
//              pyramid level, right window slid over +-5 px, parabola through the three costs around the minimum;
//   filter     matches whose SAD is >= 1.5 * 1.4 * median SAD are dropped.
// The patches are read from the padded pyramid levels resident in the handle: upstream's window may reach 10 px left
// of the right level's ROI (its bounds test is off by 2L), which lands in the reflect-101 frame both there and here.
// Launch-latency sized work (about 1000 keypoints per pair): no attempt at a roofline.
#include "orbb_internal.cuh"

namespace orbb {

#define ST_THREADS 256
#define ST_TILE 128

struct StereoArgs {
    float sf[ORBB_MAX_LEVELS], inv_sf[ORBB_MAX_LEVELS];
    int n_levels;
    float mbf, min_d, max_d;
};

__device__ __forceinline__ int pix(const LevelDev &L, int frame, int x, int y) {
    return L.img[(size_t)frame * L.frame_stride + (size_t)(ORBB_BORDER + y) * L.pitch + ORBB_ROI_X0 + x];
}

// blockIdx.y = stereo pair (left = resident frame 2p, right = 2p + 1), blockIdx.x = a block of ST_THREADS left keypoints;
// thread = one left keypoint.  (One CTA per pair walked all left keypoints: 64 pairs kept 64 of 148 SMs busy.)
__global__ void __launch_bounds__(ST_THREADS)
k_stereo_match(const LevelDev *__restrict__ levels, const orbb_keypoint *__restrict__ kp, const uint4 *__restrict__ desc,
               const int *__restrict__ counts, int max_kp, const StereoArgs A, float *__restrict__ uright,
               float *__restrict__ depth, int *__restrict__ sad) {
    __shared__ uint4 s_d[ST_TILE * 2];
    __shared__ float s_u[ST_TILE];
    __shared__ int s_lo[ST_TILE], s_hi[ST_TILE], s_oct[ST_TILE];
    const int p = blockIdx.y, fl = 2 * p, fr = 2 * p + 1;
    const size_t rowL = (size_t)fl * max_kp, rowR = (size_t)fr * max_kp, rowO = (size_t)p * max_kp;
    const int nl = min(counts[fl], max_kp), nr = min(counts[fr], max_kp);
    for (int i0 = blockIdx.x * ST_THREADS; i0 < nl; i0 += gridDim.x * ST_THREADS) {
        const int iL = i0 + threadIdx.x;
        const bool live = iL < nl;
        const orbb_keypoint kL = kp[rowL + (live ? iL : 0)];
        const uint4 qa = desc[(rowL + (live ? iL : 0)) * 2], qb = desc[(rowL + (live ? iL : 0)) * 2 + 1];
        const int row = (int)kL.y, levelL = kL.octave;
        const float minU = __fsub_rn(kL.x, A.max_d), maxU = __fsub_rn(kL.x, A.min_d);
        int best = 100, best_i = -1;  // ORBmatcher::TH_HIGH
        for (int tb = 0; tb < nr; tb += ST_TILE) {
            const int cnt = min(ST_TILE, nr - tb);
            __syncthreads();
            for (int i = threadIdx.x; i < cnt * 2; i += ST_THREADS) s_d[i] = desc[(rowR + tb) * 2 + i];
            for (int i = threadIdx.x; i < cnt; i += ST_THREADS) {
                const orbb_keypoint &k = kp[rowR + tb + i];
                const float r = __fmul_rn(2.0f, A.sf[min(max(k.octave, 0), A.n_levels - 1)]);
                s_u[i] = k.x; s_oct[i] = k.octave;
                s_lo[i] = (int)floorf(__fsub_rn(k.y, r)); s_hi[i] = (int)ceilf(__fadd_rn(k.y, r));
            }
            __syncthreads();
            if (live && !(maxU < 0.0f))
                for (int t = 0; t < cnt; ++t) {
                    if (row < s_lo[t] || row > s_hi[t]) continue;
                    if (s_oct[t] < levelL - 1 || s_oct[t] > levelL + 1) continue;
                    const float uR = s_u[t];
                    if (!(uR >= minU && uR <= maxU)) continue;
                    const uint4 a = s_d[2 * t], b = s_d[2 * t + 1];
                    const int d = __popc(qa.x ^ a.x) + __popc(qa.y ^ a.y) + __popc(qa.z ^ a.z) + __popc(qa.w ^ a.w) +
                                  __popc(qb.x ^ b.x) + __popc(qb.y ^ b.y) + __popc(qb.z ^ b.z) + __popc(qb.w ^ b.w);
                    if (d < best) { best = d; best_i = tb + t; }
                }
        }
        if (!live) continue;
        float out_u = -1.0f, out_z = -1.0f;
        int out_sad = -1;
        if (best_i >= 0 && best < 75) {  // thOrbDist = (TH_HIGH + TH_LOW) / 2
            const int oct = min(max(levelL, 0), A.n_levels - 1);
            const LevelDev &L = levels[oct];
            const float uR0 = kp[rowR + best_i].x, isf = A.inv_sf[oct];
            const float su = roundf(__fmul_rn(kL.x, isf)), sv = roundf(__fmul_rn(kL.y, isf)), sr = roundf(__fmul_rn(uR0, isf));
            const int w = 5, Lw = 5;
            const float iniu = sr + Lw - w, endu = sr + Lw + w + 1;
            if (!(iniu < 0.0f || endu >= (float)L.w)) {
                const int xl = (int)su, yl = (int)sv, xr = (int)sr;
                const int cL = pix(L, fl, xl, yl);
                int best_sad = 0x7fffffff, best_inc = 0;
                float dists[11];
#pragma unroll 1
                for (int inc = -Lw; inc <= Lw; ++inc) {
                    const int cR = pix(L, fr, xr + inc, yl);
                    int s = 0;
                    for (int dy = -w; dy <= w; ++dy)
                        for (int dx = -w; dx <= w; ++dx)
                            s += abs((pix(L, fl, xl + dx, yl + dy) - cL) - (pix(L, fr, xr + inc + dx, yl + dy) - cR));
                    if (s < best_sad) { best_sad = s; best_inc = inc; }
                    dists[Lw + inc] = (float)s;
                }
                if (best_inc != -Lw && best_inc != Lw) {
                    float d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
                    for (int k = 1; k < 10; ++k)
                        if (k == Lw + best_inc) { d1 = dists[k - 1]; d2 = dists[k]; d3 = dists[k + 1]; }
                    const float deltaR = __fdiv_rn(__fsub_rn(d1, d3), __fmul_rn(2.0f, __fsub_rn(__fadd_rn(d1, d3), __fmul_rn(2.0f, d2))));
                    if (!(deltaR < -1.0f || deltaR > 1.0f)) {
                        float bestuR = __fmul_rn(A.sf[oct], __fadd_rn(__fadd_rn(sr, (float)best_inc), deltaR));
                        float disparity = __fsub_rn(kL.x, bestuR);
                        if (disparity >= A.min_d && disparity < A.max_d) {
                            if (disparity <= 0.0f) {
                                disparity = 0.01f;
                                bestuR = (float)((double)kL.x - 0.01);
                            }
                            out_z = __fdiv_rn(A.mbf, disparity);
                            out_u = bestuR;
                            out_sad = best_sad;
                        }
                    }
                }
            }
        }
        uright[rowO + iL] = out_u; depth[rowO + iL] = out_z; sad[rowO + iL] = out_sad;
    }
}

// median filter of the SAD costs: rank selection by counting ((cost, index) order), then the 2.1 x median cut
__global__ void __launch_bounds__(ST_THREADS)
k_stereo_median(const int *__restrict__ counts, int max_kp, float *__restrict__ uright, float *__restrict__ depth,
                const int *__restrict__ sad, int *__restrict__ n_stereo) {
    extern __shared__ int s_sad[];  // [max_kp]
    __shared__ int s_m, s_median, s_left;
    const int p = blockIdx.x, nl = min(counts[2 * p], max_kp);
    const size_t rowO = (size_t)p * max_kp;
    if (threadIdx.x == 0) { s_m = 0; s_median = 0; s_left = 0; }
    __syncthreads();
    int mine = 0;
    for (int i = threadIdx.x; i < nl; i += ST_THREADS) { const int v = sad[rowO + i]; s_sad[i] = v; mine += v >= 0; }
    atomicAdd(&s_m, mine);
    __syncthreads();
    const int m = s_m;
    if (m > 0) {
        for (int i = threadIdx.x; i < nl; i += ST_THREADS) {
            const int v = s_sad[i];
            if (v < 0) continue;
            int rank = 0;
            for (int j = 0; j < nl; ++j) {
                const int u = s_sad[j];
                rank += (u >= 0) && (u < v || (u == v && j < i));
            }
            if (rank == m / 2) s_median = v;
        }
        __syncthreads();
        const float th = __fmul_rn(__fmul_rn(1.5f, 1.4f), (float)s_median);
        int kept = 0;
        for (int i = threadIdx.x; i < nl; i += ST_THREADS) {
            const int v = s_sad[i];
            if (v < 0) continue;
            if ((float)v < th) ++kept;
            else { uright[rowO + i] = -1.0f; depth[rowO + i] = -1.0f; }
        }
        atomicAdd(&s_left, kept);
        __syncthreads();
    }
    if (threadIdx.x == 0 && n_stereo) n_stereo[p] = s_left;
}

cudaError_t launch_stereo(const LevelDev *d_levels, const float *sf, const float *inv_sf, int n_levels, const orbb_keypoint *d_kp,
                          const uint8_t *d_desc, const int *d_counts, int max_kp, int n_pairs, float mbf, float fx,
                          float *d_uright, float *d_depth, int *d_sad, int *d_nstereo, cudaStream_t st) {
    StereoArgs A{};
    for (int l = 0; l < n_levels && l < ORBB_MAX_LEVELS; ++l) { A.sf[l] = sf[l]; A.inv_sf[l] = inv_sf[l]; }
    A.n_levels = n_levels;
    const float mb = mbf / fx;  // upstream: mb = mbf / fx; minZ = mb; minD = 0; maxD = mbf / minZ
    A.mbf = mbf; A.min_d = 0.0f; A.max_d = mbf / mb;
    k_stereo_match<<<dim3((max_kp + ST_THREADS - 1) / ST_THREADS, n_pairs), ST_THREADS, 0, st>>>(d_levels, d_kp, reinterpret_cast<const uint4 *>(d_desc), d_counts, max_kp, A,
                                                   d_uright, d_depth, d_sad);
    const size_t smem = sizeof(int) * (size_t)max_kp;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_stereo_median, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    k_stereo_median<<<n_pairs, ST_THREADS, smem, st>>>(d_counts, max_kp, d_uright, d_depth, d_sad, d_nstereo);
    return cudaGetLastError();
}

}  // namespace orbb
