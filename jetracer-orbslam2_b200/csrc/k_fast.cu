// k_fast.cu -- FAST-9 detection with the per-cell 20/7 threshold fallback and 3x3 score NMS
// (replaces Jetracer::fast_gpu_calc_corner_response + grid_nms, reference src/cuda/fast.cu:150-372
// and src/cuda/nms.cu:86-296; semantics = upstream ComputeKeyPointsOctTree's per-cell cv::FAST calls,
// SURVEY.md A.3).
//
// Work item = one upstream 30-px cell, processed by ONE WARP with warp-synchronous code only (no
// block barriers): the cell's (cw+6)x(ch+6) window is staged in shared memory with aligned 32-bit
// loads, then
//   phase 1  antipodal-pair precheck at the low threshold, ballot-compacted into a queue,
//   phase 2  exact threshold-free arc score m = max over 9-arcs of min|d| (3-input min/max network,
//            VIMNMX3), corners (m > t_lo) written to a score tile and re-compacted,
//   phase 3  3x3 strict NMS inside the cell (the score tile has a zero frame: neighbours outside the
//            cell's tested range count as 0, exactly like cv::FAST on the cell window),
//   phase 4  survivors are appended to the (frame, level) candidate list, one atomicAdd per 32.
// Phases 1-3 run at the high threshold first and are repeated at the low threshold only when the
// cell came out empty -- upstream's FAST(iniTh) -> FAST(minTh) fallback, decided after NMS.
// No score map ever goes to HBM (the DUMP instantiation exists only for the parity tests).
// Bound: SM issue slots / shared-memory bandwidth, not HBM.
#include "orbb_internal.cuh"

namespace orbb {

#define FAST_WARPS 4

__device__ __forceinline__ int min3(int a, int b, int c) { return __vimin3_s32(a, b, c); }
__device__ __forceinline__ int max3(int a, int b, int c) { return __vimax3_s32(a, b, c); }

// m = max(A, B), A = max over the 16 9-arcs of min(v - ring), B = same for (ring - v).
__device__ __forceinline__ int arc_score(const uint8_t *p, int tp) {
    const int v = p[0];
    int d[16];
    d[0] = v - p[3 * tp];       d[1] = v - p[3 * tp + 1];   d[2] = v - p[2 * tp + 2];   d[3] = v - p[tp + 3];
    d[4] = v - p[3];            d[5] = v - p[-tp + 3];      d[6] = v - p[-2 * tp + 2];  d[7] = v - p[-3 * tp + 1];
    d[8] = v - p[-3 * tp];      d[9] = v - p[-3 * tp - 1];  d[10] = v - p[-2 * tp - 2]; d[11] = v - p[-tp - 3];
    d[12] = v - p[-3];          d[13] = v - p[tp - 3];      d[14] = v - p[2 * tp - 2];  d[15] = v - p[3 * tp - 1];
    int lo3[16], hi3[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        lo3[k] = min3(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
        hi3[k] = max3(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
    }
    int a = -256, b = 256;
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
        const int l0 = min3(lo3[k], lo3[(k + 3) & 15], lo3[(k + 6) & 15]);
        const int l1 = min3(lo3[k + 1], lo3[(k + 4) & 15], lo3[(k + 7) & 15]);
        a = max3(a, l0, l1);
        const int h0 = max3(hi3[k], hi3[(k + 3) & 15], hi3[(k + 6) & 15]);
        const int h1 = max3(hi3[k + 1], hi3[(k + 4) & 15], hi3[(k + 7) & 15]);
        b = min3(b, h0, h1);
    }
    return max(a, -b);
}

template <bool DUMP>
__global__ void __launch_bounds__(FAST_WARPS * 32)
k_fast_cells(const LevelDev *__restrict__ levels, const CellEntry *__restrict__ cells, int n_cells, int n_levels,
             int *__restrict__ cand_count, int t_lo, int t_hi, FastSmemCfg cfg, int frame_base,
             uint8_t *__restrict__ dump, const long long *__restrict__ dump_off) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cell_id = blockIdx.x * FAST_WARPS + warp;
    if (cell_id >= n_cells) return;
    const int frame = blockIdx.y + frame_base;
    const CellEntry c = cells[cell_id];
    const LevelDev &L = levels[c.level];

    uint8_t *tile = smem + (size_t)warp * cfg.warp_bytes;
    uint8_t *score = tile + cfg.tile_pitch * cfg.tile_rows;
    uint16_t *queue = reinterpret_cast<uint16_t *>(score + cfg.score_pitch * cfg.score_rows);
    const int tp = cfg.tile_pitch, sp = cfg.score_pitch;
    const int cw = c.cw, ch = c.ch;

    // ---- stage the window: rows y0-3 .. y0+ch+2, columns from the 4-aligned address at/below x0-3
    const int xa = (c.x0 - 3) & ~3, off = (c.x0 - 3) - xa;
    const int nwords = (off + cw + 6 + 3) >> 2;
    {
        const uint8_t *roi = L.img + (size_t)frame * L.frame_stride + (size_t)ORBB_BORDER * L.pitch + ORBB_ROI_X0;
        const int rpi = 32 / nwords;  // rows per iteration (nwords <= 32 guaranteed by the host)
        const int lr = lane / nwords, lw = lane - lr * nwords;
        const uint8_t *src = roi + (ptrdiff_t)(c.y0 - 3) * L.pitch + xa + 4 * lw;
        for (int r = lr; r < ch + 6; r += rpi)
            if (lr < rpi)
                reinterpret_cast<uint32_t *>(tile + r * tp)[lw] =
                    *reinterpret_cast<const uint32_t *>(src + (ptrdiff_t)r * L.pitch);
        for (int i = lane; i < (sp * cfg.score_rows) >> 2; i += 32) reinterpret_cast<uint32_t *>(score)[i] = 0;
    }
    __syncwarp();

    const int npix = cw * ch;
    const unsigned inv_cw = (1u << 20) / (unsigned)cw + 1u;  // exact floor(idx/cw) for idx*cw < 2^20
    const unsigned lt_mask = (1u << lane) - 1u;

    // Upstream order: FAST(ini) on the cell; only if that leaves nothing, FAST(min).  Running the high
    // threshold first rejects most pixels in the precheck (the low-threshold pass is rare on textured input).
    int kn = 0;
    for (int pass = DUMP ? 1 : 0; pass < 2; ++pass) {
        const int thr = pass == 0 ? t_hi : t_lo;
        if (pass == 1 && !DUMP && t_lo == t_hi) break;
        // ---- phase 1: antipodal-pair precheck (any 9-arc holds one pixel of every antipodal pair)
        int qn = 0;
        for (int base = 0; base < npix; base += 32) {
            const int idx = base + lane;
            bool ok = false;
            if (idx < npix) {
                const int y = (int)(((unsigned)idx * inv_cw) >> 20), x = idx - y * cw;
                const uint8_t *p = tile + (y + 3) * tp + x + 3 + off;
                const int v = p[0], lo = v - thr, hi = v + thr;
                const int r0 = p[3 * tp], r8 = p[-3 * tp], r4 = p[3], r12 = p[-3];
                const bool dark = ((r0 < lo) | (r8 < lo)) & ((r4 < lo) | (r12 < lo));
                const bool bright = ((r0 > hi) | (r8 > hi)) & ((r4 > hi) | (r12 > hi));
                ok = dark | bright;
            }
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            if (ok) queue[qn + __popc(m & lt_mask)] = (uint16_t)idx;
            qn += __popc(m);
        }
        __syncwarp();

        // ---- phase 2: exact arc score; corners (m > thr) go to the score tile and stay queued
        int cn = 0;
        for (int base = 0; base < qn; base += 32) {
            const int i = base + lane;
            int idx = 0, m = 0;
            if (i < qn) {
                idx = queue[i];
                const int y = (int)(((unsigned)idx * inv_cw) >> 20), x = idx - y * cw;
                m = arc_score(tile + (y + 3) * tp + x + 3 + off, tp);
                m = m > thr ? m : 0;
                if (m) score[(y + 1) * sp + x + 1] = (uint8_t)min(m, 255);
            }
            __syncwarp();
            const unsigned bm = __ballot_sync(0xffffffffu, m != 0);
            if (m) queue[cn + __popc(bm & lt_mask)] = (uint16_t)idx;
            cn += __popc(bm);
            __syncwarp();
        }

        if (DUMP) {
            uint8_t *out = dump + dump_off[c.level];
            for (int idx = lane; idx < npix; idx += 32) {
                const int y = idx / cw, x = idx - y * cw;
                out[(size_t)(c.y0 + y) * L.w + c.x0 + x] = score[(y + 1) * sp + x + 1];
            }
            return;  // parity dump only: no candidates are emitted
        }

        // ---- phase 3: strict 3x3 NMS inside the cell (scores written by an earlier pass stay valid:
        // the arc score does not depend on the threshold)
        kn = 0;
        for (int base = 0; base < cn; base += 32) {
            const int i = base + lane;
            int idx = 0;
            bool keep = false;
            if (i < cn) {
                idx = queue[i];
                const int y = (int)(((unsigned)idx * inv_cw) >> 20), x = idx - y * cw;
                const uint8_t *s = score + (y + 1) * sp + x + 1;
                const int v = s[0];
                const int n0 = max3(s[-sp - 1], s[-sp], s[-sp + 1]);
                const int n1 = max3(s[-1], s[1], s[sp - 1]);
                const int n2 = max3(s[sp], s[sp + 1], n0);
                keep = v > max(n1, n2);
            }
            __syncwarp();
            const unsigned bm = __ballot_sync(0xffffffffu, keep);
            if (keep) queue[kn + __popc(bm & lt_mask)] = (uint16_t)idx;
            kn += __popc(bm);
            __syncwarp();
        }
        if (kn > 0) break;  // cv::FAST(ini) found keypoints: no fallback for this cell
    }

    // ---- phase 4: per-cell threshold decision + append to the (frame, level) candidate list
    int *counter = cand_count + frame * n_levels + c.level;
    uint32_t *cand = L.cand + (size_t)frame * L.cand_cap;
    for (int base = 0; base < kn; base += 32) {
        const int i = base + lane;
        bool emit = false;
        uint32_t packed = 0;
        if (i < kn) {
            const int idx = queue[i];
            const int y = (int)(((unsigned)idx * inv_cw) >> 20), x = idx - y * cw;
            const int m = score[(y + 1) * sp + x + 1];
            emit = true;
            // coordinates relative to (minBorderX, minBorderY) = (16,16), as upstream's vToDistributeKeys
            packed = (uint32_t)(c.x0 + x - ORBB_MIN_BORDER) | ((uint32_t)(c.y0 + y - ORBB_MIN_BORDER) << 12) |
                     ((uint32_t)m << 24);
        }
        const unsigned bm = __ballot_sync(0xffffffffu, emit);
        if (bm) {
            int slot = 0;
            if (lane == 0) slot = atomicAdd(counter, __popc(bm));
            slot = __shfl_sync(0xffffffffu, slot, 0);
            if (emit) {
                const int dst = slot + __popc(bm & lt_mask);
                if (dst < L.cand_cap) cand[dst] = packed;
            }
        }
    }
}

cudaError_t launch_fast(const LevelDev *d_levels, const CellEntry *d_cells, int n_cells, int n_levels,
                        int *d_cand_count, int t_lo, int t_hi, const FastSmemCfg &cfg, int frame_base,
                        int n_frames, cudaStream_t st) {
    dim3 grid((n_cells + FAST_WARPS - 1) / FAST_WARPS, n_frames);
    const size_t smem = (size_t)cfg.warp_bytes * FAST_WARPS;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_fast_cells<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    k_fast_cells<false><<<grid, FAST_WARPS * 32, smem, st>>>(d_levels, d_cells, n_cells, n_levels, d_cand_count,
                                                            t_lo, t_hi, cfg, frame_base, nullptr, nullptr);
    return cudaGetLastError();
}

// parity-test variant: one frame, dumps the per-pixel score (m > t_lo ? m : 0) of every cell
cudaError_t launch_fast_dump(const LevelDev *d_levels, const CellEntry *d_cells, int n_cells, int n_levels, int t_lo,
                             int t_hi, const FastSmemCfg &cfg, int frame, uint8_t *d_dump,
                             const long long *d_dump_off, cudaStream_t st) {
    dim3 grid((n_cells + FAST_WARPS - 1) / FAST_WARPS, 1);
    const size_t smem = (size_t)cfg.warp_bytes * FAST_WARPS;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_fast_cells<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    k_fast_cells<true><<<grid, FAST_WARPS * 32, smem, st>>>(d_levels, d_cells, n_cells, n_levels, nullptr, t_lo, t_hi,
                                                           cfg, frame, d_dump, d_dump_off);
    return cudaGetLastError();
}

}  // namespace orbb
