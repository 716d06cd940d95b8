// orbb_internal.cuh -- structures shared by the host API and the sm_100a kernels.
//
// HBM layout (one handle, batch of B frames, level l):
//   img[l]   : B x (h_l+38) rows x pitch_l bytes.  pitch_l = roundup(32 + w_l + 19, 128).  A row is
//              [13 unused][19 border][w_l ROI][19 border][pad]: ROI pixel 0 sits at byte 32, so ROI
//              rows are 16-byte aligned for 128-bit vector access; the (w+38)x(h+38) "padded level"
//              of upstream ComputePyramid is the view starting at byte 13.
//   blur[l]  : B x h_l rows x pitch_l bytes, ROI pixel 0 at byte 0 (7x7 Gaussian of the ROI).
//   cand[l]  : B x cand_cap_l packed u32 (x_rel:12 | y_rel:12 | m:8) written by the FAST kernel.
//   key/idx ping-pong + sd/head scratch for the quadtree kernel, sel[l] : B x sel_cap_l packed u32.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/orbb200.h"

#define ORBB_ROI_X0 32  // byte offset of ROI column 0 inside a padded row
#define ORBB_PAD_X0 13  // byte offset of padded column 0 (= ORBB_ROI_X0 - 19)
#define ORBB_BORDER 19
#define ORBB_MIN_BORDER 16  // upstream minBorderX/Y = EDGE_THRESHOLD - 3

namespace orbb {

// Device-visible description of one pyramid level (array of nlevels lives in HBM).
struct LevelDev {
    int w, h, pitch, rows;       // ROI size, bytes per row, rows per frame (h + 38)
    long long frame_stride;      // bytes between frames in img
    long long blur_stride;       // bytes between frames in blur
    uint8_t *img, *blur;
    // cv::resize tables for producing THIS level from level-1 (unused for level 0)
    const int *xofs;             // [w] source column
    const short2 *xalpha;        // [w] (a0,a1), 11-bit fixed point
    const int2 *yrows;           // [h] (clipped row0,row1)
    const short2 *ybeta;         // [h] (b0,b1)
    int src_w, src_h, area2x;
    // table-driven resize (k_resize_rows): per 4-byte group of the padded output row and per padded output row
    const uint4 *rs_h;           // [2*rs_nq]: {byte offset A, PRMT selector A, byte offset B, selector B}, {4 x (a0 | a1<<16)}
    const int4 *rs_v;            // [h+38]: {src row0, src row1, b0<<16, b1<<16} (reflect-101 already applied)
    int rs_nq, rs_ok;            // groups per row (first group = padded byte 12); 0 => fall back to k_resize
    // per-cell FAST grid (tested range starts at ROI (19,19))
    int w_cell, h_cell, n_cell_x, n_cell_y;
    // quadtree
    int nfeat, n_ini, depth, key_bits;
    const uint32_t *xkey, *ykey; // [w-32], [h-32]: Morton-spread path bits (+root) per coordinate
    const uint16_t *xord, *yord; // [w-32], [h-32]: (cell index << 6 | offset in cell)
    int cand_cap, sel_cap;
    uint32_t *cand;              // B x cand_cap
    uint2 *kv_a, *kv_b;          // B x cand_cap each: (path key, candidate index) pairs, radix ping-pong
    uint8_t *sd;                 // B x cand_cap: split depth of adjacent sorted keys
    uint32_t *sel;               // B x sel_cap packed selected keys
    // quadtree cell table (k_octree fast path), filled by the FAST kernel while it emits the candidates and cleared
    // by the quadtree kernel: per depth-tbl_dc tree cell the number of candidates and the best one
    int tbl_dc, tbl_cells;       // cells = 2^(root_bits + 2*tbl_dc); 0 = no table for this level
    uint32_t *tbl_cnt;           // B x tbl_cells
    unsigned long long *tbl_best;  // B x tbl_cells: (response << 56) | ((0x3ffffff - upstream order) << 24) | (x | y << 12)
    float scale, patch_size;     // mvScaleFactor[l], (float)(int)(31*scale)
};

// k_rgbd.cu, used by orbb_stage.cu: a lone frame's depth gate + 3-D lift + windowed match in one launch, and the pair compaction
cudaError_t launch_gate_match_lone(const uint32_t *d_aligned, const orbb_intrinsics &in, const orbb_keypoint *kp_raw,
                                   const uint8_t *desc_raw, const int *count_raw, int max_kp, orbb_keypoint *kp_out,
                                   uint8_t *desc_out, double *points, int *valid_out, int *count_blk, const uint8_t *q_desc,
                                   const float *q_pos, const int *q_count, float max_px, int max_hamming, int *out_idx,
                                   int *out_dist, cudaStream_t st);
cudaError_t launch_compact_pairs(const int *idx, const int *q_counts, int n_frames, int max_kp, const double *q_points,
                                 const double *t_points, const void *t_xy, int t_stride, double *prev_out,
                                 double *curr_out, uint16_t *xy_out, int *n_matched, cudaStream_t st);
// orbb_api.cu, used by orbb_stage.cu: see the definition
void note_replay(orbb_handle *h, int n_frames, long long launches);
// k_rgbd.cu, used by orbb_stage.cu
cudaError_t launch_stage_carry(orbb_keypoint *kp, uint8_t *desc, double *pts, int *valid, int carry, int max_kp, int *counts_dst,
                               const int *counts_src, int n_counts, cudaStream_t st);

#ifdef __CUDACC__
// Programmatic dependent launch (PDL): a kernel launched with launch_pdl() may become resident while its predecessor
// in the stream is still running; it must call pdl_wait() before it touches anything the predecessor writes
// (griddepcontrol.wait returns once the predecessor grid has completed and its writes are visible).  A kernel
// calls pdl_trigger() as early as it likes to allow ITS successor to be scheduled.  Both are no-ops in a normal
// launch.  Used for the level0 -> resize x7 -> FAST chain, where a lone frame's kernels are shorter than a launch gap.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// Adds one packed candidate (x:12 | y:12 | m:8, coordinates relative to the 16-px border) to its tree cell.
__device__ __forceinline__ void oct_bin_candidate(const LevelDev &L, int frame, uint32_t c) {
    const uint32_t kx = __ldg(&L.xkey[c & 0xfffu]), ky = __ldg(&L.ykey[(c >> 12) & 0xfffu]);
    const uint32_t xo = __ldg(&L.xord[c & 0xfffu]), yo = __ldg(&L.yord[(c >> 12) & 0xfffu]);
    const uint32_t cell = (kx | ky) >> (2 * (L.depth - L.tbl_dc));
    // upstream candidate order = (cell row, cell col, y in cell, x in cell); unique per pixel
    const uint32_t ord = ((yo >> 6) << 19) | ((xo >> 6) << 12) | ((yo & 63u) << 6) | (xo & 63u);
    const unsigned long long v = ((unsigned long long)(c >> 24) << 56) | ((unsigned long long)(0x3ffffffu - ord) << 24) |
                                 (c & 0xffffffu);
    const size_t at = (size_t)frame * L.tbl_cells + cell;
    atomicAdd(&L.tbl_cnt[at], 1u);
    atomicMax(&L.tbl_best[at], v);
}
#endif

// k_pyramid_fused (several pyramid levels per launch, small batches): rectangles of one tile, per level of its group
// (k = 0 .. n-1).  x coordinates are in "c" units (c = padded column + 1: c = 0 is byte 12 of a padded row, aligned
// 4-byte words are c = 4i .. 4i+3) and multiples of 4; y in padded rows.
#define ORBB_PF_GROUP 4
struct PfTile {
    short sx0, sy0, sw, sh;  // source window: image ROI coordinates (group 0) or c / padded-row units of the level below
    short nx0[ORBB_PF_GROUP], ny0[ORBB_PF_GROUP], nw[ORBB_PF_GROUP], nh[ORBB_PF_GROUP];  // pixels to COMPUTE
    short ox0[ORBB_PF_GROUP], oy0[ORBB_PF_GROUP], ow[ORBB_PF_GROUP], oh[ORBB_PF_GROUP];  // pixels to WRITE (partition)
};

struct CellEntry {  // one FAST work item = one upstream 30-px cell
    int16_t level, x0, y0, cw, ch, pad0;  // tested-range origin (ROI coords) and size
    uint32_t inv_nux;                      // (1 << 20) / ceil(cw / 4) + 1: floor(u / units per row) by multiply + shift
};

struct FastSmemCfg {
    int tile_pitch, tile_rows;    // bytes, rows of the per-warp image tile (manual staging)
    int tma_pitch;                // row bytes of the TMA box (16-byte aligned start => up to 15 bytes of phase)
    int score_pitch, score_rows;  // per-warp score tile (1-px zero frame); pitch == tile pitch (one offset addresses both)
    int tile_bytes, score_bytes;  // sizes of the two tiles inside the warp's block, multiples of 16
    int queue_len;                // u16 entries
    int warp_bytes;               // total per warp (multiple of 16)
};

}  // namespace orbb
