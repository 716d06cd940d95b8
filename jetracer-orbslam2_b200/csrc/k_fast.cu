// k_fast.cu -- FAST-9 detection with the per-cell 20/7 threshold fallback and 3x3 score NMS
// (replaces Jetracer::fast_gpu_calc_corner_response + grid_nms, reference src/cuda/fast.cu:150-372
// and src/cuda/nms.cu:86-296; semantics = upstream ComputeKeyPointsOctTree's per-cell cv::FAST calls,
// SURVEY.md A.3).
//
// Work item = one upstream 30-px cell, processed by ONE WARP with warp-synchronous code only (no
// block barriers): the cell's (cw+6)x(ch+6) window is staged in shared memory with aligned 32-bit
// loads, then
//   phase 1  packed antipodal-pair precheck (4 pixels per lane, VABSDIFF4 + SWAR compare), ballot-compacted
//            into a queue of pixel indices,
//   phase 2  exact threshold-free arc score m = max over 9-arcs of min|d|: both polarities packed in
//            s16x2 (one IMAD per ring pixel), 3-input min/max network (VIMNMX3.U16x2); corners (m > thr)
//            written to a score tile and re-compacted,
//   phase 3  3x3 strict NMS inside the cell (the score tile has a zero frame: neighbours outside the
//            cell's tested range count as 0, exactly like cv::FAST on the cell window),
//   phase 4  survivors are appended to the (frame, level) candidate list, one atomicAdd per 32.
// Phases 1-3 run at the high threshold first and are repeated at the low threshold only when the
// cell came out empty -- upstream's FAST(iniTh) -> FAST(minTh) fallback, decided after NMS.
// No score map ever goes to HBM (the DUMP instantiation exists only for the parity tests).
// Bound: SM issue slots / shared-memory bandwidth, not HBM.
#include <cuda.h>

#include <cstdlib>

#include "orbb_internal.cuh"

namespace orbb {

#define FAST_WARPS 4

// one tensor map per pyramid level (array in device global memory): 3-D u8 tensor (row bytes, rows, frames),
// box = one cell window
struct TmaMaps { CUtensorMap m[ORBB_MAX_LEVELS]; };

__device__ __forceinline__ int min3(int a, int b, int c) { return __vimin3_s32(a, b, c); }
__device__ __forceinline__ int max3(int a, int b, int c) { return __vimax3_s32(a, b, c); }

// m = max(A, B), A = max over the 16 9-arcs of min(v - ring), B = same for (ring - v).
// Both chains run at once in packed s16x2: P_k = (v - r_k + 256) | ((r_k - v + 256) << 16), which is a single
// IMAD per ring pixel, P_k = base + r_k * 0xFFFF with base = (v + 256) | ((256 - v) << 16): both halves stay in
// [1, 511], so no borrow crosses the halves.  Then min over 9 = two layers of 3-input min (VIMNMX3.U16x2) and
// the max over the 16 arcs is a 3-input max tree: 40 min/max instructions for both polarities.
__device__ __forceinline__ int arc_score(const uint8_t *p, int tp) {
    const unsigned v = p[0];
    const unsigned base = (v + 256u) | ((256u - v) << 16);
    unsigned d[16];
#define RING(k, off) d[k] = base + (unsigned)p[off] * 0xFFFFu
    RING(0, 3 * tp);       RING(1, 3 * tp + 1);   RING(2, 2 * tp + 2);    RING(3, tp + 3);
    RING(4, 3);            RING(5, -tp + 3);      RING(6, -2 * tp + 2);   RING(7, -3 * tp + 1);
    RING(8, -3 * tp);      RING(9, -3 * tp - 1);  RING(10, -2 * tp - 2);  RING(11, -tp - 3);
    RING(12, -3);          RING(13, tp - 3);      RING(14, 2 * tp - 2);   RING(15, 3 * tp - 1);
#undef RING
    unsigned lo3[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) lo3[k] = __vimin3_u16x2(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
    unsigned best = 0;
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
        const unsigned l0 = __vimin3_u16x2(lo3[k], lo3[(k + 3) & 15], lo3[(k + 6) & 15]);
        const unsigned l1 = __vimin3_u16x2(lo3[k + 1], lo3[(k + 4) & 15], lo3[(k + 7) & 15]);
        best = __vimax3_u16x2(best, l0, l1);
    }
    return (int)max(best & 0xffffu, best >> 16) - 256;
}

// TP: the tile's row pitch in bytes as a compile-time constant (0 = read it from cfg): with it the 16 ring offsets of
// the arc score and the +-3-row offsets of the precheck are instruction immediates off ONE base register, which takes
// about a dozen address instructions out of every scored pixel.  The score tile has the same pitch.
template <bool DUMP, bool TMA, int TP>
__global__ void __launch_bounds__(FAST_WARPS * 32, 9)
k_fast_cells(const CUtensorMap *__restrict__ maps, const LevelDev *__restrict__ levels, const CellEntry *__restrict__ cells, int n_cells, int n_levels,
             int *__restrict__ cand_count, int t_lo, int t_hi, FastSmemCfg cfg, int frame_base,
             uint8_t *__restrict__ dump, const long long *__restrict__ dump_off) {
    extern __shared__ __align__(128) uint8_t smem[];  // 128-byte aligned: TMA destination
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cell_id = blockIdx.x * FAST_WARPS + warp;
    if (cell_id >= n_cells) return;
    const int frame = blockIdx.y + frame_base;
    const CellEntry c = cells[cell_id];
    const LevelDev &L = levels[c.level];

    uint8_t *tile = smem + (size_t)warp * cfg.warp_bytes;
    uint8_t *score = tile + cfg.tile_bytes;  // 16-byte aligned
    uint16_t *queue = reinterpret_cast<uint16_t *>(score + cfg.score_bytes);
    const int tp = TMA ? cfg.tma_pitch : (TP ? TP : cfg.tile_pitch), sp = tp, tpw = tp >> 2;  // score tile: same pitch as the window tile
    const int cw = c.cw, ch = c.ch;
    const uint32_t *tile32 = reinterpret_cast<const uint32_t *>(tile);
    pdl_wait();  // launched as a programmatic dependent of the last pyramid kernel: the levels are complete from here on

    // ---- stage the window, RE-ALIGNED: shared-memory byte column k <-> image column x0 - 4 + k, so the
    // cell's first tested pixel sits at byte 4 of each row and groups of 4 pixels are whole 32-bit words.
    // Global loads stay aligned (ROI rows are 16-byte aligned); a funnel shift moves the bytes into place.
    // tested pixel x lives at tile byte column x + off.  Manual staging re-aligns (off = 4); the TMA box must
    // start on a 16-byte boundary, so its rows keep a byte phase of (x0 - 4) & 15 that phase 1 absorbs.  (Measured
    // in round 2: a box whose first coordinate is NOT a multiple of 16 bytes -- which would let TMA do the
    // re-alignment -- makes the copy fault with "an illegal instruction was encountered".)
    const int off = TMA ? 4 + ((c.x0 - 4) & 15) : 4;
    if (TMA) {
        // One bulk-tensor copy per cell: the TMA unit fetches the (box_w x box_h) window at byte column
        // ROI_X0 + x0 - 4 of padded row BORDER + y0 - 3 straight into shared memory (no registers, no per-lane
        // address math, arbitrary byte phase), and signals the warp's mbarrier with the byte count.
        __shared__ __align__(8) unsigned long long s_bar[FAST_WARPS];
        const unsigned bar = (unsigned)__cvta_generic_to_shared(&s_bar[warp]);
        const unsigned dst = (unsigned)__cvta_generic_to_shared(tile);
        if (lane == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            // the tensor maps live in global memory and were written by the host (cudaMemcpy): acquire them for
            // the tensormap proxy before the first use in this warp
            asm volatile("fence.proxy.tensormap::generic.acquire.sys [%0], 128;" ::"l"(maps + c.level) : "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(tp * cfg.tile_rows) : "memory");
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                ::"r"(dst), "l"(maps + c.level), "r"((ORBB_ROI_X0 + c.x0 - 4) & ~15), "r"(ORBB_BORDER + c.y0 - 3), "r"(frame), "r"(bar)
                : "memory");
        }
        for (int i = lane; i < cfg.score_bytes >> 4; i += 32) reinterpret_cast<uint4 *>(score)[i] = make_uint4(0u, 0u, 0u, 0u);
        __syncwarp();
        unsigned done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(bar) : "memory");
    } else {
        // 128-bit staging WITH re-alignment.  Each lane loads ONE aligned 16-byte chunk of a window row (ROI rows are
        // 128-byte aligned by layout) -- 4 lanes per row, 8 when a row spans more than four chunks -- and every row of
        // the lane is in flight before the first one is used (one exposed L2 round trip per cell).  The lane shifts its
        // own four words right by the byte phase of the window (the word after the chunk comes from the next lane: one
        // shuffle) and stores them at the tile position its chunk lands on, word 4 * chunk - (word offset of the
        // window inside its first chunk): the re-alignment by whole words is an address, not a data movement (round 1
        // loaded both chunks per lane, the first round-2 form selected the four output words with a uniform switch after
        // four shuffles: 14 % of the kernel's instructions, profiles/r02b_k_fast_cells_phases.txt).  32-bit stores:
        // the tile keeps an odd word pitch, so the byte loads of the arc score (pixels of many rows in one warp
        // instruction) spread over all banks; a 16-byte pitch would fold rows 8 apart onto the same banks.
        const int gx = c.x0 - 4, xa16 = gx & ~15, wo = (gx - xa16) >> 2, sh = (gx & 3) * 8;
        const int nwords = (cw + 7 + 3) >> 2;               // window words per row
        const int nld = ((wo + nwords) >> 2) + 1;           // chunks per row that hold them (the funnel shift reads one word ahead)
        const int nrows = ch + 6;
        const int lps = nld <= 4 ? 2 : 3, rpi = 32 >> lps;  // log2(lanes per row), rows per step
        const int g = lane & ((1 << lps) - 1), rl = lane >> lps;
        // a chunk that would reach past the padded row is never part of a window (those lanes re-read chunk 0 and store nothing)
        const bool in_row = 16 * g + 16 <= L.pitch - (ORBB_ROI_X0 + xa16);
        const uint8_t *src = L.img + (size_t)frame * L.frame_stride + (size_t)(ORBB_BORDER + c.y0 - 3) * L.pitch + ORBB_ROI_X0 + xa16 + (in_row ? 16 * g : 0);
        asm volatile("" : "+l"(src));                       // one pointer register pair, not re-derived per load
        const unsigned pitch = (unsigned)L.pitch;
        const int k0 = 4 * g - wo;                          // tile word of the lane's first output
        unsigned keep = 0;                                  // which of the four outputs land inside the tile row
#pragma unroll
        for (int i = 0; i < 4; ++i) keep |= (in_row && g < nld && k0 + i >= 0 && k0 + i < tpw) ? 1u << i : 0u;
        uint32_t *t0 = reinterpret_cast<uint32_t *>(tile) + k0;
        constexpr int NF = 5;                               // rows in flight per lane: a 38-row window at 8 rows per step is one round
        for (int rb = 0; rb < nrows; rb += NF * rpi) {      // rb: first row of the round, warp-uniform (the shuffles need every lane)
            const int r0 = rb + rl;
            uint4 a[NF];
#pragma unroll
            for (int u = 0; u < NF; ++u)                    // rows past the window re-read its last row
                a[u] = __ldg(reinterpret_cast<const uint4 *>(src + (unsigned)min(r0 + u * rpi, nrows - 1) * pitch));
#pragma unroll
            for (int u = 0; u < NF; ++u) {
                const unsigned nx = __shfl_down_sync(0xffffffffu, a[u].x, 1);  // first word of the chunk to the right
                const int r = r0 + u * rpi;
                const unsigned m = r < nrows ? keep : 0u;
                uint32_t *t = t0 + r * tpw;
                if (m & 1u) t[0] = __funnelshift_r(a[u].x, a[u].y, sh);
                if (m & 2u) t[1] = __funnelshift_r(a[u].y, a[u].z, sh);
                if (m & 4u) t[2] = __funnelshift_r(a[u].z, a[u].w, sh);
                if (m & 8u) t[3] = __funnelshift_r(a[u].w, nx, sh);
            }
        }
        for (int i = lane; i < cfg.score_bytes >> 4; i += 32) reinterpret_cast<uint4 *>(score)[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncwarp();

    const int npix = cw * ch;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int nux = (cw + 3) >> 2, nunits = nux * ch;        // 4-pixel units per row / per cell
    const unsigned inv_nux = c.inv_nux;  // (1 << 20) / nux + 1, tabulated by the host: exact floor(u / nux) for u * nux < 2^20
    // Queue entries are y * tp + x, cell-relative: the byte offset of the pixel inside the window tile AND inside the score
    // tile (same pitch), so phases 2 and 3 add them to one base pointer; only the emitted survivors are decoded back to (y, x).

    // Upstream order: FAST(ini) on the cell; only if that leaves nothing, FAST(min).  Running the high
    // threshold first rejects most pixels in the precheck (the low-threshold pass is rare on textured input).
    int kn = 0;
    for (int pass = DUMP ? 1 : 0; pass < 2; ++pass) {
        const int thr = pass == 0 ? t_hi : t_lo;
        if (pass == 1 && !DUMP && t_lo == t_hi) break;
        // ---- phase 1: packed precheck, 4 pixels per lane and step.  Necessary condition for a 9-arc: every antipodal
        // pair holds a pixel with |d| > thr; tested on the pairs (0,8) and (4,12): VABSDIFF4, then per byte
        // "a >= K" as bit 7 of (a + (128 - K)) | a  (K = min(thr + 1, 128); a carry out of a byte can only add 1 to
        // its neighbour, i.e. let a few more pixels through -- exactness comes from phase 2).
        //
        // Lane = tile row, step = unit column: every lane walks along ITS row, keeps the flags of the row in a 64-bit
        // register (4 bits per unit) and slides the left / centre / right words through registers (3 loads per unit;
        // with the odd word pitch of the tile the 32 rows of a step sit in 32 different banks: no replays).  The queue
        // positions then come from ONE warp prefix sum of the per-lane counts -- not from four ballots per 32 units:
        // on B200 VOTE and POPC issue at 16 lanes/clk/SM, a quarter of the ALU rate (tools/pipe_probe.cu,
        // profiles/r02_pipe_probe.txt), and the ballot form spent 12 of them per 128 pixels.  Rows beyond the 32nd
        // (cells up to 63 px high) go through a second round in linear unit order.  The queue order (row, unit, pixel
        // for round A) is raster order; nothing downstream depends on it.
        const unsigned cadd = (unsigned)(128 - min(thr + 1, 128)) * 0x01010101u;
        const unsigned last_mask = (cw & 3) ? (0x80808080u >> (8 * (4 - (cw & 3)))) : 0x80808080u;  // ragged last unit of a row
        // precheck of one unit from its five words: bit 7 of byte b set <=> pixel b passes
        auto precheck = [&](unsigned cc, unsigned up, unsigned dn, unsigned lf, unsigned rt) -> unsigned {
            const unsigned a0 = __vabsdiffu4(dn, cc), a8 = __vabsdiffu4(up, cc);
            const unsigned a4 = __vabsdiffu4(rt, cc), a12 = __vabsdiffu4(lf, cc);
            return ((a0 + cadd) | a0 | (a8 + cadd) | a8) & ((a4 + cadd) | a4 | (a12 + cadd) | a12);
        };
        int qn = 0;
        const int rows_a = min(ch, 32), units_a = rows_a * nux;
        for (int round = 0; round < 2; ++round) {
            const int U = round == 0 ? nux : (nunits - units_a + 31) >> 5;  // steps of this round (<= 16: 64 flag bits)
            if (U <= 0) break;
            unsigned wlo = 0, whi = 0;
            if (round == 0 && !TMA) {
                // rows past the cell repeat row 0 (same addresses as lane 0: a broadcast, not a bank conflict)
                const uint32_t *row = tile32 + ((lane < rows_a ? lane : 0) + 3) * tpw;  // row[0] = left word of unit 0
                const unsigned live = lane < rows_a ? 0x80808080u : 0u;
                unsigned left = row[0], cc = row[1];
                // rows +-2 slide the same way: they feed the two diagonal pairs (2,10) and (6,14)
                unsigned lp = row[2 * tpw], cp = row[1 + 2 * tpw], lm = row[-2 * tpw], cm = row[1 - 2 * tpw];
                for (int j = 0; j < U; ++j) {
                    const unsigned right = row[j + 2], dn = row[j + 1 + 3 * tpw], up = row[j + 1 - 3 * tpw];
                    const unsigned rp = row[j + 2 + 2 * tpw], rm = row[j + 2 - 2 * tpw];
                    const unsigned rt = __funnelshift_r(cc, right, 24), lf = __funnelshift_r(left, cc, 8);
                    unsigned flags = precheck(cc, up, dn, lf, rt) & (j == U - 1 ? last_mask & live : live);
                    flags &= precheck(cc, __funnelshift_r(lm, cm, 16), __funnelshift_r(cp, rp, 16),   // (-2,-2) | (+2,+2)
                                      __funnelshift_r(lp, cp, 16), __funnelshift_r(cm, rm, 16));      // (-2,+2) | (+2,-2)
                    left = cc; cc = right; lp = cp; cp = rp; lm = cm; cm = rm;
                    // bits 7 / 15 / 23 / 31 -> one nibble pushed into the TOP of the lane's 64-bit flag word: the high half of
                    // the product holds the four bits at 0..3 (no two partial products meet at or below them; what lies
                    // above is dropped by the funnel shift), two funnel shifts push them in.  After U steps pixel x of
                    // the row sits at bit x + 64 - 4 * U.
                    const unsigned nib = __umulhi(flags, 0x02040810u);
                    wlo = __funnelshift_r(wlo, whi, 4);
                    whi = __funnelshift_r(whi, nib, 4);
                }
            } else {
                // generic form (second round, TMA staging): unit = first + step * 32 + lane, index math per step
                const int first = round == 0 ? 0 : units_a, total = round == 0 ? units_a : nunits;
                for (int i = 0; i < U; ++i) {
                    const int ur = round == 0 ? lane * nux + i : first + i * 32 + lane;
                    const bool valid = round == 0 ? lane < rows_a : ur < total;
                    const int u = valid ? ur : first + i;
                    const int y = (int)(((unsigned)u * inv_nux) >> 20), j = u - y * nux;
                    unsigned cc, up, dn, lf, rt;
                    if (TMA) {  // arbitrary byte phase: every 4-pixel window is a funnel shift of two words
                        const int bc = off + 4 * j;  // byte column of the unit's first pixel
                        const uint32_t *r0 = tile32 + (y + 3) * tpw;
                        const int wc = bc >> 2, sc = (bc & 3) * 8;
                        cc = __funnelshift_r(r0[wc], r0[wc + 1], sc);
                        up = __funnelshift_r(r0[wc - 3 * tpw], r0[wc + 1 - 3 * tpw], sc);
                        dn = __funnelshift_r(r0[wc + 3 * tpw], r0[wc + 1 + 3 * tpw], sc);
                        const int bl = bc - 3, br = bc + 3;
                        lf = __funnelshift_r(r0[bl >> 2], r0[(bl >> 2) + 1], (bl & 3) * 8);
                        rt = __funnelshift_r(r0[br >> 2], r0[(br >> 2) + 1], (br & 3) * 8);
                    } else {
                        const uint32_t *row = tile32 + (y + 3) * tpw + j;  // row[0] = left word, row[1] = centre, row[2] = right
                        cc = row[1];
                        dn = row[1 + 3 * tpw]; up = row[1 - 3 * tpw];
                        rt = __funnelshift_r(cc, row[2], 24);
                        lf = __funnelshift_r(row[0], cc, 8);
                    }
                    const unsigned vm = valid ? (j == nux - 1 ? last_mask : 0x80808080u) : 0u;
                    const unsigned nib = __umulhi(precheck(cc, up, dn, lf, rt) & vm, 0x02040810u);
                    wlo = __funnelshift_r(wlo, whi, 4);
                    whi = __funnelshift_r(whi, nib, 4);
                }
            }
            // queue positions: exclusive prefix sum of the lanes' counts (5 shuffles per round)
            const int c_own = __popc(wlo) + __popc(whi);
            int incl = c_own;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            uint16_t *qp = queue + qn + incl - c_own;
            qn += __shfl_sync(0xffffffffu, incl, 31);
            // walk the SET bits only (most significant first) and store the flagged pixels as y * tp + x: a lane spends
            // one short iteration per flagged pixel of its row (~6 of 30 on the bench texture; the warp runs as long as its
            // fullest row) instead of a fixed 30 instructions per 4-pixel unit
            const int xoff = 4 * U - 64;  // bit position -> 4 * step + pixel
            const int row_e = lane * tp + xoff;
#pragma unroll
            for (int half = 1; half >= 0; --half) {
                unsigned w = half ? whi : wlo;
                while (w) {
                    const int b = 31 - __clz(w);
                    w &= ~(1u << b);
                    int e = row_e + 32 * half + b;
                    if (round != 0) {
                        const int pos = 32 * half + b + xoff, u = units_a + (pos >> 2) * 32 + lane;
                        const int y = (int)(((unsigned)u * inv_nux) >> 20);
                        e = y * (tp - 4 * nux) + 4 * u + (pos & 3);  // y * tp + 4 * j + pixel
                    }
                    *qp++ = (uint16_t)e;
                }
            }
        }
        __syncwarp();

        // ---- phase 2: exact arc score; corners (m > thr) go to the score tile and stay queued
        int cn = 0;
        for (int base = 0; base < qn; base += 32) {
            const int i = base + lane;
            int idx = 0, m = 0;
            if (i < qn) {
                idx = queue[i];
                m = arc_score(tile + 3 * tp + off + idx, tp);
                m = m > thr ? m : 0;
                if (m) score[sp + 1 + idx] = (uint8_t)min(m, 255);
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
                const uint8_t *s = score + sp + 1 + idx;
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
            const int y = idx / tp, x = idx - y * tp;  // compile-time pitch: a multiply-high
            const int m = score[sp + 1 + idx];
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
                if (dst < L.cand_cap) {
                    cand[dst] = packed;
                    // the quadtree kernel's cell table is filled here: fire-and-forget L2 atomics spread over every SM
                    // instead of shared-memory atomics inside the one CTA that owns the (frame, level)
                    if (L.tbl_cells) oct_bin_candidate(L, frame, packed);
                }
            }
        }
    }
}

// ---- Small grids (a lone frame, the reference's operating mode): ONE CTA OF FOUR WARPS PER CELL.  With fewer cells
// than the GPU has warp slots, the warp-per-cell kernel above is a chain of ~2 000 dependent warp instructions per cell
// (10-11 us for one 848x480 level on an otherwise idle GPU).  Here the four warps of a CTA share the cell: the window rows
// of the staging, the unit columns of the precheck and the entries of the arc-score / NMS / emit queues are dealt out
// over 128 threads, with one block barrier between phases and shared-memory counters for the queue lengths.  Same
// tiles, same arithmetic, same per-cell threshold decision; queue order differs, which nothing downstream depends on
// (candidates reach their list through atomics in the warp-per-cell kernel too).  Corners and NMS survivors ping-pong
// between two queues (in place, one warp's compaction would overwrite entries another warp has not read yet).
template <int TP>
__global__ void __launch_bounds__(128)
k_fast_cell_cta(const LevelDev *__restrict__ levels, const CellEntry *__restrict__ cells, int n_cells, int n_levels,
                int *__restrict__ cand_count, int t_lo, int t_hi, FastSmemCfg cfg, int frame_base) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ int s_cnt[2][3];  // per pass: queued pixels, corners, NMS survivors
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cell_id = blockIdx.x;
    const int frame = blockIdx.y + frame_base;
    const CellEntry c = cells[cell_id];
    const LevelDev &L = levels[c.level];
    uint8_t *tile = smem;
    uint8_t *score = tile + cfg.tile_bytes;
    uint16_t *queue = reinterpret_cast<uint16_t *>(score + cfg.score_bytes);
    uint16_t *queue2 = queue + cfg.queue_len;
    constexpr int tp = TP, sp = TP, tpw = TP >> 2;
    const int cw = c.cw, ch = c.ch;
    const uint32_t *tile32 = reinterpret_cast<const uint32_t *>(tile);
    if (tid < 6) (&s_cnt[0][0])[tid] = 0;
    pdl_wait();

    {   // ---- staging (see k_fast_cells): the rows of a step are dealt out over the four warps
        const int gx = c.x0 - 4, xa16 = gx & ~15, wo = (gx - xa16) >> 2, sh = (gx & 3) * 8;
        const int nwords = (cw + 7 + 3) >> 2;
        const int nld = ((wo + nwords) >> 2) + 1;
        const int nrows = ch + 6;
        const int lps = nld <= 4 ? 2 : 3, rpi = 32 >> lps;
        const int g = lane & ((1 << lps) - 1), rl = lane >> lps;
        const bool in_row = 16 * g + 16 <= L.pitch - (ORBB_ROI_X0 + xa16);
        const uint8_t *src = L.img + (size_t)frame * L.frame_stride + (size_t)(ORBB_BORDER + c.y0 - 3) * L.pitch + ORBB_ROI_X0 + xa16 + (in_row ? 16 * g : 0);
        asm volatile("" : "+l"(src));
        const unsigned pitch = (unsigned)L.pitch;
        const int k0 = 4 * g - wo;
        unsigned keep = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) keep |= (in_row && g < nld && k0 + i >= 0 && k0 + i < tpw) ? 1u << i : 0u;
        uint32_t *t0 = reinterpret_cast<uint32_t *>(tile) + k0;
        constexpr int NF = 3;  // rows in flight per lane: 4 warps x 8 (4) rows x 3 = 96 (48) rows per round
        for (int rb = 0; rb < nrows; rb += NF * 4 * rpi) {
            const int r0 = rb + warp * rpi + rl;
            uint4 a[NF];
#pragma unroll
            for (int u = 0; u < NF; ++u)
                a[u] = __ldg(reinterpret_cast<const uint4 *>(src + (unsigned)min(r0 + u * 4 * rpi, nrows - 1) * pitch));
#pragma unroll
            for (int u = 0; u < NF; ++u) {
                const unsigned nx = __shfl_down_sync(0xffffffffu, a[u].x, 1);
                const int r = r0 + u * 4 * rpi;
                const unsigned m = r < nrows ? keep : 0u;
                uint32_t *t = t0 + r * tpw;
                if (m & 1u) t[0] = __funnelshift_r(a[u].x, a[u].y, sh);
                if (m & 2u) t[1] = __funnelshift_r(a[u].y, a[u].z, sh);
                if (m & 4u) t[2] = __funnelshift_r(a[u].z, a[u].w, sh);
                if (m & 8u) t[3] = __funnelshift_r(a[u].w, nx, sh);
            }
        }
        for (int i = tid; i < cfg.score_bytes >> 4; i += 128) reinterpret_cast<uint4 *>(score)[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();

    const unsigned lt_mask = (1u << lane) - 1u;
    const int nux = (cw + 3) >> 2, nunits = nux * ch;
    const unsigned inv_nux = c.inv_nux;
    int kn = 0;
    uint16_t *qsurv = queue;
    for (int pass = 0; pass < 2; ++pass) {
        const int thr = pass == 0 ? t_hi : t_lo;
        if (pass == 1 && t_lo == t_hi) break;
        int *cnt = s_cnt[pass];
        // ---- phase 1: precheck (k_fast_cells: lane = tile row, sliding word windows); warp w takes the unit columns
        // [w * Uw, (w + 1) * Uw) of round A and the steps w, w + 4, ... of round B
        const unsigned cadd = (unsigned)(128 - min(thr + 1, 128)) * 0x01010101u;
        const unsigned last_mask = (cw & 3) ? (0x80808080u >> (8 * (4 - (cw & 3)))) : 0x80808080u;
        auto precheck = [&](unsigned cc, unsigned up, unsigned dn, unsigned lf, unsigned rt) -> unsigned {
            const unsigned a0 = __vabsdiffu4(dn, cc), a8 = __vabsdiffu4(up, cc);
            const unsigned a4 = __vabsdiffu4(rt, cc), a12 = __vabsdiffu4(lf, cc);
            return ((a0 + cadd) | a0 | (a8 + cadd) | a8) & ((a4 + cadd) | a4 | (a12 + cadd) | a12);
        };
        const int rows_a = min(ch, 32), units_a = rows_a * nux;
        for (int round = 0; round < 2; ++round) {
            int nst;  // steps of this warp in this round (<= 16: 64 flag bits)
            int jb = 0;
            unsigned wlo = 0, whi = 0;
            if (round == 0) {
                const int Uw = (nux + 3) >> 2;
                jb = warp * Uw;
                nst = max(0, min(nux, jb + Uw) - jb);
                if (nst > 0) {
                    const uint32_t *row = tile32 + ((lane < rows_a ? lane : 0) + 3) * tpw + jb;  // row[0] = left word of unit jb
                    const unsigned live = lane < rows_a ? 0x80808080u : 0u;
                    unsigned left = row[0], cc = row[1];
                    unsigned lp = row[2 * tpw], cp = row[1 + 2 * tpw], lm = row[-2 * tpw], cm = row[1 - 2 * tpw];
                    for (int j = 0; j < nst; ++j) {
                        const unsigned right = row[j + 2], dn = row[j + 1 + 3 * tpw], up = row[j + 1 - 3 * tpw];
                        const unsigned rp = row[j + 2 + 2 * tpw], rm = row[j + 2 - 2 * tpw];
                        const unsigned rt = __funnelshift_r(cc, right, 24), lf = __funnelshift_r(left, cc, 8);
                        unsigned flags = precheck(cc, up, dn, lf, rt) & (jb + j == nux - 1 ? last_mask & live : live);
                        flags &= precheck(cc, __funnelshift_r(lm, cm, 16), __funnelshift_r(cp, rp, 16),
                                          __funnelshift_r(lp, cp, 16), __funnelshift_r(cm, rm, 16));
                        left = cc; cc = right; lp = cp; cp = rp; lm = cm; cm = rm;
                        const unsigned nib = __umulhi(flags, 0x02040810u);
                        wlo = __funnelshift_r(wlo, whi, 4);
                        whi = __funnelshift_r(whi, nib, 4);
                    }
                }
            } else {
                const int U = (nunits - units_a + 31) >> 5;  // steps of round B in all
                nst = U > warp ? (U - warp + 3) >> 2 : 0;
                for (int s = 0; s < nst; ++s) {
                    const int ur = units_a + (warp + 4 * s) * 32 + lane;
                    const bool valid = ur < nunits;
                    const int u = valid ? ur : units_a;
                    const int y = (int)(((unsigned)u * inv_nux) >> 20), j = u - y * nux;
                    const uint32_t *row = tile32 + (y + 3) * tpw + j;
                    const unsigned cc = row[1], dn = row[1 + 3 * tpw], up = row[1 - 3 * tpw];
                    const unsigned rt = __funnelshift_r(cc, row[2], 24), lf = __funnelshift_r(row[0], cc, 8);
                    const unsigned vm = valid ? (j == nux - 1 ? last_mask : 0x80808080u) : 0u;
                    const unsigned nib = __umulhi(precheck(cc, up, dn, lf, rt) & vm, 0x02040810u);
                    wlo = __funnelshift_r(wlo, whi, 4);
                    whi = __funnelshift_r(whi, nib, 4);
                }
            }
            if (round == 1 && nunits <= units_a) break;  // block-uniform: no round B in this cell
            // queue positions: warp prefix sum of the lanes' counts + one shared-memory atomic per warp
            const int c_own = __popc(wlo) + __popc(whi);
            int incl = c_own;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            int wbase = 0;
            if (lane == 31 && incl) wbase = atomicAdd(&cnt[0], incl);
            wbase = __shfl_sync(0xffffffffu, wbase, 31);
            uint16_t *qp = queue + wbase + incl - c_own;
            const int xoff = 4 * nst - 64;  // bit position -> 4 * step + pixel
            const int row_e = lane * tp + xoff + 4 * jb;
#pragma unroll
            for (int half = 1; half >= 0; --half) {
                unsigned w = half ? whi : wlo;
                while (w) {
                    const int b = 31 - __clz(w);
                    w &= ~(1u << b);
                    int e = row_e + 32 * half + b;
                    if (round != 0) {
                        const int pos = 32 * half + b + xoff, u = units_a + (warp + 4 * (pos >> 2)) * 32 + lane;
                        const int y = (int)(((unsigned)u * inv_nux) >> 20);
                        e = y * (tp - 4 * nux) + 4 * u + (pos & 3);
                    }
                    *qp++ = (uint16_t)e;
                }
            }
        }
        __syncthreads();
        const int qn = cnt[0];

        // ---- phase 2: exact arc score; corners go to the score tile and to the second queue
        for (int base = 0; base < qn; base += 128) {  // block-uniform trip count
            const int i = base + tid;
            int idx = 0, m = 0;
            if (i < qn) {
                idx = queue[i];
                m = arc_score(tile + 3 * tp + 4 + idx, tp);
                m = m > thr ? m : 0;
                if (m) score[sp + 1 + idx] = (uint8_t)min(m, 255);
            }
            const unsigned bm = __ballot_sync(0xffffffffu, m != 0);
            int wb = 0;
            if (lane == 0 && bm) wb = atomicAdd(&cnt[1], __popc(bm));
            wb = __shfl_sync(0xffffffffu, wb, 0);
            if (m) queue2[wb + __popc(bm & lt_mask)] = (uint16_t)idx;
        }
        __syncthreads();
        const int cn = cnt[1];

        // ---- phase 3: strict 3x3 NMS inside the cell; survivors go back to the first queue
        for (int base = 0; base < cn; base += 128) {
            const int i = base + tid;
            int idx = 0;
            bool keep = false;
            if (i < cn) {
                idx = queue2[i];
                const uint8_t *s = score + sp + 1 + idx;
                const int v = s[0];
                const int n0 = max3(s[-sp - 1], s[-sp], s[-sp + 1]);
                const int n1 = max3(s[-1], s[1], s[sp - 1]);
                const int n2 = max3(s[sp], s[sp + 1], n0);
                keep = v > max(n1, n2);
            }
            const unsigned bm = __ballot_sync(0xffffffffu, keep);
            int wb = 0;
            if (lane == 0 && bm) wb = atomicAdd(&cnt[2], __popc(bm));
            wb = __shfl_sync(0xffffffffu, wb, 0);
            if (keep) queue[wb + __popc(bm & lt_mask)] = (uint16_t)idx;
        }
        __syncthreads();
        kn = cnt[2];
        if (kn > 0) break;  // cv::FAST(ini) found keypoints: no fallback for this cell
    }

    // ---- phase 4: append to the (frame, level) candidate list, one atomicAdd per warp and 32 survivors
    int *counter = cand_count + frame * n_levels + c.level;
    uint32_t *cand = L.cand + (size_t)frame * L.cand_cap;
    for (int base = 0; base < kn; base += 128) {
        const int i = base + tid;
        bool emit = false;
        uint32_t packed = 0;
        if (i < kn) {
            const int idx = qsurv[i];
            const int y = idx / tp, x = idx - y * tp;
            const int m = score[sp + 1 + idx];
            emit = true;
            packed = (uint32_t)(c.x0 + x - ORBB_MIN_BORDER) | ((uint32_t)(c.y0 + y - ORBB_MIN_BORDER) << 12) | ((uint32_t)m << 24);
        }
        const unsigned bm = __ballot_sync(0xffffffffu, emit);
        if (bm) {
            int slot = 0;
            if (lane == 0) slot = atomicAdd(counter, __popc(bm));
            slot = __shfl_sync(0xffffffffu, slot, 0);
            if (emit) {
                const int dst = slot + __popc(bm & lt_mask);
                if (dst < L.cand_cap) {
                    cand[dst] = packed;
                    if (L.tbl_cells) oct_bin_candidate(L, frame, packed);
                }
            }
        }
    }
}

template <int TP>
static cudaError_t launch_fast_cta_t(const LevelDev *d_levels, const CellEntry *d_cells, int n_cells, int n_levels,
                                     int *d_cand_count, int t_lo, int t_hi, const FastSmemCfg &cfg, int frame_base, int n_frames,
                                     cudaStream_t st) {
    const size_t smem = (size_t)cfg.tile_bytes + cfg.score_bytes + 4 * (size_t)cfg.queue_len;  // two queues
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_fast_cell_cta<TP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    return launch_pdl(k_fast_cell_cta<TP>, dim3(n_cells, n_frames), dim3(128), smem, st, d_levels, d_cells, n_cells, n_levels,
                      d_cand_count, t_lo, t_hi, cfg, frame_base);
}

template <bool DUMP, bool TMA, int TP>
static cudaError_t launch_fast_t(const CUtensorMap *maps, const LevelDev *d_levels, const CellEntry *d_cells, int n_cells,
                                 int n_levels, int *d_cand_count, int t_lo, int t_hi, const FastSmemCfg &cfg,
                                 int frame_base, int n_frames, uint8_t *d_dump, const long long *d_dump_off,
                                 cudaStream_t st) {
    dim3 grid((n_cells + FAST_WARPS - 1) / FAST_WARPS, n_frames);
    const size_t smem = (size_t)cfg.warp_bytes * FAST_WARPS;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_fast_cells<DUMP, TMA, TP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    return launch_pdl(k_fast_cells<DUMP, TMA, TP>, grid, dim3(FAST_WARPS * 32), smem, st, maps, d_levels, d_cells, n_cells, n_levels,
                      d_cand_count, t_lo, t_hi, cfg, frame_base, d_dump, d_dump_off);
}

cudaError_t launch_fast(const void *tma_maps, const LevelDev *d_levels, const CellEntry *d_cells, int n_cells,
                        int n_levels, int *d_cand_count, int t_lo, int t_hi, const FastSmemCfg &cfg, int frame_base,
                        int n_frames, cudaStream_t st) {
    if (tma_maps)
        return launch_fast_t<false, true, 0>(static_cast<const CUtensorMap *>(tma_maps), d_levels, d_cells, n_cells, n_levels,
                                             d_cand_count, t_lo, t_hi, cfg, frame_base, n_frames, nullptr, nullptr, st);
    // small grids: one CTA of four warps per cell while every cell still gets its own CTA slot (148 SMs x 8 CTAs);
    // ORBB_FAST_CTA=0 keeps the warp-per-cell kernel (diagnostics / A-B timing)
    static const bool cta_ok = !(getenv("ORBB_FAST_CTA") && atoi(getenv("ORBB_FAST_CTA")) == 0);
    if (cta_ok && (long long)n_cells * n_frames <= 148 * 8) {
#define ORBB_FAST_CTA_TP(P)                                                                                                 \
    case P: return launch_fast_cta_t<P>(d_levels, d_cells, n_cells, n_levels, d_cand_count, t_lo, t_hi, cfg, frame_base, n_frames, st)
        switch (cfg.tile_pitch) {
            ORBB_FAST_CTA_TP(44); ORBB_FAST_CTA_TP(52); ORBB_FAST_CTA_TP(60); ORBB_FAST_CTA_TP(68); ORBB_FAST_CTA_TP(76);
            default: break;
        }
#undef ORBB_FAST_CTA_TP
    }
#define ORBB_FAST_TP(P)                                                                                                     \
    case P: return launch_fast_t<false, false, P>(nullptr, d_levels, d_cells, n_cells, n_levels, d_cand_count, t_lo, t_hi, cfg, \
                                                  frame_base, n_frames, nullptr, nullptr, st)
    {
        switch (cfg.tile_pitch) {  // the pitches orbb_create produces: round_up(widest cell + 12, 4) | 4
            ORBB_FAST_TP(44); ORBB_FAST_TP(52); ORBB_FAST_TP(60); ORBB_FAST_TP(68); ORBB_FAST_TP(76);
            default: break;
        }
    }
#undef ORBB_FAST_TP
    return launch_fast_t<false, false, 0>(nullptr, d_levels, d_cells, n_cells, n_levels, d_cand_count, t_lo, t_hi, cfg,
                                          frame_base, n_frames, nullptr, nullptr, st);
}

// parity-test variant: one frame, dumps the per-pixel score (m > t_lo ? m : 0) of every cell
cudaError_t launch_fast_dump(const void *tma_maps, const LevelDev *d_levels, const CellEntry *d_cells, int n_cells,
                             int n_levels, int t_lo, int t_hi, const FastSmemCfg &cfg, int frame, uint8_t *d_dump,
                             const long long *d_dump_off, cudaStream_t st) {
    if (tma_maps)
        return launch_fast_t<true, true, 0>(static_cast<const CUtensorMap *>(tma_maps), d_levels, d_cells, n_cells, n_levels,
                                            nullptr, t_lo, t_hi, cfg, frame, 1, d_dump, d_dump_off, st);
    return launch_fast_t<true, false, 0>(nullptr, d_levels, d_cells, n_cells, n_levels, nullptr, t_lo, t_hi, cfg, frame, 1,
                                         d_dump, d_dump_off, st);
}

// Build the per-level tensor maps (driver entry point fetched through the runtime: no -lcuda needed).
// Returns false when the driver cannot encode them; the caller then uses the manual staging path.
bool build_fast_tma_maps(void *out_maps, const LevelDev *h_levels, int n_levels, int max_batch, const FastSmemCfg &cfg) {
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                 const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
        qres != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return false;
    }
    TmaMaps *maps = static_cast<TmaMaps *>(out_maps);
    for (int l = 0; l < n_levels; ++l) {
        const LevelDev &L = h_levels[l];
        const cuuint64_t dims[3] = {(cuuint64_t)L.pitch, (cuuint64_t)L.rows, (cuuint64_t)max_batch};
        const cuuint64_t strides[2] = {(cuuint64_t)L.pitch, (cuuint64_t)L.frame_stride};
        const cuuint32_t box[3] = {(cuuint32_t)cfg.tma_pitch, (cuuint32_t)cfg.tile_rows, 1u};
        const cuuint32_t estr[3] = {1u, 1u, 1u};
        const CUresult r = reinterpret_cast<EncodeFn>(fn)(&maps->m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, L.img, dims, strides,
                                                           box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return false;
    }
    return true;
}

size_t fast_tma_maps_bytes() { return sizeof(TmaMaps); }

}  // namespace orbb
