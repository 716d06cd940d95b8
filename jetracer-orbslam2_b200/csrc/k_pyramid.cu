// k_pyramid.cu -- ComputePyramid stage (replaces Jetracer::pyramid_create_levels,
// reference src/cuda/pyramid.cu:31-84 and its 2x box kernel :6-29) and the 7x7 Gaussian that
// replaces gaussian_blur_3x3 (reference src/cuda/gaussian_blur_3x3.cu:15-72).
//
// Semantics = upstream ORBextractor::ComputePyramid on OpenCV 4.13:
// level l is cv::resize(level l-1, INTER_LINEAR) in 11-bit fixed point, each level framed by a
// 19-px BORDER_REFLECT_101 border.  The weight/offset tables are built once on the host exactly
// like OpenCV builds them (float32/double mix), so device arithmetic is integer only.
//
// Roofline: HBM-bound by design (every level is written once and read once); see DESIGN.md for the
// measured distance to it.
#include "orbb_internal.cuh"

namespace orbb {

__device__ __forceinline__ int reflect101(int p, int len) {
    // one reflection is enough: |p| excursion is <= 19 and every level is >= 62 px
    p = p < 0 ? -p : p;
    return p >= len ? 2 * len - 2 - p : p;
}

// ---- level 0: copy the caller's frames into the padded layout + border.
// Work items of a frame are ordered so that warps do not diverge: first every 16-byte chunk that lies entirely
// inside the ROI (one 128-bit load + one 128-bit store; ROI rows are 16-byte aligned by layout), then the 32-bit
// words that touch the reflect-101 frame or the ragged right end, assembled bytewise through the reflect map.
// With an unaligned source (pointer, pitch or stride not a multiple of 16) n_int == 0 and every word takes the
// bytewise route.
__global__ void __launch_bounds__(256) k_level0(const uint8_t *__restrict__ in, size_t in_pitch, size_t in_stride,
                                                LevelDev L, int n_int, int n_edge, int frame_base) {
    pdl_trigger();  // the first resize kernel may be scheduled now; it waits for this grid before reading level 0
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int frame = blockIdx.y + frame_base;
    const uint8_t *fsrc = in + (size_t)blockIdx.y * in_stride;
    uint8_t *fdst = L.img + (size_t)frame * L.frame_stride;
    const int n_interior = n_int * L.rows;
    if (i < n_interior) {
        const int py = i / n_int, c = i - py * n_int;
        const int y = reflect101(py - ORBB_BORDER, L.h);
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(fsrc + (size_t)y * in_pitch + 16 * c));
        *reinterpret_cast<uint4 *>(fdst + (size_t)py * L.pitch + ORBB_ROI_X0 + 16 * c) = v;
        return;
    }
    const int e = i - n_interior;
    if (e >= n_edge * L.rows) return;
    const int py = e / n_edge, k = e - py * n_edge;
    // words 3..7 hold the left frame (bytes 12..31); the rest continue after the last interior chunk
    const int wd = k < 5 ? 3 + k : 8 + 4 * n_int + (k - 5);
    const uint8_t *src = fsrc + (size_t)reflect101(py - ORBB_BORDER, L.h) * in_pitch;
    uint32_t out = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int px = min(max(4 * wd + j - ORBB_PAD_X0, 0), L.w + 2 * ORBB_BORDER - 1);
        out |= (uint32_t)__ldg(src + reflect101(px - ORBB_BORDER, L.w)) << (8 * j);
    }
    *reinterpret_cast<uint32_t *>(fdst + (size_t)py * L.pitch + 4 * wd) = out;
}

// ---- level l from level l-1: tiled two-phase bilinear resize (cv::resize INTER_LINEAR, 8U).
// A CTA produces a 64-byte x 32-row tile of the PADDED output row layout (aligned 32-bit stores; the
// reflect-101 frame is produced in the same pass by reflecting the output coordinate).  The source
// window is staged in shared memory with coalesced 32-bit loads, then
//   H phase: T[src_row][dst_col] = (S[sx]*a0 + S[sx+1]*a1) >> 4   (u16, thread = dst column, the
//            column's (sx,a0,a1) lives in registers for all source rows)
//   V phase: dst = (((b0*T0)>>16) + ((b1*T1)>>16) + 2) >> 2        (thread = 4 columns x 4 rows)
// Exact 2x shrinks take OpenCV's INTER_AREA route: (a+b+c+d+2)>>2, same two phases with unit weights.
#define RS_TW 64
#define RS_TH 32
#define RS_THREADS 128

__global__ void __launch_bounds__(RS_THREADS) k_resize(const LevelDev *__restrict__ levels, int l, int src_pitch,
                                                       int src_rows_max, int frame_base) {
    extern __shared__ __align__(16) uint8_t rs_smem[];
    __shared__ int s_red[8];
    __shared__ int4 s_rowtab[RS_TH];  // per dst row of the tile: (src row0, src row1, b0, b1)
    const LevelDev &L = levels[l];
    const LevelDev &S = levels[l - 1];
    uint8_t *s_src = rs_smem;                                                         // [src_rows_max][src_pitch]
    uint16_t *s_t = reinterpret_cast<uint16_t *>(rs_smem + (((size_t)src_rows_max * src_pitch + 15) & ~(size_t)15));  // [src_rows_max][RS_TW]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int bx0 = blockIdx.x * RS_TW, py0 = blockIdx.y * RS_TH, frame = blockIdx.z + frame_base;
    const int pw = L.w + 2 * ORBB_BORDER, ph = L.h + 2 * ORBB_BORDER;
    const bool area = L.area2x != 0;

    // per-column coefficients (thread = dst column c)
    const int c = tid & (RS_TW - 1);
    int px = bx0 + c - ORBB_PAD_X0;
    px = min(max(px, 0), pw - 1);
    const int rx = reflect101(px - ORBB_BORDER, L.w);
    int sx, sx1, a0, a1;
    if (area) { sx = 2 * rx; sx1 = sx + 1; a0 = 1; a1 = 1; }
    else {
        sx = __ldg(&L.xofs[rx]);
        sx1 = min(sx + 1, S.w - 1);
        const short2 a = __ldg(&L.xalpha[rx]);
        a0 = a.x; a1 = a.y;
    }
    // per-row source rows (threads 0..RS_TH-1 each own one dst row of the tile)
    int r0 = 0x7fffffff, r1 = -1;
    if (tid < RS_TH) {
        int4 rt = make_int4(0, 0, 0, 0);
        if (py0 + tid < ph) {
            const int ry = reflect101(py0 + tid - ORBB_BORDER, L.h);
            if (area) { r0 = 2 * ry; r1 = r0 + 1; }
            else {
                const int2 yr = __ldg(&L.yrows[ry]);
                const short2 bw = __ldg(&L.ybeta[ry]);
                r0 = yr.x; r1 = yr.y; rt.z = bw.x; rt.w = bw.y;
            }
            rt.x = r0; rt.y = r1;
        }
        s_rowtab[tid] = rt;
    }
    // block-wide extents of the source window
    int cmin = __reduce_min_sync(0xffffffffu, sx), cmax = __reduce_max_sync(0xffffffffu, sx1);
    int rmin = __reduce_min_sync(0xffffffffu, r0), rmax = __reduce_max_sync(0xffffffffu, r1);
    if (lane == 0) { s_red[warp] = cmin; s_red[4 + warp] = cmax; }
    __syncthreads();
    cmin = min(min(s_red[0], s_red[1]), min(s_red[2], s_red[3]));
    cmax = max(max(s_red[4], s_red[5]), max(s_red[6], s_red[7]));
    __syncthreads();
    if (warp == 0 && lane == 0) { s_red[0] = rmin; s_red[1] = rmax; }  // rows live in warp 0 (RS_TH == 32)
    __syncthreads();
    rmin = s_red[0]; rmax = s_red[1];
    const int sa = cmin & ~3;
    const int nwords = (cmax - sa + 4) >> 2, nrows = rmax - rmin + 1;

    // stage the source window
    const uint8_t *sroi = S.img + (size_t)frame * S.frame_stride + (size_t)ORBB_BORDER * S.pitch + ORBB_ROI_X0;
    {
        const unsigned inv_nw = (1u << 20) / (unsigned)nwords + 1u;  // exact i / nwords for i * nwords < 2^20
        for (int i = tid; i < nrows * nwords; i += RS_THREADS) {
            const int r = (int)(((unsigned)i * inv_nw) >> 20), wd = i - r * nwords;
            reinterpret_cast<uint32_t *>(s_src + (size_t)r * src_pitch)[wd] =
                *reinterpret_cast<const uint32_t *>(sroi + (size_t)(rmin + r) * S.pitch + sa + 4 * wd);
        }
    }
    __syncthreads();
    // H phase
    {
        const uint8_t *p0 = s_src + (size_t)(tid >> 6) * src_pitch + (sx - sa);
        const int d1 = sx1 - sx, step = (RS_THREADS / RS_TW) * src_pitch;
        uint16_t *t = s_t + (tid >> 6) * RS_TW + c;
        const int sh = area ? 0 : 4;
#pragma unroll 4
        for (int r = tid >> 6; r < nrows; r += RS_THREADS / RS_TW) {
            *t = (uint16_t)((p0[0] * a0 + p0[d1] * a1) >> sh);
            p0 += step;
            t += (RS_THREADS / RS_TW) * RS_TW;
        }
    }
    __syncthreads();
    // V phase: thread = (quad q of 4 columns, row group g)
    const int q = tid & 15, g = tid >> 4;
    uint8_t *drow = L.img + (size_t)frame * L.frame_stride + bx0 + 4 * q;
    if (bx0 + 4 * q >= L.pitch) return;
#pragma unroll
    for (int k = 0; k < RS_TH / 8; ++k) {
        const int py = py0 + g + 8 * k;
        if (py >= ph) break;
        const int4 rt = s_rowtab[g + 8 * k];
        const int y0 = rt.x, y1 = rt.y, b0 = rt.z, b1 = rt.w;
        const uint2 ta = *reinterpret_cast<const uint2 *>(s_t + (y0 - rmin) * RS_TW + 4 * q);
        const uint2 tb = *reinterpret_cast<const uint2 *>(s_t + (y1 - rmin) * RS_TW + 4 * q);
        const int t0[4] = {(int)(ta.x & 0xffff), (int)(ta.x >> 16), (int)(ta.y & 0xffff), (int)(ta.y >> 16)};
        const int t1[4] = {(int)(tb.x & 0xffff), (int)(tb.x >> 16), (int)(tb.y & 0xffff), (int)(tb.y >> 16)};
        uint32_t out = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int v = area ? ((t0[j] + t1[j] + 2) >> 2)
                               : ((((b0 * t0[j]) >> 16) + ((b1 * t1[j]) >> 16) + 2) >> 2);
            out |= (uint32_t)v << (8 * j);
        }
        *reinterpret_cast<uint32_t *>(drow + (size_t)py * L.pitch) = out;
    }
}

// ---- level l from level l-1, table-driven form (the default; k_resize above is the fallback for scale steps
// whose 2-pixel source span does not fit an 8-byte window, i.e. scale factors > 2).
// thread = one aligned 4-byte group of one PADDED output row; a CTA covers 128 bytes x RS_ROWS consecutive rows so
// that the source rows shared by neighbouring output rows hit in L1.  No shared memory, no barriers, no dependent
// chains: every per-column and per-row quantity is precomputed on the host (rs_h / rs_v, with the reflect-101
// frame folded in), so a thread issues its 8 aligned 32-bit loads (two 8-byte windows on each of the two source
// rows) up front, gathers [S[sx],S[sx+1]] of two pixels with one PRMT, does the horizontal pass with IDP.2A
// against (a0 | a1<<16) and the vertical pass with IMAD.HI against b<<16.
#define RS_ROWS 8
struct ResizeArgs {  // by value: lives in the constant bank, no dependent global loads before the tables
    const uint8_t *src;  // source level, frame 0, padded row ORBB_BORDER (= ROI row 0), byte 0
    uint8_t *dst;        // destination level, frame 0, padded row 0, byte 12 (first group)
    long long src_stride, dst_stride;
    const uint4 *rs_h;
    const int4 *rs_v;
    int src_pitch, dst_pitch, nq, rows;
};

template <bool AREA>
__global__ void __launch_bounds__(32 * RS_ROWS) k_resize_rows(const ResizeArgs A, int frame_base) {
    const int q = blockIdx.x * 32 + threadIdx.x;
    const int py0 = blockIdx.y * (2 * RS_ROWS) + threadIdx.y;
    pdl_trigger();
    if (q >= A.nq || py0 >= A.rows) return;
    const int frame = blockIdx.z + frame_base;
    const uint4 hx = __ldg(A.rs_h + 2 * q), hw = __ldg(A.rs_h + 2 * q + 1);
    const uint8_t *sbase = A.src + (size_t)frame * A.src_stride;
    uint8_t *dbase = A.dst + (size_t)frame * A.dst_stride + 4 * q;
    constexpr int SH = AREA ? 0 : 4;
    // two output rows per thread (py0 and py0 + RS_ROWS): 16 independent loads in flight
    const int py1 = min(py0 + RS_ROWS, A.rows - 1);
    int4 v[2] = {__ldg(A.rs_v + py0), __ldg(A.rs_v + py1)};
    pdl_wait();  // the tables above are constant (ptxas still sinks their loads below the wait); the source level is
                 // written by the previous kernel
    unsigned w[2][8];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const uint8_t *p0 = sbase + (size_t)v[r].x * A.src_pitch, *p1 = sbase + (size_t)v[r].y * A.src_pitch;
        w[r][0] = __ldg(reinterpret_cast<const unsigned *>(p0 + hx.x));
        w[r][1] = __ldg(reinterpret_cast<const unsigned *>(p0 + hx.x + 4));
        w[r][2] = __ldg(reinterpret_cast<const unsigned *>(p0 + hx.z));
        w[r][3] = __ldg(reinterpret_cast<const unsigned *>(p0 + hx.z + 4));
        w[r][4] = __ldg(reinterpret_cast<const unsigned *>(p1 + hx.x));
        w[r][5] = __ldg(reinterpret_cast<const unsigned *>(p1 + hx.x + 4));
        w[r][6] = __ldg(reinterpret_cast<const unsigned *>(p1 + hx.z));
        w[r][7] = __ldg(reinterpret_cast<const unsigned *>(p1 + hx.z + 4));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        unsigned ta[4], tb[4];
        {
            const unsigned pa = __byte_perm(w[r][0], w[r][1], hx.y), pb = __byte_perm(w[r][2], w[r][3], hx.w);
            ta[0] = __dp2a_lo(hw.x, pa, 0u) >> SH; ta[1] = __dp2a_hi(hw.y, pa, 0u) >> SH;
            ta[2] = __dp2a_lo(hw.z, pb, 0u) >> SH; ta[3] = __dp2a_hi(hw.w, pb, 0u) >> SH;
        }
        {
            const unsigned pa = __byte_perm(w[r][4], w[r][5], hx.y), pb = __byte_perm(w[r][6], w[r][7], hx.w);
            tb[0] = __dp2a_lo(hw.x, pa, 0u) >> SH; tb[1] = __dp2a_hi(hw.y, pa, 0u) >> SH;
            tb[2] = __dp2a_lo(hw.z, pb, 0u) >> SH; tb[3] = __dp2a_hi(hw.w, pb, 0u) >> SH;
        }
        unsigned s[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            s[j] = AREA ? ta[j] + tb[j] + 2u
                        : __umulhi((unsigned)v[r].z, ta[j]) + __umulhi((unsigned)v[r].w, tb[j]) + 2u;
        // s < 1024: pack pairs in 16-bit halves, shift both at once, gather bytes 0 and 2
        const unsigned lo = (s[0] | (s[1] << 16)) >> 2, hi = (s[2] | (s[3] << 16)) >> 2;
        const int py = r == 0 ? py0 : py0 + RS_ROWS;
        if (py < A.rows) *reinterpret_cast<unsigned *>(dbase + (size_t)py * A.dst_pitch) = __byte_perm(lo, hi, 0x6420);
    }
}

// ---- several levels per launch for SMALL batches (a lone frame is the reference's operating mode).  A frame's
// per-level kernels are each shorter than a launch gap (4-5 us per level for ~1 us of work, even as programmatic
// dependents: every level costs a full L2 round trip chain).  Here a CTA owns a tile of the FIRST level of a group of
// up to PF_GROUP levels and derives its share of every level of the group in shared memory: the source window is
// staged once, level l+1 is computed from the level-l tile that is still in shared memory, and only the pixels the
// CTA owns are written out.  The halo a tile needs from its neighbours is recomputed, not exchanged (no inter-CTA
// synchronisation): 20-60 % more pixel operations, which a lone frame can afford (most SMs idle otherwise) and a
// large batch cannot -- large batches keep the per-level kernels above.  Every pixel value is the same integer
// formula whoever computes it, so the output is identical.  x coordinates are in "c" units: c = padded column + 1,
// so that c = 0 is byte 12 of a padded row and aligned 4-byte words are c = 4i .. 4i+3 (same grouping as k_resize_rows).
#define PF_THREADS 1024  // the per-warp row loops are latency bound: 32 warps per tile, one or two rows each
#define PF_WARPS (PF_THREADS / 32)
__global__ void __launch_bounds__(PF_THREADS) k_pyramid_fused(const uint8_t *__restrict__ in, size_t in_pitch, size_t in_stride,
                                                       const LevelDev *__restrict__ levels, int g0, int ng,
                                                       const PfTile *__restrict__ tiles, int frame_base) {
    extern __shared__ __align__(16) uint8_t pf_smem[];
    pdl_trigger();
    const PfTile &T = tiles[blockIdx.x];  // read through the pointer: the per-level arrays are indexed at run time
    const int frame = blockIdx.y + frame_base, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // Shared memory: [row tables][column tables][source window][level tiles ...].  The tables hold, for every row /
    // column this CTA computes on every level of the group, the source offsets (already relative to the source tile
    // in shared memory) and the fixed-point weights -- loaded ONCE, in parallel with the source window, so the
    // per-pixel loops below contain no global load (a dependent table load per pixel made the first version of this
    // kernel latency bound: 17 us per launch).
    int ntab_r = 0, ntab_c = 0;
    for (int k = 0; k < ng; ++k) { ntab_r += T.nh[k]; ntab_c += T.nw[k]; }
    int4 *rowt = reinterpret_cast<int4 *>(pf_smem);                       // {offset of source row 0, of source row 1, b0, b1}
    int2 *colt = reinterpret_cast<int2 *>(pf_smem + 16 * (size_t)ntab_r); // {offset of source column, a0 | a1 << 16}
    const int sw = T.sw, sh = T.sh, sp = (sw + 3) & ~3;
    uint8_t *S = pf_smem + 16 * (size_t)ntab_r + 8 * (size_t)ntab_c;      // source window (pitch = sw rounded up to 4)
    // ---- tables (constant data: may be read before the grid dependency resolves)
    {
        int rbase = 0, cbase = 0, src_x0 = T.sx0, src_y0 = T.sy0, src_p = sp;
        for (int k = 0; k < ng; ++k) {
            const LevelDev &L = levels[g0 + k];
            const int nw = T.nw[k], nh = T.nh[k], nx0 = T.nx0[k], ny0 = T.ny0[k];
            const int lw = L.w, lh = L.h, pw = lw + 2 * ORBB_BORDER, ph = lh + 2 * ORBB_BORDER;
            const bool copy0 = g0 == 0 && k == 0, area = L.area2x != 0;
            for (int y = tid; y < nh; y += PF_THREADS) {
                const int ry = reflect101(min(ny0 + y, ph - 1) - ORBB_BORDER, lh);
                int4 e;
                if (copy0) e = make_int4((ry - src_y0) * src_p, 0, 0, 0);
                else {
                    int r0, r1, b0 = 0, b1 = 0;
                    if (area) { r0 = 2 * ry; r1 = r0 + 1; }
                    else {
                        const int2 rr = __ldg(&L.yrows[ry]);
                        const short2 b = __ldg(&L.ybeta[ry]);
                        r0 = rr.x; r1 = rr.y; b0 = b.x; b1 = b.y;
                    }
                    // source pixel (ROI col sx, ROI row r) of the level below: c = sx + BORDER + 1, padded row r + BORDER
                    e = make_int4((r0 + ORBB_BORDER - src_y0) * src_p, (r1 + ORBB_BORDER - src_y0) * src_p, b0, b1);
                }
                rowt[rbase + y] = e;
            }
            for (int x = tid; x < nw; x += PF_THREADS) {
                const int rx = reflect101(min(max(nx0 + x - 1, 0), pw - 1) - ORBB_BORDER, lw);
                int2 e;
                if (copy0) e = make_int2(rx - src_x0, 0);
                else if (area) e = make_int2(2 * rx + ORBB_BORDER + 1 - src_x0, 1 | (1 << 16));
                else {
                    const short2 a = __ldg(&L.xalpha[rx]);
                    e = make_int2(__ldg(&L.xofs[rx]) + ORBB_BORDER + 1 - src_x0, (int)(unsigned short)a.x | ((int)a.y << 16));
                }
                colt[cbase + x] = e;
            }
            rbase += nh; cbase += nw;
            src_x0 = nx0; src_y0 = ny0; src_p = nw;  // the next level reads this level's tile
        }
    }
    pdl_wait();  // group >= 1 reads a level written by the previous launch
    // ---- source window: rows go to warps, columns to lanes (no index divisions)
    if (g0 == 0) {
        const uint8_t *src = in + (size_t)blockIdx.y * in_stride + (size_t)T.sy0 * in_pitch + T.sx0;
        for (int y = warp; y < sh; y += PF_WARPS)
            for (int x = lane; x < sw; x += 32) S[y * sp + x] = __ldg(src + (size_t)y * in_pitch + x);
    } else {
        const LevelDev &P = levels[g0 - 1];
        const uint8_t *src = P.img + (size_t)frame * P.frame_stride + (size_t)T.sy0 * P.pitch + 12 + T.sx0;  // sx0 multiple of 4
        const int wpr = sp >> 2;
        for (int y = warp; y < sh; y += PF_WARPS)
            for (int xw = lane; xw < wpr; xw += 32)
                reinterpret_cast<uint32_t *>(S + y * sp)[xw] = __ldcg(reinterpret_cast<const uint32_t *>(src + (size_t)y * P.pitch) + xw);
    }
    __syncthreads();
    const uint8_t *src_tile = S;
    uint8_t *D = S + sp * sh;
    int rbase = 0, cbase = 0;
    for (int k = 0; k < ng; ++k) {
        const LevelDev &L = levels[g0 + k];
        const int nw = T.nw[k], nh = T.nh[k], nx0 = T.nx0[k], ny0 = T.ny0[k];  // nw, nx0 multiples of 4
        const bool copy0 = g0 == 0 && k == 0, area = L.area2x != 0;
        for (int y = warp; y < nh; y += PF_WARPS) {
            const int4 rt = rowt[rbase + y];
            uint8_t *drow = D + y * nw;
            const uint8_t *p0 = src_tile + rt.x, *p1 = src_tile + rt.y;
            if (copy0) {  // level 0: the frame itself + its reflect-101 border
                for (int x = lane; x < nw; x += 32) drow[x] = p0[colt[cbase + x].x];
            } else if (area) {
                for (int x = lane; x < nw; x += 32) {
                    const int sx = colt[cbase + x].x;
                    drow[x] = (uint8_t)((p0[sx] + p0[sx + 1] + p1[sx] + p1[sx + 1] + 2) >> 2);
                }
            } else {
                for (int x = lane; x < nw; x += 32) {
                    const int2 ct = colt[cbase + x];
                    const int a0 = (short)(ct.y & 0xffff), a1 = ct.y >> 16;
                    const int t0 = p0[ct.x] * a0 + p0[ct.x + 1] * a1, t1 = p1[ct.x] * a0 + p1[ct.x + 1] * a1;
                    drow[x] = (uint8_t)((((rt.z * (t0 >> 4)) >> 16) + ((rt.w * (t1 >> 4)) >> 16) + 2) >> 2);
                }
            }
        }
        __syncthreads();
        // ---- write the owned rectangle, one aligned word per lane and step
        {
            uint8_t *dst = L.img + (size_t)frame * L.frame_stride + 12 + T.ox0[k];
            const int owq = T.ow[k] >> 2, oh = T.oh[k], oy0 = T.oy0[k];
            const uint8_t *drow0 = D + (oy0 - ny0) * nw + (T.ox0[k] - nx0);
            for (int y = warp; y < oh; y += PF_WARPS)
                for (int xw = lane; xw < owq; xw += 32)
                    *reinterpret_cast<uint32_t *>(dst + (size_t)(oy0 + y) * L.pitch + 4 * xw) =
                        *reinterpret_cast<const uint32_t *>(drow0 + y * nw + 4 * xw);
        }
        // the level just computed is the next level's source (it stays in shared memory; D moves on)
        src_tile = D;
        D += nw * nh;
        rbase += nh; cbase += nw;
    }
}

cudaError_t launch_pyramid_fused(const uint8_t *d_in, size_t pitch, size_t stride, const LevelDev *d_levels, int g0, int ng,
                                 const void *d_tiles, int n_tiles, size_t smem, int frame_base, int n_frames, cudaStream_t st) {
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_pyramid_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    return launch_pdl(k_pyramid_fused, dim3(n_tiles, n_frames), dim3(PF_THREADS), smem, st, d_in, pitch, stride, d_levels, g0, ng,
                      static_cast<const PfTile *>(d_tiles), frame_base);
}

size_t pyramid_fused_tile_bytes() { return sizeof(PfTile); }

// ---- 7x7 sigma=2 Gaussian of the ROI, OpenCV 4.13 fixed point: k = [18,34,48,56,48,34,18]/256 per
// pass, dst = (V + 32768) >> 16.  Reads the padded level, so REFLECT_101 at the ROI
// edge is already materialised by the border.
// One thread owns 4 adjacent columns and walks BLUR_R output rows down the image.  Per input row it
// issues three aligned, coalesced 32-bit loads (ROI rows are 16-byte aligned by layout), forms the
// byte windows with funnel shifts and does the horizontal pass with IDP.4A (2 per pixel); the vertical
// pass runs on a 7-deep register window, so no shared memory and no intermediate image exist.
#define BLUR_R 16
struct BlurIndex { int first[ORBB_MAX_LEVELS + 1]; };  // per-frame work-item prefix over levels

__global__ void __launch_bounds__(128) k_blur(const LevelDev *__restrict__ levels, int n_levels, BlurIndex bi,
                                              int frame_base) {
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= bi.first[n_levels]) return;
    int l = 0;
    while (item >= bi.first[l + 1]) ++l;
    const LevelDev &L = levels[l];
    const int frame = blockIdx.y + frame_base;
    const int wpr = (L.w + 3) >> 2;
    const int local = item - bi.first[l];
    const int strip = local / wpr, word = local - strip * wpr;
    const int y0 = strip * BLUR_R, x = word * 4;
    const uint8_t *roi = L.img + (size_t)frame * L.frame_stride + (size_t)ORBB_BORDER * L.pitch + ORBB_ROI_X0;
    uint8_t *dst = L.blur + (size_t)frame * L.blur_stride + x;
    const unsigned W1 = 18u | (34u << 8) | (48u << 16) | (56u << 24), W2 = 48u | (34u << 8) | (18u << 16);
    const int ymax = L.h + ORBB_BORDER - 1;
    unsigned h0[7], h1[7], h2[7], h3[7];
#pragma unroll
    for (int r = 0; r < BLUR_R + 6; ++r) {
        const int yy = min(y0 - 3 + r, ymax);
        const uint32_t *p = reinterpret_cast<const uint32_t *>(roi + (ptrdiff_t)yy * L.pitch + x - 4);
        const unsigned A = __ldg(p), B = __ldg(p + 1), C = __ldg(p + 2);
        const unsigned n0 = __dp4a(__funnelshift_r(A, B, 8), W1, __dp4a(__funnelshift_r(B, C, 8), W2, 0u));
        const unsigned n1 = __dp4a(__funnelshift_r(A, B, 16), W1, __dp4a(__funnelshift_r(B, C, 16), W2, 0u));
        const unsigned n2 = __dp4a(__funnelshift_r(A, B, 24), W1, __dp4a(__funnelshift_r(B, C, 24), W2, 0u));
        const unsigned n3 = __dp4a(B, W1, __dp4a(C, W2, 0u));
#pragma unroll
        for (int k = 0; k < 6; ++k) { h0[k] = h0[k + 1]; h1[k] = h1[k + 1]; h2[k] = h2[k + 1]; h3[k] = h3[k + 1]; }
        h0[6] = n0; h1[6] = n1; h2[6] = n2; h3[6] = n3;
        if (r >= 6) {
            const int y = y0 + r - 6;
            if (y < L.h) {
                const unsigned v0 = 32768u + 18u * (h0[0] + h0[6]) + 34u * (h0[1] + h0[5]) + 48u * (h0[2] + h0[4]) + 56u * h0[3];
                const unsigned v1 = 32768u + 18u * (h1[0] + h1[6]) + 34u * (h1[1] + h1[5]) + 48u * (h1[2] + h1[4]) + 56u * h1[3];
                const unsigned v2 = 32768u + 18u * (h2[0] + h2[6]) + 34u * (h2[1] + h2[5]) + 48u * (h2[2] + h2[4]) + 56u * h2[3];
                const unsigned v3 = 32768u + 18u * (h3[0] + h3[6]) + 34u * (h3[1] + h3[5]) + 48u * (h3[2] + h3[4]) + 56u * h3[3];
                const unsigned lo = __byte_perm(v0, v1, 0x0062), hi = __byte_perm(v2, v3, 0x0062);
                *reinterpret_cast<uint32_t *>(dst + (size_t)y * L.pitch) = __byte_perm(lo, hi, 0x5410);
            }
        }
    }
}

// ---------------------------------------------------------------- host launchers
cudaError_t launch_level0(const uint8_t *d_in, size_t pitch, size_t stride, const LevelDev &L0, int frame_base,
                          int n_frames, cudaStream_t st) {
    const int aligned16 = ((reinterpret_cast<uintptr_t>(d_in) | pitch | stride) & 15) == 0;
    const int n_int = aligned16 ? L0.w / 16 : 0;
    const int last_word = (ORBB_ROI_X0 + L0.w + ORBB_BORDER + 3) / 4;  // one past the last word holding a padded pixel
    const int n_edge = 5 + last_word - (8 + 4 * n_int);
    const int items = (n_int + n_edge) * L0.rows;
    dim3 grid((items + 255) / 256, n_frames);
    k_level0<<<grid, 256, 0, st>>>(d_in, pitch, stride, L0, n_int, n_edge, frame_base);
    return cudaGetLastError();
}

cudaError_t launch_resize(const LevelDev *d_levels, const LevelDev *h_levels, int l, int frame_base, int n_frames,
                          cudaStream_t st) {
    const LevelDev &Lh = h_levels[l], &Sh = h_levels[l - 1];
    if (Lh.rs_ok) {
        ResizeArgs A;
        A.src = Sh.img + (size_t)ORBB_BORDER * Sh.pitch;
        A.dst = Lh.img + 12;
        A.src_stride = Sh.frame_stride; A.dst_stride = Lh.frame_stride;
        A.rs_h = Lh.rs_h; A.rs_v = Lh.rs_v;
        A.src_pitch = Sh.pitch; A.dst_pitch = Lh.pitch; A.nq = Lh.rs_nq; A.rows = Lh.rows;
        dim3 grid((Lh.rs_nq + 31) / 32, (Lh.rows + 2 * RS_ROWS - 1) / (2 * RS_ROWS), n_frames), block(32, RS_ROWS);
        if (Lh.area2x) return launch_pdl(k_resize_rows<true>, grid, block, 0, st, A, frame_base);
        return launch_pdl(k_resize_rows<false>, grid, block, 0, st, A, frame_base);
    }
    // worst-case source window of one tile (+ slack for clamping/alignment)
    const int src_cols = (int)((double)RS_TW * Sh.w / Lh.w) + 8;
    const int src_pitch = ((src_cols + 3) / 4 + 1) * 4;
    const int src_rows = (int)((double)RS_TH * Sh.h / Lh.h) + 5;
    const size_t smem = (((size_t)src_rows * src_pitch + 15) & ~(size_t)15) + (size_t)src_rows * RS_TW * 2;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_resize, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    dim3 grid((Lh.pitch + RS_TW - 1) / RS_TW, (Lh.rows + RS_TH - 1) / RS_TH, n_frames);
    k_resize<<<grid, RS_THREADS, smem, st>>>(d_levels, l, src_pitch, src_rows, frame_base);
    return cudaGetLastError();
}

cudaError_t launch_blur(const LevelDev *d_levels, const LevelDev *h_levels, int n_levels, int frame_base,
                        int n_frames, cudaStream_t st) {
    BlurIndex bi;
    bi.first[0] = 0;
    for (int l = 0; l < n_levels; ++l)
        bi.first[l + 1] = bi.first[l] + ((h_levels[l].h + BLUR_R - 1) / BLUR_R) * ((h_levels[l].w + 3) >> 2);
    for (int l = n_levels + 1; l <= ORBB_MAX_LEVELS; ++l) bi.first[l] = bi.first[n_levels];
    dim3 grid((bi.first[n_levels] + 127) / 128, n_frames);
    k_blur<<<grid, 128, 0, st>>>(d_levels, n_levels, bi, frame_base);
    return cudaGetLastError();
}

}  // namespace orbb
