// orbb_api.cu -- C ABI of liborbb200.so (include/orbb200.h): handle, geometry tables, stage
// sequencing.  Host-side mirror of the stage-calling block of SlamGpuPipeline::buildStream
// (reference src/SlamGpuPipeline/buildStream.cpp:208-341 buffer prologue, :416-460 stage order) and
// of upstream ORBextractor's constructor.  All device memory is allocated once here;
// nothing on the per-batch path allocates, frees or synchronises (except the *_host / debug calls).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "orbb_internal.cuh"

namespace orbb {
cudaError_t launch_level0(const uint8_t *, size_t, size_t, const LevelDev &, int, int, cudaStream_t);
cudaError_t launch_resize(const LevelDev *, const LevelDev *, int, int, int, cudaStream_t);
cudaError_t launch_pyramid_fused(const uint8_t *, size_t, size_t, const LevelDev *, int, int, const void *, int, size_t, int, int,
                                 cudaStream_t);
cudaError_t launch_blur(const LevelDev *, const LevelDev *, int, int, int, cudaStream_t);
cudaError_t launch_fast(const void *, const LevelDev *, const CellEntry *, int, int, int *, int, int, const FastSmemCfg &,
                        int, int, cudaStream_t);
bool build_fast_tma_maps(void *, const LevelDev *, int, int, const FastSmemCfg &);
size_t fast_tma_maps_bytes();
cudaError_t launch_fast_dump(const void *, const LevelDev *, const CellEntry *, int, int, int, int, const FastSmemCfg &, int,
                             uint8_t *, const long long *, cudaStream_t);
cudaError_t launch_octree_bin(const LevelDev *, int, const int *, int, int, int, cudaStream_t);
cudaError_t launch_octree(const LevelDev *, int, const int *, int *, int, int, int, int, int, int, int, int,
                          cudaStream_t);
size_t octree_dyn_smem(int, int, int);
cudaError_t launch_match_windowed(const uint8_t *, const void *, int, int, const uint8_t *, const void *, int, int, float, int,
                                  int *, int *, int *, cudaStream_t);
cudaError_t launch_angle_orb(const LevelDev *, int, const int *, const int8_t *, const int *, const int *, int, int, int,
                             orbb_keypoint *, uint8_t *, int *, int, cudaStream_t);
cudaError_t launch_fast_angle_pos(float *, const float *, const uint8_t *, int, int, int, int, cudaStream_t);
cudaError_t launch_calc_orb_pos(const float *, const float *, uint8_t *, const uint8_t *, int, int, int, int, const int8_t *,
                                cudaStream_t);
cudaError_t launch_detect_export(const LevelDev *, int, const int *, const int *, const int *, int, int, float *, float *, int *,
                                 int *, int *, int, cudaStream_t);
cudaError_t launch_match(const uint8_t *, const uint8_t *, const int *, const int *, int, int, int, int, int, int4 *,
                         int, int, float, int *, int *, uint8_t *, int *, cudaStream_t, const int * = nullptr, int = 0,
                         uint8_t * = nullptr, size_t = 0, bool = true, long long * = nullptr);
int matcher_kind();
bool matcher_pre();
cudaError_t launch_popc_rate(int, int, long long *, unsigned *, cudaStream_t);
cudaError_t launch_imma_rate(int, int, long long *, unsigned *, cudaStream_t);
cudaError_t launch_align(const uint16_t *, int, float, const orbb_intrinsics &, const orbb_intrinsics &, const orbb_extrinsics &,
                         uint32_t *, cudaStream_t);
cudaError_t launch_kp_to_point(const uint32_t *, const orbb_intrinsics &, int, const orbb_keypoint *, const uint8_t *,
                               const int *, int, orbb_keypoint *, uint8_t *, double *, int *, cudaStream_t);
cudaError_t launch_reproject(const double *, const int *, int, int, const double *, const orbb_intrinsics &, float *,
                             cudaStream_t);
cudaError_t launch_compact_pairs(const int *, const int *, int, int, const double *, const double *, const void *, int,
                                 double *, double *, uint16_t *, int *, cudaStream_t);
cudaError_t launch_match_projection(const uint8_t *, const float *, const orbb_keypoint *, const int *, const uint8_t *,
                                    const orbb_keypoint *, const int *, int, int, float, int, int, const float *, int, int *, int *,
                                    int *, cudaStream_t);
cudaError_t launch_stereo(const LevelDev *, const float *, const float *, int, const orbb_keypoint *, const uint8_t *, const int *,
                          int, int, float, float, float *, float *, int *, int *, cudaStream_t);
cudaError_t launch_rgb_to_gray(const uint8_t *, size_t, size_t, int, int, int, uint8_t *, size_t, size_t, cudaStream_t);
cudaError_t launch_match_windowed_batch(const uint8_t *, const void *, int, const int *, const uint8_t *, const void *, int,
                                        const int *, int, int, float, int, int *, int *, cudaStream_t);
}  // namespace orbb

using namespace orbb;

#define ORBB_MAX_CHUNKS 16  // pipeline depth of orbb_extract_batch_host
// Split-T scratch of the matcher: one int4 per (train split, query), allocated ONCE in orbb_create.  This bounds
// n_split * nq beyond nq itself (pick_split asks for at most 64 CTAs per SM of 256 queries each, plus one round-up).
#define ORBB_MATCH_SPLIT_ROWS ((size_t)64 * 148 * 256 + 256)
#define ORBB_MATCH_EXP_ROWS ((size_t)1 << 18)  // 64 MB: train sets up to 262 144 rows take the pre-expanded path

static const int8_t k_pattern_host[1024] = {
#include "../../include/orb_pattern_31.inc"
};

struct orbb_handle {
    orbb_params p{};
    int w = 0, h = 0, max_batch = 0, device = 0, nlevels = 0;
    LevelDev lv[ORBB_MAX_LEVELS]{};
    LevelDev *d_levels = nullptr;
    float sf[ORBB_MAX_LEVELS]{}, inv_sf[ORBB_MAX_LEVELS]{};
    CellEntry *d_cells = nullptr;
    int n_cells = 0;
    int cell_first[ORBB_MAX_LEVELS + 1] = {};  // cells are stored level by level: level l owns [cell_first[l], cell_first[l+1])
    FastSmemCfg fcfg{};
    void *tma_maps = nullptr;  // DEVICE array of CUtensorMap, one per level (nullptr: FAST stages manually)
    int t_lo = 7, t_hi = 20;
    int *d_cand_count = nullptr, *d_sel_count = nullptr;
    int8_t *d_pattern = nullptr;
    int *d_slot_level = nullptr, *d_slot_base = nullptr;
    int n_slots = 0, sel_cap_max = 0, pcap = 0, pcap2 = 0, max_kp = 0;
    int4 *d_partial = nullptr;
    size_t partial_cap = 0;
    uint8_t *d_train_exp = nullptr;  // the train set as tcgen05 B tiles (256 B per row), ORBB_MATCH_EXP_ROWS rows
    size_t train_exp_rows = 0;
    int *d_stereo_sad = nullptr;  // SAD cost per left keypoint (stereo matcher scratch), (max_batch/2) x max_kp
    size_t stereo_cap = 0;
    // staging of the *_host entry points, double-buffered so consecutive batches overlap
    uint8_t *d_in2[2] = {nullptr, nullptr};
    orbb_keypoint *d_kp2[2] = {nullptr, nullptr};
    uint8_t *d_desc2[2] = {nullptr, nullptr};
    int *d_counts2[2] = {nullptr, nullptr};
    cudaEvent_t ev_ticket[2] = {nullptr, nullptr}, ev_tail[2] = {nullptr, nullptr};
    long long n_submitted = 0;
    long long n_launches = 0;  // kernels launched (not memsets/copies)
    uint8_t *d_dump = nullptr;
    long long *d_dump_off = nullptr;
    std::vector<long long> dump_off;
    int n_frames_last = 0;
    cudaStream_t s_in = nullptr, s_out = nullptr;  // copy streams of the pipelined *_host entry point
    cudaStream_t s_comp[2] = {nullptr, nullptr};   // alternating compute streams (chunk tails overlap)
    cudaStream_t s_side = nullptr;                 // side stream of the device entry point (blur under FAST/quadtree)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaStream_t s_dev[7] = {};                    // extra streams of the device entry point (batch split in parts)
    cudaEvent_t ev_dev[7] = {};
    int dev_split = 2;
    // CUDA graphs of the small-batch device entry point (launch bound: 14 kernels for ~60 us of work per frame)
    struct GraphSlot {
        const void *img = nullptr; size_t pitch = 0, stride = 0; int n = 0; void *kp = nullptr, *desc = nullptr, *cnt = nullptr;
        int max_kp = 0; cudaGraphExec_t exec = nullptr; long long launches = 0, stamp = 0;
    } graphs[4];
    long long graph_clock = 0;
    int use_graphs = 1;  // small batches replay a captured graph (ORBB_GRAPH=0: plain stream launches); see orbb_extract_batch_device
    int graph_miss_streak = 0;
    int last_host_frames = -1, last_host_latency = -1;  // chunk layout of the previous host submission
    // several pyramid levels per launch (k_pyramid_fused), small batches only: per group of <= 4 levels a tile table
    struct FusedGroup { int g0 = 0, ng = 0, n_tiles = 0; size_t smem = 0; PfTile *d_tiles = nullptr; };
    std::vector<FusedGroup> fused;
    // Opt-in (ORBB_FUSED_PYRAMID=<largest batch>, default 0 = never).  Measured on B200, one 848x480 frame, 8 levels:
    // the two fused launches take 15.5 + 14.1 us against 2.9 + 7 x 4.4 us for level0 + seven resizes that overlap as
    // programmatic dependents; the call latency is 79 us fused vs 65-68 us per level.  The tile pyramids are latency
    // bound (dependent shared-memory loads, ~2 IPC per SM, 1.5-2.2 x redundant halo pixels), so fewer launches do not
    // pay here; kept because it is bit-exact, tested, and the starting point if the per-pixel cost can be halved.
    int fused_max_frames = 0;
    cudaEvent_t ev_fence = nullptr, ev_done[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> ev_in, ev_comp;
    std::vector<void *> allocs;
    char cuda_err[256] = {0};
};

#define CK(h, call)                                                                                   \
    do {                                                                                              \
        cudaError_t e__ = (call);                                                                     \
        if (e__ != cudaSuccess) {                                                                     \
            snprintf((h)->cuda_err, sizeof((h)->cuda_err), "%s:%d %s: %s", __FILE__, __LINE__, #call, \
                     cudaGetErrorString(e__));                                                        \
            return ORBB_ERR_CUDA;                                                                     \
        }                                                                                             \
    } while (0)

template <typename T>
static cudaError_t dalloc(orbb_handle *h, T **out, size_t count) {
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T));
    if (e == cudaSuccess) {
        h->allocs.push_back(p);
        *out = static_cast<T *>(p);
    }
    return e;
}

template <typename T>
static cudaError_t upload(orbb_handle *h, T **out, const std::vector<T> &v) {
    cudaError_t e = dalloc(h, out, v.size());
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*out, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
}

static inline int cv_round(float v) { return (int)lrintf(v); }
static inline size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline uint32_t spread_bits(uint32_t v) {  // bit i -> bit 2i
    uint32_t r = 0;
    for (int i = 0; i < 16; ++i) r |= ((v >> i) & 1u) << (2 * i);
    return r;
}

extern "C" const char *orbb_strerror(int s) {
    switch (s) {
        case ORBB_OK: return "ok";
        case ORBB_ERR_INVALID: return "invalid argument";
        case ORBB_ERR_NO_DEVICE: return "no usable CUDA device (this library has no CPU path)";
        case ORBB_ERR_CUDA: return "CUDA runtime error";
        case ORBB_ERR_TOO_SMALL: return "image too small: a pyramid level would be under 62 px";
        case ORBB_ERR_CAPACITY: return "capacity exceeded";
        case ORBB_ERR_SHAPE: return "unsupported image shape";
        default: return "unknown status";
    }
}

extern "C" const char *orbb_last_cuda_error(const orbb_handle *h) { return h ? h->cuda_err : ""; }

extern "C" int orbb_destroy(orbb_handle *h) {
    if (!h) return ORBB_OK;
    cudaSetDevice(h->device);
    for (void *p : h->allocs) cudaFree(p);
    for (cudaEvent_t e : h->ev_in) cudaEventDestroy(e);
    for (cudaEvent_t e : h->ev_comp) cudaEventDestroy(e);
    if (h->s_in) cudaStreamDestroy(h->s_in);
    if (h->s_out) cudaStreamDestroy(h->s_out);
    for (int i = 0; i < 2; ++i) {
        if (h->s_comp[i]) cudaStreamDestroy(h->s_comp[i]);
        if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]);
    }
    if (h->ev_fence) cudaEventDestroy(h->ev_fence);
    if (h->s_side) cudaStreamDestroy(h->s_side);
    for (auto &g : h->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    for (int i = 0; i < 7; ++i) {
        if (h->s_dev[i]) cudaStreamDestroy(h->s_dev[i]);
        if (h->ev_dev[i]) cudaEventDestroy(h->ev_dev[i]);
    }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    for (int i = 0; i < 2; ++i) {
        if (h->ev_ticket[i]) cudaEventDestroy(h->ev_ticket[i]);
        if (h->ev_tail[i]) cudaEventDestroy(h->ev_tail[i]);
    }
    delete h;
    return ORBB_OK;
}

// cv::resize(INTER_LINEAR) coefficient tables, built exactly like OpenCV imgproc/resize.cpp
static void build_resize_tables(int sw, int sh, int dw, int dh, std::vector<int> &xofs, std::vector<short2> &xalpha,
                                std::vector<int2> &yrows, std::vector<short2> &ybeta) {
    const double inv_x = (double)dw / sw, inv_y = (double)dh / sh;
    const double scale_x = 1. / inv_x, scale_y = 1. / inv_y;
    xofs.resize(dw); xalpha.resize(dw); yrows.resize(dh); ybeta.resize(dh);
    for (int dx = 0; dx < dw; ++dx) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = (int)floorf(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        xalpha[dx] = make_short2((short)cv_round((1.f - fx) * 2048.f), (short)cv_round(fx * 2048.f));
    }
    for (int dy = 0; dy < dh; ++dy) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = (int)floorf(fy);
        fy -= sy;
        ybeta[dy] = make_short2((short)cv_round((1.f - fy) * 2048.f), (short)cv_round(fy * 2048.f));
        yrows[dy] = make_int2(std::min(std::max(sy, 0), sh - 1), std::min(std::max(sy + 1, 0), sh - 1));
    }
}

// Tile tables of k_pyramid_fused.  Levels are grouped by ORBB_PF_GROUP; a tile owns a rectangle of the group's first
// level, and on every further level the rectangle between the same scaled boundaries (rounded to multiples of 4 in c
// units, so all stores are aligned words and the rectangles partition each padded level).  The rectangle a tile must
// COMPUTE on level l is what it owns plus the source footprint of what it computes on level l+1.
static bool build_fused_tiles(orbb_handle *h, const std::vector<std::vector<int>> &xofs, const std::vector<std::vector<int2>> &yrows,
                              std::vector<std::vector<PfTile>> &out_tiles) {
    const int nl = h->nlevels;
    auto refl = [](int p, int len) { p = p < 0 ? -p : p; return p >= len ? 2 * len - 2 - p : p; };
    for (int g0 = 0; g0 < nl; g0 += ORBB_PF_GROUP) {
        const int ng = std::min(ORBB_PF_GROUP, nl - g0);
        for (int k = (g0 == 0 ? 1 : 0); k < ng; ++k)
            if (!h->lv[g0 + k].rs_ok && !h->lv[g0 + k].area2x) return false;  // scale step > 2: tiled fallback kernel only
        const LevelDev &F = h->lv[g0];
        // tile size on the group's first level: enough tiles to cover the SMs for ONE frame, at most ~8 KB of pixels each
        const int cw_total = (F.w + 2 * ORBB_BORDER + 1 + 3) & ~3, ph_total = F.h + 2 * ORBB_BORDER;
        static const int cand_tiles[4][2] = {{64, 32}, {64, 16}, {32, 16}, {32, 8}};
        int tw = 16, th = 8;
        for (const auto &c : cand_tiles)
            if ((long long)((cw_total + c[0] - 1) / c[0]) * ((ph_total + c[1] - 1) / c[1]) >= 120) { tw = c[0]; th = c[1]; break; }
        const int ntx = (cw_total + tw - 1) / tw, nty = (ph_total + th - 1) / th;
        std::vector<PfTile> tiles;
        size_t smem_max = 0;
        for (int ty = 0; ty < nty; ++ty)
            for (int tx = 0; tx < ntx; ++tx) {
                PfTile T{};
                int nx0[ORBB_PF_GROUP], nx1[ORBB_PF_GROUP], ny0[ORBB_PF_GROUP], ny1[ORBB_PF_GROUP];
                for (int k = 0; k < ng; ++k) {  // owned rectangles: scaled tile boundaries
                    const LevelDev &L = h->lv[g0 + k];
                    const int cwl = (L.w + 2 * ORBB_BORDER + 1 + 3) & ~3, phl = L.h + 2 * ORBB_BORDER;
                    auto bx = [&](int t) { return t >= ntx ? cwl : (int)(((long long)t * tw * cwl / cw_total) + 2) / 4 * 4; };
                    auto by = [&](int t) { return t >= nty ? phl : (int)((long long)t * th * phl / ph_total); };
                    T.ox0[k] = (short)bx(tx); T.ow[k] = (short)(bx(tx + 1) - bx(tx));
                    T.oy0[k] = (short)by(ty); T.oh[k] = (short)(by(ty + 1) - by(ty));
                }
                for (int k = ng - 1; k >= 0; --k) {  // compute rectangles, last level first
                    int x0 = T.ox0[k], x1 = T.ox0[k] + T.ow[k], y0 = T.oy0[k], y1 = T.oy0[k] + T.oh[k];
                    if (k < ng - 1 && nx1[k + 1] > nx0[k + 1] && ny1[k + 1] > ny0[k + 1]) {
                        const LevelDev &N = h->lv[g0 + k + 1];  // the level that reads level g0 + k
                        const int pwn = N.w + 2 * ORBB_BORDER, phn = N.h + 2 * ORBB_BORDER;
                        int fx0 = 1 << 30, fx1 = -1, fy0 = 1 << 30, fy1 = -1;
                        for (int c = nx0[k + 1]; c < nx1[k + 1]; ++c) {
                            const int rx = refl(std::min(std::max(c - 1, 0), pwn - 1) - ORBB_BORDER, N.w);
                            const int sx = N.area2x ? 2 * rx : xofs[g0 + k + 1][rx];
                            fx0 = std::min(fx0, sx + ORBB_BORDER + 1); fx1 = std::max(fx1, sx + ORBB_BORDER + 1 + 2);
                        }
                        for (int py = ny0[k + 1]; py < ny1[k + 1]; ++py) {
                            const int ry = refl(std::min(py, phn - 1) - ORBB_BORDER, N.h);
                            const int r0 = N.area2x ? 2 * ry : yrows[g0 + k + 1][ry].x, r1 = N.area2x ? 2 * ry + 1 : yrows[g0 + k + 1][ry].y;
                            fy0 = std::min(fy0, std::min(r0, r1) + ORBB_BORDER); fy1 = std::max(fy1, std::max(r0, r1) + ORBB_BORDER + 1);
                        }
                        if (T.ow[k] <= 0 || T.oh[k] <= 0) { x0 = fx0; x1 = fx1; y0 = fy0; y1 = fy1; }
                        else { x0 = std::min(x0, fx0); x1 = std::max(x1, fx1); y0 = std::min(y0, fy0); y1 = std::max(y1, fy1); }
                    }
                    nx0[k] = x0 & ~3; nx1[k] = (x1 + 3) & ~3; ny0[k] = y0; ny1[k] = y1;
                    if (nx1[k] <= nx0[k] || ny1[k] <= ny0[k]) { nx1[k] = nx0[k]; ny1[k] = ny0[k]; }
                    T.nx0[k] = (short)nx0[k]; T.nw[k] = (short)(nx1[k] - nx0[k]); T.ny0[k] = (short)ny0[k]; T.nh[k] = (short)(ny1[k] - ny0[k]);
                }
                // source window of the first level of the group
                int sx0 = 1 << 30, sx1 = -1, sy0 = 1 << 30, sy1 = -1;
                {
                    const LevelDev &N = h->lv[g0];
                    const int pwn = N.w + 2 * ORBB_BORDER, phn = N.h + 2 * ORBB_BORDER;
                    for (int c = nx0[0]; c < nx1[0]; ++c) {
                        const int rx = refl(std::min(std::max(c - 1, 0), pwn - 1) - ORBB_BORDER, N.w);
                        if (g0 == 0) { sx0 = std::min(sx0, rx); sx1 = std::max(sx1, rx + 1); }
                        else {
                            const int sx = N.area2x ? 2 * rx : xofs[g0][rx];
                            sx0 = std::min(sx0, sx + ORBB_BORDER + 1); sx1 = std::max(sx1, sx + ORBB_BORDER + 1 + 2);
                        }
                    }
                    for (int py = ny0[0]; py < ny1[0]; ++py) {
                        const int ry = refl(std::min(py, phn - 1) - ORBB_BORDER, N.h);
                        if (g0 == 0) { sy0 = std::min(sy0, ry); sy1 = std::max(sy1, ry + 1); }
                        else {
                            const int r0 = N.area2x ? 2 * ry : yrows[g0][ry].x, r1 = N.area2x ? 2 * ry + 1 : yrows[g0][ry].y;
                            sy0 = std::min(sy0, std::min(r0, r1) + ORBB_BORDER); sy1 = std::max(sy1, std::max(r0, r1) + ORBB_BORDER + 1);
                        }
                    }
                    if (sx1 < 0 || sy1 < 0) { sx0 = sx1 = sy0 = sy1 = 0; }
                    if (g0 > 0) { sx0 &= ~3; sx1 = (sx1 + 3) & ~3; }
                }
                T.sx0 = (short)sx0; T.sy0 = (short)sy0; T.sw = (short)(sx1 - sx0); T.sh = (short)(sy1 - sy0);
                size_t bytes = (size_t)((T.sw + 3) & ~3) * T.sh;
                for (int k = 0; k < ng; ++k) bytes += (size_t)T.nw[k] * T.nh[k] + 8 * (size_t)T.nw[k] + 16 * (size_t)T.nh[k];  // tile + tables
                smem_max = std::max(smem_max, bytes);
                tiles.push_back(T);
            }
        if (smem_max + 64 > 200 * 1024) return false;
        orbb_handle::FusedGroup G;
        G.g0 = g0; G.ng = ng; G.n_tiles = (int)tiles.size(); G.smem = (smem_max + 15) & ~(size_t)15;
        if (getenv("ORBB_FUSED_DEBUG")) {
            long long need = 0, own = 0;
            for (const PfTile &T : tiles) for (int k = 0; k < ng; ++k) { need += (long long)T.nw[k] * T.nh[k]; own += (long long)T.ow[k] * T.oh[k]; }
            fprintf(stderr, "fused pyramid group %d..%d: %d tiles of %dx%d (grid %dx%d), smem %zu B, computed / written pixels %.2f\n", g0,
                    g0 + ng - 1, G.n_tiles, tw, th, ntx, nty, G.smem, (double)need / (double)std::max(own, 1LL));
            const PfTile &T = tiles[tiles.size() / 2];
            for (int k = 0; k < ng; ++k) fprintf(stderr, "  mid tile level %d: need %dx%d @(%d,%d) own %dx%d @(%d,%d)\n", g0 + k, T.nw[k], T.nh[k], T.nx0[k], T.ny0[k], T.ow[k], T.oh[k], T.ox0[k], T.oy0[k]);
            fprintf(stderr, "  source %dx%d @(%d,%d)\n", T.sw, T.sh, T.sx0, T.sy0);
        }
        h->fused.push_back(G);
        out_tiles.push_back(tiles);
    }
    return true;
}

extern "C" int orbb_create(orbb_handle **out, const orbb_params *params, int width, int height, int max_batch,
                           int device) {
    if (!out || !params) return ORBB_ERR_INVALID;
    *out = nullptr;
    if (params->nlevels < 1 || params->nlevels > ORBB_MAX_LEVELS || !(params->scale_factor > 1.0f) ||
        params->nfeatures < 1 || max_batch < 1 || width < 1 || height < 1 || width > 4096 || height > 4096)
        return ORBB_ERR_INVALID;
    if (max_batch > 65535) return ORBB_ERR_CAPACITY;  // the frame index rides in gridDim.y / gridDim.z
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return ORBB_ERR_NO_DEVICE; }
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return ORBB_ERR_NO_DEVICE;
    if (device >= ndev || cudaSetDevice(device) != cudaSuccess) return ORBB_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10) return ORBB_ERR_NO_DEVICE;

    orbb_handle *h = new (std::nothrow) orbb_handle();
    if (!h) return ORBB_ERR_INVALID;
    h->p = *params; h->w = width; h->h = height; h->max_batch = max_batch; h->device = device;
    h->nlevels = params->nlevels;
    const int nl = h->nlevels, B = max_batch;
    // thresholds: cv::FAST clamps to [0,255]; ini < min behaves like min := ini (K(t2) subset K(t1))
    h->t_hi = std::min(std::max(params->ini_th_fast, 0), 255);
    h->t_lo = std::min(std::min(std::max(params->min_th_fast, 0), 255), h->t_hi);

    // ---- upstream constructor: scale chain and features per level (float32 arithmetic)
    h->sf[0] = 1.0f;
    for (int i = 1; i < nl; ++i) h->sf[i] = h->sf[i - 1] * params->scale_factor;
    for (int i = 0; i < nl; ++i) h->inv_sf[i] = 1.0f / h->sf[i];
    int nfeat[ORBB_MAX_LEVELS];
    {
        const float factor = 1.0f / params->scale_factor;
        float nd = params->nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nl));
        int sum = 0;
        for (int l = 0; l < nl - 1; ++l) {
            nfeat[l] = cv_round(nd);
            sum += nfeat[l];
            nd *= factor;
        }
        nfeat[nl - 1] = std::max(params->nfeatures - sum, 0);
    }

#define CKC(call)                                                                                          \
    do {                                                                                                   \
        cudaError_t e__ = (call);                                                                          \
        if (e__ != cudaSuccess) {                                                                          \
            fprintf(stderr, "orbb_create: %s failed: %s\n", #call, cudaGetErrorString(e__));               \
            orbb_destroy(h);                                                                               \
            return ORBB_ERR_CUDA;                                                                          \
        }                                                                                                  \
    } while (0)

    std::vector<CellEntry> cells;
    std::vector<int> slot_level, slot_base(nl);
    std::vector<std::vector<int>> h_xofs(nl);    // host copies of the resize tables: the fused-pyramid tile builder needs
    std::vector<std::vector<int2>> h_yrows(nl);  // every level's source footprint
    int max_cw = 1, max_ch = 1;
    h->dump_off.resize(nl + 1);
    long long dump_total = 0;
    for (int l = 0; l < nl; ++l) {
        LevelDev &L = h->lv[l];
        L.w = cv_round((float)width * h->inv_sf[l]);
        L.h = cv_round((float)height * h->inv_sf[l]);
        if (L.w - 32 < 30 || L.h - 32 < 30) { orbb_destroy(h); return ORBB_ERR_TOO_SMALL; }
        L.pitch = (int)round_up((size_t)ORBB_ROI_X0 + L.w + ORBB_BORDER, 128);
        L.rows = L.h + 2 * ORBB_BORDER;
        L.frame_stride = (long long)round_up((size_t)L.pitch * L.rows, 256);
        L.blur_stride = (long long)round_up((size_t)L.pitch * L.h, 256);
        CKC(dalloc(h, &L.img, (size_t)L.frame_stride * B));
        CKC(dalloc(h, &L.blur, (size_t)L.blur_stride * B));
        CKC(cudaMemset(L.img, 0, (size_t)L.frame_stride * B));
        L.scale = h->sf[l];
        L.patch_size = (float)(int)(31 * h->sf[l]);
        L.nfeat = nfeat[l];
        h->dump_off[l] = dump_total;
        dump_total += (long long)L.w * L.h;
        // resize tables
        if (l > 0) {
            const LevelDev &S = h->lv[l - 1];
            L.src_w = S.w; L.src_h = S.h;
            L.area2x = (S.w == 2 * L.w && S.h == 2 * L.h) ? 1 : 0;
            std::vector<int> xofs; std::vector<short2> xa; std::vector<int2> yr; std::vector<short2> yb;
            build_resize_tables(S.w, S.h, L.w, L.h, xofs, xa, yr, yb);
            h_xofs[l] = xofs; h_yrows[l] = yr;
            int *dxofs; short2 *dxa; int2 *dyr; short2 *dyb;
            CKC(upload(h, &dxofs, xofs)); CKC(upload(h, &dxa, xa)); CKC(upload(h, &dyr, yr)); CKC(upload(h, &dyb, yb));
            L.xofs = dxofs; L.xalpha = dxa; L.yrows = dyr; L.ybeta = dyb;
            // tables of the row-walking resize kernel: one entry per aligned 4-byte group of the padded row
            // (first group = byte 12, which holds padded columns -1..2) and one per padded row
            auto refl = [](int p, int len) { p = p < 0 ? -p : p; return p >= len ? 2 * len - 2 - p : p; };
            const int pw = L.w + 2 * ORBB_BORDER, ph = L.h + 2 * ORBB_BORDER;
            const int nq = (ORBB_ROI_X0 + L.w + ORBB_BORDER - 12 + 3) / 4;
            std::vector<uint4> rsh(2 * (size_t)nq);
            bool ok = true;
            for (int q = 0; q < nq; ++q) {
                int sxj[4]; unsigned wj[4];
                for (int j = 0; j < 4; ++j) {
                    const int px = std::min(std::max(12 + 4 * q + j - ORBB_PAD_X0, 0), pw - 1);
                    const int rx = refl(px - ORBB_BORDER, L.w);
                    if (L.area2x) { sxj[j] = 2 * rx; wj[j] = 1u | (1u << 16); }
                    else { sxj[j] = xofs[rx]; wj[j] = (unsigned)(uint16_t)xa[rx].x | ((unsigned)(uint16_t)xa[rx].y << 16); }
                }
                unsigned off[2], sel[2];
                for (int pr = 0; pr < 2; ++pr) {
                    const int s0 = sxj[2 * pr], s1 = sxj[2 * pr + 1];
                    const int base = std::min(s0, s1) & ~3;
                    const int n[4] = {s0 - base, s0 + 1 - base, s1 - base, s1 + 1 - base};
                    sel[pr] = 0;
                    for (int k = 0; k < 4; ++k) { if (n[k] > 7) ok = false; sel[pr] |= (unsigned)(n[k] & 7) << (4 * k); }
                    off[pr] = (unsigned)(ORBB_ROI_X0 + base);
                }
                rsh[2 * q] = make_uint4(off[0], sel[0], off[1], sel[1]);
                rsh[2 * q + 1] = make_uint4(wj[0], wj[1], wj[2], wj[3]);
            }
            std::vector<int4> rsv((size_t)ph);
            for (int py = 0; py < ph; ++py) {
                const int ry = refl(py - ORBB_BORDER, L.h);
                if (L.area2x) rsv[py] = make_int4(2 * ry, 2 * ry + 1, 0, 0);
                else rsv[py] = make_int4(yr[ry].x, yr[ry].y, (int)yb[ry].x << 16, (int)yb[ry].y << 16);
            }
            uint4 *drsh; int4 *drsv;
            CKC(upload(h, &drsh, rsh)); CKC(upload(h, &drsv, rsv));
            L.rs_h = drsh; L.rs_v = drsv; L.rs_nq = nq;
            L.rs_ok = (ok && !getenv("ORBB_RESIZE_TILED")) ? 1 : 0;
        }
        h->cell_first[l] = (int)cells.size();
        // ---- per-cell FAST grid (upstream ComputeKeyPointsOctTree)
        const int W = L.w - 32, H = L.h - 32;  // maxBorder - minBorder
        const float fw = (float)W, fh = (float)H;
        const int nCols = (int)(fw / 30.f), nRows = (int)(fh / 30.f);
        L.w_cell = (int)ceilf(fw / nCols);
        L.h_cell = (int)ceilf(fh / nRows);
        L.n_cell_x = L.n_cell_y = 0;
        int cap = 0;
        for (int i = 0; i < nRows; ++i) {
            const int y0 = ORBB_BORDER + i * L.h_cell, ch = std::min(L.h_cell, L.h - ORBB_BORDER - y0);
            if (ch <= 0) continue;
            L.n_cell_y = i + 1;
            for (int j = 0; j < nCols; ++j) {
                const int x0 = ORBB_BORDER + j * L.w_cell, cw = std::min(L.w_cell, L.w - ORBB_BORDER - x0);
                if (cw <= 0) continue;
                L.n_cell_x = std::max(L.n_cell_x, j + 1);
                CellEntry c{};
                c.level = (int16_t)l; c.x0 = (int16_t)x0; c.y0 = (int16_t)y0; c.cw = (int16_t)cw; c.ch = (int16_t)ch;
                c.inv_nux = (1u << 20) / (unsigned)((cw + 3) >> 2) + 1u;
                cells.push_back(c);
                max_cw = std::max(max_cw, cw); max_ch = std::max(max_ch, ch);
                cap += ((cw + 1) / 2) * ((ch + 1) / 2);  // strict 3x3 maxima: at most one per 2x2 block
            }
        }
        if (L.w_cell > 63 || L.h_cell > 63) { orbb_destroy(h); return ORBB_ERR_SHAPE; }
        L.cand_cap = (int)round_up((size_t)cap + 32, 64);
        if (L.cand_cap >= (1 << 22)) { orbb_destroy(h); return ORBB_ERR_SHAPE; }
        // ---- quadtree tables (upstream DistributeOctTree / DivideNode)
        L.n_ini = (int)roundf((float)W / H);
        if (L.n_ini < 1) { orbb_destroy(h); return ORBB_ERR_SHAPE; }
        const float hX = (float)W / L.n_ini;
        int root_bits = 0;
        while ((1 << root_bits) < L.n_ini) ++root_bits;
        int max_dim = H;
        for (int r = 0; r < L.n_ini; ++r) max_dim = std::max(max_dim, (int)(hX * (float)(r + 1)) - (int)(hX * (float)r));
        int D = 1;
        while ((1 << D) < max_dim) ++D;
        D += 1;  // one more level parts keys sitting on a root's right edge from its last column
        L.depth = D;
        L.key_bits = 2 * D + root_bits;
        if (L.key_bits > 32 || D > 15) { orbb_destroy(h); return ORBB_ERR_SHAPE; }
        std::vector<uint32_t> xkey(W), ykey(H);
        std::vector<uint16_t> xord(W), yord(H);
        for (int x = 0; x < W; ++x) {
            int root = (int)((float)x / hX);
            root = std::min(root, L.n_ini - 1);
            int lo = (int)(hX * (float)root), hi = (int)(hX * (float)(root + 1));
            uint32_t bits = 0;
            for (int d = 0; d < D; ++d) {
                const int half = (int)ceilf((float)(hi - lo) / 2);
                const int mid = lo + half;
                if (x < mid) { hi = mid; bits = bits << 1; }
                else { lo = mid; bits = (bits << 1) | 1u; }
            }
            xkey[x] = ((uint32_t)root << (2 * D)) | spread_bits(bits);
            const int t = std::max(x - 3, 0), j = t / L.w_cell;
            xord[x] = (uint16_t)((j << 6) | (t - j * L.w_cell));
        }
        for (int y = 0; y < H; ++y) {
            int lo = 0, hi = H;
            uint32_t bits = 0;
            for (int d = 0; d < D; ++d) {
                const int half = (int)ceilf((float)(hi - lo) / 2);
                const int mid = lo + half;
                if (y < mid) { hi = mid; bits = bits << 1; }
                else { lo = mid; bits = (bits << 1) | 1u; }
            }
            ykey[y] = spread_bits(bits) << 1;
            const int t = std::max(y - 3, 0), i = t / L.h_cell;
            yord[y] = (uint16_t)((i << 6) | (t - i * L.h_cell));
        }
        uint32_t *dxk, *dyk; uint16_t *dxo, *dyo;
        CKC(upload(h, &dxk, xkey)); CKC(upload(h, &dyk, ykey)); CKC(upload(h, &dxo, xord)); CKC(upload(h, &dyo, yord));
        L.xkey = dxk; L.ykey = dyk; L.xord = dxo; L.yord = dyo;
        L.sel_cap = std::max(L.nfeat, 4 * L.n_ini) + 8;
        const size_t cc = (size_t)L.cand_cap * B;
        CKC(dalloc(h, &L.cand, cc)); CKC(dalloc(h, &L.kv_a, cc)); CKC(dalloc(h, &L.kv_b, cc));
        CKC(dalloc(h, &L.sd, cc));
        CKC(dalloc(h, &L.sel, (size_t)L.sel_cap * B));
        // cell table of the quadtree fast path: the shallowest depth with >= 4 x quota cells (the selection stops
        // near the depth with ~quota nodes and may part one level deeper), at most 4096 cells and depth 6
        L.tbl_dc = 0; L.tbl_cells = 0;
        for (int dd = 1; dd <= std::min(D, 6) && (1 << (root_bits + 2 * dd)) <= 4096; ++dd) {
            L.tbl_dc = dd; L.tbl_cells = 1 << (root_bits + 2 * dd);
            if (L.tbl_cells >= 4 * L.nfeat) break;
        }
        if (L.tbl_cells) {
            CKC(dalloc(h, &L.tbl_cnt, (size_t)L.tbl_cells * B));
            CKC(dalloc(h, &L.tbl_best, (size_t)L.tbl_cells * B));
            CKC(cudaMemset(L.tbl_cnt, 0, sizeof(uint32_t) * (size_t)L.tbl_cells * B));
            CKC(cudaMemset(L.tbl_best, 0, sizeof(unsigned long long) * (size_t)L.tbl_cells * B));
        }
        slot_base[l] = (int)slot_level.size();
        for (int s = 0; s < L.sel_cap; ++s) slot_level.push_back(l);
        h->sel_cap_max = std::max(h->sel_cap_max, L.sel_cap);
    }
    h->dump_off[nl] = dump_total;
    h->n_cells = (int)cells.size(); h->n_slots = (int)slot_level.size();
    for (int l = nl; l <= ORBB_MAX_LEVELS; ++l) h->cell_first[l] = h->n_cells;
    h->max_kp = h->n_slots;
    h->pcap = h->sel_cap_max; h->pcap2 = 1;
    while (h->pcap2 < h->pcap) h->pcap2 <<= 1;
    if (octree_dyn_smem(h->sel_cap_max, h->pcap, h->pcap2) > 200 * 1024) { orbb_destroy(h); return ORBB_ERR_CAPACITY; }
    h->fcfg.tile_pitch = (int)round_up(max_cw + 12, 4) | 4;  // odd number of words: rows never share a bank pattern
    h->fcfg.tma_pitch = (int)round_up(max_cw + 10 + 15 + 4, 16);  // 16-byte aligned box start: up to 15 bytes of phase
    h->fcfg.tile_rows = max_ch + 6;
    h->fcfg.score_rows = max_ch + 2;
    h->fcfg.queue_len = (int)round_up((size_t)max_cw * max_ch, 8);
    // TMA staging of the FAST windows is opt-in (ORBB_FAST_TMA=1): measured on B200 it is not faster than the
    // manual aligned-load + funnel-shift staging (0.70 vs 0.66 ms per 256 frames; the 16-byte aligned box costs
    // 1.4 KB more shared memory per warp), see DESIGN.md.
    const bool want_tma = getenv("ORBB_FAST_TMA") && atoi(getenv("ORBB_FAST_TMA")) != 0;
    if (!want_tma) h->fcfg.tma_pitch = 0;
    // the score tile shares the window tile's pitch, so ONE queue entry (y * pitch + x) addresses both
    h->fcfg.score_pitch = std::max(h->fcfg.tile_pitch, h->fcfg.tma_pitch);
    h->fcfg.tile_bytes = (int)round_up((size_t)h->fcfg.score_pitch * h->fcfg.tile_rows, 16);
    h->fcfg.score_bytes = (int)round_up((size_t)h->fcfg.score_pitch * h->fcfg.score_rows, 16);
    h->fcfg.warp_bytes = (int)round_up((size_t)h->fcfg.tile_bytes + h->fcfg.score_bytes + 2 * (size_t)h->fcfg.queue_len, 128);  // 128-byte aligned TMA destination per warp
    if (want_tma && h->fcfg.tma_pitch <= 256 && h->fcfg.tile_rows <= 256) {
        std::vector<unsigned char> store(fast_tma_maps_bytes() + 128);
        void *hm = reinterpret_cast<void *>((reinterpret_cast<uintptr_t>(store.data()) + 63) & ~(uintptr_t)63);
        if (build_fast_tma_maps(hm, h->lv, nl, B, h->fcfg)) {
            unsigned char *dm = nullptr;
            CKC(dalloc(h, &dm, fast_tma_maps_bytes()));
            CKC(cudaMemcpy(dm, hm, fast_tma_maps_bytes(), cudaMemcpyHostToDevice));
            h->tma_maps = dm;
        }
    }
    {
        if (const char *e = getenv("ORBB_FUSED_PYRAMID")) h->fused_max_frames = std::max(atoi(e), 0);
        std::vector<std::vector<PfTile>> ft;
        if (h->fused_max_frames > 0 && build_fused_tiles(h, h_xofs, h_yrows, ft)) {
            for (size_t g = 0; g < ft.size(); ++g) CKC(upload(h, &h->fused[g].d_tiles, ft[g]));
        } else {
            h->fused.clear();
        }
    }
    CKC(upload(h, &h->d_cells, cells));
    CKC(upload(h, &h->d_slot_level, slot_level));
    CKC(upload(h, &h->d_slot_base, slot_base));
    {
        std::vector<int8_t> pat(k_pattern_host, k_pattern_host + 1024);
        CKC(upload(h, &h->d_pattern, pat));
        std::vector<LevelDev> lv(h->lv, h->lv + nl);
        CKC(upload(h, &h->d_levels, lv));
        CKC(upload(h, &h->d_dump_off, h->dump_off));
    }
    CKC(dalloc(h, &h->d_cand_count, (size_t)B * nl));
    CKC(dalloc(h, &h->d_sel_count, (size_t)B * nl));
    CKC(cudaMemset(h->d_cand_count, 0, sizeof(int) * (size_t)B * nl));
    CKC(cudaMemset(h->d_sel_count, 0, sizeof(int) * (size_t)B * nl));
    for (int i = 0; i < 2; ++i) {
        CKC(dalloc(h, &h->d_in2[i], (size_t)width * height * B));
        CKC(dalloc(h, &h->d_kp2[i], (size_t)h->max_kp * B));
        CKC(dalloc(h, &h->d_desc2[i], (size_t)h->max_kp * B * 32));
        CKC(dalloc(h, &h->d_counts2[i], (size_t)B));
        CKC(cudaEventCreateWithFlags(&h->ev_ticket[i], cudaEventDisableTiming));
        CKC(cudaEventCreateWithFlags(&h->ev_tail[i], cudaEventDisableTiming));
    }
    CKC(dalloc(h, &h->d_dump, (size_t)dump_total));
    // matcher / stereo scratch, allocated once (no call on the per-batch path allocates): split-T partial results for
    // the query sets this handle can produce (larger ones are chunked), SAD costs for max_batch/2 stereo pairs
    h->partial_cap = ORBB_MATCH_SPLIT_ROWS + std::max<size_t>((size_t)B * h->max_kp, (size_t)1 << 18);
    CKC(dalloc(h, &h->d_partial, h->partial_cap));
    if (matcher_kind() == 2 && matcher_pre()) {
        h->train_exp_rows = ORBB_MATCH_EXP_ROWS;
        CKC(dalloc(h, &h->d_train_exp, h->train_exp_rows * 256));
    }
    h->stereo_cap = (size_t)std::max(B / 2, 1) * h->max_kp;
    CKC(dalloc(h, &h->d_stereo_sad, h->stereo_cap));
    CKC(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
    CKC(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        CKC(cudaStreamCreateWithFlags(&h->s_comp[i], cudaStreamNonBlocking));
        CKC(cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
    }
    CKC(cudaEventCreateWithFlags(&h->ev_fence, cudaEventDisableTiming));
    CKC(cudaStreamCreateWithFlags(&h->s_side, cudaStreamNonBlocking));
    for (int i = 0; i < 7; ++i) {
        CKC(cudaStreamCreateWithFlags(&h->s_dev[i], cudaStreamNonBlocking));
        CKC(cudaEventCreateWithFlags(&h->ev_dev[i], cudaEventDisableTiming));
    }
    if (const char *e = getenv("ORBB_DEV_SPLIT")) h->dev_split = std::min(std::max(atoi(e), 1), 8);
    if (const char *e = getenv("ORBB_GRAPH")) h->use_graphs = atoi(e);
    CKC(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    h->ev_in.resize(ORBB_MAX_CHUNKS); h->ev_comp.resize(ORBB_MAX_CHUNKS);
    for (int i = 0; i < ORBB_MAX_CHUNKS; ++i) {
        CKC(cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
        CKC(cudaEventCreateWithFlags(&h->ev_comp[i], cudaEventDisableTiming));
    }
    CKC(cudaDeviceSynchronize());
#undef CKC
    *out = h;
    return ORBB_OK;
}

// ---------------------------------------------------------------- getters
extern "C" int orbb_get_levels(const orbb_handle *h) { return h ? h->nlevels : ORBB_ERR_INVALID; }

extern "C" int orbb_get_scale_factors(const orbb_handle *h, float *scale, float *inv_scale, float *sigma2,
                                      float *inv_sigma2) {
    if (!h) return ORBB_ERR_INVALID;
    for (int l = 0; l < h->nlevels; ++l) {
        const float s2 = l == 0 ? 1.0f : h->sf[l] * h->sf[l];
        if (scale) scale[l] = h->sf[l];
        if (inv_scale) inv_scale[l] = h->inv_sf[l];
        if (sigma2) sigma2[l] = s2;
        if (inv_sigma2) inv_sigma2[l] = 1.0f / s2;
    }
    return ORBB_OK;
}

extern "C" int orbb_get_features_per_level(const orbb_handle *h, int32_t *nfeat) {
    if (!h || !nfeat) return ORBB_ERR_INVALID;
    for (int l = 0; l < h->nlevels; ++l) nfeat[l] = h->lv[l].nfeat;
    return ORBB_OK;
}

extern "C" int orbb_max_keypoints_per_frame(const orbb_handle *h) { return h ? h->max_kp : ORBB_ERR_INVALID; }

extern "C" long long orbb_get_launch_count(const orbb_handle *h) { return h ? h->n_launches : -1; }

extern "C" int orbb_get_level(const orbb_handle *h, int frame, int level, orbb_level *out) {
    if (!h || !out || level < 0 || level >= h->nlevels || frame < 0 || frame >= h->max_batch) return ORBB_ERR_INVALID;
    const LevelDev &L = h->lv[level];
    out->width = L.w; out->height = L.h; out->pitch = L.pitch; out->roi_offset = ORBB_BORDER;
    out->padded = L.img + (size_t)frame * L.frame_stride + ORBB_PAD_X0;
    out->blurred = L.blur + (size_t)frame * L.blur_stride;
    out->scale = h->sf[level]; out->inv_scale = h->inv_sf[level]; out->nfeatures = L.nfeat;
    return ORBB_OK;
}

// ---------------------------------------------------------------- stages
// Every stage works on a frame range [f0, f0+n) of the resident batch so that the host entry point can
// pipeline chunks; the public stage calls use the whole batch of the last upload.
static int run_upload(orbb_handle *h, const uint8_t *d_images, size_t pitch, size_t stride, int f0, int n, cudaStream_t st) {
    CK(h, cudaMemsetAsync(h->d_cand_count + (size_t)f0 * h->nlevels, 0, sizeof(int) * (size_t)n * h->nlevels, st));
    CK(h, launch_level0(d_images, pitch, stride, h->lv[0], f0, n, st));
    h->n_launches += 1;
    return ORBB_OK;
}
static int run_pyramid(orbb_handle *h, int f0, int n, cudaStream_t st) {
    for (int l = 1; l < h->nlevels; ++l) CK(h, launch_resize(h->d_levels, h->lv, l, f0, n, st));
    h->n_launches += h->nlevels - 1;
    return ORBB_OK;
}
static int run_fast(orbb_handle *h, int f0, int n, cudaStream_t st) {
    CK(h, launch_fast(h->tma_maps, h->d_levels, h->d_cells, h->n_cells, h->nlevels, h->d_cand_count, h->t_lo, h->t_hi, h->fcfg, f0, n, st));
    h->n_launches += 1;
    return ORBB_OK;
}
static int run_distribute(orbb_handle *h, int f0, int n, cudaStream_t st) {
    static const bool split = getenv("ORBB_OCT_SPLIT") != nullptr;  // diagnostics: one launch per level
    if (split) {
        for (int l = 0; l < h->nlevels; ++l)
            CK(h, launch_octree(h->d_levels, h->nlevels, h->d_cand_count, h->d_sel_count, l, 1, f0, n, -1, h->sel_cap_max,
                                h->pcap, h->pcap2, st));
        h->n_launches += h->nlevels;
        return ORBB_OK;
    }
    CK(h, launch_octree(h->d_levels, h->nlevels, h->d_cand_count, h->d_sel_count, 0, h->nlevels, f0, n, -1,
                        h->sel_cap_max, h->pcap, h->pcap2, st));
    h->n_launches += 1;
    return ORBB_OK;
}
// FAST + quadtree restricted to levels [l0, l1): the cells of a level are contiguous and every (frame, level) has its
// own candidate list, counters and cell table, so disjoint level ranges may run concurrently on different streams
static int run_detect_levels(orbb_handle *h, int l0, int l1, int f0, int n, cudaStream_t st) {
    const int c0 = h->cell_first[l0], c1 = h->cell_first[l1];
    if (c1 > c0)
        CK(h, launch_fast(h->tma_maps, h->d_levels, h->d_cells + c0, c1 - c0, h->nlevels, h->d_cand_count, h->t_lo, h->t_hi, h->fcfg,
                          f0, n, st));
    CK(h, launch_octree(h->d_levels, h->nlevels, h->d_cand_count, h->d_sel_count, l0, l1 - l0, f0, n, -1, h->sel_cap_max,
                        h->pcap, h->pcap2, st));
    h->n_launches += 2;
    return ORBB_OK;
}
static int run_blur(orbb_handle *h, int f0, int n, cudaStream_t st) {
    CK(h, launch_blur(h->d_levels, h->lv, h->nlevels, f0, n, st));
    h->n_launches += 1;
    return ORBB_OK;
}
static int run_angle_orb(orbb_handle *h, int f0, int n, orbb_keypoint *d_kp, uint8_t *d_desc, int32_t *d_counts, int max_kp,
                         cudaStream_t st) {
    CK(h, launch_angle_orb(h->d_levels, h->nlevels, h->d_sel_count, h->d_pattern, h->d_slot_level, h->d_slot_base,
                           h->n_slots, f0, n, d_kp, d_desc, d_counts, max_kp, st));
    h->n_launches += 1;
    return ORBB_OK;
}
static int run_all(orbb_handle *h, const uint8_t *d_images, size_t pitch, size_t stride, int f0, int n, orbb_keypoint *d_kp,
                   uint8_t *d_desc, int32_t *d_counts, int max_kp, cudaStream_t st, cudaStream_t side = nullptr,
                   cudaEvent_t ev_fork = nullptr, cudaEvent_t ev_join = nullptr) {
    int rc;
    // Small batches (a lone frame is the reference's operating mode): the chain level0 -> 7 resizes -> FAST -> quadtree
    // -> angle/rBRIEF is latency bound end to end (each kernel a few us on a fraction of the SMs).  Detection of a level
    // needs only that level: FAST + quadtree of level 0 -- the longest of them -- move to a side stream right after the
    // level-0 kernel, those of levels 1-3 to a second one once level 3 exists, and both run NEXT TO the rest of the pyramid
    // and the detection of the small levels; the blur (needs the pyramid only) runs on another stream.  Same kernels,
    // same results.  Only as part of a captured graph: with plain stream operations the cross-stream joins cost more
    // than the overlap gains.  One 848x480 frame, 8 levels, inside the graph: one chain 65 us, groups {0-3}{4-7} 55 us,
    // {0}{1-3}{4-7} 47 us (finer splits change nothing: the level-0 chain level0 -> FAST -> quadtree -> angle/rBRIEF is
    // what remains, profiles/r02b_single_frame.txt).  ORBB_LONE_GROUPS="b0,b1,.." overrides the group boundaries.
    static const int lone_split = getenv("ORBB_LONE_SPLIT") ? atoi(getenv("ORBB_LONE_SPLIT")) : 4;
    if (side && h->use_graphs && n <= lone_split && h->nlevels >= 4 && (h->fused.empty() || n > h->fused_max_frames)) {
        // level groups [0, b0), [b0, b1), ..., [b_last, nlevels): every group but the last detects on its own side stream as
        // soon as its levels exist; the last one stays on `st` behind the rest of the pyramid
        int bounds[4], nb = 0;
        {
            const char *e = getenv("ORBB_LONE_GROUPS");  // e.g. "1,4": level 0 alone, levels 1-3, the rest
            std::string spec = e ? e : "1,4";
            for (size_t pos = 0; pos < spec.size() && nb < 3;) {
                const int v = atoi(spec.c_str() + pos);
                if (v > (nb ? bounds[nb - 1] : 0) && v < h->nlevels) bounds[nb++] = v;
                const size_t c = spec.find(',', pos);
                if (c == std::string::npos) break;
                pos = c + 1;
            }
            if (nb == 0) bounds[nb++] = 1;
        }
        if ((rc = run_upload(h, d_images, pitch, stride, f0, n, st))) return rc;
        int built = 1, g0 = 0;  // levels [0, built) exist
        for (int gi = 0; gi < nb; ++gi) {
            for (; built < bounds[gi]; ++built) CK(h, launch_resize(h->d_levels, h->lv, built, f0, n, st));
            cudaStream_t sg = h->s_dev[gi];
            CK(h, cudaEventRecord(h->ev_dev[2 * gi], st));
            CK(h, cudaStreamWaitEvent(sg, h->ev_dev[2 * gi], 0));
            if ((rc = run_detect_levels(h, g0, bounds[gi], f0, n, sg))) return rc;
            CK(h, cudaEventRecord(h->ev_dev[2 * gi + 1], sg));
            g0 = bounds[gi];
        }
        for (; built < h->nlevels; ++built) CK(h, launch_resize(h->d_levels, h->lv, built, f0, n, st));
        h->n_launches += h->nlevels - 1;
        CK(h, cudaEventRecord(ev_fork, st));
        CK(h, cudaStreamWaitEvent(side, ev_fork, 0));
        if ((rc = run_blur(h, f0, n, side))) return rc;
        CK(h, cudaEventRecord(ev_join, side));
        if ((rc = run_detect_levels(h, g0, h->nlevels, f0, n, st))) return rc;
        for (int gi = 0; gi < nb; ++gi) CK(h, cudaStreamWaitEvent(st, h->ev_dev[2 * gi + 1], 0));
        CK(h, cudaStreamWaitEvent(st, ev_join, 0));
        return run_angle_orb(h, f0, n, d_kp, d_desc, d_counts, max_kp, st);
    }
    if (!h->fused.empty() && n <= h->fused_max_frames) {
        // small batch: level 0 and every further level in ceil(nlevels / 4) launches (k_pyramid_fused)
        CK(h, cudaMemsetAsync(h->d_cand_count + (size_t)f0 * h->nlevels, 0, sizeof(int) * (size_t)n * h->nlevels, st));
        for (const auto &G : h->fused)
            CK(h, launch_pyramid_fused(d_images, pitch, stride, h->d_levels, G.g0, G.ng, G.d_tiles, G.n_tiles, G.smem, f0, n, st));
        h->n_launches += (long long)h->fused.size();
    } else {
        if ((rc = run_upload(h, d_images, pitch, stride, f0, n, st))) return rc;
        if ((rc = run_pyramid(h, f0, n, st))) return rc;
    }
    if ((rc = run_fast(h, f0, n, st))) return rc;
    // The blur depends only on the pyramid.  FAST saturates the issue slots, the quadtree kernel does not
    // (barrier/latency bound), so with a side stream the blur is forked to run under the quadtree.
    if (side) {
        CK(h, cudaEventRecord(ev_fork, st));
        CK(h, cudaStreamWaitEvent(side, ev_fork, 0));
        if ((rc = run_blur(h, f0, n, side))) return rc;
        CK(h, cudaEventRecord(ev_join, side));
    }
    if ((rc = run_distribute(h, f0, n, st))) return rc;
    if (side) CK(h, cudaStreamWaitEvent(st, ev_join, 0));
    else if ((rc = run_blur(h, f0, n, st))) return rc;
    return run_angle_orb(h, f0, n, d_kp, d_desc, d_counts, max_kp, st);
}

extern "C" int orbb_stage_upload(orbb_handle *h, const uint8_t *d_images, size_t pitch, size_t frame_stride,
                                 int n_frames, void *stream) {
    if (!h || !d_images || n_frames < 1 || pitch < (size_t)h->w) return ORBB_ERR_INVALID;
    if (n_frames > 1 && frame_stride < pitch * (size_t)h->h) return ORBB_ERR_INVALID;  // frames would overlap
    if (n_frames > h->max_batch) return ORBB_ERR_CAPACITY;
    CK(h, cudaSetDevice(h->device));
    h->n_frames_last = n_frames;
    return run_upload(h, d_images, pitch, frame_stride, 0, n_frames, static_cast<cudaStream_t>(stream));
}

extern "C" int orbb_pyramid_create_levels(orbb_handle *h, void *stream) {
    if (!h || h->n_frames_last < 1) return ORBB_ERR_INVALID;
    return run_pyramid(h, 0, h->n_frames_last, static_cast<cudaStream_t>(stream));
}

extern "C" int orbb_detect_fast(orbb_handle *h, void *stream) {
    if (!h || h->n_frames_last < 1) return ORBB_ERR_INVALID;
    // re-runnable on the resident batch: the candidate lists restart from zero (the whole-extractor path clears
    // them in run_upload instead, which keeps the pyramid -> FAST chain free of a memset node).  The quadtree cell
    // tables are cleared by the quadtree kernel that consumed them; a second detect_fast WITHOUT a distribute in
    // between leaves them double-counted, which the quadtree kernel detects (table total != list length) and answers
    // with its general sorted-key path -- same selection.
    CK(h, cudaMemsetAsync(h->d_cand_count, 0, sizeof(int) * (size_t)h->n_frames_last * h->nlevels, static_cast<cudaStream_t>(stream)));
    return run_fast(h, 0, h->n_frames_last, static_cast<cudaStream_t>(stream));
}

extern "C" int orbb_detect_distribute(orbb_handle *h, void *stream) {
    if (!h || h->n_frames_last < 1) return ORBB_ERR_INVALID;
    return run_distribute(h, 0, h->n_frames_last, static_cast<cudaStream_t>(stream));
}

extern "C" int orbb_detect(orbb_handle *h, void *stream) {
    const int rc = orbb_detect_fast(h, stream);
    return rc ? rc : orbb_detect_distribute(h, stream);
}

extern "C" int orbb_gaussian_blur(orbb_handle *h, void *stream) {
    if (!h || h->n_frames_last < 1) return ORBB_ERR_INVALID;
    return run_blur(h, 0, h->n_frames_last, static_cast<cudaStream_t>(stream));
}

extern "C" int orbb_compute_angle_and_orb(orbb_handle *h, orbb_keypoint *d_kp, uint8_t *d_desc, int32_t *d_counts,
                                          int max_kp, void *stream) {
    if (!h || !d_kp || !d_desc || !d_counts || max_kp < 1 || h->n_frames_last < 1) return ORBB_ERR_INVALID;
    return run_angle_orb(h, 0, h->n_frames_last, d_kp, d_desc, d_counts, max_kp, static_cast<cudaStream_t>(stream));
}

extern "C" int orbb_compute_fast_angle(orbb_handle *h, float *d_angle, const float *d_pos_xy, const uint8_t *d_image,
                                       int image_pitch, int image_width, int image_height, int keypoints_num, void *stream) {
    if (!h || !d_angle || !d_pos_xy || !d_image || keypoints_num < 0 || image_width < 1 || image_height < 1 ||
        image_pitch < image_width || (reinterpret_cast<uintptr_t>(d_pos_xy) & 7))
        return ORBB_ERR_INVALID;
    CK(h, cudaSetDevice(h->device));
    CK(h, launch_fast_angle_pos(d_angle, d_pos_xy, d_image, image_pitch, image_width, image_height, keypoints_num,
                                static_cast<cudaStream_t>(stream)));
    h->n_launches += keypoints_num > 0;
    return ORBB_OK;
}

extern "C" int orbb_calc_orb(orbb_handle *h, const float *d_angle, const float *d_pos_xy, uint8_t *d_desc,
                             const uint8_t *d_blurred_image, int image_pitch, int image_width, int image_height,
                             int keypoints_num, void *stream) {
    if (!h || !d_angle || !d_pos_xy || !d_desc || !d_blurred_image || keypoints_num < 0 || image_width < 1 ||
        image_height < 1 || image_pitch < image_width || (reinterpret_cast<uintptr_t>(d_pos_xy) & 7))
        return ORBB_ERR_INVALID;
    CK(h, cudaSetDevice(h->device));
    CK(h, launch_calc_orb_pos(d_angle, d_pos_xy, d_desc, d_blurred_image, image_pitch, image_width, image_height,
                              keypoints_num, h->d_pattern, static_cast<cudaStream_t>(stream)));
    h->n_launches += keypoints_num > 0;
    return ORBB_OK;
}

extern "C" int orbb_detect_export(orbb_handle *h, float *d_pos_xy, float *d_score, int32_t *d_level, int32_t *d_level_counts,
                                  int32_t *d_counts, int max_kp, void *stream) {
    if (!h || max_kp < 1 || h->n_frames_last < 1 || (reinterpret_cast<uintptr_t>(d_pos_xy) & 7)) return ORBB_ERR_INVALID;
    CK(h, cudaSetDevice(h->device));
    CK(h, launch_detect_export(h->d_levels, h->nlevels, h->d_sel_count, h->d_slot_level, h->d_slot_base, h->n_slots,
                               h->n_frames_last, d_pos_xy, d_score, d_level, d_level_counts, d_counts, max_kp,
                               static_cast<cudaStream_t>(stream)));
    h->n_launches += 1;
    return ORBB_OK;
}

extern "C" int orbb_extract_batch_device(orbb_handle *h, const uint8_t *d_images, size_t pitch, size_t frame_stride,
                                         int n_frames, orbb_keypoint *d_kp, uint8_t *d_desc, int32_t *d_counts,
                                         int max_kp, void *stream) {
    if (!h || !d_images || !d_kp || !d_desc || !d_counts || max_kp < 1 || n_frames < 1 || pitch < (size_t)h->w)
        return ORBB_ERR_INVALID;
    if (n_frames > 1 && frame_stride < pitch * (size_t)h->h) return ORBB_ERR_INVALID;  // frames would overlap
    if (n_frames > h->max_batch) return ORBB_ERR_CAPACITY;
    CK(h, cudaSetDevice(h->device));
    h->n_frames_last = n_frames;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n_frames < 64) {
        if (!h->use_graphs)
            return run_all(h, d_images, pitch, frame_stride, 0, n_frames, d_kp, d_desc, d_counts, max_kp, st, h->s_side,
                           h->ev_fork, h->ev_join);
        // Small batches are launch / latency bound (a lone frame: ~17 stream operations for ~60 us of GPU work), so the
        // whole stage sequence -- memset, pyramid chain, detection of the big levels forked next to the rest of the
        // pyramid, blur on a third stream, angle/rBRIEF after the joins -- is captured once per argument set into a
        // CUDA graph and replayed.  A caller that cycles through a few buffer sets (the reference's slots own fixed
        // buffers, buildStream.cpp:208-341) hits the 4-entry cache; a miss re-captures and patches the evicted
        // executable graph in place (cudaGraphExecUpdate: same topology).  Eight misses in a row mean the caller passes
        // fresh pointers every call: the handle then goes back to plain stream launches for good.  Measured on B200,
        // one 848x480 frame: 8 levels / 1200 kp 65 -> 58 us per call (53 back to back), 1 level / 405 kp 45 -> 38 us
        // (34 back to back), host time inside the call 34 -> 5 us.  ORBB_GRAPH=0 disables.
        orbb_handle::GraphSlot *slot = nullptr, *victim = &h->graphs[0];
        for (auto &g : h->graphs) {
            if (g.exec && g.img == d_images && g.pitch == pitch && g.stride == frame_stride && g.n == n_frames &&
                g.kp == d_kp && g.desc == d_desc && g.cnt == d_counts && g.max_kp == max_kp)
                slot = &g;
            if (g.stamp < victim->stamp) victim = &g;
        }
        if (slot) h->graph_miss_streak = 0;
        else if (++h->graph_miss_streak > 8 && h->graphs[0].exec) {
            h->use_graphs = 0;
            return run_all(h, d_images, pitch, frame_stride, 0, n_frames, d_kp, d_desc, d_counts, max_kp, st, h->s_side,
                           h->ev_fork, h->ev_join);
        }
        if (!slot) {
            slot = victim;
            cudaStream_t cs = h->s_comp[1];  // capture on an internal stream: the caller's may be the legacy stream
            const long long l0 = h->n_launches;
            CK(h, cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
            const int rc = run_all(h, d_images, pitch, frame_stride, 0, n_frames, d_kp, d_desc, d_counts, max_kp, cs,
                                   h->s_side, h->ev_fork, h->ev_join);
            cudaGraph_t graph = nullptr;
            const cudaError_t ce = cudaStreamEndCapture(cs, &graph);
            slot->launches = h->n_launches - l0;
            h->n_launches = l0;
            if (rc || ce != cudaSuccess) {
                if (graph) cudaGraphDestroy(graph);
                if (slot->exec) { cudaGraphExecDestroy(slot->exec); slot->exec = nullptr; }
                if (rc) return rc;
                CK(h, ce);
            }
            cudaError_t ie = cudaErrorUnknown;
            if (slot->exec && slot->n == n_frames) {  // same node topology: patch the kernel / memset arguments in place
                cudaGraphExecUpdateResultInfo info;
                ie = cudaGraphExecUpdate(slot->exec, graph, &info);
                if (ie != cudaSuccess) cudaGetLastError();
            }
            if (ie != cudaSuccess) {
                if (slot->exec) { cudaGraphExecDestroy(slot->exec); slot->exec = nullptr; }
                ie = cudaGraphInstantiate(&slot->exec, graph, 0);
            }
            cudaGraphDestroy(graph);
            CK(h, ie);
            slot->img = d_images; slot->pitch = pitch; slot->stride = frame_stride; slot->n = n_frames;
            slot->kp = d_kp; slot->desc = d_desc; slot->cnt = d_counts; slot->max_kp = max_kp;
        }
        slot->stamp = ++h->graph_clock;
        CK(h, cudaGraphLaunch(slot->exec, st));
        h->n_launches += slot->launches;
        return ORBB_OK;
    }
    // Large batches: parts of the batch on separate streams.  FAST saturates the issue slots while the quadtree
    // kernel is latency bound, so the parts interleave (one part's quadtree runs under another part's FAST).
    const int parts = std::max(1, std::min(h->dev_split, n_frames / 32));
    CK(h, cudaEventRecord(h->ev_fence, st));
    for (int k = 0, f0 = 0; k < parts; ++k) {
        const int n = (n_frames - f0) / (parts - k);
        cudaStream_t sk = k == 0 ? st : h->s_dev[k - 1];
        if (k) CK(h, cudaStreamWaitEvent(sk, h->ev_fence, 0));
        const int rc = run_all(h, d_images + frame_stride * f0, pitch, frame_stride, f0, n, d_kp, d_desc, d_counts, max_kp, sk);
        if (rc) return rc;
        if (k) {
            CK(h, cudaEventRecord(h->ev_dev[k - 1], sk));
            CK(h, cudaStreamWaitEvent(st, h->ev_dev[k - 1], 0));
        }
        f0 += n;
    }
    return ORBB_OK;
}

// Internal (same library, orbb_stage.cu)
namespace orbb {
// a replay of a graph that holds entry points of this handle: the handle's bookkeeping follows
void note_replay(orbb_handle *h, int n_frames, long long launches) {
    h->n_frames_last = n_frames;
    h->n_launches += launches;
}
}  // namespace orbb

// Host entry points: chunks of the batch flow through the handle's streams (H2D copy stream -> two alternating
// compute streams -> D2H copy stream) linked by events, so PCIe transfers overlap the kernels; staging buffers
// are double-buffered by ticket parity so the next batch's H2D runs under the current batch's kernels.
extern "C" int orbb_wait(orbb_handle *h, int ticket) {
    if (!h || ticket < 0 || ticket >= h->n_submitted) return ORBB_ERR_INVALID;
    if (ticket < h->n_submitted - 2) return ORBB_OK;  // older than the ring: already waited for at submit time
    CK(h, cudaEventSynchronize(h->ev_ticket[ticket & 1]));
    return ORBB_OK;
}

// latency_mode (the blocking call): chunk sizes taper towards the end of the batch (see the schedule below), so that
// little work is left once the last copy has landed.  Throughput mode (async API, batches back to back): four quarter-batch chunks (>= 32 frames each,
// ORBB_HOST_PARTS overrides) -- the next batch's H2D already runs under this batch's kernels.  The path is bound by the
// PCIe copies (1.43 ms of H2D against 1.22 ms of kernels per 256 frames of 640x480), so what chunking decides is the
// pipeline's fill and drain: with halves a run of 20 batches reached 95 % of the box's measured copy bound, with
// quarters 99-100 % (165 k -> 172 k frames/s; six or eight chunks measure the same).
static int submit_host(orbb_handle *h, const uint8_t *h_images, size_t pitch, size_t frame_stride, int n_frames,
                       orbb_keypoint *h_kp, uint8_t *h_desc, int32_t *h_counts, int max_kp, void *stream,
                       bool latency_mode) {
    if (!h || !h_images || !h_kp || !h_desc || !h_counts || max_kp < 1 || pitch < (size_t)h->w) return ORBB_ERR_INVALID;
    if (n_frames < 1 || n_frames > h->max_batch) return ORBB_ERR_CAPACITY;
    if (n_frames > 1 && frame_stride < pitch * (size_t)h->h) return ORBB_ERR_INVALID;  // frames would overlap
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(h, cudaSetDevice(h->device));
    const int ticket = (int)h->n_submitted, par = ticket & 1;
    if (ticket >= 2) CK(h, cudaEventSynchronize(h->ev_ticket[par]));  // batch ticket-2 owned this parity's buffers
    h->n_frames_last = n_frames;
    const size_t fw = (size_t)h->w, fsz = fw * h->h;
    const int mk = std::min(max_kp, h->max_kp);
    uint8_t *d_in = h->d_in2[par];
    orbb_keypoint *d_kp = h->d_kp2[par];
    uint8_t *d_desc = h->d_desc2[par];
    int *d_counts = h->d_counts2[par];
    // everything already queued on the caller's stream happens-before the pipeline.  The level / scratch buffers are
    // per frame slot and shared by consecutive batches: when this batch is chunked exactly like the previous one,
    // every frame slot stays on the same compute stream and stream order is enough; otherwise each compute stream
    // also waits for the other one's previous tail.
    const bool same_layout = ticket >= 1 && h->last_host_frames == n_frames && h->last_host_latency == (int)latency_mode;
    h->last_host_frames = n_frames; h->last_host_latency = (int)latency_mode;
    CK(h, cudaEventRecord(h->ev_fence, st));
    CK(h, cudaStreamWaitEvent(h->s_in, h->ev_fence, 0));
    for (int i = 0; i < 2; ++i) {
        CK(h, cudaStreamWaitEvent(h->s_comp[i], h->ev_fence, 0));
        if (ticket >= 1 && !same_layout) CK(h, cudaStreamWaitEvent(h->s_comp[i], h->ev_tail[i ^ 1], 0));
    }
    static const int host_parts = getenv("ORBB_HOST_PARTS") ? std::max(1, atoi(getenv("ORBB_HOST_PARTS"))) : 4;
    // chunk sizes.  Throughput mode: equal parts.  Latency mode (a lone blocking batch): the copies are the longer
    // pipeline (1.43 ms of H2D vs ~1.3 ms of kernels per 256 frames), so the call ends one chunk's kernels + D2H after the
    // LAST copy -- the schedule therefore TAPERS: ..., 64, 64, 48, 32, 16 from the end, the first chunk takes the rest.
    // (Round 1 grew the chunks 16, 32, 64, 64, ...: with the kernels as fast as the copies now, its last 64-frame chunk
    // was still computing 0.3 ms after the last copy: 134 k -> 140 k frames/s for blocking 256-frame calls.)
    int sizes[ORBB_MAX_CHUNKS], n_chunks = 0;
    if (n_frames <= 32) sizes[n_chunks++] = n_frames;
    else if (!latency_mode) {
        const int per = std::max(32, (n_frames + host_parts - 1) / host_parts);
        for (int left = n_frames; left > 0;) {
            const int n = n_chunks == ORBB_MAX_CHUNKS - 1 ? left : std::min(per, left);
            sizes[n_chunks++] = n; left -= n;
        }
    } else {
        int tail[ORBB_MAX_CHUNKS], nt = 0, left = n_frames;
        for (int want = 16; left > 0 && nt < ORBB_MAX_CHUNKS - 1; want = std::min(want + 16, 64)) {
            const int n = left - want < 16 ? left : want;  // never leave a sliver for the first chunk
            tail[nt++] = n; left -= n;
        }
        if (left > 0) tail[nt++] = left;
        for (int i = nt - 1; i >= 0; --i) sizes[n_chunks++] = tail[i];
    }
    for (int c = 0, f0 = 0; c < n_chunks; ++c) {
        const int n = sizes[c];
        cudaStream_t sc = h->s_comp[c & 1];
        uint8_t *din = d_in + fsz * f0;
        const uint8_t *src = h_images + frame_stride * f0;
        if (pitch == fw && frame_stride == fsz) {
            CK(h, cudaMemcpyAsync(din, src, fsz * n, cudaMemcpyHostToDevice, h->s_in));
        } else if (frame_stride == pitch * (size_t)h->h) {
            CK(h, cudaMemcpy2DAsync(din, fw, src, pitch, fw, (size_t)h->h * n, cudaMemcpyHostToDevice, h->s_in));
        } else {
            for (int f = 0; f < n; ++f)
                CK(h, cudaMemcpy2DAsync(din + fsz * f, fw, src + frame_stride * f, pitch, fw, h->h, cudaMemcpyHostToDevice,
                                        h->s_in));
        }
        CK(h, cudaEventRecord(h->ev_in[c], h->s_in));
        CK(h, cudaStreamWaitEvent(sc, h->ev_in[c], 0));
        int rc = run_all(h, din, fw, fsz, f0, n, d_kp, d_desc, d_counts, mk, sc);
        if (rc) return rc;
        CK(h, cudaEventRecord(h->ev_comp[c], sc));
        CK(h, cudaStreamWaitEvent(h->s_out, h->ev_comp[c], 0));
        CK(h, cudaMemcpyAsync(h_counts + f0, d_counts + f0, sizeof(int) * n, cudaMemcpyDeviceToHost, h->s_out));
        CK(h, cudaMemcpy2DAsync(h_kp + (size_t)f0 * max_kp, sizeof(orbb_keypoint) * max_kp, d_kp + (size_t)f0 * mk,
                                sizeof(orbb_keypoint) * mk, sizeof(orbb_keypoint) * mk, n, cudaMemcpyDeviceToHost, h->s_out));
        CK(h, cudaMemcpy2DAsync(h_desc + 32 * (size_t)f0 * max_kp, 32 * (size_t)max_kp, d_desc + 32 * (size_t)f0 * mk,
                                32 * (size_t)mk, 32 * (size_t)mk, n, cudaMemcpyDeviceToHost, h->s_out));
        f0 += n;
    }
    CK(h, cudaEventRecord(h->ev_tail[0], h->s_comp[0]));
    CK(h, cudaEventRecord(h->ev_tail[1], h->s_comp[1]));
    // completion is observed through orbb_wait (the caller's stream is NOT made to wait: that would serialise
    // consecutive batches through the fence above)
    CK(h, cudaEventRecord(h->ev_ticket[par], h->s_out));
    h->n_submitted++;
    return ticket;
}

extern "C" int orbb_extract_batch_host_async(orbb_handle *h, const uint8_t *h_images, size_t pitch, size_t frame_stride,
                                             int n_frames, orbb_keypoint *h_kp, uint8_t *h_desc, int32_t *h_counts,
                                             int max_kp, void *stream) {
    return submit_host(h, h_images, pitch, frame_stride, n_frames, h_kp, h_desc, h_counts, max_kp, stream, false);
}

extern "C" int orbb_extract_batch_host(orbb_handle *h, const uint8_t *h_images, size_t pitch, size_t frame_stride,
                                       int n_frames, orbb_keypoint *h_kp, uint8_t *h_desc, int32_t *h_counts,
                                       int max_kp, void *stream) {
    const int ticket = submit_host(h, h_images, pitch, frame_stride, n_frames, h_kp, h_desc, h_counts, max_kp, stream, true);
    if (ticket < 0) return ticket;
    return orbb_wait(h, ticket);
}

// ---------------------------------------------------------------- matcher
// The split-T scratch (ORBB_MATCH_SPLIT_ROWS + the handle's own keypoint capacity) is allocated in orbb_create; larger
// query sets are processed in query chunks that fit, so no call allocates or synchronises.  The scratch is per
// handle: matcher calls on one handle must be issued on one stream (or be ordered by the caller).

// Split of the train range across blockIdx.y.  Measured on B200: the kernel keeps gaining until about 64 CTAs per SM
// are queued (796 -> 886 Gpairs/s for 257 k x 50 k): small CTAs even out the tail and keep every SM's POPC pipe fed.
static int pick_split(int qblocks_total, long long nt) {
    // The tcgen05 kernel wants LONG scans instead: a CTA pays for expanding its 256 queries into shared memory once, and its
    // epilogue skips every 32-column chunk that cannot improve a row's best -- which is most of them only after the
    // first ~1000 columns of a scan.  A few CTAs per SM still even out the tail.
    // (Measured, 257 k x 50 k / 100 k x 200 k, 1-NN: 64 CTAs per SM 4197 / 4428 Gpairs/s, 16: 4530 / 4649, 4: 4629 / 4287 --
    // the last one is a 5.3-wave tail.)  So: the FEWEST splits whose CTA count fills its last wave of one-CTA-per-SM to
    // 95 %, and at least four CTAs per SM queued; profiles/r02i_umma_split_sweep.txt.
    static const int per_sm_env = getenv("ORBB_MATCH_CTAS") ? std::max(atoi(getenv("ORBB_MATCH_CTAS")), 1) : 0;
    const int per_sm = per_sm_env ? per_sm_env : (matcher_kind() == 2 ? 4 : 64);
    int want = (per_sm * 148 + qblocks_total - 1) / std::max(qblocks_total, 1);
    if (matcher_kind() == 2 && !per_sm_env) {
        for (int s = want; s <= std::min(want + 8, 64); ++s) {
            const long long ctas = (long long)std::max(qblocks_total, 1) * s, waves = (ctas + 147) / 148;
            if (ctas * 100 >= waves * 148 * 95) { want = s; break; }
        }
    }
    want = (int)std::min<long long>(want, std::max<long long>(1, nt / 128));  // at least one shared-memory tile of train rows per split
    want = std::max(want, (int)((nt + (1 << 22) - 1) >> 22));  // packed keys hold 22 index bits per split
    return std::max(1, std::min(want, 64));
}

extern "C" int orbb_match_knn(orbb_handle *h, const uint8_t *d_query, int nq, const uint8_t *d_train, int nt, int k,
                              float ratio, int32_t *d_idx, int32_t *d_dist, uint8_t *d_accept, int32_t *d_naccept,
                              void *stream) {
    if (!h || !d_query || !d_train || !d_idx || !d_dist || nq < 0 || nt < 0 || k < 1 || k > 2) return ORBB_ERR_INVALID;
    if ((reinterpret_cast<uintptr_t>(d_query) | reinterpret_cast<uintptr_t>(d_train)) & 15) return ORBB_ERR_INVALID;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(h, cudaSetDevice(h->device));
    if (d_naccept) CK(h, cudaMemsetAsync(d_naccept, 0, sizeof(int), st));
    if (nq == 0) return ORBB_OK;
    if (nt > 64 * (1 << 22)) return ORBB_ERR_CAPACITY;  // 268 M train rows per call
    const int max_chunk = (int)std::min<size_t>((h->partial_cap - ORBB_MATCH_SPLIT_ROWS) & ~(size_t)255, (size_t)1 << 30);
    for (int q0 = 0; q0 < nq;) {
        int n = std::min(nq - q0, max_chunk);
        int n_split = pick_split((n + 255) / 256, nt);
        if ((size_t)n_split * n > h->partial_cap) {  // only when a huge train set forces many splits
            n = (int)((h->partial_cap / n_split) & ~(size_t)255);
            n_split = pick_split((n + 255) / 256, nt);
        }
        CK(h, launch_match(d_query + 32 * (size_t)q0, d_train, nullptr, nullptr, 1, n, n, nt, n_split, h->d_partial, n, k, ratio,
                           d_idx + 2 * (size_t)q0, d_dist + 2 * (size_t)q0, d_accept ? d_accept + q0 : nullptr, d_naccept, st,
                           nullptr, 0, h->d_train_exp, h->train_exp_rows, q0 == 0, &h->n_launches));
        q0 += n;
    }
    return ORBB_OK;
}

// Batch form for the map-matching use (BASELINE cfg 5): the query sets are the frames of an extraction output
// ([n_frames][max_kp][32] with device-side counts), all against ONE train set.  blockIdx.z = frame; rows past a
// frame's count report "no match" (-1), so the outputs keep the fixed [n_frames][max_kp] stride of the inputs and
// can be gathered across GPUs without a compaction step or a host round trip.
extern "C" int orbb_match_knn_batch(orbb_handle *h, const uint8_t *d_query, const int32_t *d_q_counts, int n_frames,
                                    int max_kp, const uint8_t *d_train, int nt, int k, float ratio, int32_t *d_idx,
                                    int32_t *d_dist, uint8_t *d_accept, int32_t *d_naccept, void *stream) {
    if (!h || !d_query || !d_q_counts || !d_train || !d_idx || !d_dist || n_frames < 0 || max_kp < 1 || nt < 0 || k < 1 ||
        k > 2)
        return ORBB_ERR_INVALID;
    if ((reinterpret_cast<uintptr_t>(d_query) | reinterpret_cast<uintptr_t>(d_train)) & 15) return ORBB_ERR_INVALID;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(h, cudaSetDevice(h->device));
    if (d_naccept) CK(h, cudaMemsetAsync(d_naccept, 0, sizeof(int), st));
    if (n_frames == 0) return ORBB_OK;
    if (nt > 64 * (1 << 22)) return ORBB_ERR_CAPACITY;
    const int qb = (max_kp + 255) / 256;
    for (int f0 = 0; f0 < n_frames;) {  // frame chunks: gridDim.z <= 65535 and n_split x rows within the scratch
        int n = std::min(n_frames - f0, 65535);
        int n_split = pick_split(qb * n, nt);
        if ((size_t)n_split * n * max_kp > h->partial_cap) {
            n = (int)std::min<size_t>(n, std::max<size_t>(h->partial_cap / ((size_t)n_split * max_kp), 1));
            n_split = pick_split(qb * n, nt);
            if ((size_t)n_split * n * max_kp > h->partial_cap) return ORBB_ERR_CAPACITY;  // a single frame does not fit
        }
        const size_t r0 = (size_t)f0 * max_kp;
        CK(h, launch_match(d_query + 32 * r0, d_train, nullptr, nullptr, n, n * max_kp, max_kp, nt, n_split, h->d_partial,
                           n * max_kp, k, ratio, d_idx + 2 * r0, d_dist + 2 * r0, d_accept ? d_accept + r0 : nullptr,
                           d_naccept, st, d_q_counts + f0, max_kp, h->d_train_exp, h->train_exp_rows, f0 == 0, &h->n_launches));
        f0 += n;
    }
    return ORBB_OK;
}

extern "C" int orbb_match_knn_segmented(orbb_handle *h, const uint8_t *d_query, const int32_t *d_q_offsets,
                                        const uint8_t *d_train, const int32_t *d_t_offsets, int nseg, int nq_total,
                                        int max_q_per_seg, int max_t_per_seg, int k, float ratio, int32_t *d_idx,
                                        int32_t *d_dist, uint8_t *d_accept, void *stream) {
    if (!h || !d_query || !d_train || !d_q_offsets || !d_t_offsets || !d_idx || !d_dist || nseg < 1 || nseg > 65535 ||
        nq_total < 0 || max_q_per_seg < 1 || max_t_per_seg < 0 || k < 1 || k > 2)
        return ORBB_ERR_INVALID;
    if ((reinterpret_cast<uintptr_t>(d_query) | reinterpret_cast<uintptr_t>(d_train)) & 15) return ORBB_ERR_INVALID;
    if (nq_total == 0) return ORBB_OK;
    if (max_t_per_seg > 64 * (1 << 22)) return ORBB_ERR_CAPACITY;
    CK(h, cudaSetDevice(h->device));
    const int qblocks = (max_q_per_seg + 255) / 256 * nseg;
    int n_split = pick_split(qblocks, std::max(max_t_per_seg, 1));
    // the sizes are the caller's (no device read-back, no synchronisation): fewer splits if the scratch would not hold
    // n_split x nq_total rows, as long as every split still indexes its train rows with 22 bits
    const int min_split = std::max(1, (int)(((long long)max_t_per_seg + (1 << 22) - 1) >> 22));
    if ((size_t)n_split * nq_total > h->partial_cap) n_split = (int)(h->partial_cap / (size_t)nq_total);
    if (n_split < min_split) return ORBB_ERR_CAPACITY;
    CK(h, launch_match(d_query, d_train, d_q_offsets, d_t_offsets, nseg, nq_total, max_q_per_seg, 0, n_split,
                       h->d_partial, nq_total, k, ratio, d_idx, d_dist, d_accept, nullptr, static_cast<cudaStream_t>(stream)));
    h->n_launches += 2;
    return ORBB_OK;
}

extern "C" int orbb_match_windowed(orbb_handle *h, const uint8_t *d_query, const void *d_query_xy, int q_xy_stride, int nq,
                                   const uint8_t *d_train, const void *d_train_xy, int t_xy_stride, int nt, float max_px,
                                   int max_hamming, int32_t *d_idx, int32_t *d_dist, int32_t *d_nmatched, void *stream) {
    if (!h || !d_query || !d_query_xy || !d_train || !d_train_xy || !d_idx || !d_dist || nq < 0 || nt < 0 ||
        q_xy_stride < 8 || t_xy_stride < 8 || (q_xy_stride & 3) || (t_xy_stride & 3) || max_hamming < 0)
        return ORBB_ERR_INVALID;
    if ((reinterpret_cast<uintptr_t>(d_query) | reinterpret_cast<uintptr_t>(d_train)) & 15) return ORBB_ERR_INVALID;
    if ((reinterpret_cast<uintptr_t>(d_query_xy) | reinterpret_cast<uintptr_t>(d_train_xy)) & 3) return ORBB_ERR_INVALID;
    if (nq == 0) return ORBB_OK;
    CK(h, launch_match_windowed(d_query, d_query_xy, q_xy_stride, nq, d_train, d_train_xy, t_xy_stride, nt, max_px,
                                std::min(max_hamming, 257), d_idx, d_dist, d_nmatched, static_cast<cudaStream_t>(stream)));
    h->n_launches += 1;
    return ORBB_OK;
}

// ---------------------------------------------------------------- RGB-D association
static bool intrin_ok(const orbb_intrinsics *in) {
    if (!in || in->width < 1 || in->height < 1 || in->width > 16384 || in->height > 16384) return false;
    // FTHETA / KANNALA_BRANDT4 need atan/tan chains (or are ignored by the reference's copy of rsutil.h): refused
    return in->model == ORBB_DISTORTION_NONE || in->model == ORBB_DISTORTION_MODIFIED_BROWN_CONRADY ||
           in->model == ORBB_DISTORTION_INVERSE_BROWN_CONRADY || in->model == ORBB_DISTORTION_BROWN_CONRADY;
}

extern "C" int orbb_align_depth_to_other(orbb_handle *h, const uint16_t *d_depth, int n_frames, float depth_scale,
                                         const orbb_intrinsics *depth_intrin, const orbb_intrinsics *other_intrin,
                                         const orbb_extrinsics *depth_to_other, uint32_t *d_aligned_out, void *stream) {
    if (!h || !d_depth || !d_aligned_out || !depth_to_other || n_frames < 1 || !intrin_ok(depth_intrin) ||
        !intrin_ok(other_intrin))
        return ORBB_ERR_INVALID;
    // the reference asserts these out (cuda-align.cu:63-64): a forward-distorted image cannot be deprojected
    if (depth_intrin->model == ORBB_DISTORTION_MODIFIED_BROWN_CONRADY) return ORBB_ERR_INVALID;
    if (reinterpret_cast<uintptr_t>(d_aligned_out) & 15) return ORBB_ERR_INVALID;
    CK(h, cudaSetDevice(h->device));
    CK(h, launch_align(d_depth, n_frames, depth_scale, *depth_intrin, *other_intrin, *depth_to_other, d_aligned_out,
                       static_cast<cudaStream_t>(stream)));
    h->n_launches += 2;
    return ORBB_OK;
}

extern "C" int orbb_keypoint_pixel_to_point(orbb_handle *h, const uint32_t *d_aligned_depth, const orbb_intrinsics *other_intrin,
                                            int n_frames, const orbb_keypoint *d_kp_in, const uint8_t *d_desc_in,
                                            const int32_t *d_counts_in, int max_kp, orbb_keypoint *d_kp_out,
                                            uint8_t *d_desc_out, double *d_points, int32_t *d_valid_counts, void *stream) {
    if (!h || !d_aligned_depth || !d_kp_in || !d_desc_in || !d_counts_in || !d_kp_out || !d_desc_out || !d_points ||
        !d_valid_counts || n_frames < 1 || max_kp < 1 || !intrin_ok(other_intrin))
        return ORBB_ERR_INVALID;
    if (other_intrin->model == ORBB_DISTORTION_MODIFIED_BROWN_CONRADY) return ORBB_ERR_INVALID;  // cuda-align.cu:90-91
    if (d_kp_in == d_kp_out || d_desc_in == d_desc_out) return ORBB_ERR_INVALID;
    if ((reinterpret_cast<uintptr_t>(d_desc_in) | reinterpret_cast<uintptr_t>(d_desc_out)) & 15) return ORBB_ERR_INVALID;
    CK(h, cudaSetDevice(h->device));
    CK(h, launch_kp_to_point(d_aligned_depth, *other_intrin, n_frames, d_kp_in, d_desc_in, d_counts_in, max_kp, d_kp_out,
                             d_desc_out, d_points, d_valid_counts, static_cast<cudaStream_t>(stream)));
    h->n_launches += 1;
    return ORBB_OK;
}

extern "C" int orbb_reproject_points(orbb_handle *h, const double *d_points, const int32_t *d_counts, int n_frames,
                                     int max_kp, const double *d_T, const orbb_intrinsics *intrin, float *d_pos_out,
                                     void *stream) {
    if (!h || !d_points || !d_counts || !d_pos_out || n_frames < 1 || max_kp < 1 || !intrin_ok(intrin)) return ORBB_ERR_INVALID;
    if (reinterpret_cast<uintptr_t>(d_pos_out) & 7) return ORBB_ERR_INVALID;
    CK(h, cudaSetDevice(h->device));
    CK(h, launch_reproject(d_points, d_counts, n_frames, max_kp, d_T, *intrin, d_pos_out, static_cast<cudaStream_t>(stream)));
    h->n_launches += 1;
    return ORBB_OK;
}

extern "C" int orbb_match_windowed_batch(orbb_handle *h, const uint8_t *d_query, const float *d_query_xy,
                                         const int32_t *d_q_counts, const uint8_t *d_train, const void *d_train_xy,
                                         int t_xy_stride, const int32_t *d_t_counts, int n_frames, int max_kp, float max_px,
                                         int max_hamming, int32_t *d_idx, int32_t *d_dist, const double *d_query_points,
                                         const double *d_train_points, double *d_prev_matched, double *d_curr_matched,
                                         uint16_t *d_xy_u16, int32_t *d_nmatched, void *stream) {
    if (!h || !d_query || !d_query_xy || !d_q_counts || !d_train || !d_train_xy || !d_t_counts || !d_idx || !d_dist ||
        n_frames < 1 || max_kp < 1 || t_xy_stride < 8 || (t_xy_stride & 3) || max_hamming < 0)
        return ORBB_ERR_INVALID;
    if ((reinterpret_cast<uintptr_t>(d_query) | reinterpret_cast<uintptr_t>(d_train)) & 15) return ORBB_ERR_INVALID;
    if ((reinterpret_cast<uintptr_t>(d_query_xy) | reinterpret_cast<uintptr_t>(d_train_xy)) & 3) return ORBB_ERR_INVALID;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(h, cudaSetDevice(h->device));
    CK(h, launch_match_windowed_batch(d_query, d_query_xy, 8, d_q_counts, d_train, d_train_xy, t_xy_stride, d_t_counts,
                                      n_frames, max_kp, max_px, std::min(max_hamming, 257), d_idx, d_dist, st));
    h->n_launches += 1;
    if (d_nmatched || d_prev_matched || d_curr_matched || d_xy_u16) {
        CK(h, launch_compact_pairs(d_idx, d_q_counts, n_frames, max_kp, d_query_points, d_train_points, d_train_xy,
                                   t_xy_stride, d_prev_matched, d_curr_matched, d_xy_u16, d_nmatched, st));
        h->n_launches += 1;
    }
    return ORBB_OK;
}

extern "C" int orbb_match_projection_batch(orbb_handle *h, const uint8_t *d_query_desc, const float *d_query_uv,
                                           const orbb_keypoint *d_query_kp, const int32_t *d_q_counts,
                                           const uint8_t *d_train_desc, const orbb_keypoint *d_train_kp,
                                           const int32_t *d_t_counts, int n_frames, int max_kp, float th, int th_high,
                                           int check_orientation, int32_t *d_idx, int32_t *d_dist, int32_t *d_nmatched,
                                           void *stream) {
    if (!h || !d_query_desc || !d_query_uv || !d_query_kp || !d_q_counts || !d_train_desc || !d_train_kp || !d_t_counts ||
        !d_idx || !d_dist || n_frames < 1 || max_kp < 1 || !(th > 0.0f) || th_high < 0)
        return ORBB_ERR_INVALID;
    if ((reinterpret_cast<uintptr_t>(d_query_desc) | reinterpret_cast<uintptr_t>(d_train_desc)) & 15) return ORBB_ERR_INVALID;
    if (reinterpret_cast<uintptr_t>(d_query_uv) & 7) return ORBB_ERR_INVALID;
    CK(h, cudaSetDevice(h->device));
    CK(h, launch_match_projection(d_query_desc, d_query_uv, d_query_kp, d_q_counts, d_train_desc, d_train_kp, d_t_counts,
                                  n_frames, max_kp, th, std::min(th_high, 256), check_orientation, h->sf, h->nlevels, d_idx,
                                  d_dist, d_nmatched, static_cast<cudaStream_t>(stream)));
    h->n_launches += 2;
    return ORBB_OK;
}

extern "C" int orbb_compute_stereo_matches(orbb_handle *h, const orbb_keypoint *d_kp, const uint8_t *d_desc,
                                           const int32_t *d_counts, int max_kp, int n_pairs, float bf, float fx,
                                           float *d_uright, float *d_depth, int32_t *d_nstereo, void *stream) {
    if (!h || !d_kp || !d_desc || !d_counts || !d_uright || !d_depth || max_kp < 1 || n_pairs < 1 || !(bf > 0.0f) ||
        !(fx > 0.0f))
        return ORBB_ERR_INVALID;
    if (2 * n_pairs > h->n_frames_last) return ORBB_ERR_CAPACITY;  // the pairs' pyramids must be resident
    if (reinterpret_cast<uintptr_t>(d_desc) & 15) return ORBB_ERR_INVALID;
    CK(h, cudaSetDevice(h->device));
    if ((size_t)n_pairs * max_kp > h->stereo_cap) return ORBB_ERR_CAPACITY;  // scratch sized once in orbb_create
    CK(h, launch_stereo(h->d_levels, h->sf, h->inv_sf, h->nlevels, d_kp, d_desc, d_counts, max_kp, n_pairs, bf, fx, d_uright,
                        d_depth, h->d_stereo_sad, d_nstereo, static_cast<cudaStream_t>(stream)));
    h->n_launches += 2;
    return ORBB_OK;
}

extern "C" int orbb_rgb_to_grayscale(orbb_handle *h, const uint8_t *d_rgb, size_t rgb_pitch, size_t rgb_frame_stride, int width,
                                     int height, int n_frames, uint8_t *d_gray, size_t gray_pitch, size_t gray_frame_stride,
                                     void *stream) {
    if (!h || !d_rgb || !d_gray || width < 1 || height < 1 || n_frames < 1 || rgb_pitch < 3 * (size_t)width ||
        gray_pitch < (size_t)width)
        return ORBB_ERR_INVALID;
    CK(h, cudaSetDevice(h->device));
    CK(h, launch_rgb_to_gray(d_rgb, rgb_pitch, rgb_frame_stride, width, height, n_frames, d_gray, gray_pitch,
                             gray_frame_stride, static_cast<cudaStream_t>(stream)));
    h->n_launches += 1;
    return ORBB_OK;
}

// ---------------------------------------------------------------- debug / parity access
// Overwrites every scratch buffer of the handle that carries no state between calls (pyramid levels incl. their
// unused pad bytes, blurred levels, candidate lists, quadtree sort / selection scratch, host-path staging, matcher and
// stereo scratch) with `value`.  A following extraction must give the same bytes as before: the stand-in for
// compute-sanitizer's initcheck on pools where the sanitizer is not available (tests/test_gpu_contract.py).  The
// quadtree cell tables, the candidate / selection counters and the chain counters are state (zero / monotonic by
// contract) and are left alone.  Synchronises.
extern "C" int orbb_debug_poison(orbb_handle *h, int value) {
    if (!h) return ORBB_ERR_INVALID;
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaDeviceSynchronize());
    const size_t B = (size_t)h->max_batch;
    for (int l = 0; l < h->nlevels; ++l) {
        LevelDev &L = h->lv[l];
        CK(h, cudaMemset(L.img, value, (size_t)L.frame_stride * B));
        CK(h, cudaMemset(L.blur, value, (size_t)L.blur_stride * B));
        CK(h, cudaMemset(L.cand, value, sizeof(uint32_t) * (size_t)L.cand_cap * B));
        CK(h, cudaMemset(L.kv_a, value, sizeof(uint2) * (size_t)L.cand_cap * B));
        CK(h, cudaMemset(L.kv_b, value, sizeof(uint2) * (size_t)L.cand_cap * B));
        CK(h, cudaMemset(L.sd, value, (size_t)L.cand_cap * B));
        CK(h, cudaMemset(L.sel, value, sizeof(uint32_t) * (size_t)L.sel_cap * B));
    }
    for (int i = 0; i < 2; ++i) {
        CK(h, cudaMemset(h->d_in2[i], value, (size_t)h->w * h->h * B));
        CK(h, cudaMemset(h->d_kp2[i], value, sizeof(orbb_keypoint) * (size_t)h->max_kp * B));
        CK(h, cudaMemset(h->d_desc2[i], value, (size_t)h->max_kp * B * 32));
        CK(h, cudaMemset(h->d_counts2[i], value, sizeof(int) * B));
    }
    CK(h, cudaMemset(h->d_partial, value, sizeof(int4) * h->partial_cap));
    CK(h, cudaMemset(h->d_stereo_sad, value, sizeof(int) * h->stereo_cap));
    CK(h, cudaMemset(h->d_dump, value, (size_t)h->dump_off[h->nlevels]));
    CK(h, cudaDeviceSynchronize());
    return ORBB_OK;
}

// POPC lanes per clock per SM, measured: one 1024-thread CTA per SM running register-only POPC chains (k_popc_rate);
// the median over the CTAs' own cycle counts is clock independent.  Synchronises.
extern "C" int orbb_debug_popc_rate(orbb_handle *h, double *popc_per_clk_per_sm) {
    if (!h || !popc_per_clk_per_sm) return ORBB_ERR_INVALID;
    CK(h, cudaSetDevice(h->device));
    cudaDeviceProp prop;
    CK(h, cudaGetDeviceProperties(&prop, h->device));
    const int n = prop.multiProcessorCount, iters = 20000;
    long long *d_cyc = reinterpret_cast<long long *>(h->d_partial);  // scratch: n x 8 bytes + 4
    unsigned *d_sink = reinterpret_cast<unsigned *>(d_cyc + n);
    for (int rep = 0; rep < 2; ++rep) CK(h, launch_popc_rate(n, iters, d_cyc, d_sink, 0));  // first run warms the clocks
    CK(h, cudaDeviceSynchronize());
    std::vector<long long> cyc((size_t)n);
    CK(h, cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * n, cudaMemcpyDeviceToHost));
    std::sort(cyc.begin(), cyc.end());
    *popc_per_clk_per_sm = 1024.0 * 8.0 * iters / (double)std::max<long long>(cyc[n / 2], 1);
    h->n_launches += 2;
    return ORBB_OK;
}

extern "C" int orbb_debug_matcher_kind(void) { return matcher_kind(); }

// int8 tensor-core MMAs (mma.sync m16n8k32 = IMMA.16832.S8.S8) per clock per SM, measured the same way (k_imma_rate):
// the roof of the tensor-core matcher, 16 descriptor pairs per MMA.  Synchronises.
extern "C" int orbb_debug_imma_rate(orbb_handle *h, double *imma_per_clk_per_sm) {
    if (!h || !imma_per_clk_per_sm) return ORBB_ERR_INVALID;
    CK(h, cudaSetDevice(h->device));
    cudaDeviceProp prop;
    CK(h, cudaGetDeviceProperties(&prop, h->device));
    const int n = prop.multiProcessorCount, iters = 4000;
    long long *d_cyc = reinterpret_cast<long long *>(h->d_partial);  // scratch: n x 8 bytes + 4
    unsigned *d_sink = reinterpret_cast<unsigned *>(d_cyc + n);
    for (int rep = 0; rep < 2; ++rep) CK(h, launch_imma_rate(n, iters, d_cyc, d_sink, 0));  // first run warms the clocks
    CK(h, cudaDeviceSynchronize());
    std::vector<long long> cyc((size_t)n);
    CK(h, cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * n, cudaMemcpyDeviceToHost));
    std::sort(cyc.begin(), cyc.end());
    *imma_per_clk_per_sm = 32.0 * 8.0 * iters / (double)std::max<long long>(cyc[n / 2], 1);
    h->n_launches += 2;
    return ORBB_OK;
}

extern "C" int orbb_debug_get_padded(orbb_handle *h, int frame, int level, uint8_t *host_out) {
    if (!h || !host_out || level < 0 || level >= h->nlevels || frame < 0 || frame >= h->max_batch) return ORBB_ERR_INVALID;
    const LevelDev &L = h->lv[level];
    CK(h, cudaDeviceSynchronize());
    CK(h, cudaMemcpy2D(host_out, L.w + 38, L.img + (size_t)frame * L.frame_stride + ORBB_PAD_X0, L.pitch, L.w + 38,
                       L.h + 38, cudaMemcpyDeviceToHost));
    return ORBB_OK;
}

extern "C" int orbb_debug_get_blurred(orbb_handle *h, int frame, int level, uint8_t *host_out) {
    if (!h || !host_out || level < 0 || level >= h->nlevels || frame < 0 || frame >= h->max_batch) return ORBB_ERR_INVALID;
    const LevelDev &L = h->lv[level];
    CK(h, cudaDeviceSynchronize());
    CK(h, cudaMemcpy2D(host_out, L.w, L.blur + (size_t)frame * L.blur_stride, L.pitch, L.w, L.h, cudaMemcpyDeviceToHost));
    return ORBB_OK;
}

extern "C" int orbb_debug_get_scores(orbb_handle *h, int frame, int level, uint8_t *host_out) {
    if (!h || !host_out || level < 0 || level >= h->nlevels || frame < 0 || frame >= h->max_batch) return ORBB_ERR_INVALID;
    const LevelDev &L = h->lv[level];
    CK(h, cudaMemset(h->d_dump, 0, (size_t)h->dump_off[h->nlevels]));
    CK(h, launch_fast_dump(h->tma_maps, h->d_levels, h->d_cells, h->n_cells, h->nlevels, h->t_lo, h->t_hi, h->fcfg, frame, h->d_dump,
                           h->d_dump_off, 0));
    CK(h, cudaDeviceSynchronize());
    CK(h, cudaMemcpy(host_out, h->d_dump + h->dump_off[level], (size_t)L.w * L.h, cudaMemcpyDeviceToHost));
    return ORBB_OK;
}

static int download_packed(orbb_handle *h, const uint32_t *d_src, int n, int32_t *host_xyr, int max_n) {
    std::vector<uint32_t> tmp((size_t)std::max(n, 1));
    CK(h, cudaMemcpy(tmp.data(), d_src, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n && i < max_n; ++i) {
        host_xyr[3 * i] = (int)(tmp[i] & 0xfffu);
        host_xyr[3 * i + 1] = (int)((tmp[i] >> 12) & 0xfffu);
        host_xyr[3 * i + 2] = (int)(tmp[i] >> 24) - 1;  // response = arc score - 1
    }
    return n;
}

extern "C" int orbb_debug_get_candidates(orbb_handle *h, int frame, int level, int32_t *host_xyr, int max_n) {
    if (!h || !host_xyr || level < 0 || level >= h->nlevels || frame < 0 || frame >= h->max_batch) return ORBB_ERR_INVALID;
    const LevelDev &L = h->lv[level];
    int n = 0;
    CK(h, cudaDeviceSynchronize());
    CK(h, cudaMemcpy(&n, h->d_cand_count + frame * h->nlevels + level, sizeof(int), cudaMemcpyDeviceToHost));
    n = std::min(n, L.cand_cap);
    return download_packed(h, L.cand + (size_t)frame * L.cand_cap, n, host_xyr, max_n);
}

extern "C" int orbb_debug_get_selected(orbb_handle *h, int frame, int level, int32_t *host_xyr, int max_n) {
    if (!h || !host_xyr || level < 0 || level >= h->nlevels || frame < 0 || frame >= h->max_batch) return ORBB_ERR_INVALID;
    const LevelDev &L = h->lv[level];
    int n = 0;
    CK(h, cudaDeviceSynchronize());
    CK(h, cudaMemcpy(&n, h->d_sel_count + frame * h->nlevels + level, sizeof(int), cudaMemcpyDeviceToHost));
    n = std::min(n, L.sel_cap);
    return download_packed(h, L.sel + (size_t)frame * L.sel_cap, n, host_xyr, max_n);
}

extern "C" int orbb_debug_distribute(orbb_handle *h, int level, const int32_t *host_xyr, int n, int quota,
                                     int32_t *host_out_xyr, int max_out) {
    if (!h || !host_xyr || !host_out_xyr || level < 0 || level >= h->nlevels || n < 0 || quota < 0) return ORBB_ERR_INVALID;
    const LevelDev &L = h->lv[level];
    if (n > L.cand_cap || std::max(quota, 4 * L.n_ini) + 8 > L.sel_cap) return ORBB_ERR_CAPACITY;
    std::vector<uint32_t> packed((size_t)std::max(n, 1));
    for (int i = 0; i < n; ++i) {
        const int x = host_xyr[3 * i], y = host_xyr[3 * i + 1], r = host_xyr[3 * i + 2];
        if (x < 0 || x >= L.w - 32 || y < 0 || y >= L.h - 32 || r < 0 || r > 254) return ORBB_ERR_INVALID;
        packed[i] = (uint32_t)x | ((uint32_t)y << 12) | ((uint32_t)(r + 1) << 24);
    }
    CK(h, cudaDeviceSynchronize());
    CK(h, cudaMemcpy(L.cand, packed.data(), sizeof(uint32_t) * (size_t)n, cudaMemcpyHostToDevice));
    CK(h, cudaMemcpy(h->d_cand_count + level, &n, sizeof(int), cudaMemcpyHostToDevice));
    CK(h, launch_octree_bin(h->d_levels, h->nlevels, h->d_cand_count, level, 0, 1, 0));  // what the FAST kernel does while emitting
    CK(h, launch_octree(h->d_levels, h->nlevels, h->d_cand_count, h->d_sel_count, level, 1, 0, 1, quota, h->sel_cap_max,
                        h->pcap, h->pcap2, 0));
    return orbb_debug_get_selected(h, 0, level, host_out_xyr, max_out);
}
