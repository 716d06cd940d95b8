// orbb_stage.cu -- the SlamGpuPipeline slot body rewritten around the handle: host-side mirror of
// reference src/SlamGpuPipeline/buildStream.cpp:345-660 for batches of consecutive RGB-D frames.
//
// Streams (the reference uses stream / align_stream / nvjpeg_stream with four cudaStreamSynchronize per frame):
//   s_in     H2D of gray, depth and the T matrices of batch k (double-buffered device staging)
//   s_align  align_depth_to_other of batch k                      -- runs next to the extraction, as :376-394
//   s_main   extraction -> depth gate / 3-D lift -> reprojection -> windowed match + compaction -> carry row
//   s_out    D2H of the batch's results into pinned host memory (double-buffered)
// linked by events only; the host blocks in orbb_rgbd_stage_wait and, with two batches already in flight, in
// submit.  Device rows: every per-keypoint array has max_batch + 1 frame rows; row 0 carries the last frame of the
// previous batch, so "previous frame of frame f" is simply row f and "current" is row f + 1.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "orbb_internal.cuh"

struct orbb_rgbd_stage {
    orbb_rgbd_config cfg{};
    orbb_handle *h = nullptr;
    int device = 0, B = 0, max_kp = 0;
    size_t gray_bytes = 0, depth_px = 0, img_px = 0;
    cudaStream_t s_in = nullptr, s_align = nullptr, s_main = nullptr, s_out = nullptr;
    cudaEvent_t ev_gfork = nullptr, ev_gjoin = nullptr, ev_in[2] = {}, ev_gray[2] = {}, ev_out[2] = {}, ev_main[2] = {}, ev_align = nullptr, ev_gate = nullptr;
    // device
    uint8_t *d_gray[2] = {};
    uint16_t *d_depth[2] = {};
    double *d_T[2] = {};
    uint32_t *d_aligned = nullptr;
    orbb_keypoint *d_kp_raw = nullptr, *d_kp = nullptr;
    uint8_t *d_desc_raw = nullptr, *d_desc = nullptr;
    int *d_counts_raw = nullptr, *d_counts_blk = nullptr, *d_valid = nullptr, *d_idx = nullptr, *d_dist = nullptr, *d_nm = nullptr;
    double *d_pts = nullptr, *d_prev_m = nullptr, *d_curr_m = nullptr;
    float *d_pos = nullptr;
    uint16_t *d_xy = nullptr;
    // every result array is a slice of ONE device block, mirrored by one pinned host block per parity, so that a
    // full batch comes back with a single D2H copy (nine separate copies cost ~8 us of latency each)
    uint8_t *d_block = nullptr, *h_block[2] = {nullptr, nullptr};
    size_t block_bytes = 0;
    // pinned host results, by ticket parity
    struct Host {
        int *counts, *valid, *matched;
        orbb_keypoint *kp;
        uint8_t *desc;
        double *pts, *prev_m, *curr_m, *T;
        uint16_t *xy;
        int n_frames;
    } host[2] = {};
    std::vector<void *> dev_allocs, host_allocs;
    long long n_submitted = 0;
    bool gate_recorded = false;
    int carry_from = 0;  // > 0: row `carry_from` of the result block (the previous batch's last frame) still has to become row 0
    // Small batches (a lone frame per wake-up is the reference's operating mode) replay a captured CUDA graph of
    // everything after the H2D copies; one graph per (parity, frame count, pose given)
    struct FrameGraph { int p, n, has_T; cudaGraphExec_t exec; long long launches; };
    std::vector<FrameGraph> graphs;
    int use_graph = 1, graph_max_frames = 4;
    int lone_fused = 1;  // max_batch-1 graph path: gate + lift + match in one launch, reprojection under the extraction
    // diagnostics (ORBB_STAGE_PROF=1): timing events at the phase boundaries of the last submit, printed by wait()
    bool prof = false;
    cudaEvent_t pe[8] = {};
};
#define SPROF(s, k, stream) do { if ((s)->prof) cudaEventRecord((s)->pe[k], stream); } while (0)

#define SCK(s, call)                                                                                       \
    do {                                                                                                   \
        cudaError_t e__ = (call);                                                                          \
        if (e__ != cudaSuccess) {                                                                          \
            fprintf(stderr, "orbb_rgbd_stage: %s:%d %s: %s\n", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return ORBB_ERR_CUDA;                                                                          \
        }                                                                                                  \
    } while (0)
#define SRC(call)                 \
    do {                          \
        const int rc__ = (call);  \
        if (rc__ < 0) return rc__; \
    } while (0)

template <typename T>
static cudaError_t sdev(orbb_rgbd_stage *s, T **out, size_t count) {
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T));
    if (e == cudaSuccess) { s->dev_allocs.push_back(p); *out = static_cast<T *>(p); }
    return e;
}
template <typename T>
static cudaError_t shost(orbb_rgbd_stage *s, T **out, size_t count) {
    void *p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, std::max<size_t>(count, 1) * sizeof(T), cudaHostAllocDefault);
    if (e == cudaSuccess) { s->host_allocs.push_back(p); *out = static_cast<T *>(p); }
    return e;
}

extern "C" int orbb_rgbd_stage_destroy(orbb_rgbd_stage *s) {
    if (!s) return ORBB_OK;
    cudaSetDevice(s->device);
    cudaDeviceSynchronize();
    for (void *p : s->dev_allocs) cudaFree(p);
    for (void *p : s->host_allocs) cudaFreeHost(p);
    for (int i = 0; i < 2; ++i) {
        if (s->ev_in[i]) cudaEventDestroy(s->ev_in[i]);
        if (s->ev_gray[i]) cudaEventDestroy(s->ev_gray[i]);
        if (s->ev_out[i]) cudaEventDestroy(s->ev_out[i]);
        if (s->ev_main[i]) cudaEventDestroy(s->ev_main[i]);
    }
    if (s->ev_align) cudaEventDestroy(s->ev_align);
    if (s->ev_gate) cudaEventDestroy(s->ev_gate);
    if (s->ev_gfork) cudaEventDestroy(s->ev_gfork);
    if (s->ev_gjoin) cudaEventDestroy(s->ev_gjoin);
    for (auto &g : s->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    for (cudaStream_t st : {s->s_in, s->s_align, s->s_main, s->s_out})
        if (st) cudaStreamDestroy(st);
    orbb_destroy(s->h);
    delete s;
    return ORBB_OK;
}

extern "C" int orbb_rgbd_stage_create(orbb_rgbd_stage **out, const orbb_rgbd_config *cfg, int device) {
    if (!out || !cfg || cfg->max_batch < 1 || cfg->max_hamming_distance < 0) return ORBB_ERR_INVALID;
    *out = nullptr;
    orbb_rgbd_stage *s = new (std::nothrow) orbb_rgbd_stage();
    if (!s) return ORBB_ERR_INVALID;
    s->cfg = *cfg;
    s->B = cfg->max_batch;
    int rc = orbb_create(&s->h, &cfg->orb, cfg->image_intrin.width, cfg->image_intrin.height, cfg->max_batch, device);
    if (rc) { delete s; return rc; }
    cudaGetDevice(&s->device);
    s->max_kp = orbb_max_keypoints_per_frame(s->h);
    s->img_px = (size_t)cfg->image_intrin.width * cfg->image_intrin.height;
    s->gray_bytes = s->img_px;
    s->depth_px = (size_t)cfg->depth_intrin.width * cfg->depth_intrin.height;
    const size_t B = s->B, mk = s->max_kp, R = B + 1;
#define SCKC(call)                                                        \
    do {                                                                  \
        if ((call) != cudaSuccess) {                                      \
            fprintf(stderr, "orbb_rgbd_stage_create: %s failed: %s\n", #call, cudaGetErrorString(cudaGetLastError())); \
            orbb_rgbd_stage_destroy(s);                                   \
            return ORBB_ERR_CUDA;                                         \
        }                                                                 \
    } while (0)
    if (cfg->depth_intrin.width < 1 || cfg->depth_intrin.height < 1) { orbb_rgbd_stage_destroy(s); return ORBB_ERR_INVALID; }
    for (int i = 0; i < 2; ++i) {
        SCKC(sdev(s, &s->d_gray[i], s->gray_bytes * B));
        SCKC(sdev(s, &s->d_depth[i], s->depth_px * B));
        SCKC(sdev(s, &s->d_T[i], 16 * B));
        SCKC(cudaEventCreateWithFlags(&s->ev_in[i], cudaEventDisableTiming));
        SCKC(cudaEventCreateWithFlags(&s->ev_gray[i], cudaEventDisableTiming));
        SCKC(cudaEventCreateWithFlags(&s->ev_out[i], cudaEventDisableTiming));
        SCKC(cudaEventCreateWithFlags(&s->ev_main[i], cudaEventDisableTiming));
        orbb_rgbd_stage::Host &H = s->host[i];
        SCKC(shost(s, &H.T, B * 16));
    }
    SCKC(cudaEventCreateWithFlags(&s->ev_align, cudaEventDisableTiming));
    SCKC(cudaEventCreateWithFlags(&s->ev_gate, cudaEventDisableTiming));
    SCKC(cudaEventCreateWithFlags(&s->ev_gfork, cudaEventDisableTiming));
    SCKC(cudaEventCreateWithFlags(&s->ev_gjoin, cudaEventDisableTiming));
    if (const char *e = getenv("ORBB_STAGE_GRAPH")) s->use_graph = atoi(e);
    if (const char *e = getenv("ORBB_STAGE_FUSED")) s->lone_fused = atoi(e);
    SCKC(sdev(s, &s->d_aligned, s->img_px * B));
    SCKC(sdev(s, &s->d_kp_raw, B * mk)); SCKC(sdev(s, &s->d_desc_raw, B * mk * 32));
    SCKC(sdev(s, &s->d_pos, B * mk * 2)); SCKC(sdev(s, &s->d_idx, B * mk)); SCKC(sdev(s, &s->d_dist, B * mk));
    // the extraction's per-frame counts land OUTSIDE the result block: batch k+1's extraction is enqueued before the
    // wait for batch k's D2H of the block, so it must not write into it
    SCKC(sdev(s, &s->d_counts_raw, B));
    {
        // result block layout (256-byte aligned slices); rows 0 of kp/desc/pts/valid carry the previous batch's last frame
        size_t off = 0;
        auto slice = [&](size_t bytes) { const size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; };
        const size_t o_kp = slice(sizeof(orbb_keypoint) * R * mk), o_desc = slice(32 * R * mk), o_pts = slice(sizeof(double) * 3 * R * mk);
        const size_t o_prev = slice(sizeof(double) * 3 * B * mk), o_curr = slice(sizeof(double) * 3 * B * mk);
        const size_t o_xy = slice(sizeof(uint16_t) * 2 * B * mk), o_cnt = slice(sizeof(int) * B), o_valid = slice(sizeof(int) * R);
        const size_t o_nm = slice(sizeof(int) * B);
        s->block_bytes = off;
        SCKC(sdev(s, &s->d_block, off));
        uint8_t *d = s->d_block;
        s->d_kp = reinterpret_cast<orbb_keypoint *>(d + o_kp); s->d_desc = d + o_desc; s->d_pts = reinterpret_cast<double *>(d + o_pts);
        s->d_prev_m = reinterpret_cast<double *>(d + o_prev); s->d_curr_m = reinterpret_cast<double *>(d + o_curr);
        s->d_xy = reinterpret_cast<uint16_t *>(d + o_xy); s->d_counts_blk = reinterpret_cast<int *>(d + o_cnt);
        s->d_valid = reinterpret_cast<int *>(d + o_valid); s->d_nm = reinterpret_cast<int *>(d + o_nm);
        for (int i = 0; i < 2; ++i) {
            SCKC(shost(s, &s->h_block[i], off));
            uint8_t *hb = s->h_block[i];
            orbb_rgbd_stage::Host &H = s->host[i];
            H.kp = reinterpret_cast<orbb_keypoint *>(hb + o_kp) + mk; H.desc = hb + o_desc + 32 * mk;
            H.pts = reinterpret_cast<double *>(hb + o_pts) + 3 * mk;
            H.prev_m = reinterpret_cast<double *>(hb + o_prev); H.curr_m = reinterpret_cast<double *>(hb + o_curr);
            H.xy = reinterpret_cast<uint16_t *>(hb + o_xy); H.counts = reinterpret_cast<int *>(hb + o_cnt);
            H.valid = reinterpret_cast<int *>(hb + o_valid) + 1; H.matched = reinterpret_cast<int *>(hb + o_nm);
        }
    }
    SCKC(cudaMemset(s->d_valid, 0, sizeof(int) * R));
    for (cudaStream_t *st : {&s->s_in, &s->s_align, &s->s_main, &s->s_out})
        SCKC(cudaStreamCreateWithFlags(st, cudaStreamNonBlocking));
    SCKC(cudaDeviceSynchronize());
#undef SCKC
    *out = s;
    return ORBB_OK;
}

extern "C" orbb_handle *orbb_rgbd_stage_handle(orbb_rgbd_stage *s) { return s ? s->h : nullptr; }

extern "C" int orbb_rgbd_stage_reset(orbb_rgbd_stage *s) {
    if (!s) return ORBB_ERR_INVALID;
    SCK(s, cudaSetDevice(s->device));
    s->carry_from = 0;  // a pending carry would overwrite the reset
    SCK(s, cudaMemsetAsync(s->d_valid, 0, sizeof(int), s->s_main));  // row 0 = "no previous frame"
    return ORBB_OK;
}

// results of a small batch to the host on stream m: the whole block for a full batch, else only the rows in use
static int enqueue_results_d2h(orbb_rgbd_stage *s, int p, int n_frames, cudaStream_t m) {
    const size_t n = n_frames, mk = s->max_kp;
    orbb_rgbd_stage::Host &H = s->host[p];
    if (n_frames == s->B) {
        SCK(s, cudaMemcpyAsync(s->h_block[p], s->d_block, s->block_bytes, cudaMemcpyDeviceToHost, m));
    } else {
        SCK(s, cudaMemcpyAsync(H.counts, s->d_counts_blk, sizeof(int) * n, cudaMemcpyDeviceToHost, m));
        SCK(s, cudaMemcpyAsync(H.valid, s->d_valid + 1, sizeof(int) * n, cudaMemcpyDeviceToHost, m));
        SCK(s, cudaMemcpyAsync(H.matched, s->d_nm, sizeof(int) * n, cudaMemcpyDeviceToHost, m));
        SCK(s, cudaMemcpyAsync(H.kp, s->d_kp + mk, sizeof(orbb_keypoint) * n * mk, cudaMemcpyDeviceToHost, m));
        SCK(s, cudaMemcpyAsync(H.desc, s->d_desc + 32 * mk, 32 * n * mk, cudaMemcpyDeviceToHost, m));
        SCK(s, cudaMemcpyAsync(H.pts, s->d_pts + 3 * mk, sizeof(double) * 3 * n * mk, cudaMemcpyDeviceToHost, m));
        SCK(s, cudaMemcpyAsync(H.prev_m, s->d_prev_m, sizeof(double) * 3 * n * mk, cudaMemcpyDeviceToHost, m));
        SCK(s, cudaMemcpyAsync(H.curr_m, s->d_curr_m, sizeof(double) * 3 * n * mk, cudaMemcpyDeviceToHost, m));
        SCK(s, cudaMemcpyAsync(H.xy, s->d_xy, sizeof(uint16_t) * 2 * n * mk, cudaMemcpyDeviceToHost, m));
    }
    return ORBB_OK;
}

// The second half of one small batch -- depth gate / 3-D lift, reprojection, counts, windowed match, pair
// compaction and the D2H of the results -- enqueued on s_main: the body of the graph.  (The first half needs no graph of
// the stage's own: the alignment is one call on s_align and the extraction replays the extractor's graph.)  Same calls,
// same arguments, same results as the streamed path below.
static int enqueue_small_batch_tail(orbb_rgbd_stage *s, int p, int n_frames, bool has_T) {
    const size_t n = n_frames, mk = s->max_kp;
    cudaStream_t m = s->s_main;
    // A lone frame's "previous" points are row 0 only, carried before the graph starts (see submit): their reprojection and
    // the counts for the result block run on s_align next to the depth gate / 3-D lift of the new frame.  With more frames
    // the previous rows 1..n-1 are this batch's own, so the reprojection follows the lift.
    if (n_frames == 1 && s->lone_fused && s->max_kp <= 2048) {
        // A lone frame: the previous points were reprojected on s_align under the extraction (see submit), so what is left is
        // ONE launch for depth gate + 3-D lift + windowed match (k_gate_match_lone), the pair compaction and the D2H.
        SCK(s, orbb::launch_gate_match_lone(s->d_aligned, s->cfg.image_intrin, s->d_kp_raw, s->d_desc_raw, s->d_counts_raw, s->max_kp,
                                            s->d_kp + mk, s->d_desc + 32 * mk, s->d_pts + 3 * mk, s->d_valid + 1, s->d_counts_blk,
                                            s->d_desc, s->d_pos, s->d_valid, s->cfg.max_pixel_distance,
                                            std::min(s->cfg.max_hamming_distance, 257), s->d_idx, s->d_dist, m));
        SCK(s, orbb::launch_compact_pairs(s->d_idx, s->d_valid, 1, s->max_kp, s->d_pts, s->d_pts + 3 * mk, s->d_kp + mk,
                                          (int)sizeof(orbb_keypoint), s->d_prev_m, s->d_curr_m, s->d_xy, s->d_nm, m));
        orbb::note_replay(s->h, 1, 2);
        return enqueue_results_d2h(s, p, n_frames, m);
    }
    const bool fork = n_frames == 1;
    cudaStream_t side = fork ? s->s_align : m;
    if (fork) {
        SCK(s, cudaEventRecord(s->ev_gfork, m));
        SCK(s, cudaStreamWaitEvent(side, s->ev_gfork, 0));
    }
    SRC(orbb_keypoint_pixel_to_point(s->h, s->d_aligned, &s->cfg.image_intrin, n_frames, s->d_kp_raw, s->d_desc_raw,
                                     s->d_counts_raw, s->max_kp, s->d_kp + mk, s->d_desc + 32 * mk, s->d_pts + 3 * mk,
                                     s->d_valid + 1, m));
    SCK(s, cudaMemcpyAsync(s->d_counts_blk, s->d_counts_raw, sizeof(int) * n, cudaMemcpyDeviceToDevice, side));
    SRC(orbb_reproject_points(s->h, s->d_pts, s->d_valid, n_frames, s->max_kp, has_T ? s->d_T[p] : nullptr,
                              &s->cfg.image_intrin, s->d_pos, side));
    if (fork) {
        SCK(s, cudaEventRecord(s->ev_gjoin, side));
        SCK(s, cudaStreamWaitEvent(m, s->ev_gjoin, 0));
    }
    SRC(orbb_match_windowed_batch(s->h, s->d_desc, s->d_pos, s->d_valid, s->d_desc + 32 * mk, s->d_kp + mk,
                                  (int)sizeof(orbb_keypoint), s->d_valid + 1, n_frames, s->max_kp, s->cfg.max_pixel_distance,
                                  s->cfg.max_hamming_distance, s->d_idx, s->d_dist, s->d_pts, s->d_pts + 3 * mk,
                                  s->d_prev_m, s->d_curr_m, s->d_xy, s->d_nm, m));
    return enqueue_results_d2h(s, p, n_frames, m);
}

// The graph for this (parity, frame count, pose given): captured on first use.  nullptr = use the streamed path.
static orbb_rgbd_stage::FrameGraph *small_batch_graph(orbb_rgbd_stage *s, int p, int n_frames, bool has_T) {
    for (auto &g : s->graphs)
        if (g.p == p && g.n == n_frames && g.has_T == (int)has_T) return &g;
    if (s->graphs.size() >= 32) return nullptr;  // a caller that varies the batch size a lot: not worth more graphs
    const long long l0 = orbb_get_launch_count(s->h);
    if (cudaStreamBeginCapture(s->s_main, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    const int rc = enqueue_small_batch_tail(s, p, n_frames, has_T);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(s->s_main, &graph);
    const long long launches = orbb_get_launch_count(s->h) - l0;
    orbb::note_replay(s->h, n_frames, -launches);  // nothing ran yet
    cudaGraphExec_t exec = nullptr;
    if (rc == ORBB_OK && ce == cudaSuccess && cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) exec = nullptr;
    if (graph) cudaGraphDestroy(graph);
    if (!exec) { cudaGetLastError(); s->use_graph = 0; return nullptr; }  // fall back to the streamed path for good
    s->graphs.push_back({p, n_frames, (int)has_T, exec, launches});
    return &s->graphs.back();
}

extern "C" int orbb_rgbd_stage_submit(orbb_rgbd_stage *s, const uint8_t *h_gray, const uint16_t *h_depth, int n_frames,
                                      const double *h_T) {
    if (!s || !h_gray || !h_depth) return ORBB_ERR_INVALID;
    if (n_frames < 1 || n_frames > s->B) return ORBB_ERR_CAPACITY;
    SCK(s, cudaSetDevice(s->device));
    const int ticket = (int)s->n_submitted, p = ticket & 1;
    const size_t n = n_frames, mk = s->max_kp;
    orbb_rgbd_stage::Host &H = s->host[p];
    if (ticket >= 2) SCK(s, cudaEventSynchronize(s->ev_out[p]));  // batch ticket-2 owned this parity's buffers
    // ---- inputs
    if (getenv("ORBB_STAGE_PROF") && !s->prof) {
        s->prof = true;
        for (auto &e : s->pe) cudaEventCreate(&e);
    }
    SPROF(s, 0, s->s_in);
    SCK(s, cudaMemcpyAsync(s->d_gray[p], h_gray, s->gray_bytes * n, cudaMemcpyHostToDevice, s->s_in));
    SCK(s, cudaEventRecord(s->ev_gray[p], s->s_in));  // the extraction needs nothing else
    // small batch: the extraction (a replay of the extractor's graph, the start of the critical path) is enqueued before
    // the host spends time on the depth copy and the pose upload
    orbb_rgbd_stage::FrameGraph *g_small =
        (s->use_graph && n_frames <= s->graph_max_frames) ? small_batch_graph(s, p, n_frames, h_T != nullptr) : nullptr;
    if (g_small) {
        SCK(s, cudaStreamWaitEvent(s->s_main, s->ev_gray[p], 0));
        SRC(orbb_extract_batch_device(s->h, s->d_gray[p], (size_t)s->cfg.image_intrin.width, s->gray_bytes, n_frames,
                                      s->d_kp_raw, s->d_desc_raw, s->d_counts_raw, s->max_kp, s->s_main));
        SPROF(s, 3, s->s_main);
    }
    SCK(s, cudaMemcpyAsync(s->d_depth[p], h_depth, s->depth_px * n * sizeof(uint16_t), cudaMemcpyHostToDevice, s->s_in));
    if (h_T) {
        std::memcpy(H.T, h_T, sizeof(double) * 16 * n);
        SCK(s, cudaMemcpyAsync(s->d_T[p], H.T, sizeof(double) * 16 * n, cudaMemcpyHostToDevice, s->s_in));
    }
    SCK(s, cudaEventRecord(s->ev_in[p], s->s_in));
    SPROF(s, 1, s->s_in);
    // ---- small batch (a lone frame per wake-up is the reference's operating mode): the extraction starts as soon as the
    // gray frame is on the device and replays the extractor's graph, the alignment runs next to it once the depth frame
    // has arrived, and everything after the two -- carry, depth gate, reprojection, match, compaction, D2H -- is ONE graph
    // launch of the stage's own.  The host issues ~15 calls instead of ~40 and the GPU starts ~25 us earlier.
    {
        if (orbb_rgbd_stage::FrameGraph *g = g_small) {
            // s_align: once the previous batch has let go of the result block and the aligned-depth buffer, its last frame
            // becomes row 0 (under this batch's extraction), then the alignment as soon as the depth frames are there
            if (s->gate_recorded) SCK(s, cudaStreamWaitEvent(s->s_align, s->ev_gate, 0));
            if (ticket >= 1) SCK(s, cudaStreamWaitEvent(s->s_align, s->ev_out[p ^ 1], 0));  // a streamed batch's D2H (s_out)
            if (s->carry_from > 0)
                SCK(s, orbb::launch_stage_carry(s->d_kp, s->d_desc, s->d_pts, s->d_valid, s->carry_from, s->max_kp, nullptr, nullptr,
                                                0, s->s_align));
            SCK(s, cudaStreamWaitEvent(s->s_align, s->ev_in[p], 0));
            if (n_frames == 1 && s->lone_fused && s->max_kp <= 2048)  // previous points (row 0) -> current image, under the extraction
                SRC(orbb_reproject_points(s->h, s->d_pts, s->d_valid, 1, s->max_kp, h_T ? s->d_T[p] : nullptr, &s->cfg.image_intrin,
                                          s->d_pos, s->s_align));
            SRC(orbb_align_depth_to_other(s->h, s->d_depth[p], n_frames, s->cfg.depth_scale, &s->cfg.depth_intrin,
                                          &s->cfg.image_intrin, &s->cfg.depth_to_image, s->d_aligned, s->s_align));
            SCK(s, cudaEventRecord(s->ev_align, s->s_align));
            SPROF(s, 2, s->s_align);
            SCK(s, cudaStreamWaitEvent(s->s_main, s->ev_align, 0));
            SCK(s, cudaStreamWaitEvent(s->s_main, s->ev_in[p], 0));                         // the pose matrices
            SPROF(s, 4, s->s_main);
            SCK(s, cudaGraphLaunch(g->exec, s->s_main));
            orbb::note_replay(s->h, n_frames, g->launches);
            SPROF(s, 5, s->s_main);
            SPROF(s, 6, s->s_main);
            SCK(s, cudaEventRecord(s->ev_out[p], s->s_main));
            SCK(s, cudaEventRecord(s->ev_gate, s->s_main));  // for the next batch's alignment
            s->gate_recorded = true;
            H.n_frames = n_frames;
            s->carry_from = n_frames;
            s->n_submitted++;
            return ticket;
        }
    }
    // ---- depth alignment on its own stream; the aligned buffer is free once the previous batch's gate has read it
    SCK(s, cudaStreamWaitEvent(s->s_align, s->ev_in[p], 0));
    if (s->gate_recorded) SCK(s, cudaStreamWaitEvent(s->s_align, s->ev_gate, 0));
    SRC(orbb_align_depth_to_other(s->h, s->d_depth[p], n_frames, s->cfg.depth_scale, &s->cfg.depth_intrin,
                                  &s->cfg.image_intrin, &s->cfg.depth_to_image, s->d_aligned, s->s_align));
    SCK(s, cudaEventRecord(s->ev_align, s->s_align));
    SPROF(s, 2, s->s_align);
    // ---- extraction
    SCK(s, cudaStreamWaitEvent(s->s_main, s->ev_in[p], 0));
    SRC(orbb_extract_batch_device(s->h, s->d_gray[p], (size_t)s->cfg.image_intrin.width, s->gray_bytes, n_frames,
                                  s->d_kp_raw, s->d_desc_raw, s->d_counts_raw, s->max_kp, s->s_main));
    SPROF(s, 3, s->s_main);
    // ---- depth gate + 3-D lift into rows 1..n (the previous batch's D2H must have drained them)
    SCK(s, cudaStreamWaitEvent(s->s_main, s->ev_align, 0));
    if (ticket >= 1) SCK(s, cudaStreamWaitEvent(s->s_main, s->ev_out[p ^ 1], 0));
    // From here on the previous batch's D2H of the result block has completed, so the block may be written:
    // first the carry (the previous batch's last frame becomes row 0, read by this batch's reprojection / match
    // only), then this batch's counts and rows 1..n.
    SCK(s, orbb::launch_stage_carry(s->d_kp, s->d_desc, s->d_pts, s->d_valid, s->carry_from, s->max_kp, s->d_counts_blk,
                                    s->d_counts_raw, n_frames, s->s_main));
    SRC(orbb_keypoint_pixel_to_point(s->h, s->d_aligned, &s->cfg.image_intrin, n_frames, s->d_kp_raw, s->d_desc_raw,
                                     s->d_counts_raw, s->max_kp, s->d_kp + mk, s->d_desc + 32 * mk, s->d_pts + 3 * mk,
                                     s->d_valid + 1, s->s_main));
    SCK(s, cudaEventRecord(s->ev_gate, s->s_main));
    SPROF(s, 4, s->s_main);
    s->gate_recorded = true;
    // ---- previous (rows 0..n-1) -> current (rows 1..n): reproject, windowed match, compact the 3-D pairs
    SRC(orbb_reproject_points(s->h, s->d_pts, s->d_valid, n_frames, s->max_kp, h_T ? s->d_T[p] : nullptr,
                              &s->cfg.image_intrin, s->d_pos, s->s_main));
    SRC(orbb_match_windowed_batch(s->h, s->d_desc, s->d_pos, s->d_valid, s->d_desc + 32 * mk, s->d_kp + mk,
                                  (int)sizeof(orbb_keypoint), s->d_valid + 1, n_frames, s->max_kp, s->cfg.max_pixel_distance,
                                  s->cfg.max_hamming_distance, s->d_idx, s->d_dist, s->d_pts, s->d_pts + 3 * mk,
                                  s->d_prev_m, s->d_curr_m, s->d_xy, s->d_nm, s->s_main));
    SCK(s, cudaEventRecord(s->ev_main[p], s->s_main));
    SPROF(s, 5, s->s_main);
    // ---- results to the host
    SCK(s, cudaStreamWaitEvent(s->s_out, s->ev_main[p], 0));
    if (n_frames == s->B) {
        SCK(s, cudaMemcpyAsync(s->h_block[p], s->d_block, s->block_bytes, cudaMemcpyDeviceToHost, s->s_out));
    } else {  // partial batch: only the rows in use
        SCK(s, cudaMemcpyAsync(H.counts, s->d_counts_blk, sizeof(int) * n, cudaMemcpyDeviceToHost, s->s_out));
        SCK(s, cudaMemcpyAsync(H.valid, s->d_valid + 1, sizeof(int) * n, cudaMemcpyDeviceToHost, s->s_out));
        SCK(s, cudaMemcpyAsync(H.matched, s->d_nm, sizeof(int) * n, cudaMemcpyDeviceToHost, s->s_out));
        SCK(s, cudaMemcpyAsync(H.kp, s->d_kp + mk, sizeof(orbb_keypoint) * n * mk, cudaMemcpyDeviceToHost, s->s_out));
        SCK(s, cudaMemcpyAsync(H.desc, s->d_desc + 32 * mk, 32 * n * mk, cudaMemcpyDeviceToHost, s->s_out));
        SCK(s, cudaMemcpyAsync(H.pts, s->d_pts + 3 * mk, sizeof(double) * 3 * n * mk, cudaMemcpyDeviceToHost, s->s_out));
        SCK(s, cudaMemcpyAsync(H.prev_m, s->d_prev_m, sizeof(double) * 3 * n * mk, cudaMemcpyDeviceToHost, s->s_out));
        SCK(s, cudaMemcpyAsync(H.curr_m, s->d_curr_m, sizeof(double) * 3 * n * mk, cudaMemcpyDeviceToHost, s->s_out));
        SCK(s, cudaMemcpyAsync(H.xy, s->d_xy, sizeof(uint16_t) * 2 * n * mk, cudaMemcpyDeviceToHost, s->s_out));
    }
    SCK(s, cudaEventRecord(s->ev_out[p], s->s_out));
    SPROF(s, 6, s->s_out);
    H.n_frames = n_frames;
    // ---- carry: the batch's last frame (row n) becomes row 0 at the START of the next submit, after that submit's
    // wait for this batch's D2H -- copying it here would write row 0 of the block while the D2H above reads it
    s->carry_from = n_frames;
    s->n_submitted++;
    return ticket;
}

extern "C" int orbb_rgbd_stage_wait(orbb_rgbd_stage *s, int ticket, orbb_slam_frames *out) {
    if (!s || ticket < 0 || ticket >= s->n_submitted) return ORBB_ERR_INVALID;
    if (ticket < s->n_submitted - 2) return ORBB_ERR_INVALID;  // its buffers have been reused
    SCK(s, cudaSetDevice(s->device));
    const int p = ticket & 1;
    SCK(s, cudaEventSynchronize(s->ev_out[p]));
    if (s->prof && ticket == s->n_submitted - 1 && (ticket % 50) == 49) {
        cudaEventSynchronize(s->pe[6]);
        float t[7] = {0};
        for (int k = 1; k < 7; ++k) cudaEventElapsedTime(&t[k], s->pe[0], s->pe[k]);
        fprintf(stderr, "stage prof (us since H2D start): h2d %.1f | align %.1f | extract %.1f | gate %.1f | match %.1f | d2h %.1f\n",
                1e3f * t[1], 1e3f * t[2], 1e3f * t[3], 1e3f * t[4], 1e3f * t[5], 1e3f * t[6]);
    }
    if (out) {
        const orbb_rgbd_stage::Host &H = s->host[p];
        out->n_frames = H.n_frames; out->max_kp = s->max_kp;
        out->keypoints_count = H.counts; out->valid_keypoints_num = H.valid; out->matched_keypoints_num = H.matched;
        out->keypoints = H.kp; out->descriptors = H.desc; out->points = H.pts;
        out->previous_matched_points = H.prev_m; out->current_matched_points = H.curr_m; out->matched_xy = H.xy;
    }
    return ORBB_OK;
}
