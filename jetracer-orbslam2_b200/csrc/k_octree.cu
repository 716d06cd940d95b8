// k_octree.cu -- DistributeOctTree keypoint selection, one CTA per (frame, level).
// Replaces the "best response per 32x32 cell" selection of the reference's grid NMS
// (src/cuda/nms.cu:86-254) with upstream ORB-SLAM2 semantics.
//
// Upstream is a sequential std::list algorithm.  The B200 formulation removes the list, and in the common case
// the keys as well (see "PYRAMID PATH" below: the FAST kernel bins the candidates into the tree cells of depth
// 4-6, and the selection is read off a count / best-key pyramid over those cells).  The general path, also the
// fallback when the selection has to part keys below the table's depth:
//  * every split line of the tree depends only on the node rectangle, so the tree is a tensor
//    product of two 1-D binary trees; the host tabulates, per coordinate, the Morton-spread path
//    bits (xkey/ykey).  key = xkey[x] | ykey[y] is the key's full root-to-leaf path.
//  * a CTA-wide LSD radix sort of the keys (8-bit digits, warp-private histograms, match.any
//    ranking, block-wide prefix sums) makes every tree node a contiguous segment.
//  * sd[i] = depth at which sorted neighbours i,i+1 part.  Histograms of sd give the node count
//    L(k) and the expandable-node count E(k) of every breadth-first pass at once, so the whole
//    "split everything" phase collapses to picking the stopping depth.
//  * the order-dependent tail (upstream sorts expandable nodes by (size, pointer) and splits the
//    largest first until N nodes exist) is replayed exactly: nodes are ranked by (size, creation
//    sequence) with a shared-memory bitonic sort; a prefix sum of the per-node gains finds the cut.
//    The creation sequence of a breadth-first pass has a closed form (alternating digit complement
//    of the path prefix, because upstream pushes children to the list front and walks forward).
//  * each final node keeps its max-response key, first in upstream candidate order on ties
//    (64-bit shared-memory atomicMax over (response, ~order, index)).
// Bound: latency/issue (tiny per-CTA working sets, L2 resident); not HBM.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "orbb_internal.cuh"

namespace orbb {

#ifdef ORBB_OCT_PROF  // diagnostics build only (make EXTRA=-DORBB_OCT_PROF): phase timestamps of CTA (0,0)
#define OCT_T(k) do { if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) prof_t[k] = clock64(); } while (0)
#else
#define OCT_T(k) do { } while (0)
#endif

#define OCT_THREADS 512
#define OCT_WARPS (OCT_THREADS / 32)
#define FULL 0xffffffffu

// Count / best pyramids of the PYRAMID PATH (see the kernel): every tree level above the cell table, deepest level
// first: level k (R * 4^k nodes) starts at node offset (T - R * 4^(k+1)) / 3, so every level is 4-node aligned and a
// node's four children are one 16-byte (counts) or two 16-byte (best keys) shared-memory loads.  (T - R) / 3 <= 1365.
#define OCT_PYR_NODES 1368
struct __align__(16) OctPyr {
    unsigned long long pb[OCT_PYR_NODES];  // best key below the node
    uint32_t pc[OCT_PYR_NODES];            // keys below the node
    uint16_t pinfo[OCT_PYR_NODES];         // bits 0..2: non-empty children; bits 3..15: 1 + split rank (0 = not split)
};
struct __align__(16) OctStatic {
    union {  // the general path (radix sort) only ever runs after the pyramid path has given up
        uint32_t whist[OCT_WARPS][256];
        OctPyr pyr;
    };
    uint32_t digit_base[256];
    int hist_sd[40], hist_g[40];
    uint32_t warp_tot[OCT_WARPS];
    int ctl[16];
};
enum { C_MODE = 0, C_DEPTH, C_SIZE, C_PN, C_QN, C_CUT, C_TOTAL, C_NFINAL };

// exclusive block scan of one value per thread; returns exclusive prefix, *total = block sum
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *warp_tot, uint32_t *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();  // protect warp_tot reuse
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    // every warp scans the 16 warp totals with shuffles (lanes >= 16 carry zeros)
    const uint32_t wt = lane < OCT_WARPS ? warp_tot[lane] : 0u;
    uint32_t winc = wt;
#pragma unroll
    for (int o = 1; o < OCT_WARPS; o <<= 1) {
        const uint32_t t = __shfl_up_sync(FULL, winc, o);
        if (lane >= o) winc += t;
    }
    const uint32_t tot = __shfl_sync(FULL, winc, OCT_WARPS - 1);
    const uint32_t base = __shfl_sync(FULL, winc - wt, warp);
    *total = tot;
    return base + inc - v;
}

// One LSD pass.  Every warp owns a contiguous slice of the keys (stability).  Loads are issued OCT_RB at a time before
// the dependent match/atomic chain consumes them.  OCT_RB = 4 cuts the L2 round trips a lone CTA waits for (small
// batches: 217 -> 158 us for a 4-frame call); with every SM full of CTAs the other CTAs hide them anyway and the
// extra registers only cost occupancy, so large batches use OCT_RB = 1.
template <int OCT_RB>
__device__ __forceinline__ void radix_pass(const uint2 *__restrict__ kvin, uint2 *__restrict__ kvout, int n, int shift,
                                           OctStatic &S) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    for (int i = tid; i < OCT_WARPS * 256; i += OCT_THREADS) (&S.whist[0][0])[i] = 0;
    __syncthreads();
    const int seglen = ((n + OCT_THREADS - 1) / OCT_THREADS) * 32;  // per-warp contiguous slice
    const int s0 = min(n, warp * seglen), s1 = min(n, s0 + seglen);
    // digit histogram: warp-private counters, plain RED.ADD per lane (order is irrelevant for counts)
    for (int base = s0; base < s1; base += 32 * OCT_RB) {
        uint32_t kk[OCT_RB];
#pragma unroll
        for (int j = 0; j < OCT_RB; ++j) {
            const int i = base + 32 * j + lane;
            kk[j] = i < s1 ? kvin[i].x : 0u;
        }
#pragma unroll
        for (int j = 0; j < OCT_RB; ++j)
            if (base + 32 * j + lane < s1) atomicAdd(&S.whist[warp][(kk[j] >> shift) & 255u], 1u);
    }
    __syncthreads();
    uint32_t tot = 0;
    if (tid < 256) {
#pragma unroll
        for (int w = 0; w < OCT_WARPS; ++w) {
            const uint32_t t = S.whist[w][tid];
            S.whist[w][tid] = tot;
            tot += t;
        }
    }
    uint32_t dummy;
    const uint32_t ex = block_excl_scan(tid < 256 ? tot : 0u, S.warp_tot, &dummy);
    if (tid < 256) S.digit_base[tid] = ex;
    __syncthreads();
    // stable scatter: rank inside the warp step by MATCH.ANY, running offset per (warp, digit) in shared memory
    for (int base = s0; base < s1; base += 32 * OCT_RB) {
        uint2 kv[OCT_RB];
#pragma unroll
        for (int j = 0; j < OCT_RB; ++j) {
            const int i = base + 32 * j + lane;
            kv[j] = i < s1 ? kvin[i] : make_uint2(0u, 0u);
        }
#pragma unroll
        for (int j = 0; j < OCT_RB; ++j) {
            if (base + 32 * j >= s1) break;  // warp-uniform
            const bool valid = base + 32 * j + lane < s1;
            const unsigned d = valid ? ((kv[j].x >> shift) & 255u) : (0x1000u + lane);
            const unsigned peers = __match_any_sync(FULL, d);
            // every lane of the group reads the digit's running offset, the group leader advances it: one writer
            // per digit and step, steps are warp-synchronous, so no atomic and no broadcast are needed
            const uint32_t old = valid ? S.whist[warp][d] : 0u;
            __syncwarp();
            if (valid && (peers & lt) == 0) S.whist[warp][d] = old + (uint32_t)__popc(peers);
            __syncwarp();
            if (valid) kvout[S.digit_base[d] + old + __popc(peers & lt)] = kv[j];  // one 8-byte store per key
        }
    }
    __syncthreads();
}

// Node numbering from the head flags (1 = this key starts a node).  Every warp owns a contiguous slice of the keys;
// inside a slice the rank of a key is a ballot prefix, across slices one scan of the 16 slice totals.  emit(i, node,
// is_head) is called once per key with the 0-based node index.  Returns the number of nodes.  Two block barriers per
// call (the chunked block scan it replaces needed two per 512 keys).
template <typename F>
__device__ __forceinline__ uint32_t head_scan(const uint8_t *__restrict__ head, int n, OctStatic &S, F emit) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const int seglen = ((n + OCT_THREADS - 1) / OCT_THREADS) * 32;
    const int s0 = min(n, warp * seglen), s1 = min(n, s0 + seglen);
    uint32_t cnt = 0;
    for (int base = s0; base < s1; base += 32) {
        const int i = base + lane;
        cnt += __popc(__ballot_sync(FULL, i < s1 && head[i] != 0));
    }
    __syncthreads();  // protect warp_tot reuse
    if (lane == 0) S.warp_tot[warp] = cnt;
    __syncthreads();
    const uint32_t wt = lane < OCT_WARPS ? S.warp_tot[lane] : 0u;
    uint32_t winc = wt;
#pragma unroll
    for (int o = 1; o < OCT_WARPS; o <<= 1) {
        const uint32_t t = __shfl_up_sync(FULL, winc, o);
        if (lane >= o) winc += t;
    }
    const uint32_t total = __shfl_sync(FULL, winc, OCT_WARPS - 1);
    uint32_t run = __shfl_sync(FULL, winc - wt, warp);  // heads before this warp's slice
    for (int base = s0; base < s1; base += 32) {
        const int i = base + lane;
        const bool h = i < s1 && head[i] != 0;
        const unsigned bal = __ballot_sync(FULL, h);
        if (i < s1) emit(i, run + __popc(bal & lt) + (h ? 1u : 0u) - 1u, h);
        run += __popc(bal);
    }
    return total;
}

// bitonic sort, DESCENDING by key, m = power of two.  m <= OCT_THREADS: one element per thread in registers,
// strides < 32 with warp shuffles and only the wide strides through shared memory; larger m: all in smem.
__device__ __forceinline__ void bitonic_desc(unsigned long long *key, uint32_t *val, int m) {
    if (m <= OCT_THREADS) {
        const int i = threadIdx.x;
        unsigned long long a = i < m ? key[i] : 0ull;
        uint32_t v = i < m ? val[i] : 0u;
        for (int k = 2; k <= m; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                unsigned long long b;
                uint32_t w;
                if (j >= 32) {
                    __syncthreads();
                    if (i < m) { key[i] = a; val[i] = v; }
                    __syncthreads();
                    b = i < m ? key[i ^ j] : 0ull;
                    w = i < m ? val[i ^ j] : 0u;
                } else {
                    b = __shfl_xor_sync(FULL, a, j);
                    w = __shfl_xor_sync(FULL, v, j);
                }
                const bool lower = (i & j) == 0, desc = (i & k) == 0;
                // the lower index of a pair keeps the larger key in a descending run
                const bool take = (lower == desc) ? (a < b) : (a > b);
                if (take) { a = b; v = w; }
            }
        }
        __syncthreads();
        if (i < m) { key[i] = a; val[i] = v; }
        __syncthreads();
        return;
    }
    for (int k = 2; k <= m; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < m; i += OCT_THREADS) {
                const int p = i ^ j;
                if (p > i) {
                    const bool desc = (i & k) == 0;
                    const unsigned long long a = key[i], b = key[p];
                    if (desc ? (a < b) : (a > b)) {
                        key[i] = b; key[p] = a;
                        const uint32_t t = val[i]; val[i] = val[p]; val[p] = t;
                    }
                }
            }
            __syncthreads();
        }
    }
}

// ---- PYRAMID PATH (default).  The selection only ever looks at the tree down to the depth at which ~N nodes exist
// (3-5 for the quotas of a pyramid level), and it looks at the keys only through (a) how many lie below a tree node and
// (b) which is the best one there.  The FAST kernel bins every candidate it emits into the 4^Dc tree cells of depth Dc
// (L.tbl_cnt / L.tbl_best in global memory, cell index = path prefix); this CTA sums the table up the tree and reads
// everything off the pyramid -- no sort, no per-key pass, no per-record pass:
//   * L(k) = non-empty nodes of level k is the list size after upstream's k-th breadth-first pass, E(k) = nodes holding
//     more than one key is what that pass could still expand: the pass loop stops at the first k with L >= N,
//     L == L(k-1) or L + 3 E > N;
//   * the careful phase ranks the expandable nodes of a level by (count, creation sequence), splits the largest first
//     until N nodes exist (gain of a split = non-empty children - 1) and goes one level down if all of them were split;
//   * a node is final iff it is non-empty, not split, and sits at the stop depth or below a split parent; its keypoint
//     is the best key below it.  Final nodes are numbered in path order through a bit mask over the cells.
// tests/test_octree_pyramid_model.py restates this in numpy and checks it against the oracle on the CPU.
// Returns false (block-uniform) when the answer lies below depth Dc or the table does not describe this candidate
// list: the CTA then takes the general sorted-key path.  The caller clears the table afterwards.
__device__ __forceinline__ bool octree_pyramid(const LevelDev &L, OctStatic &S, const uint32_t *tbl_cnt, const unsigned long long *tbl_best,
                                               int n, int N, unsigned long long *skey, uint32_t *sval, int pcap2,
                                               uint32_t *sel, int *out_count, long long *prof_t) {
    (void)prof_t;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const int Dc = L.tbl_dc, T = L.tbl_cells, rb = L.key_bits - 2 * L.depth, R = 1 << rb;
    OctPyr &P = S.pyr;
    auto lvl_off = [&](int k) -> int { return (T - (R << (2 * (k + 1)))) / 3; };
    // ---- A. level Dc-1 from the table: one node (four cells) per thread and step, all loads up front
    const int M1 = T >> 2;
    uint4 w[2];
    ulonglong2 b[2][2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int j = tid + q * OCT_THREADS;
        w[q] = make_uint4(0u, 0u, 0u, 0u);
        b[q][0] = b[q][1] = make_ulonglong2(0ull, 0ull);
        if (j < M1) {
            w[q] = reinterpret_cast<const uint4 *>(tbl_cnt)[j];
            b[q][0] = reinterpret_cast<const ulonglong2 *>(tbl_best)[2 * j];
            b[q][1] = reinterpret_cast<const ulonglong2 *>(tbl_best)[2 * j + 1];
        }
    }
    if (tid < 40) { S.hist_sd[tid] = 0; S.hist_g[tid] = 0; }  // hist_sd[k] = L(k), hist_g[k] = E(k)
    if (tid < 128) S.digit_base[tid] = 0u;                    // bit mask of the cells a final node starts at
    __syncthreads();
    {
        unsigned acc_c = 0, acc_1 = 0;  // (L | E << 16) of the cell level and of level Dc-1
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int j = tid + q * OCT_THREADS;
            if (j < M1) {
                const uint32_t c1 = w[q].x + w[q].y + w[q].z + w[q].w;
                const unsigned nz = (w[q].x != 0) + (w[q].y != 0) + (w[q].z != 0) + (w[q].w != 0);
                const unsigned ne = (w[q].x > 1) + (w[q].y > 1) + (w[q].z > 1) + (w[q].w > 1);
                P.pc[j] = c1;
                P.pb[j] = max(max(b[q][0].x, b[q][0].y), max(b[q][1].x, b[q][1].y));
                P.pinfo[j] = (uint16_t)nz;
                acc_c += nz | (ne << 16);
                acc_1 += (c1 != 0 ? 1u : 0u) | (c1 > 1 ? 0x10000u : 0u);
            }
        }
        acc_c = __reduce_add_sync(FULL, acc_c);
        acc_1 = __reduce_add_sync(FULL, acc_1);
        if (lane == 0) {
            if (acc_c) { atomicAdd(&S.hist_sd[Dc], (int)(acc_c & 0xffffu)); atomicAdd(&S.hist_g[Dc], (int)(acc_c >> 16)); }
            if (acc_1) { atomicAdd(&S.hist_sd[Dc - 1], (int)(acc_1 & 0xffffu)); atomicAdd(&S.hist_g[Dc - 1], (int)(acc_1 >> 16)); }
        }
    }
    __syncthreads();
    // ---- B. the levels above: block-wide while a level has more than 64 nodes, then warp 0 alone
    auto level_up = [&](int k, int j) -> unsigned {  // node j of level k from its four children; returns L | E << 16
        const int src = lvl_off(k + 1) + 4 * j, dst = lvl_off(k) + j;
        const uint4 c4 = *reinterpret_cast<const uint4 *>(&P.pc[src]);
        const ulonglong2 b0 = *reinterpret_cast<const ulonglong2 *>(&P.pb[src]), b1 = *reinterpret_cast<const ulonglong2 *>(&P.pb[src + 2]);
        const uint32_t c = c4.x + c4.y + c4.z + c4.w;
        P.pc[dst] = c;
        P.pb[dst] = max(max(b0.x, b0.y), max(b1.x, b1.y));
        P.pinfo[dst] = (uint16_t)((c4.x != 0) + (c4.y != 0) + (c4.z != 0) + (c4.w != 0));
        return (c != 0 ? 1u : 0u) | (c > 1 ? 0x10000u : 0u);
    };
    int k = Dc - 2;
    for (; k >= 0 && (R << (2 * k)) > 64; --k) {
        const int M = R << (2 * k);  // <= 256
        unsigned acc = tid < M ? level_up(k, tid) : 0u;
        if (warp * 32 < M) {
            acc = __reduce_add_sync(FULL, acc);
            if (lane == 0 && acc) { atomicAdd(&S.hist_sd[k], (int)(acc & 0xffffu)); atomicAdd(&S.hist_g[k], (int)(acc >> 16)); }
        }
        __syncthreads();
    }
    if (warp == 0) {
        for (; k >= 0; --k) {
            const int M = R << (2 * k);  // <= 64
            unsigned acc = 0;
            for (int j = lane; j < M; j += 32) acc += level_up(k, j);
            acc = __reduce_add_sync(FULL, acc);
            if (lane == 0) { S.hist_sd[k] = (int)(acc & 0xffffu); S.hist_g[k] = (int)(acc >> 16); }
            __syncwarp();
        }
    }
    __syncthreads();
    OCT_T(1);
    // ---- C. replay of the breadth-first passes (every thread, same answer)
    {
        uint32_t tot = 0;
        for (int r = 0; r < R; ++r) tot += P.pc[lvl_off(0) + r];
        if (tot != (uint32_t)n) return false;  // the table must describe exactly this candidate list
    }
    int mode = -1, k0 = 0, size = 0;
    {
        int prev = S.hist_sd[0];
        for (int kk = 1; kk <= Dc; ++kk) {
            const int Lk = S.hist_sd[kk], Ek = S.hist_g[kk];
            if (Lk >= N || Lk == prev) { mode = 0; k0 = kk; size = Lk; break; }
            if (Lk + 3 * Ek > N) { mode = 1; k0 = kk; size = Lk; break; }
            prev = Lk;
        }
    }
    if (mode < 0) return false;  // the passes go below the table
    // ---- D. careful phase: one round per level
    int last = k0;  // deepest level that holds final nodes
    if (mode == 1) {
        const uint32_t digit_mask = 0xCCCCCCCCu & ((1u << (2 * k0)) - 1u);
        const uint32_t root_mask = (k0 & 1) ? 0u : ((uint32_t)(R - 1) << (2 * k0));
        for (int d = k0;; ++d) {
            if (d + 1 > Dc) return false;  // this round parts at depth d + 1
            const int Md = R << (2 * d), base = lvl_off(d), pbase = d > 0 ? lvl_off(d - 1) : 0;
            if (Md <= 256 && Md <= pcap2) {
                // Small level (the usual case: 64 or 128 nodes): no list, no sort, no scan.  Upstream's processing order is
                // descending (count, creation sequence); what the round needs per node is how much the nodes processed
                // BEFORE it have added to the list -- G = sum of their gains (the node is split iff size + G < N) -- and its
                // rank among them (the creation order of its children).  Both are sums over the nodes with a larger key:
                // G_ threads per node each compare it with a slice of the level, one shuffle reduction per group.  The
                // same sums give the level's total gain, so the round ends without a block-wide reduction.
                if (tid < Md) {
                    const uint32_t cnt = P.pc[base + tid];
                    const unsigned info = P.pinfo[base + tid];
                    bool live;
                    uint32_t seq;
                    if (d == k0) {
                        live = cnt > 1;
                        seq = ((uint32_t)tid ^ digit_mask ^ root_mask) & 0x7fffffffu;
                    } else {
                        const unsigned pr = P.pinfo[pbase + (tid >> 2)] >> 3;
                        live = cnt > 1 && pr != 0;
                        seq = (pr - 1u) * 4u + ((unsigned)tid & 3u);
                    }
                    // (count, sequence) is unique per live node, so the gain in the low bits never decides a comparison
                    skey[tid] = live ? ((unsigned long long)cnt << 40) | ((unsigned long long)seq << 3) | ((info & 7u) - 1u) : 0ull;
                }
                __syncthreads();
                const int G_ = min(32, min(OCT_THREADS / Md, Md)), Sl = Md / G_;
                const int node = min(tid / G_, Md - 1), part = tid & (G_ - 1);
                const unsigned long long my = skey[node];
                uint32_t rank = 0, gsum = 0, gall = 0;
                for (int i = 0; i < Sl; ++i) {
                    const unsigned long long o = skey[i * G_ + part];  // the group's lanes read consecutive keys
                    const uint32_t g = (uint32_t)o & 7u;
                    gall += g;
                    if (o > my) { ++rank; gsum += g; }
                }
                for (int o = 1; o < G_; o <<= 1) {
                    rank += __shfl_xor_sync(FULL, rank, o);
                    gsum += __shfl_xor_sync(FULL, gsum, o);
                    gall += __shfl_xor_sync(FULL, gall, o);
                }
                const bool is_live = my != 0ull && tid / G_ < Md;
                const bool split = is_live && (int)(size + gsum) < N;
                const bool any_live = __syncthreads_or(is_live);
                if (split && part == 0) P.pinfo[base + node] |= (uint16_t)((rank + 1u) << 3);
                __syncthreads();
                if (any_live) last = d + 1;
                // the cut falls inside this level iff splitting all of it reaches N; otherwise everything was split
                if (!any_live || (int)(size + gall) >= N || gall == 0) break;
                size += (int)gall;
                continue;
            }
            if (tid == 0) { S.ctl[C_PN] = 0; S.ctl[C_CUT] = 0x7fffffff; }
            __syncthreads();
            // expandable nodes of this level -> sort list keyed by (count, creation sequence); list order is irrelevant
            for (int j0 = 0; j0 < Md; j0 += OCT_THREADS) {  // block-uniform trip count (<= 2)
                const int j = j0 + tid;
                uint32_t cnt = 0, seq = 0;
                bool live = false;
                if (j < Md) {
                    cnt = P.pc[base + j];
                    if (d == k0) {  // closed-form creation sequence of the breadth-first pass that made these nodes
                        live = cnt > 1;
                        seq = ((uint32_t)j ^ digit_mask ^ root_mask) & 0x7fffffffu;
                    } else {        // children of the nodes the previous round split, in processing order of the parents
                        const unsigned pr = P.pinfo[pbase + (j >> 2)] >> 3;
                        live = cnt > 1 && pr != 0;
                        seq = (pr - 1u) * 4u + ((unsigned)j & 3u);
                    }
                }
                const unsigned bal = __ballot_sync(FULL, live);
                int p0 = 0;
                if (lane == 0 && bal) p0 = atomicAdd(&S.ctl[C_PN], __popc(bal));
                p0 = __shfl_sync(FULL, p0, 0);
                if (live) {
                    const int p = p0 + __popc(bal & lt);
                    if (p < pcap2) { skey[p] = ((unsigned long long)cnt << 32) | seq; sval[p] = (uint32_t)j; }
                }
            }
            __syncthreads();
            const int pn = S.ctl[C_PN];
            if (pn > pcap2) return false;  // cannot happen (pn <= list size < N <= pcap); keeps the sort in bounds
            int mm = 1;
            while (mm < pn) mm <<= 1;
            for (int p = pn + tid; p < mm; p += OCT_THREADS) { skey[p] = 0ull; sval[p] = 0xffffffffu; }
            __syncthreads();
            bitonic_desc(skey, sval, mm);
            // prefix sums of the gains in processing order; first rank at which the node count reaches N
            uint32_t carry = 0;
            for (int r0 = 0; r0 < pn; r0 += OCT_THREADS) {
                const int r = r0 + tid;
                const uint32_t g = r < pn ? (uint32_t)(P.pinfo[base + sval[r]] & 7u) - 1u : 0u;
                uint32_t tot;
                const uint32_t inc = block_excl_scan(g, S.warp_tot, &tot) + carry + g;
                if (r < pn) {
                    skey[r] = inc;  // reuse: inclusive gain prefix at rank r
                    if ((int)(size + inc) >= N) atomicMin(&S.ctl[C_CUT], r);
                }
                carry += tot;
            }
            __syncthreads();
            const int cut = S.ctl[C_CUT];
            const bool found = cut != 0x7fffffff;
            const int nsplit = found ? cut + 1 : pn;
            const uint32_t total = nsplit > 0 ? (uint32_t)skey[nsplit - 1] : 0u;
            for (int r = tid; r < nsplit; r += OCT_THREADS) P.pinfo[base + sval[r]] |= (uint16_t)((r + 1) << 3);
            __syncthreads();
            if (nsplit > 0) last = d + 1;
            if (found || total == 0) break;
            size += (int)total;
        }
    }
    OCT_T(2);
    // ---- E. final nodes -> bit mask over their first cells -> path-order index -> selection
    uint32_t *fmask = S.digit_base, *fpre = S.digit_base + 128;
    auto for_final = [&](auto &&f) {  // f(first cell, best key) for every final node this thread owns
        for (int kk = k0; kk <= min(last, Dc - 1); ++kk) {
            const int Mk = R << (2 * kk), base = lvl_off(kk), pbase = kk > 0 ? lvl_off(kk - 1) : 0;
            for (int j = tid; j < Mk; j += OCT_THREADS) {
                const bool fin = P.pc[base + j] != 0 && (P.pinfo[base + j] >> 3) == 0 &&
                                 (kk == k0 || (P.pinfo[pbase + (j >> 2)] >> 3) != 0);
                if (fin) f(j << (2 * (Dc - kk)), P.pb[base + j]);
            }
        }
        if (last >= Dc) {  // cells as nodes (rare): read them again, the table is still intact
            for (int j = tid; j < M1; j += OCT_THREADS)
                if (k0 == Dc || (P.pinfo[j] >> 3) != 0) {
                    const uint4 c4 = reinterpret_cast<const uint4 *>(tbl_cnt)[j];
                    const ulonglong2 b0 = reinterpret_cast<const ulonglong2 *>(tbl_best)[2 * j];
                    const ulonglong2 b1 = reinterpret_cast<const ulonglong2 *>(tbl_best)[2 * j + 1];
                    if (c4.x) f(4 * j, b0.x);
                    if (c4.y) f(4 * j + 1, b0.y);
                    if (c4.z) f(4 * j + 2, b1.x);
                    if (c4.w) f(4 * j + 3, b1.y);
                }
        }
    };
    for_final([&](int c, unsigned long long) { atomicOr(&fmask[c >> 5], 1u << (c & 31)); });
    __syncthreads();
    if (warp == 0) {  // exclusive prefix of the mask words' populations (<= 128 words, four per lane)
        uint32_t m4[4], s = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) { m4[i] = fmask[4 * lane + i]; s += __popc(m4[i]); }
        uint32_t inc = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += t;
        }
        uint32_t run = inc - s;
#pragma unroll
        for (int i = 0; i < 4; ++i) { fpre[4 * lane + i] = run; run += __popc(m4[i]); }
        if (lane == 31) S.ctl[C_NFINAL] = (int)inc;
    }
    __syncthreads();
    const int sel_cap = L.sel_cap;
    for_final([&](int c, unsigned long long bk) {
        const uint32_t idx = fpre[c >> 5] + __popc(fmask[c >> 5] & ((1u << (c & 31)) - 1u));
        if (idx < (uint32_t)sel_cap) sel[idx] = (uint32_t)(bk & 0xffffffull) | ((uint32_t)(bk >> 56) << 24);
    });
    if (tid == 0) *out_count = min(S.ctl[C_NFINAL], sel_cap);
    OCT_T(3);
    return true;
}

template <int OCT_RB, int MIN_CTAS>
__global__ void __launch_bounds__(OCT_THREADS, MIN_CTAS)
k_octree(const LevelDev *__restrict__ levels, int n_levels, const int *__restrict__ cand_count,
         int *__restrict__ sel_count, int level_base, int frame_base, int quota_override, int pcap, int pcap2,
         int sort_off, int sort_bytes, int no_fast) {
    __shared__ OctStatic S;
    extern __shared__ __align__(16) uint8_t dyn[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int level = blockIdx.x + level_base, frame = blockIdx.y + frame_base;
    const LevelDev &L = levels[level];
    const int N = quota_override >= 0 ? quota_override : L.nfeat;
#ifdef ORBB_OCT_PROF
    long long prof_t[12] = {0};
#endif
    OCT_T(0);
    int n = cand_count[frame * n_levels + level];
    n = min(n, L.cand_cap);
    int *out_count = sel_count + frame * n_levels + level;
    if (n <= 0) {
        if (tid == 0) *out_count = 0;
        return;
    }
    // dynamic shared memory carve-up
    unsigned long long *best = reinterpret_cast<unsigned long long *>(dyn);          // [sel_cap]
    unsigned long long *skey = best + L.sel_cap;                                       // [pcap2]
    uint32_t *sval = reinterpret_cast<uint32_t *>(skey + pcap2);                       // [pcap2]
    // per-node arrays, [pcap + 1] each: start (two generations), creation sequence (two generations,
    // 0xffffffff = not expandable this round), gain, processing rank (-1 = not split)
    uint32_t *nst_a = sval + pcap2, *nst_b = nst_a + (pcap + 1);
    uint32_t *seq_a = nst_b + (pcap + 1), *seq_b = seq_a + (pcap + 1);
    uint32_t *gain = seq_b + (pcap + 1);
    int *rank_of = reinterpret_cast<int *>(gain + (pcap + 1));

    const size_t fo = (size_t)frame * L.cand_cap;
    const uint32_t *cand = L.cand + fo;
    const int D = L.depth;

    // The pyramid path (above) answers from the FAST kernel's cell table when the table describes this candidate list
    // and the selection stays at depths <= Dc (it decides that itself: orbb_create sizes the table for 4 x the level's
    // quota, capped at 4096 cells, so a level with a very large quota may still part deeper); otherwise the general
    // sorted-key path below runs.  Either way the CTA clears its table before it exits.
    const int T = L.tbl_cells;
    const bool try_fast = !no_fast && T > 0 && n < L.cand_cap;  // an overflowed list no longer matches its table
    uint32_t *tbl_cnt = L.tbl_cnt + (size_t)frame * T;
    unsigned long long *tbl_best = L.tbl_best + (size_t)frame * T;

    bool done = false;
    if (try_fast && pcap2 <= 8190) {
#ifdef ORBB_OCT_PROF
        done = octree_pyramid(L, S, tbl_cnt, tbl_best, n, N, skey, sval, pcap2, L.sel + (size_t)frame * L.sel_cap, out_count, prof_t);
        if (done && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0)
            printf("oct-pyr n=%d N=%d Dc=%d T=%d | table+pyramid %lld replay+careful %lld final %lld cycles\n", n, N, L.tbl_dc, T,
                   prof_t[1] - prof_t[0], prof_t[2] - prof_t[1], prof_t[3] - prof_t[2]);
#else
        done = octree_pyramid(L, S, tbl_cnt, tbl_best, n, N, skey, sval, pcap2, L.sel + (size_t)frame * L.sel_cap, out_count, nullptr);
#endif
        __syncthreads();  // the general path reuses the shared arrays
    }
    if (!done) {
    int m;               // keys
    uint2 *kva, *kvb;    // kv[i].x = path key, kv[i].y = packed candidate
    int key_stride;
    bool in_smem;
    {
    // (path key, packed candidate) pairs, radix ping-pong.  When the launch could afford the shared memory (small
    // grids: one or two CTAs per SM) and this level's candidates fit, every per-key array lives in shared memory:
    // the scattered 8-byte stores of the sort and the strided passes over the keys then never leave the SM.
    const int n_pad = (n + 15) & ~15;
    in_smem = (size_t)n_pad * 16 <= (size_t)sort_bytes;
    kva = in_smem ? reinterpret_cast<uint2 *>(dyn + sort_off) : L.kv_a + fo;
    kvb = in_smem ? kva + n_pad : L.kv_b + fo;
    key_stride = in_smem ? n_pad : L.cand_cap;  // elements between the scratch arrays carved from kvb
    m = n;

    // ---- 1. path keys
    for (int base = tid; base < n; base += OCT_RB * OCT_THREADS) {
        uint32_t c[OCT_RB], kx[OCT_RB], ky[OCT_RB];
#pragma unroll
        for (int j = 0; j < OCT_RB; ++j) c[j] = base + j * OCT_THREADS < n ? cand[base + j * OCT_THREADS] : 0u;
#pragma unroll
        for (int j = 0; j < OCT_RB; ++j) { kx[j] = __ldg(&L.xkey[c[j] & 0xfffu]); ky[j] = __ldg(&L.ykey[(c[j] >> 12) & 0xfffu]); }
#pragma unroll
        for (int j = 0; j < OCT_RB; ++j) {
            const int i = base + j * OCT_THREADS;
            if (i < n) kva[i] = make_uint2(kx[j] | ky[j], c[j]);  // the candidate travels with its key: no gathers later
        }
    }
    __syncthreads();
    OCT_T(1);
    // ---- 2. LSD radix sort by path
    for (int shift = 0; shift < L.key_bits; shift += 8) {
        radix_pass<OCT_RB>(kva, kvb, n, shift, S);
        uint2 *t = kva; kva = kvb; kvb = t;
    }
    }
    const uint2 *kv = kva;  // sorted
    OCT_T(2);
    auto nkeys = [&](uint32_t a, uint32_t b) -> uint32_t { return b - a; };
    // free ping-pong buffers become scratch: two generations of u16 segment ids, and the head flags
    uint16_t *seg_a = reinterpret_cast<uint16_t *>(kvb), *seg_b = seg_a + key_stride;
    uint8_t *head = reinterpret_cast<uint8_t *>(seg_b + key_stride);
    uint8_t *sd = in_smem ? head + key_stride : L.sd + fo;  // 2 + 2 + 1 + 1 bytes per key fit the free 8-byte half

    // ---- 3. split depths + histograms
    if (tid < 40) { S.hist_sd[tid] = 0; S.hist_g[tid] = 0; }
    __syncthreads();
    // Per-thread packed 8-bit counters, 16 bins per histogram (depths 0..D <= 15), flushed to the shared histograms
    // once per warp (and every 255 items per thread): corners cluster, so most partings fall into two or three deep
    // bins and per-key shared atomics would serialise on them.
    unsigned long long cs_lo = 0, cs_hi = 0, cg_lo = 0, cg_hi = 0;
    auto flush_counts = [&]() {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const unsigned vs = (unsigned)((k < 8 ? cs_lo : cs_hi) >> (8 * (k & 7))) & 0xffu;
            const unsigned vg = (unsigned)((k < 8 ? cg_lo : cg_hi) >> (8 * (k & 7))) & 0xffu;
            const unsigned ts = __reduce_add_sync(FULL, vs), tg = __reduce_add_sync(FULL, vg);
            if (lane == 0) {
                if (ts) atomicAdd(&S.hist_sd[k], (int)ts);
                if (tg) atomicAdd(&S.hist_g[k], (int)tg);
            }
        }
        cs_lo = cs_hi = cg_lo = cg_hi = 0;
    };
    int since_flush = 0;
    for (int base = 0; base < m; base += OCT_THREADS) {  // block-uniform trip count
        const int i = base + tid;
        if (i < m) {
            const uint32_t kc = kv[i].x;
            unsigned l = 0, r = 0;
            if (i > 0) {
                const int hb = 31 - __clz(kv[i - 1].x ^ kc);
                l = hb >= 2 * D ? 0u : (unsigned)(D - (hb >> 1));
            }
            if (i < m - 1) {
                const int hb = 31 - __clz(kc ^ kv[i + 1].x);
                r = hb >= 2 * D ? 0u : (unsigned)(D - (hb >> 1));
                sd[i] = (uint8_t)r;
                const unsigned rb = min(r, 15u);  // r <= D <= 15 (keys are distinct pixels)
                const unsigned long long one = 1ull << (8 * (rb & 7u));
                if (rb < 8) cs_lo += one; else cs_hi += one;
            }
            // hist_g counts the single-key nodes by the depth at which they become isolated
            {
                const unsigned g = min(max(l, r), 15u);
                const unsigned long long one = 1ull << (8 * (g & 7u));
                if (g < 8) cg_lo += one; else cg_hi += one;
            }
        }
        if (++since_flush == 255) { flush_counts(); since_flush = 0; }
    }
    flush_counts();
    __syncthreads();
    OCT_T(3);
    // ---- 4. replay the breadth-first passes on the histograms
    if (tid == 0) {
        int prev = 1 + S.hist_sd[0], cumL = prev, cumS = S.hist_g[0];
        int mode = 0, depth = D + 1;
        for (int k = 1; k <= D + 1; ++k) {
            if (k <= D) { cumL += S.hist_sd[k]; cumS += S.hist_g[k]; }
            const int Ek = cumL - cumS;
            if (cumL >= N || cumL == prev) { mode = 0; depth = k; break; }
            if (cumL + 3 * Ek > N) { mode = 1; depth = k; break; }
            prev = cumL;
        }
        S.ctl[C_MODE] = mode; S.ctl[C_DEPTH] = depth; S.ctl[C_SIZE] = cumL;
        S.ctl[C_PN] = 0; S.ctl[C_QN] = 0;
    }
    __syncthreads();
    const int mode = S.ctl[C_MODE], k0 = S.ctl[C_DEPTH];
    for (int i = tid; i < m; i += OCT_THREADS) head[i] = (i == 0 || sd[i - 1] <= k0) ? 1 : 0;
    __syncthreads();
    OCT_T(4);
    if (mode == 1) {
        // ---- 5. careful phase, key-parallel.  A round works on the current head-delimited nodes: nst[] start of
        // every node, seq[] creation sequence of the nodes created by the previous round that hold > 1 key
        // (upstream's vSizeAndPointerToNode), seg[] node id of every key.
        int nseg = S.ctl[C_SIZE];  // < N <= pcap
        uint32_t *nst = nst_a, *nst2 = nst_b, *seq = seq_a, *seq2 = seq_b;
        uint16_t *seg = seg_a, *seg2 = seg_b;
        {
            head_scan(head, m, S, [&](int i, uint32_t node, bool h) {
                seg[i] = (uint16_t)node;
                if (h) nst[node] = (uint32_t)i;
            });
            if (tid == 0) nst[nseg] = (uint32_t)m;
            __syncthreads();
            // closed-form creation sequence of the breadth-first pass that made the depth-k0 nodes
            const uint32_t digit_mask = 0xCCCCCCCCu & ((1u << (2 * k0)) - 1u);
            const uint32_t root_mask = (k0 & 1) ? 0u : (((1u << (L.key_bits - 2 * D)) - 1u) << (2 * k0));
            for (int sidx = tid; sidx < nseg; sidx += OCT_THREADS) {
                const uint32_t st = nst[sidx], cnt = nkeys(st, nst[sidx + 1]);
                seq[sidx] = cnt > 1 ? (((kv[st].x >> (2 * (D - k0))) ^ digit_mask ^ root_mask) & 0x7fffffffu) : 0xffffffffu;
            }
        }
        int d = k0;
        while (true) {
            const int size = S.ctl[C_SIZE];
            for (int sidx = tid; sidx < nseg; sidx += OCT_THREADS) { gain[sidx] = 0; rank_of[sidx] = -1; }
            if (tid == 0) { S.ctl[C_PN] = 0; S.ctl[C_CUT] = 0x7fffffff; }
            __syncthreads();
            // gains: children - 1 = number of depth-(d+1) partings inside the node
            for (int i = tid; i < m - 1; i += OCT_THREADS)
                if (sd[i] == d + 1) {
                    const int sidx = seg[i];
                    if (seq[sidx] != 0xffffffffu) atomicAdd(&gain[sidx], 1u);
                }
            // expandable nodes -> sort list keyed by (count, creation sequence)
            for (int sidx = tid; sidx < nseg; sidx += OCT_THREADS)
                if (seq[sidx] != 0xffffffffu) {
                    const int p = atomicAdd(&S.ctl[C_PN], 1);
                    skey[p] = ((unsigned long long)nkeys(nst[sidx], nst[sidx + 1]) << 32) | seq[sidx];
                    sval[p] = (uint32_t)sidx;
                }
            __syncthreads();
            const int pn = S.ctl[C_PN];
            int mm = 1;
            while (mm < pn) mm <<= 1;
            for (int p = pn + tid; p < mm; p += OCT_THREADS) { skey[p] = 0ull; sval[p] = 0xffffffffu; }
            __syncthreads();
            bitonic_desc(skey, sval, mm);
            // prefix sums of the gains in processing order; first rank at which the node count reaches N
            uint32_t carry2 = 0;
            for (int base = 0; base < pn; base += OCT_THREADS) {
                const int r = base + tid;
                const uint32_t g = r < pn ? gain[sval[r]] : 0u;
                uint32_t tot;
                const uint32_t inc = block_excl_scan(g, S.warp_tot, &tot) + carry2 + g;
                if (r < pn) {
                    skey[r] = inc;  // reuse: inclusive gain prefix at rank r
                    if ((int)(size + inc) >= N) atomicMin(&S.ctl[C_CUT], r);
                }
                carry2 += tot;
            }
            __syncthreads();
            const int cut = S.ctl[C_CUT];
            const bool found = cut != 0x7fffffff;
            const int nsplit = found ? cut + 1 : pn;
            const uint32_t total = nsplit > 0 ? (uint32_t)skey[nsplit - 1] : 0u;
            for (int r = tid; r < nsplit; r += OCT_THREADS) rank_of[sval[r]] = r;
            __syncthreads();
            // split: every depth-(d+1) parting inside a split node starts a new node
            for (int i = tid; i < m - 1; i += OCT_THREADS)
                if (sd[i] == d + 1 && rank_of[seg[i]] >= 0) head[i + 1] = 1;
            __syncthreads();
            if (found || total == 0) break;
            // next round: renumber the nodes; the new expandable ones are the children (> 1 key) of split nodes,
            // created in processing order of their parents, n1..n4 inside a parent
            const int nseg2 = size + (int)total;
            head_scan(head, m, S, [&](int i, uint32_t node, bool h) {
                seg2[i] = (uint16_t)node;
                if (h) nst2[node] = (uint32_t)i;
            });
            if (tid == 0) nst2[nseg2] = (uint32_t)m;
            __syncthreads();
            for (int s2 = tid; s2 < nseg2; s2 += OCT_THREADS) {
                const uint32_t st = nst2[s2], cnt = nkeys(st, nst2[s2 + 1]);
                const int pr = rank_of[seg[st]];
                seq2[s2] = (pr >= 0 && cnt > 1) ? ((uint32_t)pr * 4u + ((kv[st].x >> (2 * (D - d - 1))) & 3u)) : 0xffffffffu;
            }
            if (tid == 0) S.ctl[C_SIZE] = nseg2;
            __syncthreads();
            { uint32_t *t = nst; nst = nst2; nst2 = t; t = seq; seq = seq2; seq2 = t; }
            { uint16_t *t = seg; seg = seg2; seg2 = t; }
            nseg = nseg2;
            ++d;
        }
    }
    OCT_T(5);

    // ---- 6. final nodes = head-delimited segments; keep the best key of each
    const uint32_t n_nodes = head_scan(head, m, S, [&](int i, uint32_t node, bool) { seg_a[i] = (uint16_t)min(node, 0xffffu); });
    const int nfinal = min((int)n_nodes, L.sel_cap);
    for (int sidx = tid; sidx < nfinal; sidx += OCT_THREADS) best[sidx] = 0ull;
    __syncthreads();
    for (int b0 = 0; b0 < m; b0 += OCT_RB * OCT_THREADS) {  // block-uniform trip count: the body shuffles
        const int base = b0 + tid;
        uint32_t sidx[OCT_RB], c[OCT_RB], xo[OCT_RB], yo[OCT_RB];
        unsigned long long vv[OCT_RB];
#pragma unroll
        for (int j = 0; j < OCT_RB; ++j) {
            const int i = base + j * OCT_THREADS;
            sidx[j] = i < m ? (uint32_t)seg_a[i] : 0xffffffffu;
            c[j] = i < m ? kv[i].y : 0u;
        }
        {
#pragma unroll
            for (int j = 0; j < OCT_RB; ++j) { xo[j] = __ldg(&L.xord[c[j] & 0xfffu]); yo[j] = __ldg(&L.yord[(c[j] >> 12) & 0xfffu]); }
#pragma unroll
            for (int j = 0; j < OCT_RB; ++j) {
                // (response, earliest upstream candidate order) decides; ord is unique per pixel, so the low 24 bits
                // (x | y << 12 of the winner) never take part in the comparison
                const uint32_t ord = ((yo[j] >> 6) << 19) | ((xo[j] >> 6) << 12) | ((yo[j] & 63u) << 6) | (xo[j] & 63u);
                vv[j] = ((unsigned long long)(c[j] >> 24) << 56) | ((unsigned long long)(0x3ffffffu - ord) << 24) |
                        (c[j] & 0xffffffu);
            }
        }
#pragma unroll
        for (int j = 0; j < OCT_RB; ++j) {
            if (b0 + j * OCT_THREADS >= m) break;  // block-uniform
            // the items are sorted by path, so a node's items are consecutive: segmented max over the warp's 32
            // items, then one atomic per (warp, node) run instead of one per item (64-bit shared atomicMax is a CAS loop)
            const uint32_t sg = sidx[j];
            unsigned long long v = sg < (uint32_t)nfinal ? vv[j] : 0ull;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long ov = __shfl_up_sync(FULL, v, o);
                const uint32_t os = __shfl_up_sync(FULL, sg, o);
                if (lane >= o && os == sg) v = max(v, ov);
            }
            const uint32_t nx = __shfl_down_sync(FULL, sg, 1);
            if (sg < (uint32_t)nfinal && (lane == 31 || nx != sg)) atomicMax(&best[sg], v);
        }
    }
    __syncthreads();
    uint32_t *sel = L.sel + (size_t)frame * L.sel_cap;
    for (int sidx = tid; sidx < nfinal; sidx += OCT_THREADS) {
        const unsigned long long b = best[sidx];
        sel[sidx] = (uint32_t)(b & 0xffffffull) | ((uint32_t)(b >> 56) << 24);
    }
    if (tid == 0) *out_count = nfinal;
    OCT_T(6);
#ifdef ORBB_OCT_PROF
    if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0)
        printf("oct n=%d m=%d N=%d D=%d Dc=%d T=%d mode=%d k0=%d in_smem=%d | build+sort %lld sd %lld replay %lld careful %lld final %lld cycles\n",
               n, m, N, D, L.tbl_dc, T, mode, k0, (int)in_smem, prof_t[2] - prof_t[0], prof_t[3] - prof_t[2],
               prof_t[4] - prof_t[3], prof_t[5] - prof_t[4], prof_t[6] - prof_t[5]);
#endif
    }  // general path
    if (T > 0) {  // hand the FAST kernel of the next batch an empty table (T is a multiple of 4)
        for (int i = tid; i < (T >> 2); i += OCT_THREADS) {
            reinterpret_cast<uint4 *>(tbl_cnt)[i] = make_uint4(0u, 0u, 0u, 0u);
            reinterpret_cast<ulonglong2 *>(tbl_best)[2 * i] = make_ulonglong2(0ull, 0ull);
            reinterpret_cast<ulonglong2 *>(tbl_best)[2 * i + 1] = make_ulonglong2(0ull, 0ull);
        }
    }
}

// The cell table of a candidate list that did not come from the FAST kernel (orbb_debug_distribute uploads one).
__global__ void k_octree_bin(const LevelDev *__restrict__ levels, int n_levels, const int *__restrict__ cand_count, int level,
                             int frame_base) {
    const LevelDev &L = levels[level];
    const int frame = blockIdx.y + frame_base;
    if (!L.tbl_cells) return;
    const int n = min(cand_count[frame * n_levels + level], L.cand_cap - 1);
    const uint32_t *cand = L.cand + (size_t)frame * L.cand_cap;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) oct_bin_candidate(L, frame, cand[i]);
}

cudaError_t launch_octree_bin(const LevelDev *d_levels, int n_levels, const int *d_cand_count, int level, int frame_base,
                              int n_frames, cudaStream_t st) {
    k_octree_bin<<<dim3(32, n_frames), 256, 0, st>>>(d_levels, n_levels, d_cand_count, level, frame_base);
    return cudaGetLastError();
}

size_t octree_dyn_smem(int sel_cap_max, int pcap, int pcap2) {
    return (size_t)sel_cap_max * 8 + (size_t)pcap2 * 12 + (size_t)(pcap + 1) * 4 * 6 + 16;
}

template <int OCT_RB, int MIN_CTAS>
static cudaError_t launch_octree_t(const LevelDev *d_levels, int n_levels, const int *d_cand_count, int *d_sel_count,
                                   int level_base, int n_launch_levels, int frame_base, int n_frames, int quota_override,
                                   size_t smem, int sort_bytes, int no_fast, int pcap, int pcap2, cudaStream_t st) {
    const int sort_off = (int)((smem + 15) & ~(size_t)15);
    const size_t total = (size_t)sort_off + (size_t)sort_bytes;
    if (total > 26 * 1024) {  // 48 KB default limit minus the 20.2 KB of static shared memory; the attribute is per device
        cudaError_t e = cudaFuncSetAttribute(k_octree<OCT_RB, MIN_CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)total);
        if (e != cudaSuccess) return e;
    }
    dim3 grid(n_launch_levels, n_frames);
    k_octree<OCT_RB, MIN_CTAS><<<grid, OCT_THREADS, total, st>>>(d_levels, n_levels, d_cand_count, d_sel_count, level_base,
                                                                 frame_base, quota_override, pcap, pcap2, sort_off, sort_bytes,
                                                                 no_fast);
    return cudaGetLastError();
}

cudaError_t launch_octree(const LevelDev *d_levels, int n_levels, const int *d_cand_count, int *d_sel_count,
                          int level_base, int n_launch_levels, int frame_base, int n_frames, int quota_override,
                          int sel_cap_max, int pcap, int pcap2, cudaStream_t st) {
    size_t smem = octree_dyn_smem(sel_cap_max, pcap, pcap2);
    if (getenv("ORBB_OCT_PAD")) smem = std::max(smem, (size_t)atoi(getenv("ORBB_OCT_PAD")));
    // diagnostics / tests: ORBB_OCT_NOFAST=1 forces the general (radix sort) path of every CTA
    const bool nofast = getenv("ORBB_OCT_NOFAST") != nullptr;
    // fewer CTAs than the GPU can hold at once: every CTA's own latency is the kernel's duration
    const long long ctas = (long long)n_launch_levels * n_frames;
    if (ctas <= 148 * 3) {
        // shared memory left per CTA when the grid is spread over the 148 SMs (227 KB each, 20.2 KB static per CTA)
        const int per_sm = (int)((ctas + 147) / 148);
        long long spare = (227 * 1024) / per_sm - 21 * 1024 - (long long)smem - 1024;
        static const bool no_smem_sort = getenv("ORBB_OCT_NOSMEM") != nullptr;
        const int sort_bytes = (no_smem_sort || spare < 32 * 1024) ? 0 : (int)std::min<long long>(spare, 176 * 1024) & ~15;
        if (ctas <= 148)  // at most one CTA per SM: registers are free, keep 8 loads per thread in flight
            return launch_octree_t<8, 1>(d_levels, n_levels, d_cand_count, d_sel_count, level_base, n_launch_levels, frame_base,
                                         n_frames, quota_override, smem, sort_bytes, nofast, pcap, pcap2, st);
        return launch_octree_t<4, 3>(d_levels, n_levels, d_cand_count, d_sel_count, level_base, n_launch_levels, frame_base,
                                     n_frames, quota_override, smem, sort_bytes, nofast, pcap, pcap2, st);
    }
    return launch_octree_t<1, 4>(d_levels, n_levels, d_cand_count, d_sel_count, level_base, n_launch_levels, frame_base,
                                 n_frames, quota_override, smem, 0, nofast, pcap, pcap2, st);
}

}  // namespace orbb
