// k_octree.cu -- DistributeOctTree keypoint selection, one CTA per (frame, level).
// Replaces the "best response per 32x32 cell" selection of the reference's grid NMS
// (src/cuda/nms.cu:86-254) with upstream ORB-SLAM2 semantics.
//
// Upstream is a sequential std::list algorithm.  The B200 formulation removes the list:
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
#include "orbb_internal.cuh"

namespace orbb {

#define OCT_THREADS 512
#define OCT_WARPS (OCT_THREADS / 32)
#define FULL 0xffffffffu

struct OctStatic {
    uint32_t whist[OCT_WARPS][256];
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

__device__ __forceinline__ void radix_pass(const uint32_t *__restrict__ kin, const uint32_t *__restrict__ iin,
                                           uint32_t *__restrict__ kout, uint32_t *__restrict__ iout, int n,
                                           int shift, OctStatic &S) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    for (int i = tid; i < OCT_WARPS * 256; i += OCT_THREADS) (&S.whist[0][0])[i] = 0;
    __syncthreads();
    const int seglen = ((n + OCT_THREADS - 1) / OCT_THREADS) * 32;  // per-warp contiguous slice
    const int s0 = min(n, warp * seglen), s1 = min(n, s0 + seglen);
    for (int base = s0; base < s1; base += 32) {
        const int i = base + lane;
        const bool valid = i < s1;
        const unsigned d = valid ? ((kin[i] >> shift) & 255u) : (0x1000u + lane);
        const unsigned peers = __match_any_sync(FULL, d);
        if (valid && (peers & lt) == 0) S.whist[warp][d] += __popc(peers);
        __syncwarp();
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
    for (int base = s0; base < s1; base += 32) {
        const int i = base + lane;
        const bool valid = i < s1;
        uint32_t k = 0, id = 0;
        if (valid) { k = kin[i]; id = iin[i]; }
        const unsigned d = valid ? ((k >> shift) & 255u) : (0x1000u + lane);
        const unsigned peers = __match_any_sync(FULL, d);
        uint32_t pos = 0;
        if (valid) pos = S.digit_base[d] + S.whist[warp][d] + __popc(peers & lt);
        __syncwarp();
        if (valid && (peers & lt) == 0) S.whist[warp][d] += __popc(peers);
        __syncwarp();
        if (valid) { kout[pos] = k; iout[pos] = id; }
    }
    __syncthreads();
}

// shared-memory bitonic sort, DESCENDING by key, m = power of two
__device__ __forceinline__ void bitonic_desc(unsigned long long *key, uint32_t *val, int m) {
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

__global__ void __launch_bounds__(OCT_THREADS)
k_octree(const LevelDev *__restrict__ levels, int n_levels, const int *__restrict__ cand_count,
         int *__restrict__ sel_count, int level_base, int frame_base, int quota_override, int pcap, int pcap2) {
    __shared__ OctStatic S;
    extern __shared__ __align__(16) uint8_t dyn[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int level = blockIdx.x + level_base, frame = blockIdx.y + frame_base;
    const LevelDev &L = levels[level];
    const int N = quota_override >= 0 ? quota_override : L.nfeat;
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
    uint32_t *p_start = sval + pcap2, *p_cnt = p_start + pcap, *p_seq = p_cnt + pcap;  // [pcap] each
    uint32_t *p_gain = p_seq + pcap;
    uint32_t *q_start = p_gain + pcap, *q_cnt = q_start + pcap, *q_seq = q_cnt + pcap;
    uint32_t *node_start = q_seq + pcap;                                               // [pcap + 1]

    const size_t fo = (size_t)frame * L.cand_cap;
    const uint32_t *cand = L.cand + fo;
    uint32_t *ka = L.key_a + fo, *kb = L.key_b + fo, *ia = L.idx_a + fo, *ib = L.idx_b + fo;
    uint8_t *sd = L.sd + fo;
    const int D = L.depth;

    // ---- 1. path keys
    for (int i = tid; i < n; i += OCT_THREADS) {
        const uint32_t c = cand[i];
        ka[i] = __ldg(&L.xkey[c & 0xfffu]) | __ldg(&L.ykey[(c >> 12) & 0xfffu]);
        ia[i] = (uint32_t)i;
    }
    __syncthreads();
    // ---- 2. LSD radix sort by path
    for (int shift = 0; shift < L.key_bits; shift += 8) {
        radix_pass(ka, ia, kb, ib, n, shift, S);
        uint32_t *t = ka; ka = kb; kb = t;
        t = ia; ia = ib; ib = t;
    }
    uint32_t *seg = kb;                                  // free ping-pong buffers become scratch
    uint8_t *head = reinterpret_cast<uint8_t *>(ib);

    // ---- 3. split depths + histograms
    if (tid < 40) { S.hist_sd[tid] = 0; S.hist_g[tid] = 0; }
    __syncthreads();
    for (int base = 0; base < n; base += OCT_THREADS) {
        const int i = base + tid;
        unsigned s = 0x100u + lane;
        if (i < n - 1) {
            const int hb = 31 - __clz(ka[i] ^ ka[i + 1]);
            s = hb >= 2 * D ? 0u : (unsigned)(D - (hb >> 1));
            sd[i] = (uint8_t)s;
        }
        const unsigned peers = __match_any_sync(FULL, s);
        if (i < n - 1 && (peers & ((1u << lane) - 1u)) == 0) atomicAdd(&S.hist_sd[s], __popc(peers));
    }
    __syncthreads();
    for (int base = 0; base < n; base += OCT_THREADS) {
        const int i = base + tid;
        unsigned g = 0x100u + lane;
        if (i < n) {
            const unsigned l = i > 0 ? sd[i - 1] : 0u, r = i < n - 1 ? sd[i] : 0u;
            g = max(l, r);
        }
        const unsigned peers = __match_any_sync(FULL, g);
        if (i < n && (peers & ((1u << lane) - 1u)) == 0) atomicAdd(&S.hist_g[g], __popc(peers));
    }
    __syncthreads();
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
    for (int i = tid; i < n; i += OCT_THREADS) head[i] = (i == 0 || sd[i - 1] <= k0) ? 1 : 0;
    __syncthreads();

    if (mode == 1) {
        // ---- 5. careful phase.  nodes at depth k0 -> node_start[]
        const int nseg = S.ctl[C_SIZE];  // < N <= pcap
        uint32_t carry = 0;
        for (int base = 0; base < n; base += OCT_THREADS) {
            const int i = base + tid;
            const uint32_t h = i < n ? head[i] : 0u;
            uint32_t tot;
            const uint32_t ex = block_excl_scan(h, S.warp_tot, &tot) + carry;
            if (h) node_start[ex] = (uint32_t)i;
            carry += tot;
        }
        if (tid == 0) node_start[nseg] = (uint32_t)n;
        __syncthreads();
        // expandable nodes with their closed-form creation sequence
        const uint32_t digit_mask = 0xCCCCCCCCu & ((k0 >= 16) ? 0xffffffffu : ((1u << (2 * k0)) - 1u));
        const uint32_t root_mask = (k0 & 1) ? 0u : (((1u << (L.key_bits - 2 * D)) - 1u) << (2 * k0));
        for (int s = tid; s < nseg; s += OCT_THREADS) {
            const uint32_t st = node_start[s], cnt = node_start[s + 1] - st;
            if (cnt > 1) {
                const int p = atomicAdd(&S.ctl[C_PN], 1);
                p_start[p] = st; p_cnt[p] = cnt;
                p_seq[p] = (ka[st] >> (2 * (D - k0))) ^ digit_mask ^ root_mask;
            }
        }
        __syncthreads();
        int d = k0;
        uint32_t *ps = p_start, *pc = p_cnt, *pq = p_seq, *qs = q_start, *qc = q_cnt, *qq = q_seq;
        while (true) {
            const int pn = S.ctl[C_PN], size = S.ctl[C_SIZE];
            int m = 1;
            while (m < pn) m <<= 1;
            // gains + sort keys
            for (int p = tid; p < m; p += OCT_THREADS) {
                if (p < pn) {
                    const uint32_t st = ps[p], cnt = pc[p];
                    uint32_t g = 0;
                    for (uint32_t i = st; i + 1 < st + cnt; ++i) g += (sd[i] == d + 1);
                    p_gain[p] = g;
                    skey[p] = ((unsigned long long)cnt << 32) | pq[p];
                } else {
                    skey[p] = 0ull;
                }
                sval[p] = (uint32_t)p;
            }
            if (tid == 0) { S.ctl[C_CUT] = 0x7fffffff; S.ctl[C_QN] = 0; }
            __syncthreads();
            bitonic_desc(skey, sval, m);
            // prefix sums of gains in processing order; first rank where size reaches N
            uint32_t carry2 = 0;
            for (int base = 0; base < pn; base += OCT_THREADS) {
                const int r = base + tid;
                const uint32_t g = r < pn ? p_gain[sval[r]] : 0u;
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
            // split the first nsplit nodes: new heads, and (if we go on) their expandable children
            for (int r = tid; r < nsplit; r += OCT_THREADS) {
                const int p = (int)sval[r];
                const uint32_t st = ps[p], en = st + pc[p];
                uint32_t cs = st;
                for (uint32_t i = st; i < en; ++i) {
                    const bool last = (i + 1 == en);
                    if (last || sd[i] == d + 1) {
                        if (!last) head[i + 1] = 1;
                        const uint32_t len = i + 1 - cs;
                        if (!found && len > 1) {
                            const int q = atomicAdd(&S.ctl[C_QN], 1);
                            qs[q] = cs; qc[q] = len;
                            qq[q] = (uint32_t)r * 4u + ((ka[cs] >> (2 * (D - d - 1))) & 3u);
                        }
                        cs = i + 1;
                    }
                }
            }
            __syncthreads();
            if (found || total == 0) break;
            if (tid == 0) { S.ctl[C_SIZE] = size + (int)total; S.ctl[C_PN] = S.ctl[C_QN]; }
            uint32_t *t;
            t = ps; ps = qs; qs = t;
            t = pc; pc = qc; qc = t;
            t = pq; pq = qq; qq = t;
            ++d;
            __syncthreads();
        }
    }

    // ---- 6. final nodes = head-delimited segments; keep the best key of each
    uint32_t carry = 0;
    for (int base = 0; base < n; base += OCT_THREADS) {
        const int i = base + tid;
        const uint32_t h = i < n ? head[i] : 0u;
        uint32_t tot;
        const uint32_t inc = block_excl_scan(h, S.warp_tot, &tot) + carry + h;
        if (i < n) seg[i] = inc - 1;
        carry += tot;
    }
    const int nfinal = min((int)carry, L.sel_cap);
    for (int s = tid; s < nfinal; s += OCT_THREADS) best[s] = 0ull;
    __syncthreads();
    for (int i = tid; i < n; i += OCT_THREADS) {
        const uint32_t s = seg[i];
        if (s < (uint32_t)nfinal) {
            const uint32_t id = ia[i], c = cand[id];
            const uint32_t xo = __ldg(&L.xord[c & 0xfffu]), yo = __ldg(&L.yord[(c >> 12) & 0xfffu]);
            const uint32_t ord = ((yo >> 6) << 19) | ((xo >> 6) << 12) | ((yo & 63u) << 6) | (xo & 63u);
            const unsigned long long v = ((unsigned long long)(c >> 24) << 48) |
                                         ((unsigned long long)(0x3ffffffu - ord) << 22) | id;
            atomicMax(&best[s], v);
        }
    }
    __syncthreads();
    uint32_t *sel = L.sel + (size_t)frame * L.sel_cap;
    for (int s = tid; s < nfinal; s += OCT_THREADS) sel[s] = cand[(uint32_t)(best[s] & 0x3fffffull)];
    if (tid == 0) *out_count = nfinal;
}

size_t octree_dyn_smem(int sel_cap_max, int pcap, int pcap2) {
    return (size_t)sel_cap_max * 8 + (size_t)pcap2 * 12 + (size_t)pcap * 4 * 7 + (size_t)(pcap + 1) * 4 + 16;
}

cudaError_t launch_octree(const LevelDev *d_levels, int n_levels, const int *d_cand_count, int *d_sel_count,
                          int level_base, int n_launch_levels, int frame_base, int n_frames, int quota_override,
                          int sel_cap_max, int pcap, int pcap2, cudaStream_t st) {
    const size_t smem = octree_dyn_smem(sel_cap_max, pcap, pcap2);
    static size_t configured = 0;
    if (smem > 32 * 1024 && smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(k_octree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    dim3 grid(n_launch_levels, n_frames);
    k_octree<<<grid, OCT_THREADS, smem, st>>>(d_levels, n_levels, d_cand_count, d_sel_count, level_base, frame_base,
                                              quota_override, pcap, pcap2);
    return cudaGetLastError();
}

}  // namespace orbb
