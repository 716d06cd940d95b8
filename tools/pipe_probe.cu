// pipe_probe.cu -- issue-rate microbenchmarks for the instructions the FAST / matcher kernels lean on (B200, sm_100a).
// Each kernel: one CTA of 1024 threads per SM, CHAINS independent register chains per thread, no memory traffic;
// prints lanes per clock per SM.  Pairs of instruction kinds are also run interleaved: if the mixed rate is the SUM of
// the single rates the two kinds issue to different pipes, if it is their harmonic combination they share one.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_probe tools/pipe_probe.cu && ./pipe_probe
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include <algorithm>

#define CHAINS 8
enum Kind { K_VIMNMX3 = 0, K_VIMNMX2, K_HMNMX2, K_IMAD, K_LOP3, K_VABSDIFF4, K_POPC, K_SHFL, K_VOTE, K_IADD3, K_DP4A, K_VIADD16, K_FMNMX,
            K_PRMT, K_SHF, K_ISETP_SEL, K_HMNMX2_BF, K_NKINDS };
static const char *kNames[] = {"VIMNMX3.U16x2", "VIMNMX.U16x2", "HMNMX2(f16x2)", "IMAD", "LOP3", "VABSDIFF4", "POPC", "SHFL.UP", "VOTE.ANY",
                               "IADD3", "IDP.4A", "VIADD.16x2", "FMNMX", "PRMT", "SHF(funnel)", "ISETP+SEL", "HMNMX2(bf16x2)"};

template <int KIND>
__device__ __forceinline__ unsigned op(unsigned a, unsigned b, unsigned c) {
    if (KIND == K_VIMNMX3) return __vimin3_u16x2(a, b, c);
    if (KIND == K_VIMNMX2) return __vminu2(a, b);
    if (KIND == K_HMNMX2) { unsigned r; asm volatile("min.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
    if (KIND == K_HMNMX2_BF) { unsigned r; asm volatile("min.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
    if (KIND == K_IMAD) return a * b + c;
    if (KIND == K_LOP3) { unsigned r; asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
    if (KIND == K_VABSDIFF4) return __vabsdiffu4(a, b);
    if (KIND == K_POPC) { unsigned r; asm volatile("popc.b32 %0, %1;" : "=r"(r) : "r"(a)); return r + b; }  // volatile: the plain chain is folded away
    if (KIND == K_SHFL) return __shfl_up_sync(0xffffffffu, a, 1) + b;
    if (KIND == K_VOTE) return __ballot_sync(0xffffffffu, a & 1) ^ b;
    if (KIND == K_IADD3) { unsigned r; asm volatile("add.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
    if (KIND == K_DP4A) return __dp4a(a, b, c);
    if (KIND == K_VIADD16) return __vadd2(a, b);
    if (KIND == K_FMNMX) return __float_as_uint(fminf(__uint_as_float(a), __uint_as_float(b)));
    if (KIND == K_PRMT) return __byte_perm(a, b, c);
    if (KIND == K_SHF) return __funnelshift_r(a, b, c);
    if (KIND == K_ISETP_SEL) return a > b ? c : a;
    return a;
}

template <int KA, int KB>
__global__ void __launch_bounds__(1024, 1) k_probe(int iters, unsigned seed, long long *cycles, unsigned *sink) {
    unsigned x[CHAINS], y[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { x[i] = 0x64006400u + (seed + threadIdx.x * (2 * i + 1)) % 0x03ff03ffu; y[i] = 0x64016402u + i * 0x00030001u + threadIdx.x; }
    const unsigned p = 0x64ff64f0u + (threadIdx.x & 7), q = seed | 1u;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            x[i] = op<KA>(x[i], p, q);
            if (KB >= 0) y[i] = op<(KB >= 0 ? KB : 0)>(y[i], q, p);
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    unsigned acc = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) acc ^= x[i] ^ y[i];
    if (acc == 0x12345678u) sink[0] = acc;
}

template <int KA, int KB>
static double run(int n_sm, long long *d_cyc, unsigned *d_sink) {
    const int iters = 4000;
    for (int rep = 0; rep < 2; ++rep) k_probe<KA, KB><<<n_sm, 1024>>>(iters, 0x9e3779b9u, d_cyc, d_sink);
    cudaDeviceSynchronize();
    std::vector<long long> c(n_sm);
    cudaMemcpy(c.data(), d_cyc, sizeof(long long) * n_sm, cudaMemcpyDeviceToHost);
    std::sort(c.begin(), c.end());
    const double n_ops = 1024.0 * CHAINS * iters * (KB >= 0 ? 2 : 1);
    return n_ops / (double)c[n_sm / 2];
}

#define SINGLE(K) printf("%-16s %7.2f lanes/clk/SM\n", kNames[K], run<K, -1>(n_sm, d_cyc, d_sink));
#define MIXED(A, B) printf("%-16s + %-16s %7.2f lanes/clk/SM (both kinds counted)\n", kNames[A], kNames[B], run<A, B>(n_sm, d_cyc, d_sink));

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int n_sm = prop.multiProcessorCount;
    long long *d_cyc; unsigned *d_sink;
    cudaMalloc(&d_cyc, sizeof(long long) * n_sm); cudaMalloc(&d_sink, 4);
    printf("# %s, %d SMs; 1024 threads/SM, %d chains/thread; note IADD/POPC/SHFL/VOTE kinds carry one extra add/xor per op\n", prop.name, n_sm, CHAINS);
    SINGLE(K_VIMNMX3) SINGLE(K_VIMNMX2) SINGLE(K_HMNMX2) SINGLE(K_HMNMX2_BF) SINGLE(K_IMAD) SINGLE(K_LOP3) SINGLE(K_VABSDIFF4) SINGLE(K_POPC)
    SINGLE(K_SHFL) SINGLE(K_VOTE) SINGLE(K_IADD3) SINGLE(K_DP4A) SINGLE(K_VIADD16) SINGLE(K_FMNMX) SINGLE(K_PRMT) SINGLE(K_SHF) SINGLE(K_ISETP_SEL)
    MIXED(K_VIMNMX3, K_HMNMX2) MIXED(K_VIMNMX3, K_IMAD) MIXED(K_VIMNMX3, K_LOP3) MIXED(K_HMNMX2, K_IMAD) MIXED(K_HMNMX2, K_LOP3)
    MIXED(K_LOP3, K_IMAD) MIXED(K_LOP3, K_POPC) MIXED(K_VIMNMX3, K_DP4A) MIXED(K_VIMNMX3, K_VIADD16) MIXED(K_VIMNMX3, K_FMNMX) MIXED(K_HMNMX2, K_DP4A)
    MIXED(K_HMNMX2, K_HMNMX2_BF)
    return 0;
}
