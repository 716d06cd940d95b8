// Issue rate of the legacy warp-level int8 tensor-core instruction (mma.sync m16n8k32 s8 -> IMMA.16832.S8.S8) on this
// GPU: every warp runs 8 independent accumulator chains with operands in registers, no memory traffic.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/imma_probe.cu -o tools/_build/imma_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int CH>
__global__ void __launch_bounds__(1024, 1) k_imma(int iters, int seed, long long *cycles, int *sink) {
    unsigned a[4], b[2];
    for (int i = 0; i < 4; ++i) a[i] = 0x01ff01ffu * (seed + threadIdx.x + i);
    for (int i = 0; i < 2; ++i) b[i] = 0xff01ff01u * (seed + threadIdx.x * 3 + i);
    int c[CH][4] = {};
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < CH; ++j)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+r"(c[j][0]), "+r"(c[j][1]), "+r"(c[j][2]), "+r"(c[j][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    __syncthreads();
    const long long t1 = clock64();
    int s = 0;
    for (int j = 0; j < CH; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    if (s == 0x7fffffff) sink[0] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int CH>
static void run(int threads, long long *d_cycles, int *d_sink) {
    const int iters = 4096;
    k_imma<CH><<<148, threads>>>(16, 1, d_cycles, d_sink);
    cudaDeviceSynchronize();
    k_imma<CH><<<148, threads>>>(iters, 1, d_cycles, d_sink);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed\n"); return; }
    long long h[148]; cudaMemcpy(h, d_cycles, sizeof h, cudaMemcpyDeviceToHost);
    double cyc = 0; for (int i = 0; i < 148; ++i) cyc += h[i]; cyc /= 148;
    const double mmas = (double)iters * CH * (threads / 32);
    printf("threads/SM %4d, %d independent accumulator chains per warp: %.2f cycles per IMMA.16832 per SM (%.3f IMMA/clk/SM, %.0f int8 MAC/clk/SM = %.2f POPS dense at 1.965 GHz x 148 SMs); one warp's chain step every %.1f cycles\n",
           threads, CH, cyc / mmas, mmas / cyc, mmas / cyc * 4096, mmas / cyc * 4096 * 2 * 1.965e9 * 148 / 1e15, cyc / iters);
}

int main() {
    long long *d_cycles; int *d_sink;
    cudaMalloc(&d_cycles, 148 * sizeof(long long)); cudaMalloc(&d_sink, 4);
    for (int threads : {128, 256, 512, 1024}) run<8>(threads, d_cycles, d_sink);
    run<1>(128, d_cycles, d_sink);   // dependent-issue latency of the accumulator chain: one warp per sub-partition, one chain
    run<2>(128, d_cycles, d_sink);
    run<4>(128, d_cycles, d_sink);
    run<2>(512, d_cycles, d_sink);   // the matcher's first shape: four warps per sub-partition, two chains each
    run<4>(256, d_cycles, d_sink);
    return 0;
}
