// Issue rate of the legacy warp-level int8 tensor-core instruction (mma.sync m16n8k32 s8 -> IMMA.16832.S8.S8) on this
// GPU: every warp runs 8 independent accumulator chains with operands in registers, no memory traffic.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/imma_probe.cu -o tools/_build/imma_probe
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(1024, 1) k_imma(int iters, int seed, long long *cycles, int *sink) {
    unsigned a[4], b[2];
    for (int i = 0; i < 4; ++i) a[i] = 0x01ff01ffu * (seed + threadIdx.x + i);
    for (int i = 0; i < 2; ++i) b[i] = 0xff01ff01u * (seed + threadIdx.x * 3 + i);
    int c[8][4] = {};
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+r"(c[j][0]), "+r"(c[j][1]), "+r"(c[j][2]), "+r"(c[j][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    __syncthreads();
    const long long t1 = clock64();
    int s = 0;
    for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    if (s == 0x7fffffff) sink[0] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
    long long *d_cycles; int *d_sink;
    cudaMalloc(&d_cycles, 148 * sizeof(long long)); cudaMalloc(&d_sink, 4);
    for (int threads : {128, 256, 512, 1024}) {
        const int iters = 4096;
        k_imma<<<148, threads>>>(16, 1, d_cycles, d_sink);
        cudaDeviceSynchronize();
        k_imma<<<148, threads>>>(iters, 1, d_cycles, d_sink);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed\n"); return 1; }
        long long h[148]; cudaMemcpy(h, d_cycles, sizeof h, cudaMemcpyDeviceToHost);
        double cyc = 0; for (int i = 0; i < 148; ++i) cyc += h[i]; cyc /= 148;
        const double mmas = (double)iters * 8 * (threads / 32);
        printf("threads/SM %4d: %.2f cycles per IMMA.16832 per SM (%.3f IMMA/clk/SM, %.0f int8 MAC/clk/SM = %.2f POPS dense at 1.965 GHz x 148 SMs)\n",
               threads, cyc / mmas, mmas / cyc, mmas / cyc * 4096, mmas / cyc * 4096 * 2 * 1.965e9 * 148 / 1e15);
    }
    return 0;
}
