// Micro-benchmark: peak rate of the byte-wise SAD instruction the motion search is built on
// (vabsdiff4.u32.u32.u32.add = one VABSDIFF4.U8.ACC per four pixel differences), so that K2's achieved
// pixel-absdiffs/s can be put against a MEASURED integer-pipe ceiling (SURVEY 8d asks for this denominator).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/sad_peak tools/sad_peak.cu && gpurun_out/sad_peak
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t sad4(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

template <int ILP>
__global__ void __launch_bounds__(1024) sad_kernel(uint32_t* out, uint32_t seed, int iters) {
    uint32_t acc[ILP], a = seed ^ threadIdx.x, b = seed * 2654435761u + blockIdx.x;
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) acc[i] = sad4(a + i, b + u, acc[i]);
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i];
    if (s == 0xdeadbeefu) out[0] = s;
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    uint32_t* d;
    cudaMalloc(&d, 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096, ILP = 8;
    const int blocks = prop.multiProcessorCount * 2;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        sad_kernel<ILP><<<blocks, 1024>>>(d, 12345u + rep, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double instr = (double)blocks * 1024 * iters * 8 * ILP;
        int clk = 0;
        cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
        printf("{\"sms\": %d, \"ms\": %.3f, \"vabsdiff4_per_s\": %.4e, \"pixel_absdiffs_per_s\": %.4e, \"per_clk_per_sm_at_%dMHz\": %.2f}\n",
               prop.multiProcessorCount, ms, instr / (ms * 1e-3), 4 * instr / (ms * 1e-3), clk / 1000,
               instr / (ms * 1e-3) / prop.multiProcessorCount / (clk * 1e3));
    }
    return 0;
}
