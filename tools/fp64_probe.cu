// fp64_probe.cu — measures the FP64 FMA peak and the dependent-issue latency of DFMA on the running GPU
// (MEASURED_PEAKS.json has no FP64 figure; DESIGN.md quotes this as "of measured (own)").
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_probe tools/fp64_probe.cu && ./fp64_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_fma(double* out, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
double run(int blocks, int threads, int iters, double* d) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_fma<ILP><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k_fma<ILP><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double* d; cudaMalloc(&d, sizeof(double) * 148 * 1024 * 8);
    const int sms = p.multiProcessorCount;
    // throughput: 8 CTAs x 256 threads per SM, ILP 8
    {
        const int iters = 20000;
        double ms = run<8>(sms * 8, 256, iters, d);
        double flops = 2.0 * 8 * iters * (double)sms * 8 * 256;
        printf("peak: %d SMs, %.1f us, %.2f TFLOP/s FP64 FMA (%.1f FMA/clk/SM at %.0f MHz nominal)\n", sms, ms * 1e3, flops / ms / 1e9,
               flops / 2 / (ms * 1e-3) / sms / (clk_khz * 1e3), clk_khz / 1e3);
    }
    // latency: one warp, one chain
    {
        const int iters = 200000;
        double ms = run<1>(1, 32, iters, d);
        printf("dependent DFMA chain, 1 warp: %.2f ns per DFMA = %.1f cycles at %.0f MHz\n", ms * 1e6 / iters, ms * 1e-3 / iters * clk_khz * 1e3, clk_khz / 1e3);
        for (int w : {1, 2, 4, 8}) {
            double m1 = run<1>(sms, 128 * w, iters / 10, d);
            double m3 = run<3>(sms, 128 * w, iters / 10, d);
            printf("  %d warps/SMSP: ILP1 %.1f cycles/DFMA-issue/warp, ILP3 %.1f cycles per 3 DFMA (pipe-bound = %d)\n", w,
                   m1 * 1e-3 / (iters / 10) * clk_khz * 1e3, m3 * 1e-3 / (iters / 10) * clk_khz * 1e3, 2 * 3 * w);
        }
    }
    return 0;
}
