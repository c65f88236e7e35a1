// FP64 pipe ceilings on sm_100a: DFMA throughput (independent chains) and dependent-issue latency.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_fp64 tools/microbench_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void dfma_tput(double* out, double a, double b, int iters) {
    double acc[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void dfma_latency(double* out, long long* cyc, double a, double b, int iters) {
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x = fma(x, a, b);
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 8);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096;
    for (int warps = 1; warps <= 16; warps *= 2) {
        dfma_tput<8><<<p.multiProcessorCount * 2, warps * 32 / 2 < 32 ? 32 : warps * 32 / 2>>>(out, 1.0000001, 1e-9, 16);
        cudaDeviceSynchronize();
        const int threads = warps * 32 / 2 < 32 ? 32 : warps * 32 / 2;
        cudaEventRecord(e0);
        dfma_tput<8><<<p.multiProcessorCount * 2, threads>>>(out, 1.0000001, 1e-9, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 8 * iters * (double)threads * p.multiProcessorCount * 2;
        printf("DFMA throughput, %2d warps/SM x 8 chains: %.2f TFLOP/s (%.3f ms)\n", threads * 2 / 32, fl / ms / 1e9, ms);
    }
    dfma_latency<<<1, 32>>>(out, cyc, 1.0000001, 1e-9, 1024);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dependent DFMA latency: %.2f cycles\n", (double)h / (1024 * 16));
    return 0;
}
