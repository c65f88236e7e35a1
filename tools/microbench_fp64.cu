// FP64 pipe ceilings on sm_100a: DFMA throughput against warps per SM and independent chains per warp, the
// dependent-issue latency of DFMA, and the latency of the reciprocal-square-root chain (seed + two Newton steps),
// a warp shuffle and a shared-memory round trip - the serial pieces of one K4 factorisation step.
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
__device__ __forceinline__ double rsqrt_pos(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double h = 0.5 * d;
    y = y * fma(-h * y, y, 1.5);
    y = y * fma(-h * y, y, 1.5);
    return y;
}
// mode 0: DFMA chain, 1: rsqrt chain, 2: shuffle chain, 3: shared-memory store -> load chain
__global__ void latency(double* out, long long* cyc, double a, double b, int iters, int mode) {
    __shared__ double buf[64];
    double x = threadIdx.x + 2.0;
    buf[threadIdx.x] = x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (mode == 0) x = fma(x, a, b);
            else if (mode == 1) x = rsqrt_pos(x) + 2.0;
            else if (mode == 2) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 5) & 31);
            else {
                buf[threadIdx.x] = x;
                __syncwarp();
                x = buf[(threadIdx.x + 1) & 31] + 1.0;
                __syncwarp();
            }
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <int CH>
void tput(double* out, int sms, int warps_per_sm) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096;
    const int threads = warps_per_sm * 32;
    dfma_tput<CH><<<sms, threads>>>(out, 1.0000001, 1e-9, 16);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    dfma_tput<CH><<<sms, threads>>>(out, 1.0000001, 1e-9, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * CH * iters * (double)threads * sms;
    printf("DFMA throughput, %2d warps/SM x %2d chains: %6.2f TFLOP/s\n", warps_per_sm, CH, fl / ms / 1e9);
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 8);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    for (int w = 4; w <= 32; w *= 2) {
        tput<1>(out, p.multiProcessorCount, w);
        tput<2>(out, p.multiProcessorCount, w);
        tput<4>(out, p.multiProcessorCount, w);
        tput<8>(out, p.multiProcessorCount, w);
        tput<16>(out, p.multiProcessorCount, w);
    }
    const char* names[4] = {"dependent DFMA", "rsqrt seed + 2 Newton + DADD", "SHFL.IDX (64-bit)", "STS + syncwarp + LDS + DADD + syncwarp"};
    for (int mode = 0; mode < 4; ++mode) {
        latency<<<1, 32>>>(out, cyc, 1.0000001, 1e-9, 1024, mode);
        cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("latency, %s: %.1f cycles\n", names[mode], (double)h / (1024 * 16));
    }
    return 0;
}
