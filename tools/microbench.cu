// FP32-pipe microbenchmarks on the current GPU: what is the real ceiling for the kNN inner loop?
//   ffma      : independent 3-register FFMA chains
//   ffma2     : packed fma.rn.f32x2 (Blackwell)
//   knn_mix   : 2 FFMA + 1 FSETP per "pair" (the filter loop body without loads)
//   knn_mix2  : 2 FFMA2 + 2 FSETP per two pairs
// Prints achieved TFLOP/s (2 flop per scalar FMA) and "pairs/s" for the mixes.
#include <cstdio>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 4096;
constexpr int CH = 16;

__global__ void k_ffma(float* out, float a, float b) {
    float acc[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) acc[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

__global__ void k_ffma2(float* out, float a, float b) {
    unsigned long long acc[CH];
    unsigned long long A, B;
    float2 fa = make_float2(a, a), fb = make_float2(b, b);
    A = *reinterpret_cast<unsigned long long*>(&fa);
    B = *reinterpret_cast<unsigned long long*>(&fb);
#pragma unroll
    for (int i = 0; i < CH; ++i) { float2 v = make_float2(threadIdx.x * 0.001f + i, i); acc[i] = *reinterpret_cast<unsigned long long*>(&v); }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) acc[i] = fma2(acc[i], A, B);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) { float2 v = *reinterpret_cast<float2*>(&acc[i]); s += v.x + v.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 8 targets per thread, samples come from registers rotated each iteration (no memory)
__global__ void k_mix(float* out, float x0, float y0, float s0, float thr0) {
    float qa[8], qb[8], thr[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) { qa[q] = threadIdx.x * 0.01f + q; qb[q] = q * 0.5f - threadIdx.x * 0.02f; thr[q] = thr0 - q; }
    int hits = 0;
    float xs = x0, ys = y0, ss = s0;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            bool any = false;
            float t[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) { t[q] = fmaf(qa[q], xs, fmaf(qb[q], ys, ss)); any |= t[q] < thr[q]; }
            if (any) {
#pragma unroll
                for (int q = 0; q < 8; ++q) hits += t[q] < thr[q];
            }
            xs += 0.25f; ys -= 0.125f; ss += 1.0f;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = hits;
}

__global__ void k_mix2(float* out, float x0, float y0, float s0, float thr0) {
    // two samples per FFMA2: (qa,qa)*(x0,x1)+(s0,s1)
    unsigned long long qa2[8], qb2[8];
    float thr[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        float2 a = make_float2(threadIdx.x * 0.01f + q, threadIdx.x * 0.01f + q);
        float2 b = make_float2(q * 0.5f - threadIdx.x * 0.02f, q * 0.5f - threadIdx.x * 0.02f);
        qa2[q] = *reinterpret_cast<unsigned long long*>(&a);
        qb2[q] = *reinterpret_cast<unsigned long long*>(&b);
        thr[q] = thr0 - q;
    }
    int hits = 0;
    float2 xs = make_float2(x0, x0 + 0.25f), ys = make_float2(y0, y0 - 0.125f), ss = make_float2(s0, s0 + 1.f);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            unsigned long long X = *reinterpret_cast<unsigned long long*>(&xs), Y = *reinterpret_cast<unsigned long long*>(&ys), S = *reinterpret_cast<unsigned long long*>(&ss);
            bool any = false;
            float2 t[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                unsigned long long r = fma2(qa2[q], X, fma2(qb2[q], Y, S));
                t[q] = *reinterpret_cast<float2*>(&r);
                any |= (t[q].x < thr[q]) | (t[q].y < thr[q]);
            }
            if (any) {
#pragma unroll
                for (int q = 0; q < 8; ++q) hits += (t[q].x < thr[q]) + (t[q].y < thr[q]);
            }
            xs.x += 0.5f; xs.y += 0.5f; ys.x -= 0.25f; ys.y -= 0.25f; ss.x += 2.f; ss.y += 2.f;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = hits;
}

template <typename F>
float time_ms(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < 5; ++i) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / 5;
}

int main() {
    cudaDeviceProp p; CHECK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("%s SMs=%d clock=%d kHz\n", p.name, sms, p.clockRate);
    for (int threads : {128, 256, 512, 1024}) {
        int blocks = sms * (2048 / threads);
        float* out; CHECK(cudaMalloc(&out, (size_t)blocks * threads * 4));
        double n = (double)blocks * threads;
        float ms = time_ms([&] { k_ffma<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
        printf("threads/SM=2048 (block %4d) ffma   : %7.2f TFLOP/s\n", threads, n * ITERS * CH * 2 / ms / 1e9);
        ms = time_ms([&] { k_ffma2<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
        printf("threads/SM=2048 (block %4d) ffma2  : %7.2f TFLOP/s\n", threads, n * ITERS * CH * 4 / ms / 1e9);
        CHECK(cudaFree(out));
    }
    for (int warps_per_sm : {8, 16, 32}) {
        int threads = 128, blocks = sms * warps_per_sm / 4;
        float* out; CHECK(cudaMalloc(&out, (size_t)blocks * threads * 4));
        double n = (double)blocks * threads;
        float ms = time_ms([&] { k_mix<<<blocks, threads>>>(out, 0.1f, 0.2f, 5.f, -1e30f); });
        double pairs = n * ITERS * 4 * 8;
        printf("warps/SM=%2d knn_mix  (2 FFMA + FSETP)/pair : %.3e pairs/s = %6.2f 'algorithmic' TFLOP/s (5/pair)\n", warps_per_sm, pairs / ms * 1e3, pairs * 5 / ms / 1e9);
        ms = time_ms([&] { k_mix2<<<blocks, threads>>>(out, 0.1f, 0.2f, 5.f, -1e30f); });
        pairs = n * ITERS * 2 * 8 * 2;
        printf("warps/SM=%2d knn_mix2 (FFMA2 form)           : %.3e pairs/s = %6.2f 'algorithmic' TFLOP/s (5/pair)\n", warps_per_sm, pairs / ms * 1e3, pairs * 5 / ms / 1e9);
        CHECK(cudaFree(out));
    }
    return 0;
}
