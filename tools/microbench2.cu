// Instruction-mix candidates for the kNN filter loop (no memory traffic): pairs/s per variant.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 2048;

__device__ __forceinline__ unsigned long long fma2s(float a, unsigned long long b, unsigned long long c) {
    // scalar a broadcast to both halves
    unsigned long long d, aa;
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(aa), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ float min3(float a, float b, float c) {
    float d; asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d;
}

// A: scalar FFMA, FMNMX3 over sample pairs, one compare per target per group of G samples
template <int QPT, int G>
__global__ void k_A(float* out, float x0, float y0, float s0, float thr0) {
    float qa[QPT], qb[QPT], thr[QPT];
#pragma unroll
    for (int q = 0; q < QPT; ++q) { qa[q] = threadIdx.x * 0.01f + q; qb[q] = q * 0.5f - threadIdx.x * 0.02f; thr[q] = thr0 - q; }
    int hits = 0;
    float xs = x0, ys = y0, ss = s0;
    for (int it = 0; it < ITERS; ++it) {
        float m[QPT];
#pragma unroll
        for (int q = 0; q < QPT; ++q) m[q] = 3e38f;
#pragma unroll
        for (int s = 0; s < G; s += 2) {
            float xa = xs + s, xb = xs + s + 1.f, ya = ys - s, yb = ys - s - 1.f, sa = ss + s, sb = ss + 2.f * s;
#pragma unroll
            for (int q = 0; q < QPT; ++q) {
                float t0 = fmaf(qa[q], xa, fmaf(qb[q], ya, sa));
                float t1 = fmaf(qa[q], xb, fmaf(qb[q], yb, sb));
                m[q] = min3(m[q], t0, t1);
            }
        }
        bool any = false;
#pragma unroll
        for (int q = 0; q < QPT; ++q) any |= m[q] < thr[q];
        if (any) hits++;
        xs += 0.25f; ys -= 0.125f; ss += 1.0f;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = hits;
}

// B: FFMA2 over sample pairs (scalar target operand), FMNMX3 on the two halves
template <int QPT, int G>
__global__ void k_B(float* out, float x0, float y0, float s0, float thr0) {
    float qa[QPT], qb[QPT], thr[QPT];
#pragma unroll
    for (int q = 0; q < QPT; ++q) { qa[q] = threadIdx.x * 0.01f + q; qb[q] = q * 0.5f - threadIdx.x * 0.02f; thr[q] = thr0 - q; }
    int hits = 0;
    float2 xs = make_float2(x0, x0 + 1), ys = make_float2(y0, y0 - 1), ss = make_float2(s0, s0 + 2);
    for (int it = 0; it < ITERS; ++it) {
        float m[QPT];
#pragma unroll
        for (int q = 0; q < QPT; ++q) m[q] = 3e38f;
#pragma unroll
        for (int s = 0; s < G; s += 2) {
            float2 xv = make_float2(xs.x + s, xs.y + s), yv = make_float2(ys.x - s, ys.y - s), sv = make_float2(ss.x + s, ss.y + s);
            unsigned long long X = *reinterpret_cast<unsigned long long*>(&xv), Y = *reinterpret_cast<unsigned long long*>(&yv), S = *reinterpret_cast<unsigned long long*>(&sv);
#pragma unroll
            for (int q = 0; q < QPT; ++q) {
                unsigned long long r = fma2s(qa[q], X, fma2s(qb[q], Y, S));
                float2 t = *reinterpret_cast<float2*>(&r);
                m[q] = min3(m[q], t.x, t.y);
            }
        }
        bool any = false;
#pragma unroll
        for (int q = 0; q < QPT; ++q) any |= m[q] < thr[q];
        if (any) hits++;
        xs.x += 0.25f; xs.y += 0.25f; ys.x -= 0.125f; ys.y -= 0.125f; ss.x += 1.0f; ss.y += 1.0f;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = hits;
}

// C: FFMA2 + per-pair FSETP chain (reference point)
template <int QPT, int G>
__global__ void k_C(float* out, float x0, float y0, float s0, float thr0) {
    float qa[QPT], qb[QPT], thr[QPT];
#pragma unroll
    for (int q = 0; q < QPT; ++q) { qa[q] = threadIdx.x * 0.01f + q; qb[q] = q * 0.5f - threadIdx.x * 0.02f; thr[q] = thr0 - q; }
    int hits = 0;
    float2 xs = make_float2(x0, x0 + 1), ys = make_float2(y0, y0 - 1), ss = make_float2(s0, s0 + 2);
    for (int it = 0; it < ITERS; ++it) {
        bool any = false;
#pragma unroll
        for (int s = 0; s < G; s += 2) {
            float2 xv = make_float2(xs.x + s, xs.y + s), yv = make_float2(ys.x - s, ys.y - s), sv = make_float2(ss.x + s, ss.y + s);
            unsigned long long X = *reinterpret_cast<unsigned long long*>(&xv), Y = *reinterpret_cast<unsigned long long*>(&yv), S = *reinterpret_cast<unsigned long long*>(&sv);
#pragma unroll
            for (int q = 0; q < QPT; ++q) {
                unsigned long long r = fma2s(qa[q], X, fma2s(qb[q], Y, S));
                float2 t = *reinterpret_cast<float2*>(&r);
                any |= (t.x < thr[q]) | (t.y < thr[q]);
            }
        }
        if (any) hits++;
        xs.x += 0.25f; xs.y += 0.25f; ys.x -= 0.125f; ys.y -= 0.125f; ss.x += 1.0f; ss.y += 1.0f;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = hits;
}

template <typename F>
float time_ms(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < 5; ++i) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / 5;
}

template <int QPT, int G>
void run(int sms, float* out) {
    for (int wps : {4, 8, 16, 32}) {
        int threads = 128, blocks = sms * wps / 4;
        double pairs = (double)blocks * threads * ITERS * G * QPT;
        float a = time_ms([&] { k_A<QPT, G><<<blocks, threads>>>(out, 0.1f, 0.2f, 5.f, -1e30f); });
        float b = time_ms([&] { k_B<QPT, G><<<blocks, threads>>>(out, 0.1f, 0.2f, 5.f, -1e30f); });
        float c = time_ms([&] { k_C<QPT, G><<<blocks, threads>>>(out, 0.1f, 0.2f, 5.f, -1e30f); });
        printf("QPT=%d G=%2d warps/SM=%2d : A(FFMA+MNMX3) %6.2f  B(FFMA2+MNMX3) %6.2f  C(FFMA2+FSETP) %6.2f  algorithmic TFLOP/s\n",
               QPT, G, wps, pairs * 5 / a / 1e9, pairs * 5 / b / 1e9, pairs * 5 / c / 1e9);
    }
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    float* out; cudaMalloc(&out, (size_t)sms * 32 * 32 * 4 * 4);
    run<8, 16>(sms, out);
    run<4, 16>(sms, out);
    run<8, 8>(sms, out);
    run<8, 32>(sms, out);
    run<16, 8>(sms, out);
    return 0;
}
