// Shared device/host helpers for libclimatrix_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <type_traits>

#include "../../include/climatrix_b200.h"

namespace cmx {

constexpr unsigned kFull = 0xffffffffu;

// ---- host-side bookkeeping (thread-local) ---------------------------------
void note_launch();                 // ++launch counter

// Output redirection of the batched IDW kernels (cmx_options.peer_out): the same row segment is stored into the
// buffers of up to CMX_MAX_PEERS devices (this one included) at column offset m_off of rows m_total wide.
struct PeerOut {
    void* base[CMX_MAX_PEERS];
    int n;  // 0: plain output
    long long m_total, m_off;
};
// validated copy of the caller's options (NULL -> defaults); returns CMX_OK or CMX_ERR_BAD_ARG
int read_options(const cmx_options* opt, int* search, PeerOut* peers, int* params_on_device = nullptr);
int cuda_fail(cudaError_t e);       // records text, returns CMX_ERR_CUDA
int check_launch();                 // cudaGetLastError() -> CMX_OK / CMX_ERR_CUDA
int sm_count();                     // SMs of the current device (cached)
void timing_begin(int kind, cudaStream_t s);  // no-ops unless cmx_timing_enable(1)
void timing_end(int kind, cudaStream_t s);

#define CMX_CUDA(call)                                   \
    do {                                                 \
        cudaError_t _e = (call);                         \
        if (_e != cudaSuccess) return ::cmx::cuda_fail(_e); \
    } while (0)

// ---- neighbour-search core, shared by cmx_knn / cmx_idw_* / cmx_ok_mw_* ----
struct KnnJob {
    // samples: coordinate a pairs with the targets' row axis, b with the column axis
    const double* s_a;
    const double* s_b;
    long long n;
    // targets
    const double* t_a;
    const double* t_b;
    long long n_a, n_b;
    int grid;  // 1: dense grid (n_a rows x n_b cols), 0: explicit points (n_a == n_b == M)
    int k;
    // raw neighbour table (any may be null)
    int* idx_out;      // [M][k]
    double* dist_out;  // [M][k]  sqrt(d2)
    double* d2_out;    // [M][k]  exact squared distance
    // fused IDW epilogue (idw.py:163-180); mode 0 = off
    int idw_mode;      // 1: single slice -> out[M] (weights + weighted mean in the kernel's epilogue)
                       // 2: short batch, fused (binned search only): vals [N][n_batch] sample-major -> peers / out [n_batch][m_total]
    long long n_batch;          // mode 2
    const int* nonfinite_flag;  // mode 2: device flag, != 0 when some value is NaN / inf (null: test every value)
    PeerOut peers;              // mode 2: where the [n_batch][m_total] result goes (n >= 1)
    int value_is_f32;  // dtype of vals/out
    const void* vals;  // [N] (mode 1)
    void* out;         // [M] (mode 1)
    int k_min;
    double power;
    int search;        // CMX_SEARCH_AUTO / _BRUTE / _BINNED (cmx_options.search)
    // scratch
    void* workspace;
    size_t workspace_bytes;
};
size_t knn_workspace_bytes(int k);
// with room for the partial-list tables of the sample split the planner would choose for these sizes
size_t knn_workspace_bytes_for(int k, long long n, long long n_a, long long n_b, int grid);
int run_knn(const KnnJob& job, cudaStream_t stream);
void knn_plan(long long n, long long n_a, long long n_b, int grid, int* tile_rows, int* sample_segments);

// ---- device helpers -----------------------------------------------------------
// compile-time loop: f(std::integral_constant<int, I>{}) for I in [Begin, End)
template <int Begin, int End, typename F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (Begin < End) {
        f(std::integral_constant<int, Begin>{});
        static_for<Begin + 1, End>(f);
    }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on `bar` (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(kFull, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(kFull, v, o));
    return v;
}

// IDW weight of one neighbour before normalisation: 1 / max(d, 1e-10)^power  (idw.py:163-164)
__device__ __forceinline__ double idw_weight_from_dist(double dist, double power) {
    const double d = fmax(dist, 1e-10);
    if (power == 2.0) return 1.0 / (d * d);  // numpy evaluates x**2.0 as x*x
    return 1.0 / pow(d, power);
}
__device__ __forceinline__ double idw_raw_weight(double d2, double power) {
    return idw_weight_from_dist(sqrt(d2), power);
}

__device__ __forceinline__ void atomic_min_double(double* addr, double v) {
    unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
    unsigned long long old = *a;
    while (__longlong_as_double((long long)old) > v) {
        const unsigned long long prev = atomicCAS(a, old, (unsigned long long)__double_as_longlong(v));
        if (prev == old) break;
        old = prev;
    }
}
__device__ __forceinline__ void atomic_max_double(double* addr, double v) {
    unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
    unsigned long long old = *a;
    while (__longlong_as_double((long long)old) < v) {
        const unsigned long long prev = atomicCAS(a, old, (unsigned long long)__double_as_longlong(v));
        if (prev == old) break;
        old = prev;
    }
}

__device__ __forceinline__ double deg2rad(double v) { return v * 3.141592653589793 / 180.0; }

// pykrige core.great_circle_distance (Vincenty atan2 form), degrees in, degrees out
__device__ __forceinline__ double great_circle_deg(double lon1, double lat1, double lon2, double lat2) {
    const double la1 = deg2rad(lat1), la2 = deg2rad(lat2);
    const double dlon = deg2rad(lon1 - lon2);
    double s1, c1, s2, c2, sd, cd;
    sincos(la1, &s1, &c1);
    sincos(la2, &s2, &c2);
    sincos(dlon, &sd, &cd);
    const double a = c2 * sd;
    const double b = c1 * s2 - s1 * c2 * cd;
    return 180.0 / 3.141592653589793 * atan2(sqrt(a * a + b * b), s1 * s2 + c1 * c2 * cd);
}

// pykrige core.euclid3_to_great_circle: chord length on the unit sphere -> great-circle degrees
__device__ __forceinline__ double chord_to_great_circle_deg(double chord) {
    return 180.0 - 360.0 / 3.141592653589793 * acos(0.5 * chord);
}

template <typename V>
__device__ __forceinline__ V quiet_nan();
template <>
__device__ __forceinline__ float quiet_nan<float>() { return __int_as_float(0x7fc00000); }
template <>
__device__ __forceinline__ double quiet_nan<double>() { return __longlong_as_double(0x7ff8000000000000LL); }

}  // namespace cmx
