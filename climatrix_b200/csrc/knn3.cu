// K1b: exact 3-D k-nearest-neighbour search, one warp per target.
//
// pykrige's moving window with coordinates_type="geographic" finds neighbours by chord length
// between unit vectors (cKDTree on (cos lon cos lat, sin lon cos lat, sin lat); SURVEY.md
// Appendix A.2 steps 2 and 5), reached from the reference call site
// src/climatrix/reconstruct/kriging.py:236-241.  Sample sets of that mode are small to
// moderate, so this kernel keeps everything exact and simple: the lanes of a warp sweep the
// samples in float64 (non-contracted dx*dx + dy*dy + dz*dz, the order cKDTree accumulates),
// a ballot finds samples closer than the running k-th distance, and the warp merges them into
// its register-resident sorted list with the same bitonic sort / merge network as K1.
// FP64-pipe bound: 8 FP64 op per (target, sample) pair.  No workspace.
#include <limits.h>

#include "common.cuh"
#include "rerank.cuh"

namespace cmx {
namespace {

#define CMX_INF __longlong_as_double(0x7ff0000000000000LL)

template <int E>
__global__ void __launch_bounds__(256) knn3_kernel(const double* __restrict__ sx, const double* __restrict__ sy,
                                                   const double* __restrict__ sz, long long n,
                                                   const double* __restrict__ tx, const double* __restrict__ ty,
                                                   const double* __restrict__ tz, long long m_total, int k,
                                                   int* __restrict__ idx_out, double* __restrict__ d2_out) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long m = warp0; m < m_total; m += nwarps) {
        const double qx = tx[m], qy = ty[m], qz = tz[m];
        double ld[E];
        int li[E];
#pragma unroll
        for (int e = 0; e < E; ++e) {
            ld[e] = CMX_INF;
            li[e] = INT_MAX;
        }
        double kth = CMX_INF;  // k-th smallest so far (uniform across the warp)
        for (long long base = 0; base < n; base += 32) {
            const long long i = base + lane;
            double d2 = CMX_INF;
            if (i < n) {
                const double dx = __dsub_rn(qx, sx[i]), dy = __dsub_rn(qy, sy[i]), dz = __dsub_rn(qz, sz[i]);
                d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                if (!(d2 == d2)) d2 = CMX_INF;
            }
            // ties at the k-th distance are resolved by index inside the merge, so '<=' here
            if (__any_sync(kFull, d2 <= kth && d2 < CMX_INF)) {
                double cd = (d2 <= kth) ? d2 : CMX_INF;
                int ci = (d2 <= kth && d2 < CMX_INF) ? (int)i : INT_MAX;
                if (ci == INT_MAX) cd = CMX_INF;
                warp_sort32(cd, ci, lane);
                merge_sorted32<E>(ld, li, cd, ci, lane);
                double sel = ld[0];
#pragma unroll
                for (int e = 1; e < E; ++e)
                    if (((k - 1) >> 5) == e) sel = ld[e];
                kth = __shfl_sync(kFull, sel, (k - 1) & 31);
            }
        }
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int j = e * 32 + lane;
            if (j < k) {
                idx_out[m * k + j] = li[e];
                d2_out[m * k + j] = ld[e];
            }
        }
    }
}

}  // namespace
}  // namespace cmx

extern "C" int cmx_knn3(const double* s_x, const double* s_y, const double* s_z, int64_t n, const double* t_x,
                        const double* t_y, const double* t_z, int64_t m, int k, int32_t* idx_out, double* d2_out,
                        void* stream_) {
    using namespace cmx;
    if (!s_x || !s_y || !s_z || !t_x || !t_y || !t_z || !idx_out || !d2_out) return CMX_ERR_BAD_ARG;
    if (n < 0 || m < 0 || k < 1) return CMX_ERR_BAD_ARG;
    if (k > CMX_MAX_K) return CMX_ERR_K_TOO_LARGE;
    if ((long long)k > n) return CMX_ERR_K_GT_N;
    if (n > 0x7fffffffLL) return CMX_ERR_TOO_MANY;
    if (m == 0) return CMX_OK;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int threads = 256;
    long long blocks = (m * 32 + threads - 1) / threads;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (k <= 32)
        knn3_kernel<1><<<(unsigned)blocks, threads, 0, stream>>>(s_x, s_y, s_z, n, t_x, t_y, t_z, m, k, idx_out, d2_out);
    else if (k <= 64)
        knn3_kernel<2><<<(unsigned)blocks, threads, 0, stream>>>(s_x, s_y, s_z, n, t_x, t_y, t_z, m, k, idx_out, d2_out);
    else
        knn3_kernel<4><<<(unsigned)blocks, threads, 0, stream>>>(s_x, s_y, s_z, n, t_x, t_y, t_z, m, k, idx_out, d2_out);
    note_launch();
    return check_launch();
}
