// K3: empirical variogram over all N(N-1)/2 sample pairs.
//
// Replaces pykrige core._initialize_variogram_model's pdist + binning (reached from
// reference src/climatrix/reconstruct/kriging.py:215-224; algorithm restated in
// SURVEY.md Appendix A.1 step 3):
//   d = pairwise distance, g = 0.5 (z_i - z_j)^2,
//   bin n holds pairs with edges[n] <= d < edges[n+1]; per bin: sum d, sum g, count.
// Two passes (the bin edges depend on min/max d): cmx_variogram_minmax, then
// cmx_variogram_bins.  All arithmetic float64 on the FP64 pipe; the pair matrix is
// never materialised (pykrige's pdist needs 10 GB per array at N = 50 000).
//
// Tiling: CTA (bi, bj), bj <= bi, owns a 256 x 256 block of the lower triangle; each
// thread keeps one row sample in registers and sweeps the column samples from shared
// memory.  Per-thread bin accumulators live in (L1-resident) local memory, are
// reduced over the block in shared memory and leave the CTA as one atomic per bin.
#include "common.cuh"

namespace cmx {
namespace {

constexpr int VT = 256;        // tile edge == threads per CTA
constexpr int MAX_LAGS = 64;

template <int METRIC>
__device__ __forceinline__ double pair_key(double xi, double yi, double xj, double yj) {
    // euclidean: squared distance (sqrt is monotone, taken once at the end); geographic: the distance
    if (METRIC == CMX_METRIC_EUCLIDEAN) {
        const double dx = __dsub_rn(xi, xj), dy = __dsub_rn(yi, yj);
        return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    }
    return great_circle_deg(xi, yi, xj, yj);
}

__device__ __forceinline__ void tile_of_block(long long b, int& bi, int& bj) {
    // b = bi*(bi+1)/2 + bj, bj <= bi
    bi = (int)((sqrt(8.0 * (double)b + 1.0) - 1.0) * 0.5);
    while ((long long)bi * (bi + 1) / 2 > b) --bi;
    while ((long long)(bi + 1) * (bi + 2) / 2 <= b) ++bi;
    bj = (int)(b - (long long)bi * (bi + 1) / 2);
}

__global__ void minmax_init_kernel(double* minmax) {
    minmax[0] = __longlong_as_double(0x7ff0000000000000LL);
    minmax[1] = -__longlong_as_double(0x7ff0000000000000LL);
}

template <int METRIC>
__global__ void __launch_bounds__(VT) variogram_minmax_kernel(const double* __restrict__ x,
                                                              const double* __restrict__ y, long long n,
                                                              long long n_blocks, long long part, long long n_parts,
                                                              double* minmax) {
    __shared__ double sx[VT], sy[VT];
    __shared__ double red[2 * VT / 32];
    const int tid = threadIdx.x;
    double lo = __longlong_as_double(0x7ff0000000000000LL), hi = -lo;
    // this call's share of the lower-triangle blocks: b = part (mod n_parts)
    for (long long b = part + n_parts * blockIdx.x; b < n_blocks; b += n_parts * gridDim.x) {
        int bi, bj;
        tile_of_block(b, bi, bj);
        const long long i = (long long)bi * VT + tid;
        const long long j0 = (long long)bj * VT;
        __syncthreads();
        if (j0 + tid < n) {
            sx[tid] = x[j0 + tid];
            sy[tid] = y[j0 + tid];
        }
        __syncthreads();
        if (i < n) {
            const double xi = x[i], yi = y[i];
            const int jn = (int)((n - j0) < VT ? (n - j0) : VT);
            const int jend = bi == bj ? (tid < jn ? tid : jn) : jn;  // strictly lower triangle: j < i
            for (int j = 0; j < jend; ++j) {
                const double key = pair_key<METRIC>(xi, yi, sx[j], sy[j]);
                lo = fmin(lo, key);
                hi = fmax(hi, key);
            }
        }
    }
    lo = warp_min(lo);
    hi = warp_max(hi);
    if ((tid & 31) == 0) {
        red[(tid >> 5) * 2] = lo;
        red[(tid >> 5) * 2 + 1] = hi;
    }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < VT / 32; ++w) {
            lo = fmin(lo, red[w * 2]);
            hi = fmax(hi, red[w * 2 + 1]);
        }
        if (METRIC == CMX_METRIC_EUCLIDEAN) {
            lo = sqrt(lo);
            hi = sqrt(hi);
        }
        atomic_min_double(&minmax[0], lo);
        atomic_max_double(&minmax[1], hi);
    }
}

template <int METRIC>
__global__ void __launch_bounds__(VT) variogram_bins_kernel(const double* __restrict__ x,
                                                            const double* __restrict__ y,
                                                            const double* __restrict__ z, long long n,
                                                            long long n_blocks, long long part, long long n_parts,
                                                            int nlags, const double* __restrict__ edges, double* sum_d,
                                                            double* sum_g, unsigned long long* count) {
    __shared__ double sx[VT], sy[VT], sz[VT];
    __shared__ double s_edges[MAX_LAGS + 1];
    __shared__ double s_sd[MAX_LAGS], s_sg[MAX_LAGS];
    __shared__ unsigned long long s_cnt[MAX_LAGS];
    const int tid = threadIdx.x;
    if (tid <= nlags) s_edges[tid] = edges[tid];
    if (tid < nlags) {
        s_sd[tid] = 0.0;
        s_sg[tid] = 0.0;
        s_cnt[tid] = 0ull;
    }
    double acc_d[MAX_LAGS], acc_g[MAX_LAGS];
    unsigned acc_c[MAX_LAGS];
    for (int b = 0; b < nlags; ++b) {
        acc_d[b] = 0.0;
        acc_g[b] = 0.0;
        acc_c[b] = 0u;
    }
    __syncthreads();
    const double e0 = s_edges[0];
    const double inv_dd = nlags > 0 && s_edges[1] > e0 ? 1.0 / (s_edges[1] - e0) : 0.0;

    for (long long blk = part + n_parts * blockIdx.x; blk < n_blocks; blk += n_parts * gridDim.x) {
        int bi, bj;
        tile_of_block(blk, bi, bj);
        const long long i = (long long)bi * VT + tid;
        const long long j0 = (long long)bj * VT;
        __syncthreads();
        if (j0 + tid < n) {
            sx[tid] = x[j0 + tid];
            sy[tid] = y[j0 + tid];
            sz[tid] = z[j0 + tid];
        }
        __syncthreads();
        if (i < n) {
            const double xi = x[i], yi = y[i], zi = z[i];
            const int jn = (int)((n - j0) < VT ? (n - j0) : VT);
            const int jend = bi == bj ? (tid < jn ? tid : jn) : jn;
            for (int j = 0; j < jend; ++j) {
                double d = pair_key<METRIC>(xi, yi, sx[j], sy[j]);
                if (METRIC == CMX_METRIC_EUCLIDEAN) d = sqrt(d);
                // bin: edges[b] <= d < edges[b+1]; uniform guess, then exact fix-up against the edges
                int b = (int)((d - e0) * inv_dd);
                b = b < 0 ? 0 : (b > nlags - 1 ? nlags - 1 : b);
                while (b > 0 && d < s_edges[b]) --b;
                while (b < nlags - 1 && d >= s_edges[b + 1]) ++b;
                if (d >= s_edges[b] && d < s_edges[b + 1]) {
                    const double dz = zi - sz[j];
                    acc_d[b] += d;
                    acc_g[b] += 0.5 * (dz * dz);
                    acc_c[b] += 1u;
                }
            }
            // (32-bit per-thread counters: a thread sees at most n_blocks / gridDim.x * 256 < 2^32 pairs for n < 2^31)
        }
    }
    // block reduction per bin, then one atomic per bin and CTA
    for (int b = 0; b < nlags; ++b) {
        const double d = warp_sum(acc_d[b]);
        const double g = warp_sum(acc_g[b]);
        unsigned c = acc_c[b];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(kFull, c, o);
        if ((tid & 31) == 0 && c) {
            atomicAdd(&s_sd[b], d);
            atomicAdd(&s_sg[b], g);
            atomicAdd(&s_cnt[b], (unsigned long long)c);
        }
    }
    __syncthreads();
    if (tid < nlags && s_cnt[tid]) {
        atomicAdd(&sum_d[tid], s_sd[tid]);
        atomicAdd(&sum_g[tid], s_sg[tid]);
        atomicAdd(&count[tid], s_cnt[tid]);
    }
}

}  // namespace
}  // namespace cmx

extern "C" {

int cmx_variogram_minmax(const double* x, const double* y, int64_t n, int metric, int part, int n_parts,
                         double* minmax, void* stream_) {
    using namespace cmx;
    if (!x || !y || !minmax || n < 2 || n_parts < 1 || part < 0 || part >= n_parts) return CMX_ERR_BAD_ARG;
    if (metric != CMX_METRIC_EUCLIDEAN && metric != CMX_METRIC_GEOGRAPHIC) return CMX_ERR_UNSUPPORTED;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const long long nb = (n + VT - 1) / VT;
    const long long n_blocks = nb * (nb + 1) / 2;
    const long long mine = (n_blocks - part + n_parts - 1) / n_parts;  // blocks of this part (may be 0)
    const int grid = (int)(mine < (long long)sm_count() * 8 ? (mine > 0 ? mine : 1) : (long long)sm_count() * 8);
    minmax_init_kernel<<<1, 1, 0, stream>>>(minmax);
    note_launch();
    if (metric == CMX_METRIC_EUCLIDEAN)
        variogram_minmax_kernel<CMX_METRIC_EUCLIDEAN><<<grid, VT, 0, stream>>>(x, y, n, n_blocks, part, n_parts, minmax);
    else
        variogram_minmax_kernel<CMX_METRIC_GEOGRAPHIC><<<grid, VT, 0, stream>>>(x, y, n, n_blocks, part, n_parts, minmax);
    note_launch();
    return check_launch();
}

int cmx_variogram_bins(const double* x, const double* y, const double* z, int64_t n, int metric, int part,
                       int n_parts, int nlags, const double* edges, double* sum_d, double* sum_g,
                       unsigned long long* count, void* stream_) {
    using namespace cmx;
    if (!x || !y || !z || !edges || !sum_d || !sum_g || !count || n < 2 || nlags < 1) return CMX_ERR_BAD_ARG;
    if (n_parts < 1 || part < 0 || part >= n_parts) return CMX_ERR_BAD_ARG;
    if (nlags > MAX_LAGS) return CMX_ERR_UNSUPPORTED;
    if (metric != CMX_METRIC_EUCLIDEAN && metric != CMX_METRIC_GEOGRAPHIC) return CMX_ERR_UNSUPPORTED;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const long long nb = (n + VT - 1) / VT;
    const long long n_blocks = nb * (nb + 1) / 2;
    const long long mine = (n_blocks - part + n_parts - 1) / n_parts;
    if (mine <= 0) return CMX_OK;
    const int grid = (int)(mine < (long long)sm_count() * 8 ? mine : (long long)sm_count() * 8);
    if (metric == CMX_METRIC_EUCLIDEAN)
        variogram_bins_kernel<CMX_METRIC_EUCLIDEAN>
            <<<grid, VT, 0, stream>>>(x, y, z, n, n_blocks, part, n_parts, nlags, edges, sum_d, sum_g, count);
    else
        variogram_bins_kernel<CMX_METRIC_GEOGRAPHIC>
            <<<grid, VT, 0, stream>>>(x, y, z, n, n_blocks, part, n_parts, nlags, edges, sum_d, sum_g, count);
    note_launch();
    return check_launch();
}

}  // extern "C"
