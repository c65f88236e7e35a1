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
#include "variogram_fit.cuh"

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
    __shared__ double2 sxy[VT];
    __shared__ double red[2 * VT / 32];
    const int tid = threadIdx.x;
    double lo = __longlong_as_double(0x7ff0000000000000LL), hi = -lo;
    double lo2 = lo, hi2 = hi;  // second pair of accumulators: two independent min / max chains
    // this call's share of the lower-triangle blocks: b = part (mod n_parts)
    for (long long b = part + n_parts * blockIdx.x; b < n_blocks; b += n_parts * gridDim.x) {
        int bi, bj;
        tile_of_block(b, bi, bj);
        const long long i = (long long)bi * VT + tid;
        const long long j0 = (long long)bj * VT;
        __syncthreads();
        if (j0 + tid < n) sxy[tid] = make_double2(x[j0 + tid], y[j0 + tid]);
        __syncthreads();
        if (i < n) {
            const double xi = x[i], yi = y[i];
            const int jn = (int)((n - j0) < VT ? (n - j0) : VT);
            const int jend = bi == bj ? (tid < jn ? tid : jn) : jn;  // strictly lower triangle: j < i
            int j = 0;
#pragma unroll 4
            for (; j + 1 < jend; j += 2) {
                const double2 p0 = sxy[j], p1 = sxy[j + 1];
                const double k0 = pair_key<METRIC>(xi, yi, p0.x, p0.y), k1 = pair_key<METRIC>(xi, yi, p1.x, p1.y);
                lo = fmin(lo, k0);
                hi = fmax(hi, k0);
                lo2 = fmin(lo2, k1);
                hi2 = fmax(hi2, k1);
            }
            if (j < jend) {
                const double2 p0 = sxy[j];
                const double k0 = pair_key<METRIC>(xi, yi, p0.x, p0.y);
                lo = fmin(lo, k0);
                hi = fmax(hi, k0);
            }
        }
    }
    lo = fmin(lo, lo2);
    hi = fmax(hi, hi2);
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

// Bin index of distance d: the bins are uniform except the last one (edge dmax + 0.001), so a multiply gives a guess that
// is off by at most one; one comparison against the two edges of the guessed bin corrects it.  Every pair lies in
// exactly one bin (edges[0] = dmin, edges[nlags] > dmax); `in` only filters NaN distances.
__device__ __forceinline__ int variogram_bin(double d, double e0, double e_last, double inv_dd, int nlags,
                                             const double* __restrict__ s_edges, bool& in) {
    int b = (int)((d - e0) * inv_dd);
    b = b < 0 ? 0 : (b > nlags - 1 ? nlags - 1 : b);
    const double lo = s_edges[b], hi = s_edges[b + 1];
    b += (d >= hi) - (d < lo);
    b = b < 0 ? 0 : (b > nlags - 1 ? nlags - 1 : b);
    in = d >= e0 && d < e_last;
    return b;
}

// nlags <= SM_LAGS: the per-thread accumulators are columns of SHARED memory, acc[bin][thread] (conflict-free: the bin only
// moves a lane by multiples of 256 words), updated by load - add - store at the pair's own bin.  The register version below
// pays 9 instructions per BIN and pair (add + select for both sums, compare, count); this one pays 9 per PAIR, and the
// accumulation order of a thread - hence every sum, bit for bit - is the same.
constexpr int SM_LAGS = 32;

template <int METRIC>
__global__ void __launch_bounds__(VT) variogram_bins_smem_kernel(const double* __restrict__ x,
                                                                 const double* __restrict__ y,
                                                                 const double* __restrict__ z, long long n,
                                                                 long long n_blocks, long long part, long long n_parts,
                                                                 int nlags, const double* __restrict__ edges, double* sum_d,
                                                                 double* sum_g, unsigned long long* count,
                                                                 double* __restrict__ part_d, double* __restrict__ part_g,
                                                                 unsigned long long* __restrict__ part_c) {
    extern __shared__ double s_acc[];  // [nlags][VT] sum d | [nlags][VT] sum g | [nlags][VT] count (unsigned)
    __shared__ double2 sxy[VT];
    __shared__ double sz[VT];
    __shared__ double s_edges[SM_LAGS + 2];
    __shared__ double s_wd[VT / 32][SM_LAGS], s_wg[VT / 32][SM_LAGS];
    __shared__ unsigned s_wc[VT / 32][SM_LAGS];
    const int tid = threadIdx.x;
    double* my_d = s_acc + tid;
    double* my_g = s_acc + (size_t)nlags * VT + tid;
    unsigned* my_c = reinterpret_cast<unsigned*>(s_acc + 2 * (size_t)nlags * VT) + tid;
    if (tid <= nlags) s_edges[tid] = edges[tid];
    for (int b = 0; b < nlags; ++b) {
        my_d[b * VT] = 0.0;
        my_g[b * VT] = 0.0;
        my_c[b * VT] = 0u;
    }
    __syncthreads();
    const double e0 = s_edges[0], e_last = s_edges[nlags];
    const double inv_dd = nlags > 0 && s_edges[1] > e0 ? 1.0 / (s_edges[1] - e0) : 0.0;
    const int b_max = nlags - 1;

    for (long long blk = part + n_parts * blockIdx.x; blk < n_blocks; blk += n_parts * gridDim.x) {
        int bi, bj;
        tile_of_block(blk, bi, bj);
        const long long i = (long long)bi * VT + tid;
        const long long j0 = (long long)bj * VT;
        __syncthreads();
        if (j0 + tid < n) {
            sxy[tid] = make_double2(x[j0 + tid], y[j0 + tid]);
            sz[tid] = z[j0 + tid];
        }
        __syncthreads();
        if (i < n) {
            const double xi = x[i], yi = y[i], zi = z[i];
            const int jn = (int)((n - j0) < VT ? (n - j0) : VT);
            const int jend = bi == bj ? (tid < jn ? tid : jn) : jn;
#pragma unroll 2
            for (int j = 0; j < jend; ++j) {
                const double2 pj = sxy[j];
                double d = pair_key<METRIC>(xi, yi, pj.x, pj.y);
                if (METRIC == CMX_METRIC_EUCLIDEAN) d = sqrt(d);
                // uniform bins except the last edge: the multiply is off by at most one, the two edges of the guess decide;
                // a pair outside [e0, e_last) (only a NaN distance) is dropped, so the corrected index needs no clamp
                int b = (int)((d - e0) * inv_dd);
                b = max(0, min(b, b_max));
                const double lo = s_edges[b], hi = s_edges[b + 1];
                b += (d >= hi) - (d < lo);
                const double dz = zi - sz[j];
                const double g = 0.5 * (dz * dz);
                if (d >= e0 && d < e_last) {
                    my_d[b * VT] += d;
                    my_g[b * VT] += g;
                    my_c[b * VT] += 1u;
                }
            }
        }
    }
    // the same deterministic reduction as the register version: lanes, warps in order, CTAs in order (reduce kernel)
    for (int b = 0; b < nlags; ++b) {
        const double d = warp_sum(my_d[b * VT]);
        const double g = warp_sum(my_g[b * VT]);
        unsigned c = my_c[b * VT];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(kFull, c, o);
        if ((tid & 31) == 0) {
            s_wd[tid >> 5][b] = d;
            s_wg[tid >> 5][b] = g;
            s_wc[tid >> 5][b] = c;
        }
    }
    __syncthreads();
    if (tid < nlags) {
        double d = 0.0, g = 0.0;
        unsigned long long c = 0ull;
        for (int w = 0; w < VT / 32; ++w) {
            d += s_wd[w][tid];
            g += s_wg[w][tid];
            c += s_wc[w][tid];
        }
        if (part_d) {
            part_d[(size_t)blockIdx.x * nlags + tid] = d;
            part_g[(size_t)blockIdx.x * nlags + tid] = g;
            part_c[(size_t)blockIdx.x * nlags + tid] = c;
        } else if (c) {
            atomicAdd(&sum_d[tid], d);
            atomicAdd(&sum_g[tid], g);
            atomicAdd(&count[tid], c);
        }
    }
}

// NL = 8 / 16 / 32: the per-thread accumulators are REGISTERS, updated by predicated adds (bin == q), so the pair loop has no
// memory traffic besides the broadcast loads of the column sample - the generic version (NL = MAX_LAGS) keeps them in
// (L1-resident) local memory and is bound by those dependent load-add-store chains (ncu: long_scoreboard 18 stalls per
// issue, 106 instructions per pair).
template <int METRIC, int NL>
__global__ void __launch_bounds__(VT) variogram_bins_kernel(const double* __restrict__ x,
                                                            const double* __restrict__ y,
                                                            const double* __restrict__ z, long long n,
                                                            long long n_blocks, long long part, long long n_parts,
                                                            int nlags, const double* __restrict__ edges, double* sum_d,
                                                            double* sum_g, unsigned long long* count,
                                                            double* __restrict__ part_d, double* __restrict__ part_g,
                                                            unsigned long long* __restrict__ part_c) {
    __shared__ double2 sxy[VT];
    __shared__ double sz[VT];
    __shared__ double s_edges[MAX_LAGS + 2];
    __shared__ double s_wd[VT / 32][NL], s_wg[VT / 32][NL];
    __shared__ unsigned s_wc[VT / 32][NL];
    const int tid = threadIdx.x;
    if (tid <= nlags) s_edges[tid] = edges[tid];
    double acc_d[NL], acc_g[NL];
    unsigned acc_c[NL];
#pragma unroll
    for (int b = 0; b < NL; ++b) {
        acc_d[b] = 0.0;
        acc_g[b] = 0.0;
        acc_c[b] = 0u;
    }
    __syncthreads();
    const double e0 = s_edges[0], e_last = s_edges[nlags];
    const double inv_dd = nlags > 0 && s_edges[1] > e0 ? 1.0 / (s_edges[1] - e0) : 0.0;

    for (long long blk = part + n_parts * blockIdx.x; blk < n_blocks; blk += n_parts * gridDim.x) {
        int bi, bj;
        tile_of_block(blk, bi, bj);
        const long long i = (long long)bi * VT + tid;
        const long long j0 = (long long)bj * VT;
        __syncthreads();
        if (j0 + tid < n) {
            sxy[tid] = make_double2(x[j0 + tid], y[j0 + tid]);
            sz[tid] = z[j0 + tid];
        }
        __syncthreads();
        if (i < n) {
            const double xi = x[i], yi = y[i], zi = z[i];
            const int jn = (int)((n - j0) < VT ? (n - j0) : VT);
            const int jend = bi == bj ? (tid < jn ? tid : jn) : jn;
#pragma unroll 2
            for (int j = 0; j < jend; ++j) {
                const double2 pj = sxy[j];
                double d = pair_key<METRIC>(xi, yi, pj.x, pj.y);
                if (METRIC == CMX_METRIC_EUCLIDEAN) d = sqrt(d);
                bool in;
                const int b = variogram_bin(d, e0, e_last, inv_dd, nlags, s_edges, in);
                const double dz = zi - sz[j];
                const double g = 0.5 * (dz * dz);
                if (NL <= 32) {
#pragma unroll
                    for (int q = 0; q < NL; ++q) {
                        if (in && b == q) {
                            acc_d[q] += d;
                            acc_g[q] += g;
                            acc_c[q] += 1u;
                        }
                    }
                } else if (in) {
                    acc_d[b] += d;
                    acc_g[b] += g;
                    acc_c[b] += 1u;
                }
            }
        }
    }
    // Deterministic reduction: lanes by shuffles, the warps of the CTA in warp order, the CTAs (whose share of the pair
    // blocks is fixed by the launch geometry) in CTA order by variogram_reduce_kernel - no floating-point atomics, so
    // the sums, the fitted parameters and the kriged field are reproducible from run to run.  Without scratch
    // (part_d == nullptr) the CTA results are accumulated with atomics.
#pragma unroll
    for (int b = 0; b < NL; ++b) {
        if (b < nlags) {
            const double d = warp_sum(acc_d[b]);
            const double g = warp_sum(acc_g[b]);
            unsigned c = acc_c[b];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(kFull, c, o);
            if ((tid & 31) == 0) {
                s_wd[tid >> 5][b] = d;
                s_wg[tid >> 5][b] = g;
                s_wc[tid >> 5][b] = c;
            }
        }
    }
    __syncthreads();
    if (tid < nlags) {
        double d = 0.0, g = 0.0;
        unsigned long long c = 0ull;
        for (int w = 0; w < VT / 32; ++w) {
            d += s_wd[w][tid];
            g += s_wg[w][tid];
            c += s_wc[w][tid];
        }
        if (part_d) {
            part_d[(size_t)blockIdx.x * nlags + tid] = d;
            part_g[(size_t)blockIdx.x * nlags + tid] = g;
            part_c[(size_t)blockIdx.x * nlags + tid] = c;
        } else if (c) {
            atomicAdd(&sum_d[tid], d);
            atomicAdd(&sum_g[tid], g);
            atomicAdd(&count[tid], c);
        }
    }
}

// sums the per-CTA partial results in a fixed order and ADDS them to the outputs: one warp per bin, lane l takes the
// CTAs l, l + 32, ... in order, then a shuffle tree (the same tree every run: deterministic)
__global__ void variogram_reduce_kernel(const double* __restrict__ part_d, const double* __restrict__ part_g,
                                        const unsigned long long* __restrict__ part_c, int n_ctas, int nlags, double* sum_d,
                                        double* sum_g, unsigned long long* count) {
    const int b = blockIdx.x, lane = threadIdx.x;
    if (b >= nlags) return;
    double d = 0.0, g = 0.0;
    unsigned long long c = 0ull;
    for (int i = lane; i < n_ctas; i += 32) {
        d += part_d[(size_t)i * nlags + b];
        g += part_g[(size_t)i * nlags + b];
        c += part_c[(size_t)i * nlags + b];
    }
    d = warp_sum(d);
    g = warp_sum(g);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(kFull, c, o);
    if (lane == 0) {
        sum_d[b] += d;
        sum_g[b] += g;
        count[b] += c;
    }
}

// bin edges exactly as pykrige builds them: dmin + i * dd for i < nlags, last edge dmax + 0.001
__global__ void variogram_edges_kernel(const double* __restrict__ minmax, int nlags, double* __restrict__ edges) {
    const int i = threadIdx.x;
    const double dmin = minmax[0], dmax = minmax[1];
    const double dd = __ddiv_rn(__dsub_rn(dmax, dmin), (double)nlags);
    if (i < nlags) edges[i] = __dadd_rn(dmin, __dmul_rn((double)i, dd));
    if (i == nlags) edges[i] = __dadd_rn(dmax, 0.001);
}

// lags / semivariances of the non-empty bins, then the TRF fit - one thread (the problem has <= 64 residuals and
// 2-3 parameters; the iteration is a serial chain of dependent double-precision operations)
__global__ void variogram_fit_kernel(int model, int nlags, const double* __restrict__ sum_d, const double* __restrict__ sum_g,
                                     const unsigned long long* __restrict__ count, double* __restrict__ params,
                                     double* __restrict__ lags_out, double* __restrict__ sv_out, int* __restrict__ info) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double lags[vfit::MAXM], sv[vfit::MAXM];
    int m = 0;
    for (int b = 0; b < nlags; ++b) {
        if (count[b] > 0ull) {
            lags[m] = sum_d[b] / (double)count[b];
            sv[m] = sum_g[b] / (double)count[b];
            ++m;
        }
    }
    int nfev = 0;
    const int status = m > 0 ? vfit::variogram_fit_trf(model, m, lags, sv, params, &nfev) : -1;
    for (int b = 0; b < nlags; ++b) {
        if (lags_out) lags_out[b] = b < m ? lags[b] : 0.0;
        if (sv_out) sv_out[b] = b < m ? sv[b] : 0.0;
    }
    if (info) {
        info[0] = status;
        info[1] = m;
        info[2] = nfev;
    }
}

}  // namespace
}  // namespace cmx

extern "C" {

int cmx_variogram_fit_host(int model, const double* lags, const double* semivariance, int n, double params_out[3],
                           int* nfev_out) {
    if (!lags || !semivariance || !params_out || n < 1 || n > cmx::vfit::MAXM) return CMX_ERR_BAD_ARG;
    if (model < CMX_VARIOGRAM_LINEAR || model > CMX_VARIOGRAM_EXPONENTIAL) return CMX_ERR_UNSUPPORTED;
    const int status = cmx::vfit::variogram_fit_trf(model, n, lags, semivariance, params_out, nfev_out);
    return status < 0 ? CMX_ERR_BAD_ARG : status;
}

int cmx_variogram_edges(const double* minmax, int nlags, double* edges, void* stream_) {
    using namespace cmx;
    if (!minmax || !edges || nlags < 1) return CMX_ERR_BAD_ARG;
    if (nlags > MAX_LAGS) return CMX_ERR_UNSUPPORTED;
    variogram_edges_kernel<<<1, MAX_LAGS + 1, 0, static_cast<cudaStream_t>(stream_)>>>(minmax, nlags, edges);
    note_launch();
    return check_launch();
}

int cmx_variogram_fit(int model, int nlags, const double* sum_d, const double* sum_g, const unsigned long long* count,
                      double* params_out, double* lags_out, double* semivariance_out, int32_t* info, void* stream_) {
    using namespace cmx;
    if (!sum_d || !sum_g || !count || !params_out || nlags < 1) return CMX_ERR_BAD_ARG;
    if (nlags > MAX_LAGS) return CMX_ERR_UNSUPPORTED;
    if (model < CMX_VARIOGRAM_LINEAR || model > CMX_VARIOGRAM_EXPONENTIAL) return CMX_ERR_UNSUPPORTED;
    variogram_fit_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream_)>>>(model, nlags, sum_d, sum_g, count, params_out,
                                                                           lags_out, semivariance_out, info);
    note_launch();
    return check_launch();
}

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

size_t cmx_variogram_workspace_bytes(int nlags) {
    if (nlags < 1 || nlags > cmx::MAX_LAGS) return 0;
    return (size_t)cmx::sm_count() * 8 * nlags * (2 * sizeof(double) + sizeof(unsigned long long)) + 256;
}

int cmx_variogram_bins(const double* x, const double* y, const double* z, int64_t n, int metric, int part,
                       int n_parts, int nlags, const double* edges, double* sum_d, double* sum_g,
                       unsigned long long* count, void* workspace, size_t workspace_bytes, void* stream_) {
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
    // scratch for the deterministic two-stage reduction: [grid][nlags] x (sum_d, sum_g, count); optional
    double *part_d = nullptr, *part_g = nullptr;
    unsigned long long* part_c = nullptr;
    if (workspace) {
        unsigned char* w = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(workspace) + 15) & ~uintptr_t(15));
        const size_t need = (size_t)grid * nlags * (2 * sizeof(double) + sizeof(unsigned long long));
        if (w + need > static_cast<unsigned char*>(workspace) + workspace_bytes) return CMX_ERR_WORKSPACE;
        part_d = reinterpret_cast<double*>(w);
        part_g = part_d + (size_t)grid * nlags;
        part_c = reinterpret_cast<unsigned long long*>(part_g + (size_t)grid * nlags);
    }
#define CMX_BINS_LAUNCH(METRIC, NL)                                                                                   \
    variogram_bins_kernel<METRIC, NL><<<grid, VT, 0, stream>>>(x, y, z, n, n_blocks, part, n_parts, nlags, edges, sum_d, \
                                                               sum_g, count, part_d, part_g, part_c)
#ifndef CMX_K3_REGISTERS  // tuning builds: -DCMX_K3_REGISTERS keeps the register-accumulator kernels for nlags <= 32
    if (nlags <= SM_LAGS) {
        const size_t smem = (size_t)nlags * VT * (2 * sizeof(double) + sizeof(unsigned));
        if (metric == CMX_METRIC_EUCLIDEAN) {
            CMX_CUDA(cudaFuncSetAttribute(variogram_bins_smem_kernel<CMX_METRIC_EUCLIDEAN>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            variogram_bins_smem_kernel<CMX_METRIC_EUCLIDEAN><<<grid, VT, smem, stream>>>(
                x, y, z, n, n_blocks, part, n_parts, nlags, edges, sum_d, sum_g, count, part_d, part_g, part_c);
        } else {
            CMX_CUDA(cudaFuncSetAttribute(variogram_bins_smem_kernel<CMX_METRIC_GEOGRAPHIC>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            variogram_bins_smem_kernel<CMX_METRIC_GEOGRAPHIC><<<grid, VT, smem, stream>>>(
                x, y, z, n, n_blocks, part, n_parts, nlags, edges, sum_d, sum_g, count, part_d, part_g, part_c);
        }
    } else
#endif
    if (metric == CMX_METRIC_EUCLIDEAN) {
        if (nlags <= 8) CMX_BINS_LAUNCH(CMX_METRIC_EUCLIDEAN, 8);
        else if (nlags <= 16) CMX_BINS_LAUNCH(CMX_METRIC_EUCLIDEAN, 16);
        else if (nlags <= 32) CMX_BINS_LAUNCH(CMX_METRIC_EUCLIDEAN, 32);
        else CMX_BINS_LAUNCH(CMX_METRIC_EUCLIDEAN, MAX_LAGS);
    } else {
        if (nlags <= 8) CMX_BINS_LAUNCH(CMX_METRIC_GEOGRAPHIC, 8);
        else if (nlags <= 16) CMX_BINS_LAUNCH(CMX_METRIC_GEOGRAPHIC, 16);
        else if (nlags <= 32) CMX_BINS_LAUNCH(CMX_METRIC_GEOGRAPHIC, 32);
        else CMX_BINS_LAUNCH(CMX_METRIC_GEOGRAPHIC, MAX_LAGS);
    }
#undef CMX_BINS_LAUNCH
    note_launch();
    if (part_d) {
        variogram_reduce_kernel<<<nlags, 32, 0, stream>>>(part_d, part_g, part_c, grid, nlags, sum_d, sum_g, count);
        note_launch();
    }
    return check_launch();
}

}  // extern "C"
