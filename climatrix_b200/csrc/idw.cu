// K2: IDW weighting applied to a stored neighbour table, batched over time / level.
//
// Reference arithmetic: src/climatrix/reconstruct/idw.py:163-180, once per slice
// (the reference rejects dynamic datasets, idw.py:87-104, so a batch is a loop of
// static calls; the neighbour query does not depend on the slice and is shared).
//
//   out[t][m] = sum_j wn[m][j] * v[idx[m][j]][t]  /  sum_j wn[m][j]      (finite v only)
//   NaN where fewer than k_min finite neighbours.
//
// HBM-write bound: algorithmic bytes per target = n_batch * sizeof(V) written
// (+ the [M][k] table read once).  One CTA owns 32 consecutive targets for the
// whole batch: weights are computed once, samples' value rows are gathered with
// the lanes running along the batch axis (coalesced when values are sample-major),
// and results are transposed through shared memory so that every store is a full
// 32-target row segment.
#include "common.cuh"

namespace cmx {
namespace {

constexpr int BW_WARPS = 8;
constexpr int BW_THREADS = BW_WARPS * 32;
constexpr int BW_TARGETS = 32;  // targets per CTA

// Per (target, neighbour) the CTA keeps one 16-byte shared-memory entry {weight, byte offset of the
// sample's value row}: the inner loop is LDS.128 (broadcast) + one 64-bit add + LDG + DFMA -- the row
// offset (a 64-bit multiply) is computed once per target, not once per value.
struct __align__(16) NbrEntry {
    double w;
    unsigned long long off;
};

template <typename V>
__global__ void __launch_bounds__(BW_THREADS) idw_batch_kernel(
    const int* __restrict__ idx, const double* __restrict__ dd, int dist_is_squared, long long M, int k_table,
    int k_use, int k_min, double power, const V* __restrict__ vals, long long n_samples, long long n_batch,
    long long stride_t, long long stride_i, const int* __restrict__ nonfinite_flag, V* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    NbrEntry* s_ent = reinterpret_cast<NbrEntry*>(smem_raw);               // [32][k_use]
    double* s_wsum = reinterpret_cast<double*>(s_ent + BW_TARGETS * k_use);  // [32]
    V* s_tile = reinterpret_cast<V*>(s_wsum + BW_TARGETS);                   // [32][33]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long m0 = (long long)blockIdx.x * BW_TARGETS;
    const bool check = nonfinite_flag ? *nonfinite_flag != 0 : true;

    // ---- weights, once per target (idw.py:163-165) ----
    for (int tl = warp; tl < BW_TARGETS; tl += BW_WARPS) {
        const long long m = m0 + tl;
        double wsum = 0.0;
        for (int j = lane; j < k_use; j += 32) {
            double w = 0.0;
            long long id = 0;
            if (m < M) {
                const double d = dd[m * k_table + j];
                id = idx[m * k_table + j];
                if (id >= 0 && id < n_samples)  // the (inf, INT_MAX) sentinel of a short neighbour list carries no weight
                    w = dist_is_squared ? idw_raw_weight(d, power) : idw_weight_from_dist(d, power);
                else
                    id = 0;
            }
            s_ent[tl * k_use + j].w = w;
            s_ent[tl * k_use + j].off = (unsigned long long)(id * stride_i) * sizeof(V);
            wsum += w;
        }
        wsum = warp_sum(wsum);
        __syncwarp();
        double wn = 0.0;
        for (int j = lane; j < k_use; j += 32) {
            const double w = s_ent[tl * k_use + j].w / wsum;
            s_ent[tl * k_use + j].w = w;
            wn += w;
        }
        wn = warp_sum(wn);
        if (lane == 0) s_wsum[tl] = wn;
    }
    __syncthreads();

    // ---- batch sweep: lanes run along t ----
    for (long long t0 = 0; t0 < n_batch; t0 += 32) {
        const long long t = t0 + lane;
        const bool tin = t < n_batch;
        const char* lane_base = reinterpret_cast<const char*>(vals) + (tin ? t : 0) * stride_t * (long long)sizeof(V);
        for (int tl = warp; tl < BW_TARGETS; tl += BW_WARPS) {
            const NbrEntry* e = s_ent + tl * k_use;
            V r;
            if (!check) {
                double acc = 0.0;
#pragma unroll 8
                for (int j = 0; j < k_use; ++j) {
                    const NbrEntry q = e[j];
                    acc = fma(q.w, (double)*reinterpret_cast<const V*>(lane_base + q.off), acc);
                }
                r = (V)(acc / s_wsum[tl]);
            } else {
                double num = 0.0, den = 0.0;
                int fin = 0;
#pragma unroll 4
                for (int j = 0; j < k_use; ++j) {
                    const NbrEntry q = e[j];
                    const double v = (double)*reinterpret_cast<const V*>(lane_base + q.off);
                    if (isfinite(v) && q.w > 0.0) {
                        num = fma(q.w, v, num);
                        den += q.w;
                        ++fin;
                    }
                }
                r = (fin < k_min || den == 0.0) ? quiet_nan<V>() : (V)(num / den);
            }
            s_tile[lane * 33 + tl] = r;  // [t_local][target_local]
        }
        __syncthreads();
        // coalesced stores: one 32-target row segment per (warp, t_local)
        for (int tr = warp; tr < 32; tr += BW_WARPS) {
            const long long tt = t0 + tr;
            const long long m = m0 + lane;
            if (tt < n_batch && m < M) out[tt * M + m] = s_tile[tr * 33 + lane];
        }
        __syncthreads();
    }
}

// single slice from a table: one warp per target (HPO inner loop, small problems)
template <typename V>
__global__ void idw_apply_single_kernel(const int* __restrict__ idx, const double* __restrict__ dd,
                                        int dist_is_squared, long long M, int k_table, int k_use, int k_min,
                                        double power, const V* __restrict__ vals, long long n_samples,
                                        long long stride_i, V* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long m = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (m >= M) return;
    double w[CMX_MAX_K / 32];
    double wsum = 0.0;
#pragma unroll
    for (int e = 0; e < CMX_MAX_K / 32; ++e) {
        const int j = e * 32 + lane;
        w[e] = 0.0;
        if (j < k_use) {
            const double d = dd[m * k_table + j];
            const int id = idx[m * k_table + j];
            if (id >= 0 && id < n_samples) w[e] = dist_is_squared ? idw_raw_weight(d, power) : idw_weight_from_dist(d, power);
        }
        wsum += w[e];
    }
    wsum = warp_sum(wsum);
    double num = 0.0, den = 0.0;
    int fin = 0;
#pragma unroll
    for (int e = 0; e < CMX_MAX_K / 32; ++e) {
        const int j = e * 32 + lane;
        if (j < k_use && w[e] > 0.0) {
            const double wn = w[e] / wsum;
            const double v = (double)vals[(long long)idx[m * k_table + j] * stride_i];
            if (isfinite(v)) {
                num += v * wn;
                den += wn;
                ++fin;
            }
        }
    }
    num = warp_sum(num);
    den = warp_sum(den);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) fin += __shfl_xor_sync(kFull, fin, o);
    if (lane == 0) out[m] = (fin < k_min || den == 0.0) ? quiet_nan<V>() : (V)(num / den);
}

template <typename V>
__global__ void flag_nonfinite_kernel(const V* __restrict__ v, long long count, int* flag) {
    bool bad = false;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (long long)gridDim.x * blockDim.x)
        bad |= !isfinite((double)v[i]);
    if (__any_sync(kFull, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

template <typename V>
int idw_from_table(const int* idx, const double* dd, int dist_is_squared, long long M, int k_table, int k_use,
                   int k_min, double power, const V* vals, long long n_samples, long long n_batch, int layout,
                   V* out, int* flag_scratch, cudaStream_t stream) {
    if (M == 0 || n_batch == 0) return CMX_OK;
    const long long stride_t = layout == CMX_LAYOUT_SAMPLE_MAJOR ? 1 : n_samples;
    const long long stride_i = layout == CMX_LAYOUT_SAMPLE_MAJOR ? n_batch : 1;
    if (n_batch == 1) {
        const int threads = 256;
        const long long blocks = (M * 32 + threads - 1) / threads;
        idw_apply_single_kernel<V><<<(unsigned)blocks, threads, 0, stream>>>(idx, dd, dist_is_squared, M, k_table,
                                                                            k_use, k_min, power, vals, n_samples,
                                                                            stride_i, out);
        note_launch();
        return check_launch();
    }
    if (flag_scratch) {  // lets the batched kernel skip the per-value finiteness test
        CMX_CUDA(cudaMemsetAsync(flag_scratch, 0, sizeof(int), stream));
        const int fblocks = sm_count() * 8;
        flag_nonfinite_kernel<V><<<fblocks, 256, 0, stream>>>(vals, n_samples * n_batch, flag_scratch);
        note_launch();
    }
    const size_t smem = (size_t)BW_TARGETS * k_use * sizeof(NbrEntry) + BW_TARGETS * sizeof(double) + 32 * 33 * sizeof(V);
    const long long blocks = (M + BW_TARGETS - 1) / BW_TARGETS;
    if (smem > 48 * 1024)
        CMX_CUDA(cudaFuncSetAttribute(idw_batch_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    timing_begin(CMX_KERNEL_IDW_BATCH, stream);
    idw_batch_kernel<V><<<(unsigned)blocks, BW_THREADS, smem, stream>>>(idx, dd, dist_is_squared, M, k_table, k_use,
                                                                       k_min, power, vals, n_samples, n_batch,
                                                                       stride_t, stride_i, flag_scratch, out);
    timing_end(CMX_KERNEL_IDW_BATCH, stream);
    note_launch();
    return check_launch();
}

template <typename V>
int idw_entry(const double* s_lat, const double* s_lon, int64_t n, const V* vals, int64_t n_batch, int layout,
              const double* t_lat, int64_t n_lat, const double* t_lon, int64_t n_lon, int grid, int k, int k_min,
              double power, V* out, int32_t* idx_out, double* dist_out, void* workspace, size_t workspace_bytes,
              void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!vals || !out || n_batch < 1 || k_min < 1) return CMX_ERR_BAD_ARG;
    if (layout != CMX_LAYOUT_BATCH_MAJOR && layout != CMX_LAYOUT_SAMPLE_MAJOR) return CMX_ERR_BAD_ARG;
    if (n_lat < 0 || n_lon < 0) return CMX_ERR_BAD_ARG;
    const long long M = grid ? (long long)n_lat * n_lon : (long long)n_lat;
    const size_t knn_ws = knn_workspace_bytes(k);
    KnnJob job{};
    job.s_a = s_lat;
    job.s_b = s_lon;
    job.n = n;
    job.t_a = t_lat;
    job.t_b = t_lon;
    job.n_a = n_lat;
    job.n_b = n_lon;
    job.grid = grid;
    job.k = k;
    job.idx_out = idx_out;
    job.dist_out = dist_out;
    job.k_min = k_min;
    job.power = power;
    job.value_is_f32 = sizeof(V) == 4;
    job.workspace = workspace;
    job.workspace_bytes = workspace_bytes;
    if (n_batch == 1) {
        job.idw_mode = 1;  // fused: weights + weighted mean in the kNN kernel's epilogue
        job.vals = vals;
        job.out = out;
        return run_knn(job, stream);
    }
    // batch: neighbour table once, then the HBM-bound batched apply.  Workspace: [tables | kNN scratch]
    if (k < 1 || k > CMX_MAX_K) return k < 1 ? CMX_ERR_BAD_ARG : CMX_ERR_K_TOO_LARGE;
    const size_t extra_bytes = (cmx_idw_workspace_bytes(k, M, n_batch) + 255) & ~size_t(255);
    if (!workspace || workspace_bytes < knn_ws + extra_bytes + 256) return CMX_ERR_WORKSPACE;
    unsigned char* extra = static_cast<unsigned char*>(workspace);
    extra = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(extra) + 255) & ~uintptr_t(255));
    double* d2_tab = reinterpret_cast<double*>(extra);
    int* idx_tab = reinterpret_cast<int*>(extra + (size_t)M * k * sizeof(double));
    int* flag = reinterpret_cast<int*>(extra + (size_t)M * k * (sizeof(double) + sizeof(int)));
    job.idw_mode = 0;
    job.d2_out = d2_tab;
    if (!job.idx_out) job.idx_out = idx_tab;
    job.workspace = extra + extra_bytes;
    job.workspace_bytes = workspace_bytes - (size_t)((extra + extra_bytes) - static_cast<unsigned char*>(workspace));
    int rc = run_knn(job, stream);
    if (rc != CMX_OK) return rc;
    return idw_from_table<V>(job.idx_out, d2_tab, 1, M, k, k, k_min, power, vals, n, n_batch, layout, out, flag,
                             stream);
}

template <typename V>
int apply_entry(const int32_t* idx, const double* dist, int64_t M, int k_table, const V* vals, int64_t n,
                int64_t n_batch, int layout, int k_use, int k_min, double power, V* out, int32_t* flag_scratch,
                void* stream_) {
    if (!idx || !dist || !vals || !out) return CMX_ERR_BAD_ARG;
    if (M < 0 || n_batch < 1 || k_use < 1 || k_use > k_table || k_min < 1) return CMX_ERR_BAD_ARG;
    if (k_use > CMX_MAX_K) return CMX_ERR_K_TOO_LARGE;
    if (layout != CMX_LAYOUT_BATCH_MAJOR && layout != CMX_LAYOUT_SAMPLE_MAJOR) return CMX_ERR_BAD_ARG;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    // without scratch the batched kernel tests every value for finiteness itself
    return idw_from_table<V>(idx, dist, 0, M, k_table, k_use, k_min, power, vals, n, n_batch, layout, out,
                             flag_scratch, stream);
}

}  // namespace
}  // namespace cmx

extern "C" {

size_t cmx_idw_workspace_bytes(int k, int64_t n_targets, int64_t n_batch) {
    if (n_batch <= 1 || k < 1 || n_targets < 0) return 0;
    return (size_t)n_targets * (size_t)k * (sizeof(double) + sizeof(int)) + 1024;
}

int cmx_idw_f64(const double* s_lat, const double* s_lon, int64_t n, const double* vals, int64_t n_batch, int layout,
                const double* t_lat, int64_t n_lat, const double* t_lon, int64_t n_lon, int grid, int k, int k_min,
                double power, double* out, int32_t* idx_out, double* dist_out, void* ws, size_t ws_bytes,
                void* stream) {
    return cmx::idw_entry<double>(s_lat, s_lon, n, vals, n_batch, layout, t_lat, n_lat, t_lon, n_lon, grid, k, k_min,
                                  power, out, idx_out, dist_out, ws, ws_bytes, stream);
}
int cmx_idw_f32(const double* s_lat, const double* s_lon, int64_t n, const float* vals, int64_t n_batch, int layout,
                const double* t_lat, int64_t n_lat, const double* t_lon, int64_t n_lon, int grid, int k, int k_min,
                double power, float* out, int32_t* idx_out, double* dist_out, void* ws, size_t ws_bytes,
                void* stream) {
    return cmx::idw_entry<float>(s_lat, s_lon, n, vals, n_batch, layout, t_lat, n_lat, t_lon, n_lon, grid, k, k_min,
                                 power, out, idx_out, dist_out, ws, ws_bytes, stream);
}
int cmx_idw_apply_f64(const int32_t* idx, const double* dist, int64_t M, int k_table, const double* vals, int64_t n,
                      int64_t n_batch, int layout, int k_use, int k_min, double power, double* out,
                      int32_t* flag_scratch, void* stream) {
    return cmx::apply_entry<double>(idx, dist, M, k_table, vals, n, n_batch, layout, k_use, k_min, power, out,
                                    flag_scratch, stream);
}
int cmx_idw_apply_f32(const int32_t* idx, const double* dist, int64_t M, int k_table, const float* vals, int64_t n,
                      int64_t n_batch, int layout, int k_use, int k_min, double power, float* out,
                      int32_t* flag_scratch, void* stream) {
    return cmx::apply_entry<float>(idx, dist, M, k_table, vals, n, n_batch, layout, k_use, k_min, power, out,
                                   flag_scratch, stream);
}

}  // extern "C"
