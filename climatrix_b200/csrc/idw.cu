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
    long long stride_t, long long stride_i, const int* __restrict__ nonfinite_flag, const PeerOut po) {
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
        int real = 0;  // neighbours that exist: a short list (NaN sample coordinates) ends in (inf, INT_MAX) sentinels
        for (int j = lane; j < k_use; j += 32) {
            double w = 0.0;
            long long id = 0;
            if (m < M) {
                const double d = dd[m * k_table + j];
                id = idx[m * k_table + j];
                if (id >= 0 && id < n_samples) {  // a sentinel carries no weight
                    w = dist_is_squared ? idw_raw_weight(d, power) : idw_weight_from_dist(d, power);
                    ++real;
                } else {
                    id = 0;
                }
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
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) real += __shfl_xor_sync(kFull, real, o);
        // fewer than k_min existing neighbours -> NaN also on the path that skips the per-value finiteness test
        // (a NaN divisor), exactly like the single-slice epilogue (idw.py:179-180)
        if (lane == 0) s_wsum[tl] = real < k_min ? quiet_nan<double>() : wn;
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
            if (tt < n_batch && m < M) {
                const V r = s_tile[tr * 33 + lane];
                for (int p = 0; p < po.n; ++p) static_cast<V*>(po.base[p])[tt * po.m_total + po.m_off + m] = r;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// K2, neighbour-sharing variant (default for k <= 32).
//
// Adjacent grid targets share most of their neighbours (a run of 16 targets touches ~50-100 distinct
// samples, not 16 * 32).  The kernel above fetches every (target, neighbour) value separately and is
// bound by those gathers (ncu: LSU pipe 46 %, long-scoreboard stalls on L1 misses).  Here a CTA of 16
// consecutive targets first de-duplicates its neighbour indices in shared memory (open-addressing hash
// on the sample index; slots are numbered by FIRST OCCURRENCE in (target, rank) order through a block
// prefix sum, so the summation order - and with it every output bit - is deterministic), scatters the
// normalised weights into a dense W[union][16] table and records per warp a bit mask of the union
// entries its 4 targets use.  In the sweep a warp walks that mask: ONE value load per union sample and
// batch element feeds 4 targets (W row read by two 16-byte broadcast loads), two batch elements per
// thread, i.e. 2 global loads per 8 fma instead of 8 per 8.  CTAs whose union exceeds the table, values
// with non-finite entries and k > 32 take the per-target loop (same results).
// ------------------------------------------------------------------------------------------
constexpr int UT = 16;            // targets per CTA
constexpr int UWARPS = 4;         // 4 targets per warp
constexpr int UTHREADS = UWARPS * 32;
constexpr int UMAX = 128;         // union capacity
constexpr int UHS = 1024;         // hash slots (<= 512 entries)
constexpr int UK = 32;            // largest k for the shared sweep

template <typename V, int TT>
__device__ __forceinline__ void union_sweep_chunk(const unsigned char* s_ulist, int n_list,
                                                  const unsigned long long* s_uoff, const double* s_W,
                                                  const double* s_wsum, V* s_tile, const V* vals, long long t0,
                                                  long long n_batch, long long stride_t, int warp, int lane) {
    const char* base[TT];
#pragma unroll
    for (int tt = 0; tt < TT; ++tt) {
        const long long t = t0 + tt * 32 + lane;
        base[tt] = reinterpret_cast<const char*>(vals) + (t < n_batch ? t : 0) * stride_t * (long long)sizeof(V);
    }
    double acc[4][TT];
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int tt = 0; tt < TT; ++tt) acc[g][tt] = 0.0;
    // the list is padded to a multiple of 4 with the all-zero row UMAX: four union samples (8 loads) in flight
#pragma unroll 1
    for (int j = 0; j < n_list; j += 4) {
        const unsigned u4 = *reinterpret_cast<const unsigned*>(s_ulist + warp * (UMAX + 4) + j);
        double v[4][TT];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const unsigned long long off = s_uoff[(u4 >> (8 * q)) & 0xffu];
#pragma unroll
            for (int tt = 0; tt < TT; ++tt) v[q][tt] = (double)*reinterpret_cast<const V*>(base[tt] + off);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int u = (u4 >> (8 * q)) & 0xffu;
            const double2 w01 = *reinterpret_cast<const double2*>(s_W + u * UT + warp * 4);
            const double2 w23 = *reinterpret_cast<const double2*>(s_W + u * UT + warp * 4 + 2);
#pragma unroll
            for (int tt = 0; tt < TT; ++tt) {
                acc[0][tt] = fma(w01.x, v[q][tt], acc[0][tt]);
                acc[1][tt] = fma(w01.y, v[q][tt], acc[1][tt]);
                acc[2][tt] = fma(w23.x, v[q][tt], acc[2][tt]);
                acc[3][tt] = fma(w23.y, v[q][tt], acc[3][tt]);
            }
        }
    }
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int tt = 0; tt < TT; ++tt)
            s_tile[(tt * 32 + lane) * (UT + 1) + warp * 4 + g] = (V)(acc[g][tt] / s_wsum[warp * 4 + g]);
}

template <typename V>
__global__ void __launch_bounds__(UTHREADS, 5) idw_batch_union_kernel(
    const int* __restrict__ idx, const double* __restrict__ dd, int dist_is_squared, long long M, int k_table,
    int k_use, int k_min, double power, const V* __restrict__ vals, long long n_samples, long long n_batch,
    long long stride_t, long long stride_i, const int* __restrict__ nonfinite_flag, const PeerOut po) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    NbrEntry* s_ent = reinterpret_cast<NbrEntry*>(smem_raw);                       // [UT][k_use]
    double* s_W = reinterpret_cast<double*>(s_ent + UT * k_use);                   // [UMAX + 1][UT], row UMAX stays zero
    unsigned long long* s_uoff = reinterpret_cast<unsigned long long*>(s_W + (UMAX + 1) * UT);  // [UMAX + 1]
    double* s_wsum = reinterpret_cast<double*>(s_uoff + UMAX + 1);                 // [UT]
    // the output tile of the sweep reuses the hash storage of the set-up
    V* s_tile = reinterpret_cast<V*>(s_wsum + UT);                                 // [64][UT + 1]   (10 KB region)
    int* s_key = reinterpret_cast<int*>(s_wsum + UT);                              // [UT * UK] sample index
    int* s_hkey = s_key + UT * UK;                                                 // [UHS]
    int* s_hmin = s_hkey + UHS;                                                    // [UHS]
    unsigned (*s_mask)[4] = reinterpret_cast<unsigned (*)[4]>(s_hmin + UHS);       // [UWARPS][4]
    int* s_scan = reinterpret_cast<int*>(s_mask + UWARPS);                         // [UWARPS + 1]
    unsigned char* s_hval = reinterpret_cast<unsigned char*>(s_scan + UWARPS + 1 + 3);  // [UHS]
    unsigned char* s_ulist = s_hval + UHS;                                         // [UWARPS][UMAX + 4], 4-byte aligned rows

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long m0 = (long long)blockIdx.x * UT;
    const bool check = nonfinite_flag ? *nonfinite_flag != 0 : true;
    const bool shared_sweep = !check && k_use <= UK;

    // ---- weights, once per target (idw.py:163-165) ----
    for (int tl = warp; tl < UT; tl += UWARPS) {
        const long long m = m0 + tl;
        double wsum = 0.0;
        int real = 0;  // neighbours that exist: a short list (NaN sample coordinates) ends in (inf, INT_MAX) sentinels
        for (int j = lane; j < k_use; j += 32) {
            double w = 0.0;
            long long id = 0;
            if (m < M) {
                const double d = dd[m * k_table + j];
                id = idx[m * k_table + j];
                if (id >= 0 && id < n_samples) {  // a sentinel carries no weight
                    w = dist_is_squared ? idw_raw_weight(d, power) : idw_weight_from_dist(d, power);
                    ++real;
                } else {
                    id = 0;
                }
            }
            s_ent[tl * k_use + j].w = w;
            s_ent[tl * k_use + j].off = (unsigned long long)(id * stride_i) * sizeof(V);
            if (shared_sweep) s_key[tl * k_use + j] = (int)id;
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
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) real += __shfl_xor_sync(kFull, real, o);
        // fewer than k_min existing neighbours -> NaN also on the path that skips the per-value finiteness test
        // (a NaN divisor), exactly like the single-slice epilogue (idw.py:179-180)
        if (lane == 0) s_wsum[tl] = real < k_min ? quiet_nan<double>() : wn;
    }
    bool use_union = false;
    if (shared_sweep) {
        for (int i = tid; i < UHS; i += UTHREADS) {
            s_hkey[i] = -1;
            s_hmin[i] = 0x7fffffff;
        }
        for (int i = tid; i < (UMAX + 1) * UT; i += UTHREADS) s_W[i] = 0.0;
        if (tid == 0) s_uoff[UMAX] = 0ull;
        if (tid < UWARPS * 4) (&s_mask[0][0])[tid] = 0u;
        __syncthreads();
        // ---- de-duplicate: entries e = target * k + rank, EPT consecutive entries per thread ----
        constexpr int EPT = UT * UK / UTHREADS;  // 4
        const int n_ent = UT * k_use;
        int myh[EPT];
        bool valid[EPT];
#pragma unroll
        for (int i = 0; i < EPT; ++i) {
            const int e = tid * EPT + i;
            valid[i] = e < n_ent && s_ent[e].w > 0.0;
            myh[i] = 0;
            if (valid[i]) {
                const int key = s_key[e];
                unsigned h = ((unsigned)key * 2654435761u) >> 22;  // 10 bits
                for (;;) {
                    const int old = atomicCAS(&s_hkey[h], -1, key);
                    if (old == -1 || old == key) break;
                    h = (h + 1) & (UHS - 1);
                }
                atomicMin(&s_hmin[h], e);
                myh[i] = (int)h;
            }
        }
        __syncthreads();
        int lead = 0;
        bool is_lead[EPT];
#pragma unroll
        for (int i = 0; i < EPT; ++i) {
            is_lead[i] = valid[i] && s_hmin[myh[i]] == tid * EPT + i;
            lead += is_lead[i];
        }
        int incl = lead;  // block prefix sum of the leader counts, in entry order
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int up = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += up;
        }
        if (lane == 31) s_scan[warp] = incl;
        __syncthreads();
        int base = incl - lead, total = 0;
#pragma unroll
        for (int w = 0; w < UWARPS; ++w) {
            if (w < warp) base += s_scan[w];
            total += s_scan[w];
        }
        use_union = total <= UMAX;  // CTA-uniform
        if (use_union) {
#pragma unroll
            for (int i = 0; i < EPT; ++i) {
                if (is_lead[i]) {
                    s_hval[myh[i]] = (unsigned char)base;
                    s_uoff[base] = s_ent[tid * EPT + i].off;
                    ++base;
                }
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < EPT; ++i) {
                if (valid[i]) {
                    const int e = tid * EPT + i, g = e / k_use;
                    const int u = s_hval[myh[i]];
                    s_W[u * UT + g] = s_ent[e].w;
                    atomicOr(&s_mask[g >> 2][u >> 5], 1u << (u & 31));
                }
            }
        }
    }
    __syncthreads();

    if (use_union) {
        // ---- per warp: compact list of the union entries its 4 targets use, padded with the zero row ----
        int n_list = 0;
        for (int word = 0; word < UMAX / 32; ++word) {
            const unsigned bits = s_mask[warp][word];
            if ((bits >> lane) & 1u) s_ulist[warp * (UMAX + 4) + n_list + __popc(bits & ((1u << lane) - 1u))] = (unsigned char)(word * 32 + lane);
            n_list += __popc(bits);
        }
        if (lane < 4) s_ulist[warp * (UMAX + 4) + n_list + lane] = (unsigned char)UMAX;
        n_list = (n_list + 3) & ~3;
        __syncthreads();  // lists visible; the hash storage may now be reused as the output tile
        // ---- shared sweep: 64 batch elements per pass (two per thread), 32 in a short tail ----
        for (long long t0 = 0; t0 < n_batch; t0 += 64) {
            const bool two = n_batch - t0 > 32;
            if (two)
                union_sweep_chunk<V, 2>(s_ulist, n_list, s_uoff, s_W, s_wsum, s_tile, vals, t0, n_batch, stride_t, warp, lane);
            else
                union_sweep_chunk<V, 1>(s_ulist, n_list, s_uoff, s_W, s_wsum, s_tile, vals, t0, n_batch, stride_t, warp, lane);
            __syncthreads();
            // coalesced stores: a half-warp writes one 16-target row segment
            const int rows = two ? 64 : 32;
            for (int tr = warp * 2 + (lane >> 4); tr < rows; tr += UWARPS * 2) {
                const long long tt = t0 + tr;
                const long long m = m0 + (lane & 15);
                if (tt < n_batch && m < M) {
                    const V r = s_tile[tr * (UT + 1) + (lane & 15)];
                    for (int p = 0; p < po.n; ++p) static_cast<V*>(po.base[p])[tt * po.m_total + po.m_off + m] = r;
                }
            }
            __syncthreads();
        }
        return;
    }

    // ---- per-target sweep (non-finite values, k > 32, or a union that does not fit) ----
    for (long long t0 = 0; t0 < n_batch; t0 += 32) {
        const long long t = t0 + lane;
        const bool tin = t < n_batch;
        const char* lane_base = reinterpret_cast<const char*>(vals) + (tin ? t : 0) * stride_t * (long long)sizeof(V);
        for (int tl = warp; tl < UT; tl += UWARPS) {
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
            s_tile[lane * (UT + 1) + tl] = r;  // [t_local][target_local]
        }
        __syncthreads();
        for (int tr = warp * 2 + (lane >> 4); tr < 32; tr += UWARPS * 2) {
            const long long tt = t0 + tr;
            const long long m = m0 + (lane & 15);
            if (tt < n_batch && m < M) {
                const V r = s_tile[tr * (UT + 1) + (lane & 15)];
                for (int p = 0; p < po.n; ++p) static_cast<V*>(po.base[p])[tt * po.m_total + po.m_off + m] = r;
            }
        }
        __syncthreads();
    }
}

// Score of a field against held-out values, NaN-aware like the reference's nanmean (comparison.py:268-306): a target
// counts when prediction and truth are both finite.  Deterministic: per-CTA partials in fixed slots, one CTA adds them
// in a fixed order.  part: [n_ctas][3] = sum |d|, sum d^2, count.
constexpr int SCORE_CTAS = 1024;  // grid of the stand-alone score kernel
__device__ __forceinline__ void score_cta_partial(double ad, double sq, double cnt, double* __restrict__ part) {
    __shared__ double s_part[3][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    ad = warp_sum(ad);
    sq = warp_sum(sq);
    cnt = warp_sum(cnt);
    if (lane == 0) {
        s_part[0][warp] = ad;
        s_part[1][warp] = sq;
        s_part[2][warp] = cnt;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_part[threadIdx.x][w];
        part[(size_t)blockIdx.x * 3 + threadIdx.x] = t;
    }
}
__global__ void score_reduce_kernel(const double* __restrict__ part, long long n_ctas, double* __restrict__ sums) {
    __shared__ double s[3][256];
    double a[3] = {0.0, 0.0, 0.0};
    for (long long i = threadIdx.x; i < n_ctas; i += 256)
        for (int q = 0; q < 3; ++q) a[q] += part[i * 3 + q];
    for (int q = 0; q < 3; ++q) s[q][threadIdx.x] = a[q];
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o)
            for (int q = 0; q < 3; ++q) s[q][threadIdx.x] += s[q][threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x < 3) sums[threadIdx.x] = s[threadIdx.x][0];
}
template <typename V>
__global__ void __launch_bounds__(256) score_kernel(const V* __restrict__ pred, const V* __restrict__ truth, long long n,
                                                     double* __restrict__ part) {
    double ad = 0.0, sq = 0.0, cnt = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double d = (double)pred[i] - (double)truth[i];
        if (isfinite(d)) {
            ad += fabs(d);
            sq = fma(d, d, sq);
            cnt += 1.0;
        }
    }
    score_cta_partial(ad, sq, cnt, part);
}

// single slice from a table: one warp per target (HPO inner loop, small problems).  SCORE: the trial's error against
// held-out values is reduced in the same kernel (`out` may then be null: the field is never written).
template <typename V, bool SCORE>
__global__ void __launch_bounds__(256) idw_apply_single_kernel(const int* __restrict__ idx, const double* __restrict__ dd,
                                                                int dist_is_squared, long long M, int k_table, int k_use,
                                                                int k_min, double power, const V* __restrict__ vals,
                                                                long long n_samples, long long stride_i, V* __restrict__ out,
                                                                const V* __restrict__ truth, double* __restrict__ part) {
    const int lane = threadIdx.x & 31;
    const long long m = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const bool live = m < M;  // (warp-uniform; with SCORE every warp reaches the CTA-wide reduction below)
    if (!live && !SCORE) return;
    double w[CMX_MAX_K / 32];
    double wsum = 0.0;
#pragma unroll
    for (int e = 0; e < CMX_MAX_K / 32; ++e) {
        const int j = e * 32 + lane;
        w[e] = 0.0;
        if (live && j < k_use) {
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
    const V r = (fin < k_min || den == 0.0) ? quiet_nan<V>() : (V)(num / den);
    if (live && lane == 0 && out) out[m] = r;
    if (SCORE) {
        double ad = 0.0, sq = 0.0, cnt = 0.0;
        if (live && lane == 0) {
            const double d = (double)r - (double)truth[m];  // the field as stored (rounded to V), like the reference
            if (isfinite(d)) {
                ad = fabs(d);
                sq = d * d;
                cnt = 1.0;
            }
        }
        score_cta_partial(ad, sq, cnt, part);
    }
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
                   V* out, int* flag_scratch, cudaStream_t stream, PeerOut po = PeerOut{}) {
    if (M == 0 || n_batch == 0) return CMX_OK;
    if (po.n > 0 && n_batch == 1) return CMX_ERR_UNSUPPORTED;
    if (po.n == 0) {  // plain output
        po.base[0] = out;
        po.n = 1;
        po.m_total = M;
        po.m_off = 0;
    }
    const long long stride_t = layout == CMX_LAYOUT_SAMPLE_MAJOR ? 1 : n_samples;
    const long long stride_i = layout == CMX_LAYOUT_SAMPLE_MAJOR ? n_batch : 1;
    if (n_batch == 1) {
        const int threads = 256;
        const long long blocks = (M * 32 + threads - 1) / threads;
        idw_apply_single_kernel<V, false><<<(unsigned)blocks, threads, 0, stream>>>(
            idx, dd, dist_is_squared, M, k_table, k_use, k_min, power, vals, n_samples, stride_i, out, nullptr, nullptr);
        note_launch();
        return check_launch();
    }
    if (flag_scratch) {  // lets the batched kernel skip the per-value finiteness test
        CMX_CUDA(cudaMemsetAsync(flag_scratch, 0, sizeof(int), stream));
        const int fblocks = sm_count() * 8;
        flag_nonfinite_kernel<V><<<fblocks, 256, 0, stream>>>(vals, n_samples * n_batch, flag_scratch);
        note_launch();
    }
#ifndef CMX_K2_UNION_MIN_BATCH
#define CMX_K2_UNION_MIN_BATCH 128
#endif
    if (n_batch < CMX_K2_UNION_MIN_BATCH || k_use > UK) {  // short batches: the set-up of the shared sweep does not pay
        const size_t smem = (size_t)BW_TARGETS * k_use * sizeof(NbrEntry) + BW_TARGETS * sizeof(double) + 32 * 33 * sizeof(V);
        const long long blocks = (M + BW_TARGETS - 1) / BW_TARGETS;
        if (smem > 48 * 1024)
            CMX_CUDA(cudaFuncSetAttribute(idw_batch_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        timing_begin(CMX_KERNEL_IDW_BATCH, stream);
        idw_batch_kernel<V><<<(unsigned)blocks, BW_THREADS, smem, stream>>>(idx, dd, dist_is_squared, M, k_table, k_use,
                                                                           k_min, power, vals, n_samples, n_batch,
                                                                           stride_t, stride_i, flag_scratch, po);
    } else {
        const size_t smem = (size_t)UT * k_use * sizeof(NbrEntry) + (size_t)(UMAX + 1) * UT * sizeof(double) +
                            (UMAX + 1) * sizeof(unsigned long long) + UT * sizeof(double) +
                            (size_t)UT * UK * sizeof(int) + 2 * UHS * sizeof(int) + UWARPS * 4 * sizeof(unsigned) +
                            (UWARPS + 1 + 3) * sizeof(int) + UHS + UWARPS * (UMAX + 4) + 64;
        static_assert((size_t)UT * UK * sizeof(int) + 2 * UHS * sizeof(int) >= 64 * (UT + 1) * sizeof(double),
                      "output tile must fit in the hash storage it reuses");
        const long long blocks = (M + UT - 1) / UT;
        CMX_CUDA(cudaFuncSetAttribute(idw_batch_union_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        timing_begin(CMX_KERNEL_IDW_BATCH, stream);
        idw_batch_union_kernel<V><<<(unsigned)blocks, UTHREADS, smem, stream>>>(idx, dd, dist_is_squared, M, k_table,
                                                                               k_use, k_min, power, vals, n_samples,
                                                                               n_batch, stride_t, stride_i, flag_scratch, po);
    }
    timing_end(CMX_KERNEL_IDW_BATCH, stream);
    note_launch();
    return check_launch();
}

template <typename V>
int idw_entry(const double* s_lat, const double* s_lon, int64_t n, const V* vals, int64_t n_batch, int layout,
              const double* t_lat, int64_t n_lat, const double* t_lon, int64_t n_lon, int grid, int k, int k_min,
              double power, V* out, int32_t* idx_out, double* dist_out, void* workspace, size_t workspace_bytes,
              const cmx_options* options, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PeerOut po{};
    int search = CMX_SEARCH_AUTO;
    const int orc = read_options(options, &search, &po);
    if (orc != CMX_OK) return orc;
    if (!vals || !out || n_batch < 1 || k_min < 1) return CMX_ERR_BAD_ARG;
    if (po.n > 0 && n_batch == 1) return CMX_ERR_UNSUPPORTED;
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
    job.search = search;
    job.workspace = workspace;
    job.workspace_bytes = workspace_bytes;
    if (n_batch == 1) {
        job.idw_mode = 1;  // fused: weights + weighted mean in the kNN kernel's epilogue
        job.vals = vals;
        job.out = out;
        return run_knn(job, stream);
    }
    // batch.  Workspace: [tables | kNN scratch]
    if (k < 1 || k > CMX_MAX_K) return k < 1 ? CMX_ERR_BAD_ARG : CMX_ERR_K_TOO_LARGE;
    const size_t extra_bytes = (cmx_idw_workspace_bytes(k, M, n_batch) + 255) & ~size_t(255);
    if (!workspace || workspace_bytes < knn_ws + extra_bytes + 256) return CMX_ERR_WORKSPACE;
    unsigned char* extra = static_cast<unsigned char*>(workspace);
    extra = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(extra) + 255) & ~uintptr_t(255));
    double* d2_tab = reinterpret_cast<double*>(extra);
    int* idx_tab = reinterpret_cast<int*>(extra + (size_t)M * k * sizeof(double));
    int* flag = reinterpret_cast<int*>(extra + (size_t)M * k * (sizeof(double) + sizeof(int)));
    job.workspace = extra + extra_bytes;
    job.workspace_bytes = workspace_bytes - (size_t)((extra + extra_bytes) - static_cast<unsigned char*>(workspace));

#ifndef CMX_K2_UNION_MIN_BATCH
#define CMX_K2_UNION_MIN_BATCH 128
#endif
    // Short batches with the cell-binned search: the weighting is applied inside the search kernel (fused batch
    // epilogue, knn.cu finish_batch) - the [M][k] table is never written or read back.  Long batches amortise the table
    // and use the neighbour-sharing K2 kernel; the brute-force kernel K1 keeps the table path.  So does a sharded call
    // that stores into several ranks: the fused epilogue finishes 4 targets at a time, i.e. 32-byte segments per peer,
    // and NVLink wants K2's 128/256-byte ones (8 GPUs, C5 band: 15.7 ms fused against 1.7 + 2.5 ms through the table).
    if (search != CMX_SEARCH_BRUTE && layout == CMX_LAYOUT_SAMPLE_MAJOR && n_batch < CMX_K2_UNION_MIN_BATCH && M > 0 && po.n <= 1) {
        CMX_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), stream));
        flag_nonfinite_kernel<V><<<sm_count() * 8, 256, 0, stream>>>(vals, (long long)n * n_batch, flag);
        note_launch();
        KnnJob fused = job;
        fused.idw_mode = 2;
        fused.vals = vals;
        fused.n_batch = n_batch;
        fused.nonfinite_flag = flag;
        fused.peers = po;
        if (po.n == 0) {  // plain output
            fused.peers.base[0] = out;
            fused.peers.n = 1;
            fused.peers.m_total = M;
            fused.peers.m_off = 0;
        }
        const int rc = run_knn(fused, stream);
        if (rc != CMX_ERR_UNSUPPORTED) return rc;  // (unsupported: the binned search's scratch does not fit -> table path)
    }
    job.idw_mode = 0;
    job.d2_out = d2_tab;
    if (!job.idx_out) job.idx_out = idx_tab;
    int rc = run_knn(job, stream);
    if (rc != CMX_OK) return rc;
    return idw_from_table<V>(job.idx_out, d2_tab, 1, M, k, k, k_min, power, vals, n, n_batch, layout, out, flag, stream,
                             po);
}

template <typename V>
int apply_entry(const int32_t* idx, const double* dist, int64_t M, int k_table, const V* vals, int64_t n,
                int64_t n_batch, int layout, int k_use, int k_min, double power, V* out, int32_t* flag_scratch,
                const cmx_options* options, void* stream_) {
    PeerOut po{};
    const int orc = read_options(options, nullptr, &po);
    if (orc != CMX_OK) return orc;
    if (!idx || !dist || !vals || !out) return CMX_ERR_BAD_ARG;
    if (M < 0 || n_batch < 1 || k_use < 1 || k_use > k_table || k_min < 1) return CMX_ERR_BAD_ARG;
    if (k_use > CMX_MAX_K) return CMX_ERR_K_TOO_LARGE;
    if (layout != CMX_LAYOUT_BATCH_MAJOR && layout != CMX_LAYOUT_SAMPLE_MAJOR) return CMX_ERR_BAD_ARG;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    // without scratch the batched kernel tests every value for finiteness itself
    return idw_from_table<V>(idx, dist, 0, M, k_table, k_use, k_min, power, vals, n, n_batch, layout, out,
                             flag_scratch, stream, po);
}

inline long long idw_score_ctas(long long M) { return (M * 32 + 255) / 256; }

template <typename V>
int idw_score_entry(const int32_t* idx, const double* dist, int64_t M, int k_table, const V* vals, int64_t n, int k_use,
                    int k_min, double power, const V* truth, V* out, double* sums, void* ws, size_t ws_bytes, void* stream_) {
    if (!idx || !dist || !vals || !truth || !sums || !ws) return CMX_ERR_BAD_ARG;
    if (M < 1 || k_use < 1 || k_use > k_table || k_min < 1) return CMX_ERR_BAD_ARG;
    if (k_use > CMX_MAX_K) return CMX_ERR_K_TOO_LARGE;
    const long long ctas = idw_score_ctas(M);
    if (ws_bytes < (size_t)ctas * 3 * sizeof(double)) return CMX_ERR_WORKSPACE;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    double* part = static_cast<double*>(ws);
    idw_apply_single_kernel<V, true><<<(unsigned)ctas, 256, 0, stream>>>(idx, dist, 0, M, k_table, k_use, k_min, power, vals, n,
                                                                        1, out, truth, part);
    note_launch();
    score_reduce_kernel<<<1, 256, 0, stream>>>(part, ctas, sums);
    note_launch();
    return check_launch();
}

template <typename V>
int score_entry(const V* pred, const V* truth, int64_t n, double* sums, void* ws, size_t ws_bytes, void* stream_) {
    if (!pred || !truth || !sums || !ws || n < 1) return CMX_ERR_BAD_ARG;
    if (ws_bytes < (size_t)SCORE_CTAS * 3 * sizeof(double)) return CMX_ERR_WORKSPACE;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    double* part = static_cast<double*>(ws);
    long long ctas = (n + 255) / 256;
    if (ctas > SCORE_CTAS) ctas = SCORE_CTAS;
    score_kernel<V><<<(unsigned)ctas, 256, 0, stream>>>(pred, truth, n, part);
    note_launch();
    score_reduce_kernel<<<1, 256, 0, stream>>>(part, ctas, sums);
    note_launch();
    return check_launch();
}

}  // namespace
}  // namespace cmx

extern "C" {

size_t cmx_score_workspace_bytes(int64_t n_targets) {
    const long long a = cmx::idw_score_ctas(n_targets > 0 ? n_targets : 0), b = cmx::SCORE_CTAS;
    return (size_t)(a > b ? a : b) * 3 * sizeof(double);
}
int cmx_idw_score_f64(const int32_t* idx, const double* dist, int64_t M, int k_table, const double* vals, int64_t n, int k_use,
                      int k_min, double power, const double* truth, double* out, double* sums, void* ws, size_t ws_bytes,
                      void* stream) {
    return cmx::idw_score_entry<double>(idx, dist, M, k_table, vals, n, k_use, k_min, power, truth, out, sums, ws, ws_bytes, stream);
}
int cmx_idw_score_f32(const int32_t* idx, const double* dist, int64_t M, int k_table, const float* vals, int64_t n, int k_use,
                      int k_min, double power, const float* truth, float* out, double* sums, void* ws, size_t ws_bytes,
                      void* stream) {
    return cmx::idw_score_entry<float>(idx, dist, M, k_table, vals, n, k_use, k_min, power, truth, out, sums, ws, ws_bytes, stream);
}
int cmx_score_f64(const double* pred, const double* truth, int64_t n, double* sums, void* ws, size_t ws_bytes, void* stream) {
    return cmx::score_entry<double>(pred, truth, n, sums, ws, ws_bytes, stream);
}
int cmx_score_f32(const float* pred, const float* truth, int64_t n, double* sums, void* ws, size_t ws_bytes, void* stream) {
    return cmx::score_entry<float>(pred, truth, n, sums, ws, ws_bytes, stream);
}

size_t cmx_idw_workspace_bytes(int k, int64_t n_targets, int64_t n_batch) {
    if (n_batch <= 1 || k < 1 || n_targets < 0) return 0;
    return (size_t)n_targets * (size_t)k * (sizeof(double) + sizeof(int)) + 1024;
}

int cmx_idw_f64(const double* s_lat, const double* s_lon, int64_t n, const double* vals, int64_t n_batch, int layout,
                const double* t_lat, int64_t n_lat, const double* t_lon, int64_t n_lon, int grid, int k, int k_min,
                double power, double* out, int32_t* idx_out, double* dist_out, void* ws, size_t ws_bytes,
                const cmx_options* options, void* stream) {
    return cmx::idw_entry<double>(s_lat, s_lon, n, vals, n_batch, layout, t_lat, n_lat, t_lon, n_lon, grid, k, k_min,
                                  power, out, idx_out, dist_out, ws, ws_bytes, options, stream);
}
int cmx_idw_f32(const double* s_lat, const double* s_lon, int64_t n, const float* vals, int64_t n_batch, int layout,
                const double* t_lat, int64_t n_lat, const double* t_lon, int64_t n_lon, int grid, int k, int k_min,
                double power, float* out, int32_t* idx_out, double* dist_out, void* ws, size_t ws_bytes,
                const cmx_options* options, void* stream) {
    return cmx::idw_entry<float>(s_lat, s_lon, n, vals, n_batch, layout, t_lat, n_lat, t_lon, n_lon, grid, k, k_min,
                                 power, out, idx_out, dist_out, ws, ws_bytes, options, stream);
}
int cmx_idw_apply_f64(const int32_t* idx, const double* dist, int64_t M, int k_table, const double* vals, int64_t n,
                      int64_t n_batch, int layout, int k_use, int k_min, double power, double* out,
                      int32_t* flag_scratch, const cmx_options* options, void* stream) {
    return cmx::apply_entry<double>(idx, dist, M, k_table, vals, n, n_batch, layout, k_use, k_min, power, out,
                                    flag_scratch, options, stream);
}
int cmx_idw_apply_f32(const int32_t* idx, const double* dist, int64_t M, int k_table, const float* vals, int64_t n,
                      int64_t n_batch, int layout, int k_use, int k_min, double power, float* out,
                      int32_t* flag_scratch, const cmx_options* options, void* stream) {
    return cmx::apply_entry<float>(idx, dist, M, k_table, vals, n, n_batch, layout, k_use, k_min, power, out,
                                   flag_scratch, options, stream);
}

}  // extern "C"
