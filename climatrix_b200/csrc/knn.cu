// K1: brute-force tiled k-nearest-neighbour search with a fused IDW epilogue.
//
// Replaces scipy.spatial.cKDTree(points).query(targets, k, workers=-1)
// (reference src/climatrix/reconstruct/idw.py:155-158) and, with the epilogue,
// idw.py:163-180.
//
// Scheme (see DESIGN.md "K1"):
//   * every (target, sample) pair is evaluated on the FP32 pipe in the expanded form
//     t = |s|^2 - 2 q.s  in coordinates centred on the target tile: two packed
//     FFMA2 (fma.rn.f32x2) per two pairs, reduced with the 3-input FMNMX3 to a
//     running minimum per target over a group of 32 samples; one compare per target
//     and group against a threshold held in a register;
//   * samples are staged through shared memory by TMA bulk copies
//     (cp.async.bulk + mbarrier), re-centred and converted to fp32 per tile;
//   * a group whose minimum passes the (conservative, error-bounded) fp32 filter is
//     re-examined by the whole warp (one sample per lane, ballot-compacted pushes
//     into the target's candidate buffer); when a buffer fills, the warp re-ranks
//     it in exact, non-contracted float64 (bit-identical to cKDTree's dx*dx+dy*dy)
//     with a bitonic sort/merge over warp shuffles and tightens the threshold.
//     Ties break by sample index, so results are deterministic;
//   * thresholds start from a rigorous upper bound on the k-th neighbour distance
//     read off a summed-area table of sample counts (built by four tiny pre-pass
//     kernels), so a target sees ~2-3k candidates instead of ~k ln(N/k);
//   * the final re-rank leaves the k neighbours in registers, where the IDW
//     weights and the weighted mean are evaluated before anything is stored.
#include <limits.h>
#include <stdlib.h>

#include "common.cuh"
#include "rerank.cuh"

namespace cmx {
namespace {

constexpr int WARPS = 4;
constexpr int THREADS = WARPS * 32;
#ifndef CMX_SS
#define CMX_SS 256
#endif
#ifndef CMX_CTAS
#define CMX_CTAS 5
#endif
#ifndef CMX_UNROLL
#define CMX_UNROLL 2
#endif
#define CMX_STR2(x) #x
#define CMX_STR(x) CMX_STR2(x)
#define CMX_UNROLL_PRAGMA CMX_STR(unroll CMX_UNROLL)
constexpr int SS = CMX_SS;           // samples per (warp-private) shared-memory stage
constexpr int GRP = 32;              // samples per filter group (= one per lane in the re-examination)
#ifndef CMX_CAP
#define CMX_CAP 128
#endif
constexpr int CAP = CMX_CAP;         // candidate slots per target
constexpr int FLUSH_AT = CAP - GRP;  // re-rank a target once it holds more than this
constexpr int MAX_CTAS_PER_SM = CMX_CTAS;
// resident CTAs per SM: the 16-targets-per-thread variant (large N) needs 128 registers
constexpr int ctas_per_sm(int qpt) { return qpt >= 16 ? 4 : MAX_CTAS_PER_SM; }
// target slots of scratch per SM, enough for either variant
constexpr int SLOTS_PER_SM = (ctas_per_sm(16) * WARPS * 16 * 32 > MAX_CTAS_PER_SM * WARPS * 8 * 32)
                                 ? ctas_per_sm(16) * WARPS * 16 * 32
                                 : MAX_CTAS_PER_SM * WARPS * 8 * 32;
constexpr size_t WARP_SMEM = 2 * 2 * SS * sizeof(double) + 3 * SS * sizeof(float) + 64;  // raw x2, comp, mbarriers
#ifndef CMX_GMAX
#define CMX_GMAX 2048
#endif
#ifndef CMX_SEED_C0
#define CMX_SEED_C0 0.5
#endif
constexpr int GMAX = CMX_GMAX;       // seed grid: at most GMAX x GMAX cells
constexpr int SAT_LD = GMAX + 1;     // row stride of the summed-area table

#define CMX_INF __longlong_as_double(0x7ff0000000000000LL)
#define CMX_INF_F __int_as_float(0x7f800000)

// geometry of the sample-count grid (device resident, written by seed_geometry_kernel)
struct SeedGrid {
    double a0, b0, h, inv_h;
    int ga, gb;
    int valid;
    int pad;
};

struct KnnArgs {
    KnnJob j;
    long long M;
    int tiles_c;    // grid mode: tiles per row of tiles
    int num_tiles;
    int aligned;    // sample arrays 16-byte aligned -> TMA bulk staging
    unsigned* cand; // [ctas][TQ][CAP] candidate sample indices
    double* list_d; // [ctas][TQ][KP]  running top-KP squared distances (ascending)
    int* list_i;    // [ctas][TQ][KP]  and their sample indices
    const SeedGrid* seed;
    const int* sat; // (GMAX+1)^2 summed-area table of sample counts (exclusive prefix)
    int* tile_counter;  // dynamic tile scheduler (zeroed by seed_init_kernel)
    // sample split: a tile is swept by nseg work units, each over 1/nseg of the samples; unit results
    // (sorted partial lists) go to part_d/part_i [M][nseg][KP] and knn_merge_kernel finishes the job
    int nseg;
    double* part_d;
    int* part_i;
    // cell-binned search (K1c): samples counting-sorted by seed-grid cell (row-major cell order)
    int* bin_cursor;     // [ga * gb] fill cursors of the scatter
    double* bin_a;       // [N] sorted copies of the sample coordinates
    double* bin_b;
    int* bin_idx;        // [N] original sample index of each sorted slot
    int chunk;
    int run;             // consecutive units a warp takes before it strides on (binned search: chained bound vs L1 reuse)           // K1c: consecutive work units (targets / four-target groups) a CTA takes at a time
};

// ------------------------------------------------------------------------------------------
// seed grid pre-pass
// ------------------------------------------------------------------------------------------
__global__ void seed_init_kernel(double* bbox, int* tile_counter) {
    *tile_counter = 0;
    bbox[0] = CMX_INF;
    bbox[1] = -CMX_INF;
    bbox[2] = CMX_INF;
    bbox[3] = -CMX_INF;
}

__global__ void seed_bbox_kernel(const double* __restrict__ sa, const double* __restrict__ sb, long long n,
                                 double* bbox) {
    double lo_a = CMX_INF, hi_a = -CMX_INF, lo_b = CMX_INF, hi_b = -CMX_INF;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double a = sa[i], b = sb[i];
        if (isfinite(a) && isfinite(b)) {
            lo_a = fmin(lo_a, a);
            hi_a = fmax(hi_a, a);
            lo_b = fmin(lo_b, b);
            hi_b = fmax(hi_b, b);
        }
    }
    lo_a = warp_min(lo_a);
    hi_a = warp_max(hi_a);
    lo_b = warp_min(lo_b);
    hi_b = warp_max(hi_b);
    if ((threadIdx.x & 31) == 0 && lo_a <= hi_a) {
        atomic_min_double(&bbox[0], lo_a);
        atomic_max_double(&bbox[1], hi_a);
        atomic_min_double(&bbox[2], lo_b);
        atomic_max_double(&bbox[3], hi_b);
    }
}

__device__ __forceinline__ SeedGrid make_seed_grid(const double* bbox, long long n) {
    const double ea = bbox[1] - bbox[0], eb = bbox[3] - bbox[2];
    SeedGrid s;
    s.a0 = bbox[0];
    s.b0 = bbox[2];
    s.valid = (ea >= 0.0 && eb >= 0.0 && isfinite(ea) && isfinite(eb)) ? 1 : 0;
    // about CMX_SEED_C0 samples per cell, at most GMAX cells per axis
    double h = 0.0;
    if (s.valid) {
        const double area = (ea > 0 ? ea : 1.0) * (eb > 0 ? eb : 1.0);
        h = sqrt(CMX_SEED_C0 * area / (double)(n > 0 ? n : 1));
        h = fmax(h, fmax(ea, eb) / (double)(GMAX - 1));
        if (!(h > 0.0)) h = 1.0;
    } else {
        h = 1.0;
    }
    s.h = h;
    s.inv_h = 1.0 / h;
    int ga = (int)(ea * s.inv_h) + 1, gb = (int)(eb * s.inv_h) + 1;
    s.ga = ga < 1 ? 1 : (ga > GMAX ? GMAX : ga);
    s.gb = gb < 1 ? 1 : (gb > GMAX ? GMAX : gb);
    s.pad = 0;
    return s;
}

__global__ void seed_geometry_kernel(const double* bbox, long long n, SeedGrid* g) { *g = make_seed_grid(bbox, n); }

// zero the part of the summed-area table the grid uses ((ga + 1) x (gb + 1) entries of the 2049-wide rows) and,
// for the binned search, the ga x gb scatter cursors - instead of a 16.8 MB memset regardless of N
__global__ void seed_clear_kernel(const SeedGrid* g, int* sat, int* cursor) {
    const int ga = g->ga, gb = g->gb;
    const long long cells = (long long)(ga + 1) * (gb + 1);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(i / (gb + 1)), col = (int)(i - (long long)row * (gb + 1));
        sat[(size_t)row * SAT_LD + col] = 0;
        if (cursor && i < (long long)ga * gb) cursor[i] = 0;
    }
}

__device__ __forceinline__ int seed_cell(double x, double x0, double inv_h, int g) {
    int c = (int)floor((x - x0) * inv_h);
    return c < 0 ? 0 : (c > g - 1 ? g - 1 : c);
}

__global__ void seed_hist_kernel(const double* __restrict__ sa, const double* __restrict__ sb, long long n,
                                 const SeedGrid* g, int* sat) {
    const SeedGrid s = *g;
    if (!s.valid) return;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double a = sa[i], b = sb[i];
        if (isfinite(a) && isfinite(b)) {
            const int ia = seed_cell(a, s.a0, s.inv_h, s.ga), ib = seed_cell(b, s.b0, s.inv_h, s.gb);
            atomicAdd(&sat[(size_t)(ia + 1) * SAT_LD + (ib + 1)], 1);
        }
    }
}

// inclusive scan along b of row r = blockIdx.x + 1 (one warp per row)
__global__ void seed_rowscan_kernel(const SeedGrid* g, int* sat) {
    const int ga = g->ga, gb = g->gb;
    const int r = blockIdx.x + 1;
    if (r > ga) return;
    int* row = sat + (size_t)r * SAT_LD;
    const int lane = threadIdx.x;
    int carry = 0;
    for (int c0 = 1; c0 <= gb; c0 += 32) {
        const int c = c0 + lane;
        int v = c <= gb ? row[c] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(kFull, v, o);
            if (lane >= o) v += u;
        }
        v += carry;
        if (c <= gb) row[c] = v;
        carry = __shfl_sync(kFull, v, 31);
    }
}

// inclusive scan along a of every column (thread per column)
__global__ void seed_colscan_kernel(const SeedGrid* g, int* sat) {
    const int ga = g->ga, gb = g->gb;
    const int c = blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (c > gb) return;
    int acc = 0;
    for (int r = 1; r <= ga; ++r) {
        acc += sat[(size_t)r * SAT_LD + c];
        sat[(size_t)r * SAT_LD + c] = acc;
    }
}

// Rigorous upper bound on the squared distance from (ta, tb) to its k-th nearest sample:
// grow a box of cells around the target's cell until it holds >= k samples; every one of
// them is no farther than the farthest corner of that box.
// samples in the clipped cell box [la, ha] x [lb, hb]
__device__ __forceinline__ int seed_box_count(const SeedGrid& s, const int* __restrict__ sat, int la, int ha, int lb,
                                              int hb) {
    la = max(la, 0);
    lb = max(lb, 0);
    ha = min(ha, s.ga - 1);
    hb = min(hb, s.gb - 1);
    if (ha < la || hb < lb) return 0;
    return sat[(size_t)(ha + 1) * SAT_LD + (hb + 1)] - sat[(size_t)la * SAT_LD + (hb + 1)] -
           sat[(size_t)(ha + 1) * SAT_LD + lb] + sat[(size_t)la * SAT_LD + lb];
}

// farthest squared distance from (ta, tb) to any point of the cell box [la, ha] x [lb, hb] (clipped), inflated
__device__ __forceinline__ double seed_box_far2(const SeedGrid& s, double ta, double tb, int la, int ha, int lb, int hb) {
    la = max(la, 0);
    lb = max(lb, 0);
    ha = min(ha, s.ga - 1);
    hb = min(hb, s.gb - 1);
    const double A_lo = s.a0 + la * s.h, A_hi = s.a0 + (ha + 1) * s.h;
    const double B_lo = s.b0 + lb * s.h, B_hi = s.b0 + (hb + 1) * s.h;
    const double slack = 1e-9 * (s.h + fabs(A_lo) + fabs(A_hi) + fabs(B_lo) + fabs(B_hi));
    const double da = fmax(fabs(ta - A_lo), fabs(A_hi - ta)) + slack;
    const double db = fmax(fabs(tb - B_lo), fabs(B_hi - tb)) + slack;
    return (da * da + db * db) * (1.0 + 1e-9);
}

// Rigorous upper bound on the squared distance from (ta, tb) to its k-th nearest sample: grow a region
// of cells around the target's cell until the table says it holds >= k samples; every one of them is no
// farther than the farthest point of that region.  Two region shapes are tried and the tighter bound
// wins: a square box (area / enclosing-disc ratio 2/pi) and a plus-shaped union of two rectangles
// (ratio ~0.79), which hugs the disc better and typically cuts the candidates per target by ~20 %.
__device__ __forceinline__ double seed_bound_d2(const SeedGrid& s, const int* __restrict__ sat, double ta,
                                                double tb, int k) {
    if (!s.valid) return CMX_INF;
    const int ia = seed_cell(ta, s.a0, s.inv_h, s.ga), ib = seed_cell(tb, s.b0, s.inv_h, s.gb);
    const int wmax = max(s.ga, s.gb);
    int w = 0;
    for (;; w = w < 8 ? w + 1 : w * 2) {
        if (seed_box_count(s, sat, ia - w, ia + w, ib - w, ib + w) >= k) break;
        if (w >= wmax) return CMX_INF;  // fewer than k finite samples in total
    }
    double best = seed_box_far2(s, ta, tb, ia - w, ia + w, ib - w, ib + w);
    if (w >= 2 && w <= 64) {
        // plus shape: [ia-A, ia+A] x [ib-B, ib+B]  U  [ia-B, ia+B] x [ib-A, ib+A],  B ~ 0.6 A
        for (int A = w; A <= w + (w >> 1) + 1; ++A) {
            const int B = (3 * A + 2) / 5;
            const int cnt = seed_box_count(s, sat, ia - A, ia + A, ib - B, ib + B) +
                            seed_box_count(s, sat, ia - B, ia + B, ib - A, ib + A) -
                            seed_box_count(s, sat, ia - B, ia + B, ib - B, ib + B);
            if (cnt >= k) {
                // both rectangles are contained in the disc around the target through their far corners
                const double f1 = seed_box_far2(s, ta, tb, ia - A, ia + A, ib - B, ib + B);
                const double f2 = seed_box_far2(s, ta, tb, ia - B, ia + B, ib - A, ib + A);
                best = fmin(best, fmax(f1, f2));
                break;
            }
        }
    }
    return best;
}

// ------------------------------------------------------------------------------------------
// exact re-rank machinery
// ------------------------------------------------------------------------------------------
// Conservative fp32 filter threshold for "exact d2 <= d2k" in t-space (t = d2 - |q|^2).
// Error model (DESIGN.md "filter margin"): coordinates are rounded to fp32 after
// centring (relative 2^-24 each), S = fma(b,b,a*a), t = fma(Qa,a,fma(Qb,b,S)).
__device__ __forceinline__ float filter_threshold(double d2k, float qa, float qb) {
    if (!(d2k < CMX_INF)) return CMX_INF_F;
    const double eps = 5.9604644775390625e-8;  // 2^-24
    const double rq = fabs((double)qa) + fabs((double)qb);
    const double q2 = (double)qa * (double)qa + (double)qb * (double)qb;
    const double dk = sqrt(d2k) * (1.0 + 1e-12);
    const double span = 1.4143 * dk + 2.0 * rq;
    const double delta = 1.01 * eps * span;
    const double thr = (dk + delta) * (dk + delta) - q2 + 4.04 * eps * span * span;
    float f = __double2float_ru(thr);
    f += fabsf(f) * 2.4e-7f + 1e-37f;  // strictly above, so the test can be '<'
    return f;
}

// Tiles are per warp: QPT rows x 32 columns of a dense grid, or QPT*32 consecutive explicit points.
// ts = q*32 + lane.
template <int QPT>
__device__ __forceinline__ bool target_coords(const KnnArgs& a, int tile, int ts, double& ta, double& tb,
                                              long long& m) {
    constexpr int TQW = QPT * 32;
    if (a.j.grid) {
        const int tr = tile / a.tiles_c;
        const int tc = tile - tr * a.tiles_c;
        const long long row = (long long)tr * QPT + (ts >> 5);
        const long long col = (long long)tc * 32 + (ts & 31);
        if (row >= a.j.n_a || col >= a.j.n_b) return false;
        ta = a.j.t_a[row];
        tb = a.j.t_b[col];
        m = row * a.j.n_b + col;
    } else {
        const long long p = (long long)tile * TQW + ts;
        if (p >= a.M) return false;
        ta = a.j.t_a[p];
        tb = a.j.t_b[p];
        m = p;
    }
    return true;
}

// The k neighbours of target m sit sorted in registers (element e*32+lane): store the neighbour
// table and / or evaluate the fused IDW epilogue (idw.py:163-180).  Returns the IDW value.
template <int E>
__device__ __forceinline__ double finish_target(const KnnArgs& a, int lane, long long m, const double (&ld)[E],
                                                const int (&li)[E]) {
    const int k = a.j.k;
    double wsum = 0.0;
    double w[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int i = e * 32 + lane;
        const bool in = i < k;
        if (in) {
            if (a.j.idx_out) a.j.idx_out[m * k + i] = li[e];
            if (a.j.dist_out) a.j.dist_out[m * k + i] = sqrt(ld[e]);
            if (a.j.d2_out) a.j.d2_out[m * k + i] = ld[e];
        }
        // a slot still holding the (inf, INT_MAX) sentinel means fewer than k rankable samples were found
        // (NaN coordinates): it carries no weight and its index is never dereferenced
        const bool real = in && li[e] != INT_MAX;
        w[e] = (real && a.j.idw_mode == 1) ? idw_raw_weight(ld[e], a.j.power) : 0.0;
        wsum += w[e];
    }
    if (a.j.idw_mode != 1) return 0.0;  // table only (mode 0), or the batch epilogue follows (mode 2)
    wsum = warp_sum(wsum);  // weights /= nansum(weights)            idw.py:165
    double num = 0.0, den = 0.0;
    int fin = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int i = e * 32 + lane;
        if (i < k && li[e] != INT_MAX) {
            const double wn = w[e] / wsum;
            const double v = a.j.value_is_f32 ? (double)static_cast<const float*>(a.j.vals)[li[e]]
                                              : static_cast<const double*>(a.j.vals)[li[e]];
            if (isfinite(v)) {  // weights[~isfinite] = 0; nansum(v*w)   idw.py:167-173
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
    if (fin < a.j.k_min || den == 0.0) return quiet_nan<double>();  // idw.py:179-180
    return num / den;
}

// ------------------------------------------------------------------------------------------
// Fused batch epilogue (idw_mode 2): the weighting of idw.py:163-180 applied to a short time / level batch right where
// the neighbour list is final, so the [M][k] table is neither written nor read back (K2 does that for long batches).
// The k (normalised weight, value-row offset) pairs of the target go to the warp's shared-memory strip; then the lanes
// run along the batch axis: one 16-byte broadcast load per neighbour feeds P batch elements per lane, value rows are
// read as contiguous segments (sample-major values).  Results stay in registers: r[p] is batch element p * 32 + lane.
// ------------------------------------------------------------------------------------------
struct __align__(16) BatchNbr {
    double w;
    unsigned long long off;  // byte offset of the neighbour's value row
};

template <int E, int P>
__device__ __forceinline__ void finish_batch(const KnnArgs& a, int lane, const double (&ld)[E], const int (&li)[E],
                                             BatchNbr* __restrict__ strip, double (&r)[P]) {
    const int k = a.j.k;
    const long long T = a.j.n_batch;
    const size_t vsize = a.j.value_is_f32 ? 4 : 8;
    double w[E], wsum = 0.0;
    int real = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int i = e * 32 + lane;
        const bool ok = i < k && li[e] != INT_MAX;  // (inf, INT_MAX): fewer than k rankable samples
        w[e] = ok ? idw_raw_weight(ld[e], a.j.power) : 0.0;
        wsum += w[e];
        real += ok;
    }
    wsum = warp_sum(wsum);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) real += __shfl_xor_sync(kFull, real, o);
    double den = 0.0;
    __syncwarp();  // the strip may still be read by the previous target's sweep
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int i = e * 32 + lane;
        const double wn = w[e] / wsum;  // weights /= nansum(weights)   idw.py:165
        if (i < k) {
            BatchNbr nb;
            nb.w = wn;
            nb.off = (unsigned long long)(li[e] != INT_MAX ? li[e] : 0) * (unsigned long long)T * vsize;
            strip[i] = nb;
        }
        den += wn;
    }
    den = warp_sum(den);
    __syncwarp();
    const bool check = a.j.nonfinite_flag ? *a.j.nonfinite_flag != 0 : true;
    const char* base[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const long long t = p * 32 + lane;
        base[p] = static_cast<const char*>(a.j.vals) + (t < T ? t : 0) * (long long)vsize;
    }
    if (!check) {
        double acc[P];
#pragma unroll
        for (int p = 0; p < P; ++p) acc[p] = 0.0;
        if (a.j.value_is_f32) {
#pragma unroll 8
            for (int j = 0; j < k; ++j) {
                const BatchNbr nb = strip[j];
#pragma unroll
                for (int p = 0; p < P; ++p) acc[p] = fma(nb.w, (double)*reinterpret_cast<const float*>(base[p] + nb.off), acc[p]);
            }
        } else {
#pragma unroll 8
            for (int j = 0; j < k; ++j) {
                const BatchNbr nb = strip[j];
#pragma unroll
                for (int p = 0; p < P; ++p) acc[p] = fma(nb.w, *reinterpret_cast<const double*>(base[p] + nb.off), acc[p]);
            }
        }
        // fewer than k_min existing neighbours -> NaN, like the single-slice epilogue (idw.py:179-180)
        const double dn = real < a.j.k_min ? quiet_nan<double>() : den;
#pragma unroll
        for (int p = 0; p < P; ++p) r[p] = acc[p] / dn;
    } else {  // some value is NaN / inf: weights[~isfinite] = 0; nansum; k_min counts finite values   idw.py:167-180
        double num[P], dsum[P];
        int fin[P];
#pragma unroll
        for (int p = 0; p < P; ++p) num[p] = dsum[p] = 0.0, fin[p] = 0;
        for (int j = 0; j < k; ++j) {
            const BatchNbr nb = strip[j];
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const double v = a.j.value_is_f32 ? (double)*reinterpret_cast<const float*>(base[p] + nb.off)
                                                  : *reinterpret_cast<const double*>(base[p] + nb.off);
                if (isfinite(v) && nb.w > 0.0) {
                    num[p] = fma(nb.w, v, num[p]);
                    dsum[p] += nb.w;
                    ++fin[p];
                }
            }
        }
#pragma unroll
        for (int p = 0; p < P; ++p) r[p] = (fin[p] < a.j.k_min || dsum[p] == 0.0) ? quiet_nan<double>() : num[p] / dsum[p];
    }
}

// store G consecutive targets m0 .. m0 + G - 1 of every batch element this lane owns: one row segment per element
template <int P, int G>
__device__ __forceinline__ void store_batch(const KnnArgs& a, int lane, long long m0, int count, const double (&r)[G][P]) {
    const PeerOut& po = a.j.peers;
    const long long T = a.j.n_batch;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const long long t = p * 32 + lane;
        if (t >= T) continue;
        const long long at = t * po.m_total + po.m_off + m0;
        for (int q = 0; q < po.n; ++q) {
            if (a.j.value_is_f32) {
                float* o = static_cast<float*>(po.base[q]) + at;
                if (count == G && G == 4 && (reinterpret_cast<uintptr_t>(o) & 15u) == 0) {
                    *reinterpret_cast<float4*>(o) = make_float4((float)r[0][p], (float)r[1][p], (float)r[2][p], (float)r[3][p]);
                } else {
#pragma unroll
                    for (int g = 0; g < G; ++g)
                        if (g < count) o[g] = (float)r[g][p];
                }
            } else {
                double* o = static_cast<double*>(po.base[q]) + at;
                if (count == G && G == 4 && (reinterpret_cast<uintptr_t>(o) & 15u) == 0) {
                    *reinterpret_cast<double2*>(o) = make_double2(r[0][p], r[1][p]);
                    *reinterpret_cast<double2*>(o + 2) = make_double2(r[2][p], r[3][p]);
                } else {
#pragma unroll
                    for (int g = 0; g < G; ++g)
                        if (g < count) o[g] = r[g][p];
                }
            }
        }
    }
}

// Warp-cooperative exact re-rank of one target's candidate buffer.
// Returns the new filter threshold (mid-sweep) or the fused IDW value (final).
template <int E>
__device__ __noinline__ double rerank_target(const KnnArgs& a, int lane, long long slot, double ta, double tb,
                                             long long m, float qa, float qb, int cnt, bool has_list,
                                             bool final, int seg) {
    constexpr int KP = 32 * E;
    double ld[E];
    int li[E];
    double* Ld = a.list_d + slot * KP;
    int* Li = a.list_i + slot * KP;
    if (has_list) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            ld[e] = Ld[e * 32 + lane];
            li[e] = Li[e * 32 + lane];
        }
    } else {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            ld[e] = CMX_INF;
            li[e] = INT_MAX;
        }
    }
    const unsigned* C = a.cand + slot * CAP;
    for (int b = 0; b < cnt; b += 32) {
        double cd = CMX_INF;
        int ci = INT_MAX;
        if (b + lane < cnt) {
            const unsigned s = C[b + lane];
            // bit-identical to cKDTree: plain float64 dx*dx + dy*dy, no FMA contraction
            const double da = __dsub_rn(ta, a.j.s_a[s]);
            const double db = __dsub_rn(tb, a.j.s_b[s]);
            cd = __dadd_rn(__dmul_rn(da, da), __dmul_rn(db, db));
            ci = (int)s;
            if (!(cd == cd)) cd = CMX_INF;  // NaN coordinates never rank
        }
        warp_sort32(cd, ci, lane);
        merge_sorted32<E>(ld, li, cd, ci, lane);
    }
    const int k = a.j.k;
    if (!final) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            Ld[e * 32 + lane] = ld[e];
            Li[e * 32 + lane] = li[e];
        }
        double sel = ld[0];
#pragma unroll
        for (int e = 1; e < E; ++e)
            if (((k - 1) >> 5) == e) sel = ld[e];
        const double d2k = __shfl_sync(kFull, sel, (k - 1) & 31);
        return (double)filter_threshold(d2k, qa, qb);
    }

    if (a.nseg > 1) {  // sample split: hand the sorted partial list to the merge kernel
        double* Pd = a.part_d + (m * a.nseg + seg) * KP;
        int* Pi = a.part_i + (m * a.nseg + seg) * KP;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            Pd[e * 32 + lane] = ld[e];
            Pi[e * 32 + lane] = li[e];
        }
        return 0.0;
    }
    return finish_target<E>(a, lane, m, ld, li);
}

// two IEEE FMAs in one instruction (FFMA2): r = q * (x0, x1) + (c0, c1), q broadcast
__device__ __forceinline__ void fma2_bcast(float q, float x0, float x1, float c0, float c1, float& r0, float& r1) {
    asm("{\n"
        ".reg .b64 qa, xx, cc, rr;\n"
        "mov.b64 qa, {%2, %2};\n"
        "mov.b64 xx, {%3, %4};\n"
        "mov.b64 cc, {%5, %6};\n"
        "fma.rn.f32x2 rr, qa, xx, cc;\n"
        "mov.b64 {%0, %1}, rr;\n"
        "}"
        : "=f"(r0), "=f"(r1)
        : "f"(q), "f"(x0), "f"(x1), "f"(c0), "f"(c1));
}
__device__ __forceinline__ float min3f(float a, float b, float c) {
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));  // FMNMX3; NaN operands are ignored
    return d;
}

// COLSHARE: dense-grid targets -- the QPT targets of a thread sit in one grid column, so the
// column term  u = |s|^2 - 2 q_b s_b  is evaluated once per sample and thread and each pair
// costs a single FMA (t = u - 2 q_a s_a).  Explicit target points use the generic two-FMA form.
//
// Every warp is an independent worker with a private TMA pipeline: it owns a tile of QPT x 32
// targets, streams all samples through its own double-buffered shared-memory stage (bulk copies
// issued by lane 0, completion on the warp's two mbarriers), re-centres them on its tile and
// filters.  There is no CTA-wide barrier anywhere in the kernel.
template <int E, int QPT, bool COLSHARE>
__global__ void __launch_bounds__(THREADS, ctas_per_sm(QPT)) knn_kernel(const __grid_constant__ KnnArgs a) {
    constexpr int TQW = QPT * 32;  // targets per warp tile
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;

    // per-warp shared memory: raw[2 buffers][a|b][SS] f64, comp[SS/4][a0..3 b0..3 s0..3] f32, 2 mbarriers
    unsigned char* wbase = smem_raw + (size_t)warp * WARP_SMEM;
    double* raw = reinterpret_cast<double*>(wbase);                       // [2][2][SS]
    float* comp = reinterpret_cast<float*>(wbase + 2 * 2 * SS * sizeof(double));
    uint64_t* mbar = reinterpret_cast<uint64_t*>(wbase + 2 * 2 * SS * sizeof(double) + 3 * SS * sizeof(float));
    if (lane == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        mbar_fence_init();
    }
    __syncwarp();
    uint32_t ph0 = 0, ph1 = 0;

    const long long n = a.j.n;
    const int n_st = (int)((n + SS - 1) / SS);
    const int gwarp = blockIdx.x * WARPS + warp;
    const long long slot0 = (long long)gwarp * TQW;  // + q*32 + lane
    const SeedGrid seed = *a.seed;
    unsigned* const cand_w = a.cand + slot0 * CAP;  // this warp's candidate buffers

    // warps fetch tiles from a global counter: early finishers take the remainder, which evens out the
    // load when the tile count is not a multiple of the resident warps (multi-GPU bands)
    for (;;) {
        int unit = 0;
        if (lane == 0) unit = atomicAdd(a.tile_counter, 1);
        unit = __shfl_sync(kFull, unit, 0);
        if (unit >= a.num_tiles * a.nseg) break;
        const int seg = unit / a.num_tiles;  // segment-major: the units of one segment share L2-hot samples
        const int tile = unit - seg * a.num_tiles;
        const int st_per_seg = (n_st + a.nseg - 1) / a.nseg;
        const int st_begin = seg * st_per_seg;
        const int st_end = min(n_st, st_begin + st_per_seg);
        // ---------------- targets of this thread ----------------
        float Qa[QPT], Qb[COLSHARE ? 1 : QPT], thr[QPT];
        unsigned offp[(QPT + 3) / 4];  // candidate counts, 8 bits per target (<= CAP <= 128)
#pragma unroll
        for (int i = 0; i < (QPT + 3) / 4; ++i) offp[i] = 0u;
        unsigned validm = 0, has_list = 0;
        double ca, cb;
        {
            double ta[QPT], tb[QPT];
            double mn_a = CMX_INF, mx_a = -CMX_INF, mn_b = CMX_INF, mx_b = -CMX_INF;
#pragma unroll
            for (int q = 0; q < QPT; ++q) {
                long long m;
                ta[q] = 0.0;
                tb[q] = q ? tb[0] : 0.0;  // COLSHARE: rows beyond the grid inherit the column coordinate
                if (target_coords<QPT>(a, tile, q * 32 + lane, ta[q], tb[q], m)) {
                    validm |= 1u << q;
                    mn_a = fmin(mn_a, ta[q]);
                    mx_a = fmax(mx_a, ta[q]);
                    mn_b = fmin(mn_b, tb[q]);
                    mx_b = fmax(mx_b, tb[q]);
                }
            }
            mn_a = warp_min(mn_a);
            mx_a = warp_max(mx_a);
            mn_b = warp_min(mn_b);
            mx_b = warp_max(mx_b);
            ca = 0.5 * (mn_a + mx_a);  // tile centre: keeps |q|, |s| small where it matters
            cb = 0.5 * (mn_b + mx_b);
#pragma unroll
            for (int q = 0; q < QPT; ++q) {
                const bool ok = (validm >> q) & 1u;
                const float qa = __double2float_rn(__dsub_rn(ta[q], ca));
                const float qb = __double2float_rn(__dsub_rn(tb[q], cb));
                Qa[q] = ok ? -2.0f * qa : 0.0f;
                if (!COLSHARE) Qb[q] = ok ? -2.0f * qb : 0.0f;
                else if (q == 0) Qb[0] = -2.0f * qb;  // one column per thread
                // rigorous starting threshold from the sample-count grid (never below the true k-th distance)
                thr[q] = ok ? filter_threshold(seed_bound_d2(seed, a.sat, ta[q], tb[q], a.j.k), qa, qb) : -CMX_INF_F;
            }
        }

        // exact re-rank of target (lane L, slot q) of this warp; returns threshold or fused value
        auto rerank = [&](int q, int L, int cnt, bool has, bool final) -> double {
            double ta = 0.0, tb = 0.0;
            long long m = 0;
            target_coords<QPT>(a, tile, q * 32 + L, ta, tb, m);
            const float qa = __double2float_rn(__dsub_rn(ta, ca));
            const float qb = __double2float_rn(__dsub_rn(tb, cb));
            return rerank_target<E>(a, lane, slot0 + q * 32 + L, ta, tb, m, qa, qb, cnt, has, final, seg);
        };

        auto stage_is_bulk = [&](int st) -> bool {
            return a.aligned && (n - (long long)st * SS) >= (long long)SS;
        };
        auto issue = [&](int st) {  // lane 0: TMA bulk copies of stage st into raw buffer st & 1
            const int b = st & 1;
            double* ra = raw + b * 2 * SS;
            mbar_arrive_expect_tx(&mbar[b], 2u * SS * 8u);
            tma_load_1d(ra, a.j.s_a + (long long)st * SS, SS * 8u, &mbar[b]);
            tma_load_1d(ra + SS, a.j.s_b + (long long)st * SS, SS * 8u, &mbar[b]);
        };

        // ---------------- sweep over all samples ----------------
        if (lane == 0) {
            if (st_begin < st_end && stage_is_bulk(st_begin)) issue(st_begin);
            if (st_begin + 1 < st_end && stage_is_bulk(st_begin + 1)) issue(st_begin + 1);
        }
        for (int st = st_begin; st < st_end; ++st) {
            const long long base = (long long)st * SS;
            const int cntS = (int)((n - base) < (long long)SS ? (n - base) : (long long)SS);
            const bool bulk = stage_is_bulk(st);
            const int b = st & 1;
            const double* ra = raw + b * 2 * SS;
            if (bulk) {
                if (b == 0) {
                    mbar_wait(&mbar[0], ph0);
                    ph0 ^= 1u;
                } else {
                    mbar_wait(&mbar[1], ph1);
                    ph1 ^= 1u;
                }
            }
            // re-centre on the tile, round to fp32, precompute |s|^2; layout [i/4][a0..3 b0..3 s0..3]
#pragma unroll
            for (int i = lane; i < SS; i += 32) {
                double xa, xb;
                if (bulk) {
                    xa = ra[i];
                    xb = ra[SS + i];
                } else if (i < cntS) {
                    xa = a.j.s_a[base + i];
                    xb = a.j.s_b[base + i];
                } else {
                    xa = xb = quiet_nan<double>();  // padding never passes the filter
                }
                const float fa = __double2float_rn(__dsub_rn(xa, ca));
                const float fb = __double2float_rn(__dsub_rn(xb, cb));
                float* dst = comp + (i >> 2) * 12 + (i & 3);
                dst[0] = fa;
                dst[4] = fb;
                dst[8] = fmaf(fb, fb, fa * fa);
            }
            __syncwarp();  // stage complete; raw buffer b is free again
            if (lane == 0 && st + 2 < st_end && stage_is_bulk(st + 2)) issue(st + 2);

            const int cnt_pad = (cntS + GRP - 1) & ~(GRP - 1);
            const unsigned jbase = (unsigned)base;
#pragma unroll 1
            for (int g0 = 0; g0 < cnt_pad; g0 += GRP) {
                // ---- filter: running minimum of t per target over the group (FP32 pipe) ----
                float mn[QPT];
#pragma unroll
                for (int q = 0; q < QPT; ++q) mn[q] = CMX_INF_F;
                const float4* cg = reinterpret_cast<const float4*>(comp + (g0 >> 2) * 12);
                _Pragma(CMX_UNROLL_PRAGMA)
                for (int j = 0; j < GRP / 4; ++j) {
                    const float4 A4 = cg[3 * j + 0];
                    const float4 B4 = cg[3 * j + 1];
                    const float4 S4 = cg[3 * j + 2];
                    if (COLSHARE) {
                        float u0, u1, u2, u3;
                        fma2_bcast(Qb[0], B4.x, B4.y, S4.x, S4.y, u0, u1);
                        fma2_bcast(Qb[0], B4.z, B4.w, S4.z, S4.w, u2, u3);
#pragma unroll
                        for (int q = 0; q < QPT; ++q) {
                            float t0, t1, t2, t3;
                            fma2_bcast(Qa[q], A4.x, A4.y, u0, u1, t0, t1);
                            fma2_bcast(Qa[q], A4.z, A4.w, u2, u3, t2, t3);
                            mn[q] = min3f(min3f(mn[q], t0, t1), t2, t3);
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < QPT; ++q) {
                            float u0, u1, u2, u3, t0, t1, t2, t3;
                            fma2_bcast(Qb[q], B4.x, B4.y, S4.x, S4.y, u0, u1);
                            fma2_bcast(Qb[q], B4.z, B4.w, S4.z, S4.w, u2, u3);
                            fma2_bcast(Qa[q], A4.x, A4.y, u0, u1, t0, t1);
                            fma2_bcast(Qa[q], A4.z, A4.w, u2, u3, t2, t3);
                            mn[q] = min3f(min3f(mn[q], t0, t1), t2, t3);
                        }
                    }
                }
                // ---- groups that may hold a neighbour: the warp re-examines them, one sample per lane ----
                bool anyq = false;
#pragma unroll
                for (int q = 0; q < QPT; ++q) anyq |= mn[q] < thr[q];
                if (!__any_sync(kFull, anyq)) continue;  // the common case for large N: one vote per group
                const float* cs = comp + ((g0 + lane) >> 2) * 12 + (lane & 3);
                const float xa = cs[0], xb = cs[4], xs = cs[8];
#pragma unroll
                for (int q = 0; q < QPT; ++q) {
                    unsigned hit = __ballot_sync(kFull, mn[q] < thr[q]);
                    if (hit) {
                        do {
                            const int L = __ffs(hit) - 1;
                            hit &= hit - 1;
                            const float qa2 = __shfl_sync(kFull, Qa[q], L);
                            const float qb2 = __shfl_sync(kFull, Qb[COLSHARE ? 0 : q], L);
                            const float th = __shfl_sync(kFull, thr[q], L);
                            const int o = (int)((__shfl_sync(kFull, offp[q >> 2], L) >> ((q & 3) * 8)) & 0xffu);
                            const float t = fmaf(qa2, xa, fmaf(qb2, xb, xs));
                            const bool pass = t < th;
                            const unsigned pm = __ballot_sync(kFull, pass);
                            // 32-bit offset from the warp's candidate base: (q*32 + L)*CAP + o + rank among the passing lanes
                            if (pass) cand_w[(unsigned)((q * 32 + L) * CAP + o) + __popc(pm & lt_mask)] = jbase + (unsigned)(g0 + lane);
                            int no = o + __popc(pm);
                            float nthr = th;
                            const bool full = no > FLUSH_AT;
                            if (full) {
                                __syncwarp();
                                const bool has = (__shfl_sync(kFull, has_list, L) >> q) & 1u;
                                nthr = (float)rerank(q, L, no, has, false);
                                no = 0;
                            }
                            if (lane == L) {
                                offp[q >> 2] = (offp[q >> 2] & ~(0xffu << ((q & 3) * 8))) | ((unsigned)no << ((q & 3) * 8));
                                if (full) {
                                    thr[q] = nthr;
                                    has_list |= 1u << q;
                                }
                            }
                        } while (hit);
                    }
                }
            }
            __syncwarp();  // all lanes done with comp before the next stage overwrites it
        }

        // ---------------- final re-rank + epilogue ----------------
        double res[QPT];
        __syncwarp();
#pragma unroll
        for (int q = 0; q < QPT; ++q) {
            res[q] = 0.0;
            unsigned need = __ballot_sync(kFull, ((validm >> q) & 1u) != 0);
            while (need) {
                const int L = __ffs(need) - 1;
                need &= need - 1;
                const int cnt = (int)((__shfl_sync(kFull, offp[q >> 2], L) >> ((q & 3) * 8)) & 0xffu);
                const bool has = (__shfl_sync(kFull, has_list, L) >> q) & 1u;
                const double r = rerank(q, L, cnt, has, true);
                if (lane == L) res[q] = r;
            }
        }
        if (a.j.idw_mode == 1 && a.nseg == 1) {
#pragma unroll
            for (int q = 0; q < QPT; ++q) {
                if ((validm >> q) & 1u) {
                    double ta, tb;
                    long long m;
                    target_coords<QPT>(a, tile, q * 32 + lane, ta, tb, m);
                    if (a.j.value_is_f32)
                        static_cast<float*>(a.j.out)[m] = (float)res[q];
                    else
                        static_cast<double*>(a.j.out)[m] = res[q];
                }
            }
        }
    }
}

// Sample split, second step: one warp per target merges the nseg sorted partial lists and finishes.
template <int E>
__global__ void __launch_bounds__(256) knn_merge_kernel(const __grid_constant__ KnnArgs a) {
    constexpr int KP = 32 * E;
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long m = warp0; m < a.M; m += nwarps) {
        double ld[E];
        int li[E];
        const double* Pd = a.part_d + m * a.nseg * KP;
        const int* Pi = a.part_i + m * a.nseg * KP;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            ld[e] = Pd[e * 32 + lane];
            li[e] = Pi[e * 32 + lane];
        }
        for (int sg = 1; sg < a.nseg; ++sg) {
#pragma unroll
            for (int e = 0; e < E; ++e) {  // each 32-chunk of a sorted list is itself sorted
                const double cd = Pd[sg * KP + e * 32 + lane];
                const int ci = Pi[sg * KP + e * 32 + lane];
                merge_sorted32<E>(ld, li, cd, ci, lane);
            }
        }
        const double r = finish_target<E>(a, lane, m, ld, li);
        if (a.j.idw_mode == 1 && lane == 0) {
            if (a.j.value_is_f32)
                static_cast<float*>(a.j.out)[m] = (float)r;
            else
                static_cast<double*>(a.j.out)[m] = r;
        }
    }
}

// ------------------------------------------------------------------------------------------
// K1c: cell-binned search.  The summed-area table already holds, for every cell, how many samples lie
// in the rows before it and in its own row to the left of it, so the start of cell (r, c) in a
// row-major counting sort is  S[r][gb] + S[r+1][c] - S[r][c]  - no separate prefix scan is needed.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int bin_cell_start(const int* __restrict__ sat, int gb, int r, int c) {
    const int* S0 = sat + (size_t)r * SAT_LD;
    return S0[gb] + S0[SAT_LD + c] - S0[c];
}

__global__ void bin_scatter_kernel(const double* __restrict__ sa, const double* __restrict__ sb, long long n,
                                   const SeedGrid* g, const int* __restrict__ sat, int* cursor,
                                   double* __restrict__ bin_a, double* __restrict__ bin_b, int* __restrict__ bin_idx) {
    const SeedGrid s = *g;
    if (!s.valid) return;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double a = sa[i], b = sb[i];
        if (isfinite(a) && isfinite(b)) {
            const int ia = seed_cell(a, s.a0, s.inv_h, s.ga), ib = seed_cell(b, s.b0, s.inv_h, s.gb);
            const int pos = bin_cell_start(sat, s.gb, ia, ib) + atomicAdd(&cursor[(size_t)ia * s.gb + ib], 1);
            bin_a[pos] = a;
            bin_b[pos] = b;
            bin_idx[pos] = (int)i;
        }
    }
}

// Small sample sets (N <= SMALL_N): the whole pre-pass - bounding box, grid geometry, table clear, histogram,
// both scans and (for the binned search) the counting-sort scatter - in ONE single-CTA kernel instead of seven
// launches.  At BASELINE config 1 (N = 1000) the launches, not the work, set the latency of a reconstruct call.
constexpr int SMALL_N = 16384;
__global__ void __launch_bounds__(1024) seed_small_kernel(const double* __restrict__ sa, const double* __restrict__ sb,
                                                          int n, SeedGrid* g, int* sat, int* tile_counter, int* cursor,
                                                          double* __restrict__ bin_a, double* __restrict__ bin_b,
                                                          int* __restrict__ bin_idx) {
    __shared__ double red[4][32];
    __shared__ SeedGrid sg;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x, nwarps = blockDim.x >> 5;
    double lo_a = CMX_INF, hi_a = -CMX_INF, lo_b = CMX_INF, hi_b = -CMX_INF;
    for (int i = tid; i < n; i += nthreads) {
        const double a = sa[i], b = sb[i];
        if (isfinite(a) && isfinite(b)) {
            lo_a = fmin(lo_a, a);
            hi_a = fmax(hi_a, a);
            lo_b = fmin(lo_b, b);
            hi_b = fmax(hi_b, b);
        }
    }
    lo_a = warp_min(lo_a);
    hi_a = warp_max(hi_a);
    lo_b = warp_min(lo_b);
    hi_b = warp_max(hi_b);
    if (lane == 0) {
        red[0][warp] = lo_a;
        red[1][warp] = hi_a;
        red[2][warp] = lo_b;
        red[3][warp] = hi_b;
    }
    __syncthreads();
    if (warp == 0) {
        double bbox[4];
        bbox[0] = warp_min(lane < nwarps ? red[0][lane] : CMX_INF);
        bbox[1] = warp_max(lane < nwarps ? red[1][lane] : -CMX_INF);
        bbox[2] = warp_min(lane < nwarps ? red[2][lane] : CMX_INF);
        bbox[3] = warp_max(lane < nwarps ? red[3][lane] : -CMX_INF);
        if (lane == 0) {
            sg = make_seed_grid(bbox, n);
            *g = sg;
            *tile_counter = 0;
        }
    }
    __syncthreads();
    const SeedGrid s = sg;
    const int ga = s.ga, gb = s.gb;
    for (int i = tid; i < (ga + 1) * (gb + 1); i += nthreads) {
        const int row = i / (gb + 1), col = i - row * (gb + 1);
        sat[(size_t)row * SAT_LD + col] = 0;
        if (cursor && i < ga * gb) cursor[i] = 0;
    }
    __syncthreads();
    if (s.valid) {
        for (int i = tid; i < n; i += nthreads) {
            const double a = sa[i], b = sb[i];
            if (isfinite(a) && isfinite(b)) {
                const int ia = seed_cell(a, s.a0, s.inv_h, ga), ib = seed_cell(b, s.b0, s.inv_h, gb);
                atomicAdd(&sat[(size_t)(ia + 1) * SAT_LD + (ib + 1)], 1);
            }
        }
    }
    __syncthreads();
    for (int r = 1 + warp; r <= ga; r += nwarps) {  // inclusive scan along b, one warp per row
        int* row = sat + (size_t)r * SAT_LD;
        int carry = 0;
        for (int c0 = 1; c0 <= gb; c0 += 32) {
            const int c = c0 + lane;
            int v = c <= gb ? row[c] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(kFull, v, o);
                if (lane >= o) v += u;
            }
            v += carry;
            if (c <= gb) row[c] = v;
            carry = __shfl_sync(kFull, v, 31);
        }
    }
    __syncthreads();
    for (int c = 1 + tid; c <= gb; c += nthreads) {  // inclusive scan along a, one thread per column
        int acc = 0;
        for (int r = 1; r <= ga; ++r) {
            acc += sat[(size_t)r * SAT_LD + c];
            sat[(size_t)r * SAT_LD + c] = acc;
        }
    }
    if (!cursor || !s.valid) return;
    __syncthreads();
    for (int i = tid; i < n; i += nthreads) {
        const double a = sa[i], b = sb[i];
        if (isfinite(a) && isfinite(b)) {
            const int ia = seed_cell(a, s.a0, s.inv_h, ga), ib = seed_cell(b, s.b0, s.inv_h, gb);
            const int pos = bin_cell_start(sat, gb, ia, ib) + atomicAdd(&cursor[(size_t)ia * gb + ib], 1);
            bin_a[pos] = a;
            bin_b[pos] = b;
            bin_idx[pos] = i;
        }
    }
}

// seed_bound_d2 evaluated by a whole warp for one target: lane l tests box half-width w_l (0..8, then doubling)
// in one round, the first lane whose box holds >= k samples wins; the plus-shaped refinement tests 32 arm
// lengths in a second round and the three far-corner distances are evaluated by lanes 0..2 at once.
// Same family of rigorous bounds as seed_bound_d2 (any region with >= k samples bounds the k-th distance).
__device__ __forceinline__ double seed_bound_d2_warp(const SeedGrid& s, const int* __restrict__ sat, double ta,
                                                     double tb, int k, int lane) {
    const int ia = seed_cell(ta, s.a0, s.inv_h, s.ga), ib = seed_cell(tb, s.b0, s.inv_h, s.gb);
    const int wl = lane <= 8 ? lane : (lane < 30 ? 8 << (lane - 8) : 0x3fffffff);
    const unsigned okb = __ballot_sync(kFull, seed_box_count(s, sat, ia - wl, ia + wl, ib - wl, ib + wl) >= k);
    if (okb == 0u) return CMX_INF;  // fewer than k finite samples in total
    const int w = __shfl_sync(kFull, wl, __ffs(okb) - 1);
    int A = w, B = w;
    unsigned okp = 0u;
    if (w >= 2 && w <= 64) {
        // plus shape: [ia-A, ia+A] x [ib-B, ib+B]  U  [ia-B, ia+B] x [ib-A, ib+A],  B ~ 0.6 A
        const int Al = w + lane, Bl = (3 * Al + 2) / 5;
        const int cnt = seed_box_count(s, sat, ia - Al, ia + Al, ib - Bl, ib + Bl) +
                        seed_box_count(s, sat, ia - Bl, ia + Bl, ib - Al, ib + Al) -
                        seed_box_count(s, sat, ia - Bl, ia + Bl, ib - Bl, ib + Bl);
        okp = __ballot_sync(kFull, cnt >= k && lane <= (w >> 1) + 1);
        if (okp) {
            A = w + __ffs(okp) - 1;
            B = (3 * A + 2) / 5;
        }
    }
    // lane 0: the square; lanes 1, 2: the two rectangles of the plus (both inside the disc through their far corners)
    const int ha = lane == 0 ? w : (lane == 1 ? A : B), hb = lane == 0 ? w : (lane == 1 ? B : A);
    const double far2 = seed_box_far2(s, ta, tb, ia - ha, ia + ha, ib - hb, ib + hb);
    const double f0 = __shfl_sync(kFull, far2, 0), f1 = __shfl_sync(kFull, far2, 1), f2 = __shfl_sync(kFull, far2, 2);
    return okp ? fmin(f0, fmax(f1, f2)) : f0;
}

// One warp per target.  The seed bound D2 >= (k-th nearest distance)^2 is rigorous, so every one of the k
// nearest samples lies in the cells that the square [t - D, t + D]^2 touches; in the sorted order each cell
// row of that square is one contiguous segment.  Lanes own one cell row each to look the segments up, the
// segments are then flattened so that all 32 lanes evaluate exact float64 distances, and candidates within the
// running k-th distance go through the same bitonic sort / merge network as K1's re-rank.  The order of the
// samples inside a cell (atomics in the scatter) cannot change the result: the list is the lexicographic
// (d2, index) top-k of a candidate set that always contains the true top-k.
// Chain: the previous target this warp searched (coordinates pa, pb; exact k-th squared distance pk, inf = none).
// Every one of its k neighbours lies within d_k(prev) + |prev - target| of this target (triangle inequality), so that
// sum bounds this target's k-th distance - for adjacent grid points a tighter start than the summed-area-table box,
// and it costs three flops instead of a table walk.  Used when the step is at most half of d_k(prev).
struct SearchChain {
    double pa, pb, pk;
};
template <int E>
__device__ __forceinline__ void binned_search_target(const KnnArgs& a, const SeedGrid& s, const int* __restrict__ sat, int k,
                                                     int lane, long long m, double (&ld)[E], int (&li)[E], SearchChain& ch) {
    double ta, tb;
    if (a.j.grid) {
        const long long row = m / a.j.n_b;
        ta = a.j.t_a[row];
        tb = a.j.t_b[m - row * a.j.n_b];
    } else {
        ta = a.j.t_a[m];
        tb = a.j.t_b[m];
    }
#pragma unroll
    for (int e = 0; e < E; ++e) {
        ld[e] = CMX_INF;
        li[e] = INT_MAX;
    }
    int ra0 = 0, ra1 = -1, cb0 = 0, cb1 = -1;
    double kth = CMX_INF;
    bool empty = true;
    if (s.valid) {
        const double sa = ta - ch.pa, sb = tb - ch.pb;
        const double step2 = sa * sa + sb * sb;
        if (ch.pk < CMX_INF && step2 <= 0.25 * ch.pk) {  // (false for NaN steps and an unset chain)
            const double D = (sqrt(ch.pk) + sqrt(step2)) * (1.0 + 1e-12);
            kth = D * D * (1.0 + 1e-12);
        } else {
            kth = seed_bound_d2_warp(s, sat, ta, tb, k, lane);
        }
        if (kth < CMX_INF) {
            const double D = sqrt(kth) * (1.0 + 1e-12);
            ra0 = seed_cell(ta - D, s.a0, s.inv_h, s.ga);
            ra1 = seed_cell(ta + D, s.a0, s.inv_h, s.ga);
            cb0 = seed_cell(tb - D, s.b0, s.inv_h, s.gb);
            cb1 = seed_cell(tb + D, s.b0, s.inv_h, s.gb);
        } else {  // fewer than k finite samples (or a non-finite target): look at everything
            ra1 = s.ga - 1;
            cb1 = s.gb - 1;
        }
    }
    const int R = ra1 - ra0 + 1;
    for (int rb = 0; rb < R; rb += 32) {
        int lo = 0, len = 0;
        if (rb + lane < R) {
            const int* S0 = sat + (size_t)(ra0 + rb + lane) * SAT_LD;
            const int left = S0[SAT_LD + cb0] - S0[cb0];
            lo = S0[s.gb] + left;
            len = (S0[SAT_LD + cb1 + 1] - S0[cb1 + 1]) - left;
        }
        int pre = len;  // inclusive prefix sum of the segment lengths
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(kFull, pre, o);
            if (lane >= o) pre += u;
        }
        const int total = __shfl_sync(kFull, pre, 31);
        const int excl = pre - len;
        for (int j0 = 0; j0 < total; j0 += 32) {
            const int j = j0 + lane;
            int pos = 0;  // number of segments that end at or before j = the segment holding j
#pragma unroll
            for (int step = 16; step >= 1; step >>= 1) {
                const int v = __shfl_sync(kFull, pre, pos + step - 1);
                if (v <= j) pos += step;
            }
            const int p = __shfl_sync(kFull, lo, pos) + (j - __shfl_sync(kFull, excl, pos));
            double d2 = CMX_INF;
            if (j < total) {
                // bit-identical to cKDTree: plain float64 dx*dx + dy*dy, no FMA contraction
                const double da = __dsub_rn(ta, a.bin_a[p]);
                const double db = __dsub_rn(tb, a.bin_b[p]);
                d2 = __dadd_rn(__dmul_rn(da, da), __dmul_rn(db, db));
                if (!(d2 == d2)) d2 = CMX_INF;  // NaN never ranks
            }
            // ties at the k-th distance are resolved by index inside the merge, so '<=' here
            const bool c = d2 <= kth && d2 < CMX_INF;
            if (__any_sync(kFull, c)) {
                double cd = c ? d2 : CMX_INF;
                int ci = c ? a.bin_idx[p] : INT_MAX;
                warp_sort32(cd, ci, lane);
                if (empty) {  // first candidates of this target: the sorted chunk is the list
                    ld[0] = cd;
                    li[0] = ci;
                    empty = false;
                } else {
                    merge_sorted32<E>(ld, li, cd, ci, lane);
                }
                double sel = ld[0];
#pragma unroll
                for (int e = 1; e < E; ++e)
                    if (((k - 1) >> 5) == e) sel = ld[e];
                kth = fmin(kth, __shfl_sync(kFull, sel, (k - 1) & 31));
            }
        }
    }
    {   // hand the exact k-th distance to the next target of this warp
        double sel = ld[0];
#pragma unroll
        for (int e = 1; e < E; ++e)
            if (((k - 1) >> 5) == e) sel = ld[e];
        ch.pa = ta;
        ch.pb = tb;
        ch.pk = __shfl_sync(kFull, sel, (k - 1) & 31);  // inf when the list holds fewer than k samples
    }
}

// P = 0: neighbour table and / or the single-slice IDW epilogue.  P = 2, 4: short time / level batches (up to P * 32
// slices) with the fused batch epilogue - a warp takes FOUR consecutive targets so that every lane (= batch element)
// ends up with a 32-byte (16-byte for float32) row segment to store.
template <int E, int P>
__global__ void __launch_bounds__(256) knn_binned_kernel(const __grid_constant__ KnnArgs a) {
    const int lane = threadIdx.x & 31;
    const SeedGrid s = *a.seed;
    const int* __restrict__ sat = a.sat;
    const int k = a.j.k;
    // Work is handed out in CHUNKS of consecutive targets per CTA (not strided over the grid): consecutive targets share
    // most of their neighbours, cells and table entries, so what one iteration pulled into L1 is reused by the next
    // (the batch epilogue's value-row gathers hit L1 instead of L2).
    const int CHUNK = a.chunk;  // units (targets, or groups of four targets) per CTA and chunk; a multiple of 8
    const int wib = threadIdx.x >> 5;
    // A warp takes RUN consecutive units, then strides over the other warps' runs: consecutive targets feed the chained
    // bound (SearchChain), while the 8 warps of the CTA stay on neighbouring targets so that they share L1 lines.
    const int RUN = a.run;
    SearchChain chain{0.0, 0.0, CMX_INF};
    if constexpr (P == 0) {
        for (long long base = (long long)blockIdx.x * CHUNK; base < a.M; base += (long long)gridDim.x * CHUNK) {
            const long long cend = base + CHUNK < a.M ? base + CHUNK : a.M;
            for (long long first = base + (long long)wib * RUN; first < cend; first += 8LL * RUN)
            for (long long m = first; m < (first + RUN < cend ? first + RUN : cend); ++m) {
                double ld[E];
                int li[E];
                binned_search_target<E>(a, s, sat, k, lane, m, ld, li, chain);
                const double r = finish_target<E>(a, lane, m, ld, li);
                if (a.j.idw_mode == 1 && lane == 0) {
                    if (a.j.value_is_f32)
                        static_cast<float*>(a.j.out)[m] = (float)r;
                    else
                        static_cast<double*>(a.j.out)[m] = r;
                }
            }
        }
    } else {
        constexpr int G = 4;
        __shared__ BatchNbr strips[8][32 * E];
        __shared__ double stage[8][G][P * 32];  // the group's results: [target][batch element]
        BatchNbr* strip = strips[wib];
        double (*st)[P * 32] = stage[wib];
        const bool tables = a.j.idx_out || a.j.dist_out || a.j.d2_out;
        const long long groups = (a.M + G - 1) / G;
        for (long long base = (long long)blockIdx.x * CHUNK; base < groups; base += (long long)gridDim.x * CHUNK)
        for (long long first = base + (long long)wib * RUN; first < (base + CHUNK < groups ? base + CHUNK : groups); first += 8LL * RUN)
        for (long long g = first; g < first + RUN && g < base + CHUNK && g < groups; ++g) {
            const long long m0 = g * G;
            const int count = (int)(a.M - m0 < G ? a.M - m0 : G);
            // a run-time loop: ONE copy of the search + epilogue code (four inlined copies overflow the instruction
            // cache: ncu showed 21 no_instruction stalls per issue)
#pragma unroll 1
            for (int q = 0; q < count; ++q) {
                double ld[E];
                int li[E];
                binned_search_target<E>(a, s, sat, k, lane, m0 + q, ld, li, chain);
                if (tables) finish_target<E>(a, lane, m0 + q, ld, li);
                double r[P];
                finish_batch<E, P>(a, lane, ld, li, strip, r);
#pragma unroll
                for (int p = 0; p < P; ++p) st[q][p * 32 + lane] = r[p];
            }
            __syncwarp();
            double rr[G][P];
#pragma unroll
            for (int q = 0; q < G; ++q)
#pragma unroll
                for (int p = 0; p < P; ++p) rr[q][p] = q < count ? st[q][p * 32 + lane] : 0.0;
            store_batch<P, G>(a, lane, m0, count, rr);
            __syncwarp();
        }
    }
}

template <int E>
int launch_binned(const KnnArgs& args_, bool scatter, cudaStream_t stream) {
    const int sms = sm_count();
    KnnArgs args = args_;
    {   // chunks of 64 units when that still gives every CTA slot work, else the finest split (8 = one unit per warp)
        const long long units = args.j.idw_mode == 2 ? (args.M + 3) / 4 : args.M;
        args.chunk = units >= 64LL * sms * 8 ? 64 : (units >= 16LL * sms * 8 ? 16 : 8);
        args.run = args.chunk / 8;
        if (const char* e = getenv("CMX_K1C_RUN")) args.run = atoi(e);
        if (args.run < 1) args.run = 1;
    }
    timing_begin(CMX_KERNEL_KNN_BINNED, stream);
    if (scatter) {  // (small sample sets: already done by seed_small_kernel; the cursors were cleared by the pre-pass)
        const long long pre = (args.j.n + 255) / 256;
        const int pre_blocks = (int)(pre < (long long)sms * 8 ? pre : (long long)sms * 8);
        bin_scatter_kernel<<<pre_blocks, 256, 0, stream>>>(args.j.s_a, args.j.s_b, args.j.n, args.seed, args.sat,
                                                           args.bin_cursor, args.bin_a, args.bin_b, args.bin_idx);
        note_launch();
    }
    const long long cap = (long long)sms * 8;
    if (args.j.idw_mode == 2) {
        const long long blocks = ((args.M + 3) / 4 + args.chunk - 1) / args.chunk;  // one chunk per CTA and round
        const unsigned grid = (unsigned)(blocks < cap ? blocks : cap);
        if (args.j.n_batch <= 64)
            knn_binned_kernel<E, 2><<<grid, 256, 0, stream>>>(args);
        else
            knn_binned_kernel<E, 4><<<grid, 256, 0, stream>>>(args);
    } else {
        const long long blocks = (args.M + args.chunk - 1) / args.chunk;  // one chunk per CTA and round
        knn_binned_kernel<E, 0><<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, stream>>>(args);
    }
    timing_end(CMX_KERNEL_KNN_BINNED, stream);
    note_launch();
    return check_launch();
}

template <int E, int QPT>
int launch(const KnnArgs& args, int ctas, cudaStream_t stream) {
    const size_t smem = WARPS * WARP_SMEM;
    timing_begin(CMX_KERNEL_KNN, stream);
    if (args.j.grid) {
        CMX_CUDA(cudaFuncSetAttribute(knn_kernel<E, QPT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        knn_kernel<E, QPT, true><<<ctas, THREADS, smem, stream>>>(args);
    } else {
        CMX_CUDA(cudaFuncSetAttribute(knn_kernel<E, QPT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        knn_kernel<E, QPT, false><<<ctas, THREADS, smem, stream>>>(args);
    }
    if (args.nseg > 1) {
        const long long mblocks = (args.M * 32 + 255) / 256;
        const long long cap = (long long)sm_count() * 16;
        knn_merge_kernel<E><<<(unsigned)(mblocks < cap ? mblocks : cap), 256, 0, stream>>>(args);
        note_launch();
    }
    timing_end(CMX_KERNEL_KNN, stream);
    note_launch();
    return check_launch();
}

inline int list_slots(int k) { return k <= 32 ? 1 : (k <= 64 ? 2 : 4); }

constexpr size_t SEED_BYTES = (512 + (size_t)SAT_LD * SAT_LD * sizeof(int) + 255) & ~size_t(255);

}  // namespace

namespace {
struct Plan {
    int qpt, tiles_c, nseg;
    long long tiles;
};

// Targets per thread: 16 for large sample sets (fewest shared-memory loads per pair; throughput bound),
// 8 by default, 2 or 1 for small problems (latency bound: per-target re-ranks are serial inside a warp,
// so they are spread over more warps).  A band of a big job (multi-GPU rank, second band of the host
// pipeline) has too few tiles for that: there the sample range is split into nseg segments and tile
// height and nseg come from a small cost model.
Plan make_plan(long long n, long long n_a, long long n_b, int grid) {
    const long long M = grid ? n_a * n_b : n_a;
    const int sms = sm_count();
    auto tiles_for = [&](int qpt, int& tiles_c) -> long long {  // warp tiles: qpt rows x 32 columns
        if (grid) {
            tiles_c = (int)((n_b + 31) / 32);
            return ((n_a + qpt - 1) / qpt) * (long long)tiles_c;
        }
        tiles_c = 1;
        return (M + qpt * 32 - 1) / (qpt * 32);
    };
    int tc16 = 1, tc8 = 1, tc2 = 1, tc1 = 1;
    const long long t16 = tiles_for(16, tc16), t8 = tiles_for(8, tc8), t2 = tiles_for(2, tc2), t1 = tiles_for(1, tc1);
    const long long warps16 = (long long)sms * ctas_per_sm(16) * WARPS, warps8 = (long long)sms * ctas_per_sm(8) * WARPS;
    Plan p;
    p.nseg = 1;
    const long long n_st = (n + SS - 1) / SS;
    if (n >= 262144 && t16 >= 2 * warps16) {
        // plenty of tiles: biggest tile, no split (a little splitting only if it evens out the last round)
        p.qpt = 16, p.tiles = t16, p.tiles_c = tc16;
        while (p.nseg < 8 && p.tiles * p.nseg < 3 * warps16 && n_st / (p.nseg * 2) >= 128) p.nseg *= 2;
    } else if (n >= 262144 && t16 * 8 >= warps16) {
        // A band of a big job (one rank of a multi-GPU run, the second band of the host pipeline): fewer tiles
        // than a few rounds of warps.  Pick tile height and sample split from a cost model fitted to sweeps on
        // B200 (DESIGN.md K1 "planner"): padded rows x (1 + re-rank overhead per segment) x the cost of the
        // partially filled last round of work units (the kernel is pipe-bound, so a round with a fraction f of
        // the warps busy costs about 0.3 + 0.7 f of a full one).
        double best = 1e300;
        for (int qi = 0; qi < 2; ++qi) {
            const int q = qi ? 8 : 16;
            const long long tiles = qi ? t8 : t16, warps = qi ? warps8 : warps16;
            const long long rows = grid ? ((n_a + q - 1) / q) * q : tiles;  // padded work, in rows (or tiles)
            for (int sg = 1; sg <= 8; sg *= 2) {
                if (sg > 1 && n_st / (sg * 2) < 128) break;
                const double u = (double)tiles * sg / (double)warps;
                const double fl = floor(u), fr = u - fl;
                const double rounds = fl + (fr > 1e-9 ? 0.3 + 0.7 * fr : 0.0);
                const double cost = (double)rows * (grid ? 1.0 : q) * (1.0 + 0.01 * (1.0e6 / (double)n) * sg) * rounds / u;
                if (cost < best) {
                    best = cost;
                    p.qpt = q, p.tiles = tiles, p.tiles_c = qi ? tc8 : tc16, p.nseg = sg;
                }
            }
        }
    } else if (t8 >= 2 * warps8) {
        p.qpt = 8, p.tiles = t8, p.tiles_c = tc8;
    } else if (t2 >= warps8) {
        p.qpt = 2, p.tiles = t2, p.tiles_c = tc2;
    } else {
        p.qpt = 1, p.tiles = t1, p.tiles_c = tc1;
    }
#ifdef CMX_PLAN_ENV  // tuning builds only: CMX_FORCE_QPT / CMX_FORCE_NSEG override the planner
    if (const char* fq = getenv("CMX_FORCE_QPT")) {
        const int q = atoi(fq);
        if (q == 16) p.qpt = 16, p.tiles = t16, p.tiles_c = tc16;
        if (q == 8) p.qpt = 8, p.tiles = t8, p.tiles_c = tc8;
        p.nseg = 1;
    }
    if (const char* fs = getenv("CMX_FORCE_NSEG")) p.nseg = atoi(fs);
#endif
    return p;
}
}  // namespace

size_t knn_workspace_bytes(int k) {
    if (k < 1 || k > CMX_MAX_K) return 0;
    const size_t nslots = (size_t)sm_count() * SLOTS_PER_SM;
    const size_t kp = 32 * (size_t)list_slots(k);
    return nslots * (CAP * sizeof(unsigned) + kp * (sizeof(double) + sizeof(int))) + SEED_BYTES + 512;
}

size_t knn_workspace_bytes_for(int k, long long n, long long n_a, long long n_b, int grid) {
    const size_t base = knn_workspace_bytes(k);
    if (!base || n < 0 || n_a < 0 || n_b < 0) return base;
    const Plan p = make_plan(n, n_a, n_b, grid);
    if (p.nseg <= 1) return base;
    const long long M = grid ? n_a * n_b : n_a;
    return base + (size_t)M * p.nseg * 32 * list_slots(k) * (sizeof(double) + sizeof(int)) + 512;
}

void knn_plan(long long n, long long n_a, long long n_b, int grid, int* tile_rows, int* sample_segments) {
    const Plan p = make_plan(n, n_a, n_b, grid);
    *tile_rows = p.qpt;
    *sample_segments = p.nseg;
}

int run_knn(const KnnJob& job, cudaStream_t stream) {
    if (!job.s_a || !job.s_b || !job.t_a || !job.t_b) return CMX_ERR_BAD_ARG;
    if (job.n < 0 || job.n_a < 0 || job.n_b < 0 || job.k < 1) return CMX_ERR_BAD_ARG;
    if (!job.grid && job.n_a != job.n_b) return CMX_ERR_BAD_ARG;
    if (job.k > CMX_MAX_K) return CMX_ERR_K_TOO_LARGE;
    if ((long long)job.k > job.n) return CMX_ERR_K_GT_N;
    if (job.idw_mode == 1 && (!job.vals || !job.out || job.k_min < 1)) return CMX_ERR_BAD_ARG;
    if (job.idw_mode == 2 && (!job.vals || job.k_min < 1 || job.n_batch < 2 || job.n_batch > 128 || job.peers.n < 1))
        return CMX_ERR_BAD_ARG;
    const long long M = job.grid ? job.n_a * job.n_b : job.n_a;
    if (M == 0) return CMX_OK;
    if (job.n > 0x7fffffffLL) return CMX_ERR_TOO_MANY;
    const size_t need = knn_workspace_bytes(job.k);
    if (!job.workspace || job.workspace_bytes < need) return CMX_ERR_WORKSPACE;

    KnnArgs args;
    args.j = job;
    args.M = M;
    args.aligned = ((reinterpret_cast<uintptr_t>(job.s_a) | reinterpret_cast<uintptr_t>(job.s_b)) & 15u) == 0;
    const int sms = sm_count();
    Plan plan = make_plan(job.n, job.n_a, job.n_b, job.grid);
    if (plan.tiles > 0x7fffffffLL / 8) return CMX_ERR_TOO_MANY;
    const int qpt = plan.qpt;
    args.tiles_c = plan.tiles_c;
    args.num_tiles = (int)plan.tiles;
    const int slots = sms * ctas_per_sm(qpt);

    // workspace: [seed grid | lists d | candidates | lists i]
    const size_t kp = 32 * (size_t)list_slots(job.k);
    unsigned char* ws = static_cast<unsigned char*>(job.workspace);
    ws = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255));
    SeedGrid* seed = reinterpret_cast<SeedGrid*>(ws);
    double* bbox = reinterpret_cast<double*>(ws + 128);
    args.tile_counter = reinterpret_cast<int*>(ws + 192);
    int* sat = reinterpret_cast<int*>(ws + 512);
    ws += SEED_BYTES;
    const size_t nslots = (size_t)sms * SLOTS_PER_SM;
    args.list_d = reinterpret_cast<double*>(ws);
    args.cand = reinterpret_cast<unsigned*>(ws + nslots * kp * sizeof(double));
    args.list_i = reinterpret_cast<int*>(ws + nslots * kp * sizeof(double) + nslots * CAP * sizeof(unsigned));
    // optional partial-list tables of the sample split: used only if the caller's workspace has room
    unsigned char* part = ws + nslots * (kp * (sizeof(double) + sizeof(int)) + CAP * sizeof(unsigned));
    part = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(part) + 255) & ~uintptr_t(255));
    const size_t part_need = (size_t)M * plan.nseg * kp * (sizeof(double) + sizeof(int));
    unsigned char* ws_end = static_cast<unsigned char*>(job.workspace) + job.workspace_bytes;
    if (plan.nseg > 1 && part + part_need > ws_end) plan.nseg = 1;
    args.nseg = plan.nseg;
    args.part_d = reinterpret_cast<double*>(part);
    args.part_i = reinterpret_cast<int*>(part + (size_t)M * plan.nseg * kp * sizeof(double));
    const long long units = plan.tiles * plan.nseg;
    const long long ctas_needed = (units + WARPS - 1) / WARPS;
    const int ctas = (int)(ctas_needed < slots ? ctas_needed : slots);
    args.seed = seed;
    args.sat = sat;

    // K1c borrows K1's (unused) list / candidate scratch: cursors, sorted coordinates, original indices
    bool binned = job.search != CMX_SEARCH_BRUTE;
    {
        unsigned char* b = ws;
        args.bin_cursor = reinterpret_cast<int*>(b);
        b += (size_t)GMAX * GMAX * sizeof(int);
        args.bin_a = reinterpret_cast<double*>(b);
        b += (size_t)job.n * sizeof(double);
        args.bin_b = reinterpret_cast<double*>(b);
        b += (size_t)job.n * sizeof(double);
        args.bin_idx = reinterpret_cast<int*>(b);
        b += (size_t)job.n * sizeof(int);
        if (b > ws_end) binned = false;  // does not fit (N beyond ~5e7 with the base workspace): K1 returns the same table
    }
    if (job.idw_mode == 2 && !binned) return CMX_ERR_UNSUPPORTED;  // the fused batch epilogue lives in K1c only

    // ---- seed grid pre-pass: bbox -> geometry -> histogram -> summed-area table (-> counting sort) ----
    const bool small = job.n <= SMALL_N;
    if (small) {
        seed_small_kernel<<<1, 1024, 0, stream>>>(job.s_a, job.s_b, (int)job.n, seed, sat, args.tile_counter,
                                                  binned ? args.bin_cursor : nullptr, args.bin_a, args.bin_b, args.bin_idx);
        note_launch();
    } else {
        const int pre_blocks = (int)((job.n + 255) / 256 < (long long)sms * 8 ? (job.n + 255) / 256 : (long long)sms * 8);
        seed_init_kernel<<<1, 1, 0, stream>>>(bbox, args.tile_counter);
        seed_bbox_kernel<<<pre_blocks, 256, 0, stream>>>(job.s_a, job.s_b, job.n, bbox);
        seed_geometry_kernel<<<1, 1, 0, stream>>>(bbox, job.n, seed);
        seed_clear_kernel<<<sms * 4, 256, 0, stream>>>(seed, sat, binned ? args.bin_cursor : nullptr);
        seed_hist_kernel<<<pre_blocks, 256, 0, stream>>>(job.s_a, job.s_b, job.n, seed, sat);
        seed_rowscan_kernel<<<GMAX, 32, 0, stream>>>(seed, sat);
        seed_colscan_kernel<<<(GMAX + 127) / 128, 128, 0, stream>>>(seed, sat);
        for (int i = 0; i < 7; ++i) note_launch();
    }
    int rc = check_launch();
    if (rc != CMX_OK) return rc;

    const int e = list_slots(job.k);
    if (binned) {
        args.nseg = 1;
        if (e == 1) return launch_binned<1>(args, !small, stream);
        if (e == 2) return launch_binned<2>(args, !small, stream);
        return launch_binned<4>(args, !small, stream);
    }
#define CMX_DISPATCH(Q)                                            \
    do {                                                           \
        if (e == 1) return launch<1, Q>(args, ctas, stream);       \
        if (e == 2) return launch<2, Q>(args, ctas, stream);       \
        return launch<4, Q>(args, ctas, stream);                   \
    } while (0)
    if (qpt == 16) CMX_DISPATCH(16);
    if (qpt == 8) CMX_DISPATCH(8);
    if (qpt == 2) CMX_DISPATCH(2);
    CMX_DISPATCH(1);
#undef CMX_DISPATCH
}

}  // namespace cmx
