// K4: batched moving-window ordinary kriging -- one warp per target.
//
// Replaces pykrige OrdinaryKriging.execute(..., backend="loop", n_closest_points=k)
// (_exec_loop_moving_window), reached from the reference call site
// src/climatrix/reconstruct/kriging.py:236-241; algorithm restated in SURVEY.md
// Appendix A.2 step 5.  Per target, with its k nearest samples s_1..s_k (from K1, run
// in the kriging frame) and distances bd_j:
//
//     [ G   1 ] [x ]   [ b ]      G_ij = -gamma(|s_i - s_j|), G_ii = 0
//     [ 1^T 0 ] [mu] = [ 1 ]      b_j  = -gamma(bd_j)   (0 where bd_j <= 1e-10)
//     z = sum_j x_j Z_j ,  sigma^2 = -(x.b + mu)
//
// The bordered system is solved by block elimination on the k x k core with two
// right-hand sides (G y = b, G u = 1, mu = (1.y - 1)/(1.u), x = y - mu u), so that the
// k rows map 1:1 onto the lanes of a warp.  G is symmetric indefinite (zero diagonal):
// Gauss-Jordan elimination with partial (row) pivoting in float64, the matrix held
// column-major in shared memory (conflict-free lane-per-row updates, pivot row read by
// broadcast).  Tensor cores are not applicable: these are M independent tiny solves.
// FP64-pipe + shared-memory bound; algorithmic FLOP per target
//   ~ k^2 (5 + gamma) assemble + k^3 Gauss-Jordan + 6k.
#include "common.cuh"

namespace cmx {
namespace {

constexpr double OK_EPS = 1.0e-10;  // pykrige ok.py: eps

struct OkArgs {
    const double* s_x;
    const double* s_y;
    const double* z;    // [N] standardised values (always float64)
    int out_is_f32;     // dtype of out / sigmasq
    const int* idx;     // [M][k]
    const double* d2;   // [M][k] exact squared distances (kriging frame)
    long long M;
    long long n;  // samples (indices outside [0, n) mark a short neighbour list -> NaN)
    int k;
    int model;
    double p0, p1, p2;
    const double* params_dev;  // non-null: the three parameters live in device memory (written by the fit kernel)
    int exact_values;
    int metric;  // CMX_METRIC_EUCLIDEAN: d2 = squared Euclidean; GEOGRAPHIC: d2 = squared chord, s_x = lon, s_y = lat (degrees)
    double out_scale, out_shift;
    void* out;
    void* sigmasq;
    int* n_singular;
    // optional work list (the windows the Cholesky kernel handed over): window w of *list_count is target list[w]
    const int* list;
    const int* list_count;
};

// the variogram parameters: kernel arguments, or three doubles in device memory
__device__ __forceinline__ void load_params(const OkArgs& a, double& p0, double& p1, double& p2) {
    if (a.params_dev) {
        p0 = a.params_dev[0];
        p1 = a.params_dev[1];
        p2 = a.params_dev[2];
    } else {
        p0 = a.p0;
        p1 = a.p1;
        p2 = a.p2;
    }
}
__device__ __forceinline__ long long window_count(const OkArgs& a) { return a.list ? (long long)*a.list_count : a.M; }
__device__ __forceinline__ long long window_target(const OkArgs& a, long long w) { return a.list ? (long long)a.list[w] : w; }
__device__ __forceinline__ void store_window(const OkArgs& a, long long m, double res, double var) {
    if (a.out_is_f32) {
        static_cast<float*>(a.out)[m] = (float)res;
        if (a.sigmasq) static_cast<float*>(a.sigmasq)[m] = (float)var;
    } else {
        static_cast<double*>(a.out)[m] = res;
        if (a.sigmasq) static_cast<double*>(a.sigmasq)[m] = var;
    }
}

// pykrige/variogram_models.py
__device__ __forceinline__ double gamma_of(int model, double p0, double p1, double p2, double d) {
    switch (model) {
        case CMX_VARIOGRAM_LINEAR:  // slope * d + nugget
            return p0 * d + p1;
        case CMX_VARIOGRAM_POWER:  // scale * d^exponent + nugget
            return p0 * pow(d, p1) + p2;
        case CMX_VARIOGRAM_GAUSSIAN: {  // psill (1 - exp(-d^2 / (range 4/7)^2)) + nugget
            const double r = p1 * 4.0 / 7.0;
            return p0 * (1.0 - exp(-(d * d) / (r * r))) + p2;
        }
        case CMX_VARIOGRAM_SPHERICAL:  // psill (3d/(2r) - d^3/(2r^3)) + nugget, d <= r; psill + nugget beyond
            if (d <= p1) return p0 * ((3.0 * d) / (2.0 * p1) - (d * d * d) / (2.0 * p1 * p1 * p1)) + p2;
            return p0 + p2;
        default:  // exponential: psill (1 - exp(-d / (range/3))) + nugget
            return p0 * (1.0 - exp(-d / (p1 / 3.0))) + p2;
    }
}

// distance target -> neighbour from the stored squared distance of the neighbour search
__device__ __forceinline__ double window_distance(int metric, double d2) {
    const double d = sqrt(d2);
    return metric == CMX_METRIC_GEOGRAPHIC ? chord_to_great_circle_deg(d) : d;
}
// distance between two samples of the window (pykrige _get_kriging_matrix: cdist / great_circle_distance)
__device__ __forceinline__ double sample_distance(int metric, double xi, double yi, double xj, double yj) {
    if (metric == CMX_METRIC_GEOGRAPHIC) return great_circle_deg(xi, yi, xj, yj);
    const double dx = __dsub_rn(xi, xj), dy = __dsub_rn(yi, yj);
    return sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}

// E = ceil(k / 32) rows per lane; ld = 32 * E (column stride, in doubles)
template <int E>
__global__ void __launch_bounds__(256) ok_solve_kernel(const OkArgs a, int warps_per_cta) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int LD = 32 * E;
    const int k = a.k;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // per warp: matrix (k+2) columns x LD rows, then neighbour coordinates, then pivot rows
    double P0, P1, P2;
    load_params(a, P0, P1, P2);
    const size_t per_warp = (size_t)(k + 2) * LD * sizeof(double) + 2 * LD * sizeof(double) + LD * sizeof(int);
    unsigned char* base = smem_raw + (size_t)warp * per_warp;
    double* mat = reinterpret_cast<double*>(base);
    double* cx = mat + (size_t)(k + 2) * LD;
    double* cy = cx + LD;
    int* piv = reinterpret_cast<int*>(cy + LD);
    double* colB = mat + (size_t)k * LD;        // right-hand side b
    double* colO = mat + (size_t)(k + 1) * LD;  // right-hand side of ones

    const long long total_warps = (long long)gridDim.x * warps_per_cta;
    const long long n_windows = window_count(a);
    for (long long wi = (long long)blockIdx.x * warps_per_cta + warp; wi < n_windows; wi += total_warps) {
        const long long m = window_target(a, wi);
        // ---- gather the window ----
        double mx[E], my[E], mz[E], borig[E];
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int r = e * 32 + lane;
            mx[e] = my[e] = mz[e] = borig[e] = 0.0;
            if (r < k) {
                int id = a.idx[m * k + r];
                if (id < 0 || id >= a.n) {  // short neighbour list: duplicate row 0 -> singular window -> NaN
                    id = a.idx[m * k];
                    if (id < 0 || id >= a.n) id = 0;
                }
                mx[e] = a.s_x[id];
                my[e] = a.s_y[id];
                mz[e] = a.z[id];
                const double bd = window_distance(a.metric, a.d2[m * k + r]);
                double b = -gamma_of(a.model, P0, P1, P2, bd);
                if (a.exact_values && fabs(bd) <= OK_EPS) b = 0.0;
                borig[e] = b;
                cx[r] = mx[e];
                cy[r] = my[e];
                colB[r] = b;
                colO[r] = 1.0;
            }
        }
        __syncwarp();
        // ---- assemble G (column j = neighbour j; lanes own rows) ----
        for (int j = 0; j < k; ++j) {
            const double xj = cx[j], yj = cy[j];
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int r = e * 32 + lane;
                if (r < k) {
                    const double d = sample_distance(a.metric, mx[e], my[e], xj, yj);
                    mat[(size_t)j * LD + r] = (r == j) ? 0.0 : -gamma_of(a.model, P0, P1, P2, d);
                }
            }
        }
        __syncwarp();
        // ---- Gauss-Jordan with partial pivoting ----
        unsigned done = 0;  // bit e: my row e*32+lane has been a pivot
        bool singular = false;
        // k == 1: the core is the 1x1 zero matrix while the bordered system [[0,1],[1,0]] is regular
        // (x = 1, mu = b); handled after the loop
        for (int c = 0; c < k && k > 1; ++c) {
            double best = -1.0;
            int brow = 0x7fffffff;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int r = e * 32 + lane;
                if (r < k && !((done >> e) & 1u)) {
                    const double v = fabs(mat[(size_t)c * LD + r]);
                    if (v > best) {  // first maximum wins inside a lane (rows ascend with e)
                        best = v;
                        brow = r;
                    }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(kFull, best, o);
                const int orow = __shfl_xor_sync(kFull, brow, o);
                if (ob > best || (ob == best && orow < brow)) {
                    best = ob;
                    brow = orow;
                }
            }
            if (!(best > 0.0)) {  // zero (or NaN) column: singular window
                singular = true;
                break;
            }
            const int p = brow;
            if ((p & 31) == lane) done |= 1u << (p >> 5);
            if (lane == 0) piv[c] = p;
            // true division: identical rows (duplicate samples, zero nugget) must cancel to exact zeros so
            // that a singular window is detected, as LAPACK's exact-zero pivot test does for the reference
            const double pval = mat[(size_t)c * LD + p];
            double f[E];
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int r = e * 32 + lane;
                f[e] = (r < k && r != p) ? mat[(size_t)c * LD + r] / pval : 0.0;
            }
            __syncwarp();
            for (int cc = c + 1; cc < k + 2; ++cc) {
                const double pv = mat[(size_t)cc * LD + p];
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int r = e * 32 + lane;
                    if (r < k && r != p) mat[(size_t)cc * LD + r] = fma(-f[e], pv, mat[(size_t)cc * LD + r]);
                }
            }
            __syncwarp();
        }
        // ---- back out x, mu; estimate and variance ----
        double res = quiet_nan<double>(), var = quiet_nan<double>();
        if (k == 1) {
            const double z0 = __shfl_sync(kFull, mz[0], 0), b0 = __shfl_sync(kFull, borig[0], 0);
            res = z0 * a.out_scale + a.out_shift;
            var = -(b0 + b0);
        } else if (!singular) {
            double y[E], u[E];
            double sy = 0.0, su = 0.0;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int c = e * 32 + lane;
                y[e] = u[e] = 0.0;
                if (c < k) {
                    const int p = piv[c];
                    const double dinv = 1.0 / mat[(size_t)c * LD + p];
                    y[e] = colB[p] * dinv;
                    u[e] = colO[p] * dinv;
                    sy += y[e];
                    su += u[e];
                }
            }
            sy = warp_sum(sy);
            su = warp_sum(su);
            const double mu = (sy - 1.0) / su;
            double zz = 0.0, xb = 0.0;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const double x = y[e] - mu * u[e];
                zz = fma(x, mz[e], zz);
                xb = fma(x, borig[e], xb);
            }
            zz = warp_sum(zz);
            xb = warp_sum(xb);
            res = zz * a.out_scale + a.out_shift;
            var = -(xb + mu);
            if (!isfinite(mu)) singular = true;
        }
        if (lane == 0) {
            if (singular) {
                res = quiet_nan<double>();
                var = quiet_nan<double>();
                if (a.n_singular) atomicAdd(a.n_singular, 1);
            }
            store_window(a, m, res, var);
        }
        __syncwarp();
    }
}

// k > 32: W warps (a "team") solve one target together, one row per lane (row = wsub*32 + lane).
// Same algorithm as above; the team synchronises once per elimination step on its own named
// barrier.  Doubling the warps per window doubles the parallelism available to hide the
// shared-memory round trips, at the same shared-memory footprint per window.
__device__ __forceinline__ void team_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

template <int W>
__global__ void __launch_bounds__(512) ok_solve_team_kernel(const OkArgs a, int teams_per_cta) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int LD = 32 * W;
    const int k = a.k;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int team = warp / W, wsub = warp % W;
    const int r = wsub * 32 + lane;  // my row (and my column in the final extraction)
    double P0, P1, P2;
    load_params(a, P0, P1, P2);
    const size_t per_team = (size_t)(k + 2) * LD * sizeof(double) + 2 * LD * sizeof(double) + LD * sizeof(int) +
                            2 * W * 16 + 4 * W * sizeof(double);
    unsigned char* base = smem_raw + (size_t)team * per_team;
    double* mat = reinterpret_cast<double*>(base);
    double* cx = mat + (size_t)(k + 2) * LD;
    double* cy = cx + LD;
    double* pbest = cy + LD;                            // [2][W] partial maxima (double-buffered by step parity)
    double* red = pbest + 2 * W;                        // [4][W] partial sums
    int* prow = reinterpret_cast<int*>(red + 4 * W);    // [2][W] partial arg-max rows
    int* piv = prow + 2 * W;                            // [LD]
    double* colB = mat + (size_t)k * LD;
    double* colO = mat + (size_t)(k + 1) * LD;
    const int bar_id = 1 + team, bar_n = 32 * W;

    const long long total_teams = (long long)gridDim.x * teams_per_cta;
    const long long n_windows = window_count(a);
    for (long long wi = (long long)blockIdx.x * teams_per_cta + team; wi < n_windows; wi += total_teams) {
        const long long m = window_target(a, wi);
        double mx = 0.0, my = 0.0, mz = 0.0, borig = 0.0;
        if (r < k) {
            int id = a.idx[m * k + r];
            if (id < 0 || id >= a.n) {
                id = a.idx[m * k];
                if (id < 0 || id >= a.n) id = 0;
            }
            mx = a.s_x[id];
            my = a.s_y[id];
            mz = a.z[id];
            const double bd = window_distance(a.metric, a.d2[m * k + r]);
            double b = -gamma_of(a.model, P0, P1, P2, bd);
            if (a.exact_values && fabs(bd) <= OK_EPS) b = 0.0;
            borig = b;
            cx[r] = mx;
            cy[r] = my;
            colB[r] = b;
            colO[r] = 1.0;
        }
        team_sync(bar_id, bar_n);
        if (r < k) {
            for (int j = 0; j < k; ++j) {
                const double d = sample_distance(a.metric, mx, my, cx[j], cy[j]);
                mat[(size_t)j * LD + r] = (r == j) ? 0.0 : -gamma_of(a.model, P0, P1, P2, d);
            }
        }
        bool done = false, singular = false;
        for (int c = 0; c < k; ++c) {
            double best = -1.0;
            int brow = 0x7fffffff;
            if (r < k && !done) {
                best = fabs(mat[(size_t)c * LD + r]);
                brow = r;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(kFull, best, o);
                const int orow = __shfl_xor_sync(kFull, brow, o);
                if (ob > best || (ob == best && orow < brow)) {
                    best = ob;
                    brow = orow;
                }
            }
            const int pb = (c & 1) * W;
            if (lane == 0) {
                pbest[pb + wsub] = best;
                prow[pb + wsub] = brow;
            }
            team_sync(bar_id, bar_n);  // step c-1 updates of every warp are complete and visible
#pragma unroll
            for (int w = 0; w < W; ++w) {
                const double ob = pbest[pb + w];
                const int orow = prow[pb + w];
                if (ob > best || (ob == best && orow < brow)) {
                    best = ob;
                    brow = orow;
                }
            }
            if (!(best > 0.0)) {
                singular = true;
                break;
            }
            const int p = brow;
            if (r == p) done = true;
            if (r == 0) piv[c] = p;
            const double pval = mat[(size_t)c * LD + p];
            const bool upd = r < k && r != p;
            const double f = upd ? mat[(size_t)c * LD + r] / pval : 0.0;
            if (upd) {
                for (int cc = c + 1; cc < k + 2; ++cc)
                    mat[(size_t)cc * LD + r] = fma(-f, mat[(size_t)cc * LD + p], mat[(size_t)cc * LD + r]);
            }
        }
        team_sync(bar_id, bar_n);
        double res = quiet_nan<double>(), var = quiet_nan<double>();
        if (!singular) {
            double y = 0.0, u = 0.0;
            if (r < k) {  // r plays the role of the column index here
                const int p = piv[r];
                const double dinv = 1.0 / mat[(size_t)r * LD + p];
                y = colB[p] * dinv;
                u = colO[p] * dinv;
            }
            const double sy_w = warp_sum(y), su_w = warp_sum(u);
            if (lane == 0) {
                red[0 * W + wsub] = sy_w;
                red[1 * W + wsub] = su_w;
            }
            team_sync(bar_id, bar_n);
            double sy = 0.0, su = 0.0;
#pragma unroll
            for (int w = 0; w < W; ++w) {
                sy += red[0 * W + w];
                su += red[1 * W + w];
            }
            const double mu = (sy - 1.0) / su;
            const double x = y - mu * u;
            const double zz_w = warp_sum(x * mz), xb_w = warp_sum(x * borig);
            if (lane == 0) {
                red[2 * W + wsub] = zz_w;
                red[3 * W + wsub] = xb_w;
            }
            team_sync(bar_id, bar_n);
            double zz = 0.0, xb = 0.0;
#pragma unroll
            for (int w = 0; w < W; ++w) {
                zz += red[2 * W + w];
                xb += red[3 * W + w];
            }
            res = zz * a.out_scale + a.out_shift;
            var = -(xb + mu);
            if (!isfinite(mu)) singular = true;
        }
        if (r == 0) {
            if (singular) {
                res = quiet_nan<double>();
                var = quiet_nan<double>();
                if (a.n_singular) atomicAdd(a.n_singular, 1);
            }
            store_window(a, m, res, var);
        }
        team_sync(bar_id, bar_n);  // shared memory of the team is reused by the next window
    }
}

// ------------------------------------------------------------------------------------------
// K4 for k <= 64: the window's matrix lives in REGISTERS, one row per thread.
//
// The shared-memory kernels above move every matrix element through shared memory twice per
// elimination step (k^3 loads + k^3 stores per window); at k = 64 that traffic, not the FP64 pipe,
// sets the pace.  Here thread r keeps row r of [b | 1 | G] in registers: a[0], a[1] are the two
// right-hand sides, a[2..] the not yet eliminated columns.  Each step the rows shift left by one
// (the update writes a[j-1] = a[j] - f * pivot[j]), so the active column is always a[2] and every
// register index is a compile-time constant although the step loop is a run-time loop; the code is
// specialised per phase of 8 steps (row length shrinks by 8 per phase), which keeps the wasted
// multiply-adds on dead tail columns to ~10 %.  Only the pivot row goes through shared memory: the
// best row of each warp publishes itself, one named barrier per step, all threads read it back by
// broadcast.  Same pivot rule (largest magnitude, lowest row on ties), same true division, same fma
// as the shared-memory kernels: results are bit-identical to them.
// FP64-pipe bound: ~ sum_c (k - c + 2) ~ k^2/2 fma per thread per window.
// ------------------------------------------------------------------------------------------
struct VarioConst {
    int model;
    double p0, p1, p2, c0, c1;
};
__device__ __forceinline__ VarioConst vario_const(int model, double p0, double p1, double p2) {
    VarioConst v{model, p0, p1, p2, 0.0, 0.0};
    if (model == CMX_VARIOGRAM_GAUSSIAN) {
        const double r = p1 * 4.0 / 7.0;
        v.c0 = 1.0 / (r * r);
    } else if (model == CMX_VARIOGRAM_SPHERICAL) {
        v.c0 = 3.0 / (2.0 * p1);
        v.c1 = 1.0 / (2.0 * p1 * p1 * p1);
    } else if (model == CMX_VARIOGRAM_EXPONENTIAL) {
        v.c0 = 1.0 / (p1 / 3.0);
    }
    return v;
}
// gamma_of with the divisions by constants hoisted (differs from it by rounding only)
__device__ __forceinline__ double gamma_pre(const VarioConst& v, double d) {
    switch (v.model) {
        case CMX_VARIOGRAM_LINEAR:
            return v.p0 * d + v.p1;
        case CMX_VARIOGRAM_POWER:
            return v.p0 * pow(d, v.p1) + v.p2;
        case CMX_VARIOGRAM_GAUSSIAN:
            return v.p0 * (1.0 - exp(-(d * d) * v.c0)) + v.p2;
        case CMX_VARIOGRAM_SPHERICAL:
            if (d <= v.p1) return v.p0 * (d * v.c0 - (d * d * d) * v.c1) + v.p2;
            return v.p0 + v.p2;
        default:
            return v.p0 * (1.0 - exp(-d * v.c0)) + v.p2;
    }
}

#ifndef CMX_ROWS_SPEC_RCP
#define CMX_ROWS_SPEC_RCP 1
#endif
struct RowState {
    bool done, singular;
    double diag;
    int mycol;
};

// Pivot search for one step + publication of this warp's candidate row (its reciprocal pivot in the spare slot).
// Magnitudes of non-negative doubles order like their bit patterns, so the arg-max is two REDUX and a ballot.
template <int KC, int L>
__device__ __forceinline__ void rows_publish(const double (&a)[KC + 2], const RowState& st, int c, int k, int r, int lane,
                                             int wsub, double cand, double rinv, double* prow,
                                             unsigned long long* pbest, int* pbrow) {
    constexpr int NW = (KC + 31) / 32;
    constexpr int LS = KC + 4;  // stride of a published row: KC + 2 values, 1 / pivot, pad
    const unsigned long long bits = (st.done || r >= k) ? 0ull : (unsigned long long)__double_as_longlong(fabs(cand));
    const unsigned hi = (unsigned)(bits >> 32), lo = (unsigned)bits;
    const unsigned mh = __reduce_max_sync(kFull, hi);
    const unsigned ml = __reduce_max_sync(kFull, hi == mh ? lo : 0u);
    const unsigned win = __ballot_sync(kFull, hi == mh && lo == ml);
    const int wl = __ffs(win) - 1;  // lowest row of the warp among equal maxima
    const int buf = c & 1;
    double* slot = prow + (size_t)(buf * NW + wsub) * LS;
    if (lane == wl) {
#pragma unroll
        for (int j = 0; j < L; j += 2) *reinterpret_cast<double2*>(slot + j) = make_double2(a[j], a[j + 1]);
        slot[KC + 2] = rinv;
        pbest[buf * NW + wsub] = ((unsigned long long)mh << 32) | ml;
        pbrow[buf * NW + wsub] = wsub * 32 + wl;
    }
}

// 8 elimination steps (fewer at the end of the window) on rows of current length L = 2 + remaining columns.
// On entry the candidates of step c0 are already published.  Inside a step the first multiply-add produces the
// next active column; the next pivot search and every thread's speculative reciprocal of its own candidate are
// issued right behind it, so their latency hides under the remaining multiply-adds of the step.
template <int KC, int L>
__device__ __forceinline__ void rows_phase(double (&a)[KC + 2], RowState& st, int c0, int k, int r, int lane, int wsub,
                                           double* prow, unsigned long long* pbest, int* pbrow, int bar_id) {
    constexpr int NW = (KC + 31) / 32;
    constexpr int LS = KC + 4;
#pragma unroll 1
    for (int c = c0; c < c0 + 8 && c < k; ++c) {
        if (NW > 1)
            asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(32 * NW) : "memory");
        else
            __syncwarp();
        const int buf = c & 1;
        unsigned long long best = pbest[buf * NW];
        int brow = pbrow[buf * NW];
#pragma unroll
        for (int w = 1; w < NW; ++w) {
            const unsigned long long ob = pbest[buf * NW + w];
            if (ob > best) {  // equal maxima: the lower row (earlier warp) stays
                best = ob;
                brow = pbrow[buf * NW + w];
            }
        }
        if (best == 0ull || best > 0x7ff0000000000000ull) {  // zero column (or NaN): singular window
            st.singular = true;
            return;
        }
        const double* prd = prow + (size_t)(buf * NW + (brow >> 5)) * LS;
        const double2* pr = reinterpret_cast<const double2*>(prd);
        const double2 p23 = pr[1];
        const double pval = p23.x, pinv = prd[KC + 2];
        const bool me = r == brow;
        if (me) {
            st.done = true;
            st.diag = a[2];
            st.mycol = c;
        }
        // f = a[2] / pval, correctly rounded from the published reciprocal (one Newton step on the quotient).
        // Identical rows (duplicate samples, zero nugget) must cancel to exact zeros so that a singular window is
        // detected, as LAPACK's exact-zero pivot test does for the reference: a[2] == pval gives exactly 1 here.
        double f;
#if CMX_ROWS_SPEC_RCP
        if (best > 0x0360000000000000ull && best < 0x7c90000000000000ull) {  // ~1e-290 < |pval| < ~1e290
            const double q = a[2] * pinv;
            f = fma(fma(-q, pval, a[2]), pinv, q);
        } else {
            f = a[2] / pval;
        }
#else
        f = a[2] / pval;
        (void)pinv;
#endif
        if (me) f = 0.0;
        const double next = fma(-f, p23.y, a[3]);  // my entry of the next active column
        const bool more = c + 1 < k;
        double rinv = 0.0;
#if CMX_ROWS_SPEC_RCP
        if (more) rinv = __drcp_rn(next);
#endif
        const double2 p01 = pr[0];
        a[0] = fma(-f, p01.x, a[0]);
        a[1] = fma(-f, p01.y, a[1]);
        a[2] = next;
#pragma unroll
        for (int j = 4; j < L; j += 2) {  // 16-byte broadcast loads of the published row
            const double2 q = pr[j >> 1];
            a[j - 1] = fma(-f, q.x, a[j]);
            a[j] = fma(-f, q.y, a[j + 1]);
        }
        a[L - 1] = 0.0;
        if (more) rows_publish<KC, L>(a, st, c + 1, k, r, lane, wsub, next, rinv, prow, pbest, pbrow);
    }
}

template <int KC, int PH>
__device__ __forceinline__ void rows_phases(double (&a)[KC + 2], RowState& st, int k, int r, int lane, int wsub,
                                            double* prow, unsigned long long* pbest, int* pbrow, int bar_id) {
    if constexpr (PH == 0) rows_publish<KC, KC + 2>(a, st, 0, k, r, lane, wsub, a[2], __drcp_rn(a[2]), prow, pbest, pbrow);
    if constexpr (PH * 8 < KC) {
        if (PH * 8 < k && !st.singular) {
            rows_phase<KC, KC + 2 - 8 * PH>(a, st, PH * 8, k, r, lane, wsub, prow, pbest, pbrow, bar_id);
            rows_phases<KC, PH + 1>(a, st, k, r, lane, wsub, prow, pbest, pbrow, bar_id);
        }
    }
}

template <int KC>
__host__ __device__ constexpr size_t rows_team_smem() {
    constexpr int NW = (KC + 31) / 32;
    return (size_t)KC * NW * 32 * sizeof(double) + (size_t)4 * KC * sizeof(double) + (size_t)2 * NW * (KC + 4) * sizeof(double) + 2 * NW * sizeof(unsigned long long) +
           4 * NW * sizeof(double) + 2 * NW * sizeof(int) + 8;
}

template <int KC, int TEAMS>
__global__ void __launch_bounds__(TEAMS * ((KC + 31) / 32) * 32, KC > 32 ? 3 : 4) ok_solve_rows_kernel(const OkArgs a_) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NW = (KC + 31) / 32;
    const OkArgs& A = a_;
    const int k = A.k;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int team = warp / NW, wsub = warp % NW;
    const int r = wsub * 32 + lane;
    unsigned char* base = smem_raw + (size_t)team * ((rows_team_smem<KC>() + 15) & ~size_t(15));
    constexpr int TH = NW * 32;
    double* gmat = reinterpret_cast<double*>(base);  // [KC][TH]
    double* cx = gmat + (size_t)KC * TH;
    double* cy = cx + KC;
    double* cz = cy + KC;
    double* cb = cz + KC;
    double* prow = cb + KC;                                             // [2][NW][KC + 4]
    unsigned long long* pbest = reinterpret_cast<unsigned long long*>(prow + 2 * NW * (KC + 4));  // [2][NW]
    double* red = reinterpret_cast<double*>(pbest + 2 * NW);           // [4][NW]
    int* pbrow = reinterpret_cast<int*>(red + 4 * NW);                  // [2][NW]
    const int bar_id = 1 + team;
    auto team_barrier = [&]() {
        if (NW > 1)
            asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(32 * NW) : "memory");
        else
            __syncwarp();
    };
    double P0, P1, P2;
    load_params(A, P0, P1, P2);
    const VarioConst vc = vario_const(A.model, P0, P1, P2);

    const long long total_teams = (long long)gridDim.x * TEAMS;
    const long long n_windows = window_count(A);
    for (long long wi = (long long)blockIdx.x * TEAMS + team; wi < n_windows; wi += total_teams) {
        const long long m = window_target(A, wi);
        double mx = 0.0, my = 0.0, mz = 0.0, borig = 0.0;
        if (r < k) {
            int id = A.idx[m * k + r];
            if (id < 0 || id >= A.n) {  // short neighbour list: duplicate row 0 -> singular window -> NaN
                id = A.idx[m * k];
                if (id < 0 || id >= A.n) id = 0;
            }
            mx = A.s_x[id];
            my = A.s_y[id];
            mz = A.z[id];
            const double bd = window_distance(A.metric, A.d2[m * k + r]);
            double b = -gamma_pre(vc, bd);
            if (A.exact_values && fabs(bd) <= OK_EPS) b = 0.0;
            borig = b;
        }
        if (r < KC) {
            cx[r] = mx;
            cy[r] = my;
            cz[r] = mz;
            cb[r] = borig;
        }
        team_barrier();
        // ---- G through shared memory once, [column][row]: G is symmetric, so thread r evaluates only the
        // pairs (r, r+1..r+k/2 mod k) and stores each value at both places (strides of TH+1: conflict-free) ----
        if (r < k) {
            gmat[(size_t)r * TH + r] = 0.0;
            const int half = k >> 1;
            int j = r;
            for (int t = 1; t <= half; ++t) {
                j = j + 1 == k ? 0 : j + 1;
                const double g = -gamma_pre(vc, sample_distance(A.metric, mx, my, cx[j], cy[j]));
                gmat[(size_t)j * TH + r] = g;
                gmat[(size_t)r * TH + j] = g;
            }
        }
        team_barrier();
        // ---- my row of [b | 1 | G] into registers ----
        double a[KC + 2];
        a[0] = borig;
        a[1] = r < k ? 1.0 : 0.0;
#pragma unroll
        for (int j = 0; j < KC; ++j) a[2 + j] = (r < k && j < k) ? gmat[(size_t)j * TH + r] : 0.0;
        RowState st{false, false, 1.0, 0};
        if (k > 1) rows_phases<KC, 0>(a, st, k, r, lane, wsub, prow, pbest, pbrow, bar_id);

        // ---- back out x, mu; estimate and variance (thread r holds the row that was pivot of column mycol) ----
        double res = quiet_nan<double>(), var = quiet_nan<double>();
        bool singular = st.singular;
        if (k == 1) {
            // the core is the 1x1 zero matrix while the bordered system [[0,1],[1,0]] is regular: x = 1, mu = b
            res = cz[0] * A.out_scale + A.out_shift;
            var = -(cb[0] + cb[0]);
        } else if (!singular) {
            double y = 0.0, u = 0.0;
            if (r < k) {
                const double dinv = 1.0 / st.diag;
                y = a[0] * dinv;
                u = a[1] * dinv;
            }
            double sy = warp_sum(y), su = warp_sum(u);
            if (NW > 1) {
                if (lane == 0) {
                    red[0 * NW + wsub] = sy;
                    red[1 * NW + wsub] = su;
                }
                team_barrier();
                sy = su = 0.0;
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    sy += red[0 * NW + w];
                    su += red[1 * NW + w];
                }
            }
            const double mu = (sy - 1.0) / su;
            const double x = y - mu * u;
            double zz = 0.0, xb = 0.0;
            if (r < k) {
                zz = x * cz[st.mycol];
                xb = x * cb[st.mycol];
            }
            zz = warp_sum(zz);
            xb = warp_sum(xb);
            if (NW > 1) {
                if (lane == 0) {
                    red[2 * NW + wsub] = zz;
                    red[3 * NW + wsub] = xb;
                }
                team_barrier();
                zz = xb = 0.0;
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    zz += red[2 * NW + w];
                    xb += red[3 * NW + w];
                }
            }
            res = zz * A.out_scale + A.out_shift;
            var = -(xb + mu);
            if (!isfinite(mu)) singular = true;
        }
        if (r == 0) {
            if (singular) {
                res = quiet_nan<double>();
                var = quiet_nan<double>();
                if (A.n_singular) atomicAdd(A.n_singular, 1);
            }
            store_window(A, m, res, var);
        }
        team_barrier();  // the team's shared memory is reused by the next window
    }
}

}  // namespace
}  // namespace cmx

#include "ok_chol.cuh"
#include "ok_chol2.cuh"
#ifndef CMX_CHOL_PAIRS
#define CMX_CHOL_PAIRS 1  // 1: two columns per step (ok_chol2.cuh); 0: one column per step (ok_chol.cuh)
#endif

namespace cmx {
namespace {

// grid of a pivoted kernel: all windows, or - on the hand-over list, whose length only the device knows - a
// modest persistent grid (the list is normally empty or short)
inline long long pivoted_ctas(const OkArgs& a, int windows_per_cta, int ctas_per_sm) {
    const long long cap = (long long)sm_count() * ctas_per_sm;
    if (a.list) return cap < 2LL * sm_count() ? cap : 2LL * sm_count();
    const long long ctas = (a.M + windows_per_cta - 1) / windows_per_cta;
    return ctas > cap ? cap : ctas;
}

template <int KC>
int launch_ok_rows(const OkArgs& a, cudaStream_t stream) {
    constexpr int NW = (KC + 31) / 32;
    constexpr int TEAMS = NW == 1 ? 4 : 2;  // 128 threads per CTA
    const size_t smem = TEAMS * ((rows_team_smem<KC>() + 15) & ~size_t(15));
    CMX_CUDA(cudaFuncSetAttribute(ok_solve_rows_kernel<KC, TEAMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long ctas = pivoted_ctas(a, TEAMS, 8);
    ok_solve_rows_kernel<KC, TEAMS><<<(unsigned)ctas, TEAMS * NW * 32, smem, stream>>>(a);
    note_launch();
    return check_launch();
}

template <int W>
int launch_ok_team(const OkArgs& a, cudaStream_t stream) {
    const int k = a.k;
    const size_t per_team = (size_t)(k + 2) * 32 * W * sizeof(double) + 2 * 32 * W * sizeof(double) + 32 * W * sizeof(int) +
                            2 * W * 16 + 4 * W * sizeof(double);
    int teams = (int)((200 * 1024) / per_team);
    const int max_teams = 512 / (32 * W) < 15 ? 512 / (32 * W) : 15;  // <= 16 warps per CTA, <= 15 named barriers
    teams = teams > max_teams ? max_teams : (teams < 1 ? 1 : teams);
    const size_t smem = per_team * teams;
    CMX_CUDA(cudaFuncSetAttribute(ok_solve_team_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long ctas = pivoted_ctas(a, teams, 4);
    ok_solve_team_kernel<W><<<(unsigned)ctas, teams * W * 32, smem, stream>>>(a, teams);
    note_launch();
    return check_launch();
}

template <int E>
int launch_ok(const OkArgs& a, cudaStream_t stream) {
    const int k = a.k;
    const size_t per_warp = (size_t)(k + 2) * 32 * E * sizeof(double) + 2 * 32 * E * sizeof(double) + 32 * E * sizeof(int);
    int warps = (int)((200 * 1024) / per_warp);
    warps = warps > 8 ? 8 : (warps < 1 ? 1 : warps);
    const size_t smem = per_warp * warps;
    CMX_CUDA(cudaFuncSetAttribute(ok_solve_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long ctas = pivoted_ctas(a, warps, 8);
    ok_solve_kernel<E><<<(unsigned)ctas, warps * 32, smem, stream>>>(a, warps);
    note_launch();
    return check_launch();
}

// Gauss-Jordan with partial pivoting: all windows of `a`, or the windows on a.list
int launch_pivoted(const OkArgs& a, cudaStream_t stream) {
    const int k = a.k;
#ifndef CMX_OK_SMEM_ONLY
    if (k <= 16) return launch_ok_rows<16>(a, stream);
    if (k <= 32) return launch_ok_rows<32>(a, stream);
    if (k <= 48) return launch_ok_rows<48>(a, stream);
    if (k <= 64) return launch_ok_rows<64>(a, stream);
#endif
    if (k <= 32) return launch_ok<1>(a, stream);
    if (k <= 64) return launch_ok_team<2>(a, stream);
    return launch_ok_team<4>(a, stream);
}

inline size_t chol_scratch_bytes(long long M) { return ((size_t)(M > 0 ? M : 0) * sizeof(int) + 256 + 255) & ~size_t(255); }

// K4 on an existing neighbour table.  scratch (chol_scratch_bytes(M), may be null): hand-over list of the Cholesky path.
int solve_table(const double* s_x, const double* s_y, const double* z, long long n, const int* idx, const double* d2,
                long long M, int metric, int k, int model, const double* params, int params_on_device, int exact_values,
                double out_scale, double out_shift, void* out, void* sigmasq, int out_is_f32, int32_t* n_singular,
                void* scratch, size_t scratch_bytes, cudaStream_t stream) {
    if (M == 0) return CMX_OK;
    if (M > 0x7fffffffLL) return CMX_ERR_TOO_MANY;
    OkArgs a;
    a.s_x = s_x;
    a.s_y = s_y;
    a.z = z;
    a.out_is_f32 = out_is_f32;
    a.idx = idx;
    a.d2 = d2;
    a.M = M;
    a.n = n;
    a.k = k;
    a.model = model;
    a.params_dev = params_on_device ? params : nullptr;
    a.p0 = params_on_device ? 0.0 : params[0];
    a.p1 = params_on_device ? 0.0 : params[1];
    a.p2 = params_on_device ? 0.0 : params[2];
    a.exact_values = exact_values;
    a.metric = metric;
    a.out_scale = out_scale;
    a.out_shift = out_shift;
    a.out = out;
    a.sigmasq = sigmasq;
    a.n_singular = n_singular;
    a.list = nullptr;
    a.list_count = nullptr;
    timing_begin(CMX_KERNEL_OK_SOLVE, stream);
    int rc;
#ifndef CMX_OK_NO_CHOL
    if (k >= 2 && k <= 64 && scratch && scratch_bytes >= chol_scratch_bytes(M)) {
        unsigned char* sc = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(scratch) + 255) & ~uintptr_t(255));
        if (sc + 256 + (size_t)M * sizeof(int) > static_cast<unsigned char*>(scratch) + scratch_bytes) return CMX_ERR_WORKSPACE;
        int* fail_count = reinterpret_cast<int*>(sc);
        int* fail_list = reinterpret_cast<int*>(sc + 256);
        CMX_CUDA(cudaMemsetAsync(fail_count, 0, sizeof(int), stream));
        // k <= 48: one column per step; above: two (measured: 13.8 vs 15.1 ms at k = 48, 29.1 vs 28.6 ms at k = 64 per
        // 1.04 M windows with the same assembly; the pair kernel then gained the per-model register assembly)
        if (k <= 16) rc = launch_ok_chol<2>(a, fail_list, fail_count, stream);
        else if (k <= 32) rc = launch_ok_chol<4>(a, fail_list, fail_count, stream);
        else if (k <= 48) rc = launch_ok_chol<6>(a, fail_list, fail_count, stream);
#if CMX_CHOL_PAIRS
        else rc = launch_ok_chol2<8>(a, fail_list, fail_count, stream);
#else
        else rc = launch_ok_chol<8>(a, fail_list, fail_count, stream);
#endif
        if (rc == CMX_OK) {
            a.list = fail_list;
            a.list_count = fail_count;
            rc = launch_pivoted(a, stream);
        }
    } else
#endif
    {
        rc = launch_pivoted(a, stream);
    }
    timing_end(CMX_KERNEL_OK_SOLVE, stream);
    return rc;
}

int solve_entry(const double* s_x, const double* s_y, const double* z, int64_t n, const int32_t* idx, const double* d2,
                int64_t M, int metric, int k, int model, const double* params, int exact_values, double out_scale,
                double out_shift, void* out, void* sigmasq, int out_is_f32, int32_t* n_singular, void* ws,
                size_t ws_bytes, const cmx_options* options, void* stream) {
    int params_on_device = 0;
    const int orc = read_options(options, nullptr, nullptr, &params_on_device);
    if (orc != CMX_OK) return orc;
    if (!s_x || !s_y || !z || !idx || !d2 || !params || !out || n < 1 || M < 0 || k < 1) return CMX_ERR_BAD_ARG;
    if (k > CMX_MAX_K) return CMX_ERR_K_TOO_LARGE;
    if (model < CMX_VARIOGRAM_LINEAR || model > CMX_VARIOGRAM_EXPONENTIAL) return CMX_ERR_UNSUPPORTED;
    if (metric != CMX_METRIC_EUCLIDEAN && metric != CMX_METRIC_GEOGRAPHIC) return CMX_ERR_UNSUPPORTED;
    return solve_table(s_x, s_y, z, n, idx, d2, M, metric, k, model, params, params_on_device, exact_values, out_scale,
                       out_shift, out, sigmasq, out_is_f32, n_singular, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

int ok_entry(const double* s_x, const double* s_y, const double* z, int64_t n, const double* t_x, int64_t n_x,
             const double* t_y, int64_t n_y, int grid, int metric, int k, int model, const double* params,
             int exact_values, double out_scale, double out_shift, void* out, void* sigmasq, int out_is_f32,
             int32_t* n_singular, void* workspace, size_t workspace_bytes, const cmx_options* options, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    int search = CMX_SEARCH_AUTO, params_on_device = 0;
    const int orc = read_options(options, &search, nullptr, &params_on_device);
    if (orc != CMX_OK) return orc;
    if (!s_x || !s_y || !z || !t_x || !t_y || !params || !out) return CMX_ERR_BAD_ARG;
    if (model < CMX_VARIOGRAM_LINEAR || model > CMX_VARIOGRAM_EXPONENTIAL) return CMX_ERR_UNSUPPORTED;
    if (metric != CMX_METRIC_EUCLIDEAN) return CMX_ERR_UNSUPPORTED;  // geographic: cmx_knn3 + cmx_ok_solve_*
    if (k < 1) return CMX_ERR_BAD_ARG;
    if (k > CMX_MAX_K) return CMX_ERR_K_TOO_LARGE;
    if (n_x < 0 || n_y < 0) return CMX_ERR_BAD_ARG;
    const long long M = grid ? (long long)n_x * n_y : (long long)n_x;
    // workspace: [neighbour table | hand-over list | kNN scratch]
    if (!workspace || workspace_bytes < cmx_ok_workspace_bytes(k, M)) return CMX_ERR_WORKSPACE;
    const size_t tab_bytes = ((size_t)M * k * (sizeof(double) + sizeof(int)) + 255) & ~size_t(255);
    const size_t list_bytes = chol_scratch_bytes(M);
    unsigned char* extra = static_cast<unsigned char*>(workspace);
    extra = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(extra) + 255) & ~uintptr_t(255));
    double* d2_tab = reinterpret_cast<double*>(extra);
    int* idx_tab = reinterpret_cast<int*>(extra + (size_t)M * k * sizeof(double));
    unsigned char* list_scratch = extra + tab_bytes;

    // neighbours in the kriging frame: rows of a dense target grid run along y (lat), columns along x (lon)
    KnnJob job{};
    job.s_a = s_y;
    job.s_b = s_x;
    job.n = n;
    job.t_a = t_y;
    job.t_b = t_x;
    job.n_a = grid ? n_y : n_x;
    job.n_b = n_x;
    job.grid = grid;
    job.k = k;
    job.idx_out = idx_tab;
    job.d2_out = d2_tab;
    job.search = search;
    job.workspace = list_scratch + list_bytes;
    job.workspace_bytes = workspace_bytes - (size_t)((list_scratch + list_bytes) - static_cast<unsigned char*>(workspace));
    int rc = run_knn(job, stream);
    if (rc != CMX_OK) return rc;
    if (M == 0) return CMX_OK;

    return solve_table(s_x, s_y, z, n, idx_tab, d2_tab, M, CMX_METRIC_EUCLIDEAN, k, model, params, params_on_device,
                       exact_values, out_scale, out_shift, out, sigmasq, out_is_f32, n_singular, list_scratch, list_bytes,
                       stream);
}

}  // namespace
}  // namespace cmx

extern "C" {

size_t cmx_ok_solve_workspace_bytes(int64_t n_targets) { return cmx::chol_scratch_bytes(n_targets) + 256; }

size_t cmx_ok_workspace_bytes(int k, int64_t n_targets) {
    if (k < 1 || k > CMX_MAX_K || n_targets < 0) return 0;
    return cmx::knn_workspace_bytes(k) + (size_t)n_targets * (size_t)k * (sizeof(double) + sizeof(int)) +
           cmx::chol_scratch_bytes(n_targets) + 2048;
}

int cmx_ok_mw_f64(const double* s_x, const double* s_y, const double* z, int64_t n, const double* t_x, int64_t n_x,
                  const double* t_y, int64_t n_y, int grid, int metric, int k, int model, const double params[3],
                  int exact_values, double out_scale, double out_shift, double* out, double* sigmasq,
                  int32_t* n_singular, void* ws, size_t ws_bytes, const cmx_options* options, void* stream) {
    return cmx::ok_entry(s_x, s_y, z, n, t_x, n_x, t_y, n_y, grid, metric, k, model, params, exact_values, out_scale,
                         out_shift, out, sigmasq, 0, n_singular, ws, ws_bytes, options, stream);
}
int cmx_ok_mw_f32(const double* s_x, const double* s_y, const double* z, int64_t n, const double* t_x, int64_t n_x,
                  const double* t_y, int64_t n_y, int grid, int metric, int k, int model, const double params[3],
                  int exact_values, double out_scale, double out_shift, float* out, float* sigmasq,
                  int32_t* n_singular, void* ws, size_t ws_bytes, const cmx_options* options, void* stream) {
    return cmx::ok_entry(s_x, s_y, z, n, t_x, n_x, t_y, n_y, grid, metric, k, model, params, exact_values, out_scale,
                         out_shift, out, sigmasq, 1, n_singular, ws, ws_bytes, options, stream);
}
int cmx_ok_solve_f64(const double* s_x, const double* s_y, const double* z, int64_t n, const int32_t* idx,
                     const double* d2, int64_t M, int metric, int k, int model, const double params[3],
                     int exact_values, double out_scale, double out_shift, double* out, double* sigmasq,
                     int32_t* n_singular, void* ws, size_t ws_bytes, const cmx_options* options, void* stream) {
    return cmx::solve_entry(s_x, s_y, z, n, idx, d2, M, metric, k, model, params, exact_values, out_scale, out_shift,
                            out, sigmasq, 0, n_singular, ws, ws_bytes, options, stream);
}
int cmx_ok_solve_f32(const double* s_x, const double* s_y, const double* z, int64_t n, const int32_t* idx,
                     const double* d2, int64_t M, int metric, int k, int model, const double params[3],
                     int exact_values, double out_scale, double out_shift, float* out, float* sigmasq,
                     int32_t* n_singular, void* ws, size_t ws_bytes, const cmx_options* options, void* stream) {
    return cmx::solve_entry(s_x, s_y, z, n, idx, d2, M, metric, k, model, params, exact_values, out_scale, out_shift,
                            out, sigmasq, 1, n_singular, ws, ws_bytes, options, stream);
}

}  // extern "C"
