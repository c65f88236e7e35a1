// K4, fast path: moving-window ordinary kriging by a register-resident Cholesky factorisation.
// Included by kriging.cu (uses OkArgs, VarioConst, gamma_pre, window_distance, sample_distance).
//
// Reference: pykrige _exec_loop_moving_window (call site src/climatrix/reconstruct/kriging.py:236-241)
// solves, per target,   [ A  1 ] [x ]   [ b ]     A_ij = -gamma(|s_i - s_j|), A_ii = 0
//                       [ 1' 0 ] [mu] = [ 1 ]     b_j  = -gamma(|t - s_j|)  (0 where the distance <= 1e-10)
// by LU with partial pivoting.  A is symmetric INDEFINITE, but with 1'x = 1 the system is unchanged by
// C = A + c0 11' (then mu = mu' + c0), and for c0 above a threshold C is positive definite because -gamma
// is conditionally positive definite.  With c0 = 2 max gamma_ij of the window:
//
//     C = L L',  y' = L^-1 b,  u' = L^-1 1,  w' = L^-1 Z
//     mu' = (u'.y' - 1) / (u'.u'),  z = w'.y' - mu' w'.u',  sigma^2 = -(y'.y' - mu' u'.y' + mu' + c0)
//
// y', u', w' fall out of the factorisation itself when b, 1 and Z are appended to C as three extra ROWS
// (right-looking Cholesky of the augmented lower triangle), so no triangular solve follows: k^3/6 + 3k^2/2
// multiply-adds per window instead of the k^3/2 of the pivoted Gauss-Jordan kernel, no pivot search and no
// per-step division chain.  Windows whose pivots do not stay safely positive (near-duplicate samples with
// zero nugget, gaussian model without nugget, exponent ~2 power model: cond(C) > ~1e9) are appended to a
// list and redone by the pivoted kernel (ok_solve_rows_kernel / ok_solve_team_kernel), which also owns
// the exact-singularity semantics (NaN + n_singular).
//
// Mapping: one warp per window, no shared-memory matrix.  The lower triangle is distributed 2-D block-
// cyclically over an 8 x 4 lane grid: lane (r, c) owns elements (8a + r, 4b + c), a < NB + 1, b < 2 NB,
// b <= 2a + 1, in registers t[a][b] (KC = 8 NB is the neighbour capacity; row block NB holds the three
// extra rows).  At step j the owners of column j scale it by rsqrt(pivot) and publish it to shared
// memory twice - ordered for the lanes that need it as "row factor" L[8a + r][j] and as "column factor"
// L[4b + c][j] - and every lane updates its tile with 16-byte loads: ~(NB+1) * 2NB / 2 FMA for NB + 2NB
// loaded values.  Entries of finished rows / columns are published as zeros, so the update needs no
// predicates.  The publication of column j + 1 (pivot broadcast, rsqrt, stores) is issued right after
// the first block column of step j's update so that its latency hides under the rest of the update.
// All loops are fully unrolled: every register index is a compile-time constant.
#pragma once

namespace cmx {
namespace {

// The fully unrolled factorisation is 100-200 KB of code, far beyond the instruction caches (B300_MICROARCH.md: L0 ~6 KB,
// L1.5 32 KB).  With independent warps every warp streams that code from L2 on its own and instruction fetch becomes
// a visible cost at KC = 64 (ncu: no_instruction up to 1.3 stalls per issue).  CMX_CHOL_SYNC = 1 (tuning builds): ONE CTA
// per SM, as many warps as the register file holds, and a CTA-wide barrier at every phase of four steps, so that all
// warps of the SM walk through the code together.  Measured on B200 (profiles/r02_k4_variants.txt): 3-15 % SLOWER than
// independent warps at every capacity - the stalls are per-warp fetch latency, not shared fetch bandwidth - so it is off.
#ifndef CMX_CHOL_SYNC
#define CMX_CHOL_SYNC 0
#endif
// resident CTAs per SM / warps per CTA the kernel is compiled for (255 registers at KC = 48 / 64, 128 at KC = 32, 80 at 16)
constexpr int chol_ctas_per_sm(int nb) { return CMX_CHOL_SYNC ? 1 : (nb >= 6 ? 2 : (nb >= 4 ? 4 : 6)); }
constexpr int chol_warps(int nb) { return CMX_CHOL_SYNC ? (nb >= 6 ? 8 : (nb >= 4 ? 16 : 24)) : 4; }
__device__ __forceinline__ void chol_phase_sync() {
#if CMX_CHOL_SYNC
    __syncthreads();
#endif
}

template <int NB>
struct CholCfg {
    static constexpr int KC = 8 * NB;        // neighbour capacity
    static constexpr int NBC = 2 * NB;       // column blocks of 4
    static constexpr int NBR = NB + 1;       // row blocks of 8 (the last one: b, 1, Z rows)
    // strides (in doubles) of the published column, chosen so that the 8 (4) distinct 16-byte loads of a
    // warp fall into distinct banks: ROWP = 2 * odd (mod 16), COLP * 2 banks apart for the 4 column groups
    static constexpr int ROWP = NBR <= 6 ? 6 : 10;
    static constexpr int COLP = NBC + 2;
    static constexpr int TRI = KC * (KC - 1) / 2;  // gamma of every pair, strict lower triangle packed by rows
    static constexpr int WARP_DOUBLES = 2 * KC + 2 * KC + 2 * 8 * ROWP + 2 * 4 * COLP + 3 * KC + TRI + (TRI & 1);
    static constexpr int WARPS = chol_warps(NB);
    static constexpr size_t SMEM = (size_t)WARPS * WARP_DOUBLES * sizeof(double);
};

constexpr double CHOL_MIN_PIVOT = 1.0e-9;  // relative to c0: below it the window goes to the pivoted kernel

// gamma(d) with the model fixed at compile time (same expressions as gamma_pre)
template <int MODEL>
__device__ __forceinline__ double gamma_model(const VarioConst& v, double d) {
    if (MODEL == CMX_VARIOGRAM_LINEAR) return v.p0 * d + v.p1;
    if (MODEL == CMX_VARIOGRAM_POWER) return v.p0 * pow(d, v.p1) + v.p2;
    if (MODEL == CMX_VARIOGRAM_GAUSSIAN) return v.p0 * (1.0 - exp(-(d * d) * v.c0)) + v.p2;
    if (MODEL == CMX_VARIOGRAM_SPHERICAL) return d <= v.p1 ? v.p0 * (d * v.c0 - (d * d * d) * v.c1) + v.p2 : v.p0 + v.p2;
    return v.p0 * (1.0 - exp(-d * v.c0)) + v.p2;
}

// 1 / sqrt(d) for a positive normal d without the slow-path branch of rsqrt(): hardware seed + two Newton steps
__device__ __forceinline__ double rsqrt_pos(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double h = 0.5 * d;
    y = y * fma(-h * y, y, 1.5);
    y = y * fma(-h * y, y, 1.5);
    return y;
}

// gamma of a Euclidean pair from its SQUARED distance: no square root for the Gaussian model, the spherical
// polynomial folded to two multiply-adds; d2 = 0 gives exactly gamma(0) (the clamp only feeds the reciprocal root)
template <int MODEL>
__device__ __forceinline__ double gamma_from_d2(const VarioConst& v, double d2, double sA, double sB, double r2) {
    if (MODEL == CMX_VARIOGRAM_GAUSSIAN) return v.p0 * (1.0 - exp(-d2 * v.c0)) + v.p2;
    const double d = d2 * rsqrt_pos(fmax(d2, 1.0e-290));
    if (MODEL == CMX_VARIOGRAM_SPHERICAL) return d2 <= r2 ? fma(d, fma(-sB, d2, sA), v.p2) : v.p0 + v.p2;
    return gamma_model<MODEL>(v, d);
}

// gamma of every sample pair (i, m), m < i < k, into tri[i (i - 1) / 2 + m]; returns this lane's largest value.
// Rows p and k - p are folded into one sweep of k elements, so every 32-lane pass is full (k = 64: 64 passes for
// 2016 pairs) and the index arithmetic is a few selects.  Used for the geographic metric and for k > 48; smaller
// Euclidean windows are assembled straight into registers (chol_assemble_direct).
template <int MODEL, int METRIC, int PASSES = 2>
__device__ __forceinline__ double chol_pair_gammas(const VarioConst& vc, const double2* __restrict__ xy, double* __restrict__ tri,
                                                   int k, int lane) {
    constexpr int U = METRIC == CMX_METRIC_EUCLIDEAN ? 8 / PASSES : 2;
    double gm[U][PASSES];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int q = 0; q < PASSES; ++q) gm[u][q] = 0.0;
    const double sA = vc.p0 * vc.c0, sB = vc.p0 * vc.c1, r2 = vc.p1 * vc.p1;
    const int half = k >> 1;
    for (int p0 = 1; p0 <= half; p0 += U) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int p = p0 + u;
            const int ia = p, ib = k - p;                         // row ia: columns 0 .. p-1; row ib: columns 0 .. k-p-1
            const int count = p > half ? 0 : (ia == ib ? p : k);  // k even: the middle row stands alone
            const int offa = ia * (ia - 1) / 2, offb = ib * (ib - 1) / 2 - p;
#pragma unroll
            for (int q = 0; q < PASSES; ++q) {
                const int e = lane + 32 * q;
                const bool live = e < count, first = e < p;
                const double2 pi = xy[live ? (first ? ia : ib) : 0];
                const double2 pm = xy[live ? (first ? e : e - p) : 0];
                double g;
                if (METRIC == CMX_METRIC_EUCLIDEAN) {
                    const double dx = pi.x - pm.x, dy = pi.y - pm.y;
                    g = gamma_from_d2<MODEL>(vc, fma(dy, dy, dx * dx), sA, sB, r2);
                } else {
                    g = gamma_model<MODEL>(vc, sample_distance(METRIC, pi.x, pi.y, pm.x, pm.y));
                }
                if (live) tri[(first ? offa : offb) + e] = g;
                gm[u][q] = fmax(gm[u][q], g);  // idle slots evaluate gamma(0), never above a real pair's value
            }
        }
    }
    double gmax = 0.0;
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int q = 0; q < PASSES; ++q) gmax = fmax(gmax, gm[u][q]);
    return gmax;
}

// Euclidean metric: every lane evaluates gamma for exactly the elements it owns, straight into its registers -
// no staging, no index arithmetic, no divergence (entries of the upper triangle and of padding rows are computed
// and zeroed by a select).  The model is a template parameter, so one instantiation holds NB (NB + 1) copies of ONE
// variogram formula.  sqrt(d2) = d2 * rsqrt(d2) by hardware seed + two Newton steps (1-2 ulp; the kriging weights
// need ~1e-12).  Returns the lane's largest gamma.
template <int NB, int MODEL>
__device__ __forceinline__ double chol_assemble_direct(const VarioConst& vc, const double2* __restrict__ xy, int k, int r,
                                                       int c, double (&t)[NB + 1][2 * NB]) {
    double gmax = 0.0;
    const double sA = vc.p0 * vc.c0, sB = vc.p0 * vc.c1, r2 = vc.p1 * vc.p1;
#pragma unroll
    for (int a = 0; a < NB; ++a) {
        if (a % 2 == 0) chol_phase_sync();
        const int i = 8 * a + r;
        const double2 pi = xy[i];
#pragma unroll
        for (int b = 0; b < 2 * NB; ++b) {
            if (b <= 2 * a + 1) {
                const int mm = 4 * b + c;
                const double2 pm = xy[mm];
                const double dx = pi.x - pm.x, dy = pi.y - pm.y;
                double g = gamma_from_d2<MODEL>(vc, fma(dy, dy, dx * dx), sA, sB, r2);
                g = (mm < i && i < k) ? g : 0.0;
                t[a][b] = g;
                gmax = fmax(gmax, g);
            }
        }
    }
    return gmax;
}

template <int NB>
__global__ void __launch_bounds__(chol_warps(NB) * 32, chol_ctas_per_sm(NB))
ok_chol_kernel(const OkArgs A, int* __restrict__ fail_list, int* __restrict__ fail_count) {
    using Cfg = CholCfg<NB>;
    constexpr int KC = Cfg::KC, NBC = Cfg::NBC, NBR = Cfg::NBR, ROWP = Cfg::ROWP, COLP = Cfg::COLP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = lane >> 2, c = lane & 3;
    double* base = reinterpret_cast<double*>(smem_raw) + (size_t)warp * Cfg::WARP_DOUBLES;
    double2* xy = reinterpret_cast<double2*>(base);        // [KC] sample (x, y)
    double2* bz = xy + KC;                                  // [KC] (b, z)
    double* rowbuf = reinterpret_cast<double*>(bz + KC);    // [2][8][ROWP]  L[8a + r][j] at [r][a]
    double* colbuf = rowbuf + 2 * 8 * ROWP;                 // [2][4][COLP]  L[4(2a + h) + c][j] at [c][h * NB + a]
    double* ext = colbuf + 2 * 4 * COLP;                    // [3][KC]       y', u', w'
    double* tri = ext + 3 * KC;                             // [TRI]         gamma(i, m), m < i, at i (i - 1) / 2 + m
    const int k = A.k;
    const int k4 = (k + 3) & ~3;
    double P0, P1, P2;
    load_params(A, P0, P1, P2);
    const VarioConst vc = vario_const(A.model, P0, P1, P2);

    // the trip count is the same for every warp of the CTA (phase barriers inside): a warp beyond the last window
    // recomputes window M - 1 and stores nothing
    const long long nwarps = (long long)gridDim.x * Cfg::WARPS;
    for (long long base_m = (long long)blockIdx.x * Cfg::WARPS; base_m < A.M; base_m += nwarps) {
        const bool active = base_m + warp < A.M;
        const long long m = active ? base_m + warp : A.M - 1;
        chol_phase_sync();
        // ---- gather the window: coordinates, b, z ----
        bool bad = false;
#pragma unroll
        for (int i0 = 0; i0 < KC; i0 += 32) {
            const int i = i0 + lane;
            double x = 0.0, y = 0.0, b = 0.0, z = 0.0;
            if (i < k) {
                int id = A.idx[m * k + i];
                if (id < 0 || id >= A.n) {  // short neighbour list: the pivoted kernel reports it
                    bad = true;
                    id = 0;
                }
                x = A.s_x[id];
                y = A.s_y[id];
                z = A.z[id];
                const double bd = window_distance(A.metric, A.d2[m * k + i]);
                b = -gamma_pre(vc, bd);
                if (A.exact_values && fabs(bd) <= OK_EPS) b = 0.0;
            }
            if (i < KC) {
                xy[i] = make_double2(x, y);
                bz[i] = make_double2(b, z);
            }
        }
        bad = __any_sync(kFull, bad);
        __syncwarp();

        // ---- gamma of every sample pair of the window, staged in shared memory (run-time loops: one copy of the
        // variogram / metric code per model instead of one per matrix element) ----
        double t[NBR][NBC];
        double gmax;
        // geographic metric: staged loop (the great-circle formula is too large to replicate per element);
        // KC = 64: staged as well - the direct form (72 inlined copies) costs more in instruction fetch than it saves
        const bool geo = A.metric == CMX_METRIC_GEOGRAPHIC;
        if (geo || NB >= 8) {
            if (!geo) {
                switch (A.model) {
                    case CMX_VARIOGRAM_LINEAR: gmax = chol_pair_gammas<CMX_VARIOGRAM_LINEAR, CMX_METRIC_EUCLIDEAN>(vc, xy, tri, k, lane); break;
                    case CMX_VARIOGRAM_POWER: gmax = chol_pair_gammas<CMX_VARIOGRAM_POWER, CMX_METRIC_EUCLIDEAN>(vc, xy, tri, k, lane); break;
                    case CMX_VARIOGRAM_GAUSSIAN: gmax = chol_pair_gammas<CMX_VARIOGRAM_GAUSSIAN, CMX_METRIC_EUCLIDEAN>(vc, xy, tri, k, lane); break;
                    case CMX_VARIOGRAM_SPHERICAL: gmax = chol_pair_gammas<CMX_VARIOGRAM_SPHERICAL, CMX_METRIC_EUCLIDEAN>(vc, xy, tri, k, lane); break;
                    default: gmax = chol_pair_gammas<CMX_VARIOGRAM_EXPONENTIAL, CMX_METRIC_EUCLIDEAN>(vc, xy, tri, k, lane); break;
                }
            } else
            switch (A.model) {
                case CMX_VARIOGRAM_LINEAR: gmax = chol_pair_gammas<CMX_VARIOGRAM_LINEAR, CMX_METRIC_GEOGRAPHIC>(vc, xy, tri, k, lane); break;
                case CMX_VARIOGRAM_POWER: gmax = chol_pair_gammas<CMX_VARIOGRAM_POWER, CMX_METRIC_GEOGRAPHIC>(vc, xy, tri, k, lane); break;
                case CMX_VARIOGRAM_GAUSSIAN: gmax = chol_pair_gammas<CMX_VARIOGRAM_GAUSSIAN, CMX_METRIC_GEOGRAPHIC>(vc, xy, tri, k, lane); break;
                case CMX_VARIOGRAM_SPHERICAL: gmax = chol_pair_gammas<CMX_VARIOGRAM_SPHERICAL, CMX_METRIC_GEOGRAPHIC>(vc, xy, tri, k, lane); break;
                default: gmax = chol_pair_gammas<CMX_VARIOGRAM_EXPONENTIAL, CMX_METRIC_GEOGRAPHIC>(vc, xy, tri, k, lane); break;
            }
            __syncwarp();
#pragma unroll
            for (int a = 0; a < NB; ++a) {
                const int i = 8 * a + r;
                const int rowoff = i * (i - 1) / 2;
#pragma unroll
                for (int b = 0; b < NBC; ++b) {
                    if (b <= 2 * a + 1) {
                        const int mm = 4 * b + c;
                        t[a][b] = (mm < i && i < k) ? tri[rowoff + mm] : 0.0;
                    }
                }
            }
        } else if constexpr (NB < 8) {
            switch (A.model) {
                case CMX_VARIOGRAM_LINEAR: gmax = chol_assemble_direct<NB, CMX_VARIOGRAM_LINEAR>(vc, xy, k, r, c, t); break;
                case CMX_VARIOGRAM_POWER: gmax = chol_assemble_direct<NB, CMX_VARIOGRAM_POWER>(vc, xy, k, r, c, t); break;
                case CMX_VARIOGRAM_GAUSSIAN: gmax = chol_assemble_direct<NB, CMX_VARIOGRAM_GAUSSIAN>(vc, xy, k, r, c, t); break;
                case CMX_VARIOGRAM_SPHERICAL: gmax = chol_assemble_direct<NB, CMX_VARIOGRAM_SPHERICAL>(vc, xy, k, r, c, t); break;
                default: gmax = chol_assemble_direct<NB, CMX_VARIOGRAM_EXPONENTIAL>(vc, xy, k, r, c, t); break;
            }
        }
        gmax = warp_max(gmax);
        const double c0 = 2.0 * gmax;
        bool fail = bad || !(c0 > 0.0) || !(c0 < __longlong_as_double(0x7ff0000000000000LL));
        const double thr = CHOL_MIN_PIVOT * c0;
        // t holds my elements of gamma (strict lower triangle, zero elsewhere): C = c0 - gamma off the diagonal, c0 on
        // it; padding rows / columns (>= k) are decoupled
#pragma unroll
        for (int a = 0; a < NB; ++a) {
            const int i = 8 * a + r;
#pragma unroll
            for (int b = 0; b < NBC; ++b) {
                if (b <= 2 * a + 1) {
                    const int mm = 4 * b + c;
                    t[a][b] = (i == mm) ? c0 : ((mm < i && i < k) ? c0 - t[a][b] : 0.0);
                }
            }
        }
        // extra rows KC + 0: b, KC + 1: ones, KC + 2: Z (lanes r = 0, 1, 2), zero elsewhere
#pragma unroll
        for (int b = 0; b < NBC; ++b) {
            const int mm = 4 * b + c;
            double v = 0.0;
            if (mm < k && r < 3) {
                const double2 q = bz[mm];
                v = r == 0 ? q.x : (r == 1 ? 1.0 : q.y);
            }
            t[NB][b] = v;
        }

        double lrow[NBR], lcol[NBC];
        // Scale column J by rsqrt(pivot) and publish it (owners: lanes with c == J % 4).
        auto publish = [&](auto Jc) {
            constexpr int J = decltype(Jc)::value;
            constexpr int ph = J / 4, cj = J % 4, aj = J / 8, rj = J % 8;
            const double d = __shfl_sync(kFull, t[aj][ph], rj * 4 + cj);
            fail |= !(d > thr);
            const double rinv = rsqrt_pos(d);
            // every lane scales its own block column ph (same issue slots as a divergent branch would take, but the
            // step stays one basic block, so the scheduler can fill the latencies of this chain with the trailing
            // update); only the owners of column J keep and store the result
            const bool own = c == cj;
            double v[NBR];
#pragma unroll
            for (int a = aj; a < NBR; ++a) {
                const bool live = (a > aj) || (r > rj);  // row 8a + r > J
                v[a] = live ? t[a][ph] * rinv : 0.0;
            }
            t[NB][ph] = own ? v[NB] : t[NB][ph];  // the extra rows keep their L entries: they are y', u', w'
            if (own) {
                double* rb = rowbuf + ((J & 1) * 8 + r) * ROWP;
#pragma unroll
                for (int a = aj; a < NBR; ++a) rb[a] = v[a];
                double* cb = colbuf + ((J & 1) * 4 + (r & 3)) * COLP + (r >> 2) * NB;
#pragma unroll
                for (int a = aj; a < NB; ++a) cb[a] = v[a];
            }
        };
        // Load the published column J into lrow (rows 8a + r) and lcol (columns 4b + c).
        auto load_col = [&](auto Jc) {
            constexpr int J = decltype(Jc)::value;
            constexpr int aj = J / 8, bmin = (J + 1) / 4;  // columns 4b + c > J need b >= (J + 1) / 4
            const double* rb = rowbuf + ((J & 1) * 8 + r) * ROWP;
#pragma unroll
            for (int a = aj & ~1; a < NBR; a += 2) {
                if (a + 1 < ROWP) {
                    const double2 q = *reinterpret_cast<const double2*>(rb + a);
                    if (a >= aj) lrow[a] = q.x;
                    if (a + 1 < NBR) lrow[a + 1] = q.y;
                } else {
                    lrow[a] = rb[a];
                }
            }
            const double* cb = colbuf + ((J & 1) * 4 + c) * COLP;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
                for (int a = 0; a < NB; a += 2) {
                    // pair (a, a + 1) of half h: block columns 2a + h and 2a + 2 + h
                    if (2 * (a + 1) + h >= bmin) {
                        const double2 q = *reinterpret_cast<const double2*>(cb + h * NB + a);
                        if (2 * a + h >= bmin) lcol[2 * a + h] = q.x;
                        lcol[2 * (a + 1) + h] = q.y;
                    }
                }
            }
        };
        // Trailing update of step J with look-ahead publication of column J + 1.
        auto step = [&](auto Jc) {
            constexpr int J = decltype(Jc)::value;
            constexpr int aj = J / 8, bmin = (J + 1) / 4;
            constexpr int J1 = J + 1;
            if constexpr (J1 < KC) {
#pragma unroll
                for (int a = aj; a < NBR; ++a)
                    if (a == NB || bmin <= 2 * a + 1) t[a][bmin] = fma(-lrow[a], lcol[bmin], t[a][bmin]);
                publish(std::integral_constant<int, J1>{});
            }
#pragma unroll
            for (int a = aj; a < NBR; ++a) {
#pragma unroll
                for (int b = bmin + (J1 < KC ? 1 : 0); b < NBC; ++b)
                    if (a == NB || b <= 2 * a + 1) t[a][b] = fma(-lrow[a], lcol[b], t[a][b]);
            }
            if constexpr (J1 < KC) {
                __syncwarp();
                load_col(std::integral_constant<int, J1>{});
            }
        };

        publish(std::integral_constant<int, 0>{});
        __syncwarp();
        load_col(std::integral_constant<int, 0>{});
        static_for<0, NBC>([&](auto PHc) {
            constexpr int PH = decltype(PHc)::value;
            if (4 * PH < k4) chol_phase_sync();  // k4 is uniform over the CTA; `fail` is not
            if (4 * PH < k4 && !fail) {
                step(std::integral_constant<int, 4 * PH + 0>{});
                step(std::integral_constant<int, 4 * PH + 1>{});
                step(std::integral_constant<int, 4 * PH + 2>{});
                step(std::integral_constant<int, 4 * PH + 3>{});
            }
        });

        // ---- y', u', w' -> five dot products ----
        if (r < 3) {
#pragma unroll
            for (int b = 0; b < NBC; ++b) ext[r * KC + 4 * b + c] = t[NB][b];
        }
        __syncwarp();
        double yy = 0.0, uy = 0.0, uu = 0.0, wy = 0.0, wu = 0.0;
#pragma unroll
        for (int i0 = 0; i0 < KC; i0 += 32) {
            const int i = i0 + lane;
            if (i < KC) {
                const double y = ext[i], u = ext[KC + i], w = ext[2 * KC + i];
                yy = fma(y, y, yy);
                uy = fma(u, y, uy);
                uu = fma(u, u, uu);
                wy = fma(w, y, wy);
                wu = fma(w, u, wu);
            }
        }
        yy = warp_sum(yy);
        uy = warp_sum(uy);
        uu = warp_sum(uu);
        wy = warp_sum(wy);
        wu = warp_sum(wu);
        const double mup = (uy - 1.0) / uu;
        const double zz = wy - mup * wu;
        const double var = -((yy - mup * uy) + (mup + c0));
        if (!isfinite(mup) || !isfinite(zz)) fail = true;
        if (lane == 0 && active) {
            if (fail) {
                fail_list[atomicAdd(fail_count, 1)] = (int)m;
            } else {
                const double res = zz * A.out_scale + A.out_shift;
                if (A.out_is_f32) {
                    static_cast<float*>(A.out)[m] = (float)res;
                    if (A.sigmasq) static_cast<float*>(A.sigmasq)[m] = (float)var;
                } else {
                    static_cast<double*>(A.out)[m] = res;
                    if (A.sigmasq) static_cast<double*>(A.sigmasq)[m] = var;
                }
            }
        }
        __syncwarp();  // shared memory of the warp is reused by its next window
    }
}

template <int NB>
int launch_ok_chol(const OkArgs& a, int* fail_list, int* fail_count, cudaStream_t stream) {
    using Cfg = CholCfg<NB>;
    CMX_CUDA(cudaFuncSetAttribute(ok_chol_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    long long ctas = (a.M + Cfg::WARPS - 1) / Cfg::WARPS;
    const long long cap = (long long)sm_count() * chol_ctas_per_sm(NB) * (CMX_CHOL_SYNC ? 1 : 4);
    if (ctas > cap) ctas = cap;
    ok_chol_kernel<NB><<<(unsigned)ctas, Cfg::WARPS * 32, Cfg::SMEM, stream>>>(a, fail_list, fail_count);
    note_launch();
    return check_launch();
}

}  // namespace
}  // namespace cmx
