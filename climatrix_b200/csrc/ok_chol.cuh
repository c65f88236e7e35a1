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
    static constexpr int WARPS = 4;
    static constexpr size_t SMEM = (size_t)WARPS * WARP_DOUBLES * sizeof(double);
};

// resident CTAs per SM the kernel is compiled for: 255 registers at KC = 48 / 64, 128 at KC = 32, 80 at KC = 16
constexpr int chol_ctas_per_sm(int nb) { return nb >= 6 ? 2 : (nb >= 4 ? 4 : 6); }

constexpr double CHOL_MIN_PIVOT = 1.0e-9;  // relative to c0: below it the window goes to the pivoted kernel

template <int NB>
__global__ void __launch_bounds__(128, chol_ctas_per_sm(NB))
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
    const VarioConst vc = vario_const(A.model, A.p0, A.p1, A.p2);

    const long long nwarps = (long long)gridDim.x * Cfg::WARPS;
    for (long long m = (long long)blockIdx.x * Cfg::WARPS + warp; m < A.M; m += nwarps) {
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

        // ---- gamma of every sample pair of the window (run-time loop: ONE copy of the variogram / metric code,
        // each pair evaluated once, evenly spread over the lanes), staged in shared memory ----
        const int npairs = k * (k - 1) / 2;
        double gmax = 0.0;
        for (int e = lane; e < npairs; e += 32) {
            int i = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)e)) * 0.5f);  // row of pair e, then exact
            while (i * (i - 1) / 2 > e) --i;
            while ((i + 1) * i / 2 <= e) ++i;
            const int mm = e - i * (i - 1) / 2;
            const double2 pi = xy[i], pm = xy[mm];
            const double g = gamma_pre(vc, sample_distance(A.metric, pi.x, pi.y, pm.x, pm.y));
            tri[e] = g;
            gmax = fmax(gmax, g);
        }
        __syncwarp();
        gmax = warp_max(gmax);
        const double c0 = 2.0 * gmax;
        bool fail = bad || !(c0 > 0.0) || !(c0 < __longlong_as_double(0x7ff0000000000000LL));
        const double thr = CHOL_MIN_PIVOT * c0;
        // my elements of C = c0 - gamma off the diagonal, c0 on it; padding rows / columns (>= k) are decoupled
        double t[NBR][NBC];
#pragma unroll
        for (int a = 0; a < NB; ++a) {
            const int i = 8 * a + r;
            const int rowoff = i * (i - 1) / 2;
#pragma unroll
            for (int b = 0; b < NBC; ++b) {
                if (b <= 2 * a + 1) {
                    const int mm = 4 * b + c;
                    double v = 0.0;
                    if (mm < i && i < k) v = c0 - tri[rowoff + mm];
                    t[a][b] = (i == mm) ? c0 : v;
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
            const double rinv = rsqrt(d);
            if (c == cj) {
                double v[NBR];
#pragma unroll
                for (int a = aj; a < NBR; ++a) {
                    const bool live = (a > aj) || (r > rj);  // row 8a + r > J
                    v[a] = live ? t[a][ph] * rinv : 0.0;
                }
                t[NB][ph] = v[NB];  // the extra rows keep their L entries: they are y', u', w'
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
        if (lane == 0) {
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
    const long long cap = (long long)sm_count() * chol_ctas_per_sm(NB) * 4;
    if (ctas > cap) ctas = cap;
    ok_chol_kernel<NB><<<(unsigned)ctas, Cfg::WARPS * 32, Cfg::SMEM, stream>>>(a, fail_list, fail_count);
    note_launch();
    return check_launch();
}

}  // namespace
}  // namespace cmx
