// K4, fast path, second generation: the register Cholesky of ok_chol.cuh with TWO columns per step.
// Included by kriging.cu after ok_chol.cuh (shares gamma_model, chol_pair_gammas, rsqrt_pos, the hand-over protocol).
//
// The one-column kernel spends most of its time in the serial part of a step - publish a column, synchronise, load it
// back, update, pick the next pivot - 64 times per window (ncu: fixed-latency `wait` and scoreboard stalls, FP64 pipe
// 29 %).  Here a step eliminates a PAIR of columns (J0, J1 = J0 + 1) with a rank-2 update, so a window needs half as
// many publish / synchronise / load rounds for the same multiply-adds:
//
//   lane (r, c) of the 8 x 4 lane grid owns rows 8a + r and the column PAIR 8b + 2c + {0, 1}: t[a][b][h], b <= a,
//   plus row block NB = the three extra rows (b, 1, Z) as before.  The owners of pair P are the lanes with
//   c == P % 4; the 2 x 2 pivot block sits on two of them (rows J0 = 8 (P / 4) + 2 (P % 4) and J0 + 1) and is
//   broadcast by three shuffles:
//       rinv0 = rsqrt(x00), l10 = x10 rinv0, rinv1 = rsqrt(x11 - l10^2)
//       v0 = a0 rinv0,  v1 = (a1 - v0 l10) rinv1          for every row below the pair
//   (v0, v1) go to shared memory as ONE 16-byte entry per row, ordered once for the lanes that need them as row
//   factors and once as column factors; the update is t -= v0[row] v0[col] + v1[row] v1[col].
#pragma once

namespace cmx {
namespace {

template <int NB>
struct Chol2Cfg {
    static constexpr int KC = 8 * NB;    // neighbour capacity
    static constexpr int NBR = NB + 1;   // row blocks of 8 (the last one: b, 1, Z rows)
    static constexpr int NP = KC / 2;    // column pairs
    // published pair, row order: [r][a] entries of (v0, v1); NBR is odd, so the 8 row groups of a 16-byte load fall
    // into distinct banks.  Column order: [c][b][h] entries of (v0, v1), padded to an odd number of entries per c.
    static constexpr int ROWS = 2 * NBR;           // doubles per r
    static constexpr int COLS = 4 * NB + 2;        // doubles per c
    static constexpr int TRI = KC * (KC - 1) / 2;
    static constexpr int WARP_DOUBLES = 2 * KC + 2 * KC + 2 * 8 * ROWS + 2 * 4 * COLS + 3 * KC + TRI + (TRI & 1);
    static constexpr int WARPS = 4;
    static constexpr size_t SMEM = (size_t)WARPS * WARP_DOUBLES * sizeof(double);
};

constexpr int chol2_ctas_per_sm(int nb) { return nb >= 6 ? 2 : (nb >= 4 ? 4 : 6); }

// gamma is staged through shared memory by the folded pair sweep (chol_pair_gammas): evaluating it straight into the
// owners' registers, as the one-column kernel does below k = 64, measured slower here (29.5 vs 25.0 ms per 1.04 M
// windows at k = 64: with 144 matrix registers live there is no room left to overlap the chains).
template <int NB>
__global__ void __launch_bounds__(128, chol2_ctas_per_sm(NB))
ok_chol2_kernel(const OkArgs A, int* __restrict__ fail_list, int* __restrict__ fail_count) {
    using Cfg = Chol2Cfg<NB>;
    constexpr int KC = Cfg::KC, NBR = Cfg::NBR, NP = Cfg::NP, ROWS = Cfg::ROWS, COLS = Cfg::COLS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = lane >> 2, c = lane & 3;
    double* base = reinterpret_cast<double*>(smem_raw) + (size_t)warp * Cfg::WARP_DOUBLES;
    double2* xy = reinterpret_cast<double2*>(base);        // [KC] sample (x, y)
    double2* bz = xy + KC;                                  // [KC] (b, z)
    double* rowbuf = reinterpret_cast<double*>(bz + KC);    // [2][8][ROWS]
    double* colbuf = rowbuf + 2 * 8 * ROWS;                 // [2][4][COLS]
    double* ext = colbuf + 2 * 4 * COLS;                    // [3][KC]  y', u', w'
    double* tri = ext + 3 * KC;                             // [TRI]    gamma(i, m), m < i
    const int k = A.k;
    const int k8 = (k + 7) & ~7;
    double P0, P1, P2;
    load_params(A, P0, P1, P2);
    const VarioConst vc = vario_const(A.model, P0, P1, P2);

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

        // ---- gamma of my elements ----
        double t[NBR][NB][2];
        double gmax;
        {
            const bool geo = A.metric == CMX_METRIC_GEOGRAPHIC;
            if (!geo) {
                switch (A.model) {
                    case CMX_VARIOGRAM_LINEAR: gmax = chol_pair_gammas<CMX_VARIOGRAM_LINEAR, CMX_METRIC_EUCLIDEAN, (KC + 31) / 32>(vc, xy, tri, k, lane); break;
                    case CMX_VARIOGRAM_POWER: gmax = chol_pair_gammas<CMX_VARIOGRAM_POWER, CMX_METRIC_EUCLIDEAN, (KC + 31) / 32>(vc, xy, tri, k, lane); break;
                    case CMX_VARIOGRAM_GAUSSIAN: gmax = chol_pair_gammas<CMX_VARIOGRAM_GAUSSIAN, CMX_METRIC_EUCLIDEAN, (KC + 31) / 32>(vc, xy, tri, k, lane); break;
                    case CMX_VARIOGRAM_SPHERICAL: gmax = chol_pair_gammas<CMX_VARIOGRAM_SPHERICAL, CMX_METRIC_EUCLIDEAN, (KC + 31) / 32>(vc, xy, tri, k, lane); break;
                    default: gmax = chol_pair_gammas<CMX_VARIOGRAM_EXPONENTIAL, CMX_METRIC_EUCLIDEAN, (KC + 31) / 32>(vc, xy, tri, k, lane); break;
                }
            } else {
                switch (A.model) {
                    case CMX_VARIOGRAM_LINEAR: gmax = chol_pair_gammas<CMX_VARIOGRAM_LINEAR, CMX_METRIC_GEOGRAPHIC, (KC + 31) / 32>(vc, xy, tri, k, lane); break;
                    case CMX_VARIOGRAM_POWER: gmax = chol_pair_gammas<CMX_VARIOGRAM_POWER, CMX_METRIC_GEOGRAPHIC, (KC + 31) / 32>(vc, xy, tri, k, lane); break;
                    case CMX_VARIOGRAM_GAUSSIAN: gmax = chol_pair_gammas<CMX_VARIOGRAM_GAUSSIAN, CMX_METRIC_GEOGRAPHIC, (KC + 31) / 32>(vc, xy, tri, k, lane); break;
                    case CMX_VARIOGRAM_SPHERICAL: gmax = chol_pair_gammas<CMX_VARIOGRAM_SPHERICAL, CMX_METRIC_GEOGRAPHIC, (KC + 31) / 32>(vc, xy, tri, k, lane); break;
                    default: gmax = chol_pair_gammas<CMX_VARIOGRAM_EXPONENTIAL, CMX_METRIC_GEOGRAPHIC, (KC + 31) / 32>(vc, xy, tri, k, lane); break;
                }
            }
            __syncwarp();
#pragma unroll
            for (int a = 0; a < NB; ++a) {
                const int i = 8 * a + r;
                const int rowoff = i * (i - 1) / 2;
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    if (b <= a) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int mm = 8 * b + 2 * c + h;
                            t[a][b][h] = (mm < i && i < k) ? tri[rowoff + mm] : 0.0;
                        }
                    }
                }
            }
        }
        gmax = warp_max(gmax);
        const double c0 = 2.0 * gmax;
        bool fail = bad || !(c0 > 0.0) || !(c0 < __longlong_as_double(0x7ff0000000000000LL));
        const double thr = CHOL_MIN_PIVOT * c0;
        // C = c0 - gamma off the diagonal, c0 on it; padding rows / columns (>= k) are decoupled
#pragma unroll
        for (int a = 0; a < NB; ++a) {
            const int i = 8 * a + r;
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                if (b <= a) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int mm = 8 * b + 2 * c + h;
                        t[a][b][h] = (i == mm) ? c0 : ((mm < i && i < k) ? c0 - t[a][b][h] : 0.0);
                    }
                }
            }
        }
        // extra rows KC + 0: b, KC + 1: ones, KC + 2: Z (lanes r = 0, 1, 2), zero elsewhere
#pragma unroll
        for (int b = 0; b < NB; ++b) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int mm = 8 * b + 2 * c + h;
                double v = 0.0;
                if (mm < k && r < 3) {
                    const double2 q = bz[mm];
                    v = r == 0 ? q.x : (r == 1 ? 1.0 : q.y);
                }
                t[NB][b][h] = v;
            }
        }

        double2 lr[NBR];  // row factors (v0, v1) of rows 8a + r; the column factors are read per block column
        // Eliminate pair P inside its 2 x 2 pivot block, scale the two columns and publish them.
        auto publish = [&](auto Pc) {
            constexpr int P = decltype(Pc)::value;
            constexpr int bp = P / 4, cp = P % 4, r0 = 2 * cp, r1 = r0 + 1;  // rows J0 = 8 bp + r0, J1 = J0 + 1
            const double x00 = __shfl_sync(kFull, t[bp][bp][0], r0 * 4 + cp);
            const double x10 = __shfl_sync(kFull, t[bp][bp][0], r1 * 4 + cp);
            const double x11 = __shfl_sync(kFull, t[bp][bp][1], r1 * 4 + cp);
            const double rinv0 = rsqrt_pos(x00);
            const double l10 = x10 * rinv0;
            const double d1 = fma(-l10, l10, x11);
            fail |= !(x00 > thr) || !(d1 > thr);
            const double rinv1 = rsqrt_pos(d1);
            const bool own = c == cp;
            double2* rb = reinterpret_cast<double2*>(rowbuf + ((P & 1) * 8 + r) * ROWS);
            // as column factors: row 8a + r is column 8a + 2 (r >> 1) + (r & 1)
            double2* cb = reinterpret_cast<double2*>(colbuf + ((P & 1) * 4 + (r >> 1)) * COLS) + (r & 1);
#pragma unroll
            for (int a = bp; a < NBR; ++a) {
                const bool live = (a > bp) || (r > r1);  // rows below the pair; finished rows publish zeros
                double v0 = t[a][bp][0] * rinv0;
                double v1 = fma(-v0, l10, t[a][bp][1]) * rinv1;
                if (a == NB) {  // the extra rows keep their L entries: they are y', u', w'
                    t[NB][bp][0] = own ? v0 : t[NB][bp][0];
                    t[NB][bp][1] = own ? v1 : t[NB][bp][1];
                }
                v0 = live ? v0 : 0.0;
                v1 = live ? v1 : 0.0;
                if (own) {
                    rb[a] = make_double2(v0, v1);
                    if (a < NB) cb[2 * a] = make_double2(v0, v1);
                }
            }
        };
        auto load_rows = [&](auto Pc) {
            constexpr int P = decltype(Pc)::value;
            constexpr int bp = P / 4;
            const double2* rb = reinterpret_cast<const double2*>(rowbuf + ((P & 1) * 8 + r) * ROWS);
#pragma unroll
            for (int a = bp; a < NBR; ++a) lr[a] = rb[a];
        };
        // rank-2 update of block column b with pair P
        auto update_column = [&](auto Pc, auto Bc) {
            constexpr int P = decltype(Pc)::value, b = decltype(Bc)::value;
            const double2* cb = reinterpret_cast<const double2*>(colbuf + ((P & 1) * 4 + c) * COLS);
            const double2 c0v = cb[2 * b], c1v = cb[2 * b + 1];
#pragma unroll
            for (int a = b; a < NBR; ++a) {
                t[a][b][0] = fma(-lr[a].y, c0v.y, fma(-lr[a].x, c0v.x, t[a][b][0]));
                t[a][b][1] = fma(-lr[a].y, c1v.y, fma(-lr[a].x, c1v.x, t[a][b][1]));
            }
        };
        auto step = [&](auto Pc) {
            constexpr int P = decltype(Pc)::value;
            constexpr int bp = P / 4;
            constexpr int Pn = P + 1;
            constexpr int bn = Pn / 4;  // block column of the next pair (bp or bp + 1)
            if constexpr (Pn < NP) {
                // the next pair's block column first, so that its publish overlaps the rest of the update
                update_column(Pc, std::integral_constant<int, bn>{});
                publish(std::integral_constant<int, Pn>{});
            }
            static_for<bp, NB>([&](auto Bc) {
                constexpr int b = decltype(Bc)::value;
                if constexpr (!(Pn < NP && b == bn)) update_column(Pc, Bc);
            });
            if constexpr (Pn < NP) {
                __syncwarp();
                load_rows(std::integral_constant<int, Pn>{});
            }
        };

        publish(std::integral_constant<int, 0>{});
        __syncwarp();
        load_rows(std::integral_constant<int, 0>{});
        static_for<0, NB>([&](auto PHc) {
            constexpr int PH = decltype(PHc)::value;
            if (8 * PH < k8 && !fail) {
                step(std::integral_constant<int, 4 * PH + 0>{});
                step(std::integral_constant<int, 4 * PH + 1>{});
                step(std::integral_constant<int, 4 * PH + 2>{});
                step(std::integral_constant<int, 4 * PH + 3>{});
            }
        });

        // ---- y', u', w' -> five dot products ----
        if (r < 3) {
#pragma unroll
            for (int b = 0; b < NB; ++b)
                *reinterpret_cast<double2*>(ext + r * KC + 8 * b + 2 * c) = make_double2(t[NB][b][0], t[NB][b][1]);
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
                store_window(A, m, zz * A.out_scale + A.out_shift, var);
            }
        }
        __syncwarp();  // shared memory of the warp is reused by its next window
    }
}

template <int NB>
int launch_ok_chol2(const OkArgs& a, int* fail_list, int* fail_count, cudaStream_t stream) {
    using Cfg = Chol2Cfg<NB>;
    CMX_CUDA(cudaFuncSetAttribute(ok_chol2_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    long long ctas = (a.M + Cfg::WARPS - 1) / Cfg::WARPS;
    const long long cap = (long long)sm_count() * chol2_ctas_per_sm(NB) * 4;
    if (ctas > cap) ctas = cap;
    ok_chol2_kernel<NB><<<(unsigned)ctas, Cfg::WARPS * 32, Cfg::SMEM, stream>>>(a, fail_list, fail_count);
    note_launch();
    return check_launch();
}

}  // namespace
}  // namespace cmx
