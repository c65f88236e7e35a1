// Variogram-model fit: pykrige core._calculate_variogram_model (weight = False), i.e.
//
//     scipy.optimize.least_squares(residuals, x0, bounds=bnds, loss="soft_l1")
//
// reached from the reference call site src/climatrix/reconstruct/kriging.py:215-224 (the pykrige constructor).
// pykrige passes no tolerances, so what the reference obtains is NOT the minimiser but the point where SciPy's
// Trust Region Reflective method stops at ftol = xtol = gtol = 1e-8 - it differs from the true minimum by 1e-6 ...
// 1e-2 in the parameters (measured, tools in DESIGN.md).  Parity therefore needs the same iteration, not just the
// same objective.  This file restates that iteration for the 2-3 parameter, <= 64 residual problems of the path:
//
//   * residuals r_i = gamma_model(p, lag_i) - semivariance_i          (pykrige variogram_models.py)
//   * start point / bounds of pykrige for each model                  (pykrige core.py)
//   * Jacobian by forward differences, h = sqrt(eps) sign(x) max(1, |x|), flipped / shortened at the bounds
//   * robust loss rho(z) = 2 (sqrt(1 + z) - 1) through the usual rescaling of residuals and Jacobian rows
//   * Branch-Coleman-Li trust-region-reflective step: Coleman-Li scaling, exact trust-region sub-problem on the
//     SVD of the augmented Jacobian (secular equation by safeguarded Newton), reflected and gradient candidates
//   * radius update and the ftol / xtol / gtol tests
//
// following the published algorithm (Branch, Coleman, Li 1999) with SciPy's documented constants.  The same source
// compiles for the host (cmx_variogram_fit_host; tests/test_variogram_fit.py compares it with SciPy on the CPU) and
// for the device (variogram_fit_kernel: one thread, a few hundred microseconds, no host round trip).
#pragma once

#include <float.h>
#include <math.h>
#ifdef CMX_VFIT_TRACE
#include <stdio.h>
#endif

#ifdef __CUDACC__
#define CMX_HD __host__ __device__
#else
#define CMX_HD
#endif

namespace cmx {
namespace vfit {

constexpr int MAXM = 64;  // residuals (lags)
constexpr int MAXN = 3;   // parameters
constexpr double EPS = 2.220446049250313e-16;

enum { LINEAR = 0, POWER = 1, GAUSSIAN = 2, SPHERICAL = 3, EXPONENTIAL = 4 };

CMX_HD inline double model_value(int model, const double* p, double d) {
    switch (model) {
        case LINEAR: return p[0] * d + p[1];
        case POWER: return p[0] * pow(d, p[1]) + p[2];
        case GAUSSIAN: {
            const double r = p[1] * 4.0 / 7.0;
            return p[0] * (1.0 - exp(-(d * d) / (r * r))) + p[2];
        }
        case SPHERICAL:
            if (d <= p[1]) return p[0] * ((3.0 * d) / (2.0 * p[1]) - (d * d * d) / (2.0 * (p[1] * p[1] * p[1]))) + p[2];
            return p[0] + p[2];
        default: return p[0] * (1.0 - exp(-d / (p[1] / 3.0))) + p[2];
    }
}

struct Problem {
    int model, m, n;
    const double* lags;
    const double* sv;
    double lb[MAXN], ub[MAXN];
};

CMX_HD inline void residuals(const Problem& P, const double* x, double* f) {
    for (int i = 0; i < P.m; ++i) f[i] = model_value(P.model, x, P.lags[i]) - P.sv[i];
}

CMX_HD inline double norm_n(const double* v, int n) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += v[i] * v[i];
    return sqrt(s);
}

CMX_HD inline bool finite_d(double v) { return v - v == 0.0; }
CMX_HD inline double inf_d() { return HUGE_VAL; }

// x strictly inside the bounds; rstep > 0: pulled in by rstep * max(1, |bound|), rstep == 0: by one ulp
CMX_HD inline void make_strictly_feasible(double* x, const double* lb, const double* ub, int n, double rstep) {
    for (int i = 0; i < n; ++i) {
        int active = 0;
        if (rstep == 0.0) {
            if (x[i] <= lb[i]) active = -1;
            if (x[i] >= ub[i]) active = 1;
        } else {
            const double lower_dist = x[i] - lb[i], upper_dist = ub[i] - x[i];
            const double lt = rstep * fmax(1.0, fabs(lb[i])), ut = rstep * fmax(1.0, fabs(ub[i]));
            if (finite_d(lb[i]) && lower_dist <= fmin(upper_dist, lt)) active = -1;
            if (finite_d(ub[i]) && upper_dist <= fmin(lower_dist, ut)) active = 1;
        }
        if (active == -1) x[i] = rstep == 0.0 ? nextafter(lb[i], ub[i]) : lb[i] + rstep * fmax(1.0, fabs(lb[i]));
        if (active == 1) x[i] = rstep == 0.0 ? nextafter(ub[i], lb[i]) : ub[i] - rstep * fmax(1.0, fabs(ub[i]));
        if (x[i] < lb[i] || x[i] > ub[i]) x[i] = 0.5 * (lb[i] + ub[i]);
    }
}

// forward-difference Jacobian (column-major J[j * MAXM + i] = d f_i / d x_j), steps kept inside the bounds
CMX_HD inline void jacobian(const Problem& P, const double* x, const double* f0, double* J) {
    const double rstep = sqrt(EPS);
    double x1[MAXN], f1[MAXM];
    for (int j = 0; j < P.n; ++j) {
        const double sign = x[j] >= 0.0 ? 1.0 : -1.0;
        double h = rstep * sign * fmax(1.0, fabs(x[j]));
        const double lower_dist = x[j] - P.lb[j], upper_dist = P.ub[j] - x[j];
        if (!(P.lb[j] == -inf_d() && P.ub[j] == inf_d())) {
            const double xh = x[j] + h;
            const bool violated = xh < P.lb[j] || xh > P.ub[j];
            const bool fitting = fabs(h) <= fmax(lower_dist, upper_dist);
            if (violated && fitting) h = -h;
            if (!fitting) h = upper_dist >= lower_dist ? upper_dist : -lower_dist;
        }
        for (int q = 0; q < P.n; ++q) x1[q] = x[q];
        x1[j] = x[j] + h;
        const double dx = x1[j] - x[j];
        residuals(P, x1, f1);
        for (int i = 0; i < P.m; ++i) J[j * MAXM + i] = (f1[i] - f0[i]) / dx;
    }
}

// soft-L1: cost = 0.5 sum rho0; f and the rows of J rescaled so that 0.5 |f|^2 models the robust cost locally
CMX_HD inline double soft_l1_cost(const double* f, int m) {
    double s = 0.0;
    for (int i = 0; i < m; ++i) s += 2.0 * (sqrt(1.0 + f[i] * f[i]) - 1.0);
    return 0.5 * s;
}
CMX_HD inline void scale_for_soft_l1(double* J, double* f, int m, int n) {
    for (int i = 0; i < m; ++i) {
        const double t = 1.0 + f[i] * f[i];
        const double rho1 = pow(t, -0.5), rho2 = -0.5 * pow(t, -1.5);
        double js = rho1 + 2.0 * rho2 * (f[i] * f[i]);
        if (js < EPS) js = EPS;
        js = sqrt(js);
        f[i] *= rho1 / js;
        for (int j = 0; j < n; ++j) J[j * MAXM + i] *= js;
    }
}

CMX_HD inline void cl_scaling(const double* x, const double* g, const double* lb, const double* ub, int n, double* v, double* dv) {
    for (int i = 0; i < n; ++i) {
        v[i] = 1.0;
        dv[i] = 0.0;
        if (g[i] < 0.0 && finite_d(ub[i])) {
            v[i] = ub[i] - x[i];
            dv[i] = -1.0;
        }
        if (g[i] > 0.0 && finite_d(lb[i])) {
            v[i] = x[i] - lb[i];
            dv[i] = 1.0;
        }
    }
}

// thin SVD A = U diag(s) V^T of a (rows x n) column-major matrix (leading dimension ld) by one-sided Jacobi;
// singular values descending.  On return A holds U diag(s) (columns), V is n x n row-major V[i * MAXN + j].
CMX_HD inline void svd_jacobi(double* A, int rows, int n, int ld, double* s, double* V) {
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) V[i * MAXN + j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < n - 1; ++p) {
            for (int q = p + 1; q < n; ++q) {
                double alpha = 0.0, beta = 0.0, gamma = 0.0;
                for (int i = 0; i < rows; ++i) {
                    alpha += A[p * ld + i] * A[p * ld + i];
                    beta += A[q * ld + i] * A[q * ld + i];
                    gamma += A[p * ld + i] * A[q * ld + i];
                }
                if (gamma == 0.0 || fabs(gamma) <= 1e-300) continue;
                const double lim = EPS * sqrt(alpha * beta);
                if (fabs(gamma) <= lim) continue;
                off = fmax(off, fabs(gamma) / sqrt(alpha * beta));
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
                for (int i = 0; i < rows; ++i) {
                    const double ap = A[p * ld + i], aq = A[q * ld + i];
                    A[p * ld + i] = c * ap - sn * aq;
                    A[q * ld + i] = sn * ap + c * aq;
                }
                for (int i = 0; i < n; ++i) {
                    const double vp = V[i * MAXN + p], vq = V[i * MAXN + q];
                    V[i * MAXN + p] = c * vp - sn * vq;
                    V[i * MAXN + q] = sn * vp + c * vq;
                }
            }
        }
        if (off <= 1e-15) break;
    }
    for (int j = 0; j < n; ++j) {
        double a = 0.0;
        for (int i = 0; i < rows; ++i) a += A[j * ld + i] * A[j * ld + i];
        s[j] = sqrt(a);
    }
    for (int a = 0; a < n - 1; ++a) {  // selection sort, descending
        int best = a;
        for (int b = a + 1; b < n; ++b)
            if (s[b] > s[best]) best = b;
        if (best != a) {
            const double ts = s[a];
            s[a] = s[best];
            s[best] = ts;
            for (int i = 0; i < rows; ++i) {
                const double tv = A[a * ld + i];
                A[a * ld + i] = A[best * ld + i];
                A[best * ld + i] = tv;
            }
            for (int i = 0; i < n; ++i) {
                const double tv = V[i * MAXN + a];
                V[i * MAXN + a] = V[i * MAXN + best];
                V[i * MAXN + best] = tv;
            }
        }
    }
}

// min |J p + f|^2 subject to |p| <= Delta from the SVD (uf = U^T f, s, V); returns alpha (Levenberg-Marquardt parameter)
CMX_HD inline double solve_lsq_trust_region(int n, int m, const double* uf, const double* s, const double* V, double Delta,
                                            double initial_alpha, double* p) {
    double suf[MAXN];
    for (int i = 0; i < n; ++i) suf[i] = s[i] * uf[i];
    bool full_rank = false;
    if (m >= n) full_rank = s[n - 1] > EPS * m * s[0];
    if (full_rank) {
        double q[MAXN];
        for (int i = 0; i < n; ++i) q[i] = uf[i] / s[i];
        for (int i = 0; i < n; ++i) {
            double a = 0.0;
            for (int j = 0; j < n; ++j) a += V[i * MAXN + j] * q[j];
            p[i] = -a;
        }
        if (norm_n(p, n) <= Delta) return 0.0;
    }
    double alpha_upper = norm_n(suf, n) / Delta;
    double alpha_lower = 0.0;
    auto phi_and_derivative = [&](double alpha, double& phi, double& phi_prime) {
        double q[MAXN], sum3 = 0.0;
        for (int i = 0; i < n; ++i) {
            const double denom = s[i] * s[i] + alpha;
            q[i] = suf[i] / denom;
            sum3 += suf[i] * suf[i] / (denom * denom * denom);
        }
        const double p_norm = norm_n(q, n);
        phi = p_norm - Delta;
        phi_prime = -sum3 / p_norm;
    };
    if (full_rank) {
        double phi, phi_prime;
        phi_and_derivative(0.0, phi, phi_prime);
        alpha_lower = -phi / phi_prime;
    }
    double alpha;
    if (!full_rank && initial_alpha == 0.0)
        alpha = fmax(0.001 * alpha_upper, sqrt(alpha_lower * alpha_upper));
    else
        alpha = initial_alpha;
    for (int it = 0; it < 10; ++it) {
        if (alpha < alpha_lower || alpha > alpha_upper) alpha = fmax(0.001 * alpha_upper, sqrt(alpha_lower * alpha_upper));
        double phi, phi_prime;
        phi_and_derivative(alpha, phi, phi_prime);
        if (phi < 0.0) alpha_upper = alpha;
        const double ratio = phi / phi_prime;
        alpha_lower = fmax(alpha_lower, alpha - ratio);
        alpha -= (phi + Delta) * ratio / Delta;
        if (fabs(phi) < 0.01 * Delta) break;
    }
    double q[MAXN];
    for (int i = 0; i < n; ++i) q[i] = suf[i] / (s[i] * s[i] + alpha);
    for (int i = 0; i < n; ++i) {
        double a = 0.0;
        for (int j = 0; j < n; ++j) a += V[i * MAXN + j] * q[j];
        p[i] = -a;
    }
    const double scale = Delta / norm_n(p, n);
    for (int i = 0; i < n; ++i) p[i] *= scale;
    return alpha;
}

CMX_HD inline void matvec(const double* J, int m, int n, const double* s, double* out) {  // out = J s
    for (int i = 0; i < m; ++i) {
        double a = 0.0;
        for (int j = 0; j < n; ++j) a += J[j * MAXM + i] * s[j];
        out[i] = a;
    }
}
CMX_HD inline double dot_n(const double* a, const double* b, int n) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

// 0.5 s^T (J^T J + diag) s + g^T s
CMX_HD inline double evaluate_quadratic(const double* J, int m, int n, const double* g, const double* s, const double* diag) {
    double Js[MAXM];
    matvec(J, m, n, s, Js);
    double q = dot_n(Js, Js, m);
    for (int i = 0; i < n; ++i) q += s[i] * diag[i] * s[i];
    return 0.5 * q + dot_n(s, g, n);
}

// coefficients of the quadratic along s (starting at s0 when given): a t^2 + b t + c
CMX_HD inline void build_quadratic_1d(const double* J, int m, int n, const double* g, const double* s, const double* diag,
                                      const double* s0, double& a, double& b, double& c) {
    double v[MAXM];
    matvec(J, m, n, s, v);
    a = dot_n(v, v, m);
    for (int i = 0; i < n; ++i) a += s[i] * diag[i] * s[i];
    a *= 0.5;
    b = dot_n(g, s, n);
    c = 0.0;
    if (s0) {
        double u[MAXM];
        matvec(J, m, n, s0, u);
        b += dot_n(u, v, m);
        c = 0.5 * dot_n(u, u, m) + dot_n(g, s0, n);
        for (int i = 0; i < n; ++i) {
            b += s0[i] * diag[i] * s[i];
            c += 0.5 * s0[i] * diag[i] * s0[i];
        }
    }
}

CMX_HD inline void minimize_quadratic_1d(double a, double b, double lb, double ub, double c, double& t_best, double& y_best) {
    double t[3] = {lb, ub, 0.0};
    int nt = 2;
    if (a != 0.0) {
        const double extremum = -0.5 * b / a;
        if (lb < extremum && extremum < ub) t[nt++] = extremum;
    }
    t_best = t[0];
    y_best = t[0] * (a * t[0] + b) + c;
    for (int i = 1; i < nt; ++i) {
        const double y = t[i] * (a * t[i] + b) + c;
        if (y < y_best) {  // first minimum wins, like argmin
            y_best = y;
            t_best = t[i];
        }
    }
}

CMX_HD inline double step_size_to_bound(const double* x, const double* s, const double* lb, const double* ub, int n, int* hits) {
    double steps[MAXN], min_step = inf_d();
    for (int i = 0; i < n; ++i) {
        steps[i] = inf_d();
        if (s[i] != 0.0) steps[i] = fmax((lb[i] - x[i]) / s[i], (ub[i] - x[i]) / s[i]);
        min_step = fmin(min_step, steps[i]);
    }
    if (hits)
        for (int i = 0; i < n; ++i) hits[i] = steps[i] == min_step ? (s[i] > 0.0 ? 1 : (s[i] < 0.0 ? -1 : 0)) : 0;
    return min_step;
}

// positive root pair of |x + t s| = Delta
CMX_HD inline void intersect_trust_region(const double* x, const double* s, int n, double Delta, double& t_neg, double& t_pos) {
    const double a = dot_n(s, s, n), b = dot_n(x, s, n), c = dot_n(x, x, n) - Delta * Delta;
    const double d = sqrt(b * b - a * c);
    const double q = -(b + copysign(d, b));
    const double t1 = q / a, t2 = c / q;
    if (t1 < t2) {
        t_neg = t1;
        t_pos = t2;
    } else {
        t_neg = t2;
        t_pos = t1;
    }
}

// the trust-region-reflective candidate selection; returns the predicted reduction, step in `step`, hat-space step in `step_h`
CMX_HD inline double select_step(const double* x, const double* J_h, int m, int n, const double* diag_h, const double* g_h,
                                 double* p, double* p_h, const double* d, double Delta, const double* lb, const double* ub,
                                 double theta, double* step, double* step_h) {
    bool inside = true;
    for (int i = 0; i < n; ++i) {
        const double xp = x[i] + p[i];
        if (!(xp >= lb[i] && xp <= ub[i])) inside = false;
    }
    if (inside) {
        const double p_value = evaluate_quadratic(J_h, m, n, g_h, p_h, diag_h);
        for (int i = 0; i < n; ++i) {
            step[i] = p[i];
            step_h[i] = p_h[i];
        }
        return -p_value;
    }
    int hits[MAXN];
    const double p_stride = step_size_to_bound(x, p, lb, ub, n, hits);
    double r_h[MAXN], r[MAXN], x_on_bound[MAXN];
    for (int i = 0; i < n; ++i) {
        r_h[i] = hits[i] != 0 ? -p_h[i] : p_h[i];
        r[i] = d[i] * r_h[i];
    }
    for (int i = 0; i < n; ++i) {
        p[i] *= p_stride;
        p_h[i] *= p_stride;
        x_on_bound[i] = x[i] + p[i];
    }
    double t_neg, to_tr;
    intersect_trust_region(p_h, r_h, n, Delta, t_neg, to_tr);
    const double to_bound = step_size_to_bound(x_on_bound, r, lb, ub, n, nullptr);
    double r_stride = fmin(to_bound, to_tr), r_stride_l, r_stride_u;
    if (r_stride > 0.0) {
        r_stride_l = (1.0 - theta) * p_stride / r_stride;
        r_stride_u = r_stride == to_bound ? theta * to_bound : to_tr;
    } else {
        r_stride_l = 0.0;
        r_stride_u = -1.0;
    }
    double r_value;
    if (r_stride_l <= r_stride_u) {
        double a, b, c;
        build_quadratic_1d(J_h, m, n, g_h, r_h, diag_h, p_h, a, b, c);
        minimize_quadratic_1d(a, b, r_stride_l, r_stride_u, c, r_stride, r_value);
        for (int i = 0; i < n; ++i) {
            r_h[i] = r_h[i] * r_stride + p_h[i];
            r[i] = r_h[i] * d[i];
        }
    } else {
        r_value = inf_d();
    }
    for (int i = 0; i < n; ++i) {
        p[i] *= theta;
        p_h[i] *= theta;
    }
    const double p_value = evaluate_quadratic(J_h, m, n, g_h, p_h, diag_h);
    double ag_h[MAXN], ag[MAXN];
    for (int i = 0; i < n; ++i) {
        ag_h[i] = -g_h[i];
        ag[i] = d[i] * ag_h[i];
    }
    const double to_tr_g = Delta / norm_n(ag_h, n);
    const double to_bound_g = step_size_to_bound(x, ag, lb, ub, n, nullptr);
    double ag_stride = to_bound_g < to_tr_g ? theta * to_bound_g : to_tr_g;
    double a, b, c, ag_value;
    build_quadratic_1d(J_h, m, n, g_h, ag_h, diag_h, nullptr, a, b, c);
    minimize_quadratic_1d(a, b, 0.0, ag_stride, 0.0, ag_stride, ag_value);
    for (int i = 0; i < n; ++i) {
        ag_h[i] *= ag_stride;
        ag[i] *= ag_stride;
    }
    const double* best = ag;
    const double* best_h = ag_h;
    double value = ag_value;
    if (p_value < r_value && p_value < ag_value) {
        best = p;
        best_h = p_h;
        value = p_value;
    } else if (r_value < p_value && r_value < ag_value) {
        best = r;
        best_h = r_h;
        value = r_value;
    }
    for (int i = 0; i < n; ++i) {
        step[i] = best[i];
        step_h[i] = best_h[i];
    }
    return -value;
}

// pykrige's start point and bounds for `model`; returns the number of parameters
CMX_HD inline int setup_problem(Problem& P, int model, int m, const double* lags, const double* sv, double* x0) {
    P.model = model;
    P.m = m;
    P.lags = lags;
    P.sv = sv;
    double sv_max = sv[0], sv_min = sv[0], lag_max = lags[0], lag_min = lags[0];
    for (int i = 1; i < m; ++i) {
        sv_max = fmax(sv_max, sv[i]);
        sv_min = fmin(sv_min, sv[i]);
        lag_max = fmax(lag_max, lags[i]);
        lag_min = fmin(lag_min, lags[i]);
    }
    if (model == LINEAR) {
        P.n = 2;
        x0[0] = (sv_max - sv_min) / (lag_max - lag_min);
        x0[1] = sv_min;
        P.lb[0] = 0.0, P.lb[1] = 0.0;
        P.ub[0] = inf_d(), P.ub[1] = sv_max;
    } else if (model == POWER) {
        P.n = 3;
        x0[0] = (sv_max - sv_min) / (lag_max - lag_min);
        x0[1] = 1.1;
        x0[2] = sv_min;
        P.lb[0] = 0.0, P.lb[1] = 0.001, P.lb[2] = 0.0;
        P.ub[0] = inf_d(), P.ub[1] = 1.999, P.ub[2] = sv_max;
    } else {
        P.n = 3;
        x0[0] = sv_max - sv_min;
        x0[1] = 0.25 * lag_max;
        x0[2] = sv_min;
        P.lb[0] = 0.0, P.lb[1] = 0.0, P.lb[2] = 0.0;
        P.ub[0] = 10.0 * sv_max, P.ub[1] = lag_max, P.ub[2] = sv_max;
    }
    return P.n;
}

// status: 1 gtol, 2 ftol, 3 xtol, 4 both, 0 evaluation budget spent, -1 bad input (lb >= ub, start outside, non-finite residuals)
CMX_HD inline int variogram_fit_trf(int model, int m, const double* lags, const double* sv, double* out, int* nfev_out) {
    const double ftol = 1e-8, xtol = 1e-8, gtol = 1e-8;
    if (m < 1 || m > MAXM) return -1;
    Problem P;
    double x[MAXN];
    const int n = setup_problem(P, model, m, lags, sv, x);
    for (int i = 0; i < n; ++i) {
        if (!(P.lb[i] < P.ub[i])) return -1;
        if (!(x[i] >= P.lb[i] && x[i] <= P.ub[i])) return -1;
    }
    make_strictly_feasible(x, P.lb, P.ub, n, 1e-10);

    double f[MAXM], f_true[MAXM], J[MAXN * MAXM], g[MAXN];
    residuals(P, x, f_true);
    for (int i = 0; i < m; ++i)
        if (!finite_d(f_true[i])) return -1;
    jacobian(P, x, f_true, J);
    int nfev = 1;
    double cost = soft_l1_cost(f_true, m);
    for (int i = 0; i < m; ++i) f[i] = f_true[i];
    scale_for_soft_l1(J, f, m, n);
    for (int j = 0; j < n; ++j) g[j] = dot_n(J + j * MAXM, f, m);

    double v[MAXN], dv[MAXN];
    cl_scaling(x, g, P.lb, P.ub, n, v, dv);
    double tmp[MAXN];
    for (int i = 0; i < n; ++i) tmp[i] = x[i] / sqrt(v[i]);
    double Delta = norm_n(tmp, n);
    if (Delta == 0.0) Delta = 1.0;
    const int max_nfev = n * 100;
    double alpha = 0.0;
    int status = 0;
    bool terminated = false;

    while (true) {
        cl_scaling(x, g, P.lb, P.ub, n, v, dv);
        double g_norm = 0.0;
        for (int i = 0; i < n; ++i) g_norm = fmax(g_norm, fabs(g[i] * v[i]));
        if (g_norm < gtol) {
            status = 1;
            terminated = true;
        }
#if defined(CMX_VFIT_TRACE) && (!defined(__CUDA_ARCH__) || defined(CMX_VFIT_TRACE_DEVICE))
        printf("nfev %3d cost %.10e g_norm %.4e Delta %.6e alpha %.4e x = %.12e %.12e %.12e\n", nfev, cost, g_norm, Delta, alpha, x[0], x[1], n > 2 ? x[2] : 0.0);
#endif
        if (terminated || nfev == max_nfev) break;

        double d[MAXN], diag_h[MAXN], g_h[MAXN];
        for (int i = 0; i < n; ++i) {
            d[i] = sqrt(v[i]);
            diag_h[i] = g[i] * dv[i];
            g_h[i] = d[i] * g[i];
        }
        // augmented system [J diag(d); diag(sqrt(diag_h))], right-hand side [f; 0]
        constexpr int LD = MAXM + MAXN;
        double Ja[MAXN * LD], J_h[MAXN * MAXM];
        for (int j = 0; j < n; ++j) {
            for (int i = 0; i < m; ++i) {
                J_h[j * MAXM + i] = J[j * MAXM + i] * d[j];
                Ja[j * LD + i] = J_h[j * MAXM + i];
            }
            for (int i = 0; i < n; ++i) Ja[j * LD + m + i] = i == j ? sqrt(diag_h[j]) : 0.0;
        }
        double s[MAXN], V[MAXN * MAXN], uf[MAXN];
        svd_jacobi(Ja, m + n, n, LD, s, V);
        for (int j = 0; j < n; ++j) {  // uf = U^T [f; 0], U column j = Ja column j / s_j
            double a = 0.0;
            for (int i = 0; i < m; ++i) a += Ja[j * LD + i] * f[i];
            uf[j] = s[j] > 0.0 ? a / s[j] : 0.0;
        }
        const double theta = fmax(0.995, 1.0 - g_norm);
#if defined(CMX_VFIT_TRACE) && !defined(__CUDA_ARCH__)
        printf("   g = %.6e %.6e %.6e  v = %.4e %.4e %.4e dv = %g %g %g\n   s = %.6e %.6e %.6e uf = %.6e %.6e %.6e\n", g[0], g[1], g[2], v[0], v[1], v[2], dv[0], dv[1], dv[2], s[0], s[1], s[2], uf[0], uf[1], uf[2]);
        for (int jj = 0; jj < n; ++jj) { printf("   J col %d:", jj); for (int i = 0; i < m; ++i) printf(" %.6e", J[jj * MAXM + i]); printf("\n"); }
#endif

        double actual_reduction = -1.0, cost_new = cost, x_new[MAXN], f_new[MAXM];
        while (actual_reduction <= 0.0 && nfev < max_nfev) {
            double p_h[MAXN], p[MAXN], step[MAXN], step_h[MAXN];
            alpha = solve_lsq_trust_region(n, m, uf, s, V, Delta, alpha, p_h);
            for (int i = 0; i < n; ++i) p[i] = d[i] * p_h[i];
            const double predicted_reduction = select_step(x, J_h, m, n, diag_h, g_h, p, p_h, d, Delta, P.lb, P.ub, theta, step, step_h);
            for (int i = 0; i < n; ++i) x_new[i] = x[i] + step[i];
            make_strictly_feasible(x_new, P.lb, P.ub, n, 0.0);
            residuals(P, x_new, f_new);
            ++nfev;
            const double step_h_norm = norm_n(step_h, n);
            bool all_finite = true;
            for (int i = 0; i < m; ++i) all_finite = all_finite && finite_d(f_new[i]);
            if (!all_finite) {
                Delta = 0.25 * step_h_norm;
                continue;
            }
            cost_new = soft_l1_cost(f_new, m);
            actual_reduction = cost - cost_new;
            double ratio;
            if (predicted_reduction > 0.0)
                ratio = actual_reduction / predicted_reduction;
            else if (predicted_reduction == 0.0 && actual_reduction == 0.0)
                ratio = 1.0;
            else
                ratio = 0.0;
            double Delta_new = Delta;
            if (ratio < 0.25)
                Delta_new = 0.25 * step_h_norm;
            else if (ratio > 0.75 && step_h_norm > 0.95 * Delta)
                Delta_new = Delta * 2.0;
            const double step_norm = norm_n(step, n), x_norm = norm_n(x, n);
            const bool ftol_ok = actual_reduction < ftol * cost && ratio > 0.25;
            const bool xtol_ok = step_norm < xtol * (xtol + x_norm);
            if (ftol_ok || xtol_ok) {
                status = ftol_ok && xtol_ok ? 4 : (ftol_ok ? 2 : 3);
                terminated = true;
                break;
            }
            alpha *= Delta / Delta_new;
            Delta = Delta_new;
        }
        if (actual_reduction > 0.0) {
            for (int i = 0; i < n; ++i) x[i] = x_new[i];
            for (int i = 0; i < m; ++i) f_true[i] = f[i] = f_new[i];
            cost = cost_new;
            jacobian(P, x, f_true, J);
            scale_for_soft_l1(J, f, m, n);
            for (int j = 0; j < n; ++j) g[j] = dot_n(J + j * MAXM, f, m);
        }
    }
    for (int i = 0; i < n; ++i) out[i] = x[i];
    for (int i = n; i < MAXN; ++i) out[i] = 0.0;
    if (nfev_out) *nfev_out = nfev;
    return status;
}

}  // namespace vfit
}  // namespace cmx
