// Global ordinary kriging - the mode climatrix ships (method="ok" without a neighbour count).
//
// Reference: src/climatrix/reconstruct/kriging.py:193-204, 236-241 -> pykrige OrdinaryKriging.execute(backend =
// "vectorized" | "loop") (_exec_vector / _exec_loop): build the (N+1) x (N+1) matrix
// A = [[-gamma(D), 1], [1', 0]] (zero diagonal), invert it (scipy.linalg.inv, or pinv with pseudo_inv=True) and, per
// target t, x_t = A^-1 b_t, z_t = sum_i x_ti Z_i with b_t = [-gamma(d_t); 1] (entries forced to 0 where d <= 1e-10).
//
// climatrix keeps only z (kriging.py:236 drops the variance), and z_t = b_t' A^-1 [Z; 0] by the symmetry of A.  So the
// whole field needs ONE solve, A [w; lambda] = [Z; 0], followed by z_t = sum_i w_i b_ti + lambda: O(N^3 / 3 + M N) instead
// of O(N^3 + M N^2), and no (N+1) x M right-hand side is ever formed.  The solve:
//   * 1'w = 0, so A may be replaced by C = A + c0 11' on the N x N block (C w + lambda 1 = Z, same lambda); with
//     c0 = 2 max gamma, C is symmetric positive definite (-gamma is conditionally positive definite), hence a blocked
//     right-looking CHOLESKY without pivoting: 64 x 64 diagonal blocks factored in shared memory, panel solves, and a
//     register-tiled FP64 rank-64 update of the trailing lower triangle (the only O(N^3) kernel);
//     y = C^-1 Z, u = C^-1 1 by block forward / back substitution; lambda = 1'y / 1'u; w = y - lambda u.
//   * pseudo_inv=True, or a pivot that does not stay positive (numerically singular A): a one-sided JACOBI SVD of the
//     bordered matrix (round-robin pair ordering, one CTA per column pair), w = V S^+ U' [Z; 0] with scipy.linalg.pinv's
//     cut-off (singular values <= max(M, N) eps s_max are dropped).  N <= GK_JACOBI_MAX_N.
// Everything is float64.  Tensor cores are not used: FP64 on the CUDA cores is the B200's double-precision path.
#include "common.cuh"

namespace cmx {
namespace {

constexpr int GB = 64;                 // block size of the factorisation
constexpr int GK_JACOBI_MAX_N = 4096;  // largest bordered system the Jacobi path accepts
constexpr double GK_EPS = 1.0e-10;     // pykrige ok.py: eps (exact_values rule)

__device__ __forceinline__ double gk_gamma(int model, double p0, double p1, double p2, double d) {
    switch (model) {
        case CMX_VARIOGRAM_LINEAR: return p0 * d + p1;
        case CMX_VARIOGRAM_POWER: return p0 * pow(d, p1) + p2;
        case CMX_VARIOGRAM_GAUSSIAN: {
            const double r = p1 * 4.0 / 7.0;
            return p0 * (1.0 - exp(-(d * d) / (r * r))) + p2;
        }
        case CMX_VARIOGRAM_SPHERICAL:
            if (d <= p1) return p0 * ((3.0 * d) / (2.0 * p1) - (d * d * d) / (2.0 * p1 * p1 * p1)) + p2;
            return p0 + p2;
        default: return p0 * (1.0 - exp(-d / (p1 / 3.0))) + p2;
    }
}
__device__ __forceinline__ double gk_distance(int metric, double xi, double yi, double xj, double yj) {
    if (metric == CMX_METRIC_GEOGRAPHIC) return great_circle_deg(xi, yi, xj, yj);
    const double dx = __dsub_rn(xi, xj), dy = __dsub_rn(yi, yj);
    return sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}

// 1 / sqrt(d) for a positive normal d: hardware seed + two Newton steps (1-2 ulp), no slow-path branch
__device__ __forceinline__ double gk_rsqrt(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double h = 0.5 * d;
    y = y * fma(-h * y, y, 1.5);
    y = y * fma(-h * y, y, 1.5);
    return y;
}

struct GkParams {
    int model, metric;
    double p0, p1, p2;
};

// ---- assembly ------------------------------------------------------------------------------------------------------
// gamma(d_ij) into the lower triangle (row-major, leading dimension ld) and the largest value into *gmax
__global__ void gk_gamma_kernel(const double* __restrict__ x, const double* __restrict__ y, long long n, GkParams P,
                                double* __restrict__ C, long long ld, double* gmax) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long i = blockIdx.y;
    double g = 0.0;
    if (j < i && i < n) {
        g = gk_gamma(P.model, P.p0, P.p1, P.p2, gk_distance(P.metric, x[i], y[i], x[j], y[j]));
        C[i * ld + j] = g;
    }
    g = warp_max(g == g ? g : 0.0);
    if ((threadIdx.x & 31) == 0 && g > 0.0) atomic_max_double(gmax, g);
}
// Cholesky path: C = c0 - gamma off the diagonal, c0 on it (c0 = 2 gmax)
__global__ void gk_shift_kernel(long long n, double* __restrict__ C, long long ld, const double* gmax) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long i = blockIdx.y;
    const double c0 = 2.0 * *gmax;
    if (i < n && j <= i) C[i * ld + j] = i == j ? c0 : c0 - C[i * ld + j];
}
// Jacobi path: the full bordered matrix A (np x np, np = n + 1 padded to even; symmetric, so row- = column-major) and V = I
__global__ void gk_bordered_kernel(long long n, long long np, double* __restrict__ G, double* __restrict__ V) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long i = blockIdx.y;
    if (i >= np || j >= np) return;
    double a;
    if (i < n && j < n)
        a = i == j ? 0.0 : -(j < i ? G[i * np + j] : G[j * np + i]);  // (the lower triangle holds gamma)
    else if (i == n && j == n)
        a = 0.0;
    else if ((i == n && j < n) || (j == n && i < n))
        a = 1.0;
    else
        a = 0.0;  // padding row / column
    V[i * np + j] = i == j ? 1.0 : 0.0;
    // into the staging area behind G (gk_copy_kernel moves it): the lower triangle is still being read by other threads
    G[np * np + i * np + j] = a;
}
__global__ void gk_fill_kernel(double* __restrict__ dst, double value, long long count) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
        dst[i] = value;
}
__global__ void gk_copy_kernel(double* __restrict__ dst, const double* __restrict__ src, long long count) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

// ---- blocked Cholesky ----------------------------------------------------------------------------------------------
// diagonal block kb: unblocked Cholesky of a (up to) 64 x 64 block.  *fail |= 1 on a pivot <= thr.
// 16 x 16 threads, each keeps a 4 x 4 sub-tile (rows ty + 16 a, columns tx + 16 b) in registers; per column the owners
// publish the raw column to a double-buffered shared vector, ONE barrier, everybody scales it by the reciprocal root of
// its head and updates its own registers.
__global__ void __launch_bounds__(256) gk_potrf_kernel(double* __restrict__ C, long long ld, long long n, int kb,
                                                       const double* gmax, int* fail) {
    __shared__ double col[2][GB];
    const long long r0 = (long long)kb * GB;
    const int nb = (int)(n - r0 < GB ? n - r0 : GB);
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    double t[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int i = ty + 16 * a, m = tx + 16 * b;
            t[a][b] = (i < nb && m <= i) ? C[(r0 + i) * ld + r0 + m] : (i == m ? 1.0 : 0.0);
        }
    const double thr = 1.0e-13 * 2.0 * *gmax;
    bool bad = false;
#pragma unroll
    for (int j = 0; j < GB; ++j) {
        const int bj = j >> 4, xj = j & 15;  // column j lives in sub-tile column bj of the threads with tx == xj
        if (tx == xj) {
#pragma unroll
            for (int a = 0; a < 4; ++a) col[j & 1][ty + 16 * a] = t[a][bj];
        }
        __syncthreads();
        double d = col[j & 1][j];
        if (!(d > thr)) {
            bad = true;
            d = thr > 0.0 ? thr : 1.0;
        }
        const double rinv = gk_rsqrt(d);
        double ci[4], cm[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int i = ty + 16 * a;
            ci[a] = i > j ? col[j & 1][i] * rinv : 0.0;
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int m = tx + 16 * b;
            cm[b] = m > j ? col[j & 1][m] * rinv : 0.0;
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) t[a][b] = fma(-ci[a], cm[b], t[a][b]);  // (entries above the diagonal are never stored)
        if (tx == xj) {  // the finished column: L[j][j] = sqrt(d), L[i][j] = c_i / sqrt(d)
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int i = ty + 16 * a;
                if (i == j) t[a][bj] = d * rinv;
                else if (i > j) t[a][bj] = ci[a];
            }
        }
    }
    if (bad && tid == 0) atomicOr(fail, 1);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int i = ty + 16 * a, m = tx + 16 * b;
            if (i < nb && m <= i) C[(r0 + i) * ld + r0 + m] = t[a][b];
        }
}
// panel below the diagonal block: X L_kk' = C_ik, one thread per row
__global__ void __launch_bounds__(GB) gk_trsm_kernel(double* __restrict__ C, long long ld, long long n, int kb) {
    __shared__ double L[GB][GB + 1];
    __shared__ double dinv[GB];
    const long long r0 = (long long)kb * GB;
    for (int e = threadIdx.x; e < GB * GB; e += GB) {
        const int i = e / GB, j = e % GB;
        L[i][j] = j <= i ? C[(r0 + i) * ld + r0 + j] : 0.0;
    }
    __syncthreads();
    dinv[threadIdx.x] = 1.0 / L[threadIdx.x][threadIdx.x];
    __syncthreads();
    const long long row = r0 + GB + (long long)blockIdx.x * GB + threadIdx.x;
    if (row >= n) return;
    double x[GB];
    double* c = C + row * ld + r0;
#pragma unroll
    for (int j = 0; j < GB; ++j) x[j] = c[j];
#pragma unroll
    for (int j = 0; j < GB; ++j) {  // column-oriented: the updates of one step are independent of each other
        x[j] = x[j] * dinv[j];
#pragma unroll
        for (int m = 0; m < GB; ++m)
            if (m > j) x[m] = fma(-x[j], L[m][j], x[m]);
    }
#pragma unroll
    for (int j = 0; j < GB; ++j) c[j] = x[j];
}
// trailing update C_ij -= L_ik L_jk' over the lower-triangle tiles (bi >= bj > kb): 64 x 64 tile per CTA, 4 x 4 per thread
__global__ void __launch_bounds__(256) gk_syrk_kernel(double* __restrict__ C, long long ld, long long n, int kb) {
    extern __shared__ __align__(16) double gk_smem[];
    constexpr int LDT = GB + 2;          // padded leading dimension: 16-byte aligned rows, few bank conflicts on the fill
    double* Lt_i = gk_smem;              // [k][row] of the row tile
    double* Lt_j = gk_smem + GB * LDT;   // [k][row] of the column tile
    // linear tile index -> (bi, bj), bj <= bi, both relative to kb + 1
    const long long t = blockIdx.x;
    long long bi = (long long)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while (bi * (bi + 1) / 2 > t) --bi;
    while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
    const long long bj = t - bi * (bi + 1) / 2;
    const long long ri = ((long long)kb + 1 + bi) * GB, rj = ((long long)kb + 1 + bj) * GB, k0 = (long long)kb * GB;
    const int tid = threadIdx.x;
    for (int e = tid; e < GB * GB; e += 256) {
        const int r = e / GB, k = e % GB;
        Lt_i[k * LDT + r] = ri + r < n ? C[(ri + r) * ld + k0 + k] : 0.0;
        Lt_j[k * LDT + r] = rj + r < n ? C[(rj + r) * ld + k0 + k] : 0.0;
    }
    __syncthreads();
    const int ty = tid >> 4, tx = tid & 15;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
#pragma unroll 8
    for (int k = 0; k < GB; ++k) {
        const double2 a01 = *reinterpret_cast<const double2*>(Lt_i + k * LDT + ty * 4);
        const double2 a23 = *reinterpret_cast<const double2*>(Lt_i + k * LDT + ty * 4 + 2);
        const double2 b01 = *reinterpret_cast<const double2*>(Lt_j + k * LDT + tx * 4);
        const double2 b23 = *reinterpret_cast<const double2*>(Lt_j + k * LDT + tx * 4 + 2);
        const double av[4] = {a01.x, a01.y, a23.x, a23.y}, bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const long long row = ri + ty * 4 + a;
        if (row >= n) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const long long col = rj + tx * 4 + b;
            if (col <= row) C[row * ld + col] -= acc[a][b];
        }
    }
}

// ---- triangular solves with two right-hand sides, one launch per block step ------------------------------------------
// Every CTA of a step solves the 64 x 64 diagonal block ITSELF (one warp per right-hand side: 64 serial steps over warp
// shuffles, ~2 us - cheaper than a round trip through global memory and a grid-wide wait), CTA 0 stores the block of the
// solution, every other CTA subtracts its 64 x 64 tile times that block from its part of the working vector.
__device__ __forceinline__ void gk_load_diag(const double* __restrict__ C, long long ld, long long r0, int nb,
                                             double (*L)[GB + 1]) {
    for (int e = threadIdx.x; e < GB * GB; e += blockDim.x) {
        const int i = e / GB, j = e % GB;
        L[i][j] = (i < nb && j <= i) ? C[(r0 + i) * ld + r0 + j] : (i == j ? 1.0 : 0.0);
    }
}
// L y = b.  work [2][n]: b on entry, rows below block kb updated; y [2][n]: block kb of the solution.  Grid: 1 + blocks below.
__global__ void __launch_bounds__(128) gk_forward_step_kernel(const double* __restrict__ C, long long ld, long long n, int kb,
                                                              double* __restrict__ work, double* __restrict__ y) {
    __shared__ double L[GB][GB + 1];
    __shared__ double xs[2][GB];
    __shared__ double dinv[GB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long r0 = (long long)kb * GB;
    const int nb = (int)(n - r0 < GB ? n - r0 : GB);
    gk_load_diag(C, ld, r0, nb, L);
    __syncthreads();
    if (tid < GB) dinv[tid] = 1.0 / L[tid][tid];
    __syncthreads();
    if (warp < 2) {
        double x0 = lane < nb ? work[warp * n + r0 + lane] : 0.0;
        double x1 = lane + 32 < nb ? work[warp * n + r0 + lane + 32] : 0.0;
#pragma unroll 8
        for (int c = 0; c < GB; ++c) {
            const double xc = __shfl_sync(kFull, c < 32 ? x0 : x1, c & 31) * dinv[c];
            if (lane == (c & 31)) {
                if (c < 32) x0 = xc;
                else x1 = xc;
            }
            if (lane > c) x0 = fma(-L[lane][c], xc, x0);
            if (lane + 32 > c) x1 = fma(-L[lane + 32][c], xc, x1);
        }
        xs[warp][lane] = x0;
        xs[warp][lane + 32] = x1;
    }
    __syncthreads();
    const int q = tid >> 6, i = tid & 63;
    if (blockIdx.x == 0) {
        if (i < nb) y[q * n + r0 + i] = xs[q][i];
        return;
    }
    const long long rb = r0 + (long long)blockIdx.x * GB;  // my rows
    for (int e = tid; e < GB * GB; e += 128) {
        const int r = e / GB, c = e % GB;
        L[r][c] = rb + r < n ? C[(rb + r) * ld + r0 + c] : 0.0;
    }
    __syncthreads();
    if (rb + i < n) {
        double a = 0.0;
#pragma unroll 8
        for (int c = 0; c < GB; ++c) a = fma(L[i][c], xs[q][c], a);
        work[q * n + rb + i] -= a;
    }
}
// L' x = y.  work [2][n]: y on entry, blocks left of kb updated; x [2][n]: block kb of the solution.  Grid: 1 + kb.
__global__ void __launch_bounds__(128) gk_backward_step_kernel(const double* __restrict__ C, long long ld, long long n, int kb,
                                                               double* __restrict__ work, double* __restrict__ x) {
    __shared__ double L[GB][GB + 1];
    __shared__ double xs[2][GB];
    __shared__ double dinv[GB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long r0 = (long long)kb * GB;
    const int nb = (int)(n - r0 < GB ? n - r0 : GB);
    gk_load_diag(C, ld, r0, nb, L);
    __syncthreads();
    if (tid < GB) dinv[tid] = 1.0 / L[tid][tid];
    __syncthreads();
    if (warp < 2) {
        double x0 = lane < nb ? work[warp * n + r0 + lane] : 0.0;
        double x1 = lane + 32 < nb ? work[warp * n + r0 + lane + 32] : 0.0;
#pragma unroll 8
        for (int c = GB - 1; c >= 0; --c) {  // row c of L' is column c of L
            const double xc = __shfl_sync(kFull, c < 32 ? x0 : x1, c & 31) * dinv[c];
            if (lane == (c & 31)) {
                if (c < 32) x0 = xc;
                else x1 = xc;
            }
            if (lane < c) x0 = fma(-L[c][lane], xc, x0);
            if (lane + 32 < c) x1 = fma(-L[c][lane + 32], xc, x1);
        }
        xs[warp][lane] = x0;
        xs[warp][lane + 32] = x1;
    }
    __syncthreads();
    const int q = tid >> 6, i = tid & 63;
    if (blockIdx.x == 0) {
        if (i < nb) x[q * n + r0 + i] = xs[q][i];
        return;
    }
    const long long cb = (long long)(blockIdx.x - 1) * GB;  // my columns: a block left of the diagonal
    for (int e = tid; e < GB * GB; e += 128) {
        const int r = e / GB, c = e % GB;
        L[r][c] = r < nb ? C[(r0 + r) * ld + cb + c] : 0.0;
    }
    __syncthreads();
    double a = 0.0;
#pragma unroll 8
    for (int r = 0; r < GB; ++r) a = fma(L[r][i], xs[q][r], a);
    work[q * n + cb + i] -= a;
}
// lambda = 1'y / 1'u, w = y - lambda u  (single CTA; w_ext[n] = lambda)
__global__ void __launch_bounds__(256) gk_combine_kernel(long long n, const double* __restrict__ yu, double* __restrict__ w_ext) {
    __shared__ double red[2][8];
    const int tid = threadIdx.x;
    double sy = 0.0, su = 0.0;
    for (long long i = tid; i < n; i += 256) {
        sy += yu[i];
        su += yu[n + i];
    }
    sy = warp_sum(sy);
    su = warp_sum(su);
    if ((tid & 31) == 0) {
        red[0][tid >> 5] = sy;
        red[1][tid >> 5] = su;
    }
    __syncthreads();
    sy = su = 0.0;
    for (int w = 0; w < 8; ++w) {
        sy += red[0][w];
        su += red[1][w];
    }
    const double lambda = sy / su;
    for (long long i = tid; i < n; i += 256) w_ext[i] = yu[i] - lambda * yu[n + i];
    if (tid == 0) w_ext[n] = lambda;
}

// ---- one-sided Jacobi SVD of the bordered matrix (pseudo_inv / numerically singular systems) -------------------
// step `s` of a round-robin sweep: CTA i orthogonalises its column pair of G (np x np, column j at G + j * np) and
// applies the same rotation to V.  *moved |= 1 when some pair was still measurably non-orthogonal.
__global__ void __launch_bounds__(256) gk_jacobi_kernel(double* __restrict__ G, double* __restrict__ V, long long np, int step,
                                                        int* moved) {
    __shared__ double red[3][8];
    const int i = blockIdx.x, tid = threadIdx.x;
    const int m = (int)np - 1;  // players 0 .. m - 1 rotate, player m stays
    int p, q;
    if (i == 0) {
        p = m;
        q = step % m;
    } else {
        p = (step + i) % m;
        q = (step - i + m) % m;
    }
    double* gp = G + (long long)p * np;
    double* gq = G + (long long)q * np;
    double a = 0.0, b = 0.0, g = 0.0;
    for (long long r = tid; r < np; r += 256) {
        const double x = gp[r], y = gq[r];
        a = fma(x, x, a);
        b = fma(y, y, b);
        g = fma(x, y, g);
    }
    a = warp_sum(a);
    b = warp_sum(b);
    g = warp_sum(g);
    if ((tid & 31) == 0) {
        red[0][tid >> 5] = a;
        red[1][tid >> 5] = b;
        red[2][tid >> 5] = g;
    }
    __syncthreads();
    a = b = g = 0.0;
    for (int w = 0; w < 8; ++w) {
        a += red[0][w];
        b += red[1][w];
        g += red[2][w];
    }
    if (g == 0.0 || fabs(g) <= 1.0e-15 * sqrt(a * b)) return;  // already orthogonal (CTA-uniform)
    if (tid == 0 && fabs(g) > 1.0e-13 * sqrt(a * b)) atomicOr(moved, 1);
    const double zeta = (b - a) / (2.0 * g);
    const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
    const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
    double* vp = V + (long long)p * np;
    double* vq = V + (long long)q * np;
    for (long long r = tid; r < np; r += 256) {
        const double x = gp[r], y = gq[r];
        gp[r] = c * x - sn * y;
        gq[r] = sn * x + c * y;
        const double vx = vp[r], vy = vq[r];
        vp[r] = c * vx - sn * vy;
        vq[r] = sn * vx + c * vy;
    }
}
// coef_j = (g_j . rhs) / s_j^2 for singular values above the pinv cut-off, else 0; sigma2[j] = s_j^2 (one CTA per column)
__global__ void __launch_bounds__(256) gk_svd_norm_kernel(const double* __restrict__ G, long long np, const double* __restrict__ rhs,
                                                          double* __restrict__ sigma2, double* __restrict__ dots, double* s2max) {
    __shared__ double red[2][8];
    const int j = blockIdx.x, tid = threadIdx.x;
    const double* gj = G + (long long)j * np;
    double a = 0.0, d = 0.0;
    for (long long r = tid; r < np; r += 256) {
        a = fma(gj[r], gj[r], a);
        d = fma(gj[r], rhs[r], d);
    }
    a = warp_sum(a);
    d = warp_sum(d);
    if ((tid & 31) == 0) {
        red[0][tid >> 5] = a;
        red[1][tid >> 5] = d;
    }
    __syncthreads();
    if (tid == 0) {
        a = d = 0.0;
        for (int w = 0; w < 8; ++w) {
            a += red[0][w];
            d += red[1][w];
        }
        sigma2[j] = a;
        dots[j] = d;
        atomic_max_double(s2max, a);
    }
}
// w_ext = V coef, coef_j = dots_j / sigma2_j where sigma_j > rtol * sigma_max (rtol = np * eps, scipy.linalg.pinv)
__global__ void __launch_bounds__(256) gk_svd_apply_kernel(const double* __restrict__ V, long long np, long long n_real,
                                                           const double* __restrict__ sigma2, const double* __restrict__ dots,
                                                           const double* s2max, double* __restrict__ w_ext) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= np) return;
    const double rtol = (double)n_real * 2.220446049250313e-16;
    const double cut2 = rtol * rtol * *s2max;
    double acc = 0.0;
    for (long long j = 0; j < np; ++j) {
        const double s2 = sigma2[j];
        if (s2 > cut2 && s2 > 0.0) acc = fma(V[j * np + r], dots[j] / s2, acc);
    }
    w_ext[r] = acc;
}

// ---- prediction --------------------------------------------------------------------------------------------------
// z_t = sum_i w_i b_ti + lambda, b_ti = -gamma(d_ti) (0 where d <= 1e-10 and exact_values); one thread per target
template <typename V>
__global__ void __launch_bounds__(256) gk_predict_kernel(const double* __restrict__ sx, const double* __restrict__ sy, long long n,
                                                         const double* __restrict__ w_ext, const double* __restrict__ t_x,
                                                         long long n_x, const double* __restrict__ t_y, long long n_y, int grid,
                                                         GkParams P, int exact_values, double out_scale, double out_shift,
                                                         V* __restrict__ out) {
    __shared__ double s_x[256], s_y[256], s_w[256];
    const long long M = grid ? n_x * n_y : n_x;
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double tx = 0.0, ty = 0.0;
    if (m < M) {
        if (grid) {
            const long long row = m / n_x;
            ty = t_y[row];
            tx = t_x[m - row * n_x];
        } else {
            tx = t_x[m];
            ty = t_y[m];
        }
    }
    double acc = 0.0;
    for (long long j0 = 0; j0 < n; j0 += 256) {
        __syncthreads();
        const long long j = j0 + threadIdx.x;
        s_x[threadIdx.x] = j < n ? sx[j] : 0.0;
        s_y[threadIdx.x] = j < n ? sy[j] : 0.0;
        s_w[threadIdx.x] = j < n ? w_ext[j] : 0.0;
        __syncthreads();
        const int cnt = (int)(n - j0 < 256 ? n - j0 : 256);
        if (m < M) {
            for (int c = 0; c < cnt; ++c) {
                const double d = gk_distance(P.metric, tx, ty, s_x[c], s_y[c]);
                double b = -gk_gamma(P.model, P.p0, P.p1, P.p2, d);
                if (exact_values && fabs(d) <= GK_EPS) b = 0.0;
                acc = fma(s_w[c], b, acc);
            }
        }
    }
    if (m < M) out[m] = (V)((acc + w_ext[n]) * out_scale + out_shift);
}

size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

struct GkLayout {
    size_t small, mat, aux, total;  // byte sizes: small arrays, C (or G + staging), V, sum
};
GkLayout gk_layout(long long n, int jacobi) {
    GkLayout L;
    const long long np = (n + 2) & ~1LL;  // bordered size padded to even
    L.small = align256(256) + align256((size_t)(np + 1) * 8) * 6;  // flags/gmax | rhs[2][n] (2 slots) | w_ext | sigma2 | dots | pad
    L.mat = jacobi ? align256((size_t)np * np * 8 * 2) : align256((size_t)n * n * 8);
    L.aux = jacobi ? align256((size_t)np * np * 8) : 0;
    L.total = L.small + L.mat + L.aux + 256;
    return L;
}

template <typename V>
int gk_entry(const double* s_x, const double* s_y, const double* z, long long n, const double* t_x, long long n_x,
             const double* t_y, long long n_y, int grid, int metric, int model, const double* params, int exact_values,
             int pseudo_inv, double out_scale, double out_shift, V* out, int32_t* status_host, void* workspace,
             size_t workspace_bytes, cudaStream_t stream) {
    if (!s_x || !s_y || !z || !t_x || !t_y || !params || !out || n < 1 || n_x < 0 || n_y < 0) return CMX_ERR_BAD_ARG;
    if (model < CMX_VARIOGRAM_LINEAR || model > CMX_VARIOGRAM_EXPONENTIAL) return CMX_ERR_UNSUPPORTED;
    if (metric != CMX_METRIC_EUCLIDEAN && metric != CMX_METRIC_GEOGRAPHIC) return CMX_ERR_UNSUPPORTED;
    if (!grid && n_x != n_y) return CMX_ERR_BAD_ARG;
    if (n > 0x3fffffffLL) return CMX_ERR_TOO_MANY;
    if (pseudo_inv && n + 1 > GK_JACOBI_MAX_N) return CMX_ERR_UNSUPPORTED;
    const long long M = grid ? n_x * n_y : n_x;
    if (status_host) *status_host = 0;
    const bool can_jacobi = n + 1 <= GK_JACOBI_MAX_N;
    // one layout serves both paths: room for the Jacobi fallback whenever it is possible
    const GkLayout L = gk_layout(n, can_jacobi ? 1 : 0);
    if (!workspace || workspace_bytes < L.total) return CMX_ERR_WORKSPACE;
    unsigned char* p = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
    const long long np = (n + 2) & ~1LL;
    int* flags = reinterpret_cast<int*>(p);                   // [0] cholesky failed, [1] jacobi moved
    double* gmax = reinterpret_cast<double*>(p + 64);         // [0] max gamma, [1] max sigma^2
    p += align256(256);
    const size_t slot = align256((size_t)(np + 1) * 8);
    double* rhs = reinterpret_cast<double*>(p);               // [2][n] (two slots)
    double* w_ext = reinterpret_cast<double*>(p + 2 * slot);  // [np]
    double* sigma2 = reinterpret_cast<double*>(p + 3 * slot);
    double* dots = reinterpret_cast<double*>(p + 4 * slot);
    double* rhs_b = reinterpret_cast<double*>(p + 5 * slot);  // bordered right-hand side [Z; 0; 0]
    p += 6 * slot;
    double* C = reinterpret_cast<double*>(p);
    double* Vm = reinterpret_cast<double*>(p + L.mat);

    GkParams P{model, metric, params[0], params[1], params[2]};
    CMX_CUDA(cudaMemsetAsync(flags, 0, 256, stream));  // flags and gmax / s2max
    const dim3 tri((unsigned)((n + 255) / 256), (unsigned)n);
    bool jacobi = pseudo_inv != 0;
    if (!jacobi) {
        // ---- Cholesky of C = c0 11' - Gamma ----
        gk_gamma_kernel<<<tri, 256, 0, stream>>>(s_x, s_y, n, P, C, n, gmax);
        gk_shift_kernel<<<tri, 256, 0, stream>>>(n, C, n, gmax);
        note_launch();
        note_launch();
        const long long nblk = (n + GB - 1) / GB;
        constexpr int SYRK_SMEM = 2 * GB * (GB + 2) * 8;
        CMX_CUDA(cudaFuncSetAttribute(gk_syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SYRK_SMEM));
        for (long long kb = 0; kb < nblk; ++kb) {
            gk_potrf_kernel<<<1, 256, 0, stream>>>(C, n, n, (int)kb, gmax, flags);
            note_launch();
            const long long below = nblk - kb - 1;
            if (below > 0) {
                gk_trsm_kernel<<<(unsigned)below, GB, 0, stream>>>(C, n, n, (int)kb);
                gk_syrk_kernel<<<(unsigned)(below * (below + 1) / 2), 256, SYRK_SMEM, stream>>>(C, n, n, (int)kb);
                note_launch();
                note_launch();
            }
        }
        int failed = 0;
        CMX_CUDA(cudaMemcpyAsync(&failed, flags, sizeof(int), cudaMemcpyDeviceToHost, stream));
        CMX_CUDA(cudaStreamSynchronize(stream));
        if (failed) {
            if (!can_jacobi) {
                if (status_host) *status_host = 2;  // numerically singular and too large for the SVD path
                return CMX_OK;
            }
            jacobi = true;
            if (status_host) *status_host = 1;
        } else {
            // rhs = [Z | 1]
            CMX_CUDA(cudaMemcpyAsync(rhs, z, (size_t)n * 8, cudaMemcpyDeviceToDevice, stream));
            gk_fill_kernel<<<64, 256, 0, stream>>>(rhs + n, 1.0, n);
            note_launch();
            // forward: rhs -> y (the two slots the Jacobi path uses for sigma2 | dots); backward: y -> rhs
            double* ysol = sigma2;
            static_assert(sizeof(double) == 8, "");
            for (long long kb = 0; kb < nblk; ++kb) {
                gk_forward_step_kernel<<<(unsigned)(nblk - kb), 128, 0, stream>>>(C, n, n, (int)kb, rhs, ysol);
                note_launch();
            }
            for (long long kb = nblk - 1; kb >= 0; --kb) {
                gk_backward_step_kernel<<<(unsigned)(kb + 1), 128, 0, stream>>>(C, n, n, (int)kb, ysol, rhs);
                note_launch();
            }
            gk_combine_kernel<<<1, 256, 0, stream>>>(n, rhs, w_ext);
            note_launch();
        }
    }
    if (jacobi) {
        // ---- one-sided Jacobi SVD of the bordered matrix ----
        double* G = C;  // [np][np] (+ a second np x np staging area behind it)
        CMX_CUDA(cudaMemsetAsync(gmax, 0, 16, stream));
        const dim3 trip((unsigned)((n + 255) / 256), (unsigned)n);
        gk_gamma_kernel<<<trip, 256, 0, stream>>>(s_x, s_y, n, P, G, np, gmax);
        const dim3 full((unsigned)((np + 255) / 256), (unsigned)np);
        gk_bordered_kernel<<<full, 256, 0, stream>>>(n, np, G, Vm);
        gk_copy_kernel<<<sm_count() * 4, 256, 0, stream>>>(G, G + np * np, np * np);
        for (int i = 0; i < 3; ++i) note_launch();
        CMX_CUDA(cudaMemsetAsync(rhs_b, 0, (size_t)np * 8, stream));
        CMX_CUDA(cudaMemcpyAsync(rhs_b, z, (size_t)n * 8, cudaMemcpyDeviceToDevice, stream));
        const int pairs = (int)(np / 2), steps = (int)np - 1;
        for (int sweep = 0; sweep < 40; ++sweep) {
            CMX_CUDA(cudaMemsetAsync(flags + 1, 0, sizeof(int), stream));
            for (int s = 0; s < steps; ++s) {
                gk_jacobi_kernel<<<pairs, 256, 0, stream>>>(G, Vm, np, s, flags + 1);
                note_launch();
            }
            int moved = 0;
            CMX_CUDA(cudaMemcpyAsync(&moved, flags + 1, sizeof(int), cudaMemcpyDeviceToHost, stream));
            CMX_CUDA(cudaStreamSynchronize(stream));
            if (!moved) break;
        }
        gk_svd_norm_kernel<<<(unsigned)np, 256, 0, stream>>>(G, np, rhs_b, sigma2, dots, gmax + 1);
        gk_svd_apply_kernel<<<(unsigned)((np + 255) / 256), 256, 0, stream>>>(Vm, np, n + 1, sigma2, dots, gmax + 1, w_ext);
        note_launch();
        note_launch();
    }
    if (M > 0) {
        gk_predict_kernel<V><<<(unsigned)((M + 255) / 256), 256, 0, stream>>>(s_x, s_y, n, w_ext, t_x, n_x, t_y, n_y, grid, P,
                                                                               exact_values, out_scale, out_shift, out);
        note_launch();
    }
    return check_launch();
}

}  // namespace
}  // namespace cmx

extern "C" {

size_t cmx_ok_global_workspace_bytes(int64_t n_samples) {
    if (n_samples < 1 || n_samples > 0x3fffffffLL) return 0;
    return cmx::gk_layout(n_samples, n_samples + 1 <= cmx::GK_JACOBI_MAX_N ? 1 : 0).total + 512;
}

int cmx_ok_global_f64(const double* s_x, const double* s_y, const double* z, int64_t n, const double* t_x, int64_t n_x,
                      const double* t_y, int64_t n_y, int grid, int metric, int model, const double params[3],
                      int exact_values, int pseudo_inv, double out_scale, double out_shift, double* out, int32_t* status,
                      void* ws, size_t ws_bytes, void* stream) {
    return cmx::gk_entry<double>(s_x, s_y, z, n, t_x, n_x, t_y, n_y, grid, metric, model, params, exact_values, pseudo_inv,
                                 out_scale, out_shift, out, status, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}
int cmx_ok_global_f32(const double* s_x, const double* s_y, const double* z, int64_t n, const double* t_x, int64_t n_x,
                      const double* t_y, int64_t n_y, int grid, int metric, int model, const double params[3],
                      int exact_values, int pseudo_inv, double out_scale, double out_shift, float* out, int32_t* status,
                      void* ws, size_t ws_bytes, void* stream) {
    return cmx::gk_entry<float>(s_x, s_y, z, n, t_x, n_x, t_y, n_y, grid, metric, model, params, exact_values, pseudo_inv,
                                out_scale, out_shift, out, status, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
