/*
 * climatrix_b200 -- C ABI of the B200-native IDW / ordinary-kriging
 * reconstruction hot path (libclimatrix_b200.so, sm_100a).
 *
 * Drop-in boundary.  The reference (jamesWalczak/climatrix) is pure Python: its
 * hot path calls third-party native code through Python bindings.  Each entry
 * point below replaces one of those call sites; the Python host layer
 * (climatrix_b200/*.py) and any other FFI (ctypes stub in INTEGRATION.md) bind
 * exactly these symbols.
 *
 *   cmx_knn            <- scipy.spatial.cKDTree(points).query(q, k, workers=-1)
 *                         src/climatrix/reconstruct/idw.py:155-158
 *                         (and pykrige's moving-window tree.query)
 *   cmx_idw_f64/_f32   <- the kd-tree query + NumPy weighting epilogue
 *                         src/climatrix/reconstruct/idw.py:155-180
 *   cmx_idw_apply_*    <- idw.py:163-180 re-applied to a stored neighbour table
 *                         (HPO inner loop, src/climatrix/optim/bayesian.py:399-405)
 *   cmx_idw_score_*, cmx_score_* <- one HPO trial's score: bayesian.py:399-408 + comparison.py:268-306
 *   cmx_variogram_*    <- pykrige core._initialize_variogram_model (pdist + binning),
 *                         reached from src/climatrix/reconstruct/kriging.py:215-224
 *   cmx_ok_mw_f64/_f32 <- pykrige OrdinaryKriging.execute(backend="loop",
 *                         n_closest_points=k), call site kriging.py:236-241
 *   cmx_knn3, cmx_ok_solve_* <- the same call with coordinates_type="geographic"
 *   cmx_ok_global_*    <- the same call WITHOUT n_closest_points (global kriging, what climatrix ships)
 *   cmx_variogram_fit* <- pykrige core._calculate_variogram_model (scipy least_squares, soft_l1)
 *
 * Conventions
 *  - All data pointers are DEVICE pointers owned by the caller; nothing is
 *    allocated or freed inside.  Scratch comes from a caller-provided
 *    workspace (size from the matching *_workspace_bytes()).
 *  - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it.
 *  - Return value: CMX_OK (0) or a negative CMX_ERR_* code.  Nothing throws
 *    across the ABI.  cmx_strerror() maps codes to text.
 *  - Coordinates are always float64 (neighbour sets must match the reference's
 *    float64 kd-tree bit-exactly); `_f32`/`_f64` names the dtype of sample
 *    values and of the reconstructed field (IDW) / of the stored field (kriging,
 *    whose standardised values are always float64).
 *  - Targets: `target_is_grid != 0` -> dense grid, `t_a` is the latitude axis
 *    (n_a rows), `t_b` the longitude axis (n_b columns), target m = i*n_b + j
 *    (row-major, the order of DenseDomain.get_all_spatial_points,
 *    src/climatrix/dataset/domain.py:917-922).  Otherwise explicit points:
 *    `t_a[m], t_b[m]`, n_a == n_b == M (SparseDomain, domain.py:738).
 *  - Neighbour order and ties: ascending by (squared distance, sample index),
 *    squared distance = dx*dx + dy*dy in float64 without FMA contraction.
 *  - No hidden state: everything that changes what a call does is an argument.  Optional behaviour
 *    (search algorithm, output redirection to peer devices) travels in a `const cmx_options*`
 *    (NULL = defaults); the library keeps no per-thread or per-process mode.
 */
#ifndef CLIMATRIX_B200_H
#define CLIMATRIX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CMX_OK 0
#define CMX_ERR_BAD_ARG (-1)        /* null pointer, negative size, k < 1 ...            */
#define CMX_ERR_K_GT_N (-2)         /* more neighbours requested than samples             */
#define CMX_ERR_K_TOO_LARGE (-3)    /* k above CMX_MAX_K                                  */
#define CMX_ERR_WORKSPACE (-4)      /* workspace missing or too small                     */
#define CMX_ERR_CUDA (-5)           /* a CUDA runtime call or launch failed               */
#define CMX_ERR_UNSUPPORTED (-6)    /* option not implemented (e.g. unknown variogram)    */
#define CMX_ERR_TOO_MANY (-7)       /* more than 2^31-1 samples or targets                */

#define CMX_MAX_K 128
#define CMX_MAX_PEERS 8

/* ---- per-call options (NULL = all defaults) ----
 *
 * search: neighbour-search algorithm of cmx_knn / cmx_idw_* / cmx_ok_mw_*.  All return the same neighbour tables
 * (exact float64 distances, ties by sample index); they differ only in the work done:
 *   CMX_SEARCH_BRUTE   K1: every (target, sample) pair is evaluated (north_star's brute-force kernel).
 *   CMX_SEARCH_BINNED  K1c: samples are counting-sorted into the cells of the seed grid and each target scans
 *                      only the cells that intersect the disc of its k-th-distance bound - the device counterpart
 *                      of cKDTree's pruning (reference: reconstruct/idw.py:155-158).  Needs 20 B per sample of
 *                      the ordinary kNN workspace; if that does not fit, the call uses K1.
 *   CMX_SEARCH_AUTO    (default) the faster one for the problem: K1c whenever its scratch fits.
 *
 * peer_out / n_peers / m_total / m_offset: multi-GPU fused weighting + gather (batched IDW only, n_batch > 1).
 * With one process per GPU every rank reconstructs a band of target rows and all ranks need the whole field
 * (reference: one cm.reconstruct call returns the full grid).  Instead of a separate all-gather after the kernel,
 * the batched weighting kernels store each finished row segment directly into the [n_batch][m_total] output
 * arrays of n_peers devices (peer-mapped pointers, e.g. torch symmetric memory; NVLink stores through NVSwitch)
 * at this call's first target m_offset.  The call's own `out` argument is then unused (but must be non-null).
 * n_peers = 0: plain output.  Synchronising the ranks before the buffers are read is the caller's job.
 * A redirected call with n_batch == 1 returns CMX_ERR_UNSUPPORTED.
 *
 * params_on_device: see the kriging entry points.
 */
#define CMX_SEARCH_AUTO 0
#define CMX_SEARCH_BRUTE 1
#define CMX_SEARCH_BINNED 2
typedef struct cmx_options {
    int search;
    int n_peers;
    void* peer_out[CMX_MAX_PEERS];
    int64_t m_total;
    int64_t m_offset;
    int params_on_device; /* cmx_ok_mw_* / cmx_ok_solve_*: `params` is a DEVICE pointer to 3 doubles (e.g. written by
                             cmx_variogram_fit on the same stream) instead of a host array: no host round trip */
} cmx_options;

/* variogram models, pykrige/variogram_models.py (params: see cmx_ok_mw_*) */
#define CMX_VARIOGRAM_LINEAR 0
#define CMX_VARIOGRAM_POWER 1
#define CMX_VARIOGRAM_GAUSSIAN 2
#define CMX_VARIOGRAM_SPHERICAL 3
#define CMX_VARIOGRAM_EXPONENTIAL 4

/* distance metric of the kriging frame (pykrige coordinates_type) */
#define CMX_METRIC_EUCLIDEAN 0
#define CMX_METRIC_GEOGRAPHIC 1

/* layout of a [batch x samples] value matrix */
#define CMX_LAYOUT_BATCH_MAJOR 0  /* vals[t * n + i]  (time, point)  */
#define CMX_LAYOUT_SAMPLE_MAJOR 1 /* vals[i * n_batch + t]  (point, time) -- the fast one */

int cmx_version(void);
const char* cmx_strerror(int code);
/* last CUDA error text recorded by this library on the calling thread ("" if none) */
const char* cmx_last_cuda_error(void);

/* ---- k nearest neighbours (brute force, fp32 filter + fp64 exact re-rank) ---- */

/* Scratch bytes cmx_knn / cmx_idw_* / cmx_ok_mw_* need for `k` on the current device. */
size_t cmx_knn_workspace_bytes(int k);
/*
 * Same, plus room for the partial-list tables of the sample split the library uses when the
 * targets alone cannot fill the GPU (e.g. one rank's band of a multi-GPU job).  Passing only
 * cmx_knn_workspace_bytes() is always valid; the split is then not used.
 */
size_t cmx_knn_workspace_bytes_for(int k, int64_t n_samples, int64_t n_a, int64_t n_b, int target_is_grid);

/*
 * idx_out  [M][k] int32   sample indices, ascending by (dist, index)      (required)
 * dist_out [M][k] float64 Euclidean distances sqrt(dx*dx+dy*dy)           (may be NULL)
 * Replaces cKDTree(points).query(queries, k): idw.py:155-158.
 */
int cmx_knn(const double* s_a, const double* s_b, int64_t n_samples,
            const double* t_a, int64_t n_a, const double* t_b, int64_t n_b, int target_is_grid,
            int k, int32_t* idx_out, double* dist_out,
            void* workspace, size_t workspace_bytes, const cmx_options* options, void* stream);

/* ---- inverse distance weighting ---- */

/* Extra scratch (beyond cmx_knn_workspace_bytes) for n_batch > 1: the [M][k] weight table. */
size_t cmx_idw_workspace_bytes(int k, int64_t n_targets, int64_t n_batch);

/*
 * out[t*M + m] = IDW estimate of slice t at target m (NaN where fewer than k_min
 * finite neighbours): idw.py:155-180, looped over the batch (the reference rejects
 * dynamic datasets, idw.py:87-104, so a batch is n_batch static calls).
 * idx_out/dist_out: optional [M][k] neighbour table (NULL to skip).
 */
int cmx_idw_f64(const double* s_lat, const double* s_lon, int64_t n_samples,
                const double* vals, int64_t n_batch, int vals_layout,
                const double* t_lat, int64_t n_lat, const double* t_lon, int64_t n_lon, int target_is_grid,
                int k, int k_min, double power,
                double* out, int32_t* idx_out, double* dist_out,
                void* workspace, size_t workspace_bytes, const cmx_options* options, void* stream);
int cmx_idw_f32(const double* s_lat, const double* s_lon, int64_t n_samples,
                const float* vals, int64_t n_batch, int vals_layout,
                const double* t_lat, int64_t n_lat, const double* t_lon, int64_t n_lon, int target_is_grid,
                int k, int k_min, double power,
                float* out, int32_t* idx_out, double* dist_out,
                void* workspace, size_t workspace_bytes, const cmx_options* options, void* stream);

/*
 * Re-apply idw.py:163-180 to a stored neighbour table with other (k_use <= k_table,
 * power, k_min): the hyper-parameter search inner loop (optim/bayesian.py:399-405).
 * flag_scratch: optional 4 bytes of device scratch; when given (and n_batch > 1) one pass over
 * the values records whether any is non-finite so that the batched kernel can skip the
 * per-value test; NULL -> every value is tested.
 */
int cmx_idw_apply_f64(const int32_t* idx, const double* dist, int64_t n_targets, int k_table,
                      const double* vals, int64_t n_samples, int64_t n_batch, int vals_layout,
                      int k_use, int k_min, double power, double* out, int32_t* flag_scratch,
                      const cmx_options* options, void* stream);
int cmx_idw_apply_f32(const int32_t* idx, const double* dist, int64_t n_targets, int k_table,
                      const float* vals, int64_t n_samples, int64_t n_batch, int vals_layout,
                      int k_use, int k_min, double power, float* out, int32_t* flag_scratch,
                      const cmx_options* options, void* stream);

/*
 * Score of one hyper-parameter trial against held-out values: the loop body of climatrix's HParamFinder,
 * src/climatrix/optim/bayesian.py:380-425 (reconstruct train -> val.domain, then Comparison.compute) with the
 * metrics of src/climatrix/comparison.py:268-306 (nanmean of |d|, d^2 over d = predicted - true).
 *   sums (device, 3 doubles): sum |d|, sum d^2, number of targets where prediction and truth are both finite;
 *                              MAE = sums[0]/sums[2], MSE = sums[1]/sums[2], RMSE = sqrt(MSE).
 *   cmx_idw_score_*: weighting of a stored neighbour table (as cmx_idw_apply_*, one slice) with the error reduction
 *                    fused into the same kernel; `out` may be NULL - the field is then never written.
 *   cmx_score_*:     the reduction alone, for a field that exists (ordinary-kriging trials).
 * workspace: cmx_score_workspace_bytes(n_targets).  Deterministic (fixed partial slots, fixed order).
 */
size_t cmx_score_workspace_bytes(int64_t n_targets);
int cmx_idw_score_f64(const int32_t* idx, const double* dist, int64_t n_targets, int k_table,
                      const double* vals, int64_t n_samples, int k_use, int k_min, double power,
                      const double* truth, double* out, double* sums,
                      void* workspace, size_t workspace_bytes, void* stream);
int cmx_idw_score_f32(const int32_t* idx, const double* dist, int64_t n_targets, int k_table,
                      const float* vals, int64_t n_samples, int k_use, int k_min, double power,
                      const float* truth, float* out, double* sums,
                      void* workspace, size_t workspace_bytes, void* stream);
int cmx_score_f64(const double* pred, const double* truth, int64_t n, double* sums,
                  void* workspace, size_t workspace_bytes, void* stream);
int cmx_score_f32(const float* pred, const float* truth, int64_t n, double* sums,
                  void* workspace, size_t workspace_bytes, void* stream);

/* ---- empirical variogram (pykrige core._initialize_variogram_model) ---- */

/*
 * The pair set can be split over several calls / devices: the call handles share `part` of `n_parts` of the pair
 * blocks (part = 0, n_parts = 1: everything); partial minima / maxima / sums combine by min / max / addition
 * (one rank per GPU: an all-reduce of 2 + 3 * nlags numbers).
 * minmax[0] = min, minmax[1] = max pair distance over this call's i>j pairs (device, 2 doubles; +inf / -inf
 * when the share is empty).
 */
int cmx_variogram_minmax(const double* x, const double* y, int64_t n, int metric, int part, int n_parts,
                         double* minmax, void* stream);
/*
 * edges[nlags+1] (device): bin n holds pairs with edges[n] <= d < edges[n+1].
 * sum_d / sum_g / count [nlags] (device) are ACCUMULATED into (zero them first);
 * g = 0.5 * (z_i - z_j)^2.  nlags <= 64 (CMX_ERR_UNSUPPORTED above; climatrix's hyper-parameter range is 4..20).
 * workspace: cmx_variogram_workspace_bytes(nlags) bytes of scratch make the reduction deterministic (per-CTA partial
 * sums combined in a fixed order: reproducible fits); NULL -> floating-point atomics (sums vary in the last bits
 * from run to run).
 */
size_t cmx_variogram_workspace_bytes(int nlags);
int cmx_variogram_bins(const double* x, const double* y, const double* z, int64_t n, int metric, int part,
                       int n_parts, int nlags, const double* edges,
                       double* sum_d, double* sum_g, unsigned long long* count,
                       void* workspace, size_t workspace_bytes, void* stream);

/*
 * Variogram-model fit = pykrige core._calculate_variogram_model (weight=False):
 * scipy.optimize.least_squares(residuals, x0, bounds, loss="soft_l1") with pykrige's start point and bounds, i.e. the
 * point where SciPy's Trust Region Reflective iteration stops at its default tolerances (which is what the reference
 * obtains; it is not the exact minimiser).  The iteration is restated in csrc/variogram_fit.cuh.
 *  cmx_variogram_fit_host: plain host function (no device): n <= 64 (lag, semivariance) points -> params_out[3]
 *                          (linear: slope, nugget, 0).  Returns the termination status (1 gtol, 2 ftol, 3 xtol, 4 both,
 *                          0 evaluation budget spent) or a negative CMX_ERR_*.
 *  cmx_variogram_edges:    bin edges on the device from the device-resident minmax, with pykrige's arithmetic.
 *  cmx_variogram_fit:      the same fit on the device from the accumulated bins (empty bins dropped, as pykrige does):
 *                          params_out[3]; optional lags_out / semivariance_out [nlags] (used entries first) and
 *                          info[3] = {status, points, function evaluations}.  No host round trip between the variogram
 *                          and the kriging solves (cmx_options.params_on_device).
 */
int cmx_variogram_fit_host(int variogram_model, const double* lags, const double* semivariance, int n, double params_out[3],
                           int* nfev_out);
int cmx_variogram_edges(const double* minmax, int nlags, double* edges, void* stream);
int cmx_variogram_fit(int variogram_model, int nlags, const double* sum_d, const double* sum_g,
                      const unsigned long long* count, double* params_out, double* lags_out, double* semivariance_out,
                      int32_t* info, void* stream);

/* ---- ordinary kriging, moving window ---- */

/*
 * Per target: k nearest samples in the kriging frame (x, y already normalised and
 * anisotropy-adjusted; CMX_METRIC_GEOGRAPHIC: x = lon, y = lat in degrees, neighbours
 * by 3-D chord, distances great-circle degrees), assemble the (k+1)x(k+1) system
 * A = [[-gamma(D), 1], [1^T, 0]] (zero diagonal), b = [-gamma(d); 1] with entries
 * forced to 0 where d <= 1e-10 (exact_values), solve in float64,
 * z = sum_i x_i Z_i, sigmasq = -x.b.   pykrige ok.py _exec_loop_moving_window, reached from
 * kriging.py:236-241 (which solves by LU with partial pivoting).
 * Solver: k <= 64 -> Cholesky of the equivalent positive-definite system A + c0 11^T in registers
 * (DESIGN.md K4); windows whose pivots fall below 1e-9 c0 (cond > ~1e9), k > 64 and k == 1 -> Gauss-Jordan
 * with partial pivoting, which also defines the singular case.
 * params: linear [slope, nugget, -]; power [scale, exponent, nugget];
 *         gaussian/spherical/exponential [psill, range, nugget].
 * z [N] is float64 in both variants (kriging.py:207-214 standardises the values in float64); `_f32`/`_f64`
 * names the dtype the field and the variance are STORED in.  out[m] = z * out_scale + out_shift: the
 * de-standardisation of kriging.py:245 fused into the store (pass 1, 0 for plain pykrige output).
 * out [M] (required), sigmasq_out [M] (may be NULL); NaN + *n_singular++ where the
 * local system is singular (duplicate neighbours; the reference raises LinAlgError).
 */
int cmx_ok_mw_f64(const double* s_x, const double* s_y, const double* z, int64_t n_samples,
                  const double* t_x, int64_t n_x, const double* t_y, int64_t n_y, int target_is_grid,
                  int metric, int k, int variogram_model, const double params[3], int exact_values,
                  double out_scale, double out_shift,
                  double* out, double* sigmasq_out, int32_t* n_singular,
                  void* workspace, size_t workspace_bytes, const cmx_options* options, void* stream);
int cmx_ok_mw_f32(const double* s_x, const double* s_y, const double* z, int64_t n_samples,
                  const double* t_x, int64_t n_x, const double* t_y, int64_t n_y, int target_is_grid,
                  int metric, int k, int variogram_model, const double params[3], int exact_values,
                  double out_scale, double out_shift,
                  float* out, float* sigmasq_out, int32_t* n_singular,
                  void* workspace, size_t workspace_bytes, const cmx_options* options, void* stream);
/*
 * Geographic mode (pykrige coordinates_type="geographic"), in two calls:
 *  cmx_knn3       exact 3-D k-NN between unit vectors (cKDTree on (cos lon cos lat, sin lon cos lat,
 *                 sin lat)); explicit target points; writes sample
 *                 indices and SQUARED chord lengths, rows ascending by (d2, index).  No workspace.
 *  cmx_ok_solve_* K4 on an existing neighbour table.  metric = CMX_METRIC_GEOGRAPHIC: d2 holds squared
 *                 chords (converted by 180 - 360/pi acos(chord/2)), s_x = lon, s_y = lat in degrees and
 *                 the window matrix uses the great-circle (Vincenty atan2) distance;
 *                 CMX_METRIC_EUCLIDEAN: d2 holds squared Euclidean distances in the (s_x, s_y) frame.
 *                 workspace: cmx_ok_solve_workspace_bytes(n_targets) bytes of scratch for the list of windows
 *                 the Cholesky solver hands to the pivoted one; NULL (or too small) -> every window is solved
 *                 by the pivoted Gauss-Jordan kernel (same results within rounding; ~5x slower at k = 64).
 */
int cmx_knn3(const double* s_x, const double* s_y, const double* s_z, int64_t n_samples,
             const double* t_x, const double* t_y, const double* t_z, int64_t n_targets,
             int k, int32_t* idx_out, double* d2_out, void* stream);
size_t cmx_ok_solve_workspace_bytes(int64_t n_targets);
int cmx_ok_solve_f64(const double* s_x, const double* s_y, const double* z, int64_t n_samples,
                     const int32_t* idx, const double* d2, int64_t n_targets, int metric, int k,
                     int variogram_model, const double params[3], int exact_values,
                     double out_scale, double out_shift,
                     double* out, double* sigmasq_out, int32_t* n_singular,
                     void* workspace, size_t workspace_bytes, const cmx_options* options, void* stream);
int cmx_ok_solve_f32(const double* s_x, const double* s_y, const double* z, int64_t n_samples,
                     const int32_t* idx, const double* d2, int64_t n_targets, int metric, int k,
                     int variogram_model, const double params[3], int exact_values,
                     double out_scale, double out_shift,
                     float* out, float* sigmasq_out, int32_t* n_singular,
                     void* workspace, size_t workspace_bytes, const cmx_options* options, void* stream);

/* ---- ordinary kriging, global (the mode climatrix ships: method="ok" without a neighbour count) ----
 *
 * pykrige OrdinaryKriging.execute(style, x, y, backend="vectorized"|"loop") without n_closest_points, reached from
 * kriging.py:236-241: one (N+1) x (N+1) system for all targets.  climatrix keeps only the estimate, and
 * z_t = b_t' A^-1 [Z; 0], so ONE solve A [w; lambda] = [Z; 0] serves every target (z_t = sum_i w_i b_ti + lambda):
 * blocked Cholesky of the equivalent positive-definite matrix A + c0 11' (hand-written FP64 kernels);
 * pseudo_inv != 0 (scipy.linalg.pinv semantics: singular values <= (N+1) eps s_max dropped) or a numerically singular
 * matrix -> one-sided Jacobi SVD of the bordered matrix, available for N + 1 <= 4096 (CMX_ERR_UNSUPPORTED beyond
 * when pseudo_inv is requested).
 * Coordinates in the kriging frame like cmx_ok_mw_*; metric CMX_METRIC_GEOGRAPHIC uses great-circle distances.
 * *status (HOST int, may be NULL; the call synchronises the stream once to read the factorisation flag):
 *   0 the requested path ran (Cholesky, or the SVD when pseudo_inv != 0), 1 the Cholesky met a non-positive pivot and the
 *   call fell back to the SVD path, 2 numerically singular and too large for the SVD path (out untouched).
 */
size_t cmx_ok_global_workspace_bytes(int64_t n_samples);
int cmx_ok_global_f64(const double* s_x, const double* s_y, const double* z, int64_t n_samples,
                      const double* t_x, int64_t n_x, const double* t_y, int64_t n_y, int target_is_grid,
                      int metric, int variogram_model, const double params[3], int exact_values, int pseudo_inv,
                      double out_scale, double out_shift, double* out, int32_t* status,
                      void* workspace, size_t workspace_bytes, void* stream);
int cmx_ok_global_f32(const double* s_x, const double* s_y, const double* z, int64_t n_samples,
                      const double* t_x, int64_t n_x, const double* t_y, int64_t n_y, int target_is_grid,
                      int metric, int variogram_model, const double params[3], int exact_values, int pseudo_inv,
                      double out_scale, double out_shift, float* out, int32_t* status,
                      void* workspace, size_t workspace_bytes, void* stream);

/* total scratch for cmx_ok_mw_* (includes the kNN scratch and cmx_ok_solve_workspace_bytes) */
size_t cmx_ok_workspace_bytes(int k, int64_t n_targets);

/* Work decomposition cmx_knn (CMX_SEARCH_BRUTE) would use for these sizes on the current device: targets per thread
 * (tile height in rows: 16, 8, 2 or 1) and the number of sample segments (1 = no split).  Host-only query, for
 * benchmarks and tests of the planner (DESIGN.md K1 "Planner").  Returns CMX_OK or CMX_ERR_BAD_ARG. */
int cmx_knn_plan(int64_t n_samples, int64_t n_a, int64_t n_b, int grid, int* tile_rows, int* sample_segments);

/* ---- host transfer helper ----
 * Asynchronous strided device -> host copy (cudaMemcpy2DAsync): `rows` rows of `width_bytes` bytes from a device
 * array with row pitch src_pitch into a (page-locked) host array with row pitch dst_pitch.  Lets a band of target
 * rows [T][M_band] land in its columns of the caller's [T][M] result while the next band is still computing.
 */
int cmx_copy_rows_to_host(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width_bytes,
                          size_t rows, void* stream);

/* ---- instrumentation ---- */

/* Kernels launched by this library on the calling thread since the last reset. */
int64_t cmx_launch_count(void);
void cmx_reset_launch_count(void);

/*
 * Optional per-kernel device timing (bench.py's roofline): when enabled, the library brackets
 * its two hot kernels with CUDA events on the launching stream.
 *   kind 0 = K1 neighbour search (knn_kernel), kind 1 = K2 batched IDW apply, kind 2 = K4 kriging solve
 *   (Cholesky kernel + the pivoted kernel on its hand-over list), kind 3 = K1c cell-binned neighbour search
 *   (scatter + knn_binned_kernel).
 * cmx_timing_read synchronises on the recorded events and returns the accumulated
 * milliseconds and the number of launches since the last cmx_timing_reset.
 */
#define CMX_KERNEL_KNN 0
#define CMX_KERNEL_IDW_BATCH 1
#define CMX_KERNEL_OK_SOLVE 2
#define CMX_KERNEL_KNN_BINNED 3
void cmx_timing_enable(int on);
void cmx_timing_reset(void);
int cmx_timing_read(int kind, double* total_ms, int64_t* launches);

/* FP32 FFMA throughput of the current device, measured now (TFLOP/s, 2 flop per FMA);
 * the denominator of the K1 roofline.  Synchronous; < 0 on error. */
double cmx_fp32_peak_tflops(void);
/* FP64 DFMA throughput of the current device, measured now (TFLOP/s); the denominator of the K4 roofline. */
double cmx_fp64_peak_tflops(void);

#ifdef __cplusplus
}
#endif
#endif /* CLIMATRIX_B200_H */
