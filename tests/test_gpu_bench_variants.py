"""GPU (-m gpu): the kernel instantiations bench.py actually runs, at the sizes it runs them.

The headline numbers of BENCH/SCALE come from specific work decompositions of K1 chosen by the planner
(``cmx_knn_plan``): 16-row tiles without a sample split on the full C5 grid, 8-row tiles + 4 sample segments +
``knn_merge_kernel`` on one rank's band of an 8-GPU run, 16-row tiles + 2 segments on the second band of the host
pipeline.  Every test here first ASSERTS the plan, then compares neighbour tables (indices and distances, bit for
bit) and fields with scipy's cKDTree / the IDW oracle on sampled rows - so a change of the planner's choice fails
the suite instead of silently moving the benchmark onto an untested path.  Both searches, k = 32 and 64.
The kriging solve (K4) is checked at the C4 size (N = 50 000, k = 64) against the restated pykrige, and the
Cholesky solver against the pivoted one.
"""

import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import torch  # noqa: E402

from climatrix_b200 import _lib  # noqa: E402
from climatrix_b200 import kernels as K  # noqa: E402
from helpers import rel_err, synthetic_field  # noqa: E402
from oracle.idw_oracle import dense_points, idw_oracle, idw_oracle_batched, sparse_points  # noqa: E402
from oracle.ok_oracle import moving_window_with_parameters  # noqa: E402

DEV = "cuda:0"


def _plan(n, rows, cols, grid=1):
    lib = _lib.load()
    q, s = ctypes.c_int(0), ctypes.c_int(0)
    assert lib.cmx_knn_plan(n, rows, cols, grid, ctypes.byref(q), ctypes.byref(s)) == 0
    return q.value, s.value


@pytest.fixture(scope="module")
def c5_samples():
    """bench.py's C5 inputs: 10^6 samples, lat U(-90, 90), lon U(-180, 180), seeded."""
    rng = np.random.default_rng(1234 + 5)
    la, lo = rng.uniform(-90, 90, 1_000_000), rng.uniform(-180, 180, 1_000_000)
    v = synthetic_field(la, lo, rng)
    from scipy.spatial import cKDTree

    return la, lo, v, cKDTree(sparse_points(la, lo)), torch.as_tensor(la).to(DEV), torch.as_tensor(lo).to(DEV)


def _check_rows(tree, la, lo, v, t_lat, t_lon, rows, k, dist, idx, out=None):
    q = dense_points(t_lat[rows], t_lon)
    d0, i0 = tree.query(q, k=k, workers=-1)
    sel = (np.asarray(rows)[:, None] * len(t_lon) + np.arange(len(t_lon))[None, :]).ravel()
    sel_t = torch.as_tensor(sel, device=idx.device)
    np.testing.assert_array_equal(idx[sel_t].cpu().numpy(), i0)
    np.testing.assert_array_equal(dist[sel_t].cpu().numpy(), d0)
    if out is not None:
        ref = idw_oracle(sparse_points(la, lo), v, q, k=k, power=2.0, k_min=2)
        assert rel_err(out[sel_t].cpu().numpy(), ref) <= 1e-12


C5_LAT, C5_LON = np.linspace(-90.0, 90.0, 1801), -180.0 + np.arange(3600) * 0.1


@pytest.mark.parametrize("k", [32, 64])
def test_c5_full_grid_runs_knn_kernel_16_rows_no_split(c5_samples, k):
    """BENCH headline: knn_kernel<E, 16, true>, one sample segment (plan (16, 1))."""
    la, lo, v, tree, s_la, s_lo = c5_samples
    assert _plan(1_000_000, 1801, 3600) == (16, 1)
    T = K.Targets.make(C5_LAT, C5_LON, True, DEV)
    out, dist, idx = K.idw(s_la, s_lo, v, T, k=k, return_neighbours=True, search="brute")
    rows = [0, 1, 15, 16, 900, 1337, 1799, 1800]
    _check_rows(tree, la, lo, v, C5_LAT, C5_LON, rows, k, dist, idx, out)
    # size-independent properties over the whole 6.5 M-target field
    assert bool((dist[:, 1:] >= dist[:, :-1]).all())
    assert float(out.min()) >= v.min() - 1e-9 and float(out.max()) <= v.max() + 1e-9
    del out, dist, idx
    # the product default (cell-binned search) returns the same table
    dist_b, idx_b = K.knn(s_la, s_lo, T, k, search="binned")
    _check_rows(tree, la, lo, v, C5_LAT, C5_LON, rows, k, dist_b, idx_b)


@pytest.mark.parametrize("rows,plan", [(225, (8, 4)), (153, (16, 2)), (226, None), (450, None)])
@pytest.mark.parametrize("k", [32, 64])
def test_c5_bands_run_the_sample_split_and_merge_kernel(c5_samples, rows, plan, k):
    """One rank's band of an 8-GPU C5 run (225 / 226 rows), the second band of the host pipeline (153 rows) and a
    4-GPU band (450 rows): sample split (nseg > 1), partial lists, knn_merge_kernel."""
    la, lo, v, tree, s_la, s_lo = c5_samples
    got = _plan(1_000_000, rows, 3600)
    if plan is not None:
        assert got == plan
    assert got[1] > 1, f"band of {rows} rows no longer takes the sample-split path: plan {got}"
    r0 = 1801 - rows if rows == 153 else 3 * rows  # an interior band and the last band
    t_lat = C5_LAT[r0:r0 + rows]
    T = K.Targets.make(t_lat, C5_LON, True, DEV)
    out, dist, idx = K.idw(s_la, s_lo, v, T, k=k, return_neighbours=True, search="brute")
    check = [0, 7, 8, rows // 2, rows - 2, rows - 1]
    _check_rows(tree, la, lo, v, t_lat, C5_LON, check, k, dist, idx, out)
    # batched (table + K2) through the same split path
    vt = np.stack([v, 2.0 * v - 1.0, v * v], axis=1)
    outb = K.idw(s_la, s_lo, vt, T, k=k, layout="sample_major", search="brute")
    q = dense_points(t_lat[check[:2]], C5_LON)
    ref = idw_oracle_batched(sparse_points(la, lo), vt.T, q, k=k)
    sel = (np.asarray(check[:2])[:, None] * 3600 + np.arange(3600)[None, :]).ravel()
    assert rel_err(outb[:, torch.as_tensor(sel, device=outb.device)].cpu().numpy(), ref) <= 1e-12


def test_c3_plan_and_parity_one_slice():
    """C3 shape (10^5 samples -> 721 x 1440): the planner's small-tile choice (2, 1), both searches."""
    assert _plan(100_000, 721, 1440) == (2, 1)
    rng = np.random.default_rng(1234 + 3)
    la, lo = rng.uniform(-90, 90, 100_000), rng.uniform(0, 360, 100_000)
    v = synthetic_field(la, lo, rng)
    tl, to = np.linspace(90.0, -90.0, 721), np.arange(1440) * 0.25
    T = K.Targets.make(tl, to, True, DEV)
    from scipy.spatial import cKDTree

    tree = cKDTree(sparse_points(la, lo))
    for search in ("brute", "binned"):
        out, dist, idx = K.idw(la, lo, v, T, k=32, return_neighbours=True, search=search)
        _check_rows(tree, la, lo, v, tl, to, [0, 1, 2, 3, 360, 719, 720], 32, dist, idx, out)


# ------------------------------------------------------------------------------------------
# K4 at the C4 size, and the two solvers against each other
# ------------------------------------------------------------------------------------------
def test_c4_size_k64_solve_matches_restated_pykrige_on_sampled_rows():
    """BASELINE config 4: 50 000 samples -> 721 x 1440, k = 64, spherical: ok_chol_kernel<8> (+ hand-over list)."""
    rng = np.random.default_rng(1234 + 4)
    n = 50_000
    x, y = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    z = np.sin(3 * x) * np.cos(2 * y) + 0.3 * rng.normal(size=n)
    params = [1.05, 0.35, 0.08]  # psill, range, nugget: the magnitudes the C4 fit produces
    tx, ty = np.linspace(-1, 1, 1440), np.linspace(1, -1, 721)
    out, sig, nsing = K.ok_moving_window(x, y, z, tx, ty, True, k=64, variogram_model="spherical", params=params,
                                         return_sigmasq=True)
    assert int(nsing) == 0 and out.numel() == 721 * 1440 and bool(torch.isfinite(out).all())
    rows = [0, 360, 720]
    cols = np.arange(0, 1440, 5)
    gx, gy = np.meshgrid(tx[cols], ty[rows])
    zr, sr = moving_window_with_parameters(x, y, z, gx.ravel(), gy.ravel(), "points", k=64, variogram_model="spherical",
                                           parameters=params)
    sel = torch.as_tensor((np.asarray(rows)[:, None] * 1440 + cols[None, :]).ravel(), device=out.device)
    assert np.max(np.abs(out[sel].cpu().numpy() - zr)) <= 1e-4 * np.max(np.abs(zr))
    assert np.max(np.abs(sig[sel].cpu().numpy() - sr)) <= 1e-4 * np.max(np.abs(sr))


@pytest.mark.parametrize("model,params", [
    ("spherical", [1.1, 0.4, 0.1]), ("spherical", [1.0, 0.05, 0.0]), ("linear", [0.8, 0.05]), ("linear", [0.8, 0.0]),
    ("power", [0.7, 1.5, 0.02]), ("exponential", [0.9, 0.5, 0.0]), ("gaussian", [1.0, 0.3, 0.01]),
])
@pytest.mark.parametrize("k", [2, 3, 7, 16, 17, 31, 32, 33, 47, 48, 50, 63, 64])
def test_cholesky_solver_equals_pivoted_solver(model, params, k):
    """ok_chol_kernel<NB> (every capacity, full and partly filled) against Gauss-Jordan with partial pivoting."""
    rng = np.random.default_rng(100 * k + len(model))
    n = 4000
    x, y = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    z = np.sin(3 * x) + np.cos(2 * y) + 0.2 * rng.normal(size=n)
    tx, ty = rng.uniform(-1.1, 1.1, 3000), rng.uniform(-1.1, 1.1, 3000)
    tx[:50], ty[:50] = x[:50], y[:50]  # exact hits: the b = 0 rule
    d, idx = K.knn(y, x, K.Targets.make(ty, tx, False, DEV), k)
    d2 = d * d
    kw = dict(variogram_model=model, params=params, return_sigmasq=True, out_scale=2.5, out_shift=-1.0)
    a, sa, na = K.ok_solve(x, y, z, idx, d2, solver="auto", **kw)
    p, sp, npv = K.ok_solve(x, y, z, idx, d2, solver="pivoted", **kw)
    assert int(na) == 0 and int(npv) == 0
    scale = float(p.abs().max())
    assert float((a - p).abs().max()) <= 1e-9 * scale
    assert float((sa - sp).abs().max()) <= 1e-9 * max(float(sp.abs().max()), 1e-30)
    a32, s32, _ = K.ok_solve(x, y, z, idx, d2, solver="auto", dtype=torch.float32, **kw)
    assert a32.dtype == torch.float32 and float((a32.double() - a).abs().max()) <= 1e-6 * scale


def test_ill_conditioned_and_singular_windows_are_handed_to_the_pivoted_solver():
    """Gaussian model without nugget (cond ~ 1e17) and duplicate samples: the Cholesky path must give them up;
    results then ARE the pivoted solver's (bit for bit), singular windows are NaN and counted."""
    rng = np.random.default_rng(7)
    n = 3000
    x, y = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    x[1], y[1] = x[0], y[0]  # duplicate pair -> exactly singular windows around it (zero nugget)
    z = rng.normal(size=n)
    tx, ty = rng.uniform(-1, 1, 2000), rng.uniform(-1, 1, 2000)
    tx[0], ty[0] = x[0] + 1e-3, y[0]
    for model, params, k in [("gaussian", [1.0, 3.0, 0.0], 24), ("spherical", [1.0, 0.5, 0.0], 12), ("power", [0.7, 1.99, 0.0], 40)]:
        d, idx = K.knn(y, x, K.Targets.make(ty, tx, False, DEV), k)
        kw = dict(variogram_model=model, params=params, return_sigmasq=True)
        a, sa, na = K.ok_solve(x, y, z, idx, d * d, solver="auto", **kw)
        p, sp, npv = K.ok_solve(x, y, z, idx, d * d, solver="pivoted", **kw)
        assert int(na) == int(npv) >= 1
        nan_a, nan_p = torch.isnan(a), torch.isnan(p)
        assert torch.equal(nan_a, nan_p) and bool(nan_a[0])
        ok = ~nan_p
        if model == "spherical":  # well conditioned away from the duplicate: both solvers agree to rounding
            assert float((a[ok] - p[ok]).abs().max()) <= 1e-9 * float(p[ok].abs().max())
        else:  # ill-conditioned everywhere: (nearly) every window is handed over, so the results are identical
            same = (a[ok] == p[ok]) & (sa[ok] == sp[ok])
            assert float(same.double().mean()) >= 0.99
