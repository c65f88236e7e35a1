"""CPU restatement of PyKrige's ordinary-kriging algorithm.  **Parity unpinned.**

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

climatrix's ``method="ok"`` hands everything to the third-party package
``pykrige`` (``/root/reference/src/climatrix/reconstruct/kriging.py:191,215-241``).
``pykrige`` is an un-vendored, un-pinned dependency (``pyproject.toml:65-67``:
bare ``"pykrige"``; upstream GeoStat-Framework/PyKrige, 1.7.x line) and cannot be
installed in this image (no network).  This module restates its published
algorithm -- ``pykrige/ok.py`` (``OrdinaryKriging.__init__``,
``_get_kriging_matrix``, ``_exec_vector``, ``_exec_loop``,
``_exec_loop_moving_window``, ``execute``), ``pykrige/core.py``
(``_adjust_for_anisotropy``, ``_initialize_variogram_model``,
``_calculate_variogram_model``, ``great_circle_distance``,
``euclid3_to_great_circle``) and ``pykrige/variogram_models.py`` -- on the same
SciPy primitives (``pdist``/``cdist``, ``least_squares(loss="soft_l1")``,
``scipy.linalg.inv/solve``, ``cKDTree``) so that the same LAPACK / TRF code
paths execute.  The reference repository holds no golden numbers for it
(SURVEY.md section 8c), hence "parity unpinned".

Only the constructor arguments climatrix passes are supported; everything else
is fixed at pykrige's defaults (``weight=False``, ``anisotropy_angle=0``,
``exact_values=True``, ``pseudo_inv_type="pinv"``).

Two deliberate deviations, both numerically neutral, so the oracle stays
runnable at the BASELINE sizes:
* the empirical variogram is accumulated over row blocks of the pair matrix
  instead of materialising ``pdist`` (N=50 000 would need 10 GB per array);
  bin membership uses the identical float64 predicate;
* the moving-window loop assembles each (k+1)x(k+1) block directly instead of
  slicing it out of the global (N+1)x(N+1) matrix (20 GB at N=50 000);
  ``a_all[sel][:, sel]`` holds exactly these entries.
"""

from __future__ import annotations

import warnings

import numpy as np
import scipy.linalg
from scipy.optimize import least_squares
from scipy.spatial import cKDTree
from scipy.spatial.distance import cdist

EPS = 1.0e-10  # pykrige ok.py: class attribute ``eps``


# --------------------------------------------------------------------------- #
# variogram_models.py
# --------------------------------------------------------------------------- #
def linear_variogram_model(m, d):
    slope, nugget = float(m[0]), float(m[1])
    return slope * d + nugget


def power_variogram_model(m, d):
    scale, exponent, nugget = float(m[0]), float(m[1]), float(m[2])
    return scale * d**exponent + nugget


def gaussian_variogram_model(m, d):
    psill, range_, nugget = float(m[0]), float(m[1]), float(m[2])
    return psill * (1.0 - np.exp(-(d**2.0) / (range_ * 4.0 / 7.0) ** 2.0)) + nugget


def exponential_variogram_model(m, d):
    psill, range_, nugget = float(m[0]), float(m[1]), float(m[2])
    return psill * (1.0 - np.exp(-d / (range_ / 3.0))) + nugget


def spherical_variogram_model(m, d):
    psill, range_, nugget = float(m[0]), float(m[1]), float(m[2])
    d = np.asarray(d, dtype=np.float64)
    return np.piecewise(
        d,
        [d <= range_, d > range_],
        [
            lambda x: psill * ((3.0 * x) / (2.0 * range_) - (x**3.0) / (2.0 * range_**3.0)) + nugget,
            psill + nugget,
        ],
    )


VARIOGRAM_FUNCTIONS = {
    "linear": linear_variogram_model,
    "power": power_variogram_model,
    "gaussian": gaussian_variogram_model,
    "spherical": spherical_variogram_model,
    "exponential": exponential_variogram_model,
}


# --------------------------------------------------------------------------- #
# core.py
# --------------------------------------------------------------------------- #
def great_circle_distance(lon1, lat1, lon2, lat2):
    """Vincenty (atan2) great-circle distance in degrees; inputs in degrees."""
    lat1 = np.array(lat1) * np.pi / 180.0
    lat2 = np.array(lat2) * np.pi / 180.0
    dlon = (lon1 - lon2) * np.pi / 180.0
    c1, s1 = np.cos(lat1), np.sin(lat1)
    c2, s2 = np.cos(lat2), np.sin(lat2)
    cd = np.cos(dlon)
    return (
        180.0
        / np.pi
        * np.arctan2(
            np.sqrt((c2 * np.sin(dlon)) ** 2 + (c1 * s2 - s1 * c2 * cd) ** 2),
            s1 * s2 + c1 * c2 * cd,
        )
    )


def euclid3_to_great_circle(euclid3_distance):
    """Chord length on the unit sphere -> great-circle distance in degrees."""
    return 180.0 - 360.0 / np.pi * np.arccos(0.5 * euclid3_distance)


def adjust_for_anisotropy(X, center, scaling, angle_deg):
    """2-D ``_adjust_for_anisotropy``: centre, rotate by -angle, stretch y, un-centre."""
    center = np.asarray(center, dtype=np.float64)[None, :]
    angle = angle_deg * np.pi / 180.0
    X = np.array(X, dtype=np.float64, copy=True)
    X -= center
    stretch = np.array([[1.0, 0.0], [0.0, scaling]])
    rot = np.array([[np.cos(-angle), -np.sin(-angle)], [np.sin(-angle), np.cos(-angle)]])
    X_adj = np.dot(stretch, np.dot(rot, X.T)).T
    X_adj += center
    return X_adj


def _pair_blocks(n, block):
    for s in range(0, n, block):
        yield s, min(n, s + block)


def empirical_variogram(X, z, nlags, coordinates_type, block=2048):
    """``_initialize_variogram_model`` binning step, accumulated block-wise.

    Returns ``(lags, semivariance)`` with empty bins dropped, plus the raw
    ``(dmin, dmax, bins, sum_d, sum_g, count)`` for kernel-level parity checks.
    """
    n = X.shape[0]
    if n < 2:
        raise ValueError("need at least two samples for a variogram")

    def pair_block(s, e):
        # strictly-lower-triangle pairs (i > j) with i in [s, e)
        if coordinates_type == "euclidean":
            d = cdist(X[s:e], X[:e], "euclidean")
        else:
            d = great_circle_distance(X[s:e, 0][:, None], X[s:e, 1][:, None], X[None, :e, 0], X[None, :e, 1])
        g = 0.5 * (z[s:e, None] - z[None, :e]) ** 2.0
        ii = np.arange(s, e)[:, None]
        jj = np.arange(0, e)[None, :]
        keep = ii > jj
        return d[keep], g[keep]

    dmax, dmin = -np.inf, np.inf
    for s, e in _pair_blocks(n, block):
        d, _ = pair_block(s, e)
        if d.size:
            dmax = max(dmax, float(np.amax(d)))
            dmin = min(dmin, float(np.amin(d)))
    dd = (dmax - dmin) / nlags
    bins = [dmin + i * dd for i in range(nlags)]
    dmax_edge = dmax + 0.001
    bins.append(dmax_edge)

    sum_d = np.zeros(nlags)
    sum_g = np.zeros(nlags)
    count = np.zeros(nlags, dtype=np.int64)
    for s, e in _pair_blocks(n, block):
        d, g = pair_block(s, e)
        for b in range(nlags):
            sel = (d >= bins[b]) & (d < bins[b + 1])
            c = int(sel.sum())
            if c:
                sum_d[b] += d[sel].sum()
                sum_g[b] += g[sel].sum()
                count[b] += c
    with np.errstate(invalid="ignore", divide="ignore"):
        lags = np.where(count > 0, sum_d / count, np.nan)
        semivariance = np.where(count > 0, sum_g / count, np.nan)
    good = ~np.isnan(semivariance)
    raw = dict(dmin=dmin, dmax=dmax, bins=np.array(bins), sum_d=sum_d, sum_g=sum_g, count=count)
    return lags[good], semivariance[good], raw


def fit_variogram(lags, semivariance, variogram_model):
    """``_calculate_variogram_model`` (weight=False)."""
    fn = VARIOGRAM_FUNCTIONS[variogram_model]
    if variogram_model == "linear":
        x0 = [
            (np.amax(semivariance) - np.amin(semivariance)) / (np.amax(lags) - np.amin(lags)),
            np.amin(semivariance),
        ]
        bnds = ([0.0, 0.0], [np.inf, np.amax(semivariance)])
    elif variogram_model == "power":
        x0 = [
            (np.amax(semivariance) - np.amin(semivariance)) / (np.amax(lags) - np.amin(lags)),
            1.1,
            np.amin(semivariance),
        ]
        bnds = ([0.0, 0.001, 0.0], [np.inf, 1.999, np.amax(semivariance)])
    else:
        x0 = [
            np.amax(semivariance) - np.amin(semivariance),
            0.25 * np.amax(lags),
            np.amin(semivariance),
        ]
        bnds = (
            [0.0, 0.0, 0.0],
            [10.0 * np.amax(semivariance), np.amax(lags), np.amax(semivariance)],
        )

    def residuals(params, x, y):
        return fn(params, x) - y

    res = least_squares(residuals, x0, bounds=bnds, loss="soft_l1", args=(lags, semivariance))
    return res.x


# --------------------------------------------------------------------------- #
# ok.py
# --------------------------------------------------------------------------- #
class OrdinaryKriging:
    """Restated ``pykrige.ok.OrdinaryKriging`` (subset used by climatrix)."""

    eps = EPS

    def __init__(
        self,
        x,
        y,
        z,
        variogram_model="linear",
        nlags=6,
        anisotropy_scaling=1.0,
        anisotropy_angle=0.0,
        coordinates_type="euclidean",
        exact_values=True,
        pseudo_inv=False,
        pseudo_inv_type="pinv",
        variogram_parameters=None,
    ):
        self.pseudo_inv = bool(pseudo_inv)
        self.pseudo_inv_type = str(pseudo_inv_type)
        if self.pseudo_inv_type not in ("pinv", "pinvh"):
            raise ValueError("pseudo inv type not valid: " + str(pseudo_inv_type))
        self.exact_values = exact_values
        self.coordinates_type = coordinates_type
        if variogram_model not in VARIOGRAM_FUNCTIONS:
            raise ValueError("Specified variogram model '%s' is not supported." % variogram_model)
        self.variogram_model = variogram_model
        self.variogram_function = VARIOGRAM_FUNCTIONS[variogram_model]

        self.X_ORIG = np.atleast_1d(np.squeeze(np.array(x, copy=True, dtype=np.float64)))
        self.Y_ORIG = np.atleast_1d(np.squeeze(np.array(y, copy=True, dtype=np.float64)))
        self.Z = np.atleast_1d(np.squeeze(np.array(z, copy=True, dtype=np.float64)))

        if coordinates_type == "euclidean":
            self.XCENTER = (np.amax(self.X_ORIG) + np.amin(self.X_ORIG)) / 2.0
            self.YCENTER = (np.amax(self.Y_ORIG) + np.amin(self.Y_ORIG)) / 2.0
            self.anisotropy_scaling = anisotropy_scaling
            self.anisotropy_angle = anisotropy_angle
            adj = adjust_for_anisotropy(
                np.vstack((self.X_ORIG, self.Y_ORIG)).T,
                [self.XCENTER, self.YCENTER],
                self.anisotropy_scaling,
                self.anisotropy_angle,
            )
            self.X_ADJUSTED, self.Y_ADJUSTED = adj[:, 0], adj[:, 1]
        elif coordinates_type == "geographic":
            if anisotropy_scaling != 1.0:
                warnings.warn(
                    "Anisotropy is not compatible with geographic coordinates. "
                    "Ignoring user set anisotropy.",
                    UserWarning,
                )
            self.XCENTER = 0.0
            self.YCENTER = 0.0
            self.anisotropy_scaling = 1.0
            self.anisotropy_angle = 0.0
            self.X_ADJUSTED = self.X_ORIG
            self.Y_ADJUSTED = self.Y_ORIG
        else:
            raise ValueError("Only 'euclidean' and 'geographic' are valid values for coordinates-keyword.")

        X = np.vstack((self.X_ADJUSTED, self.Y_ADJUSTED)).T
        self.lags, self.semivariance, self.variogram_raw = empirical_variogram(
            X, self.Z, nlags, coordinates_type
        )
        if variogram_parameters is not None:
            self.variogram_model_parameters = np.asarray(variogram_parameters, dtype=np.float64)
        else:
            self.variogram_model_parameters = fit_variogram(self.lags, self.semivariance, variogram_model)

    # -- helpers ----------------------------------------------------------- #
    def _gamma(self, d):
        return self.variogram_function(self.variogram_model_parameters, d)

    def _data_distances(self, sel=None):
        xa = self.X_ADJUSTED if sel is None else self.X_ADJUSTED[sel]
        ya = self.Y_ADJUSTED if sel is None else self.Y_ADJUSTED[sel]
        if self.coordinates_type == "euclidean":
            xy = np.concatenate((xa[:, None], ya[:, None]), axis=1)
            return cdist(xy, xy, "euclidean")
        return great_circle_distance(xa[:, None], ya[:, None], xa, ya)

    def _get_kriging_matrix(self, n, sel=None):
        d = self._data_distances(sel)
        a = np.zeros((n + 1, n + 1))
        a[:n, :n] = -self._gamma(d)
        np.fill_diagonal(a, 0.0)
        a[n, :] = 1.0
        a[:, n] = 1.0
        a[n, n] = 0.0
        return a

    def _rhs(self, bd):
        b = -self._gamma(bd)
        if self.exact_values:
            b = np.where(np.absolute(bd) <= self.eps, 0.0, b)
        return b

    # -- execute ----------------------------------------------------------- #
    def execute(self, style, xpoints, ypoints, backend="vectorized", n_closest_points=None, chunk=4096):
        if style not in ("grid", "points"):
            raise ValueError("style argument must be 'grid' or 'points' (masked not restated).")
        xpts = np.atleast_1d(np.squeeze(np.array(xpoints, copy=True, dtype=np.float64)))
        ypts = np.atleast_1d(np.squeeze(np.array(ypoints, copy=True, dtype=np.float64)))
        n = self.X_ADJUSTED.shape[0]
        nx, ny = xpts.size, ypts.size
        if style == "grid":
            npt = ny * nx
            grid_x, grid_y = np.meshgrid(xpts, ypts)
            xpts = grid_x.flatten()
            ypts = grid_y.flatten()
        else:
            if xpts.size != ypts.size:
                raise ValueError("xpoints and ypoints must have same dimensions when treated as listing discrete points.")
            npt = nx

        if self.coordinates_type == "euclidean":
            adj = adjust_for_anisotropy(
                np.vstack((xpts, ypts)).T,
                [self.XCENTER, self.YCENTER],
                self.anisotropy_scaling,
                self.anisotropy_angle,
            )
            xpts, ypts = adj[:, 0], adj[:, 1]
            xy_data = np.concatenate((self.X_ADJUSTED[:, None], self.Y_ADJUSTED[:, None]), axis=1)
            xy_points = np.concatenate((xpts[:, None], ypts[:, None]), axis=1)
        else:
            lon_d = self.X_ADJUSTED[:, None] * np.pi / 180.0
            lat_d = self.Y_ADJUSTED[:, None] * np.pi / 180.0
            xy_data = np.concatenate(
                (np.cos(lon_d) * np.cos(lat_d), np.sin(lon_d) * np.cos(lat_d), np.sin(lat_d)), axis=1
            )
            lon_p = xpts[:, None] * np.pi / 180.0
            lat_p = ypts[:, None] * np.pi / 180.0
            xy_points = np.concatenate(
                (np.cos(lon_p) * np.cos(lat_p), np.sin(lon_p) * np.cos(lat_p), np.sin(lat_p)), axis=1
            )

        zvalues = np.zeros(npt)
        sigmasq = np.zeros(npt)

        if n_closest_points is not None:
            if backend != "loop":
                raise ValueError("Specified backend {} for a moving window is not supported.".format(backend))
            tree = cKDTree(xy_data)
            bd_all, bd_idx = tree.query(xy_points, k=n_closest_points, eps=0.0)
            if n_closest_points == 1:
                bd_all = bd_all[:, None]
                bd_idx = bd_idx[:, None]
            if self.coordinates_type == "geographic":
                bd_all = euclid3_to_great_circle(bd_all)
            self.last_neighbours = (bd_all, bd_idx)
            k = bd_idx.shape[1]
            for i in range(npt):
                sel = bd_idx[i]
                a = self._get_kriging_matrix(k, sel)
                b = np.zeros((k + 1, 1))
                b[:k, 0] = self._rhs(bd_all[i])
                b[k, 0] = 1.0
                x = scipy.linalg.solve(a, b)
                zvalues[i] = x[:k, 0].dot(self.Z[sel])
                sigmasq[i] = -x[:, 0].dot(b[:, 0])
        else:
            a = self._get_kriging_matrix(n)
            if backend == "vectorized":
                if self.pseudo_inv:
                    a_inv = scipy.linalg.pinv(a) if self.pseudo_inv_type == "pinv" else scipy.linalg.pinvh(a)
                else:
                    a_inv = scipy.linalg.inv(a)
            elif backend != "loop":
                raise ValueError("Specified backend {} is not supported for 2D ordinary kriging.".format(backend))
            for s in range(0, npt, chunk):
                e = min(npt, s + chunk)
                if self.coordinates_type == "euclidean":
                    bd = cdist(xy_points[s:e], xy_data, "euclidean")
                else:
                    bd = great_circle_distance(
                        xpts[s:e, None], ypts[s:e, None], self.X_ADJUSTED, self.Y_ADJUSTED
                    )
                b = np.ones((e - s, n + 1))
                b[:, :n] = self._rhs(bd)
                if backend == "vectorized":
                    x = np.dot(a_inv, b.T).T
                else:
                    # _exec_loop inverts once and does one mat-vec per target
                    if s == 0:
                        if self.pseudo_inv:
                            a_inv = scipy.linalg.pinv(a) if self.pseudo_inv_type == "pinv" else scipy.linalg.pinvh(a)
                        else:
                            a_inv = scipy.linalg.inv(a)
                    x = np.dot(a_inv, b.T).T
                zvalues[s:e] = np.sum(x[:, :n] * self.Z, axis=1)
                sigmasq[s:e] = np.sum(x * -b, axis=1)

        if style == "grid":
            zvalues = zvalues.reshape((ny, nx))
            sigmasq = sigmasq.reshape((ny, nx))
        return zvalues, sigmasq
