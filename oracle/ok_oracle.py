"""Ordinary-kriging oracle: the climatrix wrapper around (restated) pykrige.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Follows ``/root/reference/src/climatrix/reconstruct/kriging.py:147-250`` at
array level; the pykrige arithmetic comes from ``oracle/pykrige_restated.py``
(**parity unpinned** for that part -- pykrige is not vendored by the reference
nor installable here).
"""

from __future__ import annotations

import numpy as np

from .pykrige_restated import OrdinaryKriging

MAX_VECTORIZED_SIZE = 500_000  # kriging.py:79


def normalize_axis(v: np.ndarray):
    """kriging.py:147-161: min-max scale to [-1, 1] with the array's own extrema."""
    v = np.asarray(v, dtype=np.float64)
    vmax = np.nanmax(v)
    vmin = np.nanmin(v)
    scaled = (v - vmin) / (vmax - vmin)
    scaled -= 0.5
    scaled *= 2
    return vmin, vmax, scaled


def standardize_values(values: np.ndarray):
    """kriging.py:163-173 (the std == 0 branch of the reference is broken: it
    returns one array where the caller unpacks three -> ValueError/TypeError)."""
    mean = np.nanmean(values)
    std = np.nanstd(values)
    if std == 0:
        raise ValueError("reference cannot unpack _standarize_values() when std == 0 (kriging.py:166-171,212)")
    return mean, std, (values - mean) / std


def ok_oracle(
    sample_lat,
    sample_lon,
    values,
    target_lat,
    target_lon,
    target_is_sparse: bool,
    *,
    backend=None,
    nlags=6,
    anisotropy_scaling=1e-6,
    coordinates_type="euclidean",
    variogram_model="linear",
    pseudo_inv=False,
    n_closest_points=None,
    return_model: bool = False,
):
    """Reconstruct on the target; returns float64 (n_lat, n_lon) or (M,).

    ``n_closest_points`` is the extension BASELINE.json's north_star adds
    (pykrige's moving window, ``backend="loop"``); ``None`` gives the reference's
    shipped global kriging.
    """
    if backend is None:  # kriging.py:193-204
        backend = "vectorized" if (len(target_lat) * len(target_lon)) < MAX_VECTORIZED_SIZE else "loop"
    if n_closest_points is not None:
        backend = "loop"  # the only pykrige backend with a moving window
    *_, lat = normalize_axis(sample_lat)  # kriging.py:207-210
    *_, lon = normalize_axis(sample_lon)
    v_mean, v_std, z = standardize_values(np.asarray(values).astype(float).squeeze())  # :212-214
    kriging = OrdinaryKriging(  # :215-224
        x=lon,
        y=lat,
        z=z,
        nlags=nlags,
        anisotropy_scaling=anisotropy_scaling,
        coordinates_type=coordinates_type,
        variogram_model=variogram_model,
        pseudo_inv=pseudo_inv,
    )
    recon_type = "points" if target_is_sparse else "grid"  # :226
    *_, t_lat = normalize_axis(target_lat)  # :230-235 (target's OWN extrema)
    *_, t_lon = normalize_axis(target_lon)
    zvalues, _ = kriging.execute(recon_type, t_lon, t_lat, backend=backend, n_closest_points=n_closest_points)
    out = np.ma.getdata(zvalues) * v_std + v_mean  # :242-245
    if return_model:
        return out, kriging
    return out


def moving_window_with_parameters(x, y, z, xpoints, ypoints, style: str, *, k: int, variogram_model: str, parameters):
    """pykrige ``execute(style, ..., backend="loop", n_closest_points=k)`` for GIVEN variogram parameters, in the
    kriging frame (x, y already normalised; isotropic).  Skips the O(N^2) empirical variogram of the constructor -
    which does not enter ``execute`` - so that spot checks at the BASELINE sizes (N = 50 000) stay cheap.
    Returns ``(z, sigmasq)``."""
    from . import pykrige_restated as pk

    okr = OrdinaryKriging.__new__(OrdinaryKriging)
    okr.X_ADJUSTED = np.asarray(x, dtype=np.float64)
    okr.Y_ADJUSTED = np.asarray(y, dtype=np.float64)
    okr.Z = np.asarray(z, dtype=np.float64)
    okr.XCENTER = okr.YCENTER = 0.0
    okr.anisotropy_scaling, okr.anisotropy_angle = 1.0, 0.0
    okr.coordinates_type = "euclidean"
    okr.variogram_model = variogram_model
    okr.variogram_function = pk.VARIOGRAM_FUNCTIONS[variogram_model]
    okr.variogram_model_parameters = np.asarray(parameters, dtype=np.float64)
    okr.exact_values, okr.pseudo_inv, okr.pseudo_inv_type = True, False, "pinv"
    return okr.execute(style, xpoints, ypoints, backend="loop", n_closest_points=k)
