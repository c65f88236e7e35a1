"""Generate ``tests/golden/*.npz`` by running the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs ``/root/reference``):

    python -m oracle.make_golden

Each fixture holds the inputs, the keyword arguments and the array the
reference reconstructor produced (``IDWReconstructor.reconstruct`` /
``OrdinaryKrigingReconstructor.reconstruct`` executed from
``/root/reference/src`` through ``oracle/reference_loader.py``).

* ``idw_*``: fully reference-generated (real climatrix code + real scipy
  cKDTree) -> these PIN the IDW oracle and the CUDA path.
* ``ok_*``: the real climatrix wrapper (normalisation, standardisation,
  style/backend choice) around ``oracle/pykrige_restated.py`` -- pykrige is
  not installable here, so the kriging arithmetic is **parity unpinned**.
  ``ok_mw_*`` additionally inject ``n_closest_points`` into pykrige's
  ``execute`` (the moving-window mode BASELINE.json's north_star specifies and
  climatrix does not expose).
"""

from __future__ import annotations

import functools
import json
import os
from unittest import mock

import numpy as np

from oracle.reference_loader import load_reference, make_dataset, make_domain, run_reference

OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def synthetic_field(lat, lon, rng, phase=0.0):
    """smooth field + N(0,1) noise."""
    return 15.0 + 10.0 * np.cos(np.deg2rad(lat)) + 5.0 * np.sin(2.0 * np.deg2rad(lon) + phase) + rng.normal(size=lat.shape)


def _save(name, meta, **arrays):
    os.makedirs(OUT_DIR, exist_ok=True)
    path = os.path.join(OUT_DIR, name + ".npz")
    np.savez_compressed(path, meta=np.array(json.dumps(meta)), **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def idw_cases():
    cases = []
    rng = np.random.default_rng(1234)

    def rand_samples(n, r=rng):
        lat = r.uniform(-90, 90, n)
        lon = r.uniform(-180, 180, n)
        return lat, lon, synthetic_field(lat, lon, r)

    # BASELINE config 1 at full size: 1000 samples -> 1 deg grid, k=30, power=2
    lat, lon, v = rand_samples(1000)
    cases.append(("idw_c1_full", dict(k=30, power=2.0, k_min=2), lat, lon, v, np.arange(-89.5, 90, 1.0), np.arange(-179.5, 180, 1.0), True, False))
    # defaults (k=5, power=2, k_min=2), coarse grid
    lat, lon, v = rand_samples(300)
    cases.append(("idw_defaults", dict(), lat, lon, v, np.linspace(-85, 85, 18), np.linspace(-175, 175, 36), True, False))
    # powers across the allowed range (idw.py:59)
    for p in (1e-10, 0.5, 3.0, 5.0):
        lat, lon, v = rand_samples(400)
        cases.append((f"idw_power_{p:g}", dict(k=12, power=p, k_min=1), lat, lon, v, np.linspace(-80, 80, 17), np.linspace(-170, 170, 35), True, False))
    # k = 1 and k = N
    lat, lon, v = rand_samples(64)
    cases.append(("idw_k1", dict(k=1, k_min=1), lat, lon, v, np.linspace(-80, 80, 9), np.linspace(-170, 170, 18), True, False))
    cases.append(("idw_k_eq_n", dict(k=64, k_min=2), lat, lon, v, np.linspace(-80, 80, 9), np.linspace(-170, 170, 18), True, False))
    # NaN / inf sample values with several k_min
    lat, lon, v = rand_samples(500)
    v = v.copy()
    bad = rng.choice(500, 120, replace=False)
    v[bad[:80]] = np.nan
    v[bad[80:100]] = np.inf
    v[bad[100:]] = -np.inf
    for k_min in (1, 4, 8):
        cases.append((f"idw_nonfinite_kmin{k_min}", dict(k=8, power=2.0, k_min=k_min), lat, lon, v, np.linspace(-85, 85, 35), np.linspace(-175, 175, 71), True, False))
    # exact hits: some targets coincide with samples (distance clamp 1e-10, idw.py:163)
    lat, lon, v = rand_samples(200)
    t_lat = np.concatenate([lat[:50], rng.uniform(-90, 90, 50)])
    t_lon = np.concatenate([lon[:50], rng.uniform(-180, 180, 50)])
    cases.append(("idw_exact_hits_sparse_target", dict(k=6, power=2.0, k_min=2), lat, lon, v, t_lat, t_lon, True, True))
    # duplicate sample coordinates (distinct values)
    lat, lon, v = rand_samples(150)
    lat[100:] = lat[:50]
    lon[100:] = lon[:50]
    cases.append(("idw_duplicate_samples", dict(k=7, power=2.0, k_min=2), lat, lon, v, np.linspace(-85.3, 85.3, 21), np.linspace(-175.3, 175.3, 43), True, False))
    # dense source (1, n_lat, n_lon) -> sparse target; off-lattice targets avoid distance ties
    g_lat = np.arange(-80.0, 81.0, 10.0)
    g_lon = np.arange(-170.0, 171.0, 10.0)
    gl, go = np.meshgrid(g_lat, g_lon, indexing="ij")
    v = synthetic_field(gl, go, rng)[None]
    cases.append(("idw_dense_source_sparse_target", dict(k=9, power=2.0, k_min=2), g_lat, g_lon, v, rng.uniform(-80, 80, 300), rng.uniform(-170, 170, 300), False, True))
    # the reference's own 5-point test fixture (tests/unit/reconstruct/test_base_interface.py:80-101), default k=5=N
    lat = np.array([-90.0, -45.0, 0.0, 45.0, 90.0])
    lon = np.array([-180.0, -90.0, 0.0, 90.0, 180.0])
    v = rng.random((5, 1))
    cases.append(("idw_reference_fixture_sparse_to_dense", dict(), lat, lon, v, lat, lon, True, False))
    # clustered samples + regional grid: small distances, stresses the fp32 filter margin
    c_lat = 50 + rng.normal(scale=0.05, size=2000)
    c_lon = 10 + rng.normal(scale=0.05, size=2000)
    v = synthetic_field(c_lat, c_lon, rng)
    cases.append(("idw_clustered_regional", dict(k=32, power=2.0, k_min=2), c_lat, c_lon, v, np.linspace(49.8, 50.2, 41), np.linspace(9.8, 10.2, 41), True, False))
    return cases


def ok_cases():
    cases = []
    rng = np.random.default_rng(4321)

    def rand_samples(n, lat_rng=(-90, 90), lon_rng=(-180, 180)):
        lat = rng.uniform(*lat_rng, n)
        lon = rng.uniform(*lon_rng, n)
        return lat, lon, synthetic_field(lat, lon, rng)

    grid = (np.linspace(-88, 88, 23), np.linspace(-176, 176, 45))
    for model in ("linear", "power", "gaussian", "spherical", "exponential"):
        lat, lon, v = rand_samples(300)
        cases.append((f"ok_global_{model}", dict(variogram_model=model, anisotropy_scaling=1.0), None, lat, lon, v, *grid, False))
        cases.append((f"ok_mw_{model}_k16", dict(variogram_model=model, anisotropy_scaling=1.0), 16, lat, lon, v, *grid, False))
    # climatrix defaults (linear, nlags=6, anisotropy_scaling=1e-6 !)
    lat, lon, v = rand_samples(250)
    cases.append(("ok_global_defaults", dict(), None, lat, lon, v, *grid, False))
    cases.append(("ok_mw_defaults_k12", dict(), 12, lat, lon, v, *grid, False))
    # anisotropy 5.0, nlags 12, loop backend
    lat, lon, v = rand_samples(250)
    cases.append(("ok_global_aniso5_loop", dict(variogram_model="spherical", anisotropy_scaling=5.0, nlags=12, backend="loop"), None, lat, lon, v, *grid, False))
    cases.append(("ok_mw_aniso5_k32", dict(variogram_model="spherical", anisotropy_scaling=5.0, nlags=12), 32, lat, lon, v, *grid, False))
    # targets whose bounding box differs from the samples' (independent normalisation quirk, kriging.py:230-235)
    lat, lon, v = rand_samples(300, (30, 70), (-10, 40))
    cases.append(("ok_mw_bbox_mismatch_k20", dict(variogram_model="exponential", anisotropy_scaling=1.0), 20, lat, lon, v, np.linspace(35, 60, 26), np.linspace(0, 30, 31), False))
    # sparse target ("points" style) containing exact sample hits is impossible to arrange through
    # independent normalisation unless bboxes coincide: reuse the sample bbox corners
    lat, lon, v = rand_samples(200)
    t_lat = np.concatenate([lat[:60], [lat.min(), lat.max()]])
    t_lon = np.concatenate([lon[:60], [lon.min(), lon.max()]])
    t_lat[-2:] = [lat.min(), lat.max()]
    cases.append(("ok_mw_points_exact_hits_k10", dict(variogram_model="spherical", anisotropy_scaling=1.0), 10, lat, lon, v, t_lat, t_lon, True))
    cases.append(("ok_global_points_pinv", dict(variogram_model="gaussian", anisotropy_scaling=1.0, pseudo_inv=True), None, lat, lon, v, t_lat, t_lon, True))
    # geographic coordinates (great-circle metric on the normalised +-1 deg patch, reproduced as is)
    lat, lon, v = rand_samples(200)
    cases.append(("ok_global_geographic", dict(variogram_model="spherical", coordinates_type="geographic"), None, lat, lon, v, *grid, False))
    cases.append(("ok_mw_geographic_k16", dict(variogram_model="spherical", coordinates_type="geographic"), 16, lat, lon, v, *grid, False))
    # BASELINE config 2 down-scaled 10x: 1000 samples -> 2 deg grid, k=32, spherical
    lat, lon, v = rand_samples(1000)
    cases.append(("ok_mw_c2_downscaled_k32", dict(variogram_model="spherical", anisotropy_scaling=1.0, nlags=6), 32, lat, lon, v, np.arange(-89.0, 90, 2.0), np.arange(-179.0, 180, 2.0), False))
    return cases


def main():
    ref = load_reference()
    for name, kw, lat, lon, v, t_lat, t_lon, src_sparse, tgt_sparse in idw_cases():
        if src_sparse:
            src = make_dataset(ref, make_domain(ref, lat, lon, True), v)
        else:
            src = make_dataset(ref, make_domain(ref, lat, lon, False), v)
        tgt = make_domain(ref, t_lat, t_lon, tgt_sparse)
        out = run_reference(ref, "idw", src, tgt, **kw)
        meta = dict(method="idw", kwargs=kw, source_sparse=src_sparse, target_sparse=tgt_sparse,
                    generator="reference IDWReconstructor.reconstruct (climatrix src/climatrix/reconstruct/idw.py) + scipy cKDTree")
        _save(name, meta, sample_lat=lat, sample_lon=lon, values=v, target_lat=t_lat, target_lon=t_lon, expected=out)

    pk = __import__("pykrige.ok", fromlist=["OrdinaryKriging"])
    for name, kw, ncp, lat, lon, v, t_lat, t_lon, tgt_sparse in ok_cases():
        src = make_dataset(ref, make_domain(ref, lat, lon, True), v.reshape(-1, 1))
        tgt = make_domain(ref, t_lat, t_lon, tgt_sparse)
        fitted = {}
        orig_execute = pk.OrdinaryKriging.execute

        def execute(self, style, xp, yp, backend="vectorized", _ncp=ncp):
            fitted["params"] = np.array(self.variogram_model_parameters)
            fitted["lags"] = np.array(self.lags)
            fitted["semivariance"] = np.array(self.semivariance)
            if _ncp is not None:
                return orig_execute(self, style, xp, yp, backend="loop", n_closest_points=_ncp)
            return orig_execute(self, style, xp, yp, backend=backend)

        with mock.patch.object(pk.OrdinaryKriging, "execute", execute):
            out = run_reference(ref, "ok", src, tgt, **kw)
        meta = dict(method="ok", kwargs=kw, n_closest_points=ncp, source_sparse=True, target_sparse=tgt_sparse,
                    generator="reference OrdinaryKrigingReconstructor.reconstruct (kriging.py) around oracle/pykrige_restated.py; "
                              "pykrige arithmetic parity UNPINNED" + ("" if ref.pykrige_is_real else " (pykrige not installed)"))
        _save(name, meta, sample_lat=lat, sample_lon=lon, values=v, target_lat=t_lat, target_lon=t_lon, expected=out,
              variogram_parameters=fitted["params"], lags=fitted["lags"], semivariance=fitted["semivariance"])


if __name__ == "__main__":
    main()
