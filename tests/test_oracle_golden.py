"""CPU: the oracle restatement reproduces the reference-generated golden vectors.

The fixtures were produced by the UNMODIFIED reference run in the build
container (``oracle/make_golden.py``).  IDW: bit-exact (same scipy kd-tree,
same NumPy expressions).  OK: the climatrix wrapper is pinned; the pykrige
arithmetic inside is the restatement itself (parity unpinned).
"""

import numpy as np
import pytest

from conftest import golden_names, load_golden
from oracle.idw_oracle import dense_points, idw_oracle, knn_bruteforce, knn_ckdtree, sparse_points
from oracle.ok_oracle import ok_oracle


def _points(lat, lon, sparse):
    return sparse_points(lat, lon) if sparse else dense_points(lat, lon)


@pytest.mark.parametrize("name", golden_names("idw_"))
def test_idw_oracle_matches_reference(name):
    g = load_golden(name)
    m = g["meta"]
    src = _points(g["sample_lat"], g["sample_lon"], m["source_sparse"])
    tgt = _points(g["target_lat"], g["target_lon"], m["target_sparse"])
    with np.errstate(invalid="ignore"):
        out = idw_oracle(src, g["values"], tgt, **{"k": 5, "power": 2.0, "k_min": 2, **m["kwargs"]})
    expected = g["expected"]
    assert out.reshape(expected.shape).tobytes() == expected.tobytes() or np.array_equal(
        out.reshape(expected.shape), expected, equal_nan=True
    )


@pytest.mark.parametrize("name", golden_names("ok_"))
def test_ok_oracle_matches_reference_wrapper(name):
    g = load_golden(name)
    m = g["meta"]
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out, model = ok_oracle(
            g["sample_lat"], g["sample_lon"], g["values"], g["target_lat"], g["target_lon"], m["target_sparse"],
            n_closest_points=m["n_closest_points"], return_model=True, **m["kwargs"],
        )
    np.testing.assert_array_equal(model.variogram_model_parameters, g["variogram_parameters"])
    np.testing.assert_array_equal(out, g["expected"])


def test_bruteforce_knn_equals_ckdtree_on_random_data():
    rng = np.random.default_rng(7)
    pts = np.stack((rng.uniform(-90, 90, 3000), rng.uniform(-180, 180, 3000)), axis=1)
    qry = np.stack((rng.uniform(-90, 90, 2000), rng.uniform(-180, 180, 2000)), axis=1)
    for k in (1, 8, 30):
        d0, i0 = knn_ckdtree(pts, qry, k)
        d1, i1 = knn_bruteforce(pts, qry, k)
        np.testing.assert_array_equal(i0, i1)
        np.testing.assert_array_equal(d0, d1)  # bit-equal: sqrt(dx*dx + dy*dy) without FMA


def test_kriging_oracle_self_checks():
    """Properties that hold for any correct ordinary kriging."""
    rng = np.random.default_rng(11)
    lat = rng.uniform(-60, 60, 120)
    lon = rng.uniform(-120, 120, 120)
    v = 3.0 + np.sin(np.deg2rad(lat)) + rng.normal(scale=0.1, size=120)
    kw = dict(variogram_model="spherical", anisotropy_scaling=1.0)
    # exact interpolator at the sample locations (same bbox -> same normalised frame)
    out = ok_oracle(lat, lon, v, lat, lon, True, n_closest_points=16, **kw)
    np.testing.assert_allclose(out, v, rtol=0, atol=1e-9)
    # moving window with k = N equals global kriging
    t_lat, t_lon = np.linspace(-60, 60, 7), np.linspace(-120, 120, 9)
    a = ok_oracle(lat, lon, v, t_lat, t_lon, False, n_closest_points=120, **kw)
    b = ok_oracle(lat, lon, v, t_lat, t_lon, False, **kw)
    np.testing.assert_allclose(a, b, rtol=0, atol=1e-9)
    # constant + tiny noise stays near the constant (weights sum to one)
    c = ok_oracle(lat, lon, 5.0 + 1e-9 * rng.normal(size=120), t_lat, t_lon, False, n_closest_points=8, **kw)
    np.testing.assert_allclose(c, 5.0, atol=1e-6)


def _hand_cases():
    import json
    import os

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ok_hand_derived.json")) as f:
        return json.load(f)["cases"]


def test_restated_pykrige_matches_the_hand_derived_fixture():
    """tests/golden/ok_hand_derived.json was solved in exact rational arithmetic (oracle/make_hand_fixture.py) and does
    not pass through pykrige_restated.py: the restatement - global and moving-window execute() - must reproduce it."""
    from oracle import pykrige_restated as pk
    from oracle.ok_oracle import moving_window_with_parameters

    for case in _hand_cases():
        x, y, z = (np.array(case[key]) for key in ("x", "y", "values"))
        tx = np.array([t["target"][0] for t in case["targets"]])
        ty = np.array([t["target"][1] for t in case["targets"]])
        want_z = np.array([t["z"] for t in case["targets"]])
        want_s = np.array([t["sigmasq"] for t in case["targets"]])
        okr = pk.OrdinaryKriging(x, y, z, variogram_model="linear", variogram_parameters=case["variogram_parameters"])
        for backend in ("vectorized", "loop"):
            got_z, got_s = okr.execute("points", tx, ty, backend=backend)
            np.testing.assert_allclose(got_z, want_z, rtol=0, atol=1e-13)
            np.testing.assert_allclose(got_s, want_s, rtol=0, atol=1e-13)
        got_z, got_s = moving_window_with_parameters(x, y, z, tx, ty, "points", k=len(x), variogram_model="linear",
                                                     parameters=case["variogram_parameters"])
        np.testing.assert_allclose(got_z, want_z, rtol=0, atol=1e-13)
        np.testing.assert_allclose(got_s, want_s, rtol=0, atol=1e-13)
