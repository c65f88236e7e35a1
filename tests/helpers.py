"""Shared builders for the tests (datasets shaped like the reference's own fixtures)."""

import numpy as np

import climatrix_b200 as cm

REF_LAT = np.array([-90.0, -45.0, 0.0, 45.0, 90.0])
REF_LON = np.array([-180.0, -90.0, 0.0, 90.0, 180.0])
TIMES3 = np.array(["2000-01-01", "2000-01-02", "2000-01-03"], dtype="datetime64[ns]")


def sparse_dataset(lat=REF_LAT, lon=REF_LON, values=None, times=TIMES3[:1], name="field", seed=0):
    """(point, time) like tests/unit/reconstruct/test_base_interface.py:80-101."""
    rng = np.random.default_rng(seed)
    n, t = len(lat), len(times)
    if values is None:
        values = rng.random((n, t))
    values = np.asarray(values, dtype=float).reshape(n, t)
    return cm.BaseClimatrixDataset(
        cm.FieldArray(
            values,
            ("point", "time"),
            {"point": np.arange(n), "time": times, "latitude": (("point",), np.asarray(lat, float)),
             "longitude": (("point",), np.asarray(lon, float))},
            name=name,
        )
    )


def dense_dataset(lat=REF_LAT, lon=REF_LON, values=None, times=TIMES3[:1], name="field", seed=0):
    """(time, latitude, longitude) like test_base_interface.py:103-122."""
    rng = np.random.default_rng(seed)
    t = len(times)
    if values is None:
        values = rng.random((t, len(lat), len(lon)))
    values = np.asarray(values, dtype=float).reshape(t, len(lat), len(lon))
    return cm.BaseClimatrixDataset(
        cm.FieldArray(values, ("time", "latitude", "longitude"),
                      {"time": times, "latitude": np.asarray(lat, float), "longitude": np.asarray(lon, float)}, name=name)
    )


def static_sparse(lat, lon, values, name="field"):
    """(point,) only."""
    return cm.BaseClimatrixDataset(
        cm.FieldArray(np.asarray(values, float), ("point",),
                      {"point": np.arange(len(lat)), "latitude": (("point",), np.asarray(lat, float)),
                       "longitude": (("point",), np.asarray(lon, float))}, name=name)
    )


def synthetic_field(lat, lon, rng, phase=0.0):
    return 15.0 + 10.0 * np.cos(np.deg2rad(lat)) + 5.0 * np.sin(2.0 * np.deg2rad(lon) + phase) + rng.normal(size=np.shape(lat))


def rel_err(out, ref):
    """max |out-ref| / max(|ref|, eps) over finite entries; NaN patterns must match exactly."""
    out, ref = np.asarray(out, float), np.asarray(ref, float)
    assert out.shape == ref.shape, (out.shape, ref.shape)
    assert np.array_equal(np.isnan(out), np.isnan(ref)), "NaN positions differ"
    ok = ~np.isnan(ref)
    if not ok.any():
        return 0.0
    return float(np.max(np.abs(out[ok] - ref[ok]) / np.maximum(np.abs(ref[ok]), 1e-12)))
