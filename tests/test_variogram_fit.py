"""The native variogram-model fit (csrc/variogram_fit.cuh) against SciPy.

pykrige fits the variogram model with ``scipy.optimize.least_squares(..., loss="soft_l1")`` at SciPy's default
tolerances (reference call site src/climatrix/reconstruct/kriging.py:215-224).  What the reference obtains is the
point where SciPy's Trust Region Reflective iteration STOPS, which is not the exact minimiser, so the native
code restates that iteration.  Here: the host build of it (``cmx_variogram_fit_host``, no GPU needed) against

* the parameters stored in the 20 kriging fixtures (produced with SciPy by ``oracle/make_golden.py``),
* a fresh ``least_squares`` call on synthetic variograms of every model.

Agreement is to ~1e-7: both run the same iteration, but the stopping tests compare noise-level quantities
(finite-difference gradients against 1e-8), so the two can stop one iteration apart.
"""

import json
import warnings

import numpy as np
import pytest

from climatrix_b200 import kernels as K
from climatrix_b200.kriging import _variogram_value, fit_variogram
from conftest import golden_names, load_golden
from oracle import pykrige_restated as pk
from oracle.ok_oracle import normalize_axis, standardize_values

MODELS = ["linear", "power", "gaussian", "spherical", "exponential"]


@pytest.mark.parametrize("name", golden_names("ok_"))
def test_native_fit_reproduces_the_fixture_parameters(name):
    g = load_golden(name)
    kw = g["meta"]["kwargs"]
    model = kw.get("variogram_model") or "linear"
    *_, lat = normalize_axis(g["sample_lat"])
    *_, lon = normalize_axis(g["sample_lon"])
    _, _, z = standardize_values(np.asarray(g["values"]).astype(float).squeeze())
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        okr = pk.OrdinaryKriging(lon, lat, z, variogram_model=model, nlags=kw.get("nlags") or 6,
                                 anisotropy_scaling=kw.get("anisotropy_scaling") or 1e-6,
                                 coordinates_type=kw.get("coordinates_type") or "euclidean")
    got = K.fit_variogram_native(okr.lags, okr.semivariance, model)
    np.testing.assert_allclose(got, g["variogram_parameters"], rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("model", MODELS)
def test_native_fit_follows_scipy_trf_on_synthetic_variograms(model):
    rng = np.random.default_rng(MODELS.index(model))
    true = {"linear": [0.7, 0.1], "power": [0.5, 1.3, 0.05], "gaussian": [1.0, 1.2, 0.1], "spherical": [1.0, 1.5, 0.05],
            "exponential": [1.0, 1.0, 0.02]}[model]
    close = 0
    trials = 60
    for trial in range(trials):
        nl = int(rng.choice([6, 10, 20]))
        # bin means of an empirical variogram: roughly equally spaced, the first one well below lag_max / 4
        lags = (np.arange(nl) + rng.uniform(0.3, 0.7, nl)) * (2.4 / nl)
        sv = np.abs(_variogram_value(model, true, lags) * (1.0 + rng.choice([0.01, 0.05, 0.1]) * rng.normal(size=nl)))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = fit_variogram(lags, sv, model)
        got = K.fit_variogram_native(lags, sv, model)
        assert got.shape == ref.shape and np.all(np.isfinite(got))
        # never a worse fit than SciPy's stopping point (beyond the stopping tolerance itself)
        cost = lambda p: np.sum(2.0 * (np.sqrt(1.0 + (_variogram_value(model, p, lags) - sv) ** 2) - 1.0))  # noqa: E731
        assert cost(got) <= cost(ref) * (1.0 + 1e-7) + 1e-12
        close += bool(np.all(np.abs(got - ref) <= 1e-6 * np.maximum(1.0, np.abs(ref))))
    assert close >= 0.9 * trials, f"only {close} of {trials} fits within 1e-6 of SciPy's"


def test_native_fit_argument_checks():
    with pytest.raises(ValueError):
        K.fit_variogram_native(np.arange(70.0), np.arange(70.0), "linear")
    with pytest.raises(ValueError):
        K.fit_variogram_native([1.0, 2.0], [1.0], "linear")
    with pytest.raises(ValueError):  # constant semivariances: lb >= ub, SciPy raises too
        K.fit_variogram_native([1.0, 2.0, 3.0], [0.0, 0.0, 0.0], "spherical")
