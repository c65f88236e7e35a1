"""CPU: the plug-in boundary behaves like the reference's (registry, enum, hparams, errors).

Mirrors tests/unit/reconstruct/test_base_interface.py, test_idw.py, test_type.py and the
registry/hparam parts of tests/unit/optim/test_bayesian.py of the reference, on
FieldArray-backed datasets (xarray is not installed in the image).
"""

import warnings

import numpy as np
import pytest

import climatrix_b200 as cm
from climatrix_b200 import BaseReconstructor, Hyperparameter, IDWReconstructor, OrdinaryKrigingReconstructor, ReconstructionType
from climatrix_b200.exceptions import OperationNotSupportedForDynamicDatasetError, ReconstructorConfigurationFailed
from helpers import TIMES3, dense_dataset, sparse_dataset

DATASETS = {
    "sparse_static": lambda: sparse_dataset(),
    "dense_static": lambda: dense_dataset(),
    "sparse_dynamic": lambda: sparse_dataset(times=TIMES3),
    "dense_dynamic": lambda: dense_dataset(times=TIMES3),
}
all_datasets = pytest.mark.parametrize("dataset", list(DATASETS), indirect=True)


@pytest.fixture
def dataset(request):
    return DATASETS[request.param]()


class TestRegistry:
    def test_classes_by_name(self):
        assert BaseReconstructor.get("idw") is IDWReconstructor
        assert BaseReconstructor.get("ok") is OrdinaryKrigingReconstructor
        assert BaseReconstructor.get(" IdW ") is IDWReconstructor

    def test_unknown_method(self):
        with pytest.raises(ValueError, match="Unknown method"):
            BaseReconstructor.get("unknown_method")

    def test_available_methods(self):
        assert {"idw", "ok"} <= set(BaseReconstructor.get_available_methods())

    def test_later_definition_overwrites_and_enum_follows(self):
        original = BaseReconstructor._registry["idw"]
        try:
            class Other(BaseReconstructor):
                NAME = "idw"

                def reconstruct(self):
                    return None

                num_params = 0

            assert BaseReconstructor.get("idw") is Other
            assert ReconstructionType.get("idw").value is Other
        finally:
            BaseReconstructor._registry["idw"] = original
            from climatrix_b200 import plugin

            plugin._invalidate_enum()
        assert ReconstructionType.get("idw").value is IDWReconstructor


class TestReconstructionType:
    def test_values(self):
        assert ReconstructionType.IDW.value is IDWReconstructor
        assert ReconstructionType.OK.value is OrdinaryKrigingReconstructor

    def test_get_variants(self):
        rt = ReconstructionType.IDW
        assert ReconstructionType.get(rt) is rt
        for name in ("IDW", "idw", "IdW"):
            assert ReconstructionType.get(name) == ReconstructionType.IDW

    def test_get_invalid(self):
        with pytest.raises(ValueError, match="Unknown reconstruction type: invalid_type"):
            ReconstructionType.get("invalid_type")
        for bad in (123, None):
            with pytest.raises(TypeError, match="Invalid reconstruction type"):
                ReconstructionType.get(bad)

    def test_missing_value(self):
        with pytest.raises(ValueError, match="is not a valid ReconstructionType"):
            ReconstructionType("invalid_method")

    def test_list_and_iteration(self):
        assert "idw" in ReconstructionType.list()
        for rt in ReconstructionType:
            assert ReconstructionType.get(rt.name) == rt


class TestHyperparameters:
    def test_idw_specs(self):
        h = IDWReconstructor.get_hparams()
        assert set(h) == {"power", "k", "k_min"}
        assert issubclass(h["power"]["type"], float) and "bounds" not in h["power"]
        assert h["k"] == {"type": int, "bounds": (1, None), "default": 5}
        assert h["k_min"]["default"] == 2 and h["power"]["default"] == 2.0

    def test_ok_specs(self):
        h = OrdinaryKrigingReconstructor.get_hparams()
        assert h["nlags"] == {"type": int, "bounds": (4, 20), "default": 6}
        assert h["anisotropy_scaling"]["default"] == 1e-6
        assert h["variogram_model"]["values"] == ["linear", "power", "gaussian", "spherical", "exponential"]
        assert h["coordinates_type"]["default"] == "euclidean"

    def test_cast_and_warn_only(self):
        ds = sparse_dataset()
        r = IDWReconstructor(ds, ds.domain, k="3", power=1, k_min=1)
        assert r.k == 3 and isinstance(r.k, int) and r.power == 1.0 and isinstance(r.power, float)
        with pytest.raises(TypeError, match="Cannot convert 'abc' to int for parameter 'k'"):
            IDWReconstructor(ds, ds.domain, k="abc")
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            r = OrdinaryKrigingReconstructor(ds, ds.domain, nlags=50, variogram_model="cubic")
        assert r.nlags == 50 and r.variogram_model == "cubic"  # out of range: warned, not rejected
        assert sum("above the maximum bound" in str(x.message) for x in w) == 1
        assert sum("not in valid values" in str(x.message) for x in w) == 1

    def test_update_and_restore_bounds(self):
        try:
            IDWReconstructor.update_bounds(bounds={"k": (2, 9), "nope": (0, 1)}, values={})
            assert IDWReconstructor.k.bounds == (2, 9)
            with pytest.raises(TypeError):
                IDWReconstructor.update_bounds(bounds=[1, 2])
        finally:
            IDWReconstructor.k.restore_default_bounds()
        assert IDWReconstructor.k.bounds == (1, None)

    def test_descriptor_rules(self):
        with pytest.raises(ValueError, match="Cannot specify both"):
            Hyperparameter[int](default=1, bounds=(0, 2), values=[1])
        with pytest.raises(ValueError, match="Lower bound must be less"):
            Hyperparameter[int](default=1, bounds=(3, 2))
        with pytest.raises(ValueError, match="numeric types"):
            Hyperparameter[str](default="a", bounds=(0, 1))

    def test_falsy_kriging_values_become_defaults(self):
        ds = sparse_dataset()
        r = OrdinaryKrigingReconstructor(ds, ds.domain, nlags=0, anisotropy_scaling=0.0)
        assert r.nlags == 6 and r.anisotropy_scaling == 1e-6  # kriging.py:141-144 `x or default`


@pytest.mark.parametrize("cls", [IDWReconstructor, OrdinaryKrigingReconstructor])
class TestConstructorContract:
    @all_datasets
    def test_stores_inputs(self, cls, dataset):
        try:
            r = cls(dataset, dataset.domain)
        except NotImplementedError as e:
            pytest.skip(f"Unsupported configuration: {e}")
        assert r.dataset is dataset and r.target_domain is dataset.domain
        assert r.target_domain.is_sparse == dataset.domain.is_sparse
        assert callable(r.reconstruct) and r.num_params == dataset.domain.get_all_spatial_points().shape[0]

    @all_datasets
    def test_invalid_dataset(self, cls, dataset):
        with pytest.raises(TypeError, match="dataset must be a BaseClimatrixDataset object"):
            cls("not_a_dataset", dataset.domain)

    @all_datasets
    def test_invalid_domain(self, cls, dataset):
        with pytest.raises(TypeError, match="domain must be a Domain object"):
            cls(dataset, "not_a_domain")


class TestIDWConfiguration:
    @all_datasets
    def test_k_min_negative(self, dataset):
        with pytest.raises(ReconstructorConfigurationFailed, match="k_min must be >= 1"):
            IDWReconstructor(dataset, dataset.domain, k_min=-1, k=1)

    @all_datasets
    def test_k_min_bigger_than_k(self, dataset):
        with pytest.raises(ReconstructorConfigurationFailed, match="k_min must be <= k"):
            IDWReconstructor(dataset, dataset.domain, k_min=10, k=2)

    def test_k_zero(self):
        ds = sparse_dataset()
        with pytest.raises(ReconstructorConfigurationFailed):
            IDWReconstructor(ds, ds.domain, k=0, k_min=0)

    def test_dynamic_refused_like_the_reference_on_request(self):
        ds = sparse_dataset(times=TIMES3)
        with pytest.raises(OperationNotSupportedForDynamicDatasetError, match="only supports datasets with"):
            IDWReconstructor(ds, ds.domain, allow_dynamic=False)
        assert issubclass(OperationNotSupportedForDynamicDatasetError, NotImplementedError)
        assert issubclass(ReconstructorConfigurationFailed, ValueError)


class TestKrigingConfiguration:
    def test_dense_source_rejected(self):
        ds = dense_dataset()
        with pytest.raises(OperationNotSupportedForDynamicDatasetError, match="dense dataset"):
            OrdinaryKrigingReconstructor(ds, ds.domain)

    def test_dynamic_rejected(self):
        ds = sparse_dataset(times=TIMES3)
        with pytest.raises(OperationNotSupportedForDynamicDatasetError):
            OrdinaryKrigingReconstructor(ds, ds.domain)

    def test_k_alias(self):
        ds = sparse_dataset()
        assert OrdinaryKrigingReconstructor(ds, ds.domain, k=3).n_closest_points == 3
        assert OrdinaryKrigingReconstructor(ds, ds.domain, n_closest_points=4).n_closest_points == 4
        with pytest.raises(ValueError):
            OrdinaryKrigingReconstructor(ds, ds.domain, k=3, n_closest_points=4)


class TestDomainShim:
    def test_spatial_points_layout(self):
        dd = dense_dataset().domain
        pts = dd.get_all_spatial_points()
        assert pts.shape == (25, 2)
        np.testing.assert_array_equal(pts[:5, 0], -90.0)       # latitude-major (domain.py:917-922)
        np.testing.assert_array_equal(pts[:5, 1], dd.longitude.values)
        sp = sparse_dataset().domain.get_all_spatial_points()
        np.testing.assert_array_equal(sp[:, 0], [-90, -45, 0, 45, 90])

    def test_from_lat_lon(self):
        d = cm.Domain.from_lat_lon(lat=slice(-10, 10, 5), lon=slice(0, 20, 10))
        assert not d.is_sparse and d.shape == (5, 3) and d.size == 15
        s = cm.Domain.from_lat_lon(lat=np.arange(3.0), lon=np.arange(3.0), kind="sparse")
        assert s.is_sparse and s.size == 3
        with pytest.raises(ValueError, match="same length"):
            cm.Domain.from_lat_lon(lat=np.arange(3.0), lon=np.arange(4.0), kind="sparse")
        with pytest.raises(ValueError, match="Unknown kind"):
            cm.Domain.from_lat_lon(kind="weird")

    def test_to_xarray_shapes_and_names(self):
        dd = cm.Domain.from_lat_lon(lat=np.arange(4.0), lon=np.arange(6.0))
        out = dd.to_xarray(np.arange(24.0), "t2m")
        assert out.values.shape == (4, 6) and out.name == "t2m" and tuple(out.dims) == ("lat", "lon")
        sd = cm.Domain.from_lat_lon(lat=np.arange(4.0), lon=np.arange(4.0), kind="sparse")
        out = sd.to_xarray(np.arange(4.0), "x")
        assert out.values.shape == (4,) and tuple(out.dims) == ("point",)
        assert cm.BaseClimatrixDataset(out).domain.is_sparse

    def test_axis_name_families(self):
        from climatrix_b200.dataset import axis_type_of

        assert axis_type_of("lat") == axis_type_of("latitude") == axis_type_of("xlat_m") == cm.AxisType.LATITUDE
        assert axis_type_of("lon") == axis_type_of("longitude") == cm.AxisType.LONGITUDE
        assert axis_type_of("valid_time") == axis_type_of("time") == cm.AxisType.TIME
        assert axis_type_of("isobaricInhPa") is None  # the reference's pattern is lowercase-only too
        assert axis_type_of("level") == axis_type_of("pressure_level") == cm.AxisType.VERTICAL
        assert axis_type_of("point") == axis_type_of("nstations") == axis_type_of("values") == cm.AxisType.POINT
        assert axis_type_of("banana") is None

    def test_foreign_input_rejected(self):
        with pytest.raises(NotImplementedError, match="can be created only based on"):
            cm.BaseClimatrixDataset(np.zeros(3))

    def test_missing_spatial_axis(self):
        with pytest.raises(ValueError, match="has no LONGITUDE axis"):
            cm.Domain(cm.FieldArray(np.zeros(3), ("lat",), {"lat": np.arange(3.0)}))

    def test_unknown_reconstruction_method(self):
        ds = sparse_dataset()
        with pytest.raises(ValueError, match="Unknown reconstruction type"):
            ds.reconstruct(ds.domain, method="nope")


def test_product_path_has_no_cpu_fallback():
    """Without a CUDA device the reconstructors must raise, not compute on the host."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    ds = sparse_dataset()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        IDWReconstructor(ds, ds.domain, k=3).reconstruct()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        OrdinaryKrigingReconstructor(ds, ds.domain, k=3, anisotropy_scaling=1.0).reconstruct()


def test_product_code_does_not_import_the_oracle():
    import os
    import re

    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "climatrix_b200")
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f


def test_ok_rejects_nlags_beyond_the_device_binning_limit():
    """The device variogram keeps at most 64 lag bins (the reference searches 4..20; pykrige has no limit): the
    constructor says so instead of failing later with NotImplementedError."""
    import climatrix_b200 as cm
    from climatrix_b200.exceptions import ReconstructorConfigurationFailed

    rng = np.random.default_rng(0)
    n = 50
    lat, lon = rng.uniform(-50, 50, n), rng.uniform(-100, 100, n)
    src = cm.BaseClimatrixDataset(cm.FieldArray(rng.normal(size=n), ("point",), {"point": np.arange(n), "latitude": (("point",), lat), "longitude": (("point",), lon)}))
    tgt = cm.Domain.from_lat_lon(np.linspace(-40, 40, 5), np.linspace(-90, 90, 7), kind="dense")
    with pytest.raises(ReconstructorConfigurationFailed):
        cm.OrdinaryKrigingReconstructor(src, tgt, nlags=65)
    cm.OrdinaryKrigingReconstructor(src, tgt, nlags=64)
