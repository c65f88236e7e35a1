"""Import the UNMODIFIED reference reconstructors from ``/root/reference``.

TEST INFRASTRUCTURE ONLY -- used by ``oracle/make_golden.py`` in the build
container (the reference tree does not exist on the GPU box, so nothing that
runs there may call this).

climatrix imports xarray / cartopy / matplotlib / seaborn / ... at module top
(``src/climatrix/dataset/base.py:11-16``), none of which is installed in this
image.  The numerical bodies of ``IDWReconstructor.reconstruct`` and
``OrdinaryKrigingReconstructor.reconstruct`` only need NumPy/SciPy plus a few
attributes of their ``dataset`` / ``target_domain`` arguments, so we

* register inert stub modules for the absent third-party packages,
* register ``oracle.pykrige_restated`` under the name ``pykrige.ok`` (pykrige
  itself is absent; for that part parity stays unpinned),
* build real reference ``SparseDomain`` / ``DenseDomain`` objects from real
  reference ``Axis`` objects (pure NumPy classes), bypassing only the xarray
  parsing in ``Domain.__new__``,
* capture the array the reconstructor hands to ``target_domain.to_xarray``.

No reference source is copied: the classes are imported from where they lie.
"""

from __future__ import annotations

import importlib
import importlib.machinery
import sys
import types
from unittest import mock

import numpy as np

REFERENCE_SRC = "/root/reference/src"

_ABSENT = (
    "xarray",
    "cartopy",
    "matplotlib",
    "seaborn",
    "netCDF4",
    "cftime",
    "optuna",
    "flask",
    "cdsapi",
    "sklearn_stub_never",
)


class _StubModule(types.ModuleType):
    """A module whose every attribute is an inert MagicMock (sub-modules too)."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        value = mock.MagicMock(name=f"{self.__name__}.{name}")
        setattr(self, name, value)
        return value


class _StubFinder:
    def find_spec(self, fullname, path=None, target=None):
        top = fullname.split(".")[0]
        if top in _ABSENT:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        mod = _StubModule(spec.name)
        mod.__path__ = []
        return mod

    def exec_module(self, module):
        if module.__name__ == "xarray":
            # accessors must stay the real classes, not mocks
            module.register_dataset_accessor = lambda name: (lambda cls: cls)
            module.register_dataarray_accessor = lambda name: (lambda cls: cls)

            class _Captured:
                """Stands in for xr.DataArray inside Domain.to_xarray."""

                def __init__(self, values=None, coords=None, dims=None, name=None, **kw):
                    self.values = np.asarray(values)
                    self.coords = coords
                    self.dims = dims
                    self.name = name

            module.DataArray = _Captured
            module.Dataset = type("Dataset", (), {})


_loaded = None


def load_reference():
    """Returns a namespace with the reference classes (cached)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    # only stub what is genuinely absent
    absent = []
    for name in _ABSENT:
        try:
            if importlib.util.find_spec(name) is None:
                absent.append(name)
        except (ImportError, ValueError):
            absent.append(name)
    globals()["_ABSENT"] = tuple(absent)
    sys.meta_path.append(_StubFinder())
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)

    # pykrige -> restated algorithm (parity unpinned for this part)
    try:
        has_pykrige = importlib.util.find_spec("pykrige") is not None
    except (ImportError, ValueError):
        has_pykrige = False
    if not has_pykrige:
        from oracle import pykrige_restated

        pk = types.ModuleType("pykrige")
        pk.__spec__ = importlib.machinery.ModuleSpec("pykrige", None, is_package=True)
        pk.__path__ = []
        pk_ok = types.ModuleType("pykrige.ok")
        pk_ok.__spec__ = importlib.machinery.ModuleSpec("pykrige.ok", None)
        pk_ok.OrdinaryKriging = pykrige_restated.OrdinaryKriging
        pk.ok = pk_ok
        sys.modules["pykrige"] = pk
        sys.modules["pykrige.ok"] = pk_ok

    import logging

    logging.disable(logging.CRITICAL)  # the reference configures a chatty stdout logger
    idw_mod = importlib.import_module("climatrix.reconstruct.idw")
    krg_mod = importlib.import_module("climatrix.reconstruct.kriging")
    domain_mod = importlib.import_module("climatrix.dataset.domain")
    axis_mod = importlib.import_module("climatrix.dataset.axis")
    base_mod = importlib.import_module("climatrix.dataset.base")
    logging.disable(logging.NOTSET)

    ns = types.SimpleNamespace(
        IDWReconstructor=idw_mod.IDWReconstructor,
        OrdinaryKrigingReconstructor=krg_mod.OrdinaryKrigingReconstructor,
        idw_module=idw_mod,
        kriging_module=krg_mod,
        SparseDomain=domain_mod.SparseDomain,
        DenseDomain=domain_mod.DenseDomain,
        Domain=domain_mod.Domain,
        AxisType=axis_mod.AxisType,
        Axis=axis_mod.Axis,  # factory: picks Latitude/Longitude/Point/Time by name
        BaseClimatrixDataset=base_mod.BaseClimatrixDataset,
        pykrige_is_real=has_pykrige,
    )
    _loaded = ns
    return ns


class _FakeDataArray:
    """The two attributes the reconstructors read from ``dataset.da``."""

    def __init__(self, values, name="field"):
        self.values = np.asarray(values)
        self.name = name


def make_domain(ref, lat, lon, sparse: bool):
    """A real reference Domain built from real reference Axis objects."""
    A = ref.AxisType
    lat = np.asarray(lat, dtype=np.float64)
    lon = np.asarray(lon, dtype=np.float64)
    if sparse:
        dom = object.__new__(ref.SparseDomain)
        dom._axes = {
            A.POINT: ref.Axis("point", np.arange(lat.size), True),
            A.LATITUDE: ref.Axis("latitude", lat, False),
            A.LONGITUDE: ref.Axis("longitude", lon, False),
        }
    else:
        dom = object.__new__(ref.DenseDomain)
        dom._axes = {
            A.LATITUDE: ref.Axis("latitude", lat, True),
            A.LONGITUDE: ref.Axis("longitude", lon, True),
        }
    return dom


def make_dataset(ref, domain, values, name="field"):
    ds = object.__new__(ref.BaseClimatrixDataset)
    ds.da = _FakeDataArray(values, name)
    ds.domain = domain
    return ds


def run_reference(ref, method: str, dataset, target_domain, **kwargs) -> np.ndarray:
    """Run the reference reconstructor's own ``reconstruct()`` and return the array
    it passes to ``target_domain.to_xarray`` (reshaped by the reference itself)."""
    cls = {"idw": ref.IDWReconstructor, "ok": ref.OrdinaryKrigingReconstructor}[method]
    module = {"idw": ref.idw_module, "ok": ref.kriging_module}[method]
    captured = {}

    def _capture(da):
        captured["da"] = da
        return da

    def _to_positive_longitude(self):
        # dataset/base.py:227-231: for a source whose longitude is not a dimension
        # (every sparse source, the only kind OK accepts) the reference re-wraps the
        # same xarray object unchanged; re-wrapping needs xarray, so return self.
        if self.domain.has_axis(ref.AxisType.LONGITUDE) and self.domain.longitude.is_dimension:
            raise NotImplementedError("dense-source longitude roll needs real xarray")
        return self

    with mock.patch.object(ref.BaseClimatrixDataset, "to_positive_longitude", _to_positive_longitude):
        rec = cls(dataset, target_domain, **kwargs)
    with mock.patch.object(module, "BaseClimatrixDataset", _capture):
        rec.reconstruct()
    return np.asarray(captured["da"].values)
