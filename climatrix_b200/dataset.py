"""Array-backed mirror of the slice of climatrix's data model the hot path touches.

The reconstructors only need, from ``dataset`` and ``target_domain``
(SURVEY.md section 8b): ``dataset.da.values`` / ``.name``, ``domain.latitude.values``,
``domain.longitude.values``, ``is_sparse``, the axis bookkeeping used by the
dynamic-dataset check, ``get_all_spatial_points()`` and ``to_xarray(values, name)``.
Reference: src/climatrix/dataset/domain.py (Domain :330, SparseDomain :706,
DenseDomain :889), dataset/axis.py, dataset/base.py (accessor :51-93,
``reconstruct`` :909-958).

xarray is optional here.  Without it a ``FieldArray`` (values + dims + coords + name)
stands in for ``xarray.DataArray``; with it, real DataArrays go in and come out and
the ``.cm`` accessor is registered (unless climatrix itself already did).  Sampling,
subsetting, plotting, arithmetic, I/O are out of scope.
"""

from __future__ import annotations

import re
from enum import Enum
from typing import ClassVar

import numpy as np

from .exceptions import AxisMatchingError, DatasetCreationError, MissingAxisError

try:  # optional
    import xarray as xr  # type: ignore
except Exception:  # noqa: BLE001
    xr = None

__all__ = ["AxisType", "Axis", "FieldArray", "Domain", "SparseDomain", "DenseDomain", "BaseClimatrixDataset"]


class AxisType(str, Enum):
    LATITUDE = "latitude"
    LONGITUDE = "longitude"
    TIME = "time"
    VERTICAL = "vertical"
    POINT = "point"

    def __str__(self) -> str:
        return self.value

    @classmethod
    def get(cls, value) -> "AxisType":
        if isinstance(value, cls):
            return value
        if not isinstance(value, str):
            raise TypeError(f"Invalid axis type: {value!r}. Expected a string or an instance of Axis.")
        try:
            return cls[value.upper()]
        except KeyError:
            raise ValueError(f"Unknown axis type: {value}") from None


# coordinate-name families climatrix recognises (dataset/axis.py:271,287,303,341-344,360)
_NAME_RULES = (
    (AxisType.LATITUDE, re.compile(r"^x?lat[a-z0-9_]*$")),
    (AxisType.LONGITUDE, re.compile(r"^x?lon[a-z0-9_]*$")),
    (AxisType.TIME, re.compile(r"^x?(valid_)?times?[0-9]*$")),
    (
        AxisType.VERTICAL,
        re.compile(r"^(z|lv_|bottom_top|sigma|h(ei)?ght|altitude|depth|isobaric|pres|level|isotherm)[a-z_]*[0-9]*$"),
    ),
    (AxisType.POINT, re.compile(r"^(point.*|values|nstation.*)$")),
)


def axis_type_of(name: str) -> AxisType | None:
    for kind, rule in _NAME_RULES:
        if rule.match(str(name)):
            return kind
    return None


class Axis:
    """One coordinate: its kind, name, values and whether it spans a dimension."""

    __slots__ = ("type", "name", "values", "is_dimension")

    def __init__(self, name: str, values, is_dimension: bool = True, kind: AxisType | None = None):
        kind = kind or axis_type_of(name)
        if kind is None:
            raise AxisMatchingError(f"No matching axis found for name: {name}.")
        self.type = kind
        self.name = name
        vals = np.atleast_1d(np.asarray([] if values is None else values))
        if kind in (AxisType.LATITUDE, AxisType.LONGITUDE) and vals.dtype.kind in "iub":
            vals = vals.astype(np.float64)
        self.values = vals
        self.is_dimension = bool(is_dimension)

    @property
    def size(self) -> int:
        return len(self.values)

    def __eq__(self, other) -> bool:
        if not isinstance(other, Axis):
            return False
        if self.name != other.name or self.size != other.size or self.is_dimension != other.is_dimension:
            return False
        if self.values.dtype.kind in "fiu" and other.values.dtype.kind in "fiu":
            return bool(np.allclose(self.values, other.values, equal_nan=True))
        return bool(np.array_equal(self.values, other.values))

    def __repr__(self) -> str:
        return f"Axis({self.type.value}, name={self.name!r}, size={self.size}, dim={self.is_dimension})"


class FieldArray:
    """Minimal labelled array (the part of ``xarray.DataArray`` the hot path reads and writes)."""

    __slots__ = ("values", "dims", "coords", "name")

    def __init__(self, values, dims, coords=None, name=None):
        self.values = np.asarray(values)
        self.dims = tuple(dims)
        if self.values.ndim != len(self.dims):
            raise DatasetCreationError(f"values have {self.values.ndim} dimensions but {len(self.dims)} dims were named")
        norm = {}
        for key, val in (coords or {}).items():
            if isinstance(val, tuple) and len(val) == 2 and not np.isscalar(val[0]) and isinstance(val[0], (tuple, list, str)):
                cdims, cvals = val
                cdims = (cdims,) if isinstance(cdims, str) else tuple(cdims)
            else:
                cvals = val
                cdims = (key,) if np.ndim(val) > 0 else ()
            norm[key] = (cdims, np.asarray(cvals))
        self.coords = norm
        self.name = name

    @property
    def shape(self):
        return self.values.shape

    def __repr__(self) -> str:
        return f"FieldArray(name={self.name!r}, dims={self.dims}, shape={self.values.shape})"


def _describe(obj):
    """(dims, {coord name: (coord dims, values)}, name) of an xarray object or FieldArray."""
    if isinstance(obj, FieldArray):
        return obj.dims, obj.coords, obj.name
    if xr is not None and isinstance(obj, (xr.DataArray, xr.Dataset)):
        if isinstance(obj, xr.Dataset):
            if len(obj.data_vars) > 1:
                raise ValueError(
                    "Dataset can be created only based on xarray.DataArray or xarray.Dataset with single "
                    "variable objects, but provided xarray.Dataset with multiple data_vars."
                )
            if len(obj.data_vars) == 1:
                obj = obj[list(obj.data_vars)[0]]
        coords = {str(k): (tuple(v.dims), np.asarray(v.values)) for k, v in obj.coords.items()}
        return tuple(str(d) for d in obj.dims), coords, getattr(obj, "name", None)
    raise NotImplementedError(
        "At the moment, dataset can be created only based on xarray.DataArray or single-variable "
        f"xarray.Dataset objects (or climatrix_b200.FieldArray), but provided {type(obj).__name__}"
    )


def _single_var(obj):
    if xr is not None and isinstance(obj, xr.Dataset) and len(obj.data_vars) == 1:
        return obj[list(obj.data_vars)[0]]
    return obj


def _match_axes(obj, sizes=None) -> dict:
    """Dimensions first (in array order), then non-dimension coordinates: domain.py:50-70."""
    dims, coords, _ = _describe(obj)
    axes: dict = {}
    for pos, dim in enumerate(dims):
        kind = axis_type_of(dim)
        if kind is None:
            continue
        if dim in coords and coords[dim][1].ndim > 0:
            vals = coords[dim][1]
        else:  # a dimension without coordinate values: xarray would give 0..n-1
            n = sizes[pos] if sizes is not None else 0
            vals = np.arange(n)
        axes[kind] = Axis(dim, vals, True, kind)
    for name, (cdims, vals) in coords.items():
        kind = axis_type_of(name)
        if kind is None or kind in axes or np.ndim(vals) == 0:
            continue
        axes[kind] = Axis(name, vals, False, kind)
    return axes


class Domain:
    """Coordinate model of a dataset or of a reconstruction target."""

    __slots__ = ("_axes",)
    is_sparse: ClassVar[bool]

    def __new__(cls, obj=None, *, axes: dict | None = None):
        if axes is None:
            obj = _single_var(obj)
            sizes = getattr(obj, "shape", None) if not (xr is not None and isinstance(obj, xr.Dataset)) else None
            axes = _match_axes(obj, sizes)
        for need in (AxisType.LATITUDE, AxisType.LONGITUDE):
            if need not in axes:
                raise ValueError(f"Dataset has no {need.name} axis")
        dense = axes[AxisType.LATITUDE].is_dimension and axes[AxisType.LONGITUDE].is_dimension
        if cls is Domain:
            cls = DenseDomain if dense else SparseDomain
        self = object.__new__(cls)
        self._axes = dict(axes)
        return self

    # -- construction ------------------------------------------------------
    @classmethod
    def from_lat_lon(cls, lat=slice(-90, 90, 0.1), lon=slice(-180, 180, 0.1), kind: str = "dense") -> "Domain":
        """Reference ``Domain.from_lat_lon`` (domain.py:373-442): slices include their stop."""

        def expand(v):
            if isinstance(v, slice):
                return np.arange(v.start, v.stop + v.step, v.step)
            return np.asarray(v)

        lat, lon = expand(lat), expand(lon)
        if len(lat) == 0:
            raise ValueError("Latitude slice produced an empty array. Check that the slice parameters (start, stop, step) are valid.")
        if len(lon) == 0:
            raise ValueError("Longitude slice produced an empty array. Check that the slice parameters (start, stop, step) are valid.")
        if kind == "dense":
            axes = {AxisType.LATITUDE: Axis("lat", lat, True), AxisType.LONGITUDE: Axis("lon", lon, True)}
        elif kind == "sparse":
            if len(lat) != len(lon):
                raise ValueError("For sparse domain, lat and lon must have the same length")
            axes = {
                AxisType.POINT: Axis("point", np.arange(len(lat)), True),
                AxisType.LATITUDE: Axis("lat", lat, False),
                AxisType.LONGITUDE: Axis("lon", lon, False),
            }
        else:
            raise ValueError(f"Unknown kind: {kind}")
        return Domain(axes=axes)

    # -- axis bookkeeping (domain.py:469-611) ------------------------------
    @property
    def axes(self) -> dict:
        return self._axes

    def _need(self, kind: AxisType) -> Axis:
        if kind not in self._axes:
            raise MissingAxisError(f"{kind.name.capitalize()} axis not found in axes ({list(self._axes)})")
        return self._axes[kind]

    latitude = property(lambda self: self._need(AxisType.LATITUDE))
    longitude = property(lambda self: self._need(AxisType.LONGITUDE))
    point = property(lambda self: self._need(AxisType.POINT))
    time = property(lambda self: self._need(AxisType.TIME))
    vertical = property(lambda self: self._need(AxisType.VERTICAL))

    @property
    def all_axes_types(self) -> list:
        return list(self._axes)

    @property
    def dims(self) -> tuple:
        return tuple(k for k, a in self._axes.items() if a.is_dimension)

    def get_size(self, axis) -> int:
        kind = AxisType.get(axis)
        return self._axes[kind].size if kind in self._axes else 1

    def has_axis(self, axis) -> bool:
        return AxisType.get(axis) in self._axes

    def get_axis(self, axis):
        return self._axes.get(AxisType.get(axis))

    @property
    def shape(self) -> tuple:
        return tuple(a.size for a in self._axes.values() if a.is_dimension)

    @property
    def size(self) -> int:
        if AxisType.POINT in self._axes:
            n = self.get_size(AxisType.POINT)
        else:
            n = self.get_size(AxisType.LATITUDE) * self.get_size(AxisType.LONGITUDE)
        if AxisType.TIME in self._axes:
            n *= self.get_size(AxisType.TIME)
        return n

    @property
    def is_dynamic(self) -> bool:
        return AxisType.TIME in self._axes and self.get_size(AxisType.TIME) > 1

    def __eq__(self, other) -> bool:
        return isinstance(other, Domain) and self._axes.keys() == other._axes.keys() and all(
            self._axes[k] == other._axes[k] for k in self._axes
        )

    __hash__ = None

    # -- what the kernels consume / produce --------------------------------
    def get_all_spatial_points(self) -> np.ndarray:
        raise NotImplementedError

    def _wrap(self, values: np.ndarray, name, coords: dict):
        """Reshape to the dimension axes (in stored order) and label: domain.py:859-886 / 1081-1098."""
        dims = self.dims
        values = np.asarray(values).reshape(tuple(self.get_size(ax) for ax in dims))
        dim_names = tuple(self._axes[ax].name for ax in dims)
        if xr is not None:
            xcoords = {k: (v[0], v[1]) if isinstance(v, tuple) else v for k, v in coords.items()}
            return xr.DataArray(values, coords=xcoords, dims=dim_names, name=name)
        return FieldArray(values, dim_names, coords, name)

    def to_xarray(self, values: np.ndarray, name: str | None = None):
        raise NotImplementedError

    def __repr__(self) -> str:
        return f"{type(self).__name__}({', '.join(repr(a) for a in self._axes.values())})"


class SparseDomain(Domain):
    is_sparse: ClassVar[bool] = True

    def get_all_spatial_points(self) -> np.ndarray:
        """(n, 2) rows of [lat, lon]: domain.py:715-738."""
        return np.stack((self.latitude.values, self.longitude.values), axis=1)

    def to_xarray(self, values, name=None):
        point = self.point.name if AxisType.POINT in self._axes else "point"
        coords = {self.latitude.name: ((point,), self.latitude.values), self.longitude.name: ((point,), self.longitude.values)}
        for kind, ax in self._axes.items():
            if ax.name in coords:
                continue
            coords[ax.name] = ((ax.name,), ax.values) if (kind is AxisType.TIME or ax.is_dimension) else ((point,), ax.values)
        return self._wrap(values, name, coords)


class DenseDomain(Domain):
    is_sparse: ClassVar[bool] = False

    def get_all_spatial_points(self) -> np.ndarray:
        """(n_lat*n_lon, 2) rows of [lat, lon], latitude-major: domain.py:898-922."""
        lat, lon = np.meshgrid(self.latitude.values, self.longitude.values, indexing="ij")
        return np.stack((lat.ravel(), lon.ravel()), axis=1)

    def to_xarray(self, values, name=None):
        coords = {ax.name: ((ax.name,), ax.values) for ax in self._axes.values()}
        return self._wrap(values, name, coords)


class BaseClimatrixDataset:
    """``da`` (labelled array) + ``domain``; entry point ``reconstruct`` (dataset/base.py:909-958)."""

    __slots__ = ("da", "domain")

    def __init__(self, obj):
        obj = _single_var(obj)
        if not isinstance(obj, FieldArray) and not (xr is not None and isinstance(obj, xr.DataArray)):
            _describe(obj)  # raises the reference's NotImplementedError for foreign types
        self.domain = Domain(obj)
        self.da = obj

    def reconstruct(self, target: Domain, *, method, **recon_kwargs):
        """``ReconstructionType.get(method).value(self, target_domain=target, **kw).reconstruct()``."""
        from . import idw, kriging  # noqa: F401  (make sure the plug-ins are registered)
        from .plugin import ReconstructionType

        chosen = ReconstructionType.get(method)
        return chosen.value(self, target_domain=target, **recon_kwargs).reconstruct()

    def to_positive_longitude(self):
        """Sparse sources are returned unchanged (dataset/base.py:227-231); the dense roll is
        xarray convenience outside the hot path."""
        lon = self.domain.get_axis(AxisType.LONGITUDE)
        if lon is None or not lon.is_dimension:
            return self
        raise NotImplementedError("to_positive_longitude for dense datasets is outside the reconstruction hot path")

    def __repr__(self) -> str:
        return f"BaseClimatrixDataset({self.da!r})"


def register_xarray_accessor(name: str = "cm") -> bool:
    """Register ``xr_obj.cm`` when xarray is present and climatrix has not done so itself."""
    if xr is None:
        return False
    if hasattr(xr.DataArray, name):
        return False
    xr.register_dataarray_accessor(name)(BaseClimatrixDataset)
    xr.register_dataset_accessor(name)(BaseClimatrixDataset)
    return True
