"""The reconstructor plug-in boundary: registry, method enum, hyper-parameter descriptors.

Host-side mirror of the reference's
  * ``BaseReconstructor``  (src/climatrix/reconstruct/base.py:16-205): registry filled by
    ``__init_subclass__``, ``get`` / ``get_hparams`` / ``get_available_methods`` /
    ``update_bounds``, type checks with the reference's messages;
  * ``ReconstructionType`` (src/climatrix/reconstruct/type.py:8-83): ``get`` / ``list``;
  * ``Hyperparameter``     (src/climatrix/optim/hyperparameter.py:14-187): typed descriptor,
    cast on assignment, out-of-range values only *warn*.
Same names, argument meaning and error behaviour; written for this package (no
reference code is copied).
"""

from __future__ import annotations

import enum
import logging
import numbers
import warnings
from abc import ABC, abstractmethod
from typing import Any, ClassVar

log = logging.getLogger(__name__)

__all__ = ["Hyperparameter", "BaseReconstructor", "ReconstructionType", "reconstruction_type"]


# --------------------------------------------------------------------------- #
# Hyper-parameter descriptor
# --------------------------------------------------------------------------- #
class Hyperparameter:
    """``name = Hyperparameter[float](default=2.0, bounds=(lo, hi))`` on a reconstructor class.

    Assigning on an instance casts to the declared type (``TypeError`` if impossible)
    and warns -- never raises -- when the value leaves ``bounds`` / ``values``.
    ``None`` means "use the default".  Class access returns the descriptor itself so
    that bounds can be edited globally (``update_bounds``), as the reference's
    hyper-parameter search does (optim/bayesian.py:37-57).
    """

    param_type: type | None = None
    _specialised: ClassVar[dict] = {}

    def __class_getitem__(cls, item):
        key = (cls, item)
        if key not in cls._specialised:
            cls._specialised[key] = type(f"{cls.__name__}[{getattr(item, '__name__', item)}]", (cls,), {"param_type": item})
        return cls._specialised[key]

    def __init__(self, *, default, bounds=None, values=None):
        if bounds is not None and values is not None:
            raise ValueError("Cannot specify both bounds and values")
        if bounds is not None:
            if len(bounds) != 2:
                raise ValueError("Bounds must be a tuple of (min_value, max_value)")
            lo, hi = bounds
            if lo is not None and hi is not None and lo >= hi:
                raise ValueError("Lower bound must be less than upper bound")
            t = type(self).param_type
            if not (isinstance(t, type) and issubclass(t, numbers.Number)):
                raise ValueError("Bounds can only be specified for numeric types")
        self.param_type = type(self).param_type
        self.default = default
        self.bounds = bounds
        self.default_bounds = bounds
        self.values = values
        self.name = None
        self.private_name = None

    def __set_name__(self, owner, name):
        self.name = name
        self.private_name = "_" + name

    def __get__(self, instance, owner=None):
        if instance is None:
            return self
        self._require_bound()
        return getattr(instance, self.private_name, self.default)

    def __set__(self, instance, value):
        self._require_bound()
        setattr(instance, self.private_name, self._validate_and_cast(value))

    def _require_bound(self):
        if self.private_name is None:
            raise RuntimeError("Hyperparameter not properly initialized")

    def _validate_and_cast(self, value):
        if value is None:
            return self.default
        if self.param_type is None:
            raise RuntimeError("Parameter type not set. Use Hyperparameter[Type] syntax.")
        try:
            cast = self.param_type(value)
        except (ValueError, TypeError) as exc:
            raise TypeError(
                f"Cannot convert {value!r} to {self.param_type.__name__} for parameter '{self.name}'"
            ) from exc
        if self.bounds is not None and isinstance(cast, numbers.Number):
            lo, hi = self.bounds
            if lo is not None and cast < lo:
                warnings.warn(f"Parameter '{self.name}' value {cast} is below the minimum bound {lo}")
            if hi is not None and cast > hi:
                warnings.warn(f"Parameter '{self.name}' value {cast} is above the maximum bound {hi}")
        if self.values is not None and cast not in self.values:
            warnings.warn(f"Parameter '{self.name}' value {cast!r} not in valid values {self.values}")
        return cast

    def get_spec(self) -> dict[str, Any]:
        spec: dict[str, Any] = {"type": self.param_type}
        for key in ("bounds", "values", "default"):
            if getattr(self, key) is not None:
                spec[key] = getattr(self, key)
        return spec

    def restore_default_bounds(self) -> None:
        self.bounds = self.default_bounds


# --------------------------------------------------------------------------- #
# Registry
# --------------------------------------------------------------------------- #
def _dataset_types() -> tuple:
    from .dataset import BaseClimatrixDataset

    types = [BaseClimatrixDataset]
    try:  # the real accessor class, when climatrix itself is installed
        from climatrix.dataset.base import BaseClimatrixDataset as _Real  # type: ignore

        types.append(_Real)
    except Exception:  # noqa: BLE001
        pass
    return tuple(types)


def _domain_types() -> tuple:
    from .dataset import Domain

    types = [Domain]
    try:
        from climatrix.dataset.domain import Domain as _Real  # type: ignore

        types.append(_Real)
    except Exception:  # noqa: BLE001
        pass
    return tuple(types)


class BaseReconstructor(ABC):
    """``Cls(dataset, target_domain, **hparams).reconstruct() -> BaseClimatrixDataset``."""

    __slots__ = ("dataset", "target_domain")

    _registry: ClassVar[dict[str, type["BaseReconstructor"]]] = {}
    NAME: ClassVar[str] = "<not set>"

    def __init__(self, dataset, target_domain) -> None:
        self.dataset = dataset
        self.target_domain = target_domain
        self._validate_types(dataset, target_domain)

    def __init_subclass__(cls, **kwargs):
        super().__init_subclass__(**kwargs)
        cls._registry[cls.NAME] = cls  # later definitions overwrite, as in the reference
        _invalidate_enum()

    @classmethod
    def get(cls, method: str) -> type["BaseReconstructor"]:
        key = method.lower().strip()
        try:
            return cls._registry[key]
        except KeyError:
            raise ValueError(
                f"Unknown method '{key}'. Available methods: {', '.join(cls._registry)}"
            ) from None

    def _validate_types(self, dataset, domain) -> None:
        if not isinstance(dataset, _dataset_types()):
            raise TypeError("dataset must be a BaseClimatrixDataset object")
        if not isinstance(domain, _domain_types()):
            raise TypeError("domain must be a Domain object")

    @abstractmethod
    def reconstruct(self):
        raise NotImplementedError

    @property
    @abstractmethod
    def num_params(self) -> int:
        raise NotImplementedError

    @classmethod
    def get_hparams(cls) -> dict[str, dict[str, Any]]:
        found = {}
        for name in dir(cls):
            attr = getattr(cls, name)
            if isinstance(attr, Hyperparameter):
                found[name] = attr.get_spec()
        return found

    @classmethod
    def get_available_methods(cls) -> list[str]:
        return list(cls._registry)

    @classmethod
    def update_bounds(cls, bounds: dict | None = None, values: dict | None = None) -> None:
        bounds = {} if bounds is None else bounds
        values = {} if values is None else values
        if not isinstance(bounds, dict) or not isinstance(values, dict):
            raise TypeError(
                "bounds and values must be dictionaries mapping parameter names to tuples or lists"
            )
        known = cls.get_hparams()
        for attr, table in (("bounds", bounds), ("values", values)):
            for name, new in table.items():
                if name in known:
                    log.debug("Updating %s for parameter '%s' to %s", attr, name, new)
                    setattr(getattr(cls, name), attr, new)


# --------------------------------------------------------------------------- #
# Method enum (rebuilt lazily whenever the registry changes -- the reference snapshots
# it once at import, type.py:8-11, which makes late registrations invisible)
# --------------------------------------------------------------------------- #
_enum_cache = None


def _invalidate_enum():
    global _enum_cache
    _enum_cache = None


def _enum_get(cls, value):
    if isinstance(value, cls):
        return value
    if not isinstance(value, str):
        raise TypeError(
            f"Invalid reconstruction type: {value!r}. Expected a string or an instance of ReconstructionType."
        )
    try:
        return cls[value.upper()]
    except KeyError:
        raise ValueError(
            f"Unknown reconstruction type: {value}. Supported types are: ({', '.join(cls.list())})."
            "Ensure that the required packages are installed."
        ) from None


def _enum_list(cls) -> list[str]:
    return [name.lower() for name in cls.__members__]


def reconstruction_type():
    """The ``ReconstructionType`` enum for the current registry (NAME.upper() -> class)."""
    global _enum_cache
    if _enum_cache is None:
        members = {name.upper(): klass for name, klass in BaseReconstructor._registry.items()}
        e = enum.Enum("ReconstructionType", members)
        e.get = classmethod(_enum_get)
        e.list = classmethod(_enum_list)
        _enum_cache = e
    return _enum_cache


class _ReconstructionTypeProxy:
    """``ReconstructionType.IDW``, ``ReconstructionType.get("idw")``, iteration -- always current."""

    def __getattr__(self, name):
        return getattr(reconstruction_type(), name)

    def __getitem__(self, name):
        return reconstruction_type()[name]

    def __call__(self, value):
        return reconstruction_type()(value)

    def __iter__(self):
        return iter(reconstruction_type())

    def __instancecheck__(self, obj):
        return isinstance(obj, reconstruction_type())


ReconstructionType = _ReconstructionTypeProxy()
