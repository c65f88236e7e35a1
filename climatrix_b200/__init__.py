"""climatrix_b200 -- B200-native IDW / ordinary-kriging reconstruction (one hot path of climatrix).

Hand-written sm_100a CUDA behind a C ABI (``include/climatrix_b200.h``) and a Python
host layer that mirrors climatrix's reconstructor plug-in API: ``method="idw"`` and
``method="ok"`` are drop-ins.  No CPU fallback: without the built library or a CUDA
device the reconstructors raise.

    from climatrix_b200 import BaseClimatrixDataset, Domain, FieldArray
    sparse = BaseClimatrixDataset(FieldArray(values, ("point",), {"latitude": ("point", lat), ...}))
    dense = sparse.reconstruct(Domain.from_lat_lon(lat=..., lon=..., kind="dense"), method="idw", k=32)
"""

__version__ = "0.1.0"

from .dataset import (  # noqa: F401
    Axis,
    AxisType,
    BaseClimatrixDataset,
    DenseDomain,
    Domain,
    FieldArray,
    SparseDomain,
    register_xarray_accessor,
)
from .exceptions import (  # noqa: F401
    OperationNotSupportedForDynamicDatasetError,
    ReconstructorConfigurationFailed,
)
from .plugin import BaseReconstructor, Hyperparameter, ReconstructionType  # noqa: F401
from .idw import IDWReconstructor  # noqa: F401
from .kriging import OrdinaryKrigingReconstructor  # noqa: F401


def set_search(algorithm: str) -> str:
    """Neighbour search used by both reconstructors: ``"auto"`` (default: the cell-binned search K1c whenever its
    scratch fits), ``"brute"`` (K1, every pair evaluated - the north_star kernel) or ``"binned"`` (K1c; same results,
    ~N/k times fewer distance evaluations).  Returns the previous setting."""
    from . import kernels

    return kernels.set_search(algorithm)


def get_search() -> str:
    from . import kernels

    return kernels.get_search()


def install_into_climatrix() -> bool:
    """Register the GPU reconstructors in a real climatrix installation.

    climatrix resolves ``method=`` through ``ReconstructionType``, an Enum snapshotted
    from ``BaseReconstructor._registry`` at first import (reconstruct/type.py:8-11), so the
    registry entry is replaced *and* the enum is rebuilt.  Returns False when climatrix
    is not importable.
    """
    try:
        import climatrix.reconstruct.base as ref_base  # type: ignore
        import climatrix.reconstruct.type as ref_type  # type: ignore
    except Exception:  # noqa: BLE001
        return False
    import enum

    ref_base.BaseReconstructor._registry["idw"] = IDWReconstructor
    ref_base.BaseReconstructor._registry["ok"] = OrdinaryKrigingReconstructor
    members = {name.upper(): cls for name, cls in ref_base.BaseReconstructor._registry.items()}
    new_enum = enum.Enum("ReconstructionType", members)
    # type.py:14-76 defines get / list / __missing__ as MODULE-LEVEL functions and attaches them as
    # classmethods; re-attach the same functions so that they bind to the rebuilt enum (copying the bound
    # methods of the old enum would keep resolving "idw" to the CPU class)
    old_enum = ref_type.ReconstructionType
    for attr in ("get", "list", "__missing__"):
        fn = getattr(ref_type, attr, None)
        if not callable(fn) or isinstance(fn, type):
            bound = getattr(old_enum, attr, None)
            fn = getattr(bound, "__func__", None)
        if fn is not None:
            setattr(new_enum, attr, classmethod(fn))
    ref_type.ReconstructionType = new_enum
    # modules that did `from climatrix.reconstruct.type import ReconstructionType` at import time hold the old enum
    import sys

    for mod in list(sys.modules.values()):
        name = getattr(mod, "__name__", "") or ""
        if name.startswith("climatrix") and getattr(mod, "ReconstructionType", None) is old_enum:
            setattr(mod, "ReconstructionType", new_enum)
    return True


register_xarray_accessor()
