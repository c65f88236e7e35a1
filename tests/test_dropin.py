"""``install_into_climatrix()``: after it, climatrix's own dispatch reaches the GPU classes.

climatrix resolves ``method=`` with ``ReconstructionType.get(method).value(...)``
(reference src/climatrix/dataset/base.py:951-958); the enum is a snapshot of
``BaseReconstructor._registry`` taken at import (reconstruct/type.py:8-11) and ``get`` / ``list`` /
``__missing__`` are module-level functions attached as classmethods (type.py:14-83).

* CPU, build container only: the REAL reference modules (imported from /root/reference through
  ``oracle/reference_loader.py``, in a subprocess because the loader edits ``sys.meta_path``) are
  patched and the reference's own ``BaseClimatrixDataset.reconstruct`` is called.
* Everywhere (and on the GPU with actual data): a stand-in ``climatrix`` package built in memory that
  uses the same registry / enum-snapshot / classmethod mechanism.
"""

import enum
import os
import subprocess
import sys
import textwrap
import types

import numpy as np
import pytest

import climatrix_b200 as cm
from conftest import ROOT, load_golden
from helpers import rel_err, sparse_dataset

REFERENCE = "/root/reference/src/climatrix"


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference tree only exists in the build container")
def test_install_into_the_real_reference_modules():
    script = textwrap.dedent(
        """
        import sys
        sys.path.insert(0, %r)
        import numpy as np
        from oracle.reference_loader import load_reference, make_dataset, make_domain
        ref = load_reference()
        import climatrix.reconstruct.type as ref_type
        import climatrix.dataset.base as ref_base
        import climatrix_b200 as cm
        cpu_idw = ref_type.ReconstructionType.get("idw").value
        assert cpu_idw is ref.IDWReconstructor
        assert cm.install_into_climatrix() is True
        RT = ref_type.ReconstructionType
        assert RT.get("idw").value is cm.IDWReconstructor, RT.get("idw").value
        assert RT.get("ok").value is cm.OrdinaryKrigingReconstructor
        assert RT.get("IDW") is RT["IDW"] and RT.get(RT.OK) is RT.OK
        assert RT["IDW"].value is cm.IDWReconstructor
        assert set(RT.list()) >= {"idw", "ok"}
        try:
            RT.get("nope")
        except ValueError as e:
            assert "Unknown reconstruction type" in str(e)
        else:
            raise AssertionError("unknown method accepted")
        try:
            RT.get(3)
        except TypeError:
            pass
        else:
            raise AssertionError("non-string accepted")
        from climatrix.reconstruct.base import BaseReconstructor
        assert BaseReconstructor.get("idw") is cm.IDWReconstructor
        # the reference's own dispatch (dataset/base.py:909-958) on reference Domain objects
        rng = np.random.default_rng(0)
        src = make_dataset(ref, make_domain(ref, rng.uniform(-90, 90, 50), rng.uniform(-180, 180, 50), True), rng.normal(size=50))
        tgt = make_domain(ref, np.linspace(-80, 80, 5), np.linspace(-170, 170, 7), False)
        import torch
        try:
            ref_base.BaseClimatrixDataset.reconstruct(src, tgt, method="idw", k=4)
        except RuntimeError as e:
            assert "no CPU fallback" in str(e) or torch.cuda.is_available(), e
        else:
            assert torch.cuda.is_available(), "the GPU class must refuse to run without a device"
        print("DROPIN-OK")
        """
    ) % ROOT
    r = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DROPIN-OK" in r.stdout, r.stdout + r.stderr


def _standin_climatrix():
    """An in-memory ``climatrix`` with the reference's plug-in mechanism and nothing else: a registry on
    ``BaseReconstructor``, an Enum snapshot of it with module-level get/list/__missing__ attached as
    classmethods, and a dataset whose ``reconstruct`` dispatches through ``ReconstructionType.get``."""
    pkg = types.ModuleType("climatrix")
    pkg.__path__ = []
    rec = types.ModuleType("climatrix.reconstruct")
    rec.__path__ = []
    base = types.ModuleType("climatrix.reconstruct.base")
    typ = types.ModuleType("climatrix.reconstruct.type")

    class BaseReconstructor:
        _registry = {}

        def __init_subclass__(cls, **kw):
            super().__init_subclass__(**kw)
            cls._registry[cls.NAME] = cls

    class CpuIDW(BaseReconstructor):
        NAME = "idw"

    class CpuOK(BaseReconstructor):
        NAME = "ok"

    base.BaseReconstructor = BaseReconstructor
    typ.ReconstructionType = enum.Enum("ReconstructionType", {n.upper(): c for n, c in BaseReconstructor._registry.items()})

    def __missing__(cls, value):
        raise ValueError(f"Unknown reconstruction method: {value}")

    def get(cls, value):
        if isinstance(value, cls):
            return value
        if not isinstance(value, str):
            raise TypeError(f"Invalid reconstruction type: {value!r}.")
        try:
            return cls[value.upper()]
        except KeyError:
            raise ValueError(f"Unknown reconstruction type: {value}. Supported types are: ({', '.join(cls.list())}).")

    def list_(cls):
        return [name.lower() for name in cls.__members__]

    typ.__missing__, typ.get, typ.list = __missing__, get, list_
    for name, fn in (("__missing__", __missing__), ("get", get), ("list", list_)):
        setattr(typ.ReconstructionType, name, classmethod(fn))

    def dispatch(dataset, target, *, method, **kw):
        rt = sys.modules["climatrix.reconstruct.type"].ReconstructionType
        return rt.get(rt.get(method)).value(dataset, target_domain=target, **kw).reconstruct()

    pkg.reconstruct, rec.base, rec.type = rec, base, typ
    mods = {"climatrix": pkg, "climatrix.reconstruct": rec, "climatrix.reconstruct.base": base, "climatrix.reconstruct.type": typ}
    return mods, dispatch, (CpuIDW, CpuOK)


@pytest.fixture
def standin(monkeypatch):
    mods, dispatch, cpu = _standin_climatrix()
    for name, mod in mods.items():
        monkeypatch.setitem(sys.modules, name, mod)
    return mods, dispatch, cpu


def test_install_rebinds_get_list_missing_on_the_rebuilt_enum(standin):
    mods, _, (cpu_idw, cpu_ok) = standin
    typ = mods["climatrix.reconstruct.type"]
    old = typ.ReconstructionType
    assert old.get("idw").value is cpu_idw
    assert cm.install_into_climatrix() is True
    new = typ.ReconstructionType
    assert new is not old
    assert new.get("idw").value is cm.IDWReconstructor and new.get("ok").value is cm.OrdinaryKrigingReconstructor
    assert new["IDW"].value is cm.IDWReconstructor and new.get(new.IDW) is new.IDW
    assert sorted(new.list()) == ["idw", "ok"]
    with pytest.raises(ValueError, match="Unknown reconstruction type"):
        new.get("nope")
    with pytest.raises(TypeError):
        new.get(1)
    assert mods["climatrix.reconstruct.base"].BaseReconstructor._registry["idw"] is cm.IDWReconstructor


def test_install_returns_false_without_climatrix(monkeypatch):
    for name in [n for n in sys.modules if n == "climatrix" or n.startswith("climatrix.")]:
        monkeypatch.delitem(sys.modules, name)
    monkeypatch.setattr(sys, "path", [p for p in sys.path if "reference" not in p])
    assert cm.install_into_climatrix() is False


@pytest.mark.gpu
def test_dispatch_through_the_patched_enum_runs_on_the_gpu(standin):
    """climatrix's ``reconstruct(target, method="idw")`` dispatch lands in the CUDA path and reproduces the
    fixture the unmodified reference produced (tests/golden/idw_defaults.npz)."""
    _, dispatch, _ = standin
    assert cm.install_into_climatrix() is True
    g = load_golden("idw_defaults")
    src = sparse_dataset(g["sample_lat"], g["sample_lon"], np.asarray(g["values"]).reshape(-1, 1))
    tgt = cm.Domain.from_lat_lon(g["target_lat"], g["target_lon"], kind="sparse" if g["meta"]["target_sparse"] else "dense")
    out = dispatch(src, tgt, method="idw", **g["meta"]["kwargs"])
    assert isinstance(out, cm.BaseClimatrixDataset)
    assert rel_err(np.asarray(out.da.values), g["expected"]) <= 1e-5
    g = load_golden("ok_mw_defaults_k12")
    src = sparse_dataset(g["sample_lat"], g["sample_lon"], np.asarray(g["values"]).reshape(-1, 1))
    tgt = cm.Domain.from_lat_lon(g["target_lat"], g["target_lon"], kind="sparse" if g["meta"]["target_sparse"] else "dense")
    out = dispatch(src, tgt, method="ok", n_closest_points=g["meta"]["n_closest_points"], **g["meta"]["kwargs"])
    got = np.asarray(out.da.values, dtype=np.float64)
    assert np.max(np.abs(got - g["expected"])) <= 1e-4 * np.max(np.abs(g["expected"]))
