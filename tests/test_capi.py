"""CPU: libclimatrix_b200.so loads and exports every symbol include/climatrix_b200.h declares.

No compute call is made here (there is no GPU in the build container); only entry points
that return before touching the device are exercised.
"""

import ctypes
import os

import pytest

from climatrix_b200 import _lib
from climatrix_b200.build import LIB_PATH, build


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB_PATH):
        build()
    return _lib.load()


def test_every_declared_symbol_is_exported_and_bound(lib):
    declared = _lib.header_symbols()
    assert len(declared) >= 17
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but missing from the library"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(declared)


def test_version_and_error_strings(lib):
    assert lib.cmx_version() >= 200
    assert lib.cmx_strerror(0) == b"ok"
    assert b"exceeds the number of samples" in lib.cmx_strerror(_lib.ERR_K_GT_N)
    assert lib.cmx_strerror(-99) == b"unknown error"


def test_workspace_queries(lib):
    assert lib.cmx_knn_workspace_bytes(0) == 0
    assert lib.cmx_knn_workspace_bytes(_lib.MAX_K + 1) == 0
    w32, w64, w128 = (lib.cmx_knn_workspace_bytes(k) for k in (32, 64, 128))
    assert 0 < w32 < w64 < w128
    assert lib.cmx_idw_workspace_bytes(32, 1000, 1) == 0
    assert lib.cmx_idw_workspace_bytes(32, 1000, 4) >= 1000 * 32 * 12
    assert lib.cmx_ok_workspace_bytes(32, 1000) > w32


def test_argument_validation_returns_codes_without_a_device(lib):
    null = ctypes.c_void_p(None)
    one = ctypes.c_void_p(16)  # never dereferenced: validation fails first
    # null pointers
    assert lib.cmx_knn(null, null, 10, null, 1, null, 1, 1, 3, null, null, null, 0, None, null) == _lib.ERR_BAD_ARG
    # k > n
    assert lib.cmx_knn(one, one, 2, one, 1, one, 1, 1, 3, one, null, one, 1 << 40, None, null) == _lib.ERR_K_GT_N
    # k above CMX_MAX_K
    assert lib.cmx_knn(one, one, 1000, one, 1, one, 1, 1, 129, one, null, one, 1 << 40, None, null) == _lib.ERR_K_TOO_LARGE
    # workspace too small
    assert lib.cmx_knn(one, one, 1000, one, 4, one, 4, 1, 8, one, null, one, 16, None, null) == _lib.ERR_WORKSPACE
    # sparse targets need matching lengths
    assert lib.cmx_knn(one, one, 1000, one, 4, one, 5, 0, 8, one, null, one, 1 << 40, None, null) == _lib.ERR_BAD_ARG
    # variogram: too many lags / unknown metric
    assert lib.cmx_variogram_bins(one, one, one, 10, 0, 0, 1, 65, one, one, one, one, null, 0, null) == _lib.ERR_UNSUPPORTED
    assert lib.cmx_variogram_workspace_bytes(6) > 0 and lib.cmx_variogram_workspace_bytes(65) == 0
    assert lib.cmx_variogram_minmax(one, one, 10, 7, 0, 1, one, null) == _lib.ERR_UNSUPPORTED
    assert lib.cmx_variogram_minmax(one, one, 10, 0, 2, 2, one, null) == _lib.ERR_BAD_ARG  # part must be < n_parts
    p = ctypes.cast((ctypes.c_double * 3)(1.0, 1.0, 0.0), ctypes.c_void_p)
    assert lib.cmx_ok_mw_f64(one, one, one, 10, one, 2, one, 2, 1, 0, 4, 9, p, 1, 1.0, 0.0, one, null, null, one, 1 << 40,
                             None, null) == _lib.ERR_UNSUPPORTED
    assert lib.cmx_ok_solve_f64(one, one, one, 10, one, one, 5, 0, 200, 0, p, 1, 1.0, 0.0, one, null, null, null, 0,
                                None, null) == _lib.ERR_K_TOO_LARGE
    assert lib.cmx_ok_solve_workspace_bytes(1000) >= 4000


def test_check_maps_codes_to_python_exceptions():
    with pytest.raises(IndexError):
        _lib.check(_lib.ERR_K_GT_N, "x")  # the reference fails with IndexError too (idw.py:167)
    with pytest.raises(ValueError):
        _lib.check(_lib.ERR_BAD_ARG, "x")
    with pytest.raises(NotImplementedError):
        _lib.check(_lib.ERR_UNSUPPORTED, "x")
    with pytest.raises(RuntimeError):
        _lib.check(_lib.ERR_CUDA, "x")
    _lib.check(0)


def test_options_are_explicit_arguments_and_validated_before_any_device_work(lib):
    """cmx_options replaces the per-thread cmx_set_search / cmx_set_output_peers of ABI 0.1: no hidden state."""
    for gone in ("cmx_set_search", "cmx_get_search", "cmx_set_output_peers"):
        assert not hasattr(lib, gone)
    null, one = ctypes.c_void_p(None), ctypes.c_void_p(16)

    def knn_with(opt):  # k > n would be reported next: proves the options were accepted
        return lib.cmx_knn(one, one, 2, one, 1, one, 1, 1, 3, one, null, one, 1 << 40, ctypes.byref(opt), null)

    opt = _lib.Options()
    for search in (_lib.SEARCH_AUTO, _lib.SEARCH_BRUTE, _lib.SEARCH_BINNED):
        opt.search = search
        assert knn_with(opt) == _lib.ERR_K_GT_N
    opt.search = 7
    assert knn_with(opt) == _lib.ERR_BAD_ARG
    opt.search = _lib.SEARCH_AUTO
    opt.n_peers = 2  # cmx_knn has no redirectable output
    opt.peer_out[0] = opt.peer_out[1] = 16
    opt.m_total, opt.m_offset = 10, 0
    assert knn_with(opt) == _lib.ERR_UNSUPPORTED

    def idw_with(opt):  # batched call; n_batch < 1 would be BAD_ARG, so use k_min = 0 as the next error
        return lib.cmx_idw_f64(one, one, 100, one, 4, 1, one, 2, one, 2, 1, 3, 0, 2.0, one, null, null, one, 1 << 40,
                               ctypes.byref(opt), null)

    opt.n_peers = 9
    assert idw_with(opt) == _lib.ERR_BAD_ARG
    opt.n_peers = 2
    opt.m_offset = 11  # beyond m_total
    assert idw_with(opt) == _lib.ERR_BAD_ARG
    opt.m_offset = 5
    opt.peer_out[1] = None
    assert idw_with(opt) == _lib.ERR_BAD_ARG
    opt.peer_out[1] = 16
    assert idw_with(opt) == _lib.ERR_BAD_ARG  # options fine -> the k_min = 0 argument error

    import climatrix_b200 as cm

    assert cm.get_search() == "auto"
    assert cm.set_search("binned") == "auto"
    assert cm.set_search("brute") == "binned"
    assert cm.set_search("auto") == "brute"
    with pytest.raises(ValueError):
        cm.set_search("kd-tree")


def test_planner_choices_for_bands_of_a_big_job():
    """make_plan's cost model must keep picking what the B200 sweep measured as best
    (profiles/r01_planner_sweep.txt); without a device the library plans for 148 SMs."""
    import ctypes

    lib = _lib.load()

    def plan(n, rows, cols, grid=1):
        q, s = ctypes.c_int(0), ctypes.c_int(0)
        assert lib.cmx_knn_plan(n, rows, cols, grid, ctypes.byref(q), ctypes.byref(s)) == 0
        return q.value, s.value

    assert plan(1_000_000, 1801, 3600) == (16, 1)      # the whole C5 grid: big tiles, no split
    assert plan(1_000_000, 153, 3600) == (16, 2)       # second band of the host pipeline
    assert plan(1_000_000, 225, 3600) == (8, 4)        # one rank of an 8-GPU run
    assert plan(1_000_000, 450, 3600)[1] in (2, 4)     # 4 GPUs: measured tie
    assert plan(100_000, 721, 1440) == (2, 1)          # C3: re-rank bound, small tiles spread the serial re-ranks
    assert plan(1_000, 180, 360)[0] in (1, 2)          # C1: latency bound, spread over the SMs
    assert lib.cmx_knn_plan(10, 1, 1, 1, None, None) == _lib.ERR_BAD_ARG
