"""Array-level API over the C ABI: torch CUDA tensors in, torch CUDA tensors out.

PyTorch is used only for device memory, streams and (elsewhere)
``torch.distributed``; every computation happens inside
``libclimatrix_b200.so`` (``include/climatrix_b200.h``).  There is no CPU path:
tensors must live on a CUDA device and the library must have been built.
"""

from __future__ import annotations

import ctypes
import os
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib

__all__ = [
    "Targets",
    "knn",
    "idw",
    "idw_apply",
    "empirical_variogram",
    "variogram_fit_device",
    "fit_variogram_native",
    "ok_moving_window",
    "ok_global",
    "knn3",
    "ok_solve",
    "launch_count",
    "reset_launch_count",
]


@dataclass
class Targets:
    """Target points: a dense (lat x lon) grid given by its axes, or explicit points.

    Mirrors ``DenseDomain`` / ``SparseDomain.get_all_spatial_points``
    (reference dataset/domain.py:917-922 / :738) without materialising the
    (M, 2) array for grids: the kernels generate grid coordinates from the axes.
    """

    lat: torch.Tensor  # float64, cuda; grid: (n_lat,), points: (M,)
    lon: torch.Tensor  # float64, cuda; grid: (n_lon,), points: (M,)
    grid: bool

    @property
    def count(self) -> int:
        return self.lat.numel() * self.lon.numel() if self.grid else self.lat.numel()

    @property
    def shape(self) -> tuple:
        return (self.lat.numel(), self.lon.numel()) if self.grid else (self.lat.numel(),)

    @staticmethod
    def make(lat, lon, grid: bool, device) -> "Targets":
        lat = _as_f64(lat, device)
        lon = _as_f64(lon, device)
        if lat.dim() != 1 or lon.dim() != 1:
            raise ValueError("target latitude/longitude must be one-dimensional")
        if not grid and lat.numel() != lon.numel():
            raise ValueError("For sparse domain, lat and lon must have the same length")
        return Targets(lat, lon, grid)

    def row_band(self, r0: int, r1: int) -> "Targets":
        """Rows [r0, r1) of a grid (or points [r0, r1)): the unit of multi-GPU sharding."""
        if self.grid:
            return Targets(self.lat[r0:r1].contiguous(), self.lon, True)
        return Targets(self.lat[r0:r1].contiguous(), self.lon[r0:r1].contiguous(), False)


def _require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError(
            "climatrix_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback."
        )
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError(f"climatrix_b200 runs on CUDA devices only, got {device}")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


def _as_f64(x, device) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch.float64).contiguous()
    return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64)).to(device, non_blocking=True)


_staging: dict = {}
UPLOAD_CHUNK_BYTES = 32 << 20
_copy_pool: tuple | None = None  # (pid, ThreadPoolExecutor)


def _copy_threads() -> int:
    """Host threads one process may spend on a staging copy: its share of the cores it is allowed to run on
    (``CMX_COPY_THREADS`` overrides; 1 = torch's own copy)."""
    def env_int(name):
        try:
            return int(os.environ.get(name) or 0)
        except ValueError:  # not a number: as if unset
            return 0

    forced = env_int("CMX_COPY_THREADS")
    if forced > 0:
        return forced
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:  # not Linux
        cores = os.cpu_count() or 1
    return max(1, min(8, cores // max(env_int("LOCAL_WORLD_SIZE"), 1)))


def _host_copy(dst: torch.Tensor, src: torch.Tensor) -> None:
    """``dst.copy_(src)`` between large host tensors on several threads.

    torch's CPU copy is parallel only over its intra-op (OpenMP) threads, and ``torchrun`` starts every rank with
    ``OMP_NUM_THREADS=1``: the staging copy of a 296 MB value array then runs at one core's memcpy rate (~45 ms
    instead of ~7 ms) and, with the short kernels of the cell-binned search, becomes the longest item of a sharded call.
    Slices are copied by a small thread pool instead (``copy_`` releases the GIL), independent of the OpenMP setting."""
    global _copy_pool
    threads, n = _copy_threads(), dst.numel()
    if threads <= 1 or torch.get_num_threads() >= threads or n * dst.element_size() < (4 << 20):
        dst.copy_(src)
        return
    if _copy_pool is None or _copy_pool[0] != os.getpid():  # (worker threads do not survive a fork)
        _copy_pool = (os.getpid(), ThreadPoolExecutor(max_workers=threads, thread_name_prefix="cmx-h2d"))
    step = -(-n // threads)
    for f in [_copy_pool[1].submit(dst[o:o + step].copy_, src[o:o + step]) for o in range(0, n, step)]:
        f.result()


def upload(x, device, dtype: torch.dtype) -> torch.Tensor:
    """Host array -> device tensor of ``dtype``.  A large pageable NumPy array goes through two page-locked 32 MB
    staging buffers: a multi-threaded CPU copy (which also converts the dtype; ``_host_copy``) fills one while the DMA
    engine drains the other - 296 MB in 6.5 ms on the B200 host against 27 ms for the driver's own pageable path
    (``tools/prof_h2d.py``).  The buffers are allocated once per dtype and kept."""
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype).contiguous()
    src = torch.as_tensor(np.ascontiguousarray(x))
    if src.numel() * src.element_size() < 2 * UPLOAD_CHUNK_BYTES:
        return src.to(device=device, dtype=dtype, non_blocking=True)
    key = (dtype, device.index)
    if key not in _staging:
        c = UPLOAD_CHUNK_BYTES // torch.empty(0, dtype=dtype).element_size()
        try:
            _staging[key] = ([torch.empty(c, dtype=dtype, pin_memory=True) for _ in range(2)],
                             [torch.cuda.Event() for _ in range(2)])
        except RuntimeError:  # page-locked memory exhausted
            return src.to(device=device, dtype=dtype)
    stage, events = _staging[key]
    flat = src.reshape(-1)
    n, c = flat.numel(), stage[0].numel()
    with torch.cuda.device(device):
        dst = torch.empty(n, dtype=dtype, device=device)
        for i, o in enumerate(range(0, n, c)):
            b = i & 1
            events[b].synchronize()  # the DMA that last read this buffer is done (no-op for a fresh event)
            m = min(c, n - o)
            _host_copy(stage[b][:m], flat[o:o + m])
            dst[o:o + m].copy_(stage[b][:m], non_blocking=True)
            events[b].record()
    return dst.view(src.shape)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


_workspaces: dict = {}


def _workspace(nbytes: int, device: torch.device) -> torch.Tensor:
    """Grow-only scratch per (device, stream): kernels on one stream run in order."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def launch_count() -> int:
    return int(_lib.load().cmx_launch_count())


def reset_launch_count() -> None:
    _lib.load().cmx_reset_launch_count()


KERNEL_KNN, KERNEL_IDW_BATCH, KERNEL_OK_SOLVE, KERNEL_KNN_BINNED = 0, 1, 2, 3

SEARCH_ALGORITHMS = {"auto": _lib.SEARCH_AUTO, "brute": _lib.SEARCH_BRUTE, "binned": _lib.SEARCH_BINNED}
_search = "auto"


def set_search(algorithm: str) -> str:
    """Default planar neighbour search of this process (every call can also name one); returns the previous one.

    ``"auto"`` (default): the faster algorithm for the problem - the cell-binned search whenever its scratch fits.
    ``"brute"``: K1, every (target, sample) pair is evaluated - the brute-force kernel of the north_star.
    ``"binned"``: K1c, samples counting-sorted into grid cells, each target scans only the cells its
    k-th-distance bound touches (the role cKDTree's pruning plays in ``reconstruct/idw.py:155-158``).
    All produce identical neighbour tables.  The choice travels to the library as an explicit per-call
    argument (``cmx_options.search``); the library itself keeps no mode.
    """
    global _search
    if algorithm not in SEARCH_ALGORITHMS:
        raise ValueError(f"unknown search algorithm {algorithm!r}; choose from {sorted(SEARCH_ALGORITHMS)}")
    prev, _search = _search, algorithm
    return prev


def get_search() -> str:
    return _search


def _options(search: str | None = None, peer_out: tuple | None = None, params_on_device: bool = False):
    """``cmx_options`` for one call: search algorithm (None -> the process default), optionally the
    peer buffers ``(pointers, m_total, m_offset)`` the batched IDW kernels store into, and whether the kriging
    entry points find their variogram parameters in device memory."""
    search = _search if search is None else search
    if search not in SEARCH_ALGORITHMS:
        raise ValueError(f"unknown search algorithm {search!r}; choose from {sorted(SEARCH_ALGORITHMS)}")
    opt = _lib.Options()
    opt.search = SEARCH_ALGORITHMS[search]
    opt.n_peers = 0
    opt.params_on_device = int(bool(params_on_device))
    if peer_out is not None:
        ptrs, m_total, m_off = peer_out
        if not 1 <= len(ptrs) <= _lib.MAX_PEERS:
            raise ValueError(f"peer_out needs 1..{_lib.MAX_PEERS} buffers")
        opt.n_peers = len(ptrs)
        for i, q in enumerate(ptrs):
            opt.peer_out[i] = int(q)
        opt.m_total, opt.m_offset = int(m_total), int(m_off)
    return ctypes.byref(opt)


def kernel_timing(enable: bool) -> None:
    """Bracket the hot kernels with CUDA events on their stream (bench.py's roofline)."""
    lib = _lib.load()
    lib.cmx_timing_enable(int(bool(enable)))
    lib.cmx_timing_reset()


def kernel_time(kind: int):
    """(total milliseconds, launches) of one kernel kind since ``kernel_timing(True)``."""
    lib = _lib.load()
    ms, n = ctypes.c_double(0.0), ctypes.c_int64(0)
    _lib.check(lib.cmx_timing_read(kind, ctypes.byref(ms), ctypes.byref(n)), "cmx_timing_read")
    return ms.value, n.value


def fp32_peak_tflops() -> float:
    _require_cuda()
    return float(_lib.load().cmx_fp32_peak_tflops())


def fp64_peak_tflops() -> float:
    _require_cuda()
    return float(_lib.load().cmx_fp64_peak_tflops())


def _knn_ws_bytes(lib, k: int, n: int, targets: "Targets") -> int:
    kk = max(1, min(int(k), _lib.MAX_K))
    return int(lib.cmx_knn_workspace_bytes_for(kk, int(n), targets.lat.numel(), targets.lon.numel(), int(targets.grid)))


def _check_samples(s_lat, s_lon):
    if s_lat.dim() != 1 or s_lat.shape != s_lon.shape:
        raise ValueError(
            "Expected a 2D NumPy array with shape (N, 2) from get_all_spatial_points(), "
            f"but got coordinates with shapes {tuple(s_lat.shape)} and {tuple(s_lon.shape)}."
        )


def knn(s_lat, s_lon, targets: Targets, k: int, *, return_distance: bool = True, search: str | None = None):
    """k nearest samples of every target: ``(dist (M,k) f64 | None, idx (M,k) int32)``.

    Replaces ``cKDTree(points).query(targets, k, workers=-1)`` (idw.py:155-158):
    Euclidean in the flat (lat, lon) plane, rows ascending by (distance, index).
    """
    lib = _lib.load()
    opt = _options(search)
    dev = _require_cuda(targets.lat.device)
    s_lat, s_lon = _as_f64(s_lat, dev), _as_f64(s_lon, dev)
    _check_samples(s_lat, s_lon)
    k = int(k)
    m = targets.count
    with torch.cuda.device(dev):
        idx = torch.empty((m, k), dtype=torch.int32, device=dev)
        dist = torch.empty((m, k), dtype=torch.float64, device=dev) if return_distance else None
        ws = _workspace(_knn_ws_bytes(lib, k, s_lat.numel(), targets), dev)
        rc = lib.cmx_knn(_ptr(s_lat), _ptr(s_lon), s_lat.numel(), _ptr(targets.lat), targets.lat.numel(),
                         _ptr(targets.lon), targets.lon.numel(), int(targets.grid), k, _ptr(idx), _ptr(dist),
                         _ptr(ws), ws.numel(), opt, _stream())
    _lib.check(rc, "cmx_knn")
    return dist, idx


def _values_layout(values: torch.Tensor, n: int, layout: str | None):
    """Returns (tensor, n_batch, layout_code).  Accepts (N,), (T,N) or (N,T)."""
    if values.dim() == 1:
        if values.numel() != n:
            raise ValueError(f"values has {values.numel()} entries for {n} samples")
        return values.contiguous(), 1, _lib.LAYOUT_BATCH_MAJOR
    if values.dim() != 2:
        raise ValueError("values must be (N,), (T, N) or (N, T)")
    if layout is None:
        layout = "sample_major" if (values.shape[0] == n and values.shape[1] != n) else "batch_major"
    if layout == "batch_major":
        if values.shape[1] != n:
            raise ValueError(f"values (T, N) has {values.shape[1]} samples, expected {n}")
        return values.contiguous(), values.shape[0], _lib.LAYOUT_BATCH_MAJOR
    if layout == "sample_major":
        if values.shape[0] != n:
            raise ValueError(f"values (N, T) has {values.shape[0]} samples, expected {n}")
        return values.contiguous(), values.shape[1], _lib.LAYOUT_SAMPLE_MAJOR
    raise ValueError(f"unknown layout {layout!r}")


def idw(s_lat, s_lon, values, targets: Targets, *, k: int = 5, power: float = 2.0, k_min: int = 2,
        dtype: torch.dtype = torch.float64, layout: str | None = None, return_neighbours: bool = False,
        out: torch.Tensor | None = None, peer_out: tuple | None = None, search: str | None = None):
    """Inverse-distance weighting on the device (reference idw.py:155-180).

    ``values``: (N,) one slice, or a time/level batch as (T, N) ("batch_major") or
    (N, T) ("sample_major", the fast layout).  Returns (M,) or (T, M) in ``dtype``;
    with ``return_neighbours`` also ``(dist, idx)``.

    ``peer_out = (pointers, m_total, m_offset)`` (batches only): the kernel stores its rows directly into the
    ``[T][m_total]`` output buffers of several devices at column ``m_offset`` (fused gather over NVLink peer
    memory, see ``sharding.PeerGather``); nothing is returned in that case.
    """
    lib = _lib.load()
    dev = _require_cuda(targets.lat.device)
    if dtype not in (torch.float64, torch.float32):
        raise TypeError("dtype must be torch.float64 or torch.float32")
    s_lat, s_lon = _as_f64(s_lat, dev), _as_f64(s_lon, dev)
    _check_samples(s_lat, s_lon)
    n = s_lat.numel()
    if not isinstance(values, torch.Tensor):
        values = torch.as_tensor(np.ascontiguousarray(values))
    if (values.device.type == "cpu" and values.dim() == 2 and values.numel() >= (1 << 22) and out is None
            and peer_out is None):
        # Host-resident batch: the values are only needed by the weighting kernel, so their host->device
        # copy runs on a side stream while the neighbour search (which needs coordinates only) computes.
        with torch.cuda.device(dev):
            main = torch.cuda.current_stream(dev)
            side = torch.cuda.Stream(dev)
            # search first: the launch is asynchronous, whereas a copy from pageable host memory keeps the host
            # busy staging - so the upload below really runs while the search kernel computes
            dist, idx = knn(s_lat, s_lon, targets, k, search=search)
            with torch.cuda.stream(side):
                vals_d = values.to(device=dev, dtype=dtype, non_blocking=True)
            main.wait_stream(side)
            vals_d.record_stream(main)
            res = idw_apply(idx, dist, vals_d, k=k, power=power, k_min=k_min, dtype=dtype, layout=layout, n_samples=n)
        return (res, dist, idx) if return_neighbours else res
    values = values.to(device=dev, dtype=dtype)
    values, n_batch, layout_code = _values_layout(values, n, layout)
    if n_batch > 1 and layout_code == _lib.LAYOUT_BATCH_MAJOR:
        # the batched kernel gathers one sample's whole series per neighbour: give it (N, T) rows
        values = values.t().contiguous()
        layout_code = _lib.LAYOUT_SAMPLE_MAJOR
    k, k_min = int(k), int(k_min)
    m = targets.count
    with torch.cuda.device(dev):
        if peer_out is not None:
            if n_batch < 2 or return_neighbours:
                raise ValueError("peer_out needs a batch of at least two slices and no neighbour output")
            out = torch.empty(1, dtype=dtype, device=dev)  # placeholder: the call's own output argument is unused
        elif out is None:
            out = torch.empty((n_batch, m) if n_batch > 1 or values.dim() == 2 else (m,), dtype=dtype, device=dev)
        elif out.dtype != dtype or out.numel() != n_batch * m or not out.is_contiguous():
            raise ValueError("out has the wrong dtype, size or is not contiguous")
        idx = dist = None
        if return_neighbours:
            idx = torch.empty((m, k), dtype=torch.int32, device=dev)
            dist = torch.empty((m, k), dtype=torch.float64, device=dev)
        kk = max(1, min(k, _lib.MAX_K))
        ws = _workspace(_knn_ws_bytes(lib, k, n, targets) + lib.cmx_idw_workspace_bytes(kk, m, n_batch) + 1024, dev)
        fn = lib.cmx_idw_f64 if dtype == torch.float64 else lib.cmx_idw_f32
        rc = fn(_ptr(s_lat), _ptr(s_lon), n, _ptr(values), n_batch, layout_code, _ptr(targets.lat),
                targets.lat.numel(), _ptr(targets.lon), targets.lon.numel(), int(targets.grid), k, k_min,
                float(power), _ptr(out), _ptr(idx), _ptr(dist), _ptr(ws), ws.numel(), _options(search, peer_out),
                _stream())
    _lib.check(rc, "cmx_idw")
    if peer_out is not None:
        return None
    if return_neighbours:
        return out, dist, idx
    return out


def idw_apply(idx: torch.Tensor, dist: torch.Tensor, values, *, k: int | None = None, power: float = 2.0,
              k_min: int = 2, dtype: torch.dtype = torch.float64, layout: str | None = None,
              n_samples: int | None = None):
    """Re-weight a stored neighbour table (hyper-parameter search inner loop:
    reference optim/bayesian.py:399-405 calls reconstruct() once per trial; the
    neighbour query only depends on max k, so it is hoisted)."""
    lib = _lib.load()
    dev = _require_cuda(idx.device)
    m, k_table = idx.shape
    k = k_table if k is None else int(k)
    if not isinstance(values, torch.Tensor):
        values = torch.as_tensor(np.ascontiguousarray(values))
    values = values.to(device=dev, dtype=dtype)
    n = int(n_samples) if n_samples is not None else (values.numel() if values.dim() == 1 else None)
    if n is None:
        raise ValueError("n_samples is required for batched values")
    values, n_batch, layout_code = _values_layout(values, n, layout)
    if n_batch > 1 and layout_code == _lib.LAYOUT_BATCH_MAJOR:
        values = values.t().contiguous()
        layout_code = _lib.LAYOUT_SAMPLE_MAJOR
    with torch.cuda.device(dev):
        out = torch.empty((n_batch, m) if values.dim() == 2 else (m,), dtype=dtype, device=dev)
        flag = torch.zeros(1, dtype=torch.int32, device=dev) if n_batch > 1 else None
        fn = lib.cmx_idw_apply_f64 if dtype == torch.float64 else lib.cmx_idw_apply_f32
        rc = fn(_ptr(idx.contiguous()), _ptr(dist.contiguous()), m, k_table, _ptr(values), n, n_batch, layout_code,
                k, int(k_min), float(power), _ptr(out), _ptr(flag), None, _stream())
    _lib.check(rc, "cmx_idw_apply")
    return out


def _score_from_sums(sums: torch.Tensor) -> dict:
    s_abs, s_sq, count = (float(v) for v in sums.cpu())
    if count == 0:
        return {"mae": float("nan"), "mse": float("nan"), "rmse": float("nan"), "count": 0}
    mse = s_sq / count
    return {"mae": s_abs / count, "mse": mse, "rmse": float(np.sqrt(mse)), "count": int(count)}


def idw_score(idx: torch.Tensor, dist: torch.Tensor, values: torch.Tensor, truth: torch.Tensor, *, k: int | None = None,
              power: float = 2.0, k_min: int = 2, return_field: bool = False):
    """One hyper-parameter trial of IDW on a stored neighbour table, scored against held-out values in the same
    kernel (reference: optim/bayesian.py:399-408 -> comparison.py:268-306, nanmean of |d| and d^2).  The field is
    not written unless ``return_field``.  Returns ``{"mae", "mse", "rmse", "count"}`` (and the field)."""
    lib = _lib.load()
    dev = _require_cuda(idx.device)
    m, k_table = idx.shape
    k = k_table if k is None else int(k)
    dtype = values.dtype
    if dtype not in (torch.float64, torch.float32) or truth.dtype != dtype:
        raise TypeError("values and truth must share a float32 or float64 dtype")
    if truth.numel() != m:
        raise ValueError("truth must hold one value per target")
    with torch.cuda.device(dev):
        sums = torch.empty(3, dtype=torch.float64, device=dev)
        out = torch.empty(m, dtype=dtype, device=dev) if return_field else None
        ws = _workspace(lib.cmx_score_workspace_bytes(m), dev)
        fn = lib.cmx_idw_score_f64 if dtype == torch.float64 else lib.cmx_idw_score_f32
        rc = fn(_ptr(idx.contiguous()), _ptr(dist.contiguous()), m, k_table, _ptr(values.contiguous()), values.numel(), k,
                int(k_min), float(power), _ptr(truth.contiguous()), _ptr(out), _ptr(sums), _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "cmx_idw_score")
    score = _score_from_sums(sums)
    return (score, out) if return_field else score


def score(pred: torch.Tensor, truth: torch.Tensor) -> dict:
    """NaN-aware MAE / MSE / RMSE of a device field against held-out values (comparison.py:268-306), reduced on the
    device in a fixed order."""
    lib = _lib.load()
    dev = _require_cuda(pred.device)
    if pred.dtype not in (torch.float64, torch.float32) or truth.dtype != pred.dtype or truth.numel() != pred.numel():
        raise TypeError("pred and truth must have the same float32/float64 dtype and size")
    with torch.cuda.device(dev):
        sums = torch.empty(3, dtype=torch.float64, device=dev)
        ws = _workspace(lib.cmx_score_workspace_bytes(pred.numel()), dev)
        fn = lib.cmx_score_f64 if pred.dtype == torch.float64 else lib.cmx_score_f32
        rc = fn(_ptr(pred.contiguous()), _ptr(truth.contiguous()), pred.numel(), _ptr(sums), _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "cmx_score")
    return _score_from_sums(sums)


# below this many samples every rank of a group bins ALL pairs itself (C2: 5e7 pairs = 0.35 ms on one GPU)
VARIOGRAM_SHARD_MIN_SAMPLES = 16384


def empirical_variogram(x, y, z, nlags: int, *, geographic: bool = False, group=None):
    """pykrige ``_initialize_variogram_model`` binning on the device.

    Returns ``(lags, semivariance)`` NumPy arrays with empty bins dropped, plus the
    raw dict (dmin, dmax, edges, sum_d, sum_g, count).  SURVEY.md Appendix A.1 step 3.

    ``group``: a ``torch.distributed`` process group (one rank per GPU, samples replicated) - every rank sweeps
    its share of the pair blocks and the 2 + 3 nlags partial results are all-reduced.
    """
    lib = _lib.load()
    dev = _require_cuda(x.device if isinstance(x, torch.Tensor) else None)
    x, y, z = _as_f64(x, dev), _as_f64(y, dev), _as_f64(z, dev)
    n = x.numel()
    metric = _lib.METRIC_GEOGRAPHIC if geographic else _lib.METRIC_EUCLIDEAN
    nlags = int(nlags)
    if nlags > 64:
        raise NotImplementedError("nlags above 64 is not supported by the variogram kernel (climatrix's range is 4..20)")
    part, n_parts, dist = 0, 1, None
    if group is not None:
        import torch.distributed as dist

        if dist.is_initialized() and dist.get_world_size(group) > 1 and n >= VARIOGRAM_SHARD_MIN_SAMPLES:
            part, n_parts = dist.get_rank(group), dist.get_world_size(group)
        else:
            dist = None  # (small sample sets: three collectives cost more than the pair sweep they would divide)
    with torch.cuda.device(dev):
        # one packed result [sum_d | sum_g | count | dmin dmax | edges] -> ONE device->host copy (one synchronisation)
        packed = torch.zeros(3 * nlags + 2 + nlags + 1, dtype=torch.float64, device=dev)
        sums = packed[: 2 * nlags].view(2, nlags)
        mm = packed[3 * nlags: 3 * nlags + 2]
        edges_t = packed[3 * nlags + 2:]
        _lib.check(lib.cmx_variogram_minmax(_ptr(x), _ptr(y), n, metric, part, n_parts, _ptr(mm), _stream()),
                   "cmx_variogram_minmax")
        if dist is not None:
            mm[0].neg_()  # one MAX all-reduce for (-min, max)
            dist.all_reduce(mm, op=dist.ReduceOp.MAX, group=group)
            mm[0].neg_()
        # bin edges exactly as pykrige builds them: dmin + n*dd, last edge dmax + 0.001 (same IEEE operations, on the device)
        _lib.check(lib.cmx_variogram_edges(_ptr(mm), nlags, _ptr(edges_t), _stream()), "cmx_variogram_edges")
        count = torch.zeros(nlags, dtype=torch.int64, device=dev)
        ws = _workspace(lib.cmx_variogram_workspace_bytes(nlags), dev)  # deterministic two-stage reduction
        _lib.check(lib.cmx_variogram_bins(_ptr(x), _ptr(y), _ptr(z), n, metric, part, n_parts, nlags, _ptr(edges_t),
                                          _ptr(sums[0]), _ptr(sums[1]), _ptr(count), _ptr(ws), ws.numel(), _stream()),
                   "cmx_variogram_bins")
        if dist is not None:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(count, op=dist.ReduceOp.SUM, group=group)
        packed[2 * nlags: 3 * nlags] = count.to(torch.float64)  # pair counts < 2^53: exact
        host = packed.cpu().numpy()
    sums_h = host[: 2 * nlags].reshape(2, nlags)
    count_h = host[2 * nlags: 3 * nlags].astype(np.int64)
    dmin, dmax = float(host[3 * nlags]), float(host[3 * nlags + 1])
    edges = host[3 * nlags + 2:].tolist()
    good = count_h > 0
    lags = sums_h[0][good] / count_h[good]
    semivariance = sums_h[1][good] / count_h[good]
    raw = dict(dmin=dmin, dmax=dmax, edges=np.array(edges), sum_d=sums_h[0], sum_g=sums_h[1], count=count_h)
    return lags, semivariance, raw


def _variogram_params(params, dev):
    """``(pointer argument, on_device, keep-alive)``: a CUDA tensor of 3 doubles is passed as it is (no host round
    trip: e.g. the output of ``variogram_fit_device``), anything else as a host array of 3 doubles."""
    if isinstance(params, torch.Tensor) and params.device.type == "cuda":
        pt = params.to(device=dev, dtype=torch.float64).contiguous()
        if pt.numel() < 3:
            raise ValueError("device-resident variogram parameters must hold 3 doubles")
        return _ptr(pt), True, pt
    vals = [float(v) for v in (params.tolist() if hasattr(params, "tolist") else list(params))]
    arr = (ctypes.c_double * 3)(*(vals + [0.0] * (3 - len(vals))))
    return ctypes.cast(arr, ctypes.c_void_p), False, arr


def fit_variogram_native(lags, semivariance, variogram_model: str):
    """pykrige ``_calculate_variogram_model`` on the host, natively: the restated SciPy TRF iteration of
    ``csrc/variogram_fit.cuh`` (``cmx_variogram_fit_host``; ~50 us instead of ~4 ms of ``scipy.optimize``).
    Returns the parameter array (2 entries for "linear", else 3)."""
    lib = _lib.load()
    lags = np.ascontiguousarray(lags, dtype=np.float64)
    sv = np.ascontiguousarray(semivariance, dtype=np.float64)
    if lags.ndim != 1 or lags.shape != sv.shape or not 1 <= lags.size <= 64:
        raise ValueError("lags and semivariance must be one-dimensional arrays of the same length (1..64)")
    out = (ctypes.c_double * 3)()
    rc = lib.cmx_variogram_fit_host(_lib.VARIOGRAM_MODELS[variogram_model], lags.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                    sv.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), int(lags.size), out, None)
    if rc < 0:
        _lib.check(rc, "cmx_variogram_fit_host")
    return np.array(out[: 2 if variogram_model == "linear" else 3])


def variogram_fit_device(x, y, z, nlags: int, variogram_model: str, *, geographic: bool = False, group=None):
    """Empirical variogram AND its model fit entirely on the device, without a host round trip:
    pair min/max (K3) -> bin edges -> bin sums (K3) -> TRF fit (one thread), all on the current stream.

    Returns a dict of device tensors: ``params`` (3 doubles; feed it to ``ok_moving_window`` / ``ok_solve``),
    ``lags`` / ``semivariance`` ([nlags], the ``info[1]`` used entries first), ``info`` = (status, points, function
    evaluations), ``minmax``, ``sums`` ([2, nlags]) and ``count``.  Reference: pykrige ``OrdinaryKriging.__init__``
    reached from kriging.py:215-224.  ``group``: shard the pair set over the ranks (all-reduce of 2 + 3 nlags numbers).
    """
    lib = _lib.load()
    dev = _require_cuda(x.device if isinstance(x, torch.Tensor) else None)
    x, y, z = _as_f64(x, dev), _as_f64(y, dev), _as_f64(z, dev)
    n, nlags = x.numel(), int(nlags)
    if nlags > 64:
        raise NotImplementedError("nlags above 64 is not supported by the variogram kernel (climatrix's range is 4..20)")
    metric = _lib.METRIC_GEOGRAPHIC if geographic else _lib.METRIC_EUCLIDEAN
    part, n_parts, dist = 0, 1, None
    if group is not None:
        import torch.distributed as dist

        if dist.is_initialized() and dist.get_world_size(group) > 1 and n >= VARIOGRAM_SHARD_MIN_SAMPLES:
            part, n_parts = dist.get_rank(group), dist.get_world_size(group)
        else:
            dist = None  # (small sample sets: three collectives cost more than the pair sweep they would divide)
    with torch.cuda.device(dev):
        mm = torch.empty(2, dtype=torch.float64, device=dev)
        _lib.check(lib.cmx_variogram_minmax(_ptr(x), _ptr(y), n, metric, part, n_parts, _ptr(mm), _stream()), "cmx_variogram_minmax")
        if dist is not None:
            mm[0].neg_()
            dist.all_reduce(mm, op=dist.ReduceOp.MAX, group=group)
            mm[0].neg_()
        edges = torch.empty(nlags + 1, dtype=torch.float64, device=dev)
        _lib.check(lib.cmx_variogram_edges(_ptr(mm), nlags, _ptr(edges), _stream()), "cmx_variogram_edges")
        sums = torch.zeros((2, nlags), dtype=torch.float64, device=dev)
        count = torch.zeros(nlags, dtype=torch.int64, device=dev)
        ws = _workspace(lib.cmx_variogram_workspace_bytes(nlags), dev)
        _lib.check(lib.cmx_variogram_bins(_ptr(x), _ptr(y), _ptr(z), n, metric, part, n_parts, nlags, _ptr(edges),
                                          _ptr(sums[0]), _ptr(sums[1]), _ptr(count), _ptr(ws), ws.numel(), _stream()),
                   "cmx_variogram_bins")
        if dist is not None:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(count, op=dist.ReduceOp.SUM, group=group)
        params = torch.zeros(3, dtype=torch.float64, device=dev)
        lags = torch.empty(nlags, dtype=torch.float64, device=dev)
        sv = torch.empty(nlags, dtype=torch.float64, device=dev)
        info = torch.zeros(3, dtype=torch.int32, device=dev)
        _lib.check(lib.cmx_variogram_fit(_lib.VARIOGRAM_MODELS[variogram_model], nlags, _ptr(sums[0]), _ptr(sums[1]), _ptr(count),
                                         _ptr(params), _ptr(lags), _ptr(sv), _ptr(info), _stream()), "cmx_variogram_fit")
    return dict(params=params, lags=lags, semivariance=sv, info=info, minmax=mm, edges=edges, sums=sums, count=count)


def ok_moving_window(s_x, s_y, z, t_x, t_y, grid: bool, *, k: int, variogram_model: str, params,
                     geographic: bool = False, exact_values: bool = True, out_scale: float = 1.0,
                     out_shift: float = 0.0, dtype: torch.dtype = torch.float64, return_sigmasq: bool = False,
                     search: str | None = None):
    """Moving-window ordinary kriging in the (already normalised / anisotropy-adjusted) frame.

    pykrige ``execute(style, xpoints, ypoints, backend="loop", n_closest_points=k)``.
    Grid targets: x = columns (n_x), y = rows (n_y); output (n_y * n_x,).  ``dtype`` is the dtype the field
    (and the variance) is stored in; the standardised values ``z`` and all arithmetic are float64.
    """
    lib = _lib.load()
    dev = _require_cuda(s_x.device if isinstance(s_x, torch.Tensor) else None)
    p, on_device, _keep = _variogram_params(params, dev)
    opt = _options(search, params_on_device=on_device)
    s_x, s_y = _as_f64(s_x, dev), _as_f64(s_y, dev)
    t_x, t_y = _as_f64(t_x, dev), _as_f64(t_y, dev)
    if not isinstance(z, torch.Tensor):
        z = torch.as_tensor(np.ascontiguousarray(z))
    z = z.to(device=dev, dtype=torch.float64).contiguous()
    n = s_x.numel()
    m = t_x.numel() * t_y.numel() if grid else t_x.numel()
    if not grid and t_x.numel() != t_y.numel():
        raise ValueError("xpoints and ypoints must have same dimensions when treated as listing discrete points.")
    model = _lib.VARIOGRAM_MODELS[variogram_model]
    k = int(k)
    with torch.cuda.device(dev):
        out = torch.empty(m, dtype=dtype, device=dev)
        sig = torch.empty(m, dtype=dtype, device=dev) if return_sigmasq else None
        nsing = torch.zeros(1, dtype=torch.int32, device=dev)
        kk = max(1, min(k, _lib.MAX_K))
        split_extra = lib.cmx_knn_workspace_bytes_for(kk, n, t_y.numel() if grid else m, t_x.numel(), int(grid)) - lib.cmx_knn_workspace_bytes(kk)
        ws = _workspace(lib.cmx_ok_workspace_bytes(kk, m) + split_extra + 1024, dev)
        fn = lib.cmx_ok_mw_f64 if dtype == torch.float64 else lib.cmx_ok_mw_f32
        rc = fn(_ptr(s_x), _ptr(s_y), _ptr(z), n, _ptr(t_x), t_x.numel(), _ptr(t_y), t_y.numel(), int(grid),
                _lib.METRIC_GEOGRAPHIC if geographic else _lib.METRIC_EUCLIDEAN, k, model, p, int(exact_values),
                float(out_scale), float(out_shift), _ptr(out), _ptr(sig), _ptr(nsing), _ptr(ws), ws.numel(), opt,
                _stream())
    _lib.check(rc, "cmx_ok_mw")
    if return_sigmasq:
        return out, sig, nsing
    return out, nsing


def ok_global(s_x, s_y, z, t_x, t_y, grid: bool, *, variogram_model: str, params, geographic: bool = False,
              exact_values: bool = True, pseudo_inv: bool = False, out_scale: float = 1.0, out_shift: float = 0.0,
              dtype: torch.dtype = torch.float64):
    """Global ordinary kriging (the mode climatrix ships): pykrige ``execute(style, x, y)`` without a moving window.

    One solve of the (N+1) x (N+1) system serves every target (``csrc/global_kriging.cu``: blocked Cholesky of the
    shifted positive-definite matrix; ``pseudo_inv`` or a numerically singular matrix -> one-sided Jacobi SVD, N <= 4095).
    Coordinates in the kriging frame.  Returns ``(field (M,), status)`` with status 0 = the requested path ran, 1 = SVD fallback after a failed Cholesky,
    2 = numerically singular and too large for the SVD path (field undefined).
    """
    lib = _lib.load()
    dev = _require_cuda(s_x.device if isinstance(s_x, torch.Tensor) else None)
    s_x, s_y, z = _as_f64(s_x, dev), _as_f64(s_y, dev), _as_f64(z, dev)
    t_x, t_y = _as_f64(t_x, dev), _as_f64(t_y, dev)
    n = s_x.numel()
    m = t_x.numel() * t_y.numel() if grid else t_x.numel()
    if not grid and t_x.numel() != t_y.numel():
        raise ValueError("xpoints and ypoints must have same dimensions when treated as listing discrete points.")
    if isinstance(params, torch.Tensor):
        params = params.detach().cpu().tolist()
    vals = [float(v) for v in (params.tolist() if hasattr(params, "tolist") else list(params))]
    p = (ctypes.c_double * 3)(*(vals + [0.0] * (3 - len(vals))))
    status = ctypes.c_int32(0)
    with torch.cuda.device(dev):
        out = torch.empty(m, dtype=dtype, device=dev)
        ws = torch.empty(int(lib.cmx_ok_global_workspace_bytes(n)), dtype=torch.uint8, device=dev)  # O(N^2): not cached
        fn = lib.cmx_ok_global_f64 if dtype == torch.float64 else lib.cmx_ok_global_f32
        rc = fn(_ptr(s_x), _ptr(s_y), _ptr(z), n, _ptr(t_x), t_x.numel(), _ptr(t_y), t_y.numel(), int(grid),
                _lib.METRIC_GEOGRAPHIC if geographic else _lib.METRIC_EUCLIDEAN, _lib.VARIOGRAM_MODELS[variogram_model], p,
                int(exact_values), int(pseudo_inv), float(out_scale), float(out_shift), _ptr(out), ctypes.byref(status),
                _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "cmx_ok_global")
    return out, int(status.value)


def knn3(s_xyz, t_xyz, k: int):
    """Exact 3-D k-NN (geographic kriging: chords between unit vectors).  ``s_xyz`` (N, 3), ``t_xyz`` (M, 3).
    Returns ``(d2 (M, k) float64 squared chord lengths, idx (M, k) int32)``, rows ascending by (d2, index)."""
    lib = _lib.load()
    dev = _require_cuda(s_xyz.device if isinstance(s_xyz, torch.Tensor) else None)
    s = [_as_f64(np.asarray(s_xyz)[:, i] if not isinstance(s_xyz, torch.Tensor) else s_xyz[:, i], dev) for i in range(3)]
    t = [_as_f64(np.asarray(t_xyz)[:, i] if not isinstance(t_xyz, torch.Tensor) else t_xyz[:, i], dev) for i in range(3)]
    n, m, k = s[0].numel(), t[0].numel(), int(k)
    with torch.cuda.device(dev):
        idx = torch.empty((m, k), dtype=torch.int32, device=dev)
        d2 = torch.empty((m, k), dtype=torch.float64, device=dev)
        rc = lib.cmx_knn3(_ptr(s[0]), _ptr(s[1]), _ptr(s[2]), n, _ptr(t[0]), _ptr(t[1]), _ptr(t[2]), m, k,
                          _ptr(idx), _ptr(d2), _stream())
    _lib.check(rc, "cmx_knn3")
    return d2, idx


def ok_solve(s_x, s_y, z, idx: torch.Tensor, d2: torch.Tensor, *, variogram_model: str, params,
             geographic: bool = False, exact_values: bool = True, out_scale: float = 1.0, out_shift: float = 0.0,
             dtype: torch.dtype = torch.float64, return_sigmasq: bool = False, solver: str = "auto"):
    """K4 on an existing neighbour table (``idx``, squared distances ``d2``): one kriging window per row.

    ``solver="auto"``: register Cholesky of the shifted (positive-definite) system with hand-over of
    ill-conditioned windows to the pivoted kernel; ``"pivoted"``: Gauss-Jordan with partial pivoting for all."""
    lib = _lib.load()
    dev = _require_cuda(idx.device)
    if solver not in ("auto", "pivoted"):
        raise ValueError("solver must be 'auto' or 'pivoted'")
    s_x, s_y = _as_f64(s_x, dev), _as_f64(s_y, dev)
    if not isinstance(z, torch.Tensor):
        z = torch.as_tensor(np.ascontiguousarray(z))
    z = z.to(device=dev, dtype=torch.float64).contiguous()
    m, k = idx.shape
    p, on_device, _keep = _variogram_params(params, dev)
    with torch.cuda.device(dev):
        out = torch.empty(m, dtype=dtype, device=dev)
        sig = torch.empty(m, dtype=dtype, device=dev) if return_sigmasq else None
        nsing = torch.zeros(1, dtype=torch.int32, device=dev)
        ws = _workspace(lib.cmx_ok_solve_workspace_bytes(m), dev) if solver == "auto" else None
        fn = lib.cmx_ok_solve_f64 if dtype == torch.float64 else lib.cmx_ok_solve_f32
        rc = fn(_ptr(s_x), _ptr(s_y), _ptr(z), s_x.numel(), _ptr(idx.contiguous()), _ptr(d2.contiguous()), m,
                _lib.METRIC_GEOGRAPHIC if geographic else _lib.METRIC_EUCLIDEAN, k, _lib.VARIOGRAM_MODELS[variogram_model],
                p, int(exact_values), float(out_scale), float(out_shift), _ptr(out), _ptr(sig), _ptr(nsing),
                _ptr(ws), ws.numel() if ws is not None else 0, _options(params_on_device=on_device), _stream())
    _lib.check(rc, "cmx_ok_solve")
    if return_sigmasq:
        return out, sig, nsing
    return out, nsing
