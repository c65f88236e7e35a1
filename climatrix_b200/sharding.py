"""Target sharding: one GPU, several GPUs in one process, or one process per GPU (NCCL).

Targets are independent given the (replicated) samples, so the path shards by
contiguous latitude-row bands of the dense target grid (or index ranges of a
sparse target), SURVEY.md section 8e.  The only exchange step is the gather of
the output slabs.

* ``idw_on_devices`` / ``ok_on_devices`` -- host arrays in, host array out; used by the
  reconstructor classes.  ``devices=[...]`` runs the bands concurrently on
  several GPUs of this process: the samples cross PCIe once (to the first device) and reach the
  others over NVLink peer copies, every band is launched asynchronously and its result streams
  into its columns of one page-locked host array.
* ``distributed_idw`` / ``distributed_ok`` -- one process per GPU under
  ``torchrun``: every rank computes its band with the CUDA kernels and the slabs
  are all-gathered over NCCL (NVLink/NVSwitch) or, for IDW batches, stored by the weighting kernel
  itself into the ranks' symmetric-memory buffers.  The band computation is
  injectable so that the partition/gather logic is testable on CPU with gloo.
"""

from __future__ import annotations

from typing import Callable

import numpy as np

__all__ = ["row_shards", "idw_on_devices", "ok_on_devices", "distributed_idw", "distributed_ok", "distributed_bands",
           "assemble_bands", "release_host_memory"]


def row_shards(n_rows: int, parts: int) -> list[tuple[int, int]]:
    """``parts`` contiguous [r0, r1) bands covering ``n_rows``; sizes differ by at most one."""
    parts = max(1, int(parts))
    base, extra = divmod(int(n_rows), parts)
    out, r0 = [], 0
    for p in range(parts):
        r1 = r0 + base + (1 if p < extra else 0)
        out.append((r0, r1))
        r0 = r1
    return out


def idw_on_devices(s_lat, s_lon, values, t_lat, t_lon, grid: bool, *, k: int, power: float, k_min: int,
                   dtype, device=None, devices=None, return_neighbours: bool = False, layout=None):
    """IDW with host (NumPy) inputs and output; H2D and D2H copies included.

    values: (N,) or (T, N).  Returns (M,) or (T, M) as a NumPy array.
    """
    import torch

    from . import kernels

    if devices is None or len(devices) <= 1:
        dev = kernels._require_cuda(devices[0] if devices else device)
        piped = _idw_two_bands(s_lat, s_lon, values, t_lat, t_lon, grid, k=k, power=power, k_min=k_min, dtype=dtype,
                               dev=dev, layout=layout) if not return_neighbours else None
        if piped is not None:
            return piped
        with torch.cuda.device(dev):
            targets = kernels.Targets.make(t_lat, t_lon, grid, dev)
            res = kernels.idw(s_lat, s_lon, values, targets, k=k, power=power, k_min=k_min, dtype=dtype,
                              return_neighbours=return_neighbours, layout=layout)
            if return_neighbours:
                out, dist, idx = res
                return _to_host(out), _to_host(dist), _to_host(idx)
            return _to_host(res)

    # several GPUs driven from this process: every band is launched before any result is awaited
    devs = [kernels._require_cuda(d) for d in devices]
    if return_neighbours:
        raise ValueError("return_neighbours is not available with several devices")
    return _idw_on_several(devs, s_lat, s_lon, values, t_lat, t_lon, grid, k=k, power=power, k_min=k_min, dtype=dtype,
                           layout=layout)


def _idw_on_several(devs, s_lat, s_lon, values, t_lat, t_lon, grid, *, k, power, k_min, dtype, layout):
    """One ``reconstruct`` call on several GPUs of this process (``IDWReconstructor(devices=[...])``).

    The host touches every input once: coordinates and values cross PCIe to the FIRST device (the values through
    the page-locked staging of ``kernels.upload``) and reach the others by peer copies over NVLink, issued on a side
    stream per device.  Order of work, nothing blocking until the end:

    1. coordinates (16 MB at C5) to every device;
    2. brute-force search on a batch: the neighbour search of every band is launched right away - it needs the
       coordinates only and is the long kernel (K1), so the upload and the replication of the values hide behind it;
    3. values to the first device, peer copies to the others;
    4. per device the weighting (on the stored table, or search + weighting in one call for the default search) and
       one strided DMA of the band into its columns of ONE page-locked ``[T, M]`` host array - eight PCIe links
       carry the result concurrently.

    Bit-identical to the single-device call (bands are independent; ``tools/check_multi_device.py``).
    """
    import torch

    from . import _lib, kernels

    lib = _lib.load()
    primary = devs[0]
    n = len(s_lat)
    if not isinstance(values, (np.ndarray, torch.Tensor)):
        values = np.asarray(values)
    probe = values if isinstance(values, torch.Tensor) else torch.as_tensor(values)  # (a view: no copy)
    _, n_batch, layout_code = kernels._values_layout(probe, n, layout)
    single = probe.dim() == 1
    lay = None if single else ("sample_major" if layout_code == 1 else "batch_major")
    n_rows = len(t_lat)
    n_cols = len(t_lon) if grid else 1
    m_total = n_rows * n_cols
    esize = 8 if dtype == torch.float64 else 4
    bands = [(dev, r0, r1) for dev, (r0, r1) in zip(devs, row_shards(n_rows, len(devs))) if r1 > r0]
    try:
        host = torch.empty((n_batch, m_total), dtype=dtype, pin_memory=True)
    except RuntimeError:
        host = torch.empty((n_batch, m_total), dtype=dtype)

    # 1. coordinates: one upload, peer copies
    with torch.cuda.device(primary):
        coords0 = [kernels._as_f64(a, primary) for a in (s_lat, s_lon, t_lat, t_lon)]
    coords = {primary: coords0}
    for dev, _, _ in bands:
        if dev not in coords:
            with torch.cuda.device(dev):
                coords[dev] = [t.to(dev, non_blocking=True) for t in coords0]

    def targets_of(dev, r0, r1):
        _, _, t_lat_d, t_lon_d = coords[dev]
        return kernels.Targets(t_lat_d[r0:r1].contiguous(), t_lon_d if grid else t_lon_d[r0:r1].contiguous(), grid)

    # 2. the long kernel first when it does not need the values.  A tall band is cut in two (like the one-device
    # pipeline, _idw_two_bands): the first part's result crosses PCIe while the second part is still searching.
    split = kernels.get_search() == "brute" and not single

    def parts_of(r0, r1):
        rows = r1 - r0
        if split and grid and rows >= 512:
            cut = r0 + (int(0.92 * rows) // 16) * 16
            return [(r0, cut), (cut, r1)]
        return [(r0, r1)]

    tables = {}
    if split:
        for dev, r0, r1 in bands:
            with torch.cuda.device(dev):
                a, b = parts_of(r0, r1)[0]
                tables[dev, r0] = kernels.knn(coords[dev][0], coords[dev][1], targets_of(dev, a, b), k)

    # 3. values: staged upload to the first device, peer copies on a side stream per device
    side = {dev: torch.cuda.Stream(dev) for dev, _, _ in bands}
    side.setdefault(primary, torch.cuda.Stream(primary))
    with torch.cuda.device(primary), torch.cuda.stream(side[primary]):
        vals0 = kernels.upload(values, primary, dtype)
        vals = {primary: vals0}
        # (torch issues a peer copy on the SOURCE device's current stream: keep that the side stream, or the copies
        # would queue behind the first device's search kernel)
        for dev, _, _ in bands:
            if dev not in vals:
                with torch.cuda.device(dev), torch.cuda.stream(side[dev]):
                    vals[dev] = vals0.to(dev, non_blocking=True)

    # 4. weighting + download, per device
    keep = []
    for dev, r0, r1 in bands:
        with torch.cuda.device(dev):
            main = torch.cuda.current_stream(dev)
            main.wait_stream(side[dev])
            vals[dev].record_stream(main)
            parts = parts_of(r0, r1)
            for i, (a, b) in enumerate(parts):
                if split:
                    dist, idx = tables[dev, r0] if i == 0 else kernels.knn(coords[dev][0], coords[dev][1], targets_of(dev, a, b), k)
                    res = kernels.idw_apply(idx, dist, vals[dev], k=k, power=power, k_min=k_min, dtype=dtype, layout=lay,
                                            n_samples=n)
                    keep.append((dist, idx, main))
                else:
                    res = kernels.idw(coords[dev][0], coords[dev][1], vals[dev], targets_of(dev, a, b), k=k, power=power,
                                      k_min=k_min, dtype=dtype, layout=lay)
                res = res.reshape(n_batch, -1)
                stream = main
                if i + 1 < len(parts):  # a later part of this band still has to search: download beside it
                    stream = torch.cuda.Stream(dev)
                    stream.wait_stream(main)
                    res.record_stream(stream)
                m_part = (b - a) * n_cols
                _lib.check(lib.cmx_copy_rows_to_host(host.data_ptr() + a * n_cols * esize, m_total * esize, res.data_ptr(),
                                                     m_part * esize, m_part * esize, n_batch, stream.cuda_stream),
                           "cmx_copy_rows_to_host")
                keep.append((res, None, stream))
    for entry in keep:
        entry[-1].synchronize()
    if primary not in [d for d, _, _ in bands]:
        side[primary].synchronize()
    out = host.numpy()
    return out.reshape(-1) if single else out


def _bands_on_devices(devs, host_inputs, band_fn, n_rows: int, n_cols: int, n_batch: int, dtype):
    """Row bands of one job on several GPUs of this process.

    ``host_inputs``: NumPy arrays / CPU tensors every device needs (samples, values, target axes).  They are
    uploaded ONCE, to the first device, and replicated to the others by peer copies (NVLink, asynchronous to the
    host).  ``band_fn(device, replicas, r0, r1)`` launches the band's kernels on ``device`` and returns its
    ``[n_batch, (r1 - r0) * n_cols]`` device tensor.  Nothing blocks until every band is in flight; each result is
    then copied (one strided DMA per device, all concurrently) into its columns of a page-locked ``[n_batch, M]`` array.
    """
    import torch

    from . import _lib

    lib = _lib.load()
    primary = devs[0]
    with torch.cuda.device(primary):
        first = [torch.as_tensor(a).to(primary, non_blocking=True) for a in host_inputs]
    bands = row_shards(n_rows, len(devs))
    m_total = n_rows * n_cols
    esize = 8 if dtype == torch.float64 else 4
    try:
        host = torch.empty((n_batch, m_total), dtype=dtype, pin_memory=True)
    except RuntimeError:
        host = torch.empty((n_batch, m_total), dtype=dtype)
    keep = []
    for dev, (r0, r1) in zip(devs, bands):
        if r1 == r0:
            continue
        with torch.cuda.device(dev):
            rep = first if dev == primary else [t.to(dev, non_blocking=True) for t in first]
            res = band_fn(dev, rep, r0, r1).reshape(n_batch, -1)
            m_band = (r1 - r0) * n_cols
            stream = torch.cuda.current_stream(dev)
            _lib.check(lib.cmx_copy_rows_to_host(host.data_ptr() + r0 * n_cols * esize, m_total * esize, res.data_ptr(),
                                                 m_band * esize, m_band * esize, n_batch, stream.cuda_stream),
                       "cmx_copy_rows_to_host")
            keep.append((rep, res, stream))
    for _, _, stream in keep:
        stream.synchronize()
    return host.numpy()


def ok_on_devices(x_s, y_s, z, t_x, t_y, grid: bool, *, k: int, variogram_model: str, params, out_scale: float,
                  out_shift: float, dtype, devices, return_sigmasq: bool = False, search=None):
    """Moving-window kriging (kriging frame, host arrays in / out) with the target rows sharded over several GPUs
    of this process.  Returns ``(field [M], sigmasq [M] | None, n_singular)``."""
    import torch

    from . import kernels

    devs = [kernels._require_cuda(d) for d in devices]
    n_rows = len(t_y) if grid else len(t_x)
    n_cols = len(t_x) if grid else 1
    counters = []

    def band(dev, rep, r0, r1):
        xs, ys, zs, tx, ty = rep
        res = kernels.ok_moving_window(xs, ys, zs, tx if grid else tx[r0:r1].contiguous(), ty[r0:r1].contiguous(), grid,
                                       k=k, variogram_model=variogram_model, params=params, out_scale=out_scale,
                                       out_shift=out_shift, dtype=dtype, return_sigmasq=return_sigmasq, search=search)
        counters.append(res[-1])
        if return_sigmasq:  # field and variance travel together: [2, band]
            return torch.stack((res[0], res[1]))
        return res[0]

    host_inputs = [np.ascontiguousarray(a, dtype=np.float64) for a in (x_s, y_s, z, t_x, t_y)]
    out = _bands_on_devices(devs, host_inputs, band, n_rows, n_cols, 2 if return_sigmasq else 1, dtype)
    n_sing = int(sum(int(c.item()) for c in counters))
    return out[0], (out[1] if return_sigmasq else None), n_sing


def _idw_two_bands(s_lat, s_lon, values, t_lat, t_lon, grid, *, k, power, k_min, dtype, dev, layout,
                   min_bytes: int = 1 << 26, stream_bands: int = 8):
    """Large host-resident batch on a dense grid, one device: hide the copies behind the kernels (or the kernels
    behind the copies).

    The result of a big reconstruction (C5: 1.9 GB) takes tens of milliseconds to cross PCIe.

    * Brute-force search (``set_search("brute")``): nothing can be copied before the search has finished, and the
      search is long.  The target rows are cut into a large first band and a small second one: the first band's
      result streams into its columns of the page-locked output (one strided DMA, ``cmx_copy_rows_to_host``) while
      the second band is still searching; the upload of the values is issued after the first search launch.
    * Cell-binned search (the default): the kernels are short and the copies are the job.  The values go up first
      (``kernels.upload``: page-locked staging, 4x the pageable rate), then ``stream_bands`` equal row bands are
      reconstructed one after the other and each band's download overlaps the next band's kernels - the call takes the
      download of the result plus one band of compute.

    Returns None when the case does not apply.
    """
    import torch

    from . import _lib, kernels

    if not grid or not isinstance(values, np.ndarray) or values.ndim != 2:
        return None
    n_rows, n_cols = len(t_lat), len(t_lon)
    n = len(s_lat)
    if layout not in ("sample_major", "batch_major"):
        return None
    n_batch = values.shape[1] if layout == "sample_major" else values.shape[0]
    esize = 8 if dtype == torch.float64 else 4
    m_total = n_rows * n_cols
    if n_batch < 2 or n_rows < 32 or m_total * n_batch * esize < min_bytes:
        return None
    brute = kernels.get_search() == "brute"
    lib = _lib.load()
    with torch.cuda.device(dev):
        main = torch.cuda.current_stream(dev)
        up, down = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        try:
            host = torch.empty((n_batch, m_total), dtype=dtype, pin_memory=True)
        except RuntimeError:
            return None
        s_lat_d, s_lon_d = kernels._as_f64(s_lat, dev), kernels._as_f64(s_lon, dev)
        t_lat_d, t_lon_d = kernels._as_f64(t_lat, dev), kernels._as_f64(t_lon, dev)
        keep, col = [], 0

        def download(band, r0, r1, stream):
            nonlocal col
            m_band = (r1 - r0) * n_cols
            if stream is down:
                down.wait_stream(main)
                band.record_stream(down)
            _lib.check(lib.cmx_copy_rows_to_host(host.data_ptr() + col * esize, m_total * esize, band.data_ptr(),
                                                 m_band * esize, m_band * esize, n_batch, stream.cuda_stream),
                       "cmx_copy_rows_to_host")
            col += m_band

        if brute:
            # small second band: its search must still cover the first band's download.  A band of a sharded job (a
            # few hundred rows) is not cut again - its search still hides the upload of the values.
            split = (int(0.92 * n_rows) // 16) * 16 if n_rows >= 512 else n_rows
            vals_d = None
            for r0, r1 in ((0, split), (split, n_rows)) if split < n_rows else ((0, n_rows),):
                tg = kernels.Targets(t_lat_d[r0:r1].contiguous(), t_lon_d, True)
                dist, idx = kernels.knn(s_lat_d, s_lon_d, tg, k)  # asynchronous launch
                if vals_d is None:  # staged by the host while the search kernel runs
                    with torch.cuda.stream(up):
                        vals_d = kernels.upload(values, dev, dtype)
                    main.wait_stream(up)
                    vals_d.record_stream(main)
                band = kernels.idw_apply(idx, dist, vals_d, k=k, power=power, k_min=k_min, dtype=dtype, layout=layout,
                                         n_samples=n)
                download(band, r0, r1, down if r1 < n_rows else main)  # the last band has nothing left to overlap with
                keep.append((band, dist, idx))
        else:
            vals_d = kernels.upload(values, dev, dtype)
            for r0, r1 in row_shards(n_rows, max(1, min(stream_bands, n_rows // 32))):
                if r1 == r0:
                    continue
                tg = kernels.Targets(t_lat_d[r0:r1].contiguous(), t_lon_d, True)
                band = kernels.idw(s_lat_d, s_lon_d, vals_d, tg, k=k, power=power, k_min=k_min, dtype=dtype, layout=layout)
                download(band, r0, r1, down)
                keep.append(band)
        down.synchronize()
        main.synchronize()
    return host.numpy()


def _to_host(t):
    """Device -> host NumPy array owned by the caller.

    Large results land directly in page-locked memory: a pinned tensor from torch's caching host
    allocator receives one asynchronous DMA copy and is handed out as the NumPy array (no second,
    pageable copy; the block returns to the allocator's cache when the array is garbage collected).
    """
    import torch

    nbytes = t.numel() * t.element_size()
    if nbytes < (1 << 22):
        return t.cpu().numpy()
    try:
        host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    except RuntimeError:  # page-locked memory exhausted: plain pageable copy
        return t.cpu().numpy()
    host.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return host.numpy()


# --------------------------------------------------------------------------- #
# one process per GPU
# --------------------------------------------------------------------------- #
def assemble_bands(gathered, bands, n_batch: int, n_cols: int):
    """[R][T][rows_max*n_cols] padded slabs -> [T][n_rows*n_cols] in the reference layout
    (time, latitude, longitude), dataset/domain.py:1081-1098."""
    import torch

    pieces = []
    for r, (r0, r1) in enumerate(bands):
        rows = r1 - r0
        if rows:
            pieces.append(gathered[r].view(n_batch, -1)[:, : rows * n_cols])
    return torch.cat(pieces, dim=1)


class PeerGather:
    """Output buffers for the fused weighting + gather: one ``[T][M]`` array per rank in symmetric (peer-mapped)
    memory, so that the batched IDW kernel of every rank stores its band into ALL ranks' arrays through NVLink
    (``cmx_set_output_peers``) and no all-gather / re-assembly pass follows the kernel.

    ``begin()`` makes sure every rank has finished reading the previous result, ``peer_out(r0)`` is what
    ``kernels.idw(..., peer_out=...)`` needs, ``finish()`` waits until every rank's stores have landed and
    returns this rank's (complete) array.  Buffers are cached per shape: the rendezvous is a collective.
    """

    _cache: dict = {}

    def __init__(self, n_batch: int, n_rows: int, n_cols: int, dtype, device, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm

        self.group = group if group is not None else dist.group.WORLD
        self.n_cols = n_cols
        self.m_total = n_rows * n_cols
        self.buf = symm.empty((n_batch, self.m_total), dtype=dtype, device=device)
        self.handle = symm.rendezvous(self.buf, self.group)
        self.ptrs = [int(p) for p in self.handle.buffer_ptrs]

    @classmethod
    def get(cls, n_batch, n_rows, n_cols, dtype, device, group=None):
        key = (n_batch, n_rows, n_cols, dtype, str(device), id(group))
        if key not in cls._cache:
            cls._cache[key] = cls(n_batch, n_rows, n_cols, dtype, device, group)
        return cls._cache[key]

    def begin(self) -> None:
        self.handle.barrier(channel=0, timeout_ms=60000)

    def peer_out(self, r0: int) -> tuple:
        return self.ptrs, self.m_total, r0 * self.n_cols

    def finish(self):
        self.handle.barrier(channel=1, timeout_ms=60000)
        return self.buf


def distributed_bands(compute: Callable, n_rows: int, n_cols: int, n_batch: int, *, dtype, device, group=None,
                      gather: str = "all"):
    """Run ``compute(r0, r1) -> tensor [n_batch, (r1-r0)*n_cols]`` on this rank's band and gather.

    gather="all": every rank returns the full [n_batch, n_rows*n_cols] tensor (all-gather);
    gather="none": returns only the local slab (the caller keeps the field sharded).
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    bands = row_shards(n_rows, world)
    r0, r1 = bands[rank]
    local = compute(r0, r1)
    if local is None:
        local = torch.empty((n_batch, 0), dtype=dtype, device=device)
    local = local.reshape(n_batch, -1)
    if gather == "none" or world == 1:
        return local
    rows_max = max(b[1] - b[0] for b in bands)
    padded = torch.zeros((n_batch, rows_max * n_cols), dtype=dtype, device=device)
    padded[:, : local.shape[1]] = local
    gathered = torch.empty((world, n_batch, rows_max * n_cols), dtype=dtype, device=device)
    dist.all_gather_into_tensor(gathered.view(-1), padded.view(-1), group=group)
    return assemble_bands(gathered, bands, n_batch, n_cols)


def distributed_idw(s_lat, s_lon, values, t_lat, t_lon, grid: bool, *, k: int, power: float = 2.0, k_min: int = 2,
                    dtype=None, layout=None, group=None, gather: str = "all", compute: Callable | None = None):
    """IDW over all ranks of ``group``: samples replicated, target rows sharded, slabs gathered.

    Inputs are device tensors (or NumPy, copied once per rank).  Returns [T, M] (or [M]
    for a single slice) on this rank's device.  ``gather``: "all" (NCCL all-gather of padded slabs + re-assembly),
    "none" (keep the field sharded) or "peer" (batches on a dense grid: the weighting kernel itself stores every
    band into all ranks' symmetric-memory buffers, no collective after it; host-resident values are uploaded first).
    With "peer" the result IS this rank's cached symmetric buffer (one per shape): the next "peer" call of the same
    shape overwrites it - ``clone()`` it to keep it.
    """
    import torch

    from . import kernels

    dtype = dtype or torch.float64
    dev = kernels._require_cuda() if compute is None else torch.device("cpu")
    n_rows = len(t_lat)
    n_cols = len(t_lon) if grid else 1
    single = not (hasattr(values, "dim") and values.dim() == 2) and not (isinstance(values, np.ndarray) and values.ndim == 2)
    n_batch = 1 if single else (values.shape[1] if layout == "sample_major" else values.shape[0])

    custom_compute = compute is not None
    if compute is None:
        def compute(r0, r1):  # noqa: E306
            if r1 == r0:
                return None
            lat_band = t_lat[r0:r1]
            lon_band = t_lon if grid else t_lon[r0:r1]
            targets = kernels.Targets.make(lat_band, lon_band, grid, dev)
            return kernels.idw(s_lat, s_lon, values, targets, k=k, power=power, k_min=k_min, dtype=dtype, layout=layout)

    if gather == "peer":
        # fused weighting + gather: every rank's kernel stores its band into all ranks' buffers (NVLink peer memory)
        import torch.distributed as dist

        if custom_compute or single or not grid or not (dist.is_initialized() and dist.get_world_size(group) > 1):
            gather = "all"  # nothing to fuse for one slice / one rank / a caller-supplied band function: plain path
        else:
            world, rank = dist.get_world_size(group), dist.get_rank(group)
            r0, r1 = row_shards(n_rows, world)[rank]
            pg = PeerGather.get(n_batch, n_rows, n_cols, dtype, dev, group)
            if not isinstance(values, torch.Tensor) or values.device.type != "cuda":
                # the kernel stores into peer memory: its inputs must be on the device before the barrier opens
                values = torch.as_tensor(np.ascontiguousarray(values)).to(device=dev, dtype=dtype)
            pg.begin()
            if r1 > r0:
                targets = kernels.Targets.make(t_lat[r0:r1], t_lon, True, dev)
                kernels.idw(s_lat, s_lon, values, targets, k=k, power=power, k_min=k_min, dtype=dtype, layout=layout,
                            peer_out=pg.peer_out(r0))
            return pg.finish()
    out = distributed_bands(compute, n_rows, n_cols, n_batch, dtype=dtype, device=dev, group=group, gather=gather)
    return out[0] if single else out


# targets x k^2 below which distributed_ok lets every rank solve the whole field (C2: 6.6e7, C4: 4.3e9)
OK_SHARD_MIN_WORK = 2.0e8


def distributed_ok(x_s, y_s, z, t_x, t_y, grid: bool, *, k: int, variogram_model: str, params, out_scale: float = 1.0,
                   out_shift: float = 0.0, dtype=None, group=None, gather: str = "all", return_sigmasq: bool = False,
                   compute: Callable | None = None):
    """Moving-window ordinary kriging over all ranks of ``group`` (BASELINE config 4): samples and the fitted
    variogram parameters replicated, target rows sharded, slabs all-gathered (8 MB at C4: NCCL).

    Arguments are in the kriging frame (normalised / anisotropy-adjusted, as ``kernels.ok_moving_window`` takes
    them); for a grid ``t_y`` are the rows.  Returns ``(field [M], sigmasq [M] | None, n_singular)`` with
    ``n_singular`` summed over the ranks.  The variogram itself is sharded by pair range
    (``kernels.empirical_variogram(..., group=group)``).  ``compute(r0, r1) -> (tensor [n_out, band], n_singular)``
    replaces the CUDA band computation (CPU tests with gloo).
    """
    import torch
    import torch.distributed as dist

    from . import kernels

    dtype = dtype or torch.float64
    dev = kernels._require_cuda() if compute is None else torch.device("cpu")
    n_rows = len(t_y) if grid else len(t_x)
    n_cols = len(t_x) if grid else 1
    n_out = 2 if return_sigmasq else 1
    sing = []

    m_all = n_rows * n_cols
    if (compute is None and dist.is_initialized() and dist.get_world_size(group) > 1 and gather != "none"
            and float(m_all) * k * k < OK_SHARD_MIN_WORK):
        # too small to shard (C2: 0.36 ms of solves on one GPU): every rank solves the whole field - identical results,
        # no collective on the path
        res = kernels.ok_moving_window(x_s, y_s, z, t_x, t_y, grid, k=k, variogram_model=variogram_model, params=params,
                                       out_scale=out_scale, out_shift=out_shift, dtype=dtype, return_sigmasq=return_sigmasq)
        return res[0], (res[1] if return_sigmasq else None), int(res[-1].item())

    if compute is None:
        def compute(r0, r1):  # noqa: E306
            res = kernels.ok_moving_window(x_s, y_s, z, t_x if grid else t_x[r0:r1], t_y[r0:r1], grid, k=k,
                                           variogram_model=variogram_model, params=params, out_scale=out_scale,
                                           out_shift=out_shift, dtype=dtype, return_sigmasq=return_sigmasq)
            field = torch.stack((res[0], res[1])) if return_sigmasq else res[0].view(1, -1)
            return field, res[-1]

    def band(r0, r1):
        if r1 == r0:
            return None
        field, n_sing = compute(r0, r1)
        sing.append(n_sing)
        return field

    out = distributed_bands(band, n_rows, n_cols, n_out, dtype=dtype, device=dev, group=group, gather=gather)
    total = torch.zeros(1, dtype=torch.int64, device=dev)
    for c in sing:
        total += torch.as_tensor(c).to(device=dev, dtype=torch.int64).reshape(-1)[:1]
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    return out[0], (out[1] if return_sigmasq else None), int(total.item())


def release_host_memory() -> None:
    """Results larger than 4 MB are handed out as NumPy arrays backed by page-locked blocks of torch's caching host
    allocator; when such an array is garbage collected its block returns to that cache and STAYS page-locked for
    reuse.  Call this to hand the cached blocks back to the operating system."""
    import torch

    torch._C._host_emptyCache()
