"""Target sharding: one GPU, several GPUs in one process, or one process per GPU (NCCL).

Targets are independent given the (replicated) samples, so the path shards by
contiguous latitude-row bands of the dense target grid (or index ranges of a
sparse target), SURVEY.md section 8e.  The only exchange step is the gather of
the output slabs.

* ``idw_on_devices``   -- host arrays in, host array out; used by the
  reconstructor classes.  ``devices=[...]`` runs the bands concurrently on
  several GPUs of this process.
* ``distributed_idw`` / ``distributed_ok`` -- one process per GPU under
  ``torchrun``: every rank computes its band with the CUDA kernels and the slabs
  are all-gathered over NCCL (NVLink/NVSwitch).  The band computation is
  injectable so that the partition/gather logic is testable on CPU with gloo.
"""

from __future__ import annotations

from typing import Callable

import numpy as np

__all__ = ["row_shards", "idw_on_devices", "distributed_idw", "distributed_bands", "assemble_bands"]


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

    # several GPUs driven from this process: launch every band, then collect
    devs = [kernels._require_cuda(d) for d in devices]
    n_rows = len(t_lat)
    bands = row_shards(n_rows, len(devs))
    pending = []
    for dev, (r0, r1) in zip(devs, bands):
        if r1 == r0:
            continue
        with torch.cuda.device(dev):
            lat_band = t_lat[r0:r1]
            lon_band = t_lon if grid else t_lon[r0:r1]
            targets = kernels.Targets.make(lat_band, lon_band, grid, dev)
            out = kernels.idw(s_lat, s_lon, values, targets, k=k, power=power, k_min=k_min, dtype=dtype, layout=layout)
            pending.append(out)
    parts = [_to_host(o) for o in pending]
    return np.concatenate(parts, axis=-1)


def _idw_two_bands(s_lat, s_lon, values, t_lat, t_lon, grid, *, k, power, k_min, dtype, dev, layout,
                   min_bytes: int = 1 << 28):
    """Large host-resident batch on a dense grid, one device: hide the device->host copy of the result.

    The result of a big reconstruction (C5: 1.9 GB) takes tens of milliseconds to cross PCIe, and none of it can
    start before the neighbour search has finished.  So the target rows are cut into a large first band and a
    small second one: the first band's result streams into its columns of the page-locked output (one strided
    DMA, ``cmx_copy_rows_to_host``) while the second band is still searching.  The upload of the values is
    issued after the first search launch, so it overlaps too.  Returns None when the case does not apply.
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
    if n_batch < 2 or n_rows < 256 or m_total * n_batch * esize < min_bytes:
        return None
    split = (int(0.92 * n_rows) // 16) * 16  # small second band: its search must still cover the first band's download
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
        keep, vals_d, col = [], None, 0
        for r0, r1 in ((0, split), (split, n_rows)):
            tg = kernels.Targets(t_lat_d[r0:r1].contiguous(), t_lon_d, True)
            dist, idx = kernels.knn(s_lat_d, s_lon_d, tg, k)  # asynchronous launch
            if vals_d is None:  # pageable upload: the host stages it while the search kernel runs
                with torch.cuda.stream(up):
                    vals_d = torch.as_tensor(values).to(device=dev, dtype=dtype, non_blocking=True)
                main.wait_stream(up)
                vals_d.record_stream(main)
            band = kernels.idw_apply(idx, dist, vals_d, k=k, power=power, k_min=k_min, dtype=dtype, layout=layout,
                                     n_samples=n)
            m_band = (r1 - r0) * n_cols
            stream = down if r1 < n_rows else main  # the last band has nothing left to overlap with
            if stream is down:
                down.wait_stream(main)
                band.record_stream(down)
            _lib.check(lib.cmx_copy_rows_to_host(host.data_ptr() + col * esize, m_total * esize, band.data_ptr(),
                                                 m_band * esize, m_band * esize, n_batch, stream.cuda_stream),
                       "cmx_copy_rows_to_host")
            keep.append((band, dist, idx))
            col += m_band
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
        return local if gather == "none" else local
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
    band into all ranks' symmetric-memory buffers, no collective after it; device-resident ``values`` only).
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
            pg.begin()
            if r1 > r0:
                targets = kernels.Targets.make(t_lat[r0:r1], t_lon, True, dev)
                kernels.idw(s_lat, s_lon, values, targets, k=k, power=power, k_min=k_min, dtype=dtype, layout=layout,
                            peer_out=pg.peer_out(r0))
            return pg.finish()  # NOTE: this rank's cached symmetric buffer; clone() to keep it across calls
    out = distributed_bands(compute, n_rows, n_cols, n_batch, dtype=dtype, device=dev, group=group, gather=gather)
    return out[0] if single else out
