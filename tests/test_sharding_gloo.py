"""CPU, world_size 2 over gloo: the multi-GPU partition + gather logic.

The band computation is injected (the CPU oracle stands in for the CUDA kernels, which
cannot run here); what is under test is ``climatrix_b200.sharding``: contiguous row bands,
padding, all-gather and re-assembly into the reference's (time, lat, lon) layout.
"""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from climatrix_b200.sharding import assemble_bands, distributed_idw, row_shards
from oracle.idw_oracle import dense_points, idw_oracle_batched, sparse_points


def test_row_shards_cover_and_balance():
    for n, p in [(721, 8), (1801, 8), (5, 8), (0, 3), (180, 1), (7, 2)]:
        bands = row_shards(n, p)
        assert len(bands) == p and bands[0][0] == 0 and bands[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(bands, bands[1:]))
        sizes = [b[1] - b[0] for b in bands]
        assert max(sizes) - min(sizes) <= 1
    assert row_shards(721, 8)[0] == (0, 91) and row_shards(721, 8)[-1] == (631, 721)


def test_assemble_bands_layout():
    bands = row_shards(5, 2)  # 3 + 2 rows
    n_cols, n_batch = 4, 3
    full = torch.arange(n_batch * 5 * n_cols, dtype=torch.float64).view(n_batch, 5 * n_cols)
    gathered = torch.zeros(2, n_batch, 3 * n_cols, dtype=torch.float64)
    for r, (r0, r1) in enumerate(bands):
        gathered[r, :, : (r1 - r0) * n_cols] = full[:, r0 * n_cols : r1 * n_cols]
    assert torch.equal(assemble_bands(gathered, bands, n_batch, n_cols), full)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_lat, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        n, t = 300, 3
        s_lat, s_lon = rng.uniform(-90, 90, n), rng.uniform(-180, 180, n)
        vals = rng.normal(size=(t, n))
        t_lat, t_lon = np.linspace(-80, 80, n_lat), np.linspace(-170, 170, 9)

        def compute(r0, r1):
            if r1 == r0:
                return None
            ref = idw_oracle_batched(sparse_points(s_lat, s_lon), vals, dense_points(t_lat[r0:r1], t_lon), k=6)
            return torch.from_numpy(ref)

        out = distributed_idw(s_lat, s_lon, vals, t_lat, t_lon, True, k=6, compute=compute, dtype=torch.float64)
        # the fused peer-store gather needs the CUDA kernels: with a caller-supplied band function it must fall
        # back to the collective gather and give the same field
        out_peer = distributed_idw(s_lat, s_lon, vals, t_lat, t_lon, True, k=6, compute=compute, dtype=torch.float64,
                                   gather="peer")
        assert torch.equal(out, out_peer)
        ref = idw_oracle_batched(sparse_points(s_lat, s_lon), vals, dense_points(t_lat, t_lon), k=6)
        q.put((rank, bool(np.array_equal(out.numpy(), ref)), tuple(out.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_lat", [7, 1])
def test_two_rank_gather_matches_single_process(n_lat):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_lat, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, same, shape in results:
        assert same, f"rank {rank} assembled a different field"
        assert shape == (3, n_lat * 9)


def _ok_worker(rank, world, port, n_lat, q):
    """distributed_ok on two gloo ranks: the restated pykrige stands in for K1 + K4 on each band."""
    import warnings

    from climatrix_b200.sharding import distributed_ok
    from oracle import pykrige_restated as pk

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(11)
        n = 200
        x, y = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
        z = np.sin(3 * x) + 0.2 * rng.normal(size=n)
        if n_lat > 3:  # a duplicate pair with zero nugget: singular windows are counted across ranks
            x[1], y[1] = x[0], y[0]
        params = [1.0, 0.5, 0.0]
        tx, ty = np.linspace(-1, 1, 6), np.linspace(-1, 1, n_lat)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            okr = pk.OrdinaryKriging(x, y, z, variogram_model="spherical", variogram_parameters=params)

        def solve(ty_band):
            zz = np.full((len(ty_band), len(tx)), np.nan)
            ss = np.full_like(zz, np.nan)
            bad = 0
            for i, yy in enumerate(ty_band):
                for j, xx in enumerate(tx):
                    try:
                        with warnings.catch_warnings():
                            warnings.simplefilter("error")
                            a, b = okr.execute("points", [xx], [yy], backend="loop", n_closest_points=8)
                        zz[i, j], ss[i, j] = a[0], b[0]
                    except Exception:  # noqa: BLE001  (LinAlgError / LinAlgWarning: singular window)
                        bad += 1
            return zz, ss, bad

        def compute(r0, r1):
            zz, ss, bad = solve(ty[r0:r1])
            return torch.from_numpy(np.stack((zz.ravel(), ss.ravel()))), torch.tensor([bad])

        field, sig, n_sing = distributed_ok(x, y, z, tx, ty, True, k=8, variogram_model="spherical", params=params,
                                            return_sigmasq=True, compute=compute)
        zz, ss, bad = solve(ty)
        same = np.array_equal(field.numpy(), zz.ravel(), equal_nan=True) and np.array_equal(sig.numpy(), ss.ravel(), equal_nan=True)
        q.put((rank, bool(same), tuple(field.shape), n_sing == bad))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_lat", [5, 1])
def test_two_rank_kriging_bands_match_single_process(n_lat):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ok_worker, args=(r, 2, port, n_lat, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, same, shape, counted in results:
        assert same, f"rank {rank} assembled a different field"
        assert shape == (n_lat * 6,)
        assert counted, f"rank {rank}: singular windows not summed over the ranks"
