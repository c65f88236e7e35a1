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
