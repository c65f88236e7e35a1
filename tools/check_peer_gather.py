"""2+ ranks (torchrun): the fused weighting + peer-store gather must equal the NCCL all-gather result bit for bit.
torchrun --nproc-per-node 2 tools/check_peer_gather.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from climatrix_b200 import kernels as K
from climatrix_b200.sharding import PeerGather, distributed_bands, row_shards

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rng = np.random.default_rng(11)
for (N, n_lat, n_lon, T, k) in [(20000, 181, 360, 130, 32), (5000, 91, 180, 37, 16), (3000, 45, 90, 3, 8)]:
    lat = torch.as_tensor(rng.uniform(-90, 90, N)).to(dev); lon = torch.as_tensor(rng.uniform(0, 360, N)).to(dev)
    vals = torch.as_tensor(rng.normal(size=(N, T))).to(dev)
    t_lat = torch.as_tensor(np.linspace(90, -90, n_lat)).to(dev); t_lon = torch.as_tensor(np.linspace(0, 359, n_lon)).to(dev)
    r0, r1 = row_shards(n_lat, world)[rank]

    def compute(a, b):
        if a == b:
            return None
        return K.idw(lat, lon, vals, K.Targets(t_lat[a:b].contiguous(), t_lon, True), k=k, layout="sample_major")

    ref = distributed_bands(compute, n_lat, n_lon, T, dtype=torch.float64, device=dev, gather="all")
    pg = PeerGather.get(T, n_lat, n_lon, torch.float64, dev)
    for rep in range(2):
        pg.begin()
        K.idw(lat, lon, vals, K.Targets(t_lat[r0:r1].contiguous(), t_lon, True), k=k, layout="sample_major",
              peer_out=pg.peer_out(r0))
        got = pg.finish()
        torch.cuda.synchronize()
        same = bool(torch.equal(got, ref.view_as(got)))
        print(f"rank {rank} N={N} T={T} rep {rep}: peer gather == NCCL gather: {same}", flush=True)
        assert same
dist.barrier()
if rank == 0:
    print("peer gather ok")
dist.destroy_process_group()
