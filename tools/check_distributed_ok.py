"""torchrun --nproc-per-node N tools/check_distributed_ok.py: OrdinaryKrigingReconstructor(group=WORLD) - pair-sharded
variogram + row bands + all-gather - against the single-rank reconstruction computed on every rank."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import climatrix_b200 as cm

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rng = np.random.default_rng(3)
n = 30_000
lat, lon = rng.uniform(-90, 90, n), rng.uniform(-180, 180, n)
v = 15 + 10 * np.cos(np.deg2rad(lat)) + rng.normal(size=n)
src = cm.BaseClimatrixDataset(cm.FieldArray(v, ("point",), {"point": np.arange(n), "latitude": (("point",), lat), "longitude": (("point",), lon)}))
tgt = cm.Domain.from_lat_lon(np.linspace(-89, 89, 181), np.linspace(-179, 179, 361), kind="dense")
kw = dict(n_closest_points=32, variogram_model="spherical", anisotropy_scaling=1.0, return_variance=True)
a = cm.OrdinaryKrigingReconstructor(src, tgt, **kw)
one = a.reconstruct().da.values
b = cm.OrdinaryKrigingReconstructor(src, tgt, group=dist.group.WORLD, **kw)
many = b.reconstruct().da.values
# The pair-sharded variogram adds the ranks' partial sums in a different order than one GPU does: the semivariances
# differ in the last bit, and the fit (finite-difference Jacobian, stops at ftol = 1e-8) moves the parameters by up to
# ~1e-7 relative.  Two ranks happen to reproduce the single-rank sums bit for bit; eight do not.
perr = np.max(np.abs(b.variogram_model_parameters - a.variogram_model_parameters) / np.maximum(np.abs(a.variogram_model_parameters), 1e-6))
err = np.max(np.abs(many - one)) / np.max(np.abs(one))
errv = np.max(np.abs(b.variance_ - a.variance_)) / np.max(np.abs(a.variance_))
print(f"rank {rank}: parameters differ by {perr:.2e}, field {err:.2e}, variance {errv:.2e}", flush=True)
assert perr <= 1e-5 and err <= 1e-6 and errv <= 1e-6, (perr, err, errv)
print(f"rank {rank}/{world}: distributed OK matches the single-rank field (max rel diff {err:.2e}, variance {errv:.2e})")
dist.destroy_process_group()
