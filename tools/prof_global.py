"""Timing of the global kriging path: python tools/prof_global.py N n_lat n_lon [pinv]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from climatrix_b200 import kernels as K
N, nlat, nlon = (int(a) for a in sys.argv[1:4])
pinv = len(sys.argv) > 4
rng = np.random.default_rng(3)
dev = torch.device("cuda:0")
x = torch.as_tensor(rng.uniform(-1, 1, N)).to(dev); y = torch.as_tensor(rng.uniform(-1, 1, N)).to(dev)
z = torch.as_tensor(np.sin(3 * x.cpu().numpy()) + 0.3 * rng.normal(size=N)).to(dev)
tx = np.linspace(-1, 1, nlon); ty = np.linspace(-1, 1, nlat)
def run(): return K.ok_global(x, y, z, tx, ty, True, variogram_model="spherical", params=[1.0, 0.5, 0.05], pseudo_inv=pinv)
run(); torch.cuda.synchronize()
t0 = time.perf_counter(); out, st = run(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
flops = N ** 3 / 3
print(f"N={N} M={nlat*nlon} pinv={pinv}: {dt*1e3:.1f} ms status {st} (Cholesky N^3/3 = {flops/1e9:.1f} GFLOP -> {flops/dt/1e12:.2f} TFLOP/s if it were all factorisation)")
