"""Phase breakdown of an ordinary-kriging step (C2-like): python tools/prof_c2.py N n_lat n_lon k"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from climatrix_b200 import kernels as K
from climatrix_b200.kriging import fit_variogram
N, nlat, nlon, k = (int(a) for a in sys.argv[1:5])
rng = np.random.default_rng(3)
dev = torch.device("cuda:0")
x = torch.as_tensor(rng.uniform(-1, 1, N)).to(dev); y = torch.as_tensor(rng.uniform(-1, 1, N)).to(dev)
zz = np.sin(3 * x.cpu().numpy()) + 0.3 * rng.normal(size=N)
z = torch.as_tensor(zz).to(dev)
tx = torch.as_tensor(np.linspace(-1, 1, nlon)).to(dev); ty = torch.as_tensor(np.linspace(-1, 1, nlat)).to(dev)
def sync(): torch.cuda.synchronize()
for rep in range(3):
    sync(); t0 = time.perf_counter()
    lags, sv, raw = K.empirical_variogram(x, y, z, 6); sync(); t1 = time.perf_counter()
    params = fit_variogram(lags, sv, "spherical"); t2 = time.perf_counter()
    out, ns = K.ok_moving_window(x, y, z, tx, ty, True, k=k, variogram_model="spherical", params=params); sync(); t3 = time.perf_counter()
print(f"N={N} M={nlat*nlon} k={k}: variogram {1e3*(t1-t0):.2f} ms, host fit {1e3*(t2-t1):.2f} ms, knn+solve {1e3*(t3-t2):.2f} ms")
