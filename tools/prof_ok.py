"""Kernel-level timing of the kriging path: python tools/prof_ok.py N n_lat n_lon k"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from climatrix_b200 import kernels as K
N, nlat, nlon, k = (int(a) for a in sys.argv[1:5])
rng = np.random.default_rng(3)
dev = torch.device("cuda:0")
x = torch.as_tensor(rng.uniform(-1, 1, N)).to(dev); y = torch.as_tensor(rng.uniform(-1, 1, N)).to(dev)
z = torch.as_tensor(rng.normal(size=N)).to(dev)
tx = np.linspace(-1, 1, nlon); ty = np.linspace(-1, 1, nlat)
def run():
    return K.ok_moving_window(x, y, z, tx, ty, True, k=k, variogram_model="spherical", params=[1.0, 0.5, 0.05])
run(); torch.cuda.synchronize()
K.kernel_timing(True)
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
knn_ms, _ = K.kernel_time(K.KERNEL_KNN)
K.kernel_timing(True); run(); torch.cuda.synchronize(); ok_ms, _ = K.kernel_time(K.KERNEL_OK_SOLVE); K.kernel_timing(False)
e2, e3 = torch.cuda.Event(True), torch.cuda.Event(True)
e2.record(); K.empirical_variogram(x, y, z, 6); e3.record(); torch.cuda.synchronize()
print(f"N={N} M={nlat*nlon} k={k}: total {e0.elapsed_time(e1):.2f} ms, knn {knn_ms:.2f} ms, solve {ok_ms:.2f} ms ({nlat*nlon/ok_ms*1e3:.3e} targets/s), variogram {e2.elapsed_time(e3):.2f} ms")
