"""Kernel-level timing of the batched IDW path: python tools/prof_batch.py N n_lat n_lon T k"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from climatrix_b200 import kernels as K
N, nlat, nlon, T, k = (int(a) for a in sys.argv[1:6])
rng = np.random.default_rng(7)
dev = torch.device("cuda:0")
lat = torch.as_tensor(rng.uniform(-90, 90, N)).to(dev); lon = torch.as_tensor(rng.uniform(0, 360, N)).to(dev)
v = torch.as_tensor(rng.normal(size=(N, T))).to(dev)
Tg = K.Targets.make(np.linspace(90, -90, nlat), np.linspace(0, 360, nlon, endpoint=False), True, dev)
def run(): return K.idw(lat, lon, v, Tg, k=k, layout="sample_major")
run(); torch.cuda.synchronize()
K.kernel_timing(True)
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record(); out = run(); e1.record(); torch.cuda.synchronize()
knn_ms = K.kernel_time(K.KERNEL_KNN)[0] + K.kernel_time(K.KERNEL_KNN_BINNED)[0]
K.kernel_timing(True); run(); torch.cuda.synchronize(); b_ms, _ = K.kernel_time(K.KERNEL_IDW_BATCH); K.kernel_timing(False)
M = nlat * nlon
print(f"N={N} M={M} T={T} k={k}: total {e0.elapsed_time(e1):.2f} ms, search(+fused epilogue) {knn_ms:.2f} ms, K2 {b_ms:.2f} ms")
