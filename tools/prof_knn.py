"""One IDW call at a chosen size, for ncu / timing.  python tools/prof_knn.py N n_lat n_lon k [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from climatrix_b200 import kernels as K
K.set_search(os.environ.get('CMX_SEARCH', 'brute'))
N, nlat, nlon, k = (int(a) for a in sys.argv[1:5])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 1
rng = np.random.default_rng(7)
dev = torch.device("cuda:0")
lat = torch.as_tensor(rng.uniform(-90, 90, N)).to(dev); lon = torch.as_tensor(rng.uniform(0, 360, N)).to(dev)
v = torch.as_tensor(rng.normal(size=N)).to(dev)
T = K.Targets.make(np.linspace(90, -90, nlat), np.linspace(0, 360, nlon, endpoint=False), True, dev)
K.idw(lat, lon, v, T, k=k); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(reps): K.idw(lat, lon, v, T, k=k)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
M = nlat * nlon
print(f"N={N} M={M} k={k}: {ms:.3f} ms  {M/ms*1e3:.3e} targets/s  {5.0*N*M/ms/1e9:.2f} TFLOP/s")
