"""Small end-to-end run of every kernel for compute-sanitizer (memcheck / racecheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from climatrix_b200 import kernels as K
rng = np.random.default_rng(0)
dev = "cuda:0"
for (N, nlat, nlon, k) in [(700, 9, 40, 5), (3000, 20, 70, 33), (1000, 5, 33, 100)]:
    lat, lon, v = rng.uniform(-90, 90, N), rng.uniform(-180, 180, N), rng.normal(size=N)
    T = K.Targets.make(np.linspace(-80, 80, nlat), np.linspace(-170, 170, nlon), True, dev)
    K.idw(lat, lon, v, T, k=k)
    K.idw(lat, lon, np.stack([v, v * 2, v + 1], 1), T, k=k, layout="sample_major")
    K.knn(lat, lon, K.Targets.make(rng.uniform(-90, 90, 300), rng.uniform(-180, 180, 300), False, dev), k)
x, y, z = rng.uniform(-1, 1, 600), rng.uniform(-1, 1, 600), rng.normal(size=600)
K.empirical_variogram(x, y, z, 6)
K.empirical_variogram(x, y, z, 6, geographic=True)
for k in (8, 40, 70):
    K.ok_moving_window(x, y, z, np.linspace(-1, 1, 12), np.linspace(-1, 1, 9), True, k=k, variogram_model="spherical", params=[1.0, 0.5, 0.05])
s = rng.normal(size=(500, 3)); s /= np.linalg.norm(s, axis=1, keepdims=True)
d2, idx = K.knn3(s, s[:100], 10)
K.ok_solve(x[:500], y[:500], z[:500], idx, d2, variogram_model="linear", params=[1.0, 0.1], geographic=True)
torch.cuda.synchronize()
print("sanitize_small done")
