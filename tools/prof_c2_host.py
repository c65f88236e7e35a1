"""Where the host time of one C2 ordinary-kriging step goes: python tools/prof_c2_host.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from climatrix_b200 import kernels as K
from climatrix_b200.kriging import _minmax_scale
rng = np.random.default_rng(2)
n, n_lat, n_lon, k = 10_000, 180, 360, 32
lat, lon = rng.uniform(-90, 90, n), rng.uniform(-180, 180, n)
vals = 15 + 10 * np.cos(np.deg2rad(lat)) + rng.normal(size=n)
t_lat, t_lon = np.linspace(-89.5, 89.5, n_lat), np.linspace(-179.5, 179.5, n_lon)
dev = torch.device("cuda:0")
vals_d = torch.as_tensor(vals).to(dev)
marks = {}
def step(record=False):
    t = [time.perf_counter()]
    def mark(name):
        if record:
            torch.cuda.synchronize()
        t.append(time.perf_counter()); marks.setdefault(name, []).append(t[-1] - t[-2])
    lat_n = torch.as_tensor(_minmax_scale(lat)).to(dev); lon_n = torch.as_tensor(_minmax_scale(lon)).to(dev); mark("normalise + upload coords")
    mean, std = float(np.nanmean(vals)), float(np.nanstd(vals)); z = (vals_d - mean) / std; mark("standardise")
    lags, sv, _ = K.empirical_variogram(lon_n, lat_n, z, 6); mark("K3 + D2H")
    params = K.fit_variogram_native(lags, sv, "spherical"); mark("fit")
    tl = torch.as_tensor(_minmax_scale(t_lat)).to(dev); to = torch.as_tensor(_minmax_scale(t_lon)).to(dev); mark("target axes")
    out, nsing = K.ok_moving_window(lon_n, lat_n, z, to, tl, True, k=k, variogram_model="spherical", params=params, out_scale=std, out_shift=mean); mark("K1c + K4 launch")
    n_s = int(nsing.item()); mark("sync (singular count)")
    return out
for _ in range(5): step()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): step()
torch.cuda.synchronize(); print(f"step: {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms")
marks.clear()
for _ in range(20): step(True)
for name, v in marks.items(): print(f"  {name:28s} {np.mean(v) * 1e3:7.3f} ms (serialised)")
