"""Host path of ONE rank's band of an 8-GPU C5 job (225 rows), on one GPU: python tools/prof_band_e2e.py [rows]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import climatrix_b200 as cm
from climatrix_b200 import kernels as K
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 225
search = sys.argv[2] if len(sys.argv) > 2 else "brute"
K.set_search(search)
rng = np.random.default_rng(1)
N, n_lat, n_lon, T, k = 1_000_000, 1801, 3600, 37, 32
lat, lon = rng.uniform(-90, 90, N), rng.uniform(0, 360, N)
vals = rng.normal(size=(N, T))
t_lat, t_lon = np.linspace(90, -90, n_lat)[:rows], np.linspace(0, 359.9, n_lon)
coords = {"point": np.arange(N), "latitude": (("point",), lat), "longitude": (("point",), lon), "level": np.arange(T, dtype=np.float64)}
src = cm.BaseClimatrixDataset(cm.FieldArray(vals, ("point", "level"), coords, name="field"))
tgt = cm.Domain.from_lat_lon(t_lat, t_lon, kind="dense")
print("torch threads", torch.get_num_threads(), "search", search)
for i in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = src.reconstruct(tgt, method="idw", k=k, power=2.0, k_min=2)
    torch.cuda.synchronize(); print(f"reconstruct band of {rows} rows: {(time.perf_counter() - t0) * 1e3:.1f} ms")
