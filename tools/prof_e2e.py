"""Timeline of the host path at C5 (two-band pipeline): python tools/prof_e2e.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from climatrix_b200 import kernels as K, _lib
from climatrix_b200.sharding import _idw_two_bands, idw_on_devices
rng = np.random.default_rng(1)
N, n_lat, n_lon, T, k = 1_000_000, 1801, 3600, 37, 32
lat, lon = rng.uniform(-90, 90, N), rng.uniform(0, 360, N)
vals = rng.normal(size=(N, T))
t_lat, t_lon = np.linspace(90, -90, n_lat), np.linspace(0, 359.9, n_lon)
dev = torch.device("cuda:0")
def full():
    return idw_on_devices(lat, lon, vals, t_lat, t_lon, True, k=k, power=2.0, k_min=2, dtype=torch.float64, layout="sample_major")
for _ in range(3): r = full()
torch.cuda.synchronize()
t0 = time.perf_counter(); r = full(); torch.cuda.synchronize(); print(f"full host path: {(time.perf_counter()-t0)*1e3:.1f} ms")
# pieces
def stamp(label, t0):
    torch.cuda.synchronize(); t = time.perf_counter(); print(f"  {label:38s} {1e3*(t-t0):8.1f} ms"); return time.perf_counter()
for rep in range(2):
    print("pieces (serialised with synchronize):")
    t = time.perf_counter()
    host = torch.empty((T, n_lat * n_lon), dtype=torch.float64, pin_memory=True); t = stamp("pinned alloc", t)
    s_lat_d, s_lon_d = K._as_f64(lat, dev), K._as_f64(lon, dev); t_lat_d, t_lon_d = K._as_f64(t_lat, dev), K._as_f64(t_lon, dev); t = stamp("coords H2D", t)
    split = (int(0.92 * n_lat) // 16) * 16
    tgA = K.Targets(t_lat_d[:split].contiguous(), t_lon_d, True)
    dA, iA = K.knn(s_lat_d, s_lon_d, tgA, k); t = stamp(f"knn band A ({split} rows)", t)
    vd = torch.as_tensor(vals).to(dev); t = stamp("values H2D (pageable)", t)
    bA = K.idw_apply(iA, dA, vd, k=k, layout="sample_major", n_samples=N); t = stamp("apply A", t)
    lib = _lib.load()
    mA = split * n_lon
    lib.cmx_copy_rows_to_host(host.data_ptr(), n_lat * n_lon * 8, bA.data_ptr(), mA * 8, mA * 8, T, torch.cuda.current_stream().cuda_stream); t = stamp("D2H band A (strided)", t)
    tgB = K.Targets(t_lat_d[split:].contiguous(), t_lon_d, True)
    dB, iB = K.knn(s_lat_d, s_lon_d, tgB, k); t = stamp(f"knn band B ({n_lat-split} rows)", t)
    bB = K.idw_apply(iB, dB, vd, k=k, layout="sample_major", n_samples=N); t = stamp("apply B", t)
    mB = (n_lat - split) * n_lon
    lib.cmx_copy_rows_to_host(host.data_ptr() + mA * 8, n_lat * n_lon * 8, bB.data_ptr(), mB * 8, mB * 8, T, torch.cuda.current_stream().cuda_stream); t = stamp("D2H band B (strided)", t)
    tg = K.Targets(t_lat_d, t_lon_d, True)
    d, i = K.knn(s_lat_d, s_lon_d, tg, k); t = stamp("knn whole grid (for comparison)", t)
    del host
