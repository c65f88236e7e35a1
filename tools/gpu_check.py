"""First-contact GPU check: kernels vs the CPU oracle on seeded inputs, with rough timings."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from climatrix_b200 import kernels as K
from oracle.idw_oracle import knn_ckdtree, knn_bruteforce, idw_oracle, idw_oracle_batched, dense_points, sparse_points
from oracle import pykrige_restated as pk

dev = torch.device("cuda:0")
print(torch.cuda.get_device_name(0), flush=True)

def field(lat, lon, rng, phase=0.0):
    return 15 + 10*np.cos(np.deg2rad(lat)) + 5*np.sin(2*np.deg2rad(lon)+phase) + rng.normal(size=lat.shape)

def timed(fn, n=3):
    torch.cuda.synchronize(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n, r

ok_all = True
def report(name, ok, extra=""):
    global ok_all
    ok_all &= bool(ok)
    print(f"[{'PASS' if ok else 'FAIL'}] {name} {extra}", flush=True)

# ---- kNN parity, several shapes ----
for (N, nlat, nlon, k, seed) in [(1000, 180, 360, 30, 1), (5000, 45, 90, 32, 2), (300, 17, 35, 5, 3), (3000, 60, 70, 64, 4), (700, 33, 31, 100, 5), (40, 9, 18, 1, 6), (5000, 7, 9, 8, 7)]:
    rng = np.random.default_rng(seed)
    lat = rng.uniform(-90, 90, N); lon = rng.uniform(-180, 180, N)
    tl = np.linspace(-89.5, 89.5, nlat); to = np.linspace(-179.5, 179.5, nlon)
    T = K.Targets.make(tl, to, True, dev)
    ms, (dist, idx) = timed(lambda: K.knn(lat, lon, T, k))
    d0, i0 = knn_ckdtree(sparse_points(lat, lon), dense_points(tl, to), k)
    same_i = np.array_equal(idx.cpu().numpy().astype(np.int64), i0)
    same_d = np.array_equal(dist.cpu().numpy(), d0)
    report(f"knn grid N={N} M={nlat*nlon} k={k}", same_i and same_d, f"idx={same_i} dist_bitexact={same_d} {ms:.3f} ms")
    if not same_i:
        bad = np.argwhere(idx.cpu().numpy() != i0)
        print("  first mismatches:", bad[:5], idx.cpu().numpy()[bad[0][0]][:8], i0[bad[0][0]][:8])

# sparse targets
rng = np.random.default_rng(11)
N, M, k = 4000, 5000, 16
lat = rng.uniform(-90, 90, N); lon = rng.uniform(-180, 180, N)
tl = rng.uniform(-90, 90, M); to = rng.uniform(-180, 180, M)
T = K.Targets.make(tl, to, False, dev)
dist, idx = K.knn(lat, lon, T, k)
d0, i0 = knn_ckdtree(sparse_points(lat, lon), sparse_points(tl, to), k)
report("knn points", np.array_equal(idx.cpu().numpy(), i0) and np.array_equal(dist.cpu().numpy(), d0))

# clustered (small distances)
rng = np.random.default_rng(12)
lat = 50 + rng.normal(scale=0.05, size=20000); lon = 10 + rng.normal(scale=0.05, size=20000)
tl = np.linspace(49.8, 50.2, 101); to = np.linspace(9.8, 10.2, 101)
T = K.Targets.make(tl, to, True, dev)
dist, idx = K.knn(lat, lon, T, 32)
d0, i0 = knn_ckdtree(sparse_points(lat, lon), dense_points(tl, to), 32)
report("knn clustered", np.array_equal(idx.cpu().numpy(), i0) and np.array_equal(dist.cpu().numpy(), d0))

# lattice ties: compare with the (d2, idx) brute-force rule
gl, go = np.meshgrid(np.arange(-20., 21, 1), np.arange(-20., 21, 1), indexing="ij")
lat, lon = gl.ravel(), go.ravel()
tl = np.arange(-10., 11, 0.5); to = np.arange(-10., 11, 0.5)
T = K.Targets.make(tl, to, True, dev)
dist, idx = K.knn(lat, lon, T, 9)
d1, i1 = knn_bruteforce(sparse_points(lat, lon), dense_points(tl, to), 9)
report("knn lattice ties (d2, idx) rule", np.array_equal(idx.cpu().numpy(), i1) and np.array_equal(dist.cpu().numpy(), d1))

# ---- fused IDW, C1 ----
rng = np.random.default_rng(1234)
N = 1000
lat = rng.uniform(-90, 90, N); lon = rng.uniform(-180, 180, N); v = field(lat, lon, rng)
tl = np.arange(-89.5, 90, 1.0); to = np.arange(-179.5, 180, 1.0)
T = K.Targets.make(tl, to, True, dev)
ms, out = timed(lambda: K.idw(lat, lon, v, T, k=30, power=2.0, k_min=2))
ref = idw_oracle(sparse_points(lat, lon), v, dense_points(tl, to), k=30, power=2.0, k_min=2)
err = np.nanmax(np.abs(out.cpu().numpy() - ref) / np.maximum(np.abs(ref), 1e-12))
report("idw fused C1 f64", err < 1e-5, f"max rel err {err:.3e}  {ms:.3f} ms -> {64800/ms*1e3:.3e} targets/s")
out32 = K.idw(lat, lon, v.astype(np.float32), T, k=30, power=2.0, k_min=2, dtype=torch.float32)
ref32 = idw_oracle(sparse_points(lat, lon), v.astype(np.float32).astype(np.float64), dense_points(tl, to), k=30)
err = np.nanmax(np.abs(out32.cpu().numpy() - ref32) / np.maximum(np.abs(ref32), 1e-12))
report("idw fused C1 f32", err < 1e-5, f"max rel err {err:.3e}")
# NaNs + other powers
vb = v.copy(); bad = rng.choice(N, 200, replace=False); vb[bad[:150]] = np.nan; vb[bad[150:]] = np.inf
for p, kmin in [(2.0, 1), (0.5, 4), (5.0, 8), (1e-10, 2)]:
    out = K.idw(lat, lon, vb, T, k=8, power=p, k_min=kmin).cpu().numpy()
    with np.errstate(invalid="ignore"):
        ref = idw_oracle(sparse_points(lat, lon), vb, dense_points(tl, to), k=8, power=p, k_min=kmin)
    nan_same = np.array_equal(np.isnan(out), np.isnan(ref))
    err = np.nanmax(np.abs(out - ref) / np.maximum(np.abs(ref), 1e-12))
    report(f"idw nonfinite p={p} kmin={kmin}", nan_same and err < 1e-5, f"nan_same={nan_same} err={err:.2e} nans={np.isnan(ref).sum()}")

# ---- batched IDW ----
Tn = 37
vt = np.stack([field(lat, lon, rng, phase=0.1*t) for t in range(Tn)])
vt[3, bad[:50]] = np.nan
for layout, arr in (("batch_major", vt), ("sample_major", np.ascontiguousarray(vt.T))):
    ms, out = timed(lambda: K.idw(lat, lon, arr, T, k=30, power=2.0, k_min=2, layout=layout))
    with np.errstate(invalid="ignore"):
        ref = idw_oracle_batched(sparse_points(lat, lon), vt, dense_points(tl, to), k=30)
    o = out.cpu().numpy()
    nan_same = np.array_equal(np.isnan(o), np.isnan(ref))
    err = np.nanmax(np.abs(o - ref) / np.maximum(np.abs(ref), 1e-12))
    report(f"idw batched T={Tn} {layout}", nan_same and err < 1e-5, f"err={err:.2e} {ms:.3f} ms")

# ---- apply from table ----
o, dist, idx = K.idw(lat, lon, v, T, k=30, return_neighbours=True)
o2 = K.idw_apply(idx, dist, v, k=12, power=3.0, k_min=2)
ref = idw_oracle(sparse_points(lat, lon), v, dense_points(tl, to), k=12, power=3.0, k_min=2)
err = np.nanmax(np.abs(o2.cpu().numpy() - ref) / np.maximum(np.abs(ref), 1e-12))
report("idw_apply k=12 p=3 from k=30 table", err < 1e-5, f"err={err:.2e}")

# ---- variogram ----
rng = np.random.default_rng(99)
N = 3000
x = rng.uniform(-1, 1, N); y = rng.uniform(-1, 1, N); z = rng.normal(size=N)
ms, (lags, sv, raw) = timed(lambda: K.empirical_variogram(x, y, z, 6), n=1)
l0, s0, r0 = pk.empirical_variogram(np.stack([x, y], 1), z, 6, "euclidean")
report("variogram euclid", raw["dmin"] == r0["dmin"] and raw["dmax"] == r0["dmax"] and np.array_equal(raw["count"], r0["count"]) and np.allclose(lags, l0, rtol=1e-12) and np.allclose(sv, s0, rtol=1e-12),
       f"dmin {raw['dmin']==r0['dmin']} dmax {raw['dmax']==r0['dmax']} count {np.array_equal(raw['count'], r0['count'])} {ms:.2f} ms")
lags, sv, raw = K.empirical_variogram(x[:800], y[:800], z[:800], 8, geographic=True)
l0, s0, r0 = pk.empirical_variogram(np.stack([x[:800], y[:800]], 1), z[:800], 8, "geographic")
report("variogram geographic", np.array_equal(raw["count"], r0["count"]) and np.allclose(lags, l0, rtol=1e-9) and np.allclose(sv, s0, rtol=1e-9), f"count {raw['count']} vs {r0['count']}")

# ---- moving-window kriging ----
for model, k in [("spherical", 32), ("linear", 16), ("power", 12), ("gaussian", 8), ("exponential", 64)]:
    rng = np.random.default_rng(5 + k)
    N = 1500
    x = rng.uniform(-1, 1, N); y = rng.uniform(-1, 1, N)
    z = np.sin(3*x) + np.cos(2*y) + 0.1*rng.normal(size=N)
    okr = pk.OrdinaryKriging(x, y, z, variogram_model=model, nlags=6, anisotropy_scaling=1.0)
    tx = np.linspace(-1, 1, 40); ty = np.linspace(-1, 1, 30)
    zr, sr = okr.execute("grid", tx, ty, backend="loop", n_closest_points=k)
    ms, (out, sig, nsing) = timed(lambda: K.ok_moving_window(okr.X_ADJUSTED, okr.Y_ADJUSTED, okr.Z, tx, ty, True, k=k, variogram_model=model, params=okr.variogram_model_parameters, return_sigmasq=True), n=1)
    o = out.cpu().numpy().reshape(30, 40)
    err = np.max(np.abs(o - zr)) / np.max(np.abs(zr))
    errs = np.max(np.abs(sig.cpu().numpy().reshape(30, 40) - sr)) / max(np.max(np.abs(sr)), 1e-30)
    report(f"ok_mw {model} k={k}", err < 1e-4, f"rel err {err:.2e} sigmasq err {errs:.2e} nsing={int(nsing)} {ms:.2f} ms")

# ---- timing at a mid-size config ----
rng = np.random.default_rng(7)
N = 100000
lat = rng.uniform(-90, 90, N); lon = rng.uniform(0, 360, N); v = field(lat, lon, rng)
tl = np.linspace(90, -90, 721); to = np.arange(0, 360, 0.25)
T = K.Targets.make(tl, to, True, dev)
slat = torch.as_tensor(lat).to(dev); slon = torch.as_tensor(lon).to(dev); sv = torch.as_tensor(v).to(dev)
ms, out = timed(lambda: K.idw(slat, slon, sv, T, k=32), n=2)
M = 721*1440
print(f"C3 single slice: {ms:.2f} ms  {M/ms*1e3:.3e} targets/s  {5.0*N*M/ms/1e9:.2f} TFLOP/s (algorithmic 5N/target)", flush=True)
# parity on a sample of rows
rows = np.array([0, 100, 360, 719, 720])
ref = idw_oracle(sparse_points(lat, lon), v, dense_points(tl[rows], to), k=32)
o = out.cpu().numpy().reshape(721, 1440)[rows].ravel()
err = np.nanmax(np.abs(o - ref) / np.maximum(np.abs(ref), 1e-12))
report("idw C3 sample rows", err < 1e-5, f"err={err:.2e}")
print("ALL PASS" if ok_all else "SOME FAILED", flush=True)
