"""One process, several GPUs: IDWReconstructor(devices=[...]) / OrdinaryKrigingReconstructor(devices=[...]) against the
single-device result (must be bit-identical: the bands are independent).  python tools/check_multi_device.py [n_devices]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import climatrix_b200 as cm

nd = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
devices = [f"cuda:{i}" for i in range(nd)]
rng = np.random.default_rng(0)
n, t = 200_000, 12
lat, lon = rng.uniform(-90, 90, n), rng.uniform(-180, 180, n)
vals = 15 + 10 * np.cos(np.deg2rad(lat))[:, None] + rng.normal(size=(n, t))
coords = {"point": np.arange(n), "level": np.arange(t, dtype=float), "latitude": (("point",), lat), "longitude": (("point",), lon)}
src = cm.BaseClimatrixDataset(cm.FieldArray(vals, ("point", "level"), coords, name="f"))
tgt = cm.Domain.from_lat_lon(np.linspace(-89, 89, 713), np.linspace(-179, 179, 1441), kind="dense")
from climatrix_b200 import kernels as K

for search in ("binned", "brute"):  # the default search (search + weighting in one call) and the split brute-force path
    K.set_search(search)
    one = src.reconstruct(tgt, method="idw", k=32).da.values
    for _ in range(3):
        t0 = time.perf_counter()
        many = cm.IDWReconstructor(src, tgt, k=32, devices=devices).reconstruct().da.values
        dt = time.perf_counter() - t0
    assert one.shape == many.shape and np.array_equal(one, many), f"IDW ({search}): devices= result differs from the single-device one"
    t0 = time.perf_counter()
    src.reconstruct(tgt, method="idw", k=32)
    dt1 = time.perf_counter() - t0
    print(f"IDW batch {t} x {713 * 1441}, {search} search, on {nd} devices: identical; {dt * 1e3:.1f} ms vs {dt1 * 1e3:.1f} ms on one")
    # batch-major values (T, N) take the same path
    srcb = cm.BaseClimatrixDataset(cm.FieldArray(np.ascontiguousarray(vals.T), ("level", "point"), coords, name="f"))
    manyb = cm.IDWReconstructor(srcb, tgt, k=32, devices=devices).reconstruct().da.values
    assert np.array_equal(one, manyb), f"IDW ({search}), batch-major values: devices= result differs"
# a tall band (>= 512 rows per device) of the brute-force path is cut in two: first part downloads beside the second search
K.set_search("brute")
tall = cm.Domain.from_lat_lon(np.linspace(-89, 89, 600 * nd + 7), np.linspace(-179, 179, 301), kind="dense")
one = src.reconstruct(tall, method="idw", k=32).da.values
many = cm.IDWReconstructor(src, tall, k=32, devices=devices).reconstruct().da.values
assert np.array_equal(one, many), "IDW (brute, tall bands): devices= result differs"
print(f"IDW batch on a {600 * nd + 7}-row grid (two-part bands), brute search: identical")
K.set_search("auto")

src1 = cm.BaseClimatrixDataset(cm.FieldArray(vals[:, 0], ("point",), {k: v for k, v in coords.items() if k != "level"}, name="f"))
one = src1.reconstruct(tgt, method="idw", k=8).da.values
many = cm.IDWReconstructor(src1, tgt, k=8, devices=devices).reconstruct().da.values
assert np.array_equal(one, many), "IDW single slice: devices= result differs"
print("IDW single slice: identical")

m = 20_000
srck = cm.BaseClimatrixDataset(cm.FieldArray(vals[:m, 0], ("point",), {"point": np.arange(m), "latitude": (("point",), lat[:m]),
                                                                    "longitude": (("point",), lon[:m])}, name="f"))
tg2 = cm.Domain.from_lat_lon(np.linspace(-89, 89, 179), np.linspace(-179, 179, 359), kind="dense")
kw = dict(n_closest_points=24, variogram_model="spherical", anisotropy_scaling=1.0, return_variance=True)
r1 = cm.OrdinaryKrigingReconstructor(srck, tg2, **kw)
one = r1.reconstruct().da.values
r2 = cm.OrdinaryKrigingReconstructor(srck, tg2, devices=devices, **kw)
many = r2.reconstruct().da.values
assert np.array_equal(one, many), "OK: devices= result differs"
assert np.array_equal(r1.variance_, r2.variance_), "OK variance differs"
print(f"OK moving window on {nd} devices: field and variance identical")
