"""Sample-split path check on one GPU: a thin band of a big job (what one rank of an 8-GPU run sees)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from climatrix_b200 import kernels as K
from oracle.idw_oracle import idw_oracle, dense_points, sparse_points
rng = np.random.default_rng(3)
N = 1_000_000
lat = rng.uniform(-90, 90, N); lon = rng.uniform(-180, 180, N); v = rng.normal(size=N)
dev = torch.device("cuda:0")
slat, slon, sv = (torch.as_tensor(a).to(dev) for a in (lat, lon, v))
for rows in (226, 450, 900):
    tl = np.linspace(-90, 90, 1801)[:rows]; to = -180 + np.arange(3600) * 0.1
    T = K.Targets.make(tl, to, True, dev)
    out, dist, idx = K.idw(slat, slon, sv, T, k=32, return_neighbours=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); K.idw(slat, slon, sv, T, k=32); e1.record(); torch.cuda.synchronize()
    sel = np.array([0, 1, rows // 2, rows - 1])
    ref, d0, i0 = idw_oracle(sparse_points(lat, lon), v, dense_points(tl[sel], to), k=32, return_neighbours=True)
    pick = (sel[:, None] * 3600 + np.arange(3600)[None, :]).ravel()
    ok = np.array_equal(idx.cpu().numpy()[pick], i0) and np.array_equal(dist.cpu().numpy()[pick], d0)
    err = np.max(np.abs(out.cpu().numpy()[pick] - ref) / np.maximum(np.abs(ref), 1e-12))
    ms = e0.elapsed_time(e1)
    print(f"rows={rows}: {ms:.1f} ms  {5.0*N*rows*3600/ms/1e9:.1f} TFLOP/s  neighbours bit-exact={ok}  idw err={err:.1e}")
# batched + kriging through the split too
T = K.Targets.make(np.linspace(-90, 90, 1801)[:226], -180 + np.arange(3600) * 0.1, True, dev)
vt = torch.as_tensor(rng.normal(size=(N, 5))).to(dev)
ob = K.idw(slat, slon, vt, T, k=32, layout="sample_major")
o1 = K.idw(slat, slon, vt[:, 2].contiguous(), T, k=32)
print("batched vs single slice through split:", float((ob[2] - o1).abs().max()))
