"""Kernel-level timing of K3: python tools/prof_variogram.py N nlags"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from climatrix_b200 import kernels as K
N, nlags = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(3)
dev = torch.device("cuda:0")
x = torch.as_tensor(rng.uniform(-1, 1, N)).to(dev); y = torch.as_tensor(rng.uniform(-1, 1, N)).to(dev)
z = torch.as_tensor(rng.normal(size=N)).to(dev)
K.empirical_variogram(x, y, z, nlags); torch.cuda.synchronize()
for _ in range(3):
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); lags, sv, raw = K.empirical_variogram(x, y, z, nlags); e1.record(); torch.cuda.synchronize()
import hashlib
h = hashlib.sha1(b"".join(np.ascontiguousarray(raw[key]).tobytes() for key in ("sum_d", "sum_g", "count"))).hexdigest()[:12]
print("lib", os.environ.get("CMX_LIB", "default"), "sums sha1", h)
print(f"N={N} nlags={nlags}: empirical_variogram {e0.elapsed_time(e1):.2f} ms ({N*(N-1)/2/e0.elapsed_time(e1)/1e6:.1f} Gpairs/s)", lags[:3], raw["count"][:3])
