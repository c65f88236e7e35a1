"""Compare two builds of K4 (CMX_LIB=... selects the library): writes / compares the kriging fields of a few windows.
python tools/check_ok_variants.py save|compare file.npz"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from climatrix_b200 import kernels as K
mode, path = sys.argv[1], sys.argv[2]
rng = np.random.default_rng(5)
dev = torch.device("cuda:0")
out = {}
for N, k, model, params in [(3000, 64, "spherical", [1.0, 0.5, 0.05]), (3000, 40, "exponential", [0.8, 0.7, 0.01]),
                            (2000, 33, "gaussian", [1.0, 0.4, 0.1]), (2000, 32, "spherical", [1.0, 0.5, 0.0]),
                            (2000, 17, "linear", [0.7, 0.02, 0.0]), (500, 9, "power", [0.9, 1.3, 0.05]),
                            (500, 2, "spherical", [1.0, 0.5, 0.05]), (500, 1, "spherical", [1.0, 0.5, 0.05]),
                            (3000, 48, "spherical", [1.0, 0.5, 0.05]), (3000, 16, "spherical", [1.0, 0.5, 0.05])]:
    x, y, z = rng.uniform(-1, 1, N), rng.uniform(-1, 1, N), rng.normal(size=N)
    x[10] = x[11]; y[10] = y[11]  # one duplicate pair
    tx, ty = np.linspace(-1, 1, 41), np.linspace(-1, 1, 37)
    for dt in (torch.float64, torch.float32):
        res, sig, nsing = K.ok_moving_window(torch.as_tensor(x).to(dev), torch.as_tensor(y).to(dev), torch.as_tensor(z).to(dev),
                                             tx, ty, True, k=k, variogram_model=model, params=params, dtype=dt, return_sigmasq=True)
        out[f"{N}_{k}_{model}_{dt}".replace("torch.", "")] = res.double().cpu().numpy()
        if sig is not None and torch.is_tensor(sig):
            out[f"{N}_{k}_{model}_{dt}_sig".replace("torch.", "")] = sig.double().cpu().numpy()
if mode == "save":
    np.savez(path, **out)
    print("saved", len(out))
else:
    ref = np.load(path)
    worst = 0.0
    for key, v in out.items():
        r = ref[key]
        nan_equal = np.array_equal(np.isnan(r), np.isnan(v))
        err = np.nanmax(np.abs(r - v) / (np.abs(r) + 1e-12)) if np.isfinite(r).any() else 0.0
        worst = max(worst, err)
        print(f"{key}: nan-pattern {'same' if nan_equal else 'DIFFERENT'}, max rel diff {err:.3e}, nans {int(np.isnan(v).sum())}")
    print("worst", worst)
