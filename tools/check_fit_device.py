"""Device fit (variogram_fit_kernel) vs host fit (cmx_variogram_fit_host) vs SciPy, on random variograms."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from climatrix_b200 import kernels as K
from climatrix_b200.kriging import fit_variogram
warnings.simplefilter("ignore")
for model in ["linear", "power", "gaussian", "spherical", "exponential"]:
    for nlags, geo in [(6, False), (12, False), (6, True)]:
        rng = np.random.default_rng(nlags + len(model))
        n = 2500
        x, y = (rng.uniform(-170, 170, n), rng.uniform(-80, 80, n)) if geo else (rng.uniform(-1, 1, n), rng.uniform(-1, 1, n))
        z = np.sin(0.03 * x if geo else 3 * x) + np.cos(0.05 * y if geo else 2 * y) + 0.3 * rng.normal(size=n)
        fit = K.variogram_fit_device(x, y, z, nlags, model, geographic=geo)
        lags, sv, raw = K.empirical_variogram(x, y, z, nlags, geographic=geo)
        st, pts, nfev = (int(v) for v in fit["info"].cpu())
        dev = fit["params"].cpu().numpy()[: 2 if model == "linear" else 3]
        host = K.fit_variogram_native(lags, sv, model)
        ref = fit_variogram(lags, sv, model)
        same_l = np.array_equal(fit["lags"].cpu().numpy()[:pts], lags)
        print(f"{model:11s} nl={nlags:2d} geo={int(geo)} st={st} nfev={nfev:3d} lags_equal={same_l}  dev-host {np.max(np.abs(dev-host)/np.maximum(1,np.abs(host))):.1e}  host-scipy {np.max(np.abs(host-ref)/np.maximum(1,np.abs(ref))):.1e}  dev-scipy {np.max(np.abs(dev-ref)/np.maximum(1,np.abs(ref))):.1e}")
        print("      dev ", dev, "\n      host", host, "\n      ref ", ref)
