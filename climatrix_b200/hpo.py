"""Hyper-parameter search inner loop on the device.

climatrix's ``HParamFinder`` calls ``train.reconstruct(val.domain, method="idw", k=..., power=...)``
once per trial and scores it against the validation values
(reference src/climatrix/optim/bayesian.py:380-425, metrics comparison.py:268-306).  The neighbour
search depends only on the largest ``k`` tried, so it runs once (K1, ``cmx_knn``); every trial is
then ONE launch on the stored table - weighting and the NaN-aware error reduction in the same kernel
(``cmx_idw_score_*``), the field is never written - plus a 3-value read-back.

Ordinary kriging (``OKSearch``): the trials vary ``nlags``, ``anisotropy_scaling``, ``coordinates_type`` and
``variogram_model`` (reference reconstruct/kriging.py:77-91).  Samples, validation points and held-out values
stay on the device; the neighbour table of the validation points is kept per (coordinates_type,
anisotropy_scaling), the empirical variogram per (..., nlags); a trial whose frame is cached is a 20 us fit,
K4 on the stored table (``cmx_ok_solve_*``) and the score reduction (``cmx_score_*``).
"""

from __future__ import annotations

import numpy as np

__all__ = ["IDWSearch", "OKSearch"]

_METRICS = ("mae", "mse", "rmse")


class IDWSearch:
    """Evaluate many ``(k, power, k_min)`` trials of IDW against held-out values.

    >>> search = IDWSearch(train_dataset, val_dataset, k_max=50)
    >>> score = search.score(k=27, power=2.87, k_min=2, metric="mae")   # lower is better
    """

    def __init__(self, train, val, k_max: int = 50, device=None):
        import torch

        from . import kernels
        from .idw import spatial_values

        self._torch = torch
        self._kernels = kernels
        pts = train.domain.get_all_spatial_points()
        vals, layout, batch = spatial_values(train)
        if batch:
            raise NotImplementedError("IDWSearch scores static datasets (one slice), like the reference's HParamFinder")
        dev = kernels._require_cuda(device)
        self.k_max = int(min(k_max, pts.shape[0]))
        tgt = val.domain
        targets = kernels.Targets.make(tgt.latitude.values, tgt.longitude.values, not tgt.is_sparse, dev)
        self.values = torch.as_tensor(np.ascontiguousarray(vals.reshape(-1), dtype=np.float64)).to(dev)
        self.dist, self.idx = kernels.knn(pts[:, 0], pts[:, 1], targets, self.k_max)
        self.truth = torch.as_tensor(np.ascontiguousarray(np.asarray(val.da.values, dtype=np.float64).reshape(-1))).to(dev)
        if self.truth.numel() != targets.count:
            raise ValueError("validation values do not match the validation domain")

    def reconstruct(self, k: int, power: float = 2.0, k_min: int = 2):
        """The IDW field on the validation domain for one trial (device tensor)."""
        if not 1 <= int(k) <= self.k_max:
            raise ValueError(f"k must be in [1, {self.k_max}]")
        if int(k_min) > int(k) or int(k_min) < 1:
            raise ValueError("k_min must be in [1, k]")
        return self._kernels.idw_apply(self.idx, self.dist, self.values, k=int(k), power=float(power), k_min=int(k_min))

    def score(self, k: int, power: float = 2.0, k_min: int = 2, metric: str = "mae") -> float:
        """NaN-aware MAE / MSE / RMSE between the trial's reconstruction and the held-out values
        (comparison.py:268-306 use ``nanmean`` the same way); weighting and reduction are one kernel."""
        if metric not in _METRICS:
            raise ValueError(f"metric must be one of {_METRICS}")
        if not 1 <= int(k) <= self.k_max:
            raise ValueError(f"k must be in [1, {self.k_max}]")
        if int(k_min) > int(k) or int(k_min) < 1:
            raise ValueError("k_min must be in [1, k]")
        return self._kernels.idw_score(self.idx, self.dist, self.values, self.truth, k=int(k), power=float(power),
                                       k_min=int(k_min))[metric]


class OKSearch:
    """Evaluate many ordinary-kriging trials against held-out values (the reference's ``HParamFinder`` loop for
    ``method="ok"``).

    >>> search = OKSearch(train_dataset, val_dataset, n_closest_points=32)
    >>> search.score(nlags=8, anisotropy_scaling=1.0, variogram_model="spherical", metric="rmse")

    ``n_closest_points=None`` scores the global system climatrix ships (one factorisation per trial)."""

    def __init__(self, train, val, n_closest_points: int | None = 32, device=None, max_cached_frames: int = 8):
        import torch

        from . import kernels
        from .kriging import _minmax_scale

        self._torch, self._kernels = torch, kernels
        dev = kernels._require_cuda(device)
        self.device = dev
        src, tgt = train.domain, val.domain
        if not src.is_sparse:
            raise ValueError("ordinary kriging needs a sparse training dataset (kriging.py:130-132)")
        raw = np.asarray(train.da.values).astype(float).squeeze()
        if raw.ndim != 1:
            raise NotImplementedError("OKSearch scores static datasets (one slice), like the reference's HParamFinder")
        self.v_mean, self.v_std = float(np.nanmean(raw)), float(np.nanstd(raw))
        if self.v_std == 0:
            raise ValueError("values have zero standard deviation")
        self.n = raw.size
        self.k = None if n_closest_points is None else int(min(n_closest_points, self.n))
        # kriging.py:207-214 / 230-235: samples and targets are min-max normalised SEPARATELY
        self._lon, self._lat = _minmax_scale(src.longitude.values), _minmax_scale(src.latitude.values)
        self._t_lon, self._t_lat = _minmax_scale(tgt.longitude.values), _minmax_scale(tgt.latitude.values)
        self._grid = not tgt.is_sparse
        self.z = torch.as_tensor(np.ascontiguousarray((raw - self.v_mean) / self.v_std)).to(dev)
        truth = np.ascontiguousarray(np.asarray(val.da.values, dtype=np.float64).reshape(-1))
        m = self._t_lat.size * self._t_lon.size if self._grid else self._t_lat.size
        if truth.size != m:
            raise ValueError("validation values do not match the validation domain")
        self.truth = torch.as_tensor(truth).to(dev)
        self._frames, self._variograms, self._max_frames = {}, {}, int(max_cached_frames)

    # -- per (coordinates_type, anisotropy_scaling): adjusted coordinates and the validation points' neighbours --
    def _frame(self, coordinates_type: str, scaling: float):
        from .kriging import _anisotropy

        torch, kernels, dev = self._torch, self._kernels, self.device
        geographic = coordinates_type == "geographic"
        key = (coordinates_type, 1.0 if geographic else float(scaling))
        if key in self._frames:
            return key, self._frames[key]
        if len(self._frames) >= self._max_frames:
            old = next(iter(self._frames))
            del self._frames[old]
            self._variograms = {k: v for k, v in self._variograms.items() if k[0] != old}
        x, y = self._lon, self._lat
        if geographic:
            xc = yc = 0.0
            x_adj, y_adj = x, y
        else:
            xc, yc = (np.amax(x) + np.amin(x)) / 2.0, (np.amax(y) + np.amin(y)) / 2.0
            x_adj, y_adj = _anisotropy(x, y, xc, yc, key[1])
        frame = {"xs": torch.as_tensor(x_adj).to(dev), "ys": torch.as_tensor(y_adj).to(dev), "geographic": geographic,
                 "xc": xc, "yc": yc, "scaling": key[1]}
        if self.k is not None:
            if geographic:
                def unit(lon_deg, lat_deg):
                    lon_r, lat_r = lon_deg * np.pi / 180.0, lat_deg * np.pi / 180.0
                    return np.stack((np.cos(lon_r) * np.cos(lat_r), np.sin(lon_r) * np.cos(lat_r), np.sin(lat_r)), axis=1)

                gx, gy = (self._t_lon, self._t_lat) if not self._grid else (a.ravel() for a in np.meshgrid(self._t_lon, self._t_lat))
                frame["d2"], frame["idx"] = kernels.knn3(unit(x_adj, y_adj), unit(gx, gy), self.k)
            else:
                if self._grid:
                    tx, _ = _anisotropy(self._t_lon, np.full_like(self._t_lon, yc), xc, yc, key[1])
                    _, ty = _anisotropy(np.full_like(self._t_lat, xc), self._t_lat, xc, yc, key[1])
                else:
                    tx, ty = _anisotropy(self._t_lon, self._t_lat, xc, yc, key[1])
                targets = kernels.Targets.make(ty, tx, self._grid, dev)
                dist, frame["idx"] = kernels.knn(frame["ys"], frame["xs"], targets, self.k)
                frame["d2"] = dist * dist
        self._frames[key] = frame
        return key, frame

    def reconstruct(self, nlags: int = 6, anisotropy_scaling: float = 1e-6, coordinates_type: str = "euclidean",
                    variogram_model: str = "linear"):
        """The kriged field on the validation domain for one trial (device tensor, de-standardised)."""
        kernels = self._kernels
        key, frame = self._frame(coordinates_type, anisotropy_scaling)
        vkey = (key, int(nlags))
        if vkey not in self._variograms:
            lags, semivariance, _ = kernels.empirical_variogram(frame["xs"], frame["ys"], self.z, int(nlags),
                                                                geographic=frame["geographic"])
            if lags.size == 0:
                raise ValueError("the empirical variogram is empty (fewer than two distinct samples)")
            self._variograms[vkey] = (lags, semivariance)
        lags, semivariance = self._variograms[vkey]
        params = kernels.fit_variogram_native(lags, semivariance, variogram_model)
        self.variogram_model_parameters = params
        if self.k is None:
            from .kriging import _anisotropy

            if frame["geographic"]:
                tx, ty = self._t_lon, self._t_lat
            elif self._grid:
                tx, _ = _anisotropy(self._t_lon, np.full_like(self._t_lon, frame["yc"]), frame["xc"], frame["yc"], frame["scaling"])
                _, ty = _anisotropy(np.full_like(self._t_lat, frame["xc"]), self._t_lat, frame["xc"], frame["yc"], frame["scaling"])
            else:
                tx, ty = _anisotropy(self._t_lon, self._t_lat, frame["xc"], frame["yc"], frame["scaling"])
            out, status = kernels.ok_global(frame["xs"], frame["ys"], self.z, tx, ty, self._grid, variogram_model=variogram_model,
                                            params=params, geographic=frame["geographic"], out_scale=self.v_std,
                                            out_shift=self.v_mean)
            if status == 2:
                raise np.linalg.LinAlgError("the kriging matrix is singular")
            return out
        out, nsing = kernels.ok_solve(frame["xs"], frame["ys"], self.z, frame["idx"], frame["d2"], variogram_model=variogram_model,
                                      params=params, geographic=frame["geographic"], out_scale=self.v_std, out_shift=self.v_mean)
        if int(nsing.item()):
            raise np.linalg.LinAlgError(f"{int(nsing.item())} kriging windows are singular")
        return out

    def score(self, nlags: int = 6, anisotropy_scaling: float = 1e-6, coordinates_type: str = "euclidean",
              variogram_model: str = "linear", metric: str = "mae") -> float:
        if metric not in _METRICS:
            raise ValueError(f"metric must be one of {_METRICS}")
        field = self.reconstruct(nlags, anisotropy_scaling, coordinates_type, variogram_model)
        return self._kernels.score(field, self.truth)[metric]
