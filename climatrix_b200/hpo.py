"""Hyper-parameter search inner loop on the device.

climatrix's ``HParamFinder`` calls ``train.reconstruct(val.domain, method="idw", k=..., power=...)``
once per trial and scores it against the validation values
(reference src/climatrix/optim/bayesian.py:380-425, metrics comparison.py:268-306).  The neighbour
search depends only on the largest ``k`` tried, so it runs once (K1, ``cmx_knn``); every trial is
then one ``cmx_idw_apply_*`` launch on the stored table plus a reduction.
"""

from __future__ import annotations

import numpy as np

__all__ = ["IDWSearch"]

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
        (comparison.py:268-306 use ``nanmean`` the same way)."""
        if metric not in _METRICS:
            raise ValueError(f"metric must be one of {_METRICS}")
        torch = self._torch
        diff = self.reconstruct(k, power, k_min) - self.truth
        if metric == "mae":
            return float(torch.nanmean(diff.abs()))
        mse = float(torch.nanmean(diff * diff))
        return mse if metric == "mse" else float(np.sqrt(mse))
