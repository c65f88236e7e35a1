"""Nearest-point matching of two sparse datasets on the device.

climatrix's ``Comparison`` pairs every predicted point with its nearest true point through
``cKDTree(true_points).query(pred_points[, distance_upper_bound])`` and differences the values
(reference src/climatrix/comparison.py:112-165).  The k = 1 query is K1 / K1c with k = 1.
"""

from __future__ import annotations

import numpy as np

__all__ = ["match_nearest", "sparse_difference"]


def match_nearest(pred_points: np.ndarray, true_points: np.ndarray, distance_threshold: float | None = None, device=None):
    """``(indices, valid_mask, distances)`` like the reference's query: for every predicted point the index of
    the nearest true point (Euclidean in the flat (lat, lon) plane, ties to the lower index); with a threshold,
    ``valid`` marks matches with ``distance < threshold`` (0.0 means machine epsilon, comparison.py:131-140)."""
    from . import kernels

    pred_points, true_points = np.asarray(pred_points, np.float64), np.asarray(true_points, np.float64)
    if pred_points.ndim != 2 or true_points.ndim != 2 or pred_points.shape[1] != 2 or true_points.shape[1] != 2:
        raise ValueError("expected (n, 2) arrays of (lat, lon)")
    dev = kernels._require_cuda(device)
    targets = kernels.Targets.make(pred_points[:, 0], pred_points[:, 1], False, dev)
    dist, idx = kernels.knn(true_points[:, 0], true_points[:, 1], targets, 1)
    dist, idx = dist.cpu().numpy().reshape(-1), idx.cpu().numpy().reshape(-1).astype(np.int64)
    if distance_threshold is None:
        return idx, np.ones(len(dist), dtype=bool), dist
    threshold = np.finfo(float).eps if distance_threshold == 0.0 else float(distance_threshold)
    return idx, dist < threshold, dist  # cKDTree's distance_upper_bound is exclusive


def sparse_difference(pred_values, pred_points, true_values, true_points, distance_threshold: float | None = None,
                      device=None):
    """``(differences, kept)``: ``pred - true[nearest]`` for the predicted points that found a partner, and their
    positions in the predicted dataset (the reference's ``_compute_sparse_diff`` without the xarray wrapping)."""
    idx, valid, _ = match_nearest(pred_points, true_points, distance_threshold, device)
    kept = np.where(valid)[0]
    pred_values, true_values = np.asarray(pred_values), np.asarray(true_values)
    return (pred_values - true_values[idx])[kept], kept
