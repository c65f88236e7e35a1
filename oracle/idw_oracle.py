"""IDW oracle: array-level restatement of the reference reconstructor.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Follows ``/root/reference/src/climatrix/reconstruct/idw.py`` line by line,
with the genuine ``scipy.spatial.cKDTree`` for the neighbour query.
"""

from __future__ import annotations

import numpy as np
from scipy.spatial import cKDTree


def dense_points(lat: np.ndarray, lon: np.ndarray) -> np.ndarray:
    """(n_lat*n_lon, 2) [lat, lon] rows, lat-major.

    Reference: ``DenseDomain.get_all_spatial_points`` dataset/domain.py:917-922.
    """
    lat_grid, lon_grid = np.meshgrid(lat, lon, indexing="ij")
    return np.stack((lat_grid.flatten(), lon_grid.flatten()), axis=1)


def sparse_points(lat: np.ndarray, lon: np.ndarray) -> np.ndarray:
    """(n, 2) [lat, lon]. Reference: ``SparseDomain.get_all_spatial_points`` domain.py:738."""
    return np.stack((lat, lon), axis=1)


def knn_ckdtree(points: np.ndarray, queries: np.ndarray, k: int, workers: int = -1):
    """The reference's neighbour query, idw.py:155-162 (tree defaults, p=2, eps=0)."""
    tree = cKDTree(points)
    dists, idxs = tree.query(queries, k=k, workers=workers)
    if k == 1:
        idxs = idxs[..., np.newaxis]
        dists = dists[..., np.newaxis]
    return dists, idxs


def knn_bruteforce(points: np.ndarray, queries: np.ndarray, k: int, chunk: int = 2048):
    """Exact float64 k-NN with the tie rule the CUDA kernel implements.

    Squared distance is ``dx*dx + dy*dy`` in float64 *without* FMA (what cKDTree
    evaluates), neighbours ordered lexicographically by
    ``(dist2, index)``.  Returns ``(sqrt(dist2), index)`` like cKDTree.
    """
    points = np.asarray(points, dtype=np.float64)
    queries = np.asarray(queries, dtype=np.float64)
    n = points.shape[0]
    m = queries.shape[0]
    if k > n:
        raise IndexError(f"k={k} exceeds the number of samples {n}")
    out_d = np.empty((m, k), dtype=np.float64)
    out_i = np.empty((m, k), dtype=np.int64)
    for s in range(0, m, chunk):
        q = queries[s : s + chunk]
        dx = q[:, None, 0] - points[None, :, 0]
        dy = q[:, None, 1] - points[None, :, 1]
        d2 = dx * dx + dy * dy
        # stable argsort == lexicographic (d2, index)
        order = np.argsort(d2, axis=1, kind="stable")[:, :k]
        out_i[s : s + chunk] = order
        out_d[s : s + chunk] = np.sqrt(np.take_along_axis(d2, order, axis=1))
    return out_d, out_i


def idw_weights_and_mean(dists, idxs, values, power, k_min):
    """idw.py:163-180 on precomputed neighbour tables; ``values`` is (N,)."""
    dists = np.maximum(dists, 1e-10)
    weights = 1.0 / np.power(dists, power)
    weights /= np.nansum(weights, axis=1, keepdims=True)

    knn_data = values[idxs]
    valid_mask = np.isfinite(knn_data)
    weights[~valid_mask] = 0.0
    weights_sum = np.nansum(weights, axis=1).squeeze()
    with np.errstate(invalid="ignore"):
        num = np.nansum(knn_data * weights, axis=1)
    # The reference calls np.divide(..., where=weights_sum != 0) without `out`,
    # leaving those slots uninitialised; they are always overwritten with NaN
    # below because weights_sum == 0 implies zero finite neighbours < k_min.
    interp_vals = np.full(num.shape, np.nan)
    np.divide(num, weights_sum, out=interp_vals, where=weights_sum != 0)
    valid_neighbor_counts = np.isfinite(knn_data).sum(axis=1)
    interp_vals[valid_neighbor_counts < k_min] = np.nan
    return interp_vals


def idw_oracle(
    sample_points: np.ndarray,
    values: np.ndarray,
    query_points: np.ndarray,
    k: int = 5,
    power: float = 2.0,
    k_min: int = 2,
    workers: int = -1,
    return_neighbours: bool = False,
):
    """Static IDW exactly as ``IDWReconstructor.reconstruct`` computes it.

    ``sample_points`` (N,2) [lat,lon]; ``values`` any shape flattening to (N,);
    ``query_points`` (M,2).  Returns (M,) float64 (plus dists/idxs on request).
    """
    values = np.asarray(values).flatten().squeeze()  # idw.py:135
    dists, idxs = knn_ckdtree(sample_points, query_points, k, workers=workers)
    out = idw_weights_and_mean(dists, idxs, values, power, k_min)
    if return_neighbours:
        return out, dists, idxs
    return out


def idw_oracle_batched(sample_points, values_tn, query_points, k=5, power=2.0, k_min=2, workers=-1):
    """Reference semantics for a time/level batch: loop the static call over slices.

    The reference rejects dynamic datasets (idw.py:87-104), so a batch is T
    independent static reconstructions.  ``values_tn`` is (T, N).  The kd-tree
    query does not depend on the slice, so it is hoisted (identical results).
    """
    values_tn = np.asarray(values_tn)
    dists, idxs = knn_ckdtree(sample_points, query_points, k, workers=workers)
    out = np.empty((values_tn.shape[0], query_points.shape[0]), dtype=np.float64)
    for t in range(values_tn.shape[0]):
        out[t] = idw_weights_and_mean(dists.copy(), idxs, values_tn[t], power, k_min)
    return out
