"""Host-side plumbing of the reconstructors (no GPU): the staging copy and the coordinate fast path.

Reference call sites: reconstruct/idw.py:150-155 (``get_all_spatial_points`` feeding the kd-tree);
dataset/domain.py:715-738 (sparse) and :898-922 (dense) for the order of the points.
"""

import numpy as np
import pytest
import torch

from climatrix_b200 import kernels
from climatrix_b200.idw import source_coordinates

from helpers import dense_dataset, sparse_dataset, static_sparse


@pytest.mark.parametrize("threads", ["1", "3", "8"])
@pytest.mark.parametrize("dst_dtype", [torch.float64, torch.float32])
def test_host_copy_equals_copy_for_every_split(monkeypatch, threads, dst_dtype):
    """``kernels._host_copy`` (the thread-pool staging copy used by ``kernels.upload``) is ``dst.copy_(src)``:
    same values for every thread count, a size that the threads do not divide, and with the dtype conversion."""
    monkeypatch.setenv("CMX_COPY_THREADS", threads)
    monkeypatch.setattr(torch, "get_num_threads", lambda: 1)  # as under torchrun (OMP_NUM_THREADS=1)
    n = (1 << 20) + 12345
    src = torch.as_tensor(np.random.default_rng(5).normal(size=n))
    dst = torch.full((n,), -7.0, dtype=dst_dtype)
    kernels._host_copy(dst, src)
    assert torch.equal(dst, src.to(dst_dtype))
    small = torch.empty(100, dtype=dst_dtype)  # below the threshold: plain copy
    kernels._host_copy(small, src[:100])
    assert torch.equal(small, src[:100].to(dst_dtype))


def test_copy_threads_is_the_share_of_the_rank(monkeypatch):
    monkeypatch.delenv("CMX_COPY_THREADS", raising=False)
    monkeypatch.setattr(kernels.os, "sched_getaffinity", lambda _pid: set(range(64)), raising=False)
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "8")
    assert kernels._copy_threads() == 8
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "32")
    assert kernels._copy_threads() == 2
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "128")
    assert kernels._copy_threads() == 1
    monkeypatch.setenv("CMX_COPY_THREADS", "5")
    assert kernels._copy_threads() == 5
    monkeypatch.setenv("CMX_COPY_THREADS", "many")  # not a number: ignored
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "")
    assert kernels._copy_threads() == 8


def test_source_coordinates_are_the_columns_of_get_all_spatial_points():
    """Sparse sources skip the (N, 2) matrix; the arrays handed to the search are the same either way."""
    rng = np.random.default_rng(3)
    lat, lon = rng.uniform(-90, 90, 257), rng.uniform(-180, 180, 257)
    for ds in (static_sparse(lat, lon, rng.normal(size=257)), sparse_dataset(lat, lon), dense_dataset()):
        pts = ds.domain.get_all_spatial_points()
        s_lat, s_lon = source_coordinates(ds.domain)
        assert s_lat.dtype == np.float64 and s_lon.dtype == np.float64
        assert s_lat.flags.c_contiguous and s_lon.flags.c_contiguous
        np.testing.assert_array_equal(s_lat, pts[:, 0])
        np.testing.assert_array_equal(s_lon, pts[:, 1])


def test_source_coordinates_rejects_a_malformed_point_matrix():
    class Broken:
        is_sparse = False

        def get_all_spatial_points(self):
            return np.zeros((4, 3))

    with pytest.raises(ValueError, match="Expected a 2D NumPy array with shape"):
        source_coordinates(Broken())
