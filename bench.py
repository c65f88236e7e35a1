#!/usr/bin/env python
"""Benchmark of the IDW / ordinary-kriging reconstruction hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c5|c3|c1|c2|c4]
                    [--search brute|binned] [--gather peer|nccl] [--no-binned] [--no-e2e] [--no-cpu-baseline]

Headline workload (BASELINE.json configs[4], the largest IDW k=32 configuration; it fits one GPU):
  IDW, 1 000 000 station samples -> 0.1 deg global grid (1801 x 3600) x 37 pressure levels, k=32, power=2.
One "step" = one reconstruction of all 37 levels (one neighbour search shared by the levels +
the batched weighting), inputs resident in HBM.  metric = target points / second (M / step time,
SURVEY.md section 8d); ``output_values_per_s`` is M*37 / step time.

Under ``torchrun`` (N > 1, one rank per GPU) the target rows are sharded over the ranks
(strong scaling: the job is fixed), samples are replicated, and every rank holds the whole
field at the end of the timed step.

The headline numbers use the brute-force neighbour search K1 (north_star); the same step with the cell-binned search
K1c (identical output) is measured right after it and reported as ``binned_search`` in the same JSON line
(``--search binned`` makes it the headline instead, ``--no-binned`` skips it).  For N > 1 ``config.gather`` says how
the bands became the full field on every rank: stored by the weighting kernel into all ranks' buffers over NVLink
peer memory (default) or an NCCL all-gather (``--gather nccl``).

``--impl reference`` times the reference's CPU implementation of the same path (scipy cKDTree +
the NumPy epilogue of idw.py:155-180, restated in oracle/ with the real scipy) on the host cores,
each step a bounded sample of the workload (see ``cpu_baseline.sample``).
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (method, N, n_lat, n_lon, T, k, description)
    "c1": ("idw", 1_000, 180, 360, 1, 30, "IDW 1,000 samples -> 1deg grid (180x360), k=30, power=2"),
    "c2": ("ok", 10_000, 180, 360, 1, 32, "OK 10,000 samples -> 1deg grid (180x360), k=32, spherical variogram"),
    "c3": ("idw", 100_000, 721, 1440, 365, 32, "IDW 100,000 samples -> 0.25deg ERA5 grid (721x1440) x 365 timesteps, k=32"),
    "c4": ("ok", 50_000, 721, 1440, 1, 64, "OK 50,000 samples -> 0.25deg grid (721x1440), k=64, spherical variogram"),
    "c5": ("idw", 1_000_000, 1801, 3600, 37, 32, "IDW 1,000,000 samples -> 0.1deg grid (1801x3600) x 37 levels, k=32, power=2"),
}


def make_inputs(name: str):
    """Seeded synthetic inputs of the BASELINE shape."""
    method, n, n_lat, n_lon, t, k, _ = WORKLOADS[name]
    rng = np.random.default_rng(1234 + int(name[1]))
    era5 = name in ("c3", "c4")
    lat = rng.uniform(-90, 90, n)
    lon = rng.uniform(0, 360, n) if era5 else rng.uniform(-180, 180, n)
    base = 15.0 + 10.0 * np.cos(np.deg2rad(lat)) + 5.0 * np.sin(2.0 * np.deg2rad(lon))
    if t == 1:
        vals = base + rng.normal(size=n)
    else:  # sample-major (point, time/level): the layout of the reference's sparse fixtures
        vals = base[:, None] + 0.5 * np.arange(t)[None, :] + rng.normal(size=(n, t))
    if era5:
        t_lat, t_lon = np.linspace(90.0, -90.0, n_lat), np.arange(n_lon) * (360.0 / n_lon)
    else:
        step = 180.0 / (n_lat - 1) if name == "c5" else 180.0 / n_lat
        if name == "c5":
            t_lat, t_lon = np.linspace(-90.0, 90.0, n_lat), -180.0 + np.arange(n_lon) * 0.1
        else:
            t_lat, t_lon = -90.0 + step * (np.arange(n_lat) + 0.5), -180.0 + (360.0 / n_lon) * (np.arange(n_lon) + 0.5)
    return dict(method=method, n=n, n_lat=n_lat, n_lon=n_lon, t=t, k=k, lat=lat, lon=lon, vals=vals,
                t_lat=t_lat, t_lon=t_lon)


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        def pump():
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except ValueError:
                continue
            for name, cell in zip(names, r[5:9]):
                if cell.lower().startswith("active"):
                    reasons.add(name)
        # "under load": the upper half of the samples (idle samples before/after the region drag the median)
        sm_sorted = sorted(sm)
        load = sm_sorted[len(sm_sorted) // 2:] if sm_sorted else []
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# CPU arm: the reference's own path (scipy cKDTree + NumPy epilogue), bounded sample
# ------------------------------------------------------------------------------------------
def cpu_reference_idw(inp, budget_rows: int | None = None):
    """Returns (targets_per_s for the WHOLE job, description of the sample, seconds spent)."""
    from scipy.spatial import cKDTree

    from oracle.idw_oracle import dense_points, idw_weights_and_mean

    n_lat, n_lon, t = inp["n_lat"], inp["n_lon"], inp["t"]
    m = n_lat * n_lon
    pts = np.stack((inp["lat"], inp["lon"]), axis=1)
    t0 = time.perf_counter()
    tree = cKDTree(pts)  # idw.py:155
    t_build = time.perf_counter() - t0
    if budget_rows is None:
        # ~3.3e5 targets/s on 8 cores: aim at ~6 s of query+epilogue
        budget_rows = max(1, min(n_lat, int(2.0e6 / n_lon)))
    rows = np.unique(np.linspace(0, n_lat - 1, budget_rows).astype(int))
    q = dense_points(inp["t_lat"][rows], inp["t_lon"])
    vals0 = inp["vals"] if t == 1 else np.ascontiguousarray(inp["vals"][:, 0])
    t0 = time.perf_counter()
    dists, idxs = tree.query(q, k=inp["k"], workers=-1)  # idw.py:158
    if inp["k"] == 1:
        dists, idxs = dists[:, None], idxs[:, None]
    with np.errstate(invalid="ignore"):
        idw_weights_and_mean(dists, idxs, vals0, 2.0, 2)  # idw.py:163-180
    t_rows = time.perf_counter() - t0
    per_target = t_rows / q.shape[0]
    # the reference rejects dynamic datasets (idw.py:87-104): a batch is T independent static calls,
    # each building its tree and running its query
    t_job = t * (t_build + per_target * m)
    sample = (f"cKDTree build on all {inp['n']} samples + query(workers=-1) + NumPy epilogue on {len(rows)} of {n_lat} "
              f"target rows ({q.shape[0]} targets), 1 of {t} slices; extrapolated to {m} targets x {t} static calls")
    return m / t_job, sample, t_build + t_rows


def cpu_reference_ok(inp, n_targets: int = 4000):
    from oracle import pykrige_restated as pk
    from oracle.ok_oracle import normalize_axis, standardize_values

    n_lat, n_lon = inp["n_lat"], inp["n_lon"]
    m = n_lat * n_lon
    t0 = time.perf_counter()
    *_, s_lat = normalize_axis(inp["lat"])
    *_, s_lon = normalize_axis(inp["lon"])
    _, _, z = standardize_values(inp["vals"])
    okr = pk.OrdinaryKriging(s_lon, s_lat, z, variogram_model="spherical", nlags=6, anisotropy_scaling=1.0)
    t_fit = time.perf_counter() - t0
    *_, t_lat = normalize_axis(inp["t_lat"])
    *_, t_lon = normalize_axis(inp["t_lon"])
    rng = np.random.default_rng(0)
    sel = rng.choice(m, size=min(n_targets, m), replace=False)
    gx, gy = np.meshgrid(t_lon, t_lat)
    t0 = time.perf_counter()
    okr.execute("points", gx.ravel()[sel], gy.ravel()[sel], backend="loop", n_closest_points=inp["k"])
    t_loop = time.perf_counter() - t0
    t_job = t_fit + t_loop / len(sel) * m
    sample = (f"restated pykrige: variogram fit on all {inp['n']} samples ({t_fit:.1f} s) + moving-window loop "
              f"(cKDTree + scipy.linalg.solve per target) on {len(sel)} random targets, extrapolated to {m}")
    return m / t_job, sample, t_fit + t_loop


def cpu_reference(inp):
    return cpu_reference_idw(inp) if inp["method"] == "idw" else cpu_reference_ok(inp)


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--search", default="brute", choices=["brute", "binned"],
                    help="neighbour search of the headline numbers: K1 brute force (north_star) or K1c cell-binned")
    ap.add_argument("--no-binned", action="store_true", help="skip the extra K1c (binned search) measurement")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"],
                    help="N > 1: how the ranks' bands become the full field on every rank: stored by the weighting kernel "
                         "itself into all ranks' buffers over NVLink peer memory (default), or an NCCL all-gather after it")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    method, n, n_lat, n_lon, t, k, desc = WORKLOADS[args.workload]
    cores = len(os.sched_getaffinity(0))
    config = {"workload": f"{args.workload}: {desc}", "samples": n, "targets": n_lat * n_lon, "batch": t, "k": k,
              "method": method, "sharding": f"target rows over {max(world, 1)} rank(s), samples replicated",
              "l2": "per-step working set (samples+values+output) exceeds the 126 MB L2; no explicit flush"}

    # ---------------- reference arm ----------------
    if args.impl == "reference":
        if rank != 0:
            return
        inp = make_inputs(args.workload)
        for _ in range(max(0, min(args.warmup, 1))):
            cpu_reference(inp)
        vals, secs = [], 0.0
        for _ in range(max(1, args.steps)):
            v, sample, s = cpu_reference(inp)
            vals.append(v)
            secs += s
        value = float(np.mean(vals))
        kind = "port"  # oracle/ restatement running the real scipy (cKDTree, LAPACK); climatrix itself is not importable here
        print(json.dumps({
            "impl": "reference", "metric": "target points/sec", "value": value, "unit": "targets/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * config["targets"] / value,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config,
            "cpu_baseline": {"value": value, "unit": "targets/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "targets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    # ---------------- our arm ----------------
    import torch

    from climatrix_b200 import kernels as K
    from climatrix_b200.sharding import distributed_bands, row_shards

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: climatrix_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist

        # NCCL prints its version banner on stdout at communicator creation; the contract is ONE JSON line
        # there, so stdout is pointed at stderr until the result is printed
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    inp = make_inputs(args.workload)
    m = n_lat * n_lon
    tdtype = torch.float64

    # inputs resident in HBM
    s_lat = torch.as_tensor(inp["lat"]).to(dev)
    s_lon = torch.as_tensor(inp["lon"]).to(dev)
    vals = torch.as_tensor(inp["vals"]).to(dev)
    t_lat_d = torch.as_tensor(inp["t_lat"]).to(dev)
    t_lon_d = torch.as_tensor(inp["t_lon"]).to(dev)
    bands = row_shards(n_lat, world)
    r0, r1 = bands[rank]
    ok_params = None
    if method == "ok":
        # variogram fit is part of the path: do it once per step (K3 + host fit) like the reference constructor
        pass

    peer = None
    if world > 1 and method == "idw" and t > 1 and args.gather == "peer":
        from climatrix_b200.sharding import PeerGather
        try:
            peer = PeerGather.get(t, n_lat, n_lon, tdtype, dev)
        except Exception as exc:  # noqa: BLE001  (no peer-mapped memory on this box: NCCL gather instead)
            print(f"[bench] symmetric memory unavailable ({type(exc).__name__}: {exc}); using the NCCL gather", file=sys.stderr)
            peer = None
        ok = torch.tensor([1 if peer is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            peer = None
    config["gather"] = ("peer stores from the weighting kernel (symmetric memory over NVLink)" if peer is not None
                        else ("NCCL all-gather" if world > 1 else "none (1 rank)"))

    def step():
        if method == "idw" and peer is not None:
            peer.begin()
            if r1 > r0:
                tg = K.Targets(t_lat_d[r0:r1].contiguous(), t_lon_d, True)
                K.idw(s_lat, s_lon, vals, tg, k=k, power=2.0, k_min=2, dtype=tdtype, layout="sample_major",
                      peer_out=peer.peer_out(r0))
            return peer.finish()
        if method == "idw":
            def compute(a, b):
                if b == a:
                    return None
                tg = K.Targets(t_lat_d[a:b].contiguous(), t_lon_d, True)
                return K.idw(s_lat, s_lon, vals, tg, k=k, power=2.0, k_min=2, dtype=tdtype,
                             layout="sample_major" if t > 1 else None)
            return distributed_bands(compute, n_lat, n_lon, t, dtype=tdtype, device=dev, gather="all")
        return ok_step(K, torch, dev, inp, s_lat, s_lon, vals, t_lat_d, t_lon_d, r0, r1, k, world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(search, steps, warmup):
        K.set_search(search)
        for _ in range(max(warmup, 0)):
            step()
        barrier()
        K.reset_launch_count()
        K.kernel_timing(True)
        sampler = ClockSampler(local_rank)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        barrier()
        ms_total = e0.elapsed_time(e1)
        r = {"clocks": sampler.stop(), "launches": K.launch_count(),
             "knn": _read_kernel(K, K.KERNEL_KNN), "knn_binned": _read_kernel(K, K.KERNEL_KNN_BINNED),
             "k2": _read_kernel(K, K.KERNEL_IDW_BATCH), "k4": _read_kernel(K, K.KERNEL_OK_SOLVE)}
        K.kernel_timing(False)
        if world > 1:
            tmax = torch.tensor([ms_total], dtype=torch.float64, device=dev)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            ms_total = float(tmax.item())
        r["ms_per_step"] = ms_total / steps
        return r

    main_run = measure(args.search, args.steps, args.warmup)
    clocks, launches, ms_per_step = main_run["clocks"], main_run["launches"], main_run["ms_per_step"]
    (knn_ms, knn_n), (k2_ms, k2_n), (k4_ms, k4_n) = main_run["knn"], main_run["k2"], main_run["k4"]
    knb_ms, knb_n = main_run["knn_binned"]
    value = m / (ms_per_step * 1e-3)

    # ---------------- roofline of the dominant kernel (K1 neighbour search, FP32 pipe) ----------------
    rows_local = r1 - r0
    flops_per_launch = 5.0 * n * rows_local * n_lon  # 5 FLOP per (target, sample) pair
    knn_avg_ms = knn_ms / max(knn_n, 1)
    achieved = flops_per_launch / (knn_avg_ms * 1e-3) / 1e12 if knn_n else None
    peak_measured = K.fp32_peak_tflops()
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    peak_nominal = sms * 128 * 2 * (clocks["sm_max_mhz"] or 1965.0) * 1e6 / 1e12
    traffic = None
    prof = os.path.join(ROOT, "profiles", "knn_traffic.json")
    if os.path.exists(prof):
        try:
            traffic = json.load(open(prof)).get(args.workload)
        except Exception:  # noqa: BLE001
            traffic = None
    roofline = {
        "kernel": "knn_kernel (K1: brute-force neighbour search + fused re-rank)",
        "bound": "fp32", "achieved": achieved, "peak": peak_measured, "unit": "TFLOP/s",
        "frac": (achieved / peak_measured) if achieved and peak_measured > 0 else None,
        "peak_source": "FFMA chain measured live in this run (cmx_fp32_peak_tflops); MEASURED_PEAKS.json has no FP32 CUDA-core figure",
        "peak_nominal": peak_nominal, "frac_of_nominal": (achieved / peak_nominal) if achieved else None,
        "algorithmic_flops_per_launch": flops_per_launch, "kernel_ms_avg": knn_avg_ms, "launches_timed": knn_n,
        "share_of_step": knn_avg_ms / ms_per_step if ms_per_step else None,
        "traffic": traffic,
    }

    # secondary kernels of the step, for context: K2 against the measured HBM copy peak, K4 as a rate
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak, hbm_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    if os.path.exists(peaks_path):
        try:
            hbm_peak, hbm_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        except Exception:  # noqa: BLE001
            pass
    other = {}
    if k2_n:
        bytes_k2 = rows_local * n_lon * (t * 8 + 12 * k)  # output written + [M][k] table read
        gbs = bytes_k2 / (k2_ms / k2_n * 1e-3) / 1e9
        other["idw_batch_kernel"] = {"bound": "hbm (nominally; on-chip gather bound at k=32, see DESIGN.md)", "achieved": gbs,
                                     "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "peak_source": hbm_src,
                                     "kernel_ms_avg": k2_ms / k2_n, "share_of_step": (k2_ms / k2_n) / ms_per_step}
    if k4_n:
        other["ok_solve_kernel"] = {"kernel_ms_avg": k4_ms / k4_n, "targets_per_s": rows_local * n_lon / (k4_ms / k4_n * 1e-3),
                                    "share_of_step": (k4_ms / k4_n) / ms_per_step}
    fused = (t == 1 and method == "idw")
    knb_bytes = 36.0 * n + rows_local * n_lon * (8.0 if fused else 12.0 * k)  # counting sort r/w + output

    def binned_entry(ms_avg, step_ms):
        gbs = knb_bytes / (ms_avg * 1e-3) / 1e9
        return {"bound": "hbm (nominally; pruned search is instruction-issue bound, see DESIGN.md K1c)", "achieved": gbs,
                "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "peak_source": hbm_src,
                "kernel_ms_avg": ms_avg, "share_of_step": ms_avg / step_ms,
                "targets_per_s": rows_local * n_lon / (ms_avg * 1e-3)}

    if knb_n:
        if not knn_n:  # headline measured with --search binned: K1c (or K2) is the dominant kernel
            ent = binned_entry(knb_ms / knb_n, ms_per_step)
            roofline = {"kernel": "knn_binned_kernel (K1c: cell-binned neighbour search, counting sort included)",
                        "bound": "hbm", "achieved": ent["achieved"], "peak": hbm_peak, "unit": "GB/s", "frac": ent["frac"],
                        "peak_source": hbm_src, "algorithmic_bytes_per_launch": knb_bytes, "kernel_ms_avg": ent["kernel_ms_avg"],
                        "launches_timed": knb_n, "share_of_step": ent["share_of_step"], "traffic": None,
                        "note": ent["bound"]}
        else:
            other["knn_binned_kernel"] = binned_entry(knb_ms / knb_n, ms_per_step)
    roofline["other_kernels"] = other

    # ---------------- end to end through the public API with host buffers ----------------
    e2e = None
    if not args.no_e2e:
        e2e = e2e_run(args, inp, rank, world, r0, r1, dev, method)

    # ---------------- the same workload with the cell-binned search (K1c) ----------------
    binned = None
    if args.search == "brute" and not args.no_binned:
        b = measure("binned", args.steps, args.warmup)
        binned = {"value": m / (b["ms_per_step"] * 1e-3), "unit": "targets/s", "ms_per_step": b["ms_per_step"],
                  "gpu_launches": int(b["launches"]),
                  "what": "identical step with cmx_set_search(CMX_SEARCH_BINNED): same neighbour tables and fields, "
                          "samples counting-sorted into grid cells, each target scans only the cells its k-th-distance bound touches"}
        if b["knn_binned"][1]:
            binned["knn_binned_kernel"] = binned_entry(b["knn_binned"][0] / b["knn_binned"][1], b["ms_per_step"])
        if b["k2"][1]:
            binned["idw_batch_kernel_ms_avg"] = b["k2"][0] / b["k2"][1]
        if b["k4"][1]:
            binned["ok_solve_kernel_ms_avg"] = b["k4"][0] / b["k4"][1]
        if not args.no_e2e:
            binned["e2e"] = e2e_run(args, inp, rank, world, r0, r1, dev, method)
        K.set_search(args.search)

    # ---------------- CPU baseline on the box's host cores (rank 0, N = 1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, sample, _ = cpu_reference(inp)
        cpu = {"value": v, "unit": "targets/s", "cores": cores, "kind": "port", "sample": sample}

    if world > 1:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
    if rank == 0:
        line = {
            "metric": "target points/sec", "value": value, "unit": "targets/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "output_values_per_s": value * t, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu, "search": args.search, "binned_search": binned,
            "notes": "fp32 distance filter + exact float64 re-rank; values/output float64; N>1: every rank ends the step with the full field (config.gather)",
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        os.dup2(2, 1)
        dist.destroy_process_group()


def _read_kernel(K, kind):
    return K.kernel_time(kind)


def ok_step(K, torch, dev, inp, s_lat, s_lon, vals, t_lat_d, t_lon_d, r0, r1, k, world):
    """Ordinary kriging step (configs c2/c4): normalise, variogram (K3) + host fit, K1 + K4 on this rank's rows."""
    from climatrix_b200.kriging import _minmax_scale, fit_variogram
    from climatrix_b200.sharding import distributed_bands

    lat_n = torch.as_tensor(_minmax_scale(inp["lat"])).to(dev)
    lon_n = torch.as_tensor(_minmax_scale(inp["lon"])).to(dev)
    v = inp["vals"]
    mean, std = float(np.nanmean(v)), float(np.nanstd(v))
    z = (vals - mean) / std
    lags, sv, _ = K.empirical_variogram(lon_n, lat_n, z, 6)
    params = fit_variogram(lags, sv, "spherical")
    tl = torch.as_tensor(_minmax_scale(inp["t_lat"])).to(dev)
    to = torch.as_tensor(_minmax_scale(inp["t_lon"])).to(dev)

    def compute(a, b):
        if b == a:
            return None
        out, _ = K.ok_moving_window(lon_n, lat_n, z, to, tl[a:b].contiguous(), True, k=k, variogram_model="spherical",
                                    params=params, out_scale=std, out_shift=mean)
        return out.view(1, -1)

    return distributed_bands(compute, inp["n_lat"], inp["n_lon"], 1, dtype=torch.float64, device=dev, gather="all")


def e2e_run(args, inp, rank, world, r0, r1, dev, method):
    """Same metric through the public reconstructor API: host (NumPy) inputs, host output, so the
    host->device copy of the inputs and the device->host copy of the field are inside the timed region."""
    import torch

    import climatrix_b200 as cm

    n, t, n_lon = inp["n"], inp["t"], inp["n_lon"]
    rows = r1 - r0
    coords = {"point": np.arange(n), "latitude": (("point",), inp["lat"]), "longitude": (("point",), inp["lon"])}
    if t > 1:
        coords["level"] = np.arange(t, dtype=np.float64)
        fa = cm.FieldArray(inp["vals"], ("point", "level"), coords, name="field")
    else:
        fa = cm.FieldArray(inp["vals"], ("point",), coords, name="field")
    src = cm.BaseClimatrixDataset(fa)
    tgt = cm.Domain.from_lat_lon(inp["t_lat"][r0:r1], inp["t_lon"], kind="dense") if rows else None
    kw = dict(k=inp["k"], power=2.0, k_min=2) if method == "idw" else dict(
        n_closest_points=inp["k"], variogram_model="spherical", nlags=6, anisotropy_scaling=1.0)

    def call():
        if tgt is None:
            return None
        return src.reconstruct(tgt, method=method, **kw)

    def barrier():
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()

    steps = max(1, min(args.steps, 3))
    for _ in range(3):  # warm-up: workspaces, and the two pinned result buffers the host allocator alternates between
        res = call()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        res = call()
    barrier()
    sec = (time.perf_counter() - t0) / steps
    if world > 1:
        import torch.distributed as dist

        tm = torch.tensor([sec], dtype=torch.float64, device=dev)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        sec = float(tm.item())
    m = inp["n_lat"] * n_lon
    h2d = (2 * n + n * t) * 8 + (rows + n_lon) * 8
    d2h = rows * n_lon * t * 8
    return {"value": m / sec, "unit": "targets/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "ms_per_step": sec * 1e3, "steps": steps,
            "api": f"BaseClimatrixDataset.reconstruct(target_domain, method='{method}', ...) with NumPy inputs and output"}


if __name__ == "__main__":
    main()
