"""Build libclimatrix_b200.so (hand-written CUDA, sm_100a only) in-tree with nvcc.

The shared library is a plain C ABI (``include/climatrix_b200.h``): no torch, no
pybind.  It lands in ``climatrix_b200/lib/`` so that it travels with the
repository snapshot to the GPU box (``*.so`` is git-ignored, not gpurun-ignored).
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.environ.get("CMX_LIB") or os.path.join(LIB_DIR, "libclimatrix_b200.so")  # CMX_LIB: tuning variants
SOURCES = ["capi.cu", "knn.cu", "knn3.cu", "idw.cu", "variogram.cu", "kriging.cu", "global_kriging.cu"]
HEADERS = ["common.cuh", "rerank.cuh", "ok_chol.cuh", "ok_chol2.cuh", "variogram_fit.cuh", os.path.join("..", "..", "include", "climatrix_b200.h")]

# variogram.cu holds the restated SciPy TRF fit: its stopping tests compare finite-difference noise with 1e-8, so the
# device build must round like the host build (no multiply-add contraction) to stop at the same iterate
EXTRA_FLAGS = {"variogram.cu": ["-fmad=false"]}

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libclimatrix_b200.so cannot be built")


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ for sm_100a and link the shared library."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = _nvcc()
    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = os.path.join(HERE, "..", "build", "obj")
    os.makedirs(obj_dir, exist_ok=True)
    from concurrent.futures import ThreadPoolExecutor

    def compile_one(src):
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *EXTRA_FLAGS.get(src, []), "-c", os.path.join(CSRC, src), "-o", obj]
        return src, obj, subprocess.run(cmd, capture_output=True, text=True)

    objs = []
    log = []
    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:  # one nvcc per source file
        for src, obj, r in pool.map(compile_one, SOURCES):
            log.append(r.stderr)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError(f"nvcc failed on {src}")
            objs.append(obj)
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH, *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    with open(os.path.join(obj_dir, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        sys.stderr.write("\n".join(log))
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
