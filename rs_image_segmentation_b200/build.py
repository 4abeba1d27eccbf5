"""Build librsx.so (the C-ABI library of include/rsx.h) in-tree with nvcc for sm_100a.

The library is compiled ahead of time next to the sources so that it travels with a snapshot of
the repository to a GPU box; nothing is JIT-compiled at import time.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "librsx.so")
# (source, object name, extra flags): the KMeans assign kernels are one source compiled once per range of D
SOURCES = [("rsx_core.cu", "rsx_core", []), ("rsx_raster_kernels.cu", "rsx_raster_kernels", []), ("rsx_glcm.cu", "rsx_glcm", []), ("rsx_stencil.cu", "rsx_stencil", []), ("rsx_pca_planar.cu", "rsx_pca_planar", []),
           ("rsx_kmeans.cu", "rsx_kmeans", []), ("rsx_kmeans_seed.cu", "rsx_kmeans_seed", [])] + [("rsx_kmeans_part.cu", f"rsx_kmeans_part{i}", [f"-DRSX_KM_PART={i}"]) for i in range(6)]
NVCC_FLAGS = ["-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "--expt-relaxed-constexpr"]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: librsx.so cannot be built on this machine")
    return exe


def _digest(paths, extra=()):
    h = hashlib.sha256()
    h.update(" ".join(extra).encode())
    for p in sorted(paths):
        h.update(p.encode())
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "rsx.h"))
    return hdrs


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    deps = _deps()
    objs, jobs = [], []
    for src, name, extra in SOURCES:
        sp = os.path.join(CSRC, src)
        op = os.path.join(OBJ, name + ".o")
        stamp = op + ".sha"
        dig = _digest([sp] + deps, extra)
        objs.append(op)
        if not force and os.path.exists(op) and os.path.exists(stamp) and open(stamp).read() == dig:
            continue
        jobs.append((sp, op, stamp, dig, extra))

    def compile_one(job):
        sp, op, stamp, dig, extra = job
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", sp, "-o", op]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {sp}:\n{r.stdout}\n{r.stderr}")
        with open(stamp, "w") as f:
            f.write(dig)
        return r.stderr

    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(os.cpu_count() or 4, 8, len(jobs))) as ex:
            for log in ex.map(compile_one, jobs):
                if verbose and log:
                    print(log, file=sys.stderr)
    if jobs or force or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
