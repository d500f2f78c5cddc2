"""Build ``libdmel_b200.so`` in-tree with nvcc for sm_100a.

    python -m dmel_codec_b200.build [--force] [-v]

The host code (``dmel_b200.cu``) and every kernel variant of the fused transform
(``fused_variant.cu`` once per (n_fft, frames per tile, CTAs per SM)) are separate
translation units compiled in parallel, then linked into one shared library.
nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels
to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.environ.get("DMEL_BUILD_OUT") or os.path.join(HERE, "libdmel_b200.so")  # DMEL_BUILD_OUT: A/B builds
OBJ_DIR = os.path.join(HERE, "_obj", os.path.splitext(os.path.basename(OUT))[0])
# (n_fft, frames per tile, CTAs per SM): must match kVariants in csrc/dmel_b200.cu
VARIANTS = [(1024, 8, 3), (1024, 16, 2), (1024, 8, 2), (1024, 16, 1), (1024, 8, 1),
            (2048, 8, 2), (2048, 16, 1), (2048, 8, 1)]
DEPS = ["dmel_b200.cu", "fused_variant.cu", "fused_ws_variant.cu", "ws_kernel.cuh", "fused_variants.h", "launch_util.cuh", "logmel_kernel.cuh",
        "fft_core.cuh", "fastdiv.cuh", "codec_kernels.cuh", "extras_kernels.cuh", "fsq_kernels.cuh", "activation_kernels.cuh",
        os.path.join("..", "..", "include", "dmel_b200.h")]
NVCC_FLAGS = ["-std=c++20", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-diag-suppress", "177"]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: set NVCC or put /usr/local/cuda/bin on PATH")


STAMP = OUT + ".srchash"  # content hash of the sources the .so was built from (mtimes do not survive copies)


def _file_hash(paths) -> str:
    h = hashlib.sha256()
    for d in paths:
        full = os.path.join(CSRC, d)
        if os.path.exists(full):
            with open(full, "rb") as f:
                h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(os.environ.get("DMEL_NVCC_EXTRA", "").encode())
    h.update(os.environ.get("DMEL_BUILD_WS", "").encode())
    return h.hexdigest()


def source_hash() -> str:
    return _file_hash(DEPS)


def is_stale() -> bool:
    if not (os.path.exists(OUT) and os.path.exists(STAMP)):
        return True
    with open(STAMP) as f:
        return f.read().strip() != source_hash()


def _compile(nvcc, src, obj, defines, verbose):
    extra = os.environ.get("DMEL_NVCC_EXTRA", "").split()  # e.g. -DDMEL_ABLATION for A/B measurements
    cmd = [nvcc, *NVCC_FLAGS, *extra, *defines, "-c", os.path.join(CSRC, src), "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed ({res.returncode}): {' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
    return res.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return OUT
    try:
        nvcc = find_nvcc()
    except RuntimeError:
        if os.path.exists(OUT) and not force:
            # no compiler on this box: use the library that travelled with the repo, but say so when its
            # sources have changed since it was built
            if is_stale():
                print(f"warning: {OUT} is older than its sources and nvcc is not available; using it as is",
                      file=sys.stderr)
            return OUT
        raise
    os.makedirs(OBJ_DIR, exist_ok=True)
    jobs = [("dmel_b200.cu", os.path.join(OBJ_DIR, "dmel_b200.o"), [])]
    for n_fft, tf, occ in VARIANTS:
        jobs.append(("fused_variant.cu", os.path.join(OBJ_DIR, f"fused_{n_fft}_{tf}_{occ}.o"),
                     [f"-DDMEL_V_NFFT={n_fft}", f"-DDMEL_V_TF={tf}", f"-DDMEL_V_OCC={occ}"]))
    if os.environ.get("DMEL_BUILD_WS"):  # the warp-specialised experiment (ws_kernel.cuh): measured slower, not built by default
        jobs.append(("fused_ws_variant.cu", os.path.join(OBJ_DIR, "fused_ws_1024.o"), []))
        jobs[0] = (jobs[0][0], jobs[0][1], ["-DDMEL_WITH_WS"])
    workers = int(os.environ.get("DMEL_BUILD_JOBS", "0")) or min(len(jobs), os.cpu_count() or 4)
    with ThreadPoolExecutor(max_workers=workers) as pool:
        logs = list(pool.map(lambda j: _compile(nvcc, j[0], j[1], j[2], verbose), jobs))
    if verbose:
        for (src, obj, _), log in zip(jobs, logs):
            print(f"== {os.path.basename(obj)}\n{log}")
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT, *[j[1] for j in jobs]]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed ({res.returncode}):\n{res.stdout}\n{res.stderr}")
    with open(STAMP, "w") as f:
        f.write(source_hash())
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
