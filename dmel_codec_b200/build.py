"""Build ``libdmel_b200.so`` in-tree with nvcc for sm_100a.

    python -m dmel_codec_b200.build [--force]

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels
to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libdmel_b200.so")
SOURCES = ["dmel_b200.cu"]
DEPS = ["dmel_b200.cu", "logmel_kernel.cuh", "fft_core.cuh", "codec_kernels.cuh",
        os.path.join("..", "..", "include", "dmel_b200.h")]
NVCC_FLAGS = ["-std=c++20", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-shared", "-Xcompiler", "-fPIC", "-diag-suppress", "177"]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: set NVCC or put /usr/local/cuda/bin on PATH")


STAMP = OUT + ".srchash"  # content hash of the sources the .so was built from (mtimes do not survive copies)


def source_hash() -> str:
    h = hashlib.sha256()
    for d in DEPS:
        with open(os.path.join(CSRC, d), "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(os.environ.get("DMEL_NVCC_EXTRA", "").encode())
    return h.hexdigest()


def is_stale() -> bool:
    if not (os.path.exists(OUT) and os.path.exists(STAMP)):
        return True
    with open(STAMP) as f:
        return f.read().strip() != source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return OUT
    try:
        find_nvcc()
    except RuntimeError:
        if os.path.exists(OUT) and not force:
            return OUT  # no compiler on this box: use the library that travelled with the repo
        raise
    extra = os.environ.get("DMEL_NVCC_EXTRA", "").split()  # e.g. -DDMEL_SCALAR_FP for A/B measurements
    cmd = [find_nvcc(), *NVCC_FLAGS, *extra, "-o", OUT, *[os.path.join(CSRC, s) for s in SOURCES]]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed ({res.returncode}):\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr)
    with open(STAMP, "w") as f:
        f.write(source_hash())
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
