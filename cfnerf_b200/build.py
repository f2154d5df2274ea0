"""Build the CUDA library in-tree: cfnerf_b200/libcfnerf_b200.so (sm_100a only, nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import glob
import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libcfnerf_b200.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
    "-shared", "-Xptxas", "-v", "--threads", "0",
]


def sources():
    return sorted(glob.glob(os.path.join(PKG_DIR, "csrc", "*.cu")))


def headers():
    return sorted(glob.glob(os.path.join(PKG_DIR, "csrc", "*.h")) + glob.glob(os.path.join(PKG_DIR, "csrc", "*.cuh")) +
                  glob.glob(os.path.join(os.path.dirname(PKG_DIR), "include", "*.h")))


def needs_build() -> bool:
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in sources() + headers())


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ into one shared library.  Returns the library path."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("CFN_NVCC_EXTRA", "").split()     # e.g. -DCFN_TC_ISSUE_STAMPS=1 for the issuer-timeline build
    cmd = [nvcc] + NVCC_FLAGS + extra + ["-o", LIB_PATH] + sources()
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = os.path.join(PKG_DIR, "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + res.stdout)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed (exit {res.returncode}); see {log}")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
