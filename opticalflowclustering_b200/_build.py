"""Compile libofc.so (sm_100a) in-tree with nvcc.  Used by __graft_entry__.build()."""
from __future__ import annotations

import glob
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
SO = os.path.join(PKG, "libofc.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build() -> bool:
    if not os.path.exists(SO):
        return True
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(os.path.dirname(PKG), "include", "ofc.h")]
    return any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return SO
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO] + sources()
    subprocess.run(cmd, check=True)
    return SO


if __name__ == "__main__":
    print(build(force=True, verbose=True))
