"""Build libstmpm.so in-tree with nvcc for sm_100a (no torch extension machinery: the product is a plain C-ABI .so)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

from . import LIB_PATH, PACKAGE_DIR, REPO_ROOT

CSRC = os.path.join(PACKAGE_DIR, "csrc")
SOURCES = [os.path.join(CSRC, "stmpm.cu")]
HEADERS = [os.path.join(CSRC, "stmpm_device.cuh"), os.path.join(REPO_ROOT, "include", "stmpm.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: cannot build libstmpm.so")
    return p


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(f) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB_PATH
    cmd = [nvcc_path(), *NVCC_FLAGS, "-o", LIB_PATH, *SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libstmpm.so")
    return LIB_PATH


if __name__ == "__main__":
    build(force=True, verbose=True)
