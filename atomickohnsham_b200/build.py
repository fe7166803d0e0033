"""Build the CUDA library in-tree: atomickohnsham_b200/libaks_b200.so (sm_100a only).

The rebuild is gated on a SHA-256 of the sources and the compiler flags (stored beside the library in
`libaks_b200.so.srchash`), not on modification times: the binary travels to the GPU box with the snapshot, where
mtimes say nothing about whether it matches the sources it travels with.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "aks_b200.cu")
LIB = os.path.join(HERE, "libaks_b200.so")
HASHFILE = LIB + ".srchash"
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def source_hash() -> str:
    """SHA-256 over every file of csrc/ and include/ (name + content, sorted) and the nvcc flags."""
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for root in (os.path.join(HERE, "csrc"), os.path.join(HERE, "..", "include")):
        for f in sorted(os.listdir(root)):
            path = os.path.join(root, f)
            if os.path.isfile(path):
                h.update(f.encode())
                h.update(open(path, "rb").read())
    return h.hexdigest()


def library_is_current() -> bool:
    try:
        return os.path.exists(LIB) and open(HASHFILE).read().strip() == source_hash()
    except OSError:
        return False


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/aks_b200.cu with nvcc for sm_100a; returns the .so path."""
    if not force and library_is_current():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    with open(HASHFILE, "w") as f:
        f.write(source_hash() + "\n")
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
