"""Builds libmce_b200.so (sm_100a) in-tree with nvcc.  `python -m cauchyfriendly_b200.build`"""
import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(PKG_DIR, "csrc", "mce_capi.cu")
OUT = os.path.join(PKG_DIR, "libmce_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "--fmad=false",            # FMA contraction changes the reference's epsilon decisions
    "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
]


def _sources():
    d = os.path.join(PKG_DIR, "csrc")
    return [os.path.join(d, f) for f in os.listdir(d)] + [os.path.join(PKG_DIR, "..", "include", "mce_b200.h")]


def build(force=False, verbose=False):
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in _sources()):
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [SRC, "-o", OUT]
    print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
