"""Build libbjx.so (hand-written sm_100a CUDA + the C ABI of include/bjx.h) in-tree with nvcc.

    python -m bijax_b200.build [--force] [--verbose]

The shared library links cudart statically and has no torch / Python dependency; the host layer binds it with ctypes.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
INCLUDE = ROOT.parent / "include"
OUT_DIR = ROOT / "_lib"
LIB = OUT_DIR / "libbjx.so"
SOURCES = ["bjx_random.cu", "bjx_elbo.cu", "bjx_lik_simt.cu", "bjx_lik_tc.cu", "bjx_gemm_tc.cu", "bjx_posterior.cu", "bjx_lowrank.cu", "bjx_p2p.cu", "bjx_minibatch.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _digest() -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(INCLUDE.glob("*.h"))):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    OUT_DIR.mkdir(exist_ok=True)
    stamp = OUT_DIR / "build.stamp"
    digest = _digest()
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == digest:
        return LIB
    nvcc = _nvcc()
    objs = []
    for src in SOURCES:
        obj = OUT_DIR / (Path(src).stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-I", str(INCLUDE), "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
        objs.append(str(obj))
    cmd = [nvcc, "-shared", "-o", str(LIB), *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcuda"]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    stamp.write_text(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
