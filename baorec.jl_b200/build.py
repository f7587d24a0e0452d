"""Builds libbaorec_b200.so (sm_100a) and the oracle's C helpers, in-tree.

    python baorec.jl_b200/build.py [--force]

Plain nvcc: one object per .cu (compiled in parallel), linked into
baorec.jl_b200/lib/libbaorec_b200.so against cuFFT and NCCL.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent
CSRC = HERE / "csrc"
LIBDIR = HERE / "lib"
OBJDIR = HERE / "build"
LIB = LIBDIR / "libbaorec_b200.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
          "-I", str(ROOT / "include"), "-I", str(CSRC)]


def _newer(target: Path, deps) -> bool:
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(t >= d.stat().st_mtime for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    LIBDIR.mkdir(exist_ok=True)
    OBJDIR.mkdir(exist_ok=True)
    srcs = sorted(CSRC.glob("*.cu"))
    hdrs = sorted(CSRC.glob("*.cuh")) + sorted((ROOT / "include").glob("*.h"))
    jobs = []
    for s in srcs:
        o = OBJDIR / (s.stem + ".o")
        if force or not _newer(o, [s] + hdrs):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        cmd = [NVCC, *ARCH, *CFLAGS, "-c", str(s), "-o", str(o)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        (OBJDIR / (s.stem + ".ptxas.log")).write_text(r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {s.name}:\n{r.stderr}")
        return s.name

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for name in ex.map(compile_one, jobs):
                if verbose:
                    print("compiled", name)
    objs = [OBJDIR / (s.stem + ".o") for s in srcs]
    if force or jobs or not _newer(LIB, objs):
        cmd = [NVCC, *ARCH, "-shared", "-o", str(LIB), *map(str, objs),
               "-L/usr/local/cuda/lib64", "-lcufft", "-lnccl",
               "-Xlinker", "-rpath,/usr/local/cuda/lib64"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stderr}")
        if verbose:
            print("linked", LIB)
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose=True)
    print(p)
