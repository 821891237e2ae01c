"""Build librt_b200.so in-tree with nvcc for sm_100a (no torch, no JIT cache).

    python -m rad_team_b200.build [--force] [--verbose]

The .so lands next to this file so that it travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
INCLUDE = PKG.parent / "include"
LIB = PKG / "librt_b200.so"

SOURCES = ["rt_host.cu", "rt_alloc.cu", "rt_env.cu", "rt_stats.cu", "rt_comm.cu"]
HEADERS = ["rt_host.h", "rt_device.cuh", "rt_env_kernels.cuh", "rt_moments.cuh"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    # bit-exact parity with the CPU oracle: never contract a*b+c into an FMA
    "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
    "-ldl",  # header-only NVTX3 loads the profiler's injection library lazily
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: librt_b200.so cannot be built (there is no CPU fallback)")


def stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = [CSRC / s for s in SOURCES + HEADERS] + [INCLUDE / "rt_b200.h", Path(__file__)]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not stale():
        return LIB
    extra = os.environ.get("RT_B200_NVCC_EXTRA", "").split()  # e.g. -DRT_STEP_MIN_CTAS=3 for A/B builds
    cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-I", str(INCLUDE), "-o", str(LIB), *[str(CSRC / s) for s in SOURCES]]
    # the image exports CC/CXX wrappers; let nvcc pick the distro host compiler
    env = dict(os.environ)
    if Path("/usr/bin/g++").exists():
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    log = (res.stdout or "") + (res.stderr or "")
    (PKG / "build.log").write_text(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + log[-6000:])
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
