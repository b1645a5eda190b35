"""Builds csrc/librlgym_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(CSRC, "librlgym_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # the gym layer must be bit-exact against the FMA-free x86 reference build
    "-fmad=false", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]


def needs_build() -> bool:
    if os.environ.get("RLG_B200_LIB"):
        return False
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".h"))]
    deps.append(os.path.join(os.path.dirname(_HERE), "include", "rlgym_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + sources() + ["-o", LIB]
    env = dict(os.environ)
    env.pop("CXX", None)  # an inherited /opt/gcc wrapper links libstdc++ statically; let nvcc pick the distro g++
    subprocess.check_call(cmd, env=env)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
