"""Builds csrc/librlgym_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(CSRC, "librlgym_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC"]
# Per-file floating-point model:
#  * engine.cu (physics + gym layer): FMA contraction + approximate division / square root — physics parity is tolerance
#    based and the role kernel is instruction-cache bound (-22 % k_roles time, profiles/r01d_ab.md); the gym layer
#    (obs, rewards, event tracker), which must be bit-exact against the FMA-free x86 reference build, uses the strict
#    s_* helpers of rl_math.h and is unaffected by these flags.
#  * everything else (collector.cu: GAE is bit-exact against the numpy oracle) stays IEEE without contraction.
#    One consequence, measured (profiles/r02d_fp_model_ab.txt): with FMA contraction the GJK / EPA of deep hitbox contacts takes
#    single threshold decisions differently and 6 of the 13 712 recorded reference ticks land up to 0.09 uu / 2.8 uu/s from the
#    reference instead of < 0.03 uu/s (no-FMA builds reproduce the host build's parity exactly); compiling only that code IEEE needs a
#    second translation unit + relocatable device code, which costs the role kernel 28 % (1.26 -> 1.61 ms), so it stays fast.
FP_FLAGS = {"engine.cu": ["-fmad=true", "-prec-div=false", "-prec-sqrt=false"]}
FP_DEFAULT = ["-fmad=false"]


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


def build(force: bool = False, verbose: bool = False, variant: str = "", extra_flags=()) -> str:
    """variant != "": an A/B or diagnostic build (extra_flags, e.g. -DRLG_PHASE_TIMING) into build_ab/lib_<variant>.so,
    selected at run time with RLG_B200_LIB; the product library is the variant-less build."""
    if not variant and not force and not needs_build():
        return LIB
    env = dict(os.environ)
    env.pop("CXX", None)  # an inherited /opt/gcc wrapper links libstdc++ statically; let nvcc pick the distro g++
    objdir = os.path.join(CSRC, "build" + ("_" + variant if variant else ""))
    lib = LIB
    if variant:
        os.makedirs(os.path.join(os.path.dirname(_HERE), "build_ab"), exist_ok=True)
        lib = os.path.join(os.path.dirname(_HERE), "build_ab", f"lib_{variant}.so")
    os.makedirs(objdir, exist_ok=True)
    procs, objs = [], []
    for src in sources():
        name = os.path.basename(src)
        obj = os.path.join(objdir, name[:-3] + ".o")
        objs.append(obj)
        cmd = [_nvcc()] + NVCC_FLAGS + FP_FLAGS.get(name, FP_DEFAULT) + list(extra_flags) + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, env=env)))
    for cmd, p in procs:
        if p.wait() != 0:
            raise subprocess.CalledProcessError(p.returncode, cmd)
    subprocess.check_call([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + objs + ["-o", lib], env=env)
    return lib


if __name__ == "__main__":
    import sys

    if len(sys.argv) > 1:  # python -m rlgymppo_cpp_b200.build VARIANT [extra nvcc flags...]
        print(build(force=True, variant=sys.argv[1], extra_flags=sys.argv[2:]))
    else:
        print(build(force=True, verbose=True))
