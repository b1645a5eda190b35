#!/usr/bin/env python
"""Opcode histogram per kernel of the built library (cuobjdump -sass): instruction count, image size, and the mnemonics that prove
the tcgen05 / TMEM / TMA paths (B200_PROFILING.md) plus local-memory and barrier traffic.  Usage: tools/sass_hist.py [lib.so] > profiles/sass_r02.txt"""
import collections, re, subprocess, sys

lib = sys.argv[1] if len(sys.argv) > 1 else "rlgymppo_cpp_b200/csrc/librlgym_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEY = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "UTCCP", "SYNCS", "LDGSTS", "LDL", "STL", "LDS", "STS", "LDG", "STG", "BAR", "FFMA", "FMUL", "FADD",
       "MUFU", "CALL", "BRA"]
kern = None; hist = {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1); hist[kern] = collections.Counter(); continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        hist[kern]["_n"] += 1
        hist[kern][m.group(1).split(".")[0]] += 1
print(f"# {lib}: SASS opcode histogram per kernel (sm_100a); n = instructions, image = n x 16 B")
for k, h in sorted(hist.items(), key=lambda kv: -kv[1]["_n"]):
    name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip().replace("(anonymous namespace)::", "").split("(")[0][-60:]
    keys = " ".join(f"{m}={h[m]}" for m in KEY if h[m])
    print(f"{name:<62} n={h['_n']:>6} image={h['_n'] * 16 // 1024:>4} KB  {keys}")
