#!/usr/bin/env python
"""Per-function / per-line hot spots of a kernel from an .ncu-rep captured with --import-source on (-lineinfo build).
Usage: tools/ncu_hotspots.py report.ncu-rep [top_n]"""
import bisect, collections, csv, os, re, subprocess, sys

rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 35
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
cur = None; hdr = None
tot = collections.Counter(); byline = collections.Counter(); stall = collections.defaultdict(collections.Counter); inst = collections.Counter(); thr = collections.Counter()
for row in rows:
    if not row: continue
    if row[0] == 'File Path': cur = row[1]; hdr = None; continue
    if row[0] == 'Function Name': continue
    if row[0] == 'Line No': hdr = row; continue
    if hdr is None or not row[0].isdigit(): continue   # source-line aggregate rows only
    d = dict(zip(hdr, row))
    g = lambda k: int(d[k]) if d.get(k, '').lstrip('-').isdigit() else 0
    s = g('# Samples'); ln = int(row[0])
    tot[cur] += s; byline[(cur, ln)] += s; inst[(cur, ln)] += g('Instructions Executed'); thr[(cur, ln)] += g('Thread Instructions Executed')
    for k in hdr:
        if k.startswith('stall_') and '(' not in k: stall[(cur, ln)][k] += g(k)
T = sum(tot.values()) or 1
print("total samples", T)
for k, v in tot.most_common(): print(f"{100*v/T:5.1f}% {os.path.basename(k)}")
def funcs(path):
    p = path if os.path.exists(path) else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "rlgymppo_cpp_b200", "csrc", os.path.basename(path))
    out = []
    for i, l in enumerate(open(p), 1):
        m = re.match(r'^(?:RL_HD|RL_HDI|__global__|static|inline|__device__|template)[^;]*?\b([A-Za-z_0-9]+)\s*\(', l)
        if m and not l.strip().endswith(';'): out.append((i, m.group(1)))
    return out
byfn = collections.Counter(); fst = collections.defaultdict(collections.Counter); fi = collections.Counter(); ft = collections.Counter(); cache = {}
for (f, ln), v in byline.items():
    if f not in cache: cache[f] = funcs(f)
    fl = cache[f]; idx = bisect.bisect_right([x[0] for x in fl], ln) - 1
    key = (os.path.basename(f), fl[idx][1] if idx >= 0 else '?')
    byfn[key] += v; fi[key] += inst[(f, ln)]; ft[key] += thr[(f, ln)]
    for k, c in stall[(f, ln)].items(): fst[key][k] += c
print("--- by function: %samples | warp-insts | avg active lanes | top stalls")
for k, v in byfn.most_common(topn):
    print(f"{100*v/T:5.1f}% {fi[k]:>10d} {ft[k]/max(fi[k],1):5.1f}  {k[0]}:{k[1]}  " + " ".join(f"{a.replace('stall_','')}={100*b/max(v,1):.0f}%" for a, b in fst[k].most_common(3)))
print("--- top lines")
for (f, ln), v in byline.most_common(topn):
    print(f"{100*v/T:5.1f}% {os.path.basename(f)}:{ln}  " + " ".join(f"{a.replace('stall_','')}={100*b/max(v,1):.0f}%" for a, b in stall[(f, ln)].most_common(2)))
