#!/usr/bin/env python
"""Code size of one kernel by source function (needs a -lineinfo build). Usage: tools/sass_size.py obj.o kernel_substr"""
import bisect, collections, os, re, subprocess, sys, tempfile
obj, kern = sys.argv[1], sys.argv[2]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
cub = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
cur_sec = None; cur = None; cnt = collections.Counter(); sec = collections.Counter()
for l in dis.splitlines():
    m = re.match(r'\s*\.section\s+(\S+)', l)
    if m: cur_sec = m.group(1); continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1), int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
        sec[cur_sec] += 1
        if cur_sec and kern in cur_sec: cnt[cur] += 1
for k, v in sec.most_common(8): print(v * 16 // 1024, 'KB', k[:100])
def funcs(path):
    out = []
    for i, l in enumerate(open(path), 1):
        m = re.match(r'^(?:RL_HD|RL_HDI|__global__|static|inline|__device__|template)[^;]*?\b([A-Za-z_0-9]+)\s*\(', l)
        if m and not l.strip().endswith(';'): out.append((i, m.group(1)))
    return out
cache = {}; byfn = collections.Counter()
for (f, ln), v in cnt.items():
    if f not in cache: cache[f] = funcs(f)
    fl = cache[f]; idx = bisect.bisect_right([x[0] for x in fl], ln) - 1
    byfn[(os.path.basename(f), fl[idx][1] if idx >= 0 else '?')] += v
tot = sum(byfn.values()); print(kern, "insts", tot, tot * 16 // 1024, "KB")
for k, v in byfn.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 40): print(f"{v*16/1024:7.1f} KB {k}")
