#!/usr/bin/env python
"""Join an ncu SASS source page with nvdisasm -gi inline chains and aggregate executed instructions of a given opcode class
by source frame.  Usage: tools/ncu_sass_join.py report.ncu-rep obj.o kernel_substr [opcode_regex] [depth]
Prints: (a) per-opcode-class totals, (b) top source frames (file:line of every frame in the inline chain) by executed count."""
import collections, csv, os, re, subprocess, sys, tempfile
rep, obj, kern = sys.argv[1], sys.argv[2], sys.argv[3]
opre = re.compile(sys.argv[4] if len(sys.argv) > 4 else r'^(STL|LDL)')
topn = int(sys.argv[5]) if len(sys.argv) > 5 else 40
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
cub = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", cub], capture_output=True, text=True).stdout
insts = []; chain = []; insec = False; fresh = True
for l in dis.splitlines():
    if l.startswith('\t.section'):
        insec = kern in l; continue
    if not insec: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        if fresh: chain = []; fresh = False
        chain.append((os.path.basename(m.group(1)), int(m.group(2)))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        insts.append((int(m.group(1), 16), m.group(2).strip(), tuple(chain))); fresh = True
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = None; data = []
for r in rows:
    if r and r[0] == 'Address': hdr = r; continue
    if hdr and len(r) == len(hdr): data.append(dict(zip(hdr, r)))
assert len(data) == len(insts), (len(data), len(insts))
def strip_pred(t): return re.sub(r'^@!?U?P\d+\s+', '', t)
tot = collections.Counter(); byframe = collections.Counter(); byleafchain = collections.Counter(); samples = collections.Counter()
for (off, text, chain), drow in zip(insts, data):
    op = strip_pred(text)
    n = int(drow['Instructions Executed'] or 0)
    cls = op.split()[0].split('.')[0]
    tot[cls] += n
    if opre.search(op):
        for fr in set(chain): byframe[fr] += n
        byleafchain[chain[-3:] if len(chain) >= 3 else chain] += n
T = sum(tot.values())
print("total warp insts", T)
for k, v in tot.most_common(25): print(f"  {k:10s} {v:12d} {100*v/T:5.1f}%")
sel = sum(v for k, v in tot.items() if opre.search(k))
print("selected", sel)
print("--- frames (any depth) by executed count of selected opcodes")
for k, v in byframe.most_common(topn): print(f"  {v:10d} {100*v/max(sel,1):5.1f}%  {k[0]}:{k[1]}")
print("--- outermost 3 frames")
for k, v in byleafchain.most_common(topn): print(f"  {v:10d} {100*v/max(sel,1):5.1f}%  " + " <- ".join(f"{a}:{b}" for a, b in k))
