#!/usr/bin/env python
"""Aggregate an ncu source-page (SASS) CSV by CUDA source line using nvdisasm -g line info.
usage: ncu_by_line.py <report.ncu-rep> <lib.so> <kernel-mangled-substring> [top]"""
import csv, os, re, subprocess, sys, tempfile, collections
rep, lib, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
dis = []                                   # one cubin per translation unit: look in all of them
for cub in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):
    dis += subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.split("\n")
line_of, cur, infunc = {}, None, False
for l in dis:
    if l.startswith(".text."):
        infunc = kern in l
    if not infunc:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.split("\n")))
hdr, out = None, []
for r in rows:
    if len(r) > 5 and r[0] == "Address":
        if hdr is not None and out:
            break
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        out.append(dict(zip(hdr, r)))
base = int(out[0]["Address"], 16)
f = lambda x: float(x) if x not in ("", None) else 0.0
agg = collections.defaultdict(lambda: collections.Counter())
tot = 0.0
for o in out:
    off = int(o["Address"], 16) - base
    key = line_of.get(off, ((None, None), ""))[0]
    n = f(o["# Samples"])
    tot += n
    agg[key]["samples"] += n
    for k in o:
        if k.startswith("stall_") and "(" not in k:
            agg[key][k[6:]] += f(o[k])
srcs = {}
def text(key):
    fn, ln = key if key else (None, None)
    if fn is None: return ""
    for d in ("flowfusion_b200/csrc", "include"):
        p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", d, fn)
        if os.path.isfile(p):
            if p not in srcs: srcs[p] = open(p).read().split("\n")
            return srcs[p][ln - 1].strip()[:80] if ln - 1 < len(srcs[p]) else ""
    return ""
print(f"total samples {tot:.0f}")
for key, c in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = sorted(((k, v) for k, v in c.items() if k != "samples"), key=lambda kv: -kv[1])[:3]
    print(f"{c['samples']/tot*100:5.2f}% {str(key):32s} {text(key):80s} {[(k, int(v)) for k, v in st]}")
