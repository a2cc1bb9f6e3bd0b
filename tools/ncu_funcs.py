#!/usr/bin/env python3
"""Aggregate an ncu report's per-SASS-instruction counters by the source FUNCTION the line belongs to
(functions of poa_grp.cuh by their definition ranges; other files by file name).
usage: tools/ncu_funcs.py report.ncu-rep kernel_substring [lib.so]"""
import csv, io, os, re, subprocess, sys, tempfile
from collections import defaultdict

rep, sym = sys.argv[1], sys.argv[2]
so = os.path.abspath(sys.argv[3] if len(sys.argv) > 3 else "c3poa_b200/libc3poa_gpu.so")
SRCFILE = os.environ.get("NCU_SRC", "poa_grp.cuh")
src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "c3poa_b200", "csrc", SRCFILE)
starts = []
for i, ln in enumerate(open(src), 1):
    m = re.match(r"^(?:C3G_FN|C3_HD __forceinline__|C3_HD inline|__global__)\s+.*?\b(c3[gs]?_\w+)\s*\(", ln)
    if m:
        starts.append((i, m.group(1)))
def func_of(f, l):
    if f != SRCFILE:
        return f
    name = SRCFILE + "(top)"
    for s, n in starts:
        if s <= l: name = n
    return name
out = subprocess.run(["ncu", "-i", os.path.abspath(rep), "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = dict(name=r[1], rows=[]); blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
plain = re.sub(r"^_Z\d+", "", sym)
blk = next(b for b in blocks if plain in b["name"])
hdr, data = blk["rows"][0], blk["rows"][1:]
ci = {h: i for i, h in enumerate(hdr)}
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=td, capture_output=True)
    cub = max((f for f in os.listdir(td) if f.endswith(".cubin")), key=lambda f: os.path.getsize(os.path.join(td, f)))
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cub)], capture_output=True, text=True).stdout
lines, in_fn, cur_line, stack = [], False, None, None
for ln in dis.splitlines():
    if ln.startswith("\t.section\t.text."):
        in_fn = sym in ln; continue
    if not in_fn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
        # inlined-at chain: attribute library lines to the innermost poa_grp.cuh line that inlined them
        m2 = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
        stack = [(os.path.basename(a), int(b)) for a, b in m2]
        continue
    if re.search(r"/\*[0-9a-f]{4,}\*/", ln) and not ln.strip().startswith("//"):
        key = cur_line
        if key and key[0] != SRCFILE and stack:
            for s in stack:
                if s[0] == SRCFILE: key = s; break
        lines.append(key)
agg = defaultdict(lambda: [0, 0, 0])
for k in range(min(len(data), len(lines))):
    r = data[k]
    f, l = lines[k] if lines[k] else ("?", 0)
    a = agg[func_of(f, l)]
    a[0] += int(r[ci["Instructions Executed"]] or 0); a[1] += int(r[ci["Thread Instructions Executed"]] or 0)
    a[2] += int(r[ci["Warp Stall Sampling (All Samples)"]] or 0)
ti = sum(a[0] for a in agg.values()); ts = sum(a[2] for a in agg.values())
print(f"# {blk['name']}: total warp-instructions {ti:,}  stall samples {ts:,}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:28s} inst {100*a[0]/max(ti,1):6.2f}%  thr/inst {a[1]/max(a[0],1):5.1f}  stall {100*a[2]/max(ts,1):6.2f}%")
