#!/usr/bin/env python3
"""Aggregate an ncu report's per-SASS-instruction counters by CUDA source line.

usage: tools/ncu_lines.py report.ncu-rep kernel_mangled_substring [lib.so] [topN]
Joins `ncu --page source --csv` (SASS order) with `nvdisasm -g` line info of the cubin in lib.so.
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def main():
    rep, sym = sys.argv[1], sys.argv[2]
    so = sys.argv[3] if len(sys.argv) > 3 else "c3poa_b200/libc3poa_gpu.so"
    topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    so = os.path.abspath(so)
    rep = os.path.abspath(rep)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # possibly several kernels: pick the block whose first row names the kernel
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = dict(name=r[1], rows=[]); blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    plain = re.sub(r"^_Z\d+", "", sym)
    kname = os.environ.get("NCU_KERNEL", plain)          # name as the report prints it (templates are demangled there)
    blk = next(b for b in blocks if kname in b["name"])
    hdr, data = blk["rows"][0], blk["rows"][1:]
    ci = {h: i for i, h in enumerate(hdr)}
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=td, capture_output=True)
        cub = max((f for f in os.listdir(td) if f.endswith(".cubin")), key=lambda f: os.path.getsize(os.path.join(td, f)))
        dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cub)], capture_output=True, text=True).stdout
    lines, in_fn, cur_line = [], False, None
    for ln in dis.splitlines():
        if ln.startswith("\t.section\t.text."):
            in_fn = sym in ln
            continue
        if not in_fn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.search(r"/\*[0-9a-f]{4,}\*/", ln) and not ln.strip().startswith("//"):
            lines.append(cur_line)
    print(f"# {blk['name']}: {len(data)} SASS rows in report, {len(lines)} in cubin")
    agg = defaultdict(lambda: [0, 0, 0, 0])
    sort_col = 3 if os.environ.get('NCU_SORT') == 'sectors' else 0
    n = min(len(data), len(lines))
    for k in range(n):
        r = data[k]
        ie = int(r[ci["Instructions Executed"]] or 0)
        te = int(r[ci["Thread Instructions Executed"]] or 0)
        ss = int(r[ci["Warp Stall Sampling (All Samples)"]] or 0)
        a = agg[lines[k]]
        a[0] += ie; a[1] += te; a[2] += ss; a[3] += int(r[ci["L2 Theoretical Sectors Global"]] or 0)
    tot_i = sum(a[0] for a in agg.values()); tot_s = sum(a[2] for a in agg.values())
    print(f"# total warp-instructions {tot_i:,}  stall samples {tot_s:,}")
    tot_x = sum(a[3] for a in agg.values())
    print(f"# L2 theoretical sectors (global) {tot_x:,}")
    print("# file:line  warp_inst  %inst  avg_threads  stall_samples  %stall  l2_sectors  %sectors")
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][sort_col])[:topn]:
        print(f"{key[0] if key else '?'}:{key[1] if key else 0:<5d} {a[0]:>14,} {100*a[0]/max(tot_i,1):6.2f}% "
              f"{a[1]/max(a[0],1):6.1f} {a[2]:>9,} {100*a[2]/max(tot_s,1):6.2f}% {a[3]:>13,} {100*a[3]/max(tot_x,1):6.2f}%")


if __name__ == "__main__":
    main()
