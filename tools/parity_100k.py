"""north_star's parity target, run on the GPU box: all reads of the bench workload (BASELINE.json configs[1], 100 000
cfg2 reads, the bench's own seed) through the fused C-ABI call in auto mode, and every one of them through the oracle
on all host threads; compares status, peaks, subread / dangling bounds, consensus bytes, DP cell counts, graph sizes.
Writes gpurun_out/r2_parity_100k.json (copied to profiles/).   usage: python tools/parity_100k.py [reads] [config]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from c3poa_b200.api import GpuConsensus, ReadBatch  # noqa: E402
from oracle import pyoracle as O  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
cfg_name = sys.argv[2] if len(sys.argv) > 2 else "cfg2_1kb_x5"
cfg = bench.CONFIGS[cfg_name]
blob, off, sp_idx, splints = bench.make_workload(cfg_name, n, bench.SEED)
sp_off = np.zeros(len(splints) + 1, dtype=np.int32)
sp_off[1:] = np.cumsum([len(x) for x in splints])
b = ReadBatch(blob, off, np.frombuffer("".join(splints).encode(), dtype=np.uint8).copy(), sp_off, np.ascontiguousarray(sp_idx, dtype=np.int32))
gpu = GpuConsensus(0, poa_mode="auto")
t0 = time.perf_counter()
out = gpu.consensus_batch(b, max_peaks=cfg["max_peaks"], cons_cap=cfg["cons_cap"])
t_gpu = time.perf_counter() - t0
given, done = gpu.lane_counts()
cores = os.cpu_count() or 1
seqs = [blob[off[i]:off[i + 1]].tobytes().decode() for i in range(n)]
t0 = time.perf_counter()
r = O.consensus_batch(seqs, splints, sp_idx, n_threads=cores, max_peaks=cfg["max_peaks"], cons_cap=cfg["cons_cap"])
t_cpu = time.perf_counter() - t0
bad = bench.compare_with_oracle(out, r, n)
db = int((out["dang_bounds"] != r["dang_bounds"]).any(axis=(1, 2)).sum())
res = out["results"]
line = {"config": cfg_name, "reads": n, "seed": bench.SEED, "mismatching_reads": bad, "dangling_bounds_mismatches": db,
        "status_hist": {int(k): int(v) for k, v in zip(*np.unique(res["status"], return_counts=True))},
        "poa_cells": int(res["poa_cells"].sum()), "group_kernel_given": given, "group_kernel_done": done,
        "gpu_first_call_s": t_gpu, "oracle_s": t_cpu, "oracle_threads": cores, "oracle": "restated (oracle/c3poa_oracle.c)",
        "fields": "status, n_peaks, peaks, n_sub, subread bounds, n_dang, dangling bounds, cons_len, consensus bytes / MSA rows, "
                  "poa_cells, poa_nodes"}
print(json.dumps(line))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(line, open(os.path.join(ROOT, "gpurun_out", f"r2_parity_{cfg_name}_{n}.json"), "w"), indent=1)
sys.exit(1 if bad or db else 0)
