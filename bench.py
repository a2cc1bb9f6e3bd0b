#!/usr/bin/env python3
"""bench.py -- reads->pre-polish-consensus throughput of the C3POa per-read hot path on B200.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one pass of the whole hot path (encode -> conk profile -> SG/call_peaks -> split ->
banded POA consensus) over one batch of synthetic R2C2 reads (BASELINE.json configs[1]: 1 kb
insert, 5 repeats, single splint, ~10 % errors, 100k reads per GPU).  `value` is measured with the
inputs resident in HBM (CUDA events around every run, max over ranks); `e2e` goes through the
C-ABI call with pinned HOST buffers, copies inside the timed region.  Reads are independent:
ranks shard by read with no collective (weak scaling: 100k reads per GPU).

`--impl reference` times the CPU path (oracle port of conk + call_peaks + split + abPOA; the real
conk / pyabpoa are not installable offline) on all host cores, on a bounded sample.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from c3poa_b200 import synth  # noqa: E402

CONFIGS = {
    # name: (insert_len, repeats, reads per GPU)
    "cfg2_1kb_x5": dict(insert_len=1000, repeats=5),
}
SEED = 20251017 + 2
OPS_PER_POA_CELL = 18      # SURVEY.md section 8(d), p = 1 predecessor
OPS_PER_CONK_CELL = 8


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--reads", type=int, default=100000, help="reads per GPU per step")
    ap.add_argument("--config", default="cfg2_1kb_x5", choices=sorted(CONFIGS))
    ap.add_argument("--cpu-sample", type=int, default=0, help="reads in the CPU-baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--poa-mode", default="auto", choices=["auto", "warp", "lane", "grp"],
                    help="POA kernel: auto = group kernel (8 lanes per read) with the warp-per-read kernel as fallback")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_workload(cfg, n_reads, seed):
    c = CONFIGS[cfg]
    blob, off, strand = synth.make_batch(n_reads, insert_len=c["insert_len"], repeats=c["repeats"], seed=seed)
    splints = [synth.SPLINT1, synth.revcomp(synth.SPLINT1)]
    return blob, off, strand.astype(np.int32), splints


def cpu_reference_run(blob, off, sp_idx, splints, n_sample, threads, steps, warmup):
    """Times the oracle port (conk + call_peaks + split + abPOA) on `n_sample` reads, all host threads."""
    from oracle import pyoracle as O
    n_sample = min(n_sample, off.size - 1)
    seqs = [blob[off[i]:off[i + 1]].tobytes().decode() for i in range(n_sample)]
    idx = sp_idx[:n_sample]
    for _ in range(warmup):
        O.consensus_batch(seqs[:max(threads, 8)], splints, idx[:max(threads, 8)], n_threads=threads)
    t0 = time.perf_counter()
    cells = 0
    for _ in range(steps):
        r = O.consensus_batch(seqs, splints, idx, n_threads=threads, cons_cap=4096)
        cells += int(r["results"]["poa_cells"].sum())
    dt = time.perf_counter() - t0
    ok = int((r["results"]["status"] == 0).sum())
    return dict(reads_per_s=n_sample * steps / dt, seconds=dt, n_sample=n_sample, poa_cells=cells, ok=ok)


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    wl_name = f"{a.config}: {a.reads} synthetic R2C2 reads/GPU, insert {CONFIGS[a.config]['insert_len']}, " \
              f"{CONFIGS[a.config]['repeats']} repeats, Splint1 (284 nt), 4/3/3 % sub/ins/del"

    # ------------------------------------------------------------------ reference arm
    if a.impl == "reference":
        if rank != 0:
            return 0
        n_sample = a.cpu_sample or min(a.reads, 150 * cores)
        blob, off, sp_idx, splints = make_workload(a.config, n_sample, SEED)
        r = cpu_reference_run(blob, off, sp_idx, splints, n_sample, cores, a.steps, a.warmup)
        line = {
            "impl": "reference", "metric": "reads_to_consensus_per_sec", "value": r["reads_per_s"], "unit": "reads/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * r["seconds"] / a.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": wl_name, "sample_reads_per_step": r["n_sample"]},
            "cpu_baseline": {"value": r["reads_per_s"], "unit": "reads/s", "cores": cores, "kind": "port",
                             "sample": f"{r['n_sample']} reads of the workload per step, {cores} threads; oracle port "
                                       "(conk/pyabpoa are not installable offline; scalar int32 DP, no SIMD)"},
            "e2e": {"value": r["reads_per_s"], "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "poa_gcups": r["poa_cells"] / r["seconds"] / 1e9,
        }
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ native arm
    from c3poa_b200.api import GpuConsensus, PinnedArray, ReadBatch, RESULT_DTYPE
    from c3poa_b200.dist import Group
    grp = Group("nccl")

    gpu = GpuConsensus(local_rank, poa_mode=a.poa_mode)
    blob, off, sp_idx, splints = make_workload(a.config, a.reads, SEED + 1000 * rank)
    n = off.size - 1
    # pinned host staging of the inputs (e2e copies come from here)
    pin_blob = PinnedArray(blob.shape, np.uint8); pin_blob.array[:] = blob
    pin_off = PinnedArray(off.shape, np.int64); pin_off.array[:] = off
    sp_join = "".join(splints).encode()
    batch = ReadBatch(pin_blob.array, pin_off.array, np.frombuffer(sp_join, dtype=np.uint8).copy(),
                      np.array([0, len(splints[0]), len(splints[0]) + len(splints[1])], dtype=np.int32), sp_idx)
    max_peaks, cons_cap = 16, 2048
    out_pin = dict(peaks=PinnedArray((n, max_peaks), np.int32), sub_bounds=PinnedArray((n, max_peaks, 2), np.int32),
                   dang_bounds=PinnedArray((n, 2, 2), np.int32), cons=PinnedArray((n, cons_cap), np.uint8),
                   results=PinnedArray((n,), RESULT_DTYPE))
    out = {k: v.array for k, v in out_pin.items()}
    kw = dict(max_peaks=max_peaks, cons_cap=cons_cap)

    barrier, allmax, allsum = grp.barrier, grp.allmax, grp.allsum

    # ---- device-resident measurement ----
    gpu.stage(batch)
    for _ in range(a.warmup):
        gpu.run(**kw)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    dev_ms, stage_ms, launches = 0.0, dict(encode_ms=0.0, conk_ms=0.0, peaks_ms=0.0, split_ms=0.0, poa_ms=0.0), 0
    t0 = time.perf_counter()
    for _ in range(a.steps):
        gpu.run(**kw)                      # returns after the stream is synchronised
        t = gpu.timings()
        dev_ms += t["total_ms"]
        launches += t["kernel_launches"]
        for k in stage_ms:
            stage_ms[k] += t[k]
    wall_s = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop()
    res = gpu.fetch(out)["results"]
    lane_given, lane_done = gpu.lane_counts()
    dev_ms_max = allmax(dev_ms)
    wall_max = allmax(wall_s)
    n_ok = int((res["status"] == 0).sum())
    n_err = int((res["status"] < 0).sum())
    poa_cells = int(res["poa_cells"].sum())
    conk_cells = int((np.diff(off) * len(splints[0])).sum())
    total_reads = allsum(float(n)) * a.steps
    value = total_reads / (dev_ms_max * 1e-3)

    # ---- end-to-end through the C ABI: pinned host -> device -> pinned host, every step ----
    for _ in range(1):
        gpu.consensus_batch(batch, out=out, **kw)
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        gpu.consensus_batch(batch, out=out, **kw)
    e2e_s = allmax(time.perf_counter() - t0)
    barrier()
    h2d = int(blob.nbytes + off.nbytes + len(sp_join) + 12 + sp_idx.nbytes)
    d2h = int(sum(v.nbytes for v in out.values()))

    # ---- roofline of the dominant kernel ----
    int_peak = gpu.int_peak_ops()
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    if os.path.exists(peaks_file):
        try:
            hbm_peak = float(json.load(open(peaks_file))["hbm_gbs"]); hbm_src = "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    poa_s = stage_ms["poa_ms"] * 1e-3 / a.steps
    conk_s = stage_ms["conk_ms"] * 1e-3 / a.steps
    poa_kernel_name = ("c3_poa_lane_kernel" if a.poa_mode == "lane" else "c3_poa_grp_kernel") if lane_done * 2 >= n else "c3_poa_kernel"
    dominant = poa_kernel_name if poa_s >= conk_s else "c3_conk_kernel"
    sb = out["sub_bounds"]
    ns = res["n_sub"]
    in_poa = ns >= 3
    sub_bases = int(((sb[:, :, 1] - sb[:, :, 0]) * (np.arange(max_peaks)[None, :] < ns[:, None]))[in_poa].sum())
    if dominant != "c3_conk_kernel":
        # 2-bit bases in + consensus out + 1 B/cell backtrack written and read once
        alg_bytes = sub_bases / 4 + int(res["cons_len"][in_poa].sum()) + 2 * poa_cells
        k_s, int_ops = poa_s, OPS_PER_POA_CELL * poa_cells
    else:
        alg_bytes = int(blob.nbytes) / 4 + 4 * int(blob.nbytes)
        k_s, int_ops = conk_s, OPS_PER_CONK_CELL * conk_cells
    achieved = alg_bytes / k_s / 1e9 if k_s > 0 else 0.0
    traffic = None
    tf = os.path.join(ROOT, "profiles", "r01_poa_lane_traffic.json" if dominant == "c3_poa_lane_kernel" else "r01_poa_traffic.json")
    if dominant != "c3_conk_kernel" and os.path.exists(tf):      # from one `ncu --set full` capture, per launch
        try:
            t = json.load(open(tf))
            traffic = (t["dram_bytes_read"] + t["dram_bytes_write"]) * (n / t["reads_per_launch"])
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic, "algorithmic_bytes": alg_bytes, "peak_source": hbm_src,
                "note": ("thread-per-read kernel: bound by memory latency at 12 warps/SM (ncu: issue slots 24 % busy, DRAM 29 % of "
                         "peak, long-scoreboard stalls); traffic is 9x the algorithmic bytes because H/E1/E2 are kept as int16 "
                         "per cell for the value-based backtrack" if dominant == "c3_poa_lane_kernel" else
                         "integer-ALU bound kernel: see roofline_int for the binding resource")}
    roofline_int = {"kernel": dominant, "achieved_ops_per_s": int_ops / k_s if k_s > 0 else 0.0,
                    "peak_ops_per_s": int_peak, "frac": (int_ops / k_s / int_peak) if (k_s > 0 and int_peak > 0) else None,
                    "ops_per_cell": OPS_PER_POA_CELL if dominant != "c3_conk_kernel" else OPS_PER_CONK_CELL,
                    "peak_source": "measured live: independent VIADDMNMX chains on all SMs (c3_measure_int_peak)"}

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only) ----
    cpu_baseline = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        n_sample = a.cpu_sample or min(n, 150 * cores)
        r = cpu_reference_run(blob, off, sp_idx, splints, n_sample, cores, 1, 1)
        cpu_baseline = {"value": r["reads_per_s"], "unit": "reads/s", "cores": cores, "kind": "port",
                        "sample": f"first {r['n_sample']} reads of the workload, {cores} threads, 1 pass "
                                  f"({r['seconds']:.1f} s); oracle port (scalar int32 DP; conk/pyabpoa not installable offline)"}

    if rank == 0:
        line = {
            "metric": "reads_to_consensus_per_sec", "value": value, "unit": "reads/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": dev_ms_max / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": wl_name, "reads_per_gpu": n, "parallelism": f"read-sharded x{world}, no collective",
                       "l2": "inputs (>700 MB/GPU/step) exceed the 126 MB L2; no flush needed"},
            "wall_ms_per_step": 1e3 * wall_max / a.steps,
            "stage_ms_per_step": {k: v / a.steps for k, v in stage_ms.items()},
            "poa_gcups": poa_cells / poa_s / 1e9 * world if poa_s > 0 else None,
            "conk_gcups": conk_cells / conk_s / 1e9 * world if conk_s > 0 else None,
            "reads_ok_rank0": n_ok, "reads_err_rank0": n_err,
            "poa_kernel": {"mode": a.poa_mode, "lane_reads_rank0": lane_given, "lane_done_rank0": lane_done},
            "e2e": {"value": total_reads / e2e_s, "unit": "reads/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roofline, "roofline_int": roofline_int,
            "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line))
    grp.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
