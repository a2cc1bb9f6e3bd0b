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
    # BASELINE.json configs[0..4]; reads = default reads per GPU (cfg1: the reference's own 1,000-read case)
    "cfg1": dict(desc="1 kb insert, 3-5 repeats, single splint", inserts=[1000], repeats=(3, 5), n_splints=1, reads=1000,
                 max_peaks=16, cons_cap=2048, cpu_per_core=150),
    "cfg2_1kb_x5": dict(desc="1 kb insert, 5 repeats, single splint", inserts=[1000], repeats=(5, 5), n_splints=1,
                        reads=100000, max_peaks=16, cons_cap=2048, cpu_per_core=150),
    "cfg3": dict(desc="500 bp insert, 15-30 repeats (deep POA graphs), single splint", inserts=[500], repeats=(15, 30),
                 n_splints=1, reads=40000, max_peaks=64, cons_cap=1536, cpu_per_core=12),
    "cfg4": dict(desc="3-5 kb inserts in 20-50 kb concatemers, 2-4 repeats (wide bands, large graphs), single splint",
                 inserts=list(range(3000, 5001, 250)), repeats=(2, 4), n_splints=1, reads=20000, max_peaks=16,
                 cons_cap=12288, flank=(300, 2500), cpu_per_core=16),
    "cfg5": dict(desc="mixed inserts {500,1000,2000,4000}, 2-10 repeats, 4 demultiplexed splints", inserts=[500, 1000, 2000, 4000],
                 repeats=(2, 10), n_splints=4, reads=125000, max_peaks=16, cons_cap=10240, cpu_per_core=24),
}
CONFIGS["cfg2"] = CONFIGS["cfg2_1kb_x5"]
SEED = 20251017 + 2
OPS_PER_POA_CELL = 18      # SURVEY.md section 8(d), p = 1 predecessor
OPS_PER_CONK_CELL = 8


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--reads", type=int, default=0, help="reads per GPU per step (weak) / in total (strong); 0 = the config's default")
    ap.add_argument("--config", default="cfg2_1kb_x5", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --reads per GPU; strong: --reads in total, split over the ranks")
    ap.add_argument("--parity-reads", type=int, default=-1,
                    help="reads of rank 0's batch compared with the oracle (-1 = the cpu_baseline sample, 0 = none)")
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
    """(blob, off, sp_idx, splints in both orientations) of n_reads synthetic reads of the config."""
    c = CONFIGS[cfg]
    if len(c["inserts"]) == 1 and c["repeats"][0] == c["repeats"][1] and c["n_splints"] == 1:
        blob, off, strand = synth.make_batch(n_reads, insert_len=c["inserts"][0], repeats=c["repeats"][0], seed=seed)
        return blob, off, strand.astype(np.int32), [synth.SPLINT1, synth.revcomp(synth.SPLINT1)]
    rng = np.random.default_rng(4)
    sp = [synth.SPLINT1] + [synth.random_seq(rng, 284).tobytes().decode() for _ in range(c["n_splints"] - 1)]
    return synth.make_mixed_batch(n_reads, c["inserts"], c["repeats"], sp, seed=seed, flank=c.get("flank", (100, 400)))


def rank_share(total, world, rank):
    """Strong scaling: reads of rank `rank` when `total` reads are split over `world` ranks (shares differ by at most one)."""
    return total // world + (1 if rank < total % world else 0)


def oracle_kind():
    """"real" when the reference's own natives (pyabpoa 1.0.5 + conk, built by `make -C oracle ref`) are importable AND the
    restated oracle agrees with them on a probe; else "restated" (parity with upstream unpinned)."""
    ref = os.path.join(ROOT, "oracle", "_ref")
    if os.path.isdir(ref) and ref not in sys.path:
        sys.path.insert(1, ref)
    try:
        import conk
        import pyabpoa
    except Exception:
        return "restated"
    from oracle import pyoracle as O
    rng = np.random.default_rng(1)
    a = synth.random_seq(rng, 900)
    g = [synth.mutate(rng, a).tobytes().decode() for _ in range(5)]
    ok = O.poa_msa(g)["cons"] == pyabpoa.msa_aligner(match=5).msa(g, True, False).cons_seq[0]
    seq = synth.make_reads(1, insert_len=600, repeats=4, seed=3)["seqs"][0]
    ok = ok and np.array_equal(np.asarray(O.conk(synth.SPLINT1, seq, 20)), np.asarray(conk.conk(synth.SPLINT1, seq, 20)))
    return "real" if ok else "restated (DISAGREES with the importable pyabpoa/conk)"


def compare_with_oracle(out, r, n_check):
    """GPU outputs of the first n_check reads vs the oracle's (same reads, same parameters): status, peaks, subread and
    dangling bounds, consensus bytes (MSA rows for 2-repeat reads), DP cell counts, graph sizes.  Returns the number of
    reads with any difference."""
    g, o = out["results"], r["results"]
    bad = np.zeros(n_check, dtype=bool)
    for f in ("status", "n_peaks", "n_sub", "n_dang", "cons_len", "poa_nodes", "poa_cells"):
        bad |= g[f][:n_check] != o[f][:n_check]
    mp = min(out["peaks"].shape[1], r["peaks"].shape[1])
    kmask = np.arange(mp)[None, :] < np.minimum(g["n_peaks"][:n_check], mp)[:, None]
    bad |= ((out["peaks"][:n_check, :mp] != r["peaks"][:n_check, :mp]) & kmask).any(axis=1)
    smask = np.arange(mp)[None, :] < np.minimum(g["n_sub"][:n_check], mp)[:, None]
    bad |= ((out["sub_bounds"][:n_check, :mp] != r["sub_bounds"][:n_check, :mp]).any(axis=2) & smask).any(axis=1)
    cc = min(out["cons"].shape[1], r["cons"].shape[1])
    L = np.where(g["status"][:n_check] == 2, 2 * g["cons_len"][:n_check], g["cons_len"][:n_check])
    cmask = np.arange(cc)[None, :] < np.minimum(L, cc)[:, None]
    bad |= ((out["cons"][:n_check, :cc] != r["cons"][:n_check, :cc]) & cmask).any(axis=1)
    return int(bad.sum())


def cpu_reference_run(blob, off, sp_idx, splints, n_sample, threads, steps, warmup, max_peaks=16, cons_cap=4096):
    """Times the oracle port (conk + call_peaks + split + abPOA) on `n_sample` reads, all host threads."""
    from oracle import pyoracle as O
    n_sample = min(n_sample, off.size - 1)
    seqs = [blob[off[i]:off[i + 1]].tobytes().decode() for i in range(n_sample)]
    idx = sp_idx[:n_sample]
    for _ in range(warmup):
        O.consensus_batch(seqs[:max(threads, 8)], splints, idx[:max(threads, 8)], n_threads=threads, max_peaks=max_peaks,
                          cons_cap=cons_cap)
    t0 = time.perf_counter()
    cells = 0
    for _ in range(steps):
        r = O.consensus_batch(seqs, splints, idx, n_threads=threads, max_peaks=max_peaks, cons_cap=cons_cap)
        cells += int(r["results"]["poa_cells"].sum())
    dt = time.perf_counter() - t0
    ok = int((r["results"]["status"] == 0).sum())
    return dict(reads_per_s=n_sample * steps / dt, seconds=dt, n_sample=n_sample, poa_cells=cells, ok=ok, out=r)


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    cfg = CONFIGS[a.config]
    reads_arg = a.reads or cfg["reads"]
    if a.scaling == "strong":
        n_rank = rank_share(reads_arg, world, rank)
        wl_size = f"{reads_arg} synthetic R2C2 reads in total, split over {world} GPU(s)"
    else:
        n_rank = reads_arg
        wl_size = f"{reads_arg} synthetic R2C2 reads/GPU"
    wl_name = f"{a.config}: {wl_size}; {cfg['desc']}; 284-nt splints, 4/3/3 % sub/ins/del, both strands"
    max_peaks, cons_cap = cfg["max_peaks"], cfg["cons_cap"]

    # ------------------------------------------------------------------ reference arm
    if a.impl == "reference":
        if rank != 0:
            return 0
        n_sample = a.cpu_sample or min(reads_arg, cfg["cpu_per_core"] * cores)
        blob, off, sp_idx, splints = make_workload(a.config, n_sample, SEED)
        r = cpu_reference_run(blob, off, sp_idx, splints, n_sample, cores, a.steps, a.warmup, max_peaks, cons_cap)
        line = {
            "impl": "reference", "metric": "reads_to_consensus_per_sec", "value": r["reads_per_s"], "unit": "reads/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * r["seconds"] / a.steps,
            "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": wl_name, "sample_reads_per_step": r["n_sample"]},
            "cpu_baseline": {"value": r["reads_per_s"], "unit": "reads/s", "cores": cores, "kind": "port",
                             "sample": f"{r['n_sample']} reads of the workload per step, {cores} threads; oracle port "
                                       "(conk/pyabpoa are not installable offline; scalar int32 DP, no SIMD)"},
            "oracle": oracle_kind(),
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
    blob, off, sp_idx, splints = make_workload(a.config, n_rank, SEED + 1000 * rank)
    n = off.size - 1
    # pinned host staging of the inputs (e2e copies come from here)
    pin_blob = PinnedArray(blob.shape, np.uint8); pin_blob.array[:] = blob
    pin_off = PinnedArray(off.shape, np.int64); pin_off.array[:] = off
    sp_join = "".join(splints).encode()
    sp_off = np.zeros(len(splints) + 1, dtype=np.int32)
    sp_off[1:] = np.cumsum([len(x) for x in splints])
    batch = ReadBatch(pin_blob.array, pin_off.array, np.frombuffer(sp_join, dtype=np.uint8).copy(), sp_off,
                      np.ascontiguousarray(sp_idx, dtype=np.int32))
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
    tkeys = ("encode_ms", "conk_ms", "peaks_ms", "split_ms", "poa_ms", "poa_dp_ms", "poa_graph_ms", "poa_warp_ms", "poa_lane_ms")
    dev_ms, stage_ms, launches, step_ms = 0.0, {k: 0.0 for k in tkeys}, 0, []
    dp_launches = graph_launches = 0
    t0 = time.perf_counter()
    for _ in range(a.steps):
        gpu.run(**kw)                      # returns after the stream is synchronised
        t = gpu.timings()
        dev_ms += t["total_ms"]
        step_ms.append(t["total_ms"])
        launches += t["kernel_launches"]
        dp_launches += t["poa_dp_launches"]; graph_launches += t["poa_graph_launches"]
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
    sp_len = np.array([len(x) for x in splints], dtype=np.int64)
    conk_cells = int((np.diff(off) * sp_len[sp_idx]).sum())
    total_reads = allsum(float(n)) * a.steps
    value = total_reads / (dev_ms_max * 1e-3)

    # ---- end-to-end through the fused C-ABI entry point (c3_consensus_batch): pinned host -> device -> pinned host ----
    for _ in range(1):
        gpu.consensus_batch(batch, out=out, **kw)
    barrier()
    e2e_steps = []
    t0 = time.perf_counter()
    for _ in range(a.steps):
        t1 = time.perf_counter()
        gpu.consensus_batch(batch, out=out, **kw)
        e2e_steps.append(time.perf_counter() - t1)
    e2e_s = allmax(time.perf_counter() - t0)
    barrier()
    h2d = int(blob.nbytes + off.nbytes + len(sp_join) + sp_off.nbytes + batch.sp_idx.nbytes)
    d2h = int(sum(v.nbytes for v in out.values()))

    # ---- roofline of the dominant kernel (per-launch CUDA events on the launch stream, summed by kernel) ----
    int_peak = gpu.int_peak_ops()
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    if os.path.exists(peaks_file):
        try:
            hbm_peak = float(json.load(open(peaks_file))["hbm_gbs"]); hbm_src = "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    per_step = {k: v / a.steps for k, v in stage_ms.items()}
    kernels = {"c3_poa_grp_dp_kernel": per_step["poa_dp_ms"], "c3_poa_graph_kernel": per_step["poa_graph_ms"],
               "c3_poa_kernel": per_step["poa_warp_ms"], "c3_poa_lane_kernel": per_step["poa_lane_ms"],
               "c3_conk_kernel": per_step["conk_ms"], "c3_peaks_kernel": per_step["peaks_ms"]}
    dominant = max(kernels, key=kernels.get)
    k_s = kernels[dominant] * 1e-3
    sb = out["sub_bounds"]
    ns = res["n_sub"]
    in_poa = ns >= 3
    sub_bases = int(((sb[:, :, 1] - sb[:, :, 0]) * (np.arange(max_peaks)[None, :] < ns[:, None]))[in_poa].sum())
    fast_share = lane_done / max(1, int((ns >= 2).sum()))
    if dominant.startswith("c3_poa"):
        # 2-bit bases in + consensus out + 1 B/cell backtrack written and read once.  When the group path ran,
        # its DP kernel computes every cell and writes the backtrack bytes; the graph kernel reads them.
        alg_bytes = sub_bases / 4 + int(res["cons_len"][in_poa].sum()) + 2 * poa_cells
        if dominant == "c3_poa_grp_dp_kernel":
            alg_bytes = sub_bases / 4 + poa_cells
        int_ops = OPS_PER_POA_CELL * poa_cells if dominant != "c3_poa_graph_kernel" else 0
        if dominant in ("c3_poa_grp_dp_kernel", "c3_poa_lane_kernel"):
            int_ops *= fast_share                                  # the cells of the reads this kernel finished
    elif dominant == "c3_conk_kernel":
        alg_bytes = int(blob.nbytes) / 4 + 4 * int(blob.nbytes)
        int_ops = OPS_PER_CONK_CELL * conk_cells
    else:
        alg_bytes = 4 * int(blob.nbytes) + 4 * max_peaks * n
        int_ops = 0
    achieved = alg_bytes / k_s / 1e9 if k_s > 0 else 0.0
    traffic, traffic_source = None, None
    tf = os.path.join(ROOT, "profiles", "r02_traffic.json")      # dram bytes per launch from `ncu` captures, by kernel
    if os.path.exists(tf):
        try:
            t = json.load(open(tf)).get(dominant)
            if t:      # dram__bytes_read.sum + dram__bytes_write.sum over all launches of the kernel in one step
                traffic = (t["dram_bytes_read"] + t["dram_bytes_write"]) * (n / t["reads"])
                traffic_source = f"static ncu capture ({t.get('capture', '?')}), scaled by reads; not re-measured by this run"
        except Exception:
            traffic = None
    n_launch = {"c3_poa_grp_dp_kernel": dp_launches, "c3_poa_graph_kernel": graph_launches}.get(dominant, a.steps) / a.steps
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic, "traffic_source": traffic_source,
                "algorithmic_bytes": alg_bytes, "peak_source": hbm_src,
                "kernel_ms_per_step": kernels[dominant], "launches_per_step": n_launch,
                "avg_launch_ms": kernels[dominant] / max(n_launch, 1),
                "note": "achieved/algorithmic_bytes/traffic are per step (= all launches of the kernel in one pass over the batch); "
                        "integer max-plus DP: the binding resource is the issue/ALU rate, see roofline_int"}
    roofline_int = {"kernel": dominant, "achieved_ops_per_s": int_ops / k_s if k_s > 0 else 0.0,
                    "peak_ops_per_s": int_peak, "frac": (int_ops / k_s / int_peak) if (k_s > 0 and int_peak > 0) else None,
                    "ops_per_cell": OPS_PER_POA_CELL if dominant != "c3_conk_kernel" else OPS_PER_CONK_CELL,
                    "poa_stage_frac": (OPS_PER_POA_CELL * poa_cells / (per_step["poa_ms"] * 1e-3) / int_peak)
                    if per_step["poa_ms"] > 0 and int_peak > 0 else None,
                    "peak_source": "measured live: independent VIADDMNMX chains on all SMs (c3_measure_int_peak)"}

    # ---- CPU baseline on this box's host cores + parity of the GPU outputs against it (rank 0, N=1 only) ----
    cpu_baseline, parity = None, None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        n_sample = a.cpu_sample or min(n, cfg["cpu_per_core"] * cores)
        r = cpu_reference_run(blob, off, sp_idx, splints, n_sample, cores, 1, 1, max_peaks, cons_cap)
        cpu_baseline = {"value": r["reads_per_s"], "unit": "reads/s", "cores": cores, "kind": "port",
                        "sample": f"first {r['n_sample']} reads of the workload, {cores} threads, 1 pass "
                                  f"({r['seconds']:.1f} s); oracle port (scalar int32 DP; conk/pyabpoa not installable offline)"}
        n_check = r["n_sample"] if a.parity_reads < 0 else min(a.parity_reads, r["n_sample"])
        if n_check > 0:
            parity = {"checked_reads": n_check, "mismatches": compare_with_oracle(out, r["out"], n_check), "oracle": oracle_kind(),
                      "fields": "status, peaks, subread bounds, consensus bytes / MSA rows, DP cell counts, graph sizes"}

    if rank == 0:
        line = {
            "metric": "reads_to_consensus_per_sec", "value": value, "unit": "reads/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": dev_ms_max / a.steps, "higher_is_better": True,
            "scaling": a.scaling, "vs_baseline": None, "dtype": "int16 DP cells (int32 where the score bound needs it), fp64 smoothing",
            "data": "synthetic",
            "config": {"workload": wl_name, "reads_rank0": n, "parallelism": f"read-sharded x{world}, no collective",
                       "l2": f"inputs ({int(blob.nbytes) >> 20} MB/GPU/step) and the POA workspace exceed the 126 MB L2; no flush needed"},
            "wall_ms_per_step": 1e3 * wall_max / a.steps,
            "step_ms_rank0": {"median": statistics.median(step_ms), "min": min(step_ms), "max": max(step_ms)},
            "stage_ms_per_step": per_step,
            "poa_gcups": poa_cells / (per_step["poa_ms"] * 1e-3) / 1e9 * world if per_step["poa_ms"] > 0 else None,
            "conk_gcups": conk_cells / (per_step["conk_ms"] * 1e-3) / 1e9 * world if per_step["conk_ms"] > 0 else None,
            "reads_ok_rank0": n_ok, "reads_err_rank0": n_err,
            "poa_kernel": {"mode": a.poa_mode, "reads_in_poa_rank0": int((ns >= 2).sum()), "group_kernel_given_rank0": lane_given,
                           "group_kernel_done_rank0": lane_done, "warp_kernel_reads_rank0": int((ns >= 2).sum()) - lane_done},
            "e2e": {"value": total_reads / e2e_s, "unit": "reads/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "entry_point": "c3_consensus_batch",
                    "step_ms_rank0": {"median": 1e3 * statistics.median(e2e_steps), "min": 1e3 * min(e2e_steps), "max": 1e3 * max(e2e_steps)}},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roofline, "roofline_int": roofline_int,
            "cpu_baseline": cpu_baseline, "parity": parity,
        }
        print(json.dumps(line))
    grp.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
