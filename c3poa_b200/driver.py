"""C3POa-compatible driver over the GPU hot path.

Keeps the reference's CLI flags, config file, c3poa.log, header naming
(>readName_avgQ_len_repeats_consLen) and Splint_N/ directory layout
(/root/reference/C3POa.py:26-84,167-173,175-272), but replaces the per-read loop of
analyze_reads (C3POa.py:110-165) by batch calls into the C ABI.  BLAT (splint assignment) and
racon (polishing) remain the reference's external steps: the PSL is read if it exists
(bin/preprocess.py:17), else `blat` is run; consensi written here are the PRE-polish abPOA
consensi unless --polish is given and racon + mappy are available.
"""
from __future__ import annotations

import argparse
import gzip
import os
import shutil
import subprocess
import sys

import numpy as np

from .api import GpuConsensus, ReadBatch, pairwise_rows
from .fastx import fastx_read, revcomp
from .ingest import FastqBatches
from .pairwise import pairwise_consensus  # noqa: F401  (2-repeat path, host side)

VERSION = "v2.2.3-b200"


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="Makes consensus sequences from R2C2 reads (GPU hot path).")
    p.add_argument("--reads", "-r", type=str)
    p.add_argument("--splint_file", "-s", type=str)
    p.add_argument("--out_path", "-o", type=str, default=os.getcwd())
    p.add_argument("--config", "-c", type=str, default="")
    p.add_argument("--lencutoff", "-l", type=int, default=1000)
    p.add_argument("--mdistcutoff", "-d", type=int, default=500)
    p.add_argument("--zero", "-z", action="store_false", default=True)
    p.add_argument("--numThreads", "-n", type=int, default=1)
    p.add_argument("--groupSize", "-g", type=int, default=1000)
    p.add_argument("--blatThreads", "-b", action="store_true", default=False)
    p.add_argument("--compress_output", "-co", action="store_true", default=False)
    p.add_argument("--version", "-v", action="version", version=VERSION)
    p.add_argument("--device", type=int, default=0, help="CUDA device ordinal")
    p.add_argument("--batch", type=int, default=50000, help="reads per GPU batch")
    p.add_argument("--gpus", type=int, default=1, help="GPUs of this node to shard the reads over (one process each)")
    p.add_argument("--inflight", type=int, default=2, help="batches in flight per GPU (handles + host threads)")
    p.add_argument("--polish", action="store_true", default=False,
                   help="polish each batch's consensi with ONE racon process (racon from the config file / PATH); the "
                        "default output is the pre-polish abPOA consensus")
    p.add_argument("--assign", choices=["psl", "gpu"], default="psl",
                   help="splint/strand per read: 'psl' = BLAT PSL as the reference (default); 'gpu' = conk profile of "
                        "every splint x strand on the GPU (no BLAT; not BLAT-equivalent)")
    p.add_argument("--assign_min_frac", type=float, default=0.05,
                   help="--assign gpu: accept the best splint if its profile maximum reaches this fraction of a perfect match")
    return p.parse_args(argv)


def config_reader(config_in):
    progs = {}
    with open(config_in) as f:
        for line in f:
            if line.startswith("#") or not line.rstrip().split():
                continue
            k, v = line.rstrip().split("\t")[:2]
            progs[k] = v
    for missing in {"racon", "blat"} - set(progs):
        progs[missing] = missing
        sys.stderr.write(f"Using {missing} from your path, not the config file.\n")
    return progs


def read_psl(align_psl, names):
    """bin/preprocess.py:22-45: best splint + strand per read (gaps < 50 and score > 50)."""
    cand = {n: [(None, 1.0, None)] for n in names}
    adapter_set = set()
    with open(align_psl) as f:
        for line in f:
            c = line.rstrip().split("\t")
            if len(c) < 14:
                continue
            read_name, adapter, strand = c[9], c[13], c[8]
            gaps, score = float(c[5]), float(c[0])
            if gaps < 50 and score > 50 and read_name in cand:
                cand[read_name].append((adapter, score, strand))
                adapter_set.add(adapter)
    adapter_dict, no_splint = {}, 0
    for name, al in cand.items():
        best = sorted(al, key=lambda x: x[1], reverse=True)[0]
        if not best[0]:
            no_splint += 1
            continue
        adapter_dict[name] = (best[0], best[2])
    return adapter_dict, adapter_set, no_splint


def run_blat(blat, reads_fastq, splint_file, tmp_dir, lencutoff):
    fa = os.path.join(tmp_dir, "R2C2_temp_for_BLAT.fasta")
    with open(fa, "w") as f:
        for name, seq, _ in fastx_read(reads_fastq):
            if len(seq) >= lencutoff:
                f.write(f">{name}\n{seq}\n")
    psl = os.path.join(tmp_dir, "splint_to_read_alignments.psl")
    with open(os.path.join(tmp_dir, "blat_messages.log"), "w") as log:
        # the reference's flags, bin/preprocess.py:71-75
        subprocess.run([blat, "-noHead", "-stepSize=1", "-t=DNA", "-q=DNA", "-minScore=15",
                        "-minIdentity=10", splint_file, fa, psl], stdout=log, stderr=log, check=True)
    os.remove(fa)
    return psl


def header(name, qual, seq_len, repeats, cons_len):
    """C3POa.py:168-171.  qual: the quality string, or its Phred sum (int) as the C++ ingest returns it."""
    qsum = qual if isinstance(qual, (int, np.integer)) else sum(ord(x) - 33 for x in qual)
    avg_qual = round(int(qsum) / seq_len, 2)
    return ">" + name + "_" + "_".join(str(x) for x in (avg_qual, seq_len, repeats, cons_len))


def compute_batch(gpu, names, blob, off, splint_dict, adapter_dict, mdist):
    """The GPU part of one batch (packed arrays; every read has a splint): returns the fused-path outputs."""
    sp_names = sorted(splint_dict)
    splints = []
    for n in sp_names:
        splints += splint_dict[n]
    idx = np.array([2 * sp_names.index(adapter_dict[n][0]) + (1 if adapter_dict[n][1] == "-" else 0) for n in names],
                   dtype=np.int32)
    sp_blob = np.frombuffer("".join(splints).encode(), dtype=np.uint8).copy()
    sp_off = np.zeros(len(splints) + 1, dtype=np.int32)
    sp_off[1:] = np.cumsum([len(s) for s in splints])
    batch = ReadBatch(np.ascontiguousarray(blob), np.ascontiguousarray(off, dtype=np.int64), sp_blob, sp_off, idx)
    max_len = int(np.diff(off).max())
    return gpu.consensus_batch(batch, min_dist=mdist, max_peaks=128, cons_cap=min(2 * max_len, 131072))


def write_batch(out, names, blob, off, qual, qual_sum, adapter_dict, handles, polish=None, zero=True):
    """The host part: consensus FASTA + subread FASTQ exactly as analyze_reads / determine_consensus write them.
    Reads with a plain consensus (status 0) are formatted by the library in one pass per splint directory
    (c3_format_batch); only the 2-repeat reads, whose quality-aware pairwise consensus is host-side Python
    (bin/consensus.py), are still handled one by one."""
    from .ingest import format_batch, pack_names
    R = out["results"]
    stats = dict(consensus=0, no_peaks=0, pairwise=0, errors=0)     # pairwise: subset of consensus (2 repeats)
    order = sorted(handles)
    gidx = {a: k for k, a in enumerate(order)}
    group = np.fromiter((gidx[adapter_dict[nm][0]] for nm in names), dtype=np.int32, count=len(names))
    names_raw, name_off = pack_names(names)
    if polish is not None:
        # f-4: one racon process per batch on (all subreads, whole-length overlaps, all pre-polish consensi); the
        # consensus records are then formatted from the polished sequences
        from .polish import polish_batch
        _, fq_all, _ = format_batch(out, names_raw, name_off, blob, qual, off, qual_sum)
        stats["polished"] = polish_batch(polish["racon"], polish["tmp_dir"], names, out, off, fq_all.tobytes(),
                                         threads=polish.get("threads", 1), tag=f"batch{polish.get('serial', 0)}")
    for k, adapter in enumerate(order):
        fa, fq, st = format_batch(out, names_raw, name_off, blob, qual, off, qual_sum, group, k)
        cons_fh, sub_fh = handles[adapter]
        cons_fh.write(fa); sub_fh.write(fq)
        stats["consensus"] += st["consensus"]; stats["no_peaks"] += st["no_peaks"]; stats["errors"] += st["errors"]
    for i in np.flatnonzero(R["status"] == 2):
        name = names[i]
        cons_fh, sub_fh = handles[adapter_dict[name][0]]
        ns, nd = int(R["n_sub"][i]), int(R["n_dang"][i])
        a0, a1 = int(off[i]), int(off[i + 1])
        if not (ns == 2 and R["cons_len"][i] > 0):
            # 0-repeat read (one peak, or every subread an outlier).  The reference's zero_repeats
            # (bin/determine_consensus.py:14-18,107-139; --zero, default on) first writes the two dangling halves as
            # @name_0 / @name_1 and then overlaps them with mappy, which is absent here: the records are written, the
            # consensus is not produced, and the count goes to c3poa.log -- never silent.
            stats["zero"] = stats.get("zero", 0) + 1
            if zero and ns == 0 and nd == 2:
                seq = blob[a0:a1].tobytes().decode()
                q = qual[a0:a1].tobytes().decode()
                db = out["dang_bounds"][i, :2]
                sub_fh.write("".join(f"@{name}_{k}\n{seq[a:b]}\n+\n{q[a:b]}\n" for k, (a, b) in enumerate(db)).encode())
                stats["zero_records"] = stats.get("zero_records", 0) + 1
            continue
        seq = blob[a0:a1].tobytes().decode()
        q = qual[a0:a1].tobytes().decode()
        sb, db = out["sub_bounds"][i, :ns], out["dang_bounds"][i, :nd]
        # 2-repeat path: abPOA pairwise MSA rows from the GPU + quality-aware consensus
        # (bin/determine_consensus.py:33-41, bin/consensus.py)
        rows = pairwise_rows(out, i)
        cons = pairwise_consensus(rows, [seq[a:b] for a, b in sb], [q[a:b] for a, b in sb])
        cons_fh.write((header(name, int(qual_sum[i]), len(seq), ns, len(cons)) + "\n" + cons + "\n").encode())
        # subreads: @name_1..n, dangling @name_0 / @name_{n+1}  (bin/determine_consensus.py:57-77)
        recs = [f"@{name}_{k + 1}\n{seq[a:b]}\n+\n{q[a:b]}\n" for k, (a, b) in enumerate(sb)]
        recs += [f"@{name}_{0 if k == 0 else ns + 1}\n{seq[a:b]}\n+\n{q[a:b]}\n" for k, (a, b) in enumerate(db)]
        sub_fh.write("".join(recs).encode())
        stats["pairwise"] += 1
        stats["consensus"] += 1
    return stats


def main(args):
    if not args.out_path.endswith("/"):
        args.out_path += "/"
    os.makedirs(args.out_path, exist_ok=True)
    progs = config_reader(args.config) if args.config else {"racon": "racon", "blat": "blat"}
    tmp_dir = args.out_path + "tmp/"
    os.makedirs(tmp_dir, exist_ok=True)
    names = []
    scan = FastqBatches(args.reads, min_len=args.lencutoff, max_reads=args.batch, pinned=False)
    for b in scan:                              # pass 1: names of the reads >= --lencutoff (C3POa.py:200-206)
        names += b["names"]
    short_reads = int(scan.n_short.value)
    scan.close()
    splint_dict = {n: [s, revcomp(s)] for n, s, _ in fastx_read(args.splint_file)}
    if args.assign == "gpu":
        adapter_dict, adapter_set, no_splint = None, set(splint_dict), 0       # assigned per batch on the GPU
    else:
        align_psl = tmp_dir + "splint_to_read_alignments.psl"
        if not os.path.exists(align_psl) or os.stat(align_psl).st_size == 0:
            print("Aligning splints to reads with blat", file=sys.stderr)
            run_blat(progs["blat"], args.reads, args.splint_file, tmp_dir, args.lencutoff)
        else:
            print("Reading existing psl file", file=sys.stderr)
        adapter_dict, adapter_set, no_splint = read_psl(align_psl, names)
    all_reads = len(names) + short_reads
    totals = _run_all(args, adapter_dict, splint_dict, adapter_set)
    no_splint += totals.pop("no_splint", 0)
    with open(args.out_path + "c3poa.log", "w") as log:     # C3POa.py:214-229
        print("C3POa version:", VERSION, file=log)
        print("Total reads:", all_reads, file=log)
        print("No splint reads:", no_splint, "({:.2f}%)".format(no_splint / max(all_reads, 1) * 100), file=log)
        print("Under len cutoff:", short_reads, "({:.2f}%)".format(short_reads / max(all_reads, 1) * 100), file=log)
        print("Total thrown away reads:", short_reads + no_splint,
              "({:.2f}%)".format((short_reads + no_splint) / max(all_reads, 1) * 100), file=log)
        print("Reads after preprocessing:", all_reads - (short_reads + no_splint), file=log)
        if getattr(args, "polish", False):
            print("Polishing: one racon process per batch on the >= 3-repeat consensi (2-repeat pairwise consensi and the",
                  "dangling ends are not polished); polished:", totals.get("polished", 0), file=log)
        else:
            print("Polishing: off (pre-polish abPOA consensi; the reference always runs racon -- use --polish)", file=log)
        # not in the reference's log: what this driver does not produce must not disappear silently
        print("Zero-repeat reads without consensus (zero_repeats needs mappy; dangling subreads written:",
              str(totals.get("zero_records", 0)) + "):", totals.get("zero", 0), file=log)
    print("GPU consensus:", totals, file=sys.stderr)
    return totals


def _run_all(args, adapter_dict, splint_dict, adapter_set):
    for adapter in adapter_set:
        os.makedirs(args.out_path + adapter, exist_ok=True)
    if args.gpus <= 1:
        return _consume(args, args.device, 0, 1, adapter_dict, splint_dict, adapter_set, final=True)
    # one process per GPU (spawn, like the reference's pool: C3POa.py:236,279); reads are sharded by index, every
    # rank writes <splint>/tmp<rank>/ and the parent concatenates (cat_files, C3POa.py:259-271)
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(args.gpus) as pool:
        parts = pool.starmap(_consume, [(args, r, r, args.gpus, adapter_dict, splint_dict, adapter_set, False)
                                        for r in range(args.gpus)])
    totals = {k: sum(p.get(k, 0) for p in parts) for k in set().union(*parts)}
    for adapter in adapter_set:
        base = args.out_path + adapter
        for fn in ("R2C2_Consensus.fasta", "R2C2_Subreads.fastq"):
            with _opener(args)(base + "/" + fn) as fh:
                for r in range(args.gpus):
                    part = f"{base}/tmp{r}/{fn}"
                    if os.path.exists(part):
                        with open(part, "rb") as src:
                            shutil.copyfileobj(src, fh)
        for r in range(args.gpus):
            shutil.rmtree(f"{base}/tmp{r}", ignore_errors=True)
    return totals


def _drain_one(pending, handles, totals, polish=None, zero=True):
    fut, names, blob, off, qual, qsum, ad = pending.pop(0)
    out = fut.result()
    if polish is not None:
        polish = dict(polish, serial=totals.get("batches", 0))
    totals["batches"] = totals.get("batches", 0) + 1
    for key, v in write_batch(out, names, blob, off, qual, qsum, ad, handles, polish, zero=zero).items():
        totals[key] = totals.get(key, 0) + v


def _opener(args):
    return (lambda p: gzip.open(p + ".gz", "wb")) if args.compress_output else (lambda p: open(p, "wb"))


def _consume(args, device, rank, world, adapter_dict, splint_dict, adapter_set, final):
    """Process the reads with index % world == rank on `device`; final=True writes the final files directly."""
    handles = {}
    for adapter in adapter_set:
        base = args.out_path + adapter
        if final:
            op = _opener(args)
        else:
            base += f"/tmp{rank}"
            os.makedirs(base, exist_ok=True)
            op = lambda p: open(p, "wb")     # noqa: E731
        handles[adapter] = (op(base + "/R2C2_Consensus.fasta"), op(base + "/R2C2_Subreads.fastq"))
    mod = int(os.environ.get("C3POA_DEVICE_MODULO", "0"))      # tests: fold ranks onto fewer GPUs
    gpu = GpuConsensus(device % mod if mod > 0 else device)
    totals = dict(consensus=0, no_peaks=0, pairwise=0, errors=0)
    k = 0
    import queue
    from concurrent.futures import ThreadPoolExecutor
    n_inflight = max(1, int(getattr(args, "inflight", 2)))
    gpus = [gpu] + [GpuConsensus(gpu.device) for _ in range(n_inflight - 1)]
    free_gpus = queue.Queue()
    for g2 in gpus:
        free_gpus.put(g2)

    def run_on_free_handle(*a):
        g3 = free_gpus.get()                # a handle is not thread-safe: one batch per handle at a time
        try:
            return compute_batch(g3, *a)
        finally:
            free_gpus.put(g3)

    pool = ThreadPoolExecutor(max_workers=n_inflight)
    pending = []
    polish = None
    if getattr(args, "polish", False):         # f-4: one racon process per batch (bin/determine_consensus.py:49-104 per read)
        progs = config_reader(args.config) if args.config else {"racon": "racon"}
        pdir = args.out_path + f"tmp/polish{rank}/"
        os.makedirs(pdir, exist_ok=True)
        polish = dict(racon=progs["racon"], tmp_dir=pdir, threads=max(1, args.numThreads))
    # staging buffers are pinned (slow to allocate: ~0.2 s per GB): no larger than the input can fill
    fsz = os.path.getsize(args.reads)
    could_hold = fsz * (8 if args.reads.endswith(".gz") else 1) + (1 << 20)
    reader = FastqBatches(args.reads, min_len=args.lencutoff, max_reads=args.batch,
                          max_bases=min(1 << 29, max(1 << 22, min(args.batch * 12000, could_hold))), nbuf=n_inflight + 2)
    gpu_assign = adapter_dict is None
    assign_gpu = GpuConsensus(gpu.device) if gpu_assign else None
    if gpu_assign:
        sp_names = sorted(splint_dict)
        cands = [s for n in sp_names for s in splint_dict[n]]
        perfect = np.array([5 * len(s) * (len(s) + 1) // 2 for s in cands], dtype=np.float64)
    held = []                                   # reads held back for a batch of their own (see below)

    def submit(names, blob, off, qual, qsum, ad):
        fut = pool.submit(run_on_free_handle, names, blob, off, splint_dict, ad, args.mdistcutoff)
        pending.append((fut, names, blob, off, qual, qsum, ad))
        while len(pending) >= len(gpus) + 1 or (pending and pending[0][0].done()):
            _drain_one(pending, handles, totals, polish, zero=bool(args.zero))

    def flush_held():
        if not held:
            return
        held.sort(key=lambda t: len(t[1]))
        h_off = np.zeros(len(held) + 1, dtype=np.int64)
        h_off[1:] = np.cumsum([len(t[1]) for t in held])
        submit([t[0] for t in held], np.concatenate([t[1] for t in held]), h_off, np.concatenate([t[2] for t in held]),
               np.array([t[3] for t in held]), {t[0]: t[4] for t in held})
        totals["held_back_long_reads"] = totals.get("held_back_long_reads", 0) + len(held)
        held.clear()

    for b in reader:
        # keep the reads that have a splint and belong to this rank (round-robin over the kept reads)
        sel = []
        if gpu_assign:
            mine = [i for i in range(b["n"]) if (k + i) % world == rank]
            k += b["n"]
            adapter_dict = {}
            if mine:
                m_off = np.zeros(len(mine) + 1, dtype=np.int64)
                m_off[1:] = np.cumsum([b["off"][i + 1] - b["off"][i] for i in mine])
                m_blob = b["blob"] if len(mine) == b["n"] else np.concatenate([b["blob"][b["off"][i]:b["off"][i + 1]] for i in mine])
                best, scores = assign_gpu.assign_splints(m_blob, m_off, cands)
                for t, i in enumerate(mine):
                    c = int(best[t])
                    if scores[c, t] >= args.assign_min_frac * perfect[c]:
                        adapter_dict[b["names"][i]] = (sp_names[c // 2], "-" if c % 2 else "+")
                        sel.append(i)
                    else:
                        totals["no_splint"] = totals.get("no_splint", 0) + 1
        else:
            for i, name in enumerate(b["names"]):
                if name in adapter_dict:
                    if k % world == rank:
                        sel.append(i)
                    k += 1
        if not sel:
            continue
        names, off = b["names"], b["off"]
        if len(sel) == b["n"]:
            blob, qual, qsum = b["blob"], b["qual"], b["qual_sum"]
        else:
            idx = np.asarray(sel)
            lens = (off[idx + 1] - off[idx])
            blob = np.concatenate([b["blob"][off[i]:off[i + 1]] for i in sel])
            qual = np.concatenate([b["qual"][off[i]:off[i + 1]] for i in sel])
            qsum = b["qual_sum"][idx]
            names = [names[i] for i in sel]
            off = np.zeros(len(sel) + 1, dtype=np.int64)
            off[1:] = np.cumsum(lens)
        ad = dict(adapter_dict) if gpu_assign else adapter_dict
        # The consensus buffers of a batch are dense ([reads][cons_cap], cons_cap from the longest read): one 100-kb read
        # in a 50 000-read batch would cost gigabytes on host and device.  Reads far longer than the rest of their batch
        # are held back and run as batches of their own.
        lens = np.diff(off)
        long_cut = max(32768, 4 * int(np.median(lens)))
        is_long = lens > long_cut
        if is_long.any() and not is_long.all():
            for i in np.flatnonzero(is_long):
                a0, a1 = int(off[i]), int(off[i + 1])
                held.append((names[i], blob[a0:a1].copy(), qual[a0:a1].copy(), qsum[i], ad[names[i]]))
            keep = np.flatnonzero(~is_long)
            blob = np.concatenate([blob[off[i]:off[i + 1]] for i in keep])
            qual = np.concatenate([qual[off[i]:off[i + 1]] for i in keep])
            qsum = qsum[keep]
            names = [names[i] for i in keep]
            off = np.zeros(len(keep) + 1, dtype=np.int64)
            off[1:] = np.cumsum(lens[keep])
        # two batches in flight per GPU (one handle + host thread each): the next batch's copies and kernels
        # overlap the tail of the previous persistent POA grid and the Python-side output writing
        submit(names, blob, off, qual, qsum, ad)
        if len(held) >= 2048:
            flush_held()
    flush_held()
    while pending:
        _drain_one(pending, handles, totals, polish, zero=bool(args.zero))
    pool.shutdown()
    for g2 in gpus[1:]:
        g2.close()
    if assign_gpu is not None:
        assign_gpu.close()
    reader.close()
    for a, b2 in handles.values():
        a.close(); b2.close()
    gpu.close()
    return totals


def cli(argv=None):
    args = parse_args(argv)
    if not args.reads or not args.splint_file:
        print("Reads (--reads/-r) and splint (--splint_file/-s) are required", file=sys.stderr)
        return 1
    main(args)
    return 0
