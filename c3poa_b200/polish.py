"""Batched polishing hand-off.

The reference polishes read by read: per read it writes three temporary files, builds a minimap2 index of one
consensus with mappy, maps the read's subreads to it, and starts one racon process
(/root/reference/bin/determine_consensus.py:49-104).  Here a whole batch goes to ONE racon process:

    targets    <tmp>/batch_abpoa.fasta      >{name}\\n{pre-polish consensus}            (:49-53)
    sequences  <tmp>/batch_subreads.fastq   @{name}_{k} subread records                (:57-77; c3_format_batch's text)
    overlaps   <tmp>/batch_overlaps.paf     one line per complete subread              (:63-66)

The overlaps need no aligner: every complete subread was aligned end to end to its read's graph by the POA stage, so
it is reported over its whole length against the whole consensus (racon realigns inside its windows; PAF carries no
CIGAR either way).  Dangling ends are left out -- the reference places them with mappy, which is not available here
-- so polished sequences are NOT claimed to equal the reference's; the acceptance test is plumbing only (file
formats, one process per batch, names round trip, unpolished targets kept).  racon itself stays external (north_star).
"""
from __future__ import annotations

import os
import subprocess

import numpy as np


def write_polish_inputs(tmp_dir, names, out, off, fastq_text, tag="batch"):
    """Writes the three racon inputs for the status-0 reads of a batch.  fastq_text: the subread FASTQ records
    c3_format_batch produced for these reads.  Returns (sequences, overlaps, targets) paths and the read indices."""
    R = out["results"]
    idx = np.flatnonzero(R["status"] == 0)
    seq_path = os.path.join(tmp_dir, f"{tag}_subreads.fastq")
    paf_path = os.path.join(tmp_dir, f"{tag}_overlaps.paf")
    tgt_path = os.path.join(tmp_dir, f"{tag}_abpoa.fasta")
    with open(seq_path, "wb") as fh:
        fh.write(fastq_text)
    sb = out["sub_bounds"]
    with open(tgt_path, "w") as ft, open(paf_path, "w") as fp:
        for i in idx:
            name, cl = names[i], int(R["cons_len"][i])
            ft.write(f">{name}\n{out['cons'][i, :cl].tobytes().decode()}\n")
            for k in range(int(R["n_sub"][i])):
                ql = int(sb[i, k, 1] - sb[i, k, 0])
                fp.write(f"{name}_{k + 1}\t{ql}\t0\t{ql}\t+\t{name}\t{cl}\t0\t{cl}\t{min(ql, cl)}\t{max(ql, cl)}\t60\n")
    return seq_path, paf_path, tgt_path, idx


def run_racon(racon, seq_path, paf_path, tgt_path, threads=1, log_path=None):
    """One racon process for the whole batch (-u keeps targets racon did not polish).  Returns {name: sequence}."""
    cmd = [racon, seq_path, paf_path, tgt_path, "-q", "5", "-t", str(max(1, threads)), "-u"]
    with open(log_path or os.devnull, "w") as log:
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=log)
    if res.returncode != 0:
        # the reference tolerates a failing racon (bin/determine_consensus.py:92-99 reads whatever came out): a batch keeps
        # its pre-polish consensi rather than aborting the run after the GPU work
        import sys
        print(f"racon exited with {res.returncode} for {tgt_path}: the batch keeps its pre-polish consensi", file=sys.stderr)
        return {}
    polished, name = {}, None
    for line in res.stdout.decode().splitlines():
        if line.startswith(">"):
            name = line[1:].split()[0]
            polished[name] = []
        elif name is not None:
            polished[name].append(line.strip())
    return {k: "".join(v) for k, v in polished.items()}


def apply_polished(names, out, idx, polished):
    """Puts the polished sequences back into the batch outputs (cons, cons_len); a target racon did not return, or one
    that no longer fits the consensus buffer, keeps its pre-polish consensus.  Returns the number replaced."""
    cap = out["cons"].shape[1]
    n = 0
    for i in idx:
        s = polished.get(names[i])
        if not s or len(s) > cap:
            continue
        b = np.frombuffer(s.encode(), dtype=np.uint8)
        out["cons"][i, :b.size] = b
        out["results"]["cons_len"][i] = b.size
        n += 1
    return n


def polish_batch(racon, tmp_dir, names, out, off, fastq_text, threads=1, tag="batch", keep=False):
    seq_path, paf_path, tgt_path, idx = write_polish_inputs(tmp_dir, names, out, off, fastq_text, tag)
    try:
        if idx.size == 0:
            return 0
        polished = run_racon(racon, seq_path, paf_path, tgt_path, threads, os.path.join(tmp_dir, "racon_messages.log"))
        return apply_polished(names, out, idx, polished)
    finally:
        if not keep:
            for p in (seq_path, paf_path, tgt_path):
                if os.path.exists(p):
                    os.remove(p)
