"""Seeded synthetic R2C2 read generator (SURVEY.md §8(d) read model).

A read is a window of the concatemer (insert + splint)^k with partial inserts at
both ends, i.i.d. errors (4 % substitution, 3 % insertion, 3 % deletion), on a
random strand.  ACGT only.  Also writes the PSL that lets the reference's
preprocess() skip BLAT (/root/reference/bin/preprocess.py:17,30-32).
"""
from __future__ import annotations

import numpy as np

SPLINT1 = (
    "TGAGGCTGATGAGTTCCATATTTGAAAAGTTTTCATCACTACTTAGTTTTTTGATAGCTTCAAGCCAGAGTTGTCTTTTTCTATCTACTCTCATACAACCAAT"
    "AAATGCTGAAATGAATTCTAAGCGGAGATCGCCTAGTGATTTTAAACTATTGCTGGCAGCATTCTTGAGTCCAATATAAAAGTATTGTGTACCTTTTGCTGGG"
    "TCAGGTTGTTCTTTAGGAGGAGTAAAAGGATCAAATGCACTAAACGAAACTGAAACAAGCGATCGAAAATATCCCTTT"
)  # Splint1, 284 nt: the sequence in the reference's splint.fasta
BASES = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTNacgtn", b"TGCANtgcan"):
    _COMP[_a] = _b


def revcomp(seq: str) -> str:
    a = np.frombuffer(seq.encode(), dtype=np.uint8)
    return _COMP[a][::-1].tobytes().decode()


def random_seq(rng: np.random.Generator, n: int) -> np.ndarray:
    return BASES[rng.integers(0, 4, size=n)]


def mutate(rng: np.random.Generator, seq: np.ndarray, sub=0.04, ins=0.03, dele=0.03) -> np.ndarray:
    """i.i.d. per-base errors; vectorised."""
    n = seq.size
    u = rng.random(n)
    keep = u >= dele
    is_sub = (u >= dele) & (u < dele + sub)
    out = seq.copy()
    ns = int(is_sub.sum())
    if ns:
        # substitute with a different base
        cur = np.searchsorted(BASES, out[is_sub])
        out[is_sub] = BASES[(cur + rng.integers(1, 4, size=ns)) % 4]
    is_ins = rng.random(n) < ins
    reps = keep.astype(np.int64) + is_ins.astype(np.int64)
    res = np.repeat(out, reps)
    # positions of inserted bases: the first copy where both keep and ins, or the only copy when deleted+ins
    starts = np.cumsum(reps) - reps
    ins_pos = starts[is_ins]
    res[ins_pos] = BASES[rng.integers(0, 4, size=ins_pos.size)]
    return res


def make_reads(n_reads: int, insert_len=1000, repeats=5, splints=None, seed=20251018,
               insert_choices=None, repeat_range=None, err=(0.04, 0.03, 0.03), flank=(100, 400),
               both_strands=True):
    """Returns dict with names, seqs (str), quals (str), splint_name, strand, truth inserts.

    repeats = number of complete insert copies (subreads); the read holds repeats+1 splints.
    """
    rng = np.random.default_rng(seed)
    if splints is None:
        splints = {"Splint1": SPLINT1}
    sp_names = list(splints)
    sp_arr = {k: np.frombuffer(v.encode(), dtype=np.uint8) for k, v in splints.items()}
    names, seqs, quals, spn, strands, truths = [], [], [], [], [], []
    for i in range(n_reads):
        il = int(rng.choice(insert_choices)) if insert_choices is not None else (
            int(rng.integers(insert_len[0], insert_len[1] + 1)) if isinstance(insert_len, tuple) else insert_len)
        k = int(rng.integers(repeat_range[0], repeat_range[1] + 1)) if repeat_range else repeats
        sname = sp_names[int(rng.integers(0, len(sp_names)))]
        sp = sp_arr[sname]
        ins = random_seq(rng, il)
        head = int(rng.integers(flank[0], min(flank[1], il) + 1))
        tail = int(rng.integers(flank[0], min(flank[1], il) + 1))
        parts = [ins[il - head:], sp]
        for _ in range(k):
            parts += [ins, sp]
        parts.append(ins[:tail])
        clean = np.concatenate(parts)
        noisy = mutate(rng, clean, *err)
        strand = "+"
        if both_strands and rng.random() < 0.5:
            noisy = _COMP[noisy][::-1]
            strand = "-"
        q = rng.integers(7, 21, size=noisy.size).astype(np.uint8) + 33
        names.append(f"r{i:07d}")
        seqs.append(noisy.tobytes().decode())
        quals.append(q.tobytes().decode())
        spn.append(sname)
        strands.append(strand)
        truths.append(ins.tobytes().decode())
    return dict(names=names, seqs=seqs, quals=quals, splint_name=spn, strand=strands, truth=truths,
                splints=dict(splints))


def write_fastq(path, names, seqs, quals):
    with open(path, "w") as f:
        for n, s, q in zip(names, seqs, quals):
            f.write(f"@{n}\n{s}\n+\n{q}\n")


def write_psl(path, names, splint_names, strands):
    """One PSL line per read: cols 0=matches, 5=gaps, 8=strand, 9=read, 13=splint."""
    with open(path, "w") as f:
        for n, s, st in zip(names, splint_names, strands):
            cols = ["0"] * 21
            cols[0], cols[5], cols[8], cols[9], cols[13] = "250", "0", st, n, s
            f.write("\t".join(cols) + "\n")


def _make_chunk(m, il, k, sp, seed, c0, err, flank, both_strands):
    rng = np.random.default_rng([seed, c0])
    ls = sp.size
    U = il + ls
    ins = BASES[rng.integers(0, 4, size=(m, il), dtype=np.uint8)]
    unit = np.concatenate([ins, np.broadcast_to(sp, (m, ls))], axis=1)
    full = np.tile(unit, (1, k + 2))
    head = rng.integers(flank[0], min(flank[1], il) + 1, size=m)
    tail = rng.integers(flank[0], min(flank[1], il) + 1, size=m)
    start = il - head
    stop = (k + 1) * U + tail
    cols = np.arange(full.shape[1], dtype=np.int32)[None, :]
    mask = (cols >= start[:, None]) & (cols < stop[:, None])
    clean = full[mask]                                   # row-major: reads stay contiguous
    clen = (stop - start).astype(np.int64)
    # errors over the whole chunk at once (16-bit uniform draws: resolution 1.5e-5)
    nb = clean.size
    sub, ins_p, dele = err
    u = rng.integers(0, 65536, size=nb, dtype=np.uint16)
    t_del, t_sub = int(round(dele * 65536)), int(round((dele + sub) * 65536))
    keep = u >= t_del
    is_sub = keep & (u < t_sub)
    out = clean
    idx_sub = np.flatnonzero(is_sub)
    if idx_sub.size:
        cur = (out[idx_sub] >> 1) & 3                      # A,C,G,T -> 0,1,3,2 (distinct)
        lut = np.array([0, 1, 3, 2], dtype=np.uint8)        # back to the index in BASES
        out[idx_sub] = BASES[(lut[cur] + rng.integers(1, 4, size=idx_sub.size, dtype=np.uint8)) % 4]
    is_ins = rng.integers(0, 65536, size=nb, dtype=np.uint16) < int(round(ins_p * 65536))
    reps = keep.astype(np.int32) + is_ins.astype(np.int32)
    noisy = np.repeat(out, reps)
    cs = np.cumsum(reps, dtype=np.int64)
    pos = cs[is_ins] - reps[is_ins]
    noisy[pos] = BASES[rng.integers(0, 4, size=pos.size, dtype=np.uint8)]
    bounds = np.concatenate(([0], np.cumsum(clen)))
    new_bounds = np.concatenate(([0], cs[bounds[1:] - 1]))
    st = (rng.random(m) < 0.5) if both_strands else np.zeros(m, dtype=bool)
    for i in np.flatnonzero(st):
        a, b = new_bounds[i], new_bounds[i + 1]
        noisy[a:b] = _COMP[noisy[a:b]][::-1]
    return noisy, np.diff(new_bounds), st


def make_batch(n_reads: int, insert_len=1000, repeats=5, seed=20251018, splint: str = SPLINT1,
               err=(0.04, 0.03, 0.03), flank=(100, 400), both_strands=True, workers=None):
    """Vectorised generator for uniform configs (same insert length and repeat count for every
    read): returns (blob uint8 ASCII, off int64[n+1], strand bool[n] (True = '-')).
    Same read model as make_reads; used where 1e5+ reads are needed quickly.  Deterministic in
    (seed, n_reads): every 4096-read chunk has its own generator."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    sp = np.frombuffer(splint.encode(), dtype=np.uint8)
    chunk = 4096
    jobs = [(min(chunk, n_reads - c0), int(insert_len), int(repeats), sp, seed, c0, err, flank, both_strands)
            for c0 in range(0, n_reads, chunk)]
    workers = workers or min(16, os.cpu_count() or 1)
    with ThreadPoolExecutor(max_workers=workers) as ex:
        parts = list(ex.map(lambda a: _make_chunk(*a), jobs))
    blob = np.concatenate([p[0] for p in parts])
    off = np.zeros(n_reads + 1, dtype=np.int64)
    off[1:] = np.cumsum(np.concatenate([p[1] for p in parts]))
    return blob, off, np.concatenate([p[2] for p in parts])


def make_mixed_batch(n_reads: int, inserts, repeat_range, splints=None, seed=20251018, err=(0.04, 0.03, 0.03),
                     flank=(100, 400), both_strands=True, workers=None):
    """Vectorised generator for mixed configs: every read draws its insert length from `inserts` (a list of lengths),
    its repeat count from repeat_range (inclusive) and its splint from `splints` (list of str; default Splint1), all
    uniformly.  Reads of one (insert, repeats, splint) class are generated together (_make_chunk) and then put back at
    their drawn places, so the batch is mixed read by read.  Returns (blob uint8 ASCII, off int64[n+1],
    sp_idx int32[n] = 2 * splint + (strand == '-'), splint list in both orientations)."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    splints = list(splints) if splints else [SPLINT1]
    sp_arr = [np.frombuffer(s.encode(), dtype=np.uint8) for s in splints]
    rng = np.random.default_rng([seed, 7])
    inserts = [int(x) for x in inserts]
    ci = rng.integers(0, len(inserts), size=n_reads)
    ck = rng.integers(repeat_range[0], repeat_range[1] + 1, size=n_reads)
    cs = rng.integers(0, len(splints), size=n_reads)
    key = (ci * 64 + ck) * 16 + cs
    order = np.argsort(key, kind="stable")
    ks, starts = np.unique(key[order], return_index=True)
    ends = list(starts[1:]) + [n_reads]
    jobs, where = [], []
    for kk, a, b in zip(ks, starts, ends):
        il, k, sp = inserts[int(kk) // 16 // 64], int(kk) // 16 % 64, int(kk) % 16
        m_max = max(64, min(4096, (192 << 20) // ((k + 2) * (il + sp_arr[sp].size))))      # <= ~192 MB of clean bases per chunk
        for c0 in range(int(a), int(b), m_max):
            c1 = min(int(b), c0 + m_max)
            jobs.append((c1 - c0, il, k, sp_arr[sp], seed, c0, err, flank, both_strands))
            where.append((order[c0:c1], sp))
    workers = workers or min(16, os.cpu_count() or 1)
    with ThreadPoolExecutor(max_workers=workers) as ex:
        parts = list(ex.map(lambda a: _make_chunk(*a), jobs))
    lens = np.zeros(n_reads, dtype=np.int64)
    sp_idx = np.zeros(n_reads, dtype=np.int32)
    for (idx, sp), (_, ln, st) in zip(where, parts):
        lens[idx] = ln
        sp_idx[idx] = 2 * sp + st.astype(np.int32)
    off = np.zeros(n_reads + 1, dtype=np.int64)
    off[1:] = np.cumsum(lens)
    blob = np.empty(int(off[-1]), dtype=np.uint8)
    for (idx, _), (nz, ln, _) in zip(where, parts):
        src = np.concatenate(([0], np.cumsum(ln)))
        for t, i in enumerate(idx):
            blob[off[i]:off[i + 1]] = nz[src[t]:src[t + 1]]
    both = []
    for s_ in splints:
        both += [s_, revcomp(s_)]
    return blob, off, sp_idx, both
